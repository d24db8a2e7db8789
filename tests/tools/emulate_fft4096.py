"""NumPy emulation of the index algebra of gomel_b200/csrc/fft4096.cuh (thread/slot layouts,
exchange patterns, twiddle tables, partner lanes, magnitude row order).  CPU-only sanity check of
the derivation; it mirrors the CUDA code line by line but is not the product."""
import numpy as np

N = 4096
T = np.arange(256)


def lanes():
    w, l = T >> 5, T & 31
    k0c = np.where(l < 16, w, np.where(w == 0, 8, 16 - w))
    k1c = np.where(l < 16, l, 31 - l)
    src = np.where(w == 0, np.where(l < 16, (16 - l) & 15, 47 - l), l ^ 16)
    return dict(base_a=T + 2 * (T >> 4), base_b=k0c * 288 + (l & 15), n0b=l & 15, base_c=k0c * 288 + k1c * 18,
                k0c=k0c, k1c=k1c, klow=k0c + 16 * k1c, src=src)


def dft16(v, inv):            # v: (256, 16) natural in/out
    k = np.arange(16)
    W = np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 16)
    return v @ W.T


def tables():
    t = T
    T1 = np.array([np.exp(-2j * np.pi * ((t * k0) % 4096) / 4096) for k0 in range(1, 16)])   # [15][256]
    T2 = np.array([[np.exp(-2j * np.pi * ((n0 * k1) % 256) / 256) for n0 in range(16)] for k1 in range(16)])
    return T1, T2


def fwd(v, L, T1, T2):
    xb = np.zeros(16 * 288, complex)
    v = dft16(v, False)
    for k0 in range(1, 16):
        v[:, k0] *= T1[k0 - 1][T]
    for k0 in range(16):
        xb[k0 * 288 + L['base_a']] = v[:, k0]
    v = np.stack([xb[L['base_b'] + r * 18] for r in range(16)], axis=1)
    v = dft16(v, False)
    for k1 in range(1, 16):
        v[:, k1] *= T2[k1][L['n0b']]
    for r in range(16):
        xb[L['base_b'] + r * 18] = v[:, r]
    v = np.stack([xb[L['base_c'] + c] for c in range(16)], axis=1)
    return dft16(v, False)


def inv(v, L, T1, T2):
    xb = np.zeros(16 * 288, complex)
    v = dft16(v, True)
    for n0 in range(1, 16):
        v[:, n0] *= np.conj(T2[n0][L['k1c']])
    for c in range(16):
        xb[L['base_c'] + c] = v[:, c]
    v = np.stack([xb[L['base_b'] + r * 18] for r in range(16)], axis=1)
    v = dft16(v, True)
    for r in range(16):
        xb[L['base_b'] + r * 18] = v[:, r]
    v = np.stack([xb[k0 * 288 + L['base_a']] for k0 in range(16)], axis=1)
    for k0 in range(1, 16):
        v[:, k0] *= np.conj(T1[k0 - 1][T])
    return dft16(v, True)


def partner(v, L):
    """P[t][k2] = Z[N-k] as fetch_partner builds it (incl. the special thread's `saved` chain)."""
    P = np.zeros_like(v)
    saved = None
    for j in range(8):
        gsrc = (T & ~31) | L['src']          # __shfl_sync source lane is relative to the warp
        a = v[gsrc, 15 - j].copy()
        b = v[gsrc, j].copy()
        # special thread t = 0
        a[0] = v[0, 0] if j == 0 else saved
        b[0] = v[0, j + 1]
        saved = v[0, 15 - j]
        P[:, j], P[:, 15 - j] = a, b
    return P


def mag_pos(k):
    if k >= 2048:
        return 2048
    k0, k1 = k & 15, (k >> 4) & 15
    rho = ((k0 & 7) << 1) | (k0 >> 3)
    return (k & ~255) | (rho << 4) | k1


def mag_unpos(pos):
    rho, k1 = (pos >> 4) & 15, pos & 15
    k0 = (rho >> 1) | ((rho & 1) << 3)
    return (pos & ~255) | (k1 << 4) | k0


def main():
    rng = np.random.default_rng(0)
    L = lanes()
    T1, T2 = tables()
    z = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    v = np.stack([z[T + 256 * m] for m in range(16)], axis=1)
    Z = fwd(v.copy(), L, T1, T2)
    ref = np.fft.fft(z)
    k = L['klow'][:, None] + 256 * np.arange(16)[None, :]
    print("fwd err", np.abs(Z - ref[k]).max())
    assert np.abs(Z - ref[k]).max() < 1e-9
    assert sorted(L['klow'].tolist()) == list(range(256))
    P = partner(Z, L)
    print("partner err", np.abs(P - ref[(N - k) % N]).max())
    assert np.abs(P - ref[(N - k) % N]).max() < 1e-9
    back = inv(Z.copy(), L, T1, T2) / N
    print("inv err", np.abs(back - v).max())
    assert np.abs(back - v).max() < 1e-12
    # magnitude row indices used by k_gl_iter
    klow = L['klow']
    idx_lo = np.array([mag_pos(int(k)) for k in klow])
    idx_hi = np.array([mag_pos(256 - int(k)) if k else 256 for k in klow])
    assert all(mag_unpos(mag_pos(k)) == k for k in range(2048))
    # staged rows are read conflict-free: 32 lanes of a warp hit 32 distinct banks (one duplicate allowed in warp 0)
    for w in range(8):
        for idx in (idx_lo, idx_hi):
            banks = idx[w * 32:(w + 1) * 32] % 32
            assert len(set(banks.tolist())) >= 31, (w, sorted(banks.tolist()))
    for t in range(256):
        for j in range(8):
            kk = klow[t] + 256 * j
            assert j * 256 + idx_lo[t] == mag_pos(kk), (t, j)
            kk2 = klow[t] + 256 * (15 - j)
            assert j * 256 + idx_hi[t] == mag_pos(N - kk2), (t, j, kk2)
    # two real frames through one transform + Griffin-Lim substitution identity
    xa, xb_ = rng.standard_normal(N), rng.standard_normal(N)
    zz = xa + 1j * xb_
    v = np.stack([zz[T + 256 * m] for m in range(16)], axis=1)
    Z = fwd(v, L, T1, T2)
    P = partner(Z, L)
    XA2 = Z + np.conj(P)
    XB2 = (Z - np.conj(P)) / 1j
    print("split err", np.abs(XA2 / 2 - np.fft.fft(xa)[k]).max(), np.abs(XB2 / 2 - np.fft.fft(xb_)[k]).max())
    MA, MB = rng.random(2049), rng.random(2049)
    kb = np.minimum(k, N - k)
    YA = MA[kb] * XA2 / np.abs(XA2)
    YB = MB[kb] * XB2 / np.abs(XB2)
    out = inv(YA + 1j * YB, L, T1, T2) / N
    ya = np.fft.irfft(MA * np.fft.rfft(xa) / np.abs(np.fft.rfft(xa)), n=N)
    yb = np.fft.irfft(MB * np.fft.rfft(xb_) / np.abs(np.fft.rfft(xb_)), n=N)
    n = T[:, None] + 256 * np.arange(16)[None, :]
    print("GL pair err", np.abs(out.real - ya[n]).max(), np.abs(out.imag - yb[n]).max())
    assert np.abs(out.real - ya[n]).max() < 1e-9 and np.abs(out.imag - yb[n]).max() < 1e-9
    # bank conflicts: each half-warp of a 64-bit access must hit 16 distinct bank pairs
    for name, base, offs in (("a", L['base_a'], [k0 * 288 for k0 in range(16)]),
                             ("b", L['base_b'], [r * 18 for r in range(16)])):
        worst = 1
        for o in offs:
            addr = base + o
            for h in range(16):
                banks = (addr[h * 16:(h + 1) * 16] * 2) % 32
                worst = max(worst, 16 // len(set(banks.tolist())))
        print("pattern", name, "worst conflict degree", worst)
        assert worst == 1
    # pattern (c) uses 128-bit accesses: a quarter-warp (8 lanes x 16 B) must cover 32 distinct banks
    worst = 1
    for q in range(8):
        addr = (L['base_c'] + 2 * q) * 2                       # first 32-bit word of each lane's float4
        assert np.all(addr % 4 == 0)                           # 16-byte aligned
        for g8 in range(32):
            words = np.concatenate([(addr[g8 * 8:(g8 + 1) * 8] + i) % 32 for i in range(4)])
            worst = max(worst, 32 // len(set(words.tolist())))
    print("pattern c (128-bit) worst conflict degree", worst)
    assert worst == 1
    # exchange 2 is warp-local: the cells a thread reads in pattern (c) were written (pattern b) by threads of its own warp
    owner = {}
    for t in range(256):
        for r in range(16):
            owner[int(L['base_b'][t]) + r * 18] = t >> 5
    for t in range(256):
        for c in range(16):
            assert owner[int(L['base_c'][t]) + c] == t >> 5
    print("exchange 2 warp-local: OK")
    print("OK")


if __name__ == "__main__":
    main()
