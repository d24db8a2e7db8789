"""Second, independent float64 restatement of the gomel hot path in NumPy (pocketfft rfft/irfft,
half-spectrum Hermitian form).  TEST INFRASTRUCTURE ONLY -- it exists to cross-check
oracle/gomel_oracle.c (the literal, full-spectrum restatement) and is never imported by the
product.  Reference lines as in gomel_oracle.h; the algebraic short-cuts used here are the ones
SURVEY.md Appendix B verified (full-spectrum Abs/Phase/Rect loop == 2049-bin C2R formulation).
"""
import numpy as np


def pad(buf, hop):
    """mel/impl.go:429-455"""
    n, mt = len(buf), 15 * hop
    if n >= mt:
        r = (n - mt) % hop
        p = hop - r - 1 if r else 0
    else:
        p = mt - n - 1
    return np.concatenate([buf, np.zeros(p)]) if p > 0 else np.asarray(buf, np.float64)


def frames_of(x, N, H):
    nf = int((len(x) - N) / H) + 1
    idx = np.arange(N)[None, :] + np.arange(nf)[:, None] * H
    return x[idx]


def stft_half(x, N, H):
    return np.fft.rfft(frames_of(x, N, H) * np.hanning(N), axis=1)      # (frames, N/2+1)


def hz_to_mel(v):
    return 1127.0 * np.log(1.0 + v / 700.0)


def mel_to_hz(v):
    return 700.0 * (np.exp(v / 1127.0) - 1.0)


def to_mel(wav, mels=192, fmin=0.0, fmax=16000.0, hop=1280, N=4096):
    """mel/mel.go:46-74"""
    B = N // 2
    X = np.abs(stft_half(pad(np.asarray(wav, np.float64), hop), N, H=hop))
    ch = np.stack([X[:, :B], X[:, 1:B + 1]], axis=2)                    # |X[j]|, |X[N-1-j]|=|X[j+1]|
    melbin = hz_to_mel(fmax) / mels
    out = np.zeros((X.shape[0], mels, 2))
    for i in range(mels):
        vallo = B * (fmin + mel_to_hz(melbin * i)) / (fmax + fmin)
        valhi = B * (fmin + mel_to_hz(melbin * (i + 1))) / (fmax + fmin)
        inlo, modlo, inhi = int(np.trunc(vallo)), vallo - np.trunc(vallo), int(np.floor(valhi))
        if vallo < 0 and inlo < 0:
            inlo, modlo, inhi = 0, 0.0, 0
        if inlo + 1 == inhi:
            out[:, i, :] = ch[:, inlo, :] * (1 - modlo) + ch[:, inhi, :] * modlo
        else:
            tot = np.zeros((X.shape[0], 2))
            for k in range(inlo, inhi):
                tot += ch[:, k, :]
            out[:, i, :] = tot / (inhi - inlo + 1)
    return np.log(np.maximum(out, 1e-5)).reshape(-1, 2)


def undomel(lin_mel, mels, B, fmin, fmax):
    """mel/impl.go:347-384 on (frames, mels, 2) linear values"""
    filterbin = hz_to_mel(fmax) / mels
    F = lin_mel.shape[0]
    out = np.zeros((F, B, 2))
    for i in range(B):
        vallo = hz_to_mel(i * (fmax + fmin) / B - fmin) / filterbin
        valhi = hz_to_mel((i + 1) * (fmax + fmin) / B - fmin) / filterbin
        inlo, modlo, inhi = int(np.trunc(vallo)), vallo - np.trunc(vallo), int(np.floor(valhi))
        if inlo == inhi:
            out[:, i, :] = lin_mel[:, inlo, :]
        elif inlo + 1 == inhi and inhi < mels:
            out[:, i, :] = lin_mel[:, inlo, :] * (1 - modlo) + lin_mel[:, inhi, :] * modlo
        else:
            tot = np.zeros((F, 2))
            for k in range(inlo, inhi):
                tot += lin_mel[:, k, :]
            out[:, i, :] = tot / (inhi - inlo + 1)
    return out


def gl_magnitudes(mel, mels=192, fmin=0.0, fmax=16000.0, N=4096, tune_mul=1.0, tune_add=0.0):
    """exp -> undomel -> undospectrum, reduced to the 2049 magnitudes Griffin-Lim actually uses
    (M[k]=ch0[k] k<2048, M[2048]=ch1[2047])."""
    B = N // 2
    lin = np.exp(np.asarray(mel, np.float64).reshape(-1, mels, 2))
    sp = (undomel(lin, mels, B, fmin, fmax) - tune_add) / tune_mul
    return np.concatenate([sp[:, :, 0], sp[:, B - 1:B, 1]], axis=1)     # (frames, 2049)


def griffin_lim(M, init, iters, hop=1280, N=4096):
    """mel/mel.go:76-139 in Hermitian form: no window-sum normalisation, Jacobi update."""
    F = M.shape[0]
    w = np.hanning(N)
    ola = N + (F - 1) * hop
    sig = np.array(init, np.float64)
    idx = np.arange(N)[None, :] + np.arange(F)[:, None] * hop
    Mabs = np.abs(M)                                                    # cmplx.Abs of Rect(real,0)
    for _ in range(iters):
        X = np.fft.rfft(sig[idx] * w, axis=1)
        mag = np.abs(X)
        unit = np.where(mag > 0, X / np.where(mag > 0, mag, 1), 1.0)
        Y = Mabs * unit
        y = np.fft.irfft(Y, n=N, axis=1) * w
        new = np.zeros(ola)
        for f in range(F):
            new[f * hop:f * hop + N] += y[f]
        sig = new
    return sig


def from_mel(mel, init, iters, mels=192, fmin=0.0, fmax=16000.0, hop=1280, N=4096):
    return griffin_lim(gl_magnitudes(mel, mels, fmin, fmax, N), init, iters, hop, N)


def to_phase(wav, num_freqs=768, hop=1280, N=4096):
    """phase/phase.go:41-70"""
    X = stft_half(pad(np.asarray(wav, np.float64), hop), N, hop)
    keep = X[:, 1:num_freqs + 1]
    return np.stack([keep.imag, keep.real], axis=2).reshape(-1, 2)


def from_phase(spec, num_freqs=768, hop=1280, N=4096, volume_boost=0.0):
    """phase/phase.go:136-153"""
    B = N // 2
    s = np.asarray(spec, np.float64).reshape(-1, num_freqs, 2)
    F = s.shape[0]
    X = np.zeros((F, B + 1), np.complex128)
    X[:, 1:num_freqs + 1] = s[:, :, 1] + 1j * s[:, :, 0]
    X[:, num_freqs + 1:] = X[:, num_freqs:num_freqs + 1]                # grow: replicate last kept bin
    X[:, B] = X[:, B].real                                              # conj write wins; C2R uses Re only
    w = np.hanning(N)
    y = np.fft.irfft(X, n=N, axis=1) * w
    ola = N + (F - 1) * hop
    out, ws = np.zeros(ola), np.zeros(ola)
    for f in range(F):
        out[f * hop:f * hop + N] += y[f]
        ws[f * hop:f * hop + N] += w * w
    thr = ws.max() * 0.5
    hi = ws > thr
    mid = (~hi) & (ws > 1e-21)
    out[hi] /= ws[hi]
    out[mid] = out[mid] / ws[mid] * (ws[mid] / thr)
    if volume_boost != 0:
        out = out * volume_boost
    return out
