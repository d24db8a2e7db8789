"""Property tests (Hypothesis) of the host-side logic around the hot path -- the tests the reference's own port
plan lists but never wrote (.kiro/specs/phase-python-port/tasks.md: pad/is_padded, shrink/grow, float16
metadata, zero-stuffing).  CPU only: the product's host helpers against the C oracle and against the stated
invariants; plus the multi-GPU bookkeeping (clip sharding, time-split partition, tiling of the C ABI)."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from gomel_b200 import _lib, codec, phase, shard, timesplit

HOPS = st.sampled_from([256, 1280, 512, 100])


@settings(max_examples=300, deadline=None)
@given(n=st.integers(1, 3_000_000), hop=HOPS)
def test_pad_length_rule(oracle, n, hop):
    """pad (mel/impl.go:429-455): below 15 hops pad to 15*hop - 1; above, to length = 15*hop + k*hop + hop-1
    unless the remainder is already 0; is_padded (:457-479) recognises exactly that length"""
    p = codec.pad_len(n, hop)
    assert p == oracle.pad_len(n, hop) and 0 <= p < max(hop, 15 * hop)
    total = n + p
    if n < 15 * hop:
        assert total == max(15 * hop - 1, n)
    else:
        r = (n - 15 * hop) % hop
        assert (p == 0) if r == 0 else ((total - 15 * hop) % hop == hop - 1)
    assert codec.is_padded(n, total, hop) and oracle.is_padded(n, total, hop)
    if p > 0:
        assert not codec.is_padded(n, total + 1, hop)
    assert len(codec.pad(np.zeros(min(n, 5000)), hop)) == min(n, 5000) + codec.pad_len(min(n, 5000), hop)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 3_000_000), geo=st.sampled_from([(4096, 1280), (2048, 256)]))
def test_frame_count_rule_of_the_c_abi(n, geo):
    """gomel_frames = pad + gossp NumFrames int((len - N)/hop) + 1 + ISTFT length N + (F-1)*hop; no GPU needed"""
    n_fft, hop = geo
    cfg = _lib.make_config(n_fft=n_fft, hop=hop)
    npad, fr, ola = _lib.frames(cfg, n)
    assert npad == n + codec.pad_len(n, hop)
    assert fr == (npad - n_fft) // hop + 1 and fr >= 1
    assert ola == n_fft + (fr - 1) * hop and ola <= npad


@settings(max_examples=100, deadline=None)
@given(frames=st.integers(1, 6), nf=st.integers(1, 2048), seed=st.integers(0, 2**31))
def test_shrink_grow(frames, nf, seed):
    """shrink keeps the first NumFreqs bins of every frame; grow replicates the LAST kept bin upward (not zeros)"""
    spec = np.random.default_rng(seed).standard_normal((frames * 2048, 2))
    small = phase.shrink(spec, 4096, nf)
    assert small.shape == (frames * nf, 2)
    assert np.array_equal(small.reshape(frames, nf, 2), spec.reshape(frames, 2048, 2)[:, :nf])
    big = phase.grow(small, 4096, nf).reshape(frames, 2048, 2)
    assert np.array_equal(big[:, :nf], small.reshape(frames, nf, 2))
    assert np.array_equal(big[:, nf:], np.repeat(small.reshape(frames, nf, 2)[:, -1:], 2048 - nf, axis=1))
    assert np.array_equal(phase.shrink(big.reshape(-1, 2), 4096, nf), small)


@settings(max_examples=300, deadline=None)
@given(v=st.floats(-65000.0, 65000.0, allow_nan=False))
def test_float16_metadata_roundtrip(oracle, v):
    """mel/impl.go:120-125 (float64 -> float32 -> float16) vs phase.py:604-620 (float64 -> float16): both within half
    an ulp of float16; the Go order equals the oracle's bit pattern"""
    go, py = codec.pack_f16_go(v), codec.pack_f16_py(v)
    assert int.from_bytes(go, "little") == oracle.f16_bits(v)
    for b in (go, py):
        back = codec.unpack_f16(b)
        assert abs(back - v) <= max(abs(v) * 2.0**-11, 2.0**-25)
    assert codec.unpack_f16(codec.pack_f16_go(codec.unpack_f16(go))) == codec.unpack_f16(go)      # idempotent


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 4000), zero_pad=st.integers(1, 50), zero_shift=st.integers(0, 7), seed=st.integers(0, 2**31))
def test_zero_stuff_upsample(n, zero_pad, zero_shift, seed):
    """phase/impl.go:509-529: after every `zero_pad` samples insert `zero_shift` zeros; kept samples x (1+shift)"""
    x = np.random.default_rng(seed).standard_normal(n)
    y = phase.zero_stuff_upsample(x, zero_pad, zero_shift)
    groups = -(-n // zero_pad)
    assert len(y) == n + groups * zero_shift
    period = zero_pad + zero_shift
    idx = np.arange(len(y))
    kept = (idx % period) < zero_pad
    assert np.array_equal(y[kept][:n], x * (1 + zero_shift)) and not y[~kept].any()
    assert phase.zero_stuff_upsample(x, 0, zero_shift) is x


@settings(max_examples=60, deadline=None)
@given(mels=st.integers(8, 256), fmax=st.sampled_from([8000.0, 11025.0, 16000.0, 22050.0]),
       bins=st.sampled_from([1024, 2048]))
def test_filterbank_tables_tile_the_axis(oracle, mels, fmax, bins):
    """domel / undomel (mel/impl.go:310-384): band i ends where band i+1 starts, nothing lies outside the
    spectrum / the mel axis, and the product-side mirror equals the oracle's tables bit for bit"""
    flo, fhi, fmod, ilo, ihi, imod = _lib.mel_tables(bins, mels, 0.0, fmax)
    lo, hi, mod = oracle.mel_fwd_tables(bins, mels, 0.0, fmax)
    assert np.array_equal(flo, lo) and np.array_equal(fhi, hi) and np.array_equal(fmod, mod)
    lo, hi, mod = oracle.mel_inv_tables(bins, mels, 0.0, fmax)
    assert np.array_equal(ilo, lo) and np.array_equal(ihi, hi) and np.array_equal(imod, mod)
    assert flo[0] == 0 and np.all(flo[1:] == fhi[:-1]) and np.all(fhi >= flo) and fhi[-1] <= bins
    assert np.all((fmod >= 0) & (fmod < 1)) and np.all((imod >= 0) & (imod < 1))
    assert ilo[0] == 0 and np.all(ilo[1:] == ihi[:-1]) and np.all(ihi >= ilo) and ihi[-1] <= mels


@given(n=st.integers(0, 5000), world=st.integers(1, 16))
def test_clip_sharding_covers_every_clip_once(n, world):
    parts = [shard.clip_range(n, r, world) for r in range(world)]
    assert [i for p in parts for i in p] == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=300, deadline=None)
@given(frames=st.integers(8, 200_000), world=st.integers(1, 8), tile=st.integers(0, 128))
def test_time_split_partition(frames, world, tile):
    """rank ranges are contiguous, start on tile boundaries and cover every frame once"""
    T = max((tile if tile > 0 else 16) + ((tile if tile > 0 else 16) & 1), 4)
    if -(-frames // T) < world:
        return
    parts = timesplit.partition(frames, world, tile)
    pos = 0
    for a, n in parts:
        assert a == pos and a % T == 0 and n > 0
        pos += n
    assert pos == frames


# ---- properties of the oracle itself (the mel path has no reference golden vectors: DESIGN.md section 2) ----------
@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31), gain=st.floats(0.25, 8.0), geo=st.sampled_from([(4096, 1280, 192, 16000.0), (2048, 256, 160, 8000.0)]))
def test_oracle_to_mel_is_homogeneous_and_shift_covariant(oracle, seed, gain, geo):
    """|X| is linear in the signal and domel is a linear band average, so above the 1e-5 floor
    exp(ToMel(g*x)) = g * exp(ToMel(x)); and delaying x by one hop moves every frame by one column"""
    n_fft, hop, mels, fmax = geo
    cfg = oracle.config(num_mels=mels, window=hop, resolut=n_fft, mel_fmax=fmax)
    x = np.random.default_rng(seed).standard_normal(20 * hop) * 0.3
    a = np.exp(oracle.to_mel(cfg, x))
    b = np.exp(oracle.to_mel(cfg, gain * x))
    ok = (a > 2e-5) & (b > 2e-5)
    assert ok.mean() > 0.9 and np.allclose(b[ok], gain * a[ok], rtol=1e-12)
    assert a.min() >= 1e-5 * (1 - 1e-15)                                  # spectral_normalize floor
    xs = np.concatenate([np.zeros(hop), x])                               # one hop later
    c = np.exp(oracle.to_mel(cfg, xs)).reshape(-1, mels, 2)
    a3 = a.reshape(-1, mels, 2)
    k = (len(x) - n_fft) // hop + 1                                       # frames made of signal samples only
    assert np.allclose(c[1:k + 1], a3[:k], rtol=1e-12, atol=1e-15)


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 2**31), frames=st.integers(1, 6))
def test_oracle_griffin_lim_structure(oracle, seed, frames):
    """mel.ISTFT (mel/mel.go:76-139): zero iterations return the start signal; the result does not depend on the
    start signal's scale (only its phases are kept)"""
    rng = np.random.default_rng(seed)
    mel = rng.uniform(-6.0, 1.0, (frames * 192, 2))
    init = rng.random(4096 + (frames - 1) * 1280) + 0.05
    assert np.array_equal(oracle.from_mel(oracle.config(gl_iters=0), mel, init), init)
    one = oracle.from_mel(oracle.config(gl_iters=1), mel, init)
    scaled = oracle.from_mel(oracle.config(gl_iters=1), mel, 3.0 * init)
    assert np.allclose(one, scaled, rtol=1e-9, atol=1e-12)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31), nf=st.sampled_from([384, 768, 836, 2048]), boost=st.sampled_from([0.0, 0.5, 2.0]))
def test_oracle_phase_roundtrip_recovers_the_interior(oracle, seed, nf, boost):
    """FromPhase(ToPhase(x)) (phase/phase.go:41-153): with all 2048 bins kept and the DC bin (dropped by the
    format) absent from x, the interior (window-sum above the 0.5*max threshold) is x itself, times VolumeBoost"""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(30000) * 0.2
    x -= x.mean()
    cfg = oracle.config(num_freqs=nf, volume_boost=boost)
    spec = oracle.to_phase(cfg, x)
    y = oracle.from_phase(cfg, spec)
    assert len(spec) % nf == 0 and len(y) == 4096 + (len(spec) // nf - 1) * 1280
    if nf == 2048:
        g = boost if boost != 0 else 1.0
        sl = slice(4096, 26000)
        # per-frame DC (mean of the windowed frame) is dropped with bin 0: small, not zero
        assert np.abs(y[sl] - g * x[sl]).max() < 0.05 * g
