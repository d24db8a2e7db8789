"""CPU tests of the oracle: golden vectors produced by the reference's own phase.py, the KATs of
the reference's test scripts, cross-check of the C restatement against the independent NumPy one,
and known-answer signals.  No GPU."""
import os

import numpy as np
import pytest

from util import rel_l2, synth_clip

GOLD = os.path.join(os.path.dirname(__file__), "golden", "phase_ref.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


# ---- pinned against the reference's phase.py (tests/golden/make_golden.py) ----------------------
@pytest.mark.parametrize("name,sr", [("a48k", 48000), ("b44k", 44100), ("c48k_short", 48000), ("hdr", 48000)])
def test_phase_golden(oracle, gold, name, sr):
    nf = int(gold[f"{name}_num_freqs"])
    cfg = oracle.config(num_freqs=nf)
    spec = oracle.to_phase(cfg, gold[f"{name}_wav"])
    assert spec.shape == gold[f"{name}_spec"].shape
    assert rel_l2(spec, gold[f"{name}_spec"]) < 1e-13
    rec = oracle.from_phase(cfg, gold[f"{name}_spec"])
    assert rec.shape == gold[f"{name}_rec"].shape
    assert rel_l2(rec, gold[f"{name}_rec"]) < 1e-13


def test_phase_golden_volume_boost(oracle, gold):
    cfg = oracle.config(num_freqs=768, volume_boost=1.666)
    assert rel_l2(oracle.from_phase(cfg, gold["a48k_spec"]), gold["a48k_rec_boost"]) < 1e-13


def test_pad_golden(oracle, gold):
    for n, padded, yes, no in zip(gold["pad_in"], gold["pad_out"], gold["is_padded"], gold["is_padded_neg"]):
        assert int(n) + oracle.pad_len(int(n), 1280) == int(padded)
        assert oracle.is_padded(int(n), int(padded), 1280) == bool(yes)
        assert oracle.is_padded(int(n), int(padded) + 1, 1280) == bool(no)


def test_zero_stuff_kats(oracle, gold):
    """test_zero_stuff.py:9-34 (the boost (1+zero_shift) is applied by the code, not by the printed text)"""
    for i in range(4):
        zp, zs = (int(x) for x in gold[f"zs{i}_args"])
        assert np.array_equal(oracle.zero_stuff(gold[f"zs{i}_in"], zp, zs), gold[f"zs{i}_out"])
    assert np.array_equal(oracle.zero_stuff(np.array([1., 2., 3., 4., 5.]), 1, 1),
                          np.array([2., 0, 4., 0, 6., 0, 8., 0, 10., 0]))
    assert oracle.pad_shift(22050) == (1, 1) and oracle.pad_shift(8000) == (1, 5) and oracle.pad_shift(44100) == (0, 0)
    assert oracle.pad_shift(32000) == (2, 1) and oracle.pad_shift(16000) == (1, 2) and oracle.pad_shift(11025) == (1, 3)


def test_float16_golden(oracle, gold):
    b = gold["f16_bytes"]
    for i, v in enumerate(gold["f16_in"]):
        bits = oracle.f16_bits(float(v))
        assert (bits & 0xFF, bits >> 8) == (int(b[2 * i]), int(b[2 * i + 1])), v
        assert oracle.f16_value(bits) == float(gold["f16_back"][i])
    assert oracle.f16_value(oracle.f16_bits(44100.0)) == 44096.0


# ---- C restatement vs the independent NumPy restatement ------------------------------------------
def test_c_vs_numpy_mel(oracle):
    from oracle import oracle_np as ONP
    wav = synth_clip(0, 1.2)
    cfg = oracle.config(gl_iters=3)
    mel = oracle.to_mel(cfg, wav)
    assert rel_l2(mel, ONP.to_mel(wav)) < 1e-12
    frames = len(mel) // 192
    init = np.random.default_rng(1).random(4096 + (frames - 1) * 1280)
    a = oracle.from_mel(cfg, mel, init)
    b = ONP.from_mel(mel, init, 3)
    assert rel_l2(a, b) < 1e-12                                          # full-spectrum loop == Hermitian form


def test_both_restatements_agree_on_the_newmel_default_geometry(oracle):
    """mel.NewMel() defaults (mel/mel.go:30-41): 160 mels, fmax 8000, Window 256, Resolut 2048, 2 iterations"""
    from oracle import oracle_np as ONP
    wav = synth_clip(3, 0.35)
    cfg = oracle.config(num_mels=160, window=256, resolut=2048, mel_fmax=8000.0, gl_iters=2)
    a = oracle.to_mel(cfg, wav)
    b = ONP.to_mel(wav, mels=160, fmax=8000.0, hop=256, N=2048)
    assert a.shape == b.shape and rel_l2(a, b) < 1e-12
    frames = len(a) // 160
    init = np.random.default_rng(6).random(2048 + (frames - 1) * 256)
    x = oracle.from_mel(cfg, a, init)
    y = ONP.from_mel(a, init, 2, mels=160, fmax=8000.0, hop=256, N=2048)
    assert x.shape == y.shape and rel_l2(x, y) < 1e-12
    # the filterbank tables of the product-side mirror at this geometry == the oracle's
    from gomel_b200 import _lib
    flo, fhi, fmod, ilo, ihi, imod = _lib.mel_tables(1024, 160, 0.0, 8000.0)
    lo, hi, mod = oracle.mel_fwd_tables(1024, 160, 0.0, 8000.0)
    assert np.array_equal(flo, lo) and np.array_equal(fhi, hi) and np.array_equal(fmod, mod)
    lo, hi, mod = oracle.mel_inv_tables(1024, 160, 0.0, 8000.0)
    assert np.array_equal(ilo, lo) and np.array_equal(ihi, hi) and np.array_equal(imod, mod)


def test_c_vs_numpy_phase(oracle):
    from oracle import oracle_np as ONP
    wav = synth_clip(1, 0.9)
    cfg = oracle.config(num_freqs=836, volume_boost=0.5)
    spec = oracle.to_phase(cfg, wav)
    assert rel_l2(spec, ONP.to_phase(wav, 836)) < 1e-13
    assert rel_l2(oracle.from_phase(cfg, spec), ONP.from_phase(spec, 836, volume_boost=0.5)) < 1e-13


# ---- definitions ---------------------------------------------------------------------------------
def test_fft_definition(oracle):
    x = np.random.default_rng(0).standard_normal(4096) + 1j * np.random.default_rng(1).standard_normal(4096)
    assert rel_l2(oracle.fft(x), np.fft.fft(x)) < 1e-13                  # forward un-normalised
    assert rel_l2(oracle.fft(x, inverse=True), np.fft.ifft(x)) < 1e-13   # inverse 1/N


def test_hann_is_symmetric_np_hanning(oracle):
    w = oracle.hann(4096)
    assert np.allclose(w, np.hanning(4096), atol=1e-15) and w[0] == 0 and abs(w[-1]) < 1e-15


def test_frame_geometry(oracle):
    cfg = oracle.config()
    assert oracle.num_frames(44100, cfg) == 32
    assert oracle.num_frames(441000, cfg) == 342
    assert oracle.num_frames(158760000, cfg) == 124029
    assert 441000 + oracle.pad_len(441000, 1280) == 441599


def test_filterbank_structure(oracle):
    """5 lerp + 187 box bands forward; 1856 copy + 191 lerp + 1 box inverse."""
    lo, hi, mod = oracle.mel_fwd_tables(2048, 192, 0.0, 16000.0)
    lerp = [i for i in range(192) if lo[i] + 1 == hi[i]]
    assert lerp == [0, 2, 4, 6, 10] and hi[-1] == 2048 and (hi - lo).max() == 36
    lo, hi, mod = oracle.mel_inv_tables(2048, 192, 0.0, 16000.0)
    copy = lo == hi
    lp = (lo + 1 == hi) & (hi < 192)
    assert copy.sum() == 1856 and lp.sum() == 191 and (~(copy | lp)).sum() == 1
    assert (lo[2047], hi[2047]) == (191, 192)                            # bin 2047 -> m[191]/2


def test_known_answers(oracle):
    cfg = oracle.config()
    # silence: every mel entry is ln(1e-5)
    m = oracle.to_mel(cfg, np.zeros(30000))
    assert np.all(m == np.log(1e-5))
    # DC: |X[0]| = sum(w) ; lowest mel band (lerp of bins 0,1) is dominated by it
    dc = oracle.to_mel(cfg, np.ones(40000))
    w = oracle.hann(4096)
    X = np.abs(np.fft.rfft(w))
    lo, hi, mod = oracle.mel_fwd_tables(2048, 192, 0.0, 16000.0)
    exp0 = X[lo[0]] * (1 - mod[0]) + X[hi[0]] * mod[0]
    assert abs(dc[0, 0] - np.log(exp0)) < 1e-12
    # bin-centred tone: phase representation peaks at that bin
    k = 200
    t = np.arange(40000)
    tone = np.cos(2 * np.pi * k * t / 4096)
    spec = oracle.to_phase(oracle.config(num_freqs=768), tone).reshape(-1, 768, 2)
    mag = np.hypot(spec[3, :, 0], spec[3, :, 1])
    assert mag.argmax() == k - 1                                         # entry j <-> bin j+1
    # zero Griffin-Lim iterations return the start signal (mel/mel.go:85)
    mel = oracle.to_mel(cfg, synth_clip(2, 0.3))
    frames = len(mel) // 192
    init = np.random.default_rng(3).random(4096 + (frames - 1) * 1280)
    assert np.array_equal(oracle.from_mel(oracle.config(gl_iters=0), mel, init), init)


def test_phase_roundtrip_is_not_identity(oracle):
    """only bins 1..NumFreqs are kept -> parity is judged against the reference output."""
    wav = np.random.default_rng(5).uniform(-1, 1, 48000)
    cfg = oracle.config(num_freqs=768)
    rec = oracle.from_phase(cfg, oracle.to_phase(cfg, wav))
    n = min(len(rec), len(wav))
    assert rel_l2(rec[:n], wav[:n]) > 0.5


def test_from_mel_rejects_ragged_input(oracle):
    with pytest.raises(ValueError):
        oracle.from_mel(oracle.config(), np.zeros((192 * 2 + 48, 2)), np.zeros(10))


# ---- image arithmetic ------------------------------------------------------------------------------
def test_dumpbuffer_and_quantise(oracle):
    rng = np.random.default_rng(9)
    buf = rng.standard_normal((192 * 5, 2))
    img = oracle.mel_image(buf, 192)
    mn, mx = buf.min(axis=0), buf.max(axis=0)
    exp = (np.trunc(255 * (buf[:, 0] - mn[0]) / (mx[0] - mn[0])).astype(np.uint16)
           | (np.trunc(255 * (buf[:, 1] - mn[1]) / (mx[1] - mn[1])).astype(np.uint16) << 8))
    assert np.array_equal(img, exp)
    const = np.full((192 * 2, 2), 3.06e-5)                               # max == min -> NaN -> int(NaN) -> 0
    assert np.all(oracle.mel_image(const, 192) == 0)
    # mel PNG pixels: round trip through quantise / dequantise stays within one quantisation step
    px = oracle.mel_quantise(buf, 192, True, 1289.4, 44100.0)
    back, samples, sr = oracle.mel_dequantise(px, True)
    step = (buf.max() - buf.min()) / 255
    assert np.abs(back - buf).max() < 1.02 * step + 0.01 * abs(buf).max()   # + float16 metadata rounding
    assert sr == 44096.0 and samples == oracle.f16_value(oracle.f16_bits(1289.4)) * 5


def test_phase_quantise_blue_wrap_and_hdr(oracle):
    rng = np.random.default_rng(10)
    buf = rng.standard_normal((768 * 3, 2))
    px8 = oracle.phase_quantise(buf, 768, False, 1000.0, 48000.0, 0, False)
    v0 = (buf[:, 0] - buf[:, 0].min()) / (buf[:, 0].max() - buf[:, 0].min())
    blue = (np.trunc(255 * (-v0)).astype(np.int64) & 0xFF).astype(np.uint8)      # uint8(int(255*(-val0)))
    got_blue = px8[:, :, 2].T.reshape(-1)
    keep = np.ones(len(blue), bool)
    keep[768 - 16:768] = False                                           # metadata bytes
    assert np.array_equal(got_blue[keep], blue[keep])
    px16 = oracle.phase_quantise(buf, 768, True, 1000.0, 48000.0, 0, True)
    assert px16.dtype == np.uint16 and px16[:, :, 3].min() == 65535
    back, _, sr = oracle.phase_dequantise(px16, True, 0, True)
    assert np.abs(back - buf).max() < 0.02 * np.abs(buf).max() and sr == 48000.0


def test_volume_boost_doubles_peak(oracle):
    """test_phase_comprehensive.py:184-195: volume boost 2.0 doubles the peak of the reconstruction"""
    wav = synth_clip(3, 0.6, sr=48000)
    spec = oracle.to_phase(oracle.config(num_freqs=768), wav)
    a = oracle.from_phase(oracle.config(num_freqs=768, volume_boost=1.0), spec)
    b = oracle.from_phase(oracle.config(num_freqs=768, volume_boost=2.0), spec)
    assert abs(np.abs(b).max() / np.abs(a).max() - 2.0) < 1e-12
    c = oracle.from_phase(oracle.config(num_freqs=768, volume_boost=0.0), spec)     # 0 = no boost (phase/phase.go:146)
    assert np.array_equal(a, c)


def test_oracle_is_clean_under_asan_ubsan(tmp_path):
    """the checker itself is checked -- every oracle entry point on small inputs (both
    geometries, edge lengths) in a build with AddressSanitizer + UndefinedBehaviorSanitizer"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "orc_san")
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined",
           "-o", exe, os.path.join(root, "tests", "c", "oracle_sanitize.c"), os.path.join(root, "oracle", "gomel_oracle.c"), "-lm"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("toolchain without sanitizer runtime")
    assert b.returncode == 0, b.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1", LD_PRELOAD=""))
    assert r.returncode == 0 and "oracle sanitize ok" in r.stdout, r.stdout + r.stderr
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr
