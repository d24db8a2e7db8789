"""GPU tests of the file-level API: WAV -> PNG -> WAV through the drop-in
classes, PNG pixel bytes against the oracle and against pixels written by the reference's own
phase.py (tests/golden), float16 metadata, trimming rules."""
import os

import numpy as np
import pytest

from util import rel_l2, synth_clip

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "phase_ref.npz")


def _mel():
    from gomel_b200 import NewMel
    m = NewMel()
    m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut, m.YReverse = 192, 0, 16000, 1280, 4096, True
    m.GriffinLimIterations, m.VolumeBoost = 2, 0.0
    return m


def test_tomelwav_png_pixels_match_oracle(tmp_path, oracle, ctx):
    """cmd/tomel configuration: ToMelWav writes the PNG the reference would (<= 1 LSB on boundary pixels)."""
    from gomel_b200 import codec
    wav = synth_clip(70, 2.0)
    wf, pf = str(tmp_path / "a.wav"), str(tmp_path / "a.png")
    codec.save_wav(wf, wav, 44100)
    buf, sr = codec.load_wav(wf)
    m = _mel()
    m.ToMelWav(wf, pf)
    px = codec.read_png(pf)
    ref_mel = oracle.to_mel(oracle.config(), buf)
    frames = len(ref_mel) // 192
    assert px.shape == (192, frames, 3)
    ref_px = oracle.mel_quantise(ref_mel, 192, True, float(len(buf) * 192) / float(len(ref_mel)), float(sr))
    d = np.abs(px[:, :, :2].astype(int) - ref_px[:, :, :2].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3              # fp32 mel vs float64 mel: boundary pixels only
    assert np.array_equal(px[:, 1:, 2], ref_px[:, 1:, 2])       # blue: zero except metadata column
    meta_ours, meta_ref = px[:8, 0, 2], ref_px[:8, 0, 2]        # y-reversed: metadata at the top-left
    assert np.array_equal(meta_ours[4:], meta_ref[4:])          # samples_in_mel, sample rate bytes
    assert np.abs(meta_ours[:4].astype(int) - meta_ref[:4].astype(int)).max() <= 1   # max/min float16 of fp32-vs-f64


def test_towavpng_matches_oracle_pipeline(tmp_path, oracle, ctx):
    from gomel_b200 import codec
    wav = synth_clip(71, 1.5)
    wf, pf, of = str(tmp_path / "b.wav"), str(tmp_path / "b.png"), str(tmp_path / "b_out.wav")
    codec.save_wav(wf, wav, 44100)
    m = _mel()
    m.ToMelWav(wf, pf)
    # decode parity: device de-quantisation == oracle loadpng arithmetic on the same pixels, bit for bit
    px = codec.read_png(pf)
    rgba = np.concatenate([px, np.full(px.shape[:2] + (1,), 255, np.uint8)], axis=2)
    obuf, osamples, osr = oracle.mel_dequantise(rgba, True)
    buf, samples, sr = codec.mel_load_png(pf, True)
    assert np.array_equal(buf, obuf) and samples == osamples and sr == osr == 44096.0
    # full ToWavPng with an injected start signal
    frames = len(buf) // 192
    ola = 4096 + (frames - 1) * 1280
    init = np.random.default_rng(2).random(ola)
    m.InitSignal = init
    m.VolumeBoost = 0.25
    m.SampleRate = 0
    m.ToWavPng(pf, of)
    assert m.SampleRate == 44096                                 # embedded float16 rate (mel/mel.go:231-233)
    got, got_sr = codec.load_wav(of)
    ref = oracle.from_mel(oracle.config(gl_iters=2), obuf + 0.25, init)
    if int(osamples) > 0 and oracle.is_padded(int(osamples), len(ref), 1280) and len(ref) > int(osamples):
        ref = ref[:int(osamples)]
    assert len(got) == len(ref)
    ref16 = np.clip(ref, -1, 1)
    assert np.abs(got * 32767.0 - ref16 * 32767.0).max() <= 1.5  # one 16-bit LSB


@pytest.mark.parametrize("ihs", [0, 2])
def test_phase_py_png_matches_reference_pixels(tmp_path, ihs, ctx):
    """pixels written by the reference's own save_image / values read by its load_image (golden)"""
    from gomel_b200 import codec
    from gomel_b200 import phase as P
    g = np.load(GOLD)
    f = str(tmp_path / "p.png")
    P.save_image(f, g["c48k_short_spec"], 768, 1289.4, 48000, True, False, ihs)
    px = codec.read_png(f)
    ref = g[f"img{ihs}_pixels"]
    assert px.shape == ref.shape
    d = np.abs(px.astype(int) - ref.astype(int))
    if ihs == 0:
        assert d.max() == 0                                      # bit-exact
    else:
        assert d.max() <= 1 and (d > 0).mean() < 1e-3            # asinh: libm vs CUDA ulps on boundary pixels
    buf, samples, sr, nf = P.load_image(f, True, False, ihs)
    assert (samples, sr, nf) == tuple(g[f"img{ihs}_meta"])
    if ihs == 0:
        assert np.array_equal(buf, g["img0_loaded"])
    else:
        assert rel_l2(buf, g["img2_loaded"]) < 1e-2              # differs only where a pixel moved by 1 LSB


@pytest.mark.parametrize("hdr,ihs", [(False, 0), (True, 0), (False, 2)])
def test_phase_go_png_matches_oracle(tmp_path, oracle, ctx, hdr, ihs):
    from gomel_b200 import codec
    rng = np.random.default_rng(5)
    spec = rng.standard_normal((768 * 7, 2)) * 20
    f = str(tmp_path / "g.png")
    codec.phase_dump_image_go(f, spec, 768, True, 1289.4, 48000.0, ihs, hdr)
    px = codec.read_png(f)
    ref = oracle.phase_quantise(spec, 768, True, 1289.4, 48000.0, ihs, hdr)
    d = np.abs(px.astype(int) - ref[:, :, :3].astype(int))
    if ihs == 0:
        assert d.max() == 0 and px.dtype == (np.uint16 if hdr else np.uint8)
    else:
        assert d[:, :, :2].max() <= 1
    pad = np.full(px.shape[:2] + (1,), 65535 if hdr else 255, px.dtype)
    obuf, osamples, osr = oracle.phase_dequantise(np.concatenate([px, pad], axis=2), True, ihs, hdr)
    buf, samples, sr = codec.phase_load_png_go(f, True, ihs, hdr)
    assert samples == osamples and sr == osr
    assert rel_l2(buf, obuf) < 1e-14


def test_phase_file_roundtrip(tmp_path, oracle, ctx):
    from gomel_b200 import Phase, codec
    wav = synth_clip(72, 1.2, sr=48000)
    wf, pf, of = str(tmp_path / "c.wav"), str(tmp_path / "c.png"), str(tmp_path / "c_out.wav")
    codec.save_wav(wf, wav, 48000)
    ph = Phase()
    ph.to_phase_wav(wf, pf)
    assert ph.num_freqs == 768
    rate = Phase().to_wav_png(pf, of)
    assert rate == 48000
    out, sr = codec.load_wav(of)
    frames = int((len(wav) + codec.pad_len(len(wav), 1280) - 4096) / 1280) + 1
    assert sr == 48000 and len(out) == min(4096 + (frames - 1) * 1280, len(wav))   # trimmed only if longer (phase.py:343)
    # 22.05 kHz input: zero-stuffed to 44.1 kHz, 836 bins
    wav2 = synth_clip(73, 1.0, sr=22050)
    codec.save_wav(wf, wav2, 22050)
    ph2 = Phase()
    ph2.to_phase_wav(wf, pf)
    assert ph2.num_freqs == 836
    assert Phase().to_wav_png(pf, of) == 22050                   # nearest standard rate to the float16 value


def test_config1_old_format_fixture_shape(tmp_path, oracle, ctx):
    """BASELINE configs[0] names the reference's glados PNG (183 x 80, older 3-channel format, every metadata
    byte 0x02).  With NumMels = 192 the Go reference panics (14,640 entries are not a multiple of 192,
    mel/impl.go:366-372): we return an error instead.  With NumMels = 80 the image decodes to the
    constant 3.06e-5 spectrogram (max == min) and runs the flat-spectrum Griffin-Lim path end to end.
    The fixture itself is not copied: an image with the same geometry and metadata bytes is synthesised."""
    from gomel_b200 import NewMel, _lib, codec
    w, h = 183, 80
    img = np.zeros((h, w, 3), np.uint8)
    rng = np.random.default_rng(0)
    img[:, :, 0] = rng.integers(0, 255, (h, w))
    img[:, :, 1] = rng.integers(0, 255, (h, w))
    img[:, :, 2] = 2                                             # every metadata byte 0x02 -> float16 0x0202 = 3.06e-5
    f = str(tmp_path / "old.png")
    codec.write_png(f, img)
    buf, samples, sr = codec.mel_load_png(f, True)
    assert buf.shape == (w * h, 2) and samples == 0.0 and abs(sr - 3.0637e-5) < 1e-8
    assert np.all(buf == buf[0, 0]) and abs(buf[0, 0] - 3.0637e-5) < 1e-8     # max == min: constant spectrogram
    m = NewMel()
    m.MelFmin, m.MelFmax, m.YReverse, m.Window, m.Resolut, m.GriffinLimIterations = 0, 16000, True, 1280, 4096, 2
    m.NumMels = 192
    with pytest.raises(_lib.GomelError):
        m.ToWavPng(f, str(tmp_path / "x.wav"))                   # the reference panics here
    m.NumMels = 80
    init = np.random.default_rng(1).random(4096 + (w - 1) * 1280)
    m.InitSignal = init
    got = m.FromMel(buf.copy())
    assert got.shape == (237056,)                                # OLA length for 183 frames
    ocfg = oracle.config(num_mels=80, gl_iters=2)
    ref = oracle.from_mel(ocfg, buf, init)
    assert rel_l2(got, ref) < 1e-4
    m.SampleRate = 44100
    m.ToWavPng(f, str(tmp_path / "old.wav"))                     # whole file path: PNG -> WAV
    out, osr = codec.load_wav(str(tmp_path / "old.wav"))
    assert osr == 44100 and len(out) == 237056


def test_batch_directory_tools_match_single_file_api(tmp_path, oracle, ctx):
    """tomel_dir / towav_dir (one batched GPU call for many files) == Mel.ToMelWav / Mel.ToWavPng per file"""
    from gomel_b200 import batch, codec
    ind, pngd, pngd1, wavd = tmp_path / "in", tmp_path / "png", tmp_path / "png1", tmp_path / "wav"
    for d in (ind, pngd1):
        d.mkdir()
    lens = {"a": 0.7, "b": 1.3, "c": 0.7, "d": 0.2}
    for k, (name, secs) in enumerate(lens.items()):
        codec.save_wav(str(ind / f"{name}.wav"), synth_clip(80 + k, secs), 44100)
    _flac_fixture(str(ind / "e.flac"), 0.5, 44100, True, 9)                  # a stereo FLAC file in the same directory
    m = _mel()
    out = batch.tomel_dir(str(ind), str(pngd), m, chunk=3)
    assert len(out) == 5
    m.ToMelFlac(str(ind / "e.flac"), str(pngd1 / "e.flac.png"))
    assert np.array_equal(codec.read_png(str(pngd / "e.flac.png")), codec.read_png(str(pngd1 / "e.flac.png")))
    os.remove(str(pngd / "e.flac.png"))
    for name in lens:
        m.ToMelWav(str(ind / f"{name}.wav"), str(pngd1 / f"{name}.wav.png"))
        a, b = codec.read_png(str(pngd / f"{name}.wav.png")), codec.read_png(str(pngd1 / f"{name}.wav.png"))
        assert a.shape == b.shape and np.array_equal(a, b)          # bit-identical pixels
    # back to audio, injected start signals so both paths are comparable
    inits = {}
    for name in lens:
        buf, _, _ = codec.mel_load_png(str(pngd / f"{name}.wav.png"), True)
        inits[f"{name}.wav.png"] = np.random.default_rng(hash(name) % 1000).random(4096 + (len(buf) // 192 - 1) * 1280)
    wavs = batch.towav_dir(str(pngd), str(wavd), m, init_signals=inits)
    assert len(wavs) == 4
    for name in lens:
        m1 = _mel()
        m1.InitSignal = inits[f"{name}.wav.png"].astype(np.float32).astype(np.float64)
        m1.ToWavPng(str(pngd / f"{name}.wav.png"), str(tmp_path / "single.wav"))
        x, _ = codec.load_wav(str(wavd / f"{name}.wav.png.wav"))
        y, _ = codec.load_wav(str(tmp_path / "single.wav"))
        assert len(x) == len(y) and np.abs(x - y).max() * 32767 <= 1.001   # float32 vs float64 mel input: <= 1 PCM LSB


# ------------------------------------------------------------------ FLAC entry points
def _flac_fixture(path, seconds, sr, stereo, seed):
    from gomel_b200 import flac
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    ch = [(9000 * np.sin(2 * np.pi * (330 + 110 * c) * t + c) + rng.normal(0, 300, n)).astype(np.int64) for c in range(2 if stereo else 1)]
    pcm = np.stack(ch, axis=1)
    flac.encode(path, pcm if stereo else pcm[:, 0], sr, bps=16, blocksize=4096, stereo_mode="mid_side" if stereo else "independent")
    return pcm


def test_tomelflac_is_tomel_of_the_go_decoded_samples(tmp_path, ctx):
    """mel.ToMelFlac (mel/mel.go:176-192): loadflac's block-wise channel concatenation scaled by 1/65536, then ToMel and
    dumpimage -- the PNG equals the one written from the same samples through the buffer API"""
    from gomel_b200 import codec
    ff, pf, qf = str(tmp_path / "s.flac"), str(tmp_path / "s.png"), str(tmp_path / "ref.png")
    pcm = _flac_fixture(ff, 1.3, 44100, True, 3)
    m = _mel()
    m.ToMelFlac(ff, pf)
    blocks = [pcm[i:i + 4096] for i in range(0, len(pcm), 4096)]
    buf = np.concatenate([b[:, c] for b in blocks for c in range(2)]) / 65536.0
    spec = m.ToMel(buf)
    codec.mel_dump_image(qf, spec, 192, True, float(len(buf) * 192) / float(len(spec)), 44100.0)
    assert np.array_equal(codec.read_png(pf), codec.read_png(qf))
    from gomel_b200 import mel as M
    assert np.array_equal(M.LoadFlac(ff), buf)
    with pytest.raises(M.ErrFileNotLoaded):
        m.ToMelFlac(str(tmp_path / "missing.flac"), pf)


def test_python_phase_flac_entry_points(tmp_path, ctx):
    """phase.py:255-318: to_tensor_flac == to_phase(zero-stuffed soundfile samples); to_phase_flac writes the PNG that
    save_image writes for that spectrogram, with the sample rate following the stuffing ratio (22050 -> 44100)"""
    from gomel_b200 import phase as P
    ff, pf, qf = str(tmp_path / "p.flac"), str(tmp_path / "p.png"), str(tmp_path / "q.png")
    pcm = _flac_fixture(ff, 1.0, 22050, True, 4)
    audio = np.mean(pcm / 32768.0, axis=1)
    ph = P.Phase()
    spec = ph.to_tensor_flac(ff)
    assert ph.num_freqs == 836 and ph.family is False
    up = P.zero_stuff_upsample(audio, 1, 1)
    want = P.Phase(sample_rate=22050).to_phase(up)
    assert spec.shape == want.shape and np.array_equal(spec, want)
    ph2 = P.Phase()
    ph2.to_phase_flac(ff, pf)
    P.save_image(qf, want, 836, float(len(up) * 836) / float(len(want)), 44100, True, False, 0)
    assert np.array_equal(P.codec.read_png(pf), P.codec.read_png(qf))
    back, samples, rate, nfq = P.load_image(pf, True, False, 0)
    assert nfq == 836 and abs(rate - 44100) < 64 and back.shape == want.shape


def test_go_phase_flac_and_wav_tools_round_trip(tmp_path, ctx, oracle):
    """cmd/tophase + cmd/fromphase on the Go names: ToPhaseFlac (1/32768, length before stuffing in the metadata),
    ToWavPng (trim only if isPadded, family main rate, truncating 16-bit writer)"""
    from gomel_b200 import codec
    from gomel_b200.cli import _phase
    ff, pf, wf = str(tmp_path / "g.flac"), str(tmp_path / "g.png"), str(tmp_path / "g.wav")
    pcm = _flac_fixture(ff, 1.1, 48000, False, 5)
    p = _phase()
    p.ToPhaseFlac(ff, pf)
    buf, samples, sr = codec.phase_load_png_go(pf, True, p.ihsPasses(), p.HDR)
    frames = len(buf) // p.num_freqs
    assert abs(samples / frames * frames - len(pcm)) < 0.002 * len(pcm) and abs(sr - 48000) < 64     # float16 metadata
    q = _phase()
    q.ToWavPng(pf, wf)
    assert q.sample_rate == 48000
    got, got_sr = codec.load_wav(wf)
    assert got_sr == 48000.0
    ref = oracle.from_phase(oracle.config(num_freqs=p.num_freqs), buf)
    n = min(len(got), len(ref))
    assert len(got) in (len(ref), int(samples))
    assert np.abs(got[:n] * 32767.0 - np.clip(ref[:n], -1, 1) * 32767.0).max() <= 1.5
