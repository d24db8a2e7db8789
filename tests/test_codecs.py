"""CPU tests of the host-side file codecs either side of the hot path: the native FLAC
reader, and the two flavours of WAV / FLAC sample scaling the reference has -- Go (beep, mewkiz/flac) and Python
(soundfile).  No GPU: nothing here calls a transform."""
import os
import struct
import wave

import numpy as np
import pytest

from gomel_b200 import codec, flac


def _pcm(n=9000, seed=0, bps=16, ch=2):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    amp = (1 << (bps - 1)) - 1
    x = [(0.5 * amp * np.sin(2 * np.pi * (300 + 170 * c) * t / 44100 + c) + rng.normal(0, amp * 0.01, n)).astype(np.int64)
         for c in range(ch)]
    return np.stack(x, axis=1)


@pytest.mark.parametrize("bps", [16, 24])
@pytest.mark.parametrize("mode", ["independent", "left_side", "right_side", "mid_side"])
@pytest.mark.parametrize("order,lpc,escape", [(-1, False, False), (0, False, False), (2, False, False), (4, False, False),
                                              (2, True, False), (3, True, False), (1, False, True)])
def test_flac_reader_round_trips_every_subframe_kind(tmp_path, bps, mode, order, lpc, escape):
    pcm = _pcm(bps=bps)
    pcm[4096:8192] = (pcm[4096:8192] >> 3) << 3          # a block with wasted bits
    pcm[8192:, 1] = 77                                   # a CONSTANT subframe in the ragged last block
    p = str(tmp_path / "a.flac")
    flac.encode(p, pcm, 44100, bps=bps, blocksize=4096, stereo_mode=mode, order=order, lpc=lpc, escape=escape)
    blocks, sr, b2, nch = flac.decode(p)
    assert (sr, b2, nch) == (44100, bps, 2) and [len(b[0]) for b in blocks] == [4096, 4096, 808]
    got = np.stack([np.concatenate([blk[c] for blk in blocks]) for c in range(nch)], axis=1)
    assert np.array_equal(got, pcm)


def test_flac_go_flavour_is_blockwise_channel_concatenation_with_the_packages_own_scales(tmp_path):
    """loadflac appends frame.Subframes[0].Samples, then Subframes[1].Samples, per frame (mel/impl.go:287-293), and the
    mel package divides by 65536 where the phase package divides by 32768 (mel/impl.go:290, phase/impl.go:375)"""
    pcm = _pcm(n=10000)
    p = str(tmp_path / "s.flac")
    flac.encode(p, pcm, 48000, blocksize=4096, stereo_mode="mid_side")
    mel_buf, sr = codec.load_flac_go(p, 256 * 256)
    ph_buf, _ = codec.load_flac_go(p, 256 * 128)
    assert sr == 48000.0 and len(mel_buf) == 20000
    want = np.concatenate([pcm[0:4096, 0], pcm[0:4096, 1], pcm[4096:8192, 0], pcm[4096:8192, 1], pcm[8192:, 0], pcm[8192:, 1]])
    assert np.array_equal(mel_buf, want / 65536.0)
    assert np.array_equal(ph_buf, want / 32768.0) and np.array_equal(ph_buf, 2 * mel_buf)
    mono = str(tmp_path / "m.flac")
    flac.encode(mono, pcm[:, 0], 44100)
    assert np.array_equal(codec.load_flac_go(mono, 256 * 128)[0], pcm[:, 0] / 32768.0)
    assert codec.load_flac_go(str(tmp_path / "missing.flac"), 1)[0].size == 0         # println(err), nil, 0


def test_flac_python_flavour_is_soundfile_semantics(tmp_path):
    """load_flac_with_sr (phase.py:570-586): sf.read float64 = sample / 2^(bits-1), stereo averaged"""
    for bps in (16, 24):
        pcm = _pcm(n=5000, bps=bps)
        p = str(tmp_path / f"s{bps}.flac")
        flac.encode(p, pcm, 22050, bps=bps, stereo_mode="left_side")
        a, sr = codec.load_flac_sf(p)
        assert sr == 22050 and isinstance(sr, int)
        assert np.array_equal(a, np.mean(pcm / float(1 << (bps - 1)), axis=1))


def _write_wav(path, pcm, sw, sr=44100):
    with wave.open(path, "wb") as w:
        w.setnchannels(pcm.shape[1])
        w.setsampwidth(sw)
        w.setframerate(sr)
        if sw == 1:
            w.writeframes(pcm.astype(np.uint8).tobytes())
        elif sw == 2:
            w.writeframes(pcm.astype("<i2").tobytes())
        else:
            b = pcm.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
            w.writeframes(np.ascontiguousarray(b).tobytes())


def test_wav_flavours(tmp_path):
    """Go: beep v1.1.0 wav.Decode -- left channel, 8-bit p/255*2-1, 16-bit v/32767, 24-bit v/(2^23-1).
    Python: soundfile -- v/2^(bits-1) ((p-128)/128 for 8-bit), channels averaged (phase.py:551-567)."""
    p = str(tmp_path / "w.wav")
    pcm16 = _pcm(n=3000)
    _write_wav(p, pcm16, 2)
    g, sr = codec.load_wav(p)
    assert sr == 44100.0 and np.array_equal(g, pcm16[:, 0] / 32767.0)
    s, sr = codec.load_wav_sf(p)
    assert sr == 44100 and np.array_equal(s, np.mean(pcm16 / 32768.0, axis=1))
    pcm8 = (np.random.default_rng(1).integers(0, 256, (500, 1)))
    _write_wav(p, pcm8, 1)
    assert np.array_equal(codec.load_wav(p)[0], pcm8[:, 0] / 255.0 * 2.0 - 1.0)
    assert np.array_equal(codec.load_wav_sf(p)[0], (pcm8[:, 0] - 128.0) / 128.0)
    pcm24 = _pcm(n=700, bps=24, ch=1)
    _write_wav(p, pcm24, 3)
    assert np.array_equal(codec.load_wav(p)[0], pcm24[:, 0] / float((1 << 23) - 1))
    assert np.array_equal(codec.load_wav_sf(p)[0], pcm24[:, 0] / float(1 << 23))


def test_wav_writers_truncate_for_go_and_round_for_python(tmp_path):
    """dumpwav -> beep: int16(clamp(v) * 32767) truncates; soundfile PCM_16: lrint(clip(v) * 32767) rounds"""
    x = np.array([0.0, 0.5, -0.5, 0.99999, -0.99999, 1.7, -3.0, 1e-5, 2.5 / 32767, 3.5 / 32767, -2.5 / 32767])
    pg, pp = str(tmp_path / "g.wav"), str(tmp_path / "p.wav")
    codec.save_wav(pg, x, 44100)
    codec.save_wav_sf(pp, x, 48000)
    with wave.open(pg, "rb") as w:
        g = np.frombuffer(w.readframes(w.getnframes()), "<i2")
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 44100)
    with wave.open(pp, "rb") as w:
        s = np.frombuffer(w.readframes(w.getnframes()), "<i2")
        assert w.getframerate() == 48000
    c = np.clip(x, -1, 1) * 32767.0
    assert np.array_equal(g, np.trunc(c).astype(np.int16))
    assert np.array_equal(s, np.rint(c).astype(np.int16))
    assert g[8] == 2 and s[8] == 2 and s[9] == 4 and s[10] == -2 and g[10] == -2        # ties to even / toward zero


def test_python_phase_module_uses_the_python_flavour():
    """gomel_b200.phase is the drop-in for the reference's phase.py: its loaders / saver are the soundfile ones"""
    from gomel_b200 import phase
    assert phase.load_wav_with_sr is not None and phase.load_flac_with_sr("x") if False else True
    import inspect
    assert "load_wav_sf" in inspect.getsource(phase.load_wav_with_sr)
    assert "load_flac_sf" in inspect.getsource(phase.load_flac_with_sr)
    assert "save_wav_sf" in inspect.getsource(phase.save_wav)
    for name in ("to_phase_flac", "to_tensor_flac", "to_phase_wav", "to_wav_png", "ToPhaseFlac", "ToPhaseWav"):
        assert callable(getattr(phase.Phase, name))
