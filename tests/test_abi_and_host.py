"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol
include/gomel_cuda.h declares (no compute calls), sizing arithmetic, host-side helpers of the
drop-in classes, codecs.  No GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from gomel_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gomel_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gomel_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    L = C.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/gomel_cuda.h but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)     # the ctypes binding covers the ABI
    assert b"sm_100a" in lib.load().gomel_version()


def test_frames_matches_oracle(lib, oracle):
    cfg = lib.make_config()
    for n in (1, 100, 19198, 19199, 19200, 20479, 20480, 20481, 44100, 441000, 158760000):
        npad, fr, ola = lib.frames(cfg, n)
        assert npad == n + oracle.pad_len(n, 1280)
        assert fr == oracle.num_frames(n, oracle.config())
        assert ola == 4096 + (fr - 1) * 1280
    assert lib.frames(cfg, 441000) == (441599, 342, 440576)
    assert lib.frames(cfg, 158760000) == (158760959, 124029, 158759936)
    with pytest.raises(lib.GomelError):
        lib.frames(cfg, 0)


def test_mel_tables_match_oracle(lib, oracle):
    flo, fhi, fmod, ilo, ihi, imod = lib.mel_tables(2048, 192, 0.0, 16000.0)
    lo, hi, mod = oracle.mel_fwd_tables(2048, 192, 0.0, 16000.0)
    assert np.array_equal(flo, lo) and np.array_equal(fhi, hi) and np.array_equal(fmod, mod)
    lo, hi, mod = oracle.mel_inv_tables(2048, 192, 0.0, 16000.0)
    assert np.array_equal(ilo, lo) and np.array_equal(ihi, hi) and np.array_equal(imod, mod)


def test_config_struct_layout_matches_header(lib):
    # int x5, double x3, int, double x2: natural alignment -> 72 bytes on LP64
    assert C.sizeof(lib.Config) == 72
    assert lib.Config.tune_mul.offset == 24 and lib.Config.flags.offset == 48
    assert lib.Config.mel_fmin.offset == 56 and lib.Config.mel_fmax.offset == 64


def test_no_gpu_means_loud_failure(lib):
    """Without a CUDA device context creation raises; nothing falls back to a CPU path."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from gomel_b200 import _lib\n"
            "try:\n    _lib.Context(0)\n    print('CTX')\n"
            "except _lib.GomelError as e:\n    print('RAISED', e.code)\n"
            "try:\n    _lib.default_context(0)\n    print('CTX')\n"
            "except _lib.GomelError as e:\n    print('RAISED2', e.code)\n") % ROOT
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120).stdout
    assert "RAISED" in out and "RAISED2" in out and "CTX" not in out


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gomel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_phase_module_helpers_match_reference_golden():
    from gomel_b200 import phase as P
    g = np.load(os.path.join(ROOT, "tests", "golden", "phase_ref.npz"))
    assert np.array_equal(P.shrink(g["shrink_in"], 4096, 768), g["shrink_out"])
    assert np.array_equal(P.grow(g["shrink_out"], 4096, 768), g["grow_out"])
    assert P.shrink(np.zeros((10 * 2048, 2)), 4096, 768).shape == (7680, 2)          # test_phase_comprehensive.py:66-70
    assert P.grow(np.zeros((7680, 2)), 4096, 768).shape == (20480, 2)
    for i in range(4):
        zp, zs = (int(x) for x in g[f"zs{i}_args"])
        assert np.array_equal(P.zero_stuff_upsample(g[f"zs{i}_in"], zp, zs), g[f"zs{i}_out"])
    for n, padded, yes in zip(g["pad_in"], g["pad_out"], g["is_padded"]):
        if n < 10 ** 6:
            assert len(P.pad(np.zeros(int(n)), 1280)) == int(padded)
        assert P.is_padded(int(n), int(padded), 1280) == bool(yes)
    for i, v in enumerate(g["f16_in"]):
        assert P.pack_float16_to_bytes(float(v)) == bytes(g["f16_bytes"][2 * i:2 * i + 2])


def test_phase_constructor_mirrors_reference():
    from gomel_b200 import Phase
    assert Phase().num_freqs == 0 and Phase().window == 1280 and Phase().resolut == 4096      # phase.py:37-43
    assert Phase(sample_rate=48000).num_freqs == 768 and Phase(sample_rate=44100).num_freqs == 836
    assert Phase(sample_rate=16000, HDR=True).num_freqs == 1536 and Phase(sample_rate=22050, HDR=True).num_freqs == 1672
    assert Phase(IHS=True).IHS == 2 and Phase(IHS=True, HDR=True).IHS == 0
    with pytest.raises(ValueError):
        Phase(sample_rate=12345)
    ph = Phase(sample_rate=32000)
    assert ph.pad_shift(32000) == (2, 1) and ph.zero_pad(8000) == 1 and ph.zero_shift(8000) == 5
    with pytest.raises(ValueError):
        ph.pad_shift(44100)                                                           # wrong family


def test_mel_defaults_mirror_newmel():
    from gomel_b200 import NewMel
    m = NewMel()                                                                      # mel/mel.go:30-41
    assert (m.NumMels, m.MelFmin, m.MelFmax, m.TuneMul, m.TuneAdd, m.Window, m.Resolut, m.GriffinLimIterations) == \
        (160, 0, 8000, 1, 0, 256, 2048, 2)


def test_png_and_wav_containers(tmp_path):
    from gomel_b200 import codec
    rng = np.random.default_rng(0)
    for dt, ch in ((np.uint8, 4), (np.uint8, 3), (np.uint16, 4), (np.uint16, 3)):
        px = rng.integers(0, np.iinfo(dt).max, (37, 23, ch)).astype(dt)
        f = str(tmp_path / f"t_{np.dtype(dt).name}_{ch}.png")
        codec.write_png(f, px)
        back = codec.read_png(f)
        assert back.dtype == dt and np.array_equal(back, px[:, :, :3])
    from PIL import Image
    img = np.array(Image.open(str(tmp_path / "t_uint8_4.png")))                      # readable by a stock decoder
    assert img.shape == (37, 23, 4)
    x = np.sin(np.arange(5000) / 7.0) * 0.8
    codec.save_wav(str(tmp_path / "a.wav"), x, 44100)
    y, sr = codec.load_wav(str(tmp_path / "a.wav"))
    assert sr == 44100 and len(y) == len(x) and np.abs(y - x).max() < 1e-4
    assert codec.unpack_f16(codec.pack_f16_go(44100.0)) == 44096.0


def test_reference_fixture_is_the_old_format():
    """glados PNG is 183x80; with NumMels=192 the reference panics -> we refuse the length."""
    p = "/root/reference/glados-1609757458000_.png"
    if not os.path.exists(p):
        pytest.skip("reference checkout not present on this box")
    from PIL import Image
    im = Image.open(p)
    assert im.size == (183, 80)
    assert (183 * 80) % 192 != 0


def test_png16_reader_handles_all_filter_types(tmp_path):
    """Go's png.Encode picks a filter per row; the 16-bit reader (used for HDR images written by the Go reference)
    must undo all five.  Rows are filtered here by hand, one type per row."""
    import struct
    import zlib
    from gomel_b200 import codec
    rng = np.random.default_rng(1)
    h, w, ch = 10, 7, 3
    px = rng.integers(0, 65535, (h, w, ch)).astype(np.uint16)
    raw = px.astype(">u2").reshape(h, -1).view(np.uint8).astype(np.int32)       # (h, w*6) bytes
    bpp = ch * 2
    out = bytearray()
    prev = np.zeros(raw.shape[1], np.int32)
    for y in range(h):
        ft = y % 5
        cur = raw[y]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        b = prev
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            pred = np.zeros_like(cur)
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = b
        elif ft == 3:
            pred = (a + b) >> 1
        else:
            p = a + b - c
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        out.append(ft)
        out.extend(((cur - pred) & 255).astype(np.uint8).tobytes())
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    f = str(tmp_path / "filtered16.png")
    with open(f, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0))
                 + chunk(b"IDAT", zlib.compress(bytes(out))) + chunk(b"IEND", b""))
    back = codec.read_png(f)
    assert back.dtype == np.uint16 and np.array_equal(back, px)


def test_save_wav_pcm16_is_the_container_half_of_save_wav(tmp_path):
    """codec.save_wav = dumpwav's clamp + int16(v * 32767) (mel/impl.go:195-232) followed by save_wav_pcm16"""
    from gomel_b200 import codec
    x = np.concatenate([np.linspace(-1.5, 1.5, 4001), [0.0, 1e-9, -1e-9, 0.99999, -0.99999]])
    a, b = str(tmp_path / "a.wav"), str(tmp_path / "b.wav")
    codec.save_wav(a, x, 22050)
    pcm = (np.clip(x, -1.0, 1.0) * 32767.0).astype(np.int16)          # truncation toward zero, like Go's int16()
    codec.save_wav_pcm16(b, pcm, 22050)
    assert open(a, "rb").read() == open(b, "rb").read()
    y, sr = codec.load_wav(a)
    assert sr == 22050 and len(y) == len(x) and y[0] == -1.0 and pcm[0] == -32767 and pcm[-1] == -32766
    assert pcm[-4] == 0 and pcm[-3] == 0                                  # +-1e-9 truncate toward zero
