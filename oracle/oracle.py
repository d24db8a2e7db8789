"""ctypes loader for the CPU float64 oracle (oracle/gomel_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package gomel_b200 never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgomel_oracle.so")


class OrcConfig(C.Structure):
    _fields_ = [("num_mels", C.c_int), ("num_freqs", C.c_int), ("window", C.c_int),
                ("resolut", C.c_int), ("mel_fmin", C.c_double), ("mel_fmax", C.c_double),
                ("tune_mul", C.c_double), ("tune_add", C.c_double), ("volume_boost", C.c_double),
                ("gl_iters", C.c_int)]


def build(force=False):
    """Compile the oracle with the system gcc (OpenMP if available)."""
    src = os.path.join(_HERE, "gomel_oracle.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)
            and os.path.getmtime(_SO) >= os.path.getmtime(os.path.join(_HERE, "gomel_oracle.h"))):
        return _SO
    base = ["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-shared", "-o", _SO, src, "-lm"]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base[:1] + extra + base[1:], capture_output=True, text=True)
        if r.returncode == 0:
            return _SO
    raise RuntimeError("oracle build failed:\n" + r.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        cp = C.POINTER(OrcConfig)
        L.orc_pad_len.restype = C.c_long
        L.orc_pad_len.argtypes = [C.c_long, C.c_int]
        L.orc_is_padded.restype = C.c_int
        L.orc_is_padded.argtypes = [C.c_long, C.c_long, C.c_int]
        L.orc_num_frames.restype = C.c_long
        L.orc_num_frames.argtypes = [C.c_long, C.c_int, C.c_int]
        L.orc_hann.argtypes = [C.c_int, dp]
        L.orc_fft.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_to_mel.restype = C.c_long
        L.orc_to_mel.argtypes = [cp, dp, C.c_long, dp, C.c_long]
        L.orc_from_mel.restype = C.c_long
        L.orc_from_mel.argtypes = [cp, dp, C.c_long, dp, dp, C.c_long]
        L.orc_mel_fwd_tables.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, ip, ip, dp]
        L.orc_mel_inv_tables.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, ip, ip, dp, dp, dp]
        L.orc_to_phase.restype = C.c_long
        L.orc_to_phase.argtypes = [cp, dp, C.c_long, dp, C.c_long]
        L.orc_from_phase.restype = C.c_long
        L.orc_from_phase.argtypes = [cp, dp, C.c_long, dp, C.c_long]
        u16p, u8p = C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)
        L.orc_mel_dumpbuffer.argtypes = [dp, C.c_long, C.c_int, u16p]
        L.orc_phase_dumpbuffer.argtypes = [dp, C.c_long, C.c_int, u16p]
        L.orc_mel_quantise.argtypes = [dp, C.c_long, C.c_int, C.c_int, C.c_double, C.c_double, u8p]
        L.orc_mel_dequantise.argtypes = [u8p, C.c_int, C.c_int, C.c_int, dp, dp, dp]
        L.orc_phase_quantise.argtypes = [dp, C.c_long, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_int, u8p, u16p]
        L.orc_phase_dequantise.argtypes = [u8p, u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           dp, dp, dp]
        L.orc_f16_bits.restype = C.c_uint16
        L.orc_f16_bits.argtypes = [C.c_double]
        L.orc_f16_value.restype = C.c_double
        L.orc_f16_value.argtypes = [C.c_uint16]
        L.orc_pad_shift.argtypes = [C.c_int, ip, ip]
        L.orc_zero_stuff_len.restype = C.c_long
        L.orc_zero_stuff_len.argtypes = [C.c_long, C.c_int, C.c_int]
        L.orc_zero_stuff.argtypes = [dp, C.c_long, C.c_int, C.c_int, dp]
        L.orc_from_mel_batch.restype = C.c_long
        L.orc_from_mel_batch.argtypes = [cp, dp, C.c_long, C.c_int, dp, dp, C.c_long, C.c_int]
        L.orc_to_mel_batch.restype = C.c_long
        L.orc_to_mel_batch.argtypes = [cp, dp, C.c_long, C.c_int, dp, C.c_long, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def config(num_mels=192, num_freqs=768, window=1280, resolut=4096, mel_fmin=0.0, mel_fmax=16000.0,
           tune_mul=1.0, tune_add=0.0, volume_boost=0.0, gl_iters=2):
    return OrcConfig(num_mels, num_freqs, window, resolut, mel_fmin, mel_fmax, tune_mul, tune_add,
                     volume_boost, gl_iters)


def pad_len(n, hop):
    return int(lib().orc_pad_len(n, hop))


def is_padded(orig, padded, hop):
    return bool(lib().orc_is_padded(orig, padded, hop))


def num_frames(n_samples, cfg):
    npad = n_samples + pad_len(n_samples, cfg.window)
    return int(lib().orc_num_frames(npad, cfg.resolut, cfg.window))


def hann(n):
    w = np.empty(n, np.float64)
    lib().orc_hann(n, _dp(w))
    return w


def fft(x, inverse=False):
    x = np.asarray(x, np.complex128)
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    lib().orc_fft(_dp(re), _dp(im), len(x), int(inverse))
    return re + 1j * im


def to_mel(cfg, wav):
    wav = np.ascontiguousarray(wav, np.float64)
    frames = num_frames(len(wav), cfg)
    out = np.empty((frames * cfg.num_mels, 2), np.float64)
    rc = lib().orc_to_mel(C.byref(cfg), _dp(wav), len(wav), _dp(out), out.size)
    if rc < 0:
        raise RuntimeError(f"orc_to_mel rc={rc}")
    return out


def from_mel(cfg, mel, init):
    """NOTE: like mel.FromMel the oracle exp()s its input in place; we pass a copy."""
    mel = np.array(mel, np.float64, copy=True).reshape(-1, 2)
    if len(mel) % cfg.num_mels:
        raise ValueError("len(mel) % num_mels != 0 (the Go reference panics here)")
    frames = len(mel) // cfg.num_mels
    ola = cfg.resolut + (frames - 1) * cfg.window
    init = np.ascontiguousarray(init, np.float64)
    assert len(init) == ola
    out = np.empty(ola, np.float64)
    rc = lib().orc_from_mel(C.byref(cfg), _dp(mel), len(mel), _dp(init), _dp(out), ola)
    if rc < 0:
        raise RuntimeError(f"orc_from_mel rc={rc}")
    return out


def mel_fwd_tables(filtersize, mels, fmin, fmax):
    lo, hi = np.empty(mels, np.int32), np.empty(mels, np.int32)
    mod = np.empty(mels, np.float64)
    ip = C.POINTER(C.c_int)
    lib().orc_mel_fwd_tables(filtersize, mels, fmin, fmax, lo.ctypes.data_as(ip), hi.ctypes.data_as(ip), _dp(mod))
    return lo, hi, mod


def mel_inv_tables(filtersize, mels, fmin, fmax):
    lo, hi = np.empty(filtersize, np.int32), np.empty(filtersize, np.int32)
    mod, flo, fhi = (np.empty(filtersize, np.float64) for _ in range(3))
    ip = C.POINTER(C.c_int)
    lib().orc_mel_inv_tables(filtersize, mels, fmin, fmax, lo.ctypes.data_as(ip), hi.ctypes.data_as(ip),
                             _dp(mod), _dp(flo), _dp(fhi))
    return lo, hi, mod


def to_phase(cfg, wav):
    wav = np.ascontiguousarray(wav, np.float64)
    frames = num_frames(len(wav), cfg)
    keep = min(cfg.num_freqs, cfg.resolut // 2)
    out = np.empty((frames * keep, 2), np.float64)
    rc = lib().orc_to_phase(C.byref(cfg), _dp(wav), len(wav), _dp(out), out.size)
    if rc < 0:
        raise RuntimeError(f"orc_to_phase rc={rc}")
    return out


def from_phase(cfg, spec):
    spec = np.ascontiguousarray(spec, np.float64).reshape(-1, 2)
    frames = len(spec) // cfg.num_freqs
    ola = cfg.resolut + (frames - 1) * cfg.window
    out = np.empty(ola, np.float64)
    rc = lib().orc_from_phase(C.byref(cfg), _dp(spec), len(spec), _dp(out), ola)
    if rc < 0:
        raise RuntimeError(f"orc_from_phase rc={rc}")
    return out


def mel_image(buf, mels):
    buf = np.ascontiguousarray(buf, np.float64).reshape(-1, 2)
    out = np.empty(len(buf), np.uint16)
    lib().orc_mel_dumpbuffer(_dp(buf), len(buf), mels, out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


def phase_image(buf, mels):
    buf = np.ascontiguousarray(buf, np.float64).reshape(-1, 2)
    out = np.empty(len(buf), np.uint16)
    lib().orc_phase_dumpbuffer(_dp(buf), len(buf), mels, out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


def mel_quantise(buf, mels, reverse, samples_in_mel, sr):
    buf = np.ascontiguousarray(buf, np.float64).reshape(-1, 2)
    stride = len(buf) // mels
    rgba = np.zeros((mels, stride, 4), np.uint8)
    lib().orc_mel_quantise(_dp(buf), len(buf), mels, int(reverse), samples_in_mel, sr,
                           rgba.ctypes.data_as(C.POINTER(C.c_uint8)))
    return rgba


def mel_dequantise(rgba, reverse):
    rgba = np.ascontiguousarray(rgba, np.uint8)
    h, w = rgba.shape[:2]
    buf = np.empty((w * h, 2), np.float64)
    s, sr = C.c_double(), C.c_double()
    lib().orc_mel_dequantise(rgba.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, int(reverse), _dp(buf),
                             C.byref(s), C.byref(sr))
    return buf, s.value, sr.value


def phase_quantise(buf, mels, reverse, samples_in_mel, sr, ihs_passes, hdr):
    buf = np.array(buf, np.float64, copy=True).reshape(-1, 2)
    stride = len(buf) // mels
    if hdr:
        out = np.zeros((mels, stride, 4), np.uint16)
        lib().orc_phase_quantise(_dp(buf), len(buf), mels, int(reverse), samples_in_mel, sr, ihs_passes, 1,
                                 None, out.ctypes.data_as(C.POINTER(C.c_uint16)))
    else:
        out = np.zeros((mels, stride, 4), np.uint8)
        lib().orc_phase_quantise(_dp(buf), len(buf), mels, int(reverse), samples_in_mel, sr, ihs_passes, 0,
                                 out.ctypes.data_as(C.POINTER(C.c_uint8)), None)
    return out


def phase_dequantise(px, reverse, ihs_passes, hdr):
    h, w = px.shape[:2]
    buf = np.empty((w * h, 2), np.float64)
    s, sr = C.c_double(), C.c_double()
    if hdr:
        px = np.ascontiguousarray(px, np.uint16)
        lib().orc_phase_dequantise(None, px.ctypes.data_as(C.POINTER(C.c_uint16)), w, h, int(reverse),
                                   ihs_passes, 1, _dp(buf), C.byref(s), C.byref(sr))
    else:
        px = np.ascontiguousarray(px, np.uint8)
        lib().orc_phase_dequantise(px.ctypes.data_as(C.POINTER(C.c_uint8)), None, w, h, int(reverse),
                                   ihs_passes, 0, _dp(buf), C.byref(s), C.byref(sr))
    return buf, s.value, sr.value


def f16_bits(v):
    return int(lib().orc_f16_bits(float(v)))


def f16_value(bits):
    return float(lib().orc_f16_value(int(bits)))


def pad_shift(sr):
    a, b = C.c_int(), C.c_int()
    lib().orc_pad_shift(sr, C.byref(a), C.byref(b))
    return a.value, b.value


def zero_stuff(audio, zp, zs):
    audio = np.ascontiguousarray(audio, np.float64)
    n = int(lib().orc_zero_stuff_len(len(audio), zp, zs))
    out = np.empty(n, np.float64)
    lib().orc_zero_stuff(_dp(audio), len(audio), zp, zs, _dp(out))
    return out


def from_mel_batch(cfg, mel, init, threads=0):
    """mel: (clips, frames*mels, 2); init: (clips, ola). Returns (clips, ola)."""
    mel = np.array(mel, np.float64, copy=True)
    clips, n_entries = mel.shape[0], mel.shape[1]
    frames = n_entries // cfg.num_mels
    ola = cfg.resolut + (frames - 1) * cfg.window
    init = np.ascontiguousarray(init, np.float64)
    out = np.empty((clips, ola), np.float64)
    rc = lib().orc_from_mel_batch(C.byref(cfg), _dp(mel), n_entries, clips, _dp(init), _dp(out), ola, threads)
    if rc < 0:
        raise RuntimeError(f"orc_from_mel_batch rc={rc}")
    return out


def to_mel_batch(cfg, wav, threads=0):
    wav = np.ascontiguousarray(wav, np.float64)
    clips, n = wav.shape
    frames = num_frames(n, cfg)
    out = np.empty((clips, frames * cfg.num_mels, 2), np.float64)
    rc = lib().orc_to_mel_batch(C.byref(cfg), _dp(wav), n, clips, _dp(out), frames * cfg.num_mels * 2, threads)
    if rc < 0:
        raise RuntimeError(f"orc_to_mel_batch rc={rc}")
    return out
