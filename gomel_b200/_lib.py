"""ctypes binding of libgomelcuda.so (include/gomel_cuda.h).  Fails loudly: a missing library,
a missing symbol or a missing GPU raises -- there is no CPU path behind this module."""
import ctypes as C
import math
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GOMEL_CUDA_LIB") or os.path.join(_HERE, "libgomelcuda.so")   # override: A/B builds

OK, E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED, E_STATE = 0, -1, -2, -3, -4, -5
Q_SINGLE_MINMAX, Q_HDR, Q_BLUE_WRAP = 1, 2, 4


class GomelError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gomel error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """gomel_config (include/gomel_cuda.h)"""
    _fields_ = [("n_fft", C.c_int), ("hop", C.c_int), ("n_mels", C.c_int), ("n_freqs", C.c_int),
                ("gl_iters", C.c_int), ("tune_mul", C.c_double), ("tune_add", C.c_double),
                ("volume_boost", C.c_double), ("flags", C.c_int), ("mel_fmin", C.c_double), ("mel_fmax", C.c_double)]


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_cp = C.POINTER(Config)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/gomel_cuda.h declares
SIGNATURES = {
    "gomel_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "gomel_ctx_destroy": (None, [_vp]),
    "gomel_last_error": (C.c_char_p, [_vp]),
    "gomel_version": (C.c_char_p, []),
    "gomel_launch_count": (C.c_ulonglong, [_vp]),
    "gomel_set_tile_frames": (C.c_int, [_vp, C.c_int]),
    "gomel_set_lead_f64": (C.c_int, [_vp, C.c_int]),
    "gomel_set_f32_tail": (C.c_int, [_vp, C.c_int]),
    "gomel_set_gl_guard": (C.c_int, [_vp, C.c_float, _fp]),
    "gomel_last_gl_guard": (C.c_int, [_vp, _ip, _ip, _fp, _fp, C.c_int]),
    "gomel_frames": (C.c_int, [_cp, C.c_long, _lp, _lp, _lp]),
    "gomel_ola_len": (C.c_long, [_cp, C.c_long]),
    "gomel_set_mel_tables": (C.c_int, [_vp, _cp, _ip, _ip, _dp, _ip, _ip, _dp]),
    "gomel_to_mel": (C.c_int, [_vp, _cp, _dp, C.c_long, _dp]),
    "gomel_from_mel": (C.c_int, [_vp, _cp, _dp, C.c_long, _dp, C.c_ulonglong, _dp]),
    "gomel_to_phase": (C.c_int, [_vp, _cp, _dp, C.c_long, _dp]),
    "gomel_from_phase": (C.c_int, [_vp, _cp, _dp, C.c_long, _dp]),
    "gomel_image": (C.c_int, [_vp, _dp, C.c_long, C.c_int, C.POINTER(C.c_ushort), _dp]),
    "gomel_quantise": (C.c_int, [_vp, _dp, C.c_long, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ushort), _dp]),
    "gomel_dequantise": (C.c_int, [_vp, C.POINTER(C.c_ushort), C.c_long, C.c_int, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_int, _dp]),
    "gomel_dev_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "gomel_dev_free": (C.c_int, [_vp, _vp]),
    "gomel_host_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "gomel_host_free": (C.c_int, [_vp, _vp]),
    "gomel_copy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "gomel_copy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "gomel_sync": (C.c_int, [_vp]),
    "gomel_timer_start": (C.c_int, [_vp]),
    "gomel_timer_stop": (C.c_int, [_vp, _fp]),
    "gomel_last_hot_kernel_ms": (C.c_int, [_vp, _fp, _ip]),
    "gomel_last_lead_kernel_ms": (C.c_int, [_vp, _fp, _ip]),
    "gomel_to_mel_dev": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, C.c_long, C.c_long, _vp]),
    "gomel_to_phase_dev": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, C.c_long, C.c_long, _vp]),
    "gomel_stft_dev": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, C.c_long, C.c_long, _vp]),
    "gomel_from_mel_dev": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, _vp, C.c_ulonglong, C.c_long, _vp]),
    "gomel_from_phase_dev": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, C.c_long, _vp]),
    "gomel_ts_create": (C.c_int, [_vp, _cp, C.c_long, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gomel_ts_create2": (C.c_int, [_vp, _cp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gomel_ts_destroy": (None, [_vp]),
    "gomel_ts_range": (C.c_int, [_vp, _lp, _lp, _lp, _lp]),
    "gomel_ts_load": (C.c_int, [_vp, _vp, _vp, C.c_ulonglong]),
    "gomel_ts_iterate": (C.c_int, [_vp, C.c_int, C.c_int]),
    "gomel_ts_halo_ptrs": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "gomel_ts_halo_elem_bytes": (C.c_int, [_vp, C.c_int]),
    "gomel_ts_lead_iters": (C.c_int, [_vp]),
    "gomel_ts_comm_stream": (_vp, [_vp]),
    "gomel_ts_comm_begin": (C.c_int, [_vp, C.c_int]),
    "gomel_ts_comm_end": (C.c_int, [_vp, C.c_int]),
    "gomel_ts_finish": (C.c_int, [_vp, C.c_int, _vp]),
    "gomel_ts_sync": (C.c_int, [_vp]),
    "gomel_nccl_unique_id": (C.c_int, [_vp, C.c_char_p]),
    "gomel_ts_nccl_init": (C.c_int, [_vp, C.c_char_p]),
    "gomel_ts_run_nccl": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "gomel_ts_phase_istft": (C.c_int, [_vp, _vp]),
    "gomel_ts_phase_halo_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "gomel_ts_phase_finish": (C.c_int, [_vp, _vp]),
    "gomel_ts_phase_run_nccl": (C.c_int, [_vp, _vp, _vp]),
    "gomel_copy_d2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "gomel_from_mel_batch_host": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, _vp, C.c_ulonglong, _vp, C.c_int]),
    "gomel_from_mel_batch_host_pcm16": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, _vp, C.c_ulonglong, _vp, C.c_int]),
    "gomel_to_mel_batch_host": (C.c_int, [_vp, _cp, _vp, C.c_int, C.c_long, _vp, C.c_int]),
}

_lib = None
_lock = threading.RLock()         # re-entrant: default_context() creates a Context (-> load()) under it


def load():
    """dlopen libgomelcuda.so and bind every declared symbol (no GPU needed for this step)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C gomel_b200/csrc`.  gomel_b200 has no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)          # AttributeError if the library does not export it
                fn.restype, fn.argtypes = res, args
            _lib = L
    return _lib


FLAG_F64 = 1          # every Griffin-Lim iteration in float64 on the fused kernel (GOMEL_FLAG_F64)
FLAG_F64_REF = 2      # round-1 strict float64 path, test instrument (GOMEL_FLAG_F64_REF)


def make_config(n_fft=4096, hop=1280, n_mels=192, n_freqs=768, gl_iters=2, tune_mul=1.0, tune_add=0.0,
                volume_boost=0.0, flags=0, mel_fmin=0.0, mel_fmax=0.0):
    """mel_fmin / mel_fmax select the registered filterbank tables (Context.set_mel_tables); both 0 = the most
    recently registered tables for (n_fft, n_mels)."""
    return Config(n_fft, hop, n_mels, n_freqs, gl_iters, tune_mul, tune_add, volume_boost, flags, mel_fmin, mel_fmax)


def frames(cfg, n_samples):
    """(n_padded, n_frames, ola_len) -- pure host arithmetic, works without a GPU."""
    a, b, c = C.c_long(), C.c_long(), C.c_long()
    rc = load().gomel_frames(C.byref(cfg), n_samples, C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise GomelError(rc, "gomel_frames: bad length")
    return a.value, b.value, c.value


# ---- mel filterbank tables, computed by the CALLER's math library ------------------------
def _hz_to_mel(v):
    return 1127.0 * math.log(1.0 + (v / 700.0))        # mel/impl.go:304-308


def _mel_to_hz(v):
    return 700.0 * (math.exp(v / 1127.0) - 1.0)        # mel/impl.go:298-302


def mel_tables(filtersize, mels, fmin, fmax):
    """The (int(inlo), int(inhi), modlo) triples of domel (mel/impl.go:313-323) and undomel
    (mel/impl.go:350-360)."""
    flo, fhi = np.empty(mels, np.int32), np.empty(mels, np.int32)
    fmod = np.empty(mels, np.float64)
    melbin = _hz_to_mel(fmax) / float(mels)
    for i in range(mels):
        vallo = float(filtersize) * (fmin + _mel_to_hz(melbin * float(i))) / (fmax + fmin)
        valhi = float(filtersize) * (fmin + _mel_to_hz(melbin * float(i + 1))) / (fmax + fmin)
        modlo, inlo = math.modf(vallo)
        inhi = math.floor(valhi)
        if inlo < 0:
            inlo, modlo, inhi = 0, 0.0, 0
        flo[i], fhi[i], fmod[i] = int(inlo), int(inhi), modlo
    ilo, ihi = np.empty(filtersize, np.int32), np.empty(filtersize, np.int32)
    imod = np.empty(filtersize, np.float64)
    filterbin = _hz_to_mel(fmax) / float(mels)
    for i in range(filtersize):
        vallo = _hz_to_mel((float(i) * (fmax + fmin) / float(filtersize)) - fmin) / filterbin
        valhi = _hz_to_mel((float(i + 1) * (fmax + fmin) / float(filtersize)) - fmin) / filterbin
        modlo, inlo = math.modf(vallo)
        inhi = math.floor(valhi)
        if inlo < 0:
            inlo, modlo, inhi = 0, 0.0, 0
        ilo[i], ihi[i], imod[i] = int(inlo), int(inhi), modlo
    return flo, fhi, fmod, ilo, ihi, imod


class Context:
    """Owns one gomel_ctx (one device, one stream)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.gomel_ctx_create(device, C.byref(h))
        if rc:
            raise GomelError(rc, "gomel_ctx_create failed (no CUDA device? gomel_b200 has no CPU fallback)")
        self.h = h
        self.device = device
        self._tables_key = None          # last set registered by set_mel_tables (the "most recent" one)
        self._registered = set()         # keys registered so far (bounded; the library keeps 64 sets, LRU)
        self._lock = threading.RLock()

    def close(self):
        if getattr(self, "h", None):
            self.lib.gomel_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc:
            raise GomelError(rc, self.lib.gomel_last_error(self.h).decode())

    def _register(self, key):
        n_fft, n_mels, fmin, fmax = key
        kcfg = make_config(n_fft=n_fft, hop=256 if n_fft == 2048 else 1280, n_mels=n_mels, mel_fmin=fmin, mel_fmax=fmax)
        flo, fhi, fmod, ilo, ihi, imod = mel_tables(n_fft // 2, n_mels, fmin, fmax)
        self.check(self.lib.gomel_set_mel_tables(
            self.h, C.byref(kcfg), flo.ctypes.data_as(_ip), fhi.ctypes.data_as(_ip), fmod.ctypes.data_as(_dp),
            ilo.ctypes.data_as(_ip), ihi.ctypes.data_as(_ip), imod.ctypes.data_as(_dp)))
        if len(self._registered) >= 48:
            self._registered.clear()
        self._registered.add(key)

    def set_mel_tables(self, cfg, fmin, fmax):
        """Registers the filterbank tables of (cfg.n_fft, cfg.n_mels, fmin, fmax) and makes them the context's most
        recent set -- the one every config with mel_fmin = mel_fmax = 0 uses.  For one configuration at a time;
        callers that share a context across threads with different configurations use use_mel_tables()."""
        key = (cfg.n_fft, cfg.n_mels, float(fmin), float(fmax))
        with self._lock:
            if self._tables_key == key:
                return
            self._register(key)
            self._tables_key = key

    def use_mel_tables(self, cfg, fmin, fmax):
        """Stamps the table key (MelFmin, MelFmax) into `cfg` and registers that set once.  The mel entry points
        look the set up by the key of the cfg they are given, so threads with different Mel configurations can
        share this context: there is no context-wide "current tables" state on this path."""
        key = (cfg.n_fft, cfg.n_mels, float(fmin), float(fmax))
        cfg.mel_fmin, cfg.mel_fmax = key[2], key[3]
        with self._lock:
            self._tables_key = None          # a keyed use may reorder the library's most-recent list
            if key not in self._registered:
                self._register(key)

    def _mel_call(self, cfg, call):
        """Runs a mel entry point; if the library has evicted the cfg's table set meanwhile, registers it again."""
        rc = call()
        if rc == E_STATE and (cfg.mel_fmin != 0 or cfg.mel_fmax != 0):
            with self._lock:
                self._register((cfg.n_fft, cfg.n_mels, cfg.mel_fmin, cfg.mel_fmax))
            rc = call()
        self.check(rc)

    def launch_count(self):
        return int(self.lib.gomel_launch_count(self.h))

    def set_tile_frames(self, t):
        self.check(self.lib.gomel_set_tile_frames(self.h, int(t)))

    def set_lead_f64(self, k):
        """minimum number of float64 lead iterations of Griffin-Lim (default 16); returns the previous value"""
        rc = self.lib.gomel_set_lead_f64(self.h, int(k))
        if rc < 0:
            self.check(rc)
        return rc

    def set_f32_tail(self, n):
        """maximum number of trailing float32 iterations (default 16; < 0 unlimited); returns the previous value"""
        rc = self.lib.gomel_set_f32_tail(self.h, int(n))
        if rc < 0:
            self.check(rc)
        return -1 if rc == 0x7fffffff else rc

    def set_gl_guard(self, threshold):
        """leverage threshold of the float32 tail's singular-bin guard (default 5e4, 0 disables); returns the previous value"""
        prev = C.c_float()
        self.check(self.lib.gomel_set_gl_guard(self.h, float(threshold), C.byref(prev)))
        return prev.value

    def last_gl_guard(self, cap=0):
        """-> (clips the guard saw, clips re-run in float64, largest leverage, leverage of the first `cap` clips)"""
        n, r, mx = C.c_int(), C.c_int(), C.c_float()
        lev = np.zeros(max(cap, 1), np.float32)
        self.check(self.lib.gomel_last_gl_guard(self.h, C.byref(n), C.byref(r), C.byref(mx), lev.ctypes.data_as(_fp), int(cap)))
        return n.value, r.value, mx.value, lev[:min(cap, n.value)]

    def set_gl_precision(self, lead, tail):
        """-> (previous lead, previous tail)"""
        return self.set_lead_f64(lead), self.set_f32_tail(tail)

    # ---- host-buffer API -------------------------------------------------------------
    def to_mel(self, cfg, wav):
        wav = np.ascontiguousarray(wav, np.float64)
        if len(wav) == 0:            # the reference's pad() grows an empty buffer to 15*Window-1 zeros; one zero pads the same
            wav = np.zeros(1)
        _, fr, _ = frames(cfg, len(wav))
        out = np.empty((fr * cfg.n_mels, 2), np.float64)
        self._mel_call(cfg, lambda: self.lib.gomel_to_mel(self.h, C.byref(cfg), wav.ctypes.data_as(_dp), len(wav),
                                                          out.ctypes.data_as(_dp)))
        return out

    def from_mel(self, cfg, mel, init=None, seed=0):
        mel = np.ascontiguousarray(mel, np.float64).reshape(-1, 2)
        if cfg.n_mels <= 0 or len(mel) == 0 or len(mel) % cfg.n_mels:
            # the Go reference strides by NumMels without a length check and panics (mel/impl.go:366-372)
            raise GomelError(E_ARG, "len(mel) is not a positive multiple of NumMels")
        fr = len(mel) // cfg.n_mels
        ola = cfg.n_fft + (fr - 1) * cfg.hop
        ip = None
        if init is not None:
            init = np.ascontiguousarray(init, np.float64)
            if len(init) != ola:
                raise GomelError(E_ARG, "init signal length != ola_len")
            ip = init.ctypes.data_as(_dp)
        out = np.empty(ola, np.float64)
        self._mel_call(cfg, lambda: self.lib.gomel_from_mel(self.h, C.byref(cfg), mel.ctypes.data_as(_dp), fr, ip, seed,
                                                            out.ctypes.data_as(_dp)))
        return out

    def to_phase(self, cfg, wav):
        wav = np.ascontiguousarray(wav, np.float64)
        if len(wav) == 0:
            wav = np.zeros(1)
        _, fr, _ = frames(cfg, len(wav))
        out = np.empty((fr * cfg.n_freqs, 2), np.float64)
        self.check(self.lib.gomel_to_phase(self.h, C.byref(cfg), wav.ctypes.data_as(_dp), len(wav),
                                           out.ctypes.data_as(_dp)))
        return out

    def from_phase(self, cfg, spec):
        spec = np.ascontiguousarray(spec, np.float64).reshape(-1, 2)
        if cfg.n_freqs <= 0 or len(spec) == 0 or len(spec) % cfg.n_freqs:
            raise GomelError(E_ARG, "len(spec) is not a positive multiple of NumFreqs")
        fr = len(spec) // cfg.n_freqs
        out = np.empty(cfg.n_fft + (fr - 1) * cfg.hop, np.float64)
        self.check(self.lib.gomel_from_phase(self.h, C.byref(cfg), spec.ctypes.data_as(_dp), fr,
                                             out.ctypes.data_as(_dp)))
        return out

    def image(self, buf, mels):
        buf = np.ascontiguousarray(buf, np.float64).reshape(-1, 2)
        n = (len(buf) // mels) * mels
        out = np.empty(n, np.uint16)
        mm = np.empty(4, np.float64)
        self.check(self.lib.gomel_image(self.h, buf.ctypes.data_as(_dp), len(buf), mels,
                                        out.ctypes.data_as(C.POINTER(C.c_ushort)), mm.ctypes.data_as(_dp)))
        return out

    def quantise(self, buf, mels, flags, ihs_passes=0):
        """-> (rgb uint16 (n,3) in buffer order, [max0,max1,min0,min1])"""
        buf = np.ascontiguousarray(buf, np.float64).reshape(-1, 2)
        n = (len(buf) // mels) * mels
        out = np.empty((n, 3), np.uint16)
        mm = np.empty(4, np.float64)
        self.check(self.lib.gomel_quantise(self.h, buf.ctypes.data_as(_dp), len(buf), mels, flags, ihs_passes,
                                           out.ctypes.data_as(C.POINTER(C.c_ushort)), mm.ctypes.data_as(_dp)))
        return out, mm

    def dequantise(self, rg, hdr, max0, max1, min0, min1, ihs_passes=0):
        rg = np.ascontiguousarray(rg, np.uint16).reshape(-1, 2)
        out = np.empty((len(rg), 2), np.float64)
        self.check(self.lib.gomel_dequantise(self.h, rg.ctypes.data_as(C.POINTER(C.c_ushort)), len(rg), int(hdr),
                                             max0, max1, min0, min1, ihs_passes, out.ctypes.data_as(_dp)))
        return out

    # ---- device-resident API ---------------------------------------------------------
    def dev_malloc(self, nbytes):
        p = _vp()
        self.check(self.lib.gomel_dev_malloc(self.h, nbytes, C.byref(p)))
        return p

    def dev_free(self, p):
        self.check(self.lib.gomel_dev_free(self.h, p))

    def host_malloc(self, nbytes):
        p = _vp()
        self.check(self.lib.gomel_host_malloc(self.h, nbytes, C.byref(p)))
        return p

    def host_free(self, p):
        self.check(self.lib.gomel_host_free(self.h, p))

    def pinned_array(self, shape, dtype=np.float32):
        """numpy view of a pinned host buffer; keep the returned owner pointer to free it."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.host_malloc(n)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), p

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.gomel_copy_h2d(self.h, dptr, arr.ctypes.data_as(_vp), arr.nbytes))
        self.sync()

    def d2h(self, arr, dptr):
        assert arr.flags["C_CONTIGUOUS"]
        self.check(self.lib.gomel_copy_d2h(self.h, arr.ctypes.data_as(_vp), dptr, arr.nbytes))
        self.sync()

    def sync(self):
        self.check(self.lib.gomel_sync(self.h))

    def timer_start(self):
        self.check(self.lib.gomel_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.lib.gomel_timer_stop(self.h, C.byref(ms)))
        return ms.value


def _hot(self):
    ms, n = C.c_float(), C.c_int()
    self.check(self.lib.gomel_last_hot_kernel_ms(self.h, C.byref(ms), C.byref(n)))
    return ms.value, n.value


Context.last_hot_kernel_ms = _hot


def _lead(self):
    ms, n = C.c_float(), C.c_int()
    self.check(self.lib.gomel_last_lead_kernel_ms(self.h, C.byref(ms), C.byref(n)))
    return ms.value, n.value


Context.last_lead_kernel_ms = _lead

_default = {}


def default_context(device=0):
    """Process-wide context per device (what the drop-in classes use)."""
    with _lock:
        ctx = _default.get(device)
        if ctx is None:
            ctx = _default[device] = Context(device)
    return ctx
