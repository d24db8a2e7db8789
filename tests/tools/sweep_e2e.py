#!/usr/bin/env python3
"""e2e pipeline sweep in one process: chunk size x smallest tapered chunk (x GOMEL_GL_STREAMS is fixed per process).
    python tests/tools/sweep_e2e.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip
from gomel_b200 import _lib

ctx = _lib.Context(0)
cfg = _lib.make_config(gl_iters=32)
ctx.set_mel_tables(cfg, 0.0, 16000.0)
clips, n = 1024, 441000
npad, frames, ola = _lib.frames(cfg, n)
nb = 16
wav = np.stack([synth_clip(c, 10.0) for c in range(nb)]).astype(np.float32)
base = np.empty((nb, frames * 192 * 2), np.float32)
ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), nb, n,
                                          base.ctypes.data_as(C.c_void_p), 16))
h_mel, o1 = ctx.pinned_array((clips, frames * 192 * 2), np.float32)
for c in range(clips):
    h_mel[c] = base[c % nb]
h_out, o2 = ctx.pinned_array((clips, ola), np.float32)
audio = clips * frames * 1280 / 44100.0


def run(chunk):
    ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(cfg), h_mel.ctypes.data_as(C.c_void_p), clips, frames,
                                                None, 1, h_out.ctypes.data_as(C.c_void_p), chunk))


combos = [(c, f, 0) for c in (256, 384, 512, 768, 1024) for f in (32, 64, 128, 256) if f <= c // 2]
if len(sys.argv) > 1 and sys.argv[1] == "first":                 # sweep the first chunk instead
    combos = [(512, 64, f1) for f1 in (32, 64, 96, 128, 192, 256)] + [(384, 64, 64), (640, 64, 128)]
for chunk, floor_, first in combos:
    os.environ["GOMEL_CHUNK_TAPER_MIN"] = str(floor_)
    if first:
        os.environ["GOMEL_CHUNK_FIRST"] = str(first)
    run(chunk); run(chunk)
    t0 = time.perf_counter()
    for _ in range(5):
        run(chunk)
    ms = (time.perf_counter() - t0) * 1e3 / 5
    print(f"chunk {chunk:5d} taper_min {floor_:4d} first {first:4d}: {ms:7.2f} ms  {audio / ms:7.1f} k audio-s/s", flush=True)
