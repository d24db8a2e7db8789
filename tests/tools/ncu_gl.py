#!/usr/bin/env python3
"""one batched FromMel (default precision policy) for ncu: `python tests/tools/ncu_gl.py [clips] [iters]`"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip                      # noqa: E402
from gomel_b200 import _lib                      # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = _lib.Context(0)
cfg = _lib.make_config(gl_iters=iters)
ctx.set_mel_tables(cfg, 0.0, 16000.0)
n = 441000
_, frames, ola = _lib.frames(cfg, n)
nb = 8
wav = np.stack([synth_clip(c, 10.0) for c in range(nb)]).astype(np.float32)
base = np.empty((nb, frames * 192 * 2), np.float32)
ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), nb, n,
                                          base.ctypes.data_as(C.c_void_p), 8))
mel = np.concatenate([base] * (clips // nb))
d_mel = ctx.dev_malloc(mel.nbytes)
d_out = ctx.dev_malloc(clips * ola * 4)
ctx.h2d(d_mel, mel)
for rep in range(2):
    ctx.timer_start()
    ctx.check(ctx.lib.gomel_from_mel_dev(ctx.h, C.byref(cfg), d_mel, clips, frames, None, 7, ola, d_out))
    ms = ctx.timer_stop()
lms, ln = ctx.last_lead_kernel_ms()
hms, hn = ctx.last_hot_kernel_ms()
print(f"{clips} clips x {frames} frames, {iters} it: {ms:.3f} ms; float64 lead {ln} it {lms:.3f} ms "
      f"({clips * frames * ln / (lms / 1e3) if ln else 0:.3e} frame-it/s); float32 {hn} it {hms:.3f} ms "
      f"({clips * frames * hn / (hms / 1e3) if hn else 0:.3e} frame-it/s)")
