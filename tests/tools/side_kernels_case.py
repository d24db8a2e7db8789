#!/usr/bin/env python3
"""runs the batched ToMel / ToPhase / FromPhase kernels once on 256 x 10 s clips (ncu target for the side kernels)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip
from gomel_b200 import _lib
ctx = _lib.Context(0)
cfg = _lib.make_config()
ctx.set_mel_tables(cfg, 0.0, 16000.0)
clips, n = 256, 441000
npad, frames, ola = _lib.frames(cfg, n)
stride = (npad + 3) & ~3
wav = np.zeros((clips, stride), np.float32)
base = np.stack([synth_clip(500 + c, 10.0) for c in range(8)]).astype(np.float32)
for c in range(clips):
    wav[c, :n] = base[c % 8]
d_sig = ctx.dev_malloc(wav.nbytes); ctx.h2d(d_sig, wav)
d_mel = ctx.dev_malloc(clips * frames * 192 * 2 * 4)
d_ph = ctx.dev_malloc(clips * frames * 768 * 2 * 4)
d_wav = ctx.dev_malloc(clips * ola * 4)
for _ in range(3):
    ctx.check(ctx.lib.gomel_to_mel_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_mel))
    ctx.check(ctx.lib.gomel_to_phase_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_ph))
    ctx.check(ctx.lib.gomel_from_phase_dev(ctx.h, C.byref(cfg), d_ph, clips, frames, ola, d_wav))
ctx.sync()
print("SIDE_DONE")
