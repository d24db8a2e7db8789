#!/usr/bin/env python3
"""times batched ToMel (256 x 10 s clips, device-resident, CUDA events) for each library given:
    python tests/tools/time_to_mel.py gomel_b200/libgomelcuda.so gomel_b200/ab/*.so"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        env = dict(os.environ, GOMEL_CUDA_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, __file__, "--child"], capture_output=True, text=True, env=env)
        print(f"{os.path.basename(lib):32s} {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
    sys.exit(0)

import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip, rel_l2
from gomel_b200 import _lib
from oracle import oracle
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
call = lambda: ctx.check(ctx.lib.gomel_to_mel_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_mel))
for _ in range(5):
    call()
ctx.sync()
best = 1e9
for _ in range(5):
    ctx.timer_start()
    for _ in range(20):
        call()
    best = min(best, ctx.timer_stop() / 20)
got = np.empty((clips, frames * 192, 2), np.float32)
ctx.d2h(got, d_mel)
ref = oracle.to_mel(oracle.config(), base[3].astype(np.float64))
print(f"to_mel {best:.4f} ms  {clips * frames / best / 1e3:.1f} Mframes/s  rel-L2 {rel_l2(np.exp(got[3]), np.exp(ref)):.2e}")
