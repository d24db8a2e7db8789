#!/usr/bin/env python3
"""throughput of the strict float64 Griffin-Lim path vs the float32 path on one long clip (GPU box)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gomel_b200 import _lib
ctx = _lib.Context(0)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
frames = int(secs * 44100 / 1280)
mel = np.random.default_rng(0).uniform(-8, 2, (frames * 192, 2))
init = np.random.default_rng(1).random(4096 + (frames - 1) * 1280)
for name, flags in (("float32", 0), ("strict float64", _lib.FLAG_F64)):
    cfg = _lib.make_config(gl_iters=32, flags=flags)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    ctx.from_mel(cfg, mel, init=init)
    t0 = time.perf_counter(); ctx.from_mel(cfg, mel, init=init); dt = time.perf_counter() - t0
    print(f"{name:15s}: {frames} frames x 32 it in {dt*1e3:8.1f} ms -> {frames*1280/44100/dt:9.0f} audio-s/s, {frames*32/dt:.3e} frame-it/s (host API incl. float64 copies)")
