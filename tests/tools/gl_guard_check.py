#!/usr/bin/env python3
"""The float32 tail's singular-bin guard on single (clip, start signal) pairs (GPU box):
    python tests/tools/gl_guard_check.py <clip> <seed0> <n> [iters] [seconds]
prints, per pair, the error against the all-float64 kernel with the guard off and on, and the recorded leverage."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from gl_modes_check import O, rel_l2, run, synth_clip      # noqa: E402
from gomel_b200 import _lib                                # noqa: E402

clip = int(sys.argv[1]); seed0 = int(sys.argv[2]); n = int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 32
seconds = float(sys.argv[5]) if len(sys.argv) > 5 else 10.0
ctx = _lib.default_context(0)
wav = synth_clip(clip, seconds)
mel = O.to_mel(O.config(), wav)
frames = len(mel) // 192
for s in range(seed0, seed0 + n):
    init = np.random.default_rng(s).random(4096 + (frames - 1) * 1280)
    ref = run(mel, init, iters, True)
    prev = ctx.set_gl_guard(1e30)                 # records, never re-runs
    off = run(mel, init, iters, False)
    _, _, lev, _ = ctx.last_gl_guard()
    ctx.set_gl_guard(prev)
    on = run(mel, init, iters, False)
    nc, nr, lev2, _ = ctx.last_gl_guard()
    print(f"clip {clip} seed {s}: guard off {rel_l2(off, ref):.2e} (leverage {lev:.3e}) | guard on {rel_l2(on, ref):.2e} "
          f"(re-run {nr}/{nc}, leverage {lev2:.3e})", flush=True)
