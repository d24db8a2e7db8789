#!/usr/bin/env python3
"""One (clip, start signal) pair under every split of float64 lead / float32 tail iterations, and the iteration at which
a float32 tail first leaves the float64 trajectory (GPU box):
    python tests/tools/gl_trace_pair.py <clip> <seed> [iters] [seconds]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from gl_modes_check import O, rel_l2, run, synth_clip      # noqa: E402

clip = int(sys.argv[1]); seed = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 32
seconds = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
wav = synth_clip(clip, seconds)
mel = O.to_mel(O.config(), wav)
frames = len(mel) // 192
init = np.random.default_rng(seed).random(4096 + (frames - 1) * 1280)
ref = run(mel, init, iters, True)
print("lead -> rel-L2 of the final signal vs all-float64")
for lead in list(range(0, iters + 1, 2)):
    print(f"  lead {lead:3d}: {rel_l2(run(mel, init, iters, False, lead=lead), ref):.3e}", flush=True)
# where along the run does the default split diverge: compare truncated runs (k iterations) of the two modes
print("k -> rel-L2 after k iterations, lead 16 vs all-float64")
for k in range(16, iters + 1, 2):
    a = run(mel, init, k, False, lead=16); b = run(mel, init, k, True)
    print(f"  k {k:3d}: {rel_l2(a, b):.3e}", flush=True)
