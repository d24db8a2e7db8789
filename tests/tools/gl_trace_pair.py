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

seed = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 32
seconds = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
n = int(round(seconds * 44100))
if sys.argv[1] == "impulses":                       # the sweep tool's impulse clip
    rng = np.random.default_rng(8)
    wav = np.zeros(n)
    wav[rng.integers(0, n, max(4, n // 11000))] = rng.uniform(-1, 1, max(4, n // 11000))
elif sys.argv[1] == "white_noise":
    wav = np.random.default_rng(77).uniform(-1, 1, n)
else:
    wav = synth_clip(int(sys.argv[1]), seconds)
mel = O.to_mel(O.config(), wav)
frames = len(mel) // 192
init = np.random.default_rng(seed).random(4096 + (frames - 1) * 1280)
ref = run(mel, init, iters, True)
print("lead -> rel-L2 of the final signal vs all-float64")
for lead in list(range(0, iters + 1, 2)) + [iters - 1]:
    print(f"  lead {lead:3d}: {rel_l2(run(mel, init, iters, False, lead=lead), ref):.3e}", flush=True)
# where along the run does the default split diverge: compare truncated runs (k iterations) of the two modes
print("k -> rel-L2 after k iterations, lead 16 vs all-float64")
for k in range(16, iters + 1):
    a = run(mel, init, k, False, lead=16); b = run(mel, init, k, True)
    print(f"  k {k:3d}: {rel_l2(a, b):.3e}", flush=True)
