#!/usr/bin/env python3
"""how does the fp32-vs-float64 Griffin-Lim deviation depend on the start signal? (GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import rel_l2, synth_clip
from oracle import oracle as O, oracle_np as ONP
from gomel_b200 import NewMel

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
wav = synth_clip(0, secs)
ocfg = O.config(gl_iters=32)
mel = O.to_mel(ocfg, wav)
frames = len(mel) // 192
ola = 4096 + (frames - 1) * 1280
M = ONP.gl_magnitudes(mel)
for seed in (5, 7, 11, 13):
    init = np.random.default_rng(seed).random(ola)
    errs = []
    for iters in (2, 8, 16, 32):
        m = NewMel(); m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut, m.GriffinLimIterations = 192, 0, 16000, 1280, 4096, iters
        m.InitSignal = init
        got = m.FromMel(mel.copy())
        ref = ONP.griffin_lim(M, init, iters)
        errs.append(rel_l2(got, ref))
    # float64 sensitivity: perturb the start signal by 1e-7 relative (what fp32 storage does) and rerun in float64
    pert = ONP.griffin_lim(M, init.astype(np.float32).astype(np.float64), 32)
    base = ONP.griffin_lim(M, init, 32)
    print(f"seed {seed:2d}: fp32-vs-f64 after 2/8/16/32 it = " + " ".join(f"{e:.2e}" for e in errs) +
          f" | float64 run with fp32-rounded start signal vs exact: {rel_l2(pert, base):.2e}", flush=True)
