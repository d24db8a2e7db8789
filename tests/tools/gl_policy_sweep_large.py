#!/usr/bin/env python3
"""Default precision policy against the all-float64 kernel on a larger population than the test-suite sweep:
8 synthetic clips + white noise + silence + impulses, 10 s each, 32 start signals each at GL-32 (352 pairs) and 12 at
GL-100 (132 pairs).  Writes gpurun_out/gl_policy_sweep_large.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from gl_modes_check import O, rel_l2, run, synth_clip      # noqa: E402

kinds = [("clip%d" % c, synth_clip(c, 10.0)) for c in range(8)]
rng = np.random.default_rng(8)
kinds.append(("white_noise", np.random.default_rng(77).uniform(-1, 1, 441000)))
kinds.append(("silence", np.zeros(441000)))
imp = np.zeros(441000)
imp[rng.integers(0, 441000, 40)] = rng.uniform(-1, 1, 40)
kinds.append(("impulses", imp))
rows = []
for iters, n_seeds in ((32, 32), (100, 12)):
    for name, wav in kinds:
        mel = O.to_mel(O.config(), wav)
        errs = []
        for s in range(n_seeds):
            init = np.random.default_rng(1000 + s).random(440576)
            e = rel_l2(run(mel, init, iters, False), run(mel, init, iters, True))
            errs.append(e)
            rows.append({"clip": name, "iters": iters, "seed": 1000 + s, "rel_l2": e})
        print(f"GL-{iters} {name}: max {max(errs):.2e} median {np.median(errs):.2e}", flush=True)
summ = {}
for iters in (32, 100):
    v = np.array([r["rel_l2"] for r in rows if r["iters"] == iters])
    summ[iters] = {"pairs": int(len(v)), "max": float(v.max()), "median": float(np.median(v)), "p99": float(np.quantile(v, 0.99)),
                   "pass_frac_1e-4": float(np.mean(v <= 1e-4))}
print(json.dumps(summ, indent=1))
json.dump({"rows": rows, "summary": summ}, open(os.path.join(ROOT, "gpurun_out", "gl_policy_sweep_large.json"), "w"), indent=1)
