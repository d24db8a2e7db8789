#!/usr/bin/env python3
"""GL-32 on the large population (11 clips x 32 start signals): rel-L2 vs the all-float64 kernel for several numbers of
float64 lead iterations.  Writes gpurun_out/gl_lead_sweep_large.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from gl_modes_check import O, rel_l2, run, synth_clip      # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 32
leads = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4,6,8,12,16,20,24".split(","))]
n_seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 32
SEED0 = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
kinds = [("clip%d" % c, synth_clip(c, 10.0)) for c in range(8)]
rng = np.random.default_rng(8)
kinds.append(("white_noise", np.random.default_rng(77).uniform(-1, 1, 441000)))
kinds.append(("silence", np.zeros(441000)))
imp = np.zeros(441000)
imp[rng.integers(0, 441000, 40)] = rng.uniform(-1, 1, 40)
kinds.append(("impulses", imp))
rows = []
for name, wav in kinds:
    mel = O.to_mel(O.config(), wav)
    for s in range(n_seeds):
        init = np.random.default_rng(SEED0 + s).random(440576)
        f64 = run(mel, init, iters, True)
        row = {"clip": name, "seed": SEED0 + s}
        for lead in leads:
            row["lead%d" % lead] = rel_l2(run(mel, init, iters, False, lead=lead), f64)
        rows.append(row)
    sub = [r for r in rows if r["clip"] == name]
    print(name, " ".join("lead%d %.1e" % (l, max(r["lead%d" % l] for r in sub)) for l in leads), flush=True)
summ = {}
for lead in leads:
    v = np.array([r["lead%d" % lead] for r in rows])
    summ["lead%d" % lead] = {"max": float(v.max()), "p99": float(np.quantile(v, 0.99)), "median": float(np.median(v)),
                             "pass_frac_1e-4": float(np.mean(v <= 1e-4)), "n_over_3e-5": int(np.sum(v > 3e-5))}
print(json.dumps(summ, indent=1))
json.dump({"iters": iters, "rows": rows, "summary": summ}, open(os.path.join(ROOT, "gpurun_out", f"gl_lead_sweep_large_{iters}_{SEED0}.json"), "w"), indent=1)
