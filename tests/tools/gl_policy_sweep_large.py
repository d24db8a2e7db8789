#!/usr/bin/env python3
"""Default precision policy against the all-float64 kernel on large populations (GPU box):
    python tests/tools/gl_policy_sweep_large.py <iters> <n_seeds> <seed0> [seconds]
11 clips (8 synthetic, white noise, silence, impulses) x n_seeds start signals; writes
gpurun_out/gl_policy_sweep_<iters>_<seed0>_<seconds>.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
from gl_modes_check import O, rel_l2, run, synth_clip      # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 32
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
seconds = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
n = int(round(seconds * 44100))
kinds = [("clip%d" % c, synth_clip(c, seconds)) for c in range(8)]
rng = np.random.default_rng(8)
kinds.append(("white_noise", np.random.default_rng(77).uniform(-1, 1, n)))
kinds.append(("silence", np.zeros(n)))
imp = np.zeros(n)
imp[rng.integers(0, n, max(4, n // 11000))] = rng.uniform(-1, 1, max(4, n // 11000))
kinds.append(("impulses", imp))
rows = []
from gomel_b200 import _lib                                  # noqa: E402
ctx = _lib.default_context(0)
guard_mode = os.environ.get("GL_SWEEP_GUARD", "default")     # "record": statistic only, nothing re-run; "off"; "default"
if guard_mode == "record":
    ctx.set_gl_guard(1e30)
elif guard_mode == "off":
    ctx.set_gl_guard(0.0)
if "GL_SWEEP_LEAD" in os.environ:                             # a split other than the default: lead L, the rest float32
    ctx.set_gl_precision(int(os.environ["GL_SWEEP_LEAD"]), -1)
if "GL_SWEEP_THR" in os.environ:
    ctx.set_gl_guard(float(os.environ["GL_SWEEP_THR"]))
for name, wav in kinds:
    mel = O.to_mel(O.config(), wav)
    frames = len(mel) // 192
    ola = 4096 + (frames - 1) * 1280
    errs = []
    for s in range(n_seeds):
        init = np.random.default_rng(seed0 + s).random(ola)
        hyb = run(mel, init, iters, False)
        _, n_rerun, lev, _ = ctx.last_gl_guard()
        e = rel_l2(hyb, run(mel, init, iters, True))
        errs.append(e)
        rows.append({"clip": name, "iters": iters, "seed": seed0 + s, "rel_l2": e, "leverage": lev, "rerun": n_rerun})
    print(f"GL-{iters} {seconds:g}s {name}: max {max(errs):.2e} median {np.median(errs):.2e}", flush=True)
v = np.array([r["rel_l2"] for r in rows])
summ = {"iters": iters, "seconds": seconds, "pairs": int(len(v)), "max": float(v.max()), "median": float(np.median(v)),
        "p99": float(np.quantile(v, 0.99)), "pass_frac_1e-4": float(np.mean(v <= 1e-4)), "misses": int(np.sum(v > 1e-4)),
        "guard": guard_mode, "rerun_frac": float(np.mean([r["rerun"] for r in rows]))}
print(json.dumps(summ))
json.dump({"rows": rows, "summary": summ}, open(os.path.join(ROOT, "gpurun_out", f"gl_policy_sweep_{iters}_{seed0}_{seconds:g}_{guard_mode}" + ("_lead" + os.environ["GL_SWEEP_LEAD"] if "GL_SWEEP_LEAD" in os.environ else "") + ".json"), "w"), indent=1)
