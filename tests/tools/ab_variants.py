#!/usr/bin/env python3
"""A/B several builds of the library in one GPU call:
    python tests/tools/ab_variants.py gomel_b200/libgomelcuda.so gomel_b200/ab/*.so
Each build runs the quick parity subset and the headline bench (no CPU leg); prints one line per build."""
import json
import os
import subprocess
import sys

for lib in sys.argv[1:]:
    env = dict(os.environ, GOMEL_CUDA_LIB=os.path.abspath(lib))
    t = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-x", "-q", "-k",
                        "stft_spectrum or from_mel_small or to_phase or from_phase_frame"], capture_output=True, text=True, env=env)
    ok = "passed" in t.stdout and "failed" not in t.stdout
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-stft"],
                         capture_output=True, text=True, env=env).stdout.strip().splitlines()
    try:
        d = json.loads(out[-1])
        r = d["roofline"]
        print(f"{os.path.basename(lib):40s} parity={'ok' if ok else 'FAIL'}  launch {r['avg_launch_ms']:.3f} ms  frac {r['frac']:.4f}  "
              f"value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}", flush=True)
    except Exception as e:      # noqa: BLE001
        print(f"{os.path.basename(lib):40s} parity={'ok' if ok else 'FAIL'}  bench failed: {e}", flush=True)
