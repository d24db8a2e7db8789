#!/usr/bin/env python3
"""tile-size sweep of the headline bench (run on the GPU box): python tests/tools/sweep_tiles.py 0 38 58 ..."""
import json
import subprocess
import sys

for t in sys.argv[1:]:
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-stft", "--tile", t],
                         capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    r = d["roofline"]
    print(f"tile {t:>4}: launch {r['avg_launch_ms']:.3f} ms  frac {r['frac']:.4f}  value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}", flush=True)
