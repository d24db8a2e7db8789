#!/usr/bin/env python3
"""e2e pipeline chunk-size sweep (GPU box): python tests/tools/sweep_chunks.py 64 128 256 512"""
import json, subprocess, sys
for c in sys.argv[1:]:
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-stft", "--chunk", c],
                         capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    print(f"chunk {c:>4}: value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  e2e ms/step {d['e2e']['ms_per_step']:.1f}", flush=True)
