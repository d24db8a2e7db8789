#!/usr/bin/env python3
"""e2e pipeline: sweep the smallest tapered chunk and the chunk size (GPU box)"""
import json, os, subprocess, sys
for floor_ in (16, 64, 128, 256):
    for chunk in (256, 512):
        env = dict(os.environ, GOMEL_CHUNK_TAPER_MIN=str(floor_))
        out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-stft", "--chunk", str(chunk)],
                             capture_output=True, text=True, env=env).stdout.strip().splitlines()
        d = json.loads(out[-1])
        print(f"taper_min {floor_:4d} chunk {chunk:4d}: value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  ({d['e2e']['ms_per_step']:.1f} ms)", flush=True)
