#!/usr/bin/env python3
"""device + e2e throughput against the automatic tiling's wave target (GOMEL_TILE_WAVES), one process per value"""
import json, os, subprocess, sys
for w in sys.argv[1:] or ("3", "4", "6", "8", "12"):
    env = dict(os.environ, GOMEL_TILE_WAVES=w)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu", "--no-stft"],
                         capture_output=True, text=True, env=env).stdout.strip().splitlines()
    d = json.loads(out[-1])
    print(f"waves {w:>3s}: value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  pcm16 {d['e2e']['pcm16']['value']:.0f}", flush=True)
