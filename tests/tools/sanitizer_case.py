#!/usr/bin/env python3
"""small end-to-end pass over every kernel for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python tests/tools/sanitizer_case.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip
from gomel_b200 import NewMel, Phase, _lib, timesplit

ctx = _lib.default_context(0)
m = NewMel(); m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut, m.GriffinLimIterations = 192, 0, 16000, 1280, 4096, 3
wav = synth_clip(0, 0.9)
for tile in (0, 4, 6):
    ctx.set_tile_frames(tile)
    mel = m.ToMel(wav)
    frames = len(mel) // 192
    m.InitSignal = np.random.default_rng(0).random(4096 + (frames - 1) * 1280)
    out = m.FromMel(mel.copy())
    ph = Phase(num_freqs=768)
    spec = ph.to_phase(wav)
    rec = ph.from_phase(spec)
    assert np.isfinite(out).all() and np.isfinite(rec).all()
ctx.set_tile_frames(0)
m.Strict = True
out64 = m.FromMel(mel.copy())
m.Strict = False
img = m.Image(mel)
rgb, mm = ctx.quantise(mel, 192, _lib.Q_SINGLE_MINMAX)
back = ctx.dequantise(rgb[:, :2], False, mm[0], mm[0], mm[2], mm[2])
# odd frame count + time split emulation
wav2 = synth_clip(1, 0.5)
mel2 = m.ToMel(wav2)
cfg = _lib.make_config(gl_iters=2)
ctx.set_mel_tables(cfg, 0.0, 16000.0)
fr2 = len(mel2) // 192
init2 = np.random.default_rng(1).random(4096 + (fr2 - 1) * 1280).astype(np.float32)
ts = timesplit.run_local(ctx, cfg, mel2, init2, 2, 2, tile_frames=4, overlap=True)
assert np.isfinite(ts).all()
print("SANITIZER_CASE_DONE", frames, fr2)
