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
m.Strict = "ref"
ref64 = m.FromMel(mel.copy())
m.Strict = False
assert np.abs(out64 - ref64).max() < 1e-9
# float64 lead iterations followed by float32 ones, several tilings; batched pipeline with a ragged last chunk
m.GriffinLimIterations = 19
import ctypes as C
for tile in (0, 4, 10):
    ctx.set_tile_frames(tile)
    hy = m.FromMel(mel.copy())
    assert np.isfinite(hy).all()
ctx.set_tile_frames(0)
cfgb = _lib.make_config(gl_iters=18)
ctx.set_mel_tables(cfgb, 0.0, 16000.0)
mb = np.stack([mel.astype(np.float32)] * 5)
ob = np.empty((5, 4096 + (frames - 1) * 1280), np.float32)
ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(cfgb), mb.ctypes.data_as(C.c_void_p), 5, frames, None, 3,
                                            ob.ctypes.data_as(C.c_void_p), 2))
assert np.isfinite(ob).all()
m.GriffinLimIterations = 3
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
cfg6 = _lib.make_config(gl_iters=18)
ts6 = timesplit.run_local(ctx, cfg6, mel2, init2, 18, 2, tile_frames=4, overlap=True)        # crosses the float64 -> float32 hand-over
assert np.isfinite(ts6).all()
pcfg = _lib.make_config(n_mels=0, n_freqs=768, gl_iters=0)
tp = timesplit.phase_istft_local(ctx, pcfg, ph.to_phase(wav2), 2, tile_frames=4)
assert np.isfinite(tp).all()
print("SANITIZER_CASE_DONE", frames, fr2)
