#!/usr/bin/env python3
"""prints the measured parity margins (run on the GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import rel_l2, synth_clip
from oracle import oracle as O
from gomel_b200 import NewMel, _lib

def gl(wav, iters, seed, tile=0):
    m = NewMel(); m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut, m.GriffinLimIterations = 192, 0, 16000, 1280, 4096, iters
    ocfg = O.config(gl_iters=iters); mel = O.to_mel(ocfg, wav); frames = len(mel)//192
    init = np.random.default_rng(seed).random(4096+(frames-1)*1280); m.InitSignal = init
    _lib.default_context(0).set_tile_frames(tile); got = m.FromMel(mel.copy()); _lib.default_context(0).set_tile_frames(0)
    return rel_l2(got, O.from_mel(ocfg, mel, init))

rng = np.random.default_rng(8)
print("GL-2   1 s clip      ", gl(synth_clip(11, 1.0), 2, 1))
print("GL-32  1.5 s clip    ", gl(synth_clip(12, 1.5), 32, 2))
print("GL-100 0.8 s clip    ", gl(synth_clip(14, 0.8), 100, 3))
print("GL-32  white noise   ", gl(rng.uniform(-1, 1, 30000), 32, 4))
print("GL-32  silence       ", gl(np.zeros(30000), 32, 5))
print("GL-100 white noise   ", gl(rng.uniform(-1, 1, 30000), 100, 6))
print("GL-32  10 s clip     ", gl(synth_clip(0, 10.0), 32, 7))
m = NewMel(); m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut = 192, 0, 16000, 1280, 4096
w = synth_clip(0, 10.0); a = m.ToMel(w); b = O.to_mel(O.config(), w)
print("ToMel 10 s linear    ", rel_l2(np.exp(a), np.exp(b)), " log ", rel_l2(a, b))
