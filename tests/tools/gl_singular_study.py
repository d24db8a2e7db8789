#!/usr/bin/env python3
"""CPU study (numpy / scipy, no GPU): does a per-bin statistic of the float64 Griffin-Lim trajectory predict the
(clip, start signal) pairs whose float32 tail leaves it?   python tests/tools/gl_singular_study.py <clip> <seed0> <n> [lead]"""
import os
import sys

import numpy as np
import scipy.fft as sf

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip                       # noqa: E402
from oracle import oracle_np as onp               # noqa: E402

N, H = 4096, 1280


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


def gl(M, init, iters, lead, want_stats=False):
    F = M.shape[0]
    w = np.hanning(N)
    ola = N + (F - 1) * H
    idx = np.arange(N)[None, :] + np.arange(F)[:, None] * H
    sig = np.array(init, np.float64)
    stats = []
    Mrms = np.sqrt(np.mean(M * M))
    for it in range(iters):
        if it == lead:
            sig = sig.astype(np.float32)
        f32 = it >= lead
        dt = np.float32 if f32 else np.float64
        fr = sig[idx] * w.astype(dt)
        X = sf.rfft(fr, axis=1)
        mag = np.abs(X)
        if want_stats and not f32:
            safe = np.maximum(mag, 1e-300)
            Xrms = np.sqrt(np.mean(mag * mag, axis=1, keepdims=True))
            q = M / safe                                  # ratio
            lev = M * Xrms / safe / Mrms                  # phase leverage of a unit absolute error, in units of the clip's rms magnitude
            k = np.unravel_index(np.argmax(lev), lev.shape)
            stats.append((float(q.max()), float(lev.max()), k, float(M[k] / Mrms), float(mag[k] / Xrms[k[0], 0])))
        unit = np.where(mag > 0, X / np.where(mag > 0, mag, 1), 1.0).astype(X.dtype)
        Y = (M.astype(dt) * unit).astype(X.dtype)
        y = sf.irfft(Y, n=N, axis=1) * w.astype(dt)
        new = np.zeros(ola, dt)
        for f in range(F):
            new[f * H:f * H + N] += y[f]
        sig = new
    return sig, stats


clip = int(sys.argv[1]); seed0 = int(sys.argv[2]); n = int(sys.argv[3])
lead = int(sys.argv[4]) if len(sys.argv) > 4 else 16
wav = synth_clip(clip, 10.0)
mel = onp.to_mel(wav)
M = np.abs(onp.gl_magnitudes(mel))
F = M.shape[0]
for s in range(seed0, seed0 + n):
    init = np.random.default_rng(s).random(N + (F - 1) * H)
    ref, st = gl(M, init, 32, 32, True)
    hyb, _ = gl(M, init, 32, lead)
    e = rel_l2(hyb, ref)
    tail = st[lead:] if lead < 32 else st
    levs = [x[1] for x in st]
    print(f"seed {s}: policy(lead {lead}) err {e:.2e} | max leverage over iterations {lead}..31: {max(levs[lead:]):.3e} "
          f"at it {lead + int(np.argmax(levs[lead:]))} | per-iteration leverage: " + " ".join(f"{v:.1e}" for v in levs), flush=True)
