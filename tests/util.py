"""Shared helpers for the tests and bench: seeded synthetic clips, rel-L2."""
import numpy as np

SR = 44100


def synth_clip(c, seconds=10.0, sr=SR, n=None):
    """clip c = sum_k a_k sin(2 pi (f_k t + r_k t^2 / 2)) + 0.05 N(0,1), peak-normalised to 0.9"""
    rng = np.random.default_rng(1000 + c)
    n = int(round(seconds * sr)) if n is None else n
    t = np.arange(n) / sr
    x = np.zeros(n)
    for _ in range(8):
        a = rng.uniform(0.05, 0.3)
        f = np.exp(rng.uniform(np.log(60.0), np.log(12000.0)))
        r = rng.uniform(-400.0, 400.0)
        x += a * np.sin(2 * np.pi * (f * t + 0.5 * r * t * t))
    x += 0.05 * rng.standard_normal(n)
    return 0.9 * x / np.abs(x).max()


def rel_l2(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    if not (np.iscomplexobj(a) or np.iscomplexobj(b)):
        a, b = a.astype(np.float64), b.astype(np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))
