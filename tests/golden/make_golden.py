#!/usr/bin/env python3
"""Generates tests/golden/phase_ref.npz by IMPORTING the reference's own phase.py
(/root/reference/phase.py, with a stub for the missing `soundfile` module) and running its
float64 NumPy implementation on small seeded inputs.  Run in the build container only
(/root/reference does not exist on the GPU box); the .npz it writes is committed.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
sys.path.insert(0, REF)
import phase as RP  # noqa: E402  (the reference module)

out = {}
rng = np.random.default_rng(20261018)


def clip(n, sr, seed):
    r = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = np.zeros(n)
    for _ in range(6):
        a, f, c = r.uniform(0.05, 0.3), np.exp(r.uniform(np.log(60), np.log(12000))), r.uniform(-400, 400)
        x += a * np.sin(2 * np.pi * (f * t + 0.5 * c * t * t))
    x += 0.05 * r.standard_normal(n)
    return 0.9 * x / np.abs(x).max()


# --- to_phase / from_phase at both families (phase.py:113-220)
for name, sr, n in (("a48k", 48000, 24000), ("b44k", 44100, 30011), ("c48k_short", 48000, 5000)):
    ph = RP.Phase(sample_rate=sr)
    wav = clip(n, sr, 77 + n)
    spec = ph.to_phase(wav)
    rec = ph.from_phase(spec)
    out[f"{name}_wav"] = wav
    out[f"{name}_num_freqs"] = np.int64(ph.num_freqs)
    out[f"{name}_spec"] = spec
    out[f"{name}_rec"] = rec
# volume boost > 0 path (phase.py:216-217)
ph = RP.Phase(sample_rate=48000, volume_boost=1.666)
out["a48k_rec_boost"] = ph.from_phase(out["a48k_spec"])
# HDR doubles num_freqs (phase.py:52); keep the frame count tiny, from_phase is a Python loop
ph = RP.Phase(sample_rate=48000, HDR=True)
wav = clip(19199, 48000, 5)
out["hdr_wav"] = wav
out["hdr_num_freqs"] = np.int64(ph.num_freqs)
out["hdr_spec"] = ph.to_phase(wav)
out["hdr_rec"] = ph.from_phase(out["hdr_spec"])

# --- pad / is_padded (phase.py:352-402)
lens = np.array([1, 100, 19198, 19199, 19200, 19201, 20479, 20480, 20481, 44100, 441000, 158760000], np.int64)
out["pad_in"] = lens
out["pad_out"] = np.array([len(RP.pad(np.zeros(int(n)), 1280)) if n < 10**6 else
                           int(n) + (1280 - (int(n) - 15 * 1280) % 1280 - 1 if (int(n) - 15 * 1280) % 1280 else 0)
                           for n in lens], np.int64)
out["is_padded"] = np.array([RP.is_padded(int(a), int(b), 1280) for a, b in zip(lens, out["pad_out"])])
out["is_padded_neg"] = np.array([RP.is_padded(int(a), int(b) + 1, 1280) for a, b in zip(lens, out["pad_out"])])

# --- shrink / grow (phase.py:430-466; KAT shapes test_phase_comprehensive.py:66-70)
sg = rng.standard_normal((3 * 2048, 2))
out["shrink_in"] = sg
out["shrink_out"] = RP.shrink(sg, 4096, 768)
out["grow_out"] = RP.grow(out["shrink_out"], 4096, 768)

# --- zero_stuff_upsample KATs (test_zero_stuff.py:9-34; the code, not the printed text, is the truth)
for i, (a, zp, zs) in enumerate((([1., 2., 3., 4., 5.], 1, 1), ([1., 2., 3.], 1, 3), ([1., 2.], 1, 5),
                                  ([1., 2., 3., 4., 5.], 2, 1))):
    out[f"zs{i}_in"] = np.array(a)
    out[f"zs{i}_args"] = np.array([zp, zs], np.int64)
    out[f"zs{i}_out"] = RP.zero_stuff_upsample(np.array(a), zp, zs)

# --- float16 packing (phase.py:604-640)
vals = np.array([0.0, 1.0, -1.0, 44100.0, 48000.0, 1289.4, 3.06e-5, 65504.0, 1e-8, -11.5129, 7.25])
out["f16_in"] = vals
out["f16_bytes"] = np.frombuffer(b"".join(RP.pack_float16_to_bytes(v) for v in vals), np.uint8)
out["f16_back"] = np.array([RP.unpack_bytes_to_float64(bytes(out["f16_bytes"][2 * i:2 * i + 2])) for i in range(len(vals))])

# --- save_image / load_image, 8-bit, with and without asinh passes (phase.py:643-852)
import tempfile  # noqa: E402
from PIL import Image as _PILImage  # noqa: E402
for tag, ihs in (("img0", 0), ("img2", 2)):
    spec = out["c48k_short_spec"]
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "x.png")
        RP.save_image(f, spec.copy(), 768, 1289.4, 48000, True, False, ihs)
        out[f"{tag}_pixels"] = np.array(_PILImage.open(f).convert("RGB"), np.uint8)
        buf, samples, sr, nf = RP.load_image(f, True, False, ihs)
        out[f"{tag}_loaded"] = buf
        out[f"{tag}_meta"] = np.array([samples, sr, nf], np.float64)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "phase_ref.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "arrays")
