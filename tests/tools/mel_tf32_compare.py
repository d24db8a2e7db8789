#!/usr/bin/env python3
"""north_star: "the triangular mel filterbank applied as a banded sparse kernel, with tensor cores used only if ncu
shows a dense TF32 projection wins within tolerance".  This script measures the alternative on a B200 (comparison
point only -- cuBLAS through torch, not on the product path):

  banded   the product: k_stft_fwd<MODE_MEL> = STFT + |X| + banded mel sums + clamp/log in one kernel, spectra never
           leave the SM.  Cost of the mel part = ToMel time - time of the same kernel writing raw half spectra.
  dense    |X| materialised in HBM [frames][2049] (it has to be: a tensor-core projection cannot sit inside the FFT
           kernel's register epilogue), then [frames x 2049] @ [2049 x 384] as TF32 / 3xTF32-equivalent fp32 GEMMs.

Accuracy is checked against a float64 projection of the same magnitudes (tolerance 1e-5 relative L2).
Writes gpurun_out/mel_tf32_compare.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import synth_clip                      # noqa: E402
from gomel_b200 import _lib                      # noqa: E402

ctx = _lib.Context(0)
cfg = _lib.make_config()
ctx.set_mel_tables(cfg, 0.0, 16000.0)
clips, n = 256, 441000
npad, frames, ola = _lib.frames(cfg, n)
stride = (npad + 3) & ~3
wav = np.zeros((clips, stride), np.float32)
base = np.stack([synth_clip(500 + c, 10.0) for c in range(8)]).astype(np.float32)
for c in range(clips):
    wav[c, :n] = base[c % 8]
d_sig = ctx.dev_malloc(wav.nbytes)
ctx.h2d(d_sig, wav)
d_mel = ctx.dev_malloc(clips * frames * 192 * 2 * 4)
spec = torch.empty((clips * frames, 2049, 2), dtype=torch.float32, device="cuda")


def timed(call, reps=10):
    for _ in range(3):
        call()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        call()
    return ctx.timer_stop() / reps


ms_tomel = timed(lambda: ctx.check(ctx.lib.gomel_to_mel_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_mel)))
ms_spec = timed(lambda: ctx.check(ctx.lib.gomel_stft_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames,
                                                         C.c_void_p(spec.data_ptr()))))
ctx.sync()
torch.cuda.synchronize()
mag = torch.linalg.vector_norm(spec, dim=2)                      # [frames, 2049] float32 in HBM
# dense weights of domel (mel/impl.go:310-345): column (mel, channel); channel 1 reads the bins shifted by one
flo, fhi, fmod, _, _, _ = _lib.mel_tables(2048, 192, 0.0, 16000.0)
W = np.zeros((2049, 384))
for i in range(192):
    lo, hi = int(flo[i]), int(fhi[i])
    if lo + 1 == hi:
        W[lo, i] += 1 - fmod[i]; W[hi, i] += fmod[i]
        W[lo + 1, 192 + i] += 1 - fmod[i]; W[hi + 1, 192 + i] += fmod[i]
    else:
        for k in range(lo, hi):
            W[k, i] += 1.0 / (hi - lo + 1)
            W[k + 1, 192 + i] += 1.0 / (hi - lo + 1)
Wd = torch.tensor(W, dtype=torch.float32, device="cuda")
sub = mag[:4096].double()
ref = sub @ torch.tensor(W, dtype=torch.float64, device="cuda")


def gemm_ms(allow_tf32):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.set_float32_matmul_precision("high" if allow_tf32 else "highest")
    for _ in range(3):
        out = mag @ Wd
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = mag @ Wd
    e1.record()
    torch.cuda.synchronize()
    err = float(torch.linalg.norm((mag[:4096] @ Wd).double() - ref) / torch.linalg.norm(ref))
    return e0.elapsed_time(e1) / 10, err


ms_tf32, err_tf32 = gemm_ms(True)
ms_fp32, err_fp32 = gemm_ms(False)
# |X| pass that a dense projection needs first: write + read of [frames][2049] floats
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m2 = torch.linalg.vector_norm(spec, dim=2)
e1.record()
torch.cuda.synchronize()
ms_mag = e0.elapsed_time(e1) / 10
# banded accuracy: product ToMel (linear domain) against the float64 projection of float64 magnitudes of the same spectra
mel = np.empty((clips * frames, 192, 2), np.float32)
ctx.d2h(mel, d_mel)
refb = (spec[:4096].double().pow(2).sum(2).sqrt() @ torch.tensor(W, dtype=torch.float64, device="cuda")).cpu().numpy()
got = np.exp(mel[:4096].astype(np.float64)).transpose(0, 2, 1).reshape(4096, 384)
refb = np.maximum(refb, 1e-5)
err_banded = float(np.linalg.norm(got - refb) / np.linalg.norm(refb))
n_frames = clips * frames
res = {
    "workload": f"{clips} x 10 s clips = {n_frames} frames, 2049 bins -> 192 mels x 2 channels",
    "banded_in_kernel": {"to_mel_ms": ms_tomel, "stft_only_ms_writing_half_spectra": ms_spec,
                         "note": "the STFT-only variant writes 16 KB of spectrum per frame and is slower than ToMel itself; the "
                                 "banded mel part costs < the whole ToMel kernel", "rel_l2_vs_float64_projection": err_banded},
    "dense_tensor_core": {"magnitude_pass_ms": ms_mag, "gemm_tf32_ms": ms_tf32, "gemm_tf32_rel_l2": err_tf32,
                          "gemm_fp32_ms": ms_fp32, "gemm_fp32_rel_l2": err_fp32,
                          "total_tf32_ms_excluding_stft": ms_mag + ms_tf32,
                          "flops": 2.0 * n_frames * 2049 * 384, "tf32_tflops": 2.0 * n_frames * 2049 * 384 / (ms_tf32 / 1e3) / 1e12},
    "tolerance": 1e-5,
}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mel_tf32_compare.json"), "w"), indent=1)
