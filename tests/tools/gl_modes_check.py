#!/usr/bin/env python3
"""Griffin-Lim precision modes on the GPU (run on a B200 box):

  check   fused float64 kernel (GOMEL_FLAG_F64) vs the round-1 strict path (GOMEL_FLAG_F64_REF) vs the CPU oracle,
          several tilings; hybrid (float64 lead iterations + float32) vs the fused float64 result
  speed   frame-iterations/s of the float64 lead kernel and of the float32 kernel on a 256-clip batch
  sweep   >= 4 clips x >= 16 start seeds, 10 s, GL-32 and GL-100: rel-L2 of lead = 0 / 1 / 2 / 3 / 4 / 6 / 8 against the
          all-float64 run; writes gpurun_out/gl_parity_sweep.json (summarised in profiles/r02_gl_parity_sweep.md)
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import rel_l2, synth_clip                      # noqa: E402
from oracle import oracle as O                           # noqa: E402
from gomel_b200 import NewMel, _lib                      # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def mel_obj(iters, strict=False):
    m = NewMel()
    m.NumMels, m.MelFmin, m.MelFmax, m.Window, m.Resolut = 192, 0, 16000, 1280, 4096
    m.GriffinLimIterations, m.Strict = iters, strict
    return m


def run(mel, init, iters, strict=False, lead=None):
    ctx = _lib.default_context(0)
    prev = ctx.set_gl_precision(lead, -1) if lead is not None else None     # exactly `lead` float64 iterations
    m = mel_obj(iters, strict)
    m.InitSignal = init
    out = m.FromMel(mel.copy())
    if prev is not None:
        ctx.set_gl_precision(*prev)
    return out


def check():
    ctx = _lib.default_context(0)
    ok = True
    for secs, iters, seed, tiles in ((0.45, 3, 2, (0, 4)), (1.5, 32, 5, (0, 6, 12)), (1.0, 100, 13, (0, 8))):
        wav = synth_clip(15, secs)
        ocfg = O.config(gl_iters=iters)
        mel = O.to_mel(ocfg, wav)
        frames = len(mel) // 192
        init = np.random.default_rng(seed).random(4096 + (frames - 1) * 1280)
        ref = O.from_mel(ocfg, mel, init)
        r64 = run(mel, init, iters, "ref")
        print(f"[{secs}s it{iters}] REF64 vs oracle {rel_l2(r64, ref):.2e}", flush=True)
        for tile in tiles:
            ctx.set_tile_frames(tile)
            f64 = run(mel, init, iters, True)
            hyb = run(mel, init, iters, False)
            f32 = run(mel, init, iters, False, lead=0)
            ctx.set_tile_frames(0)
            e = rel_l2(f64, ref)
            print(f"   tile {tile:3d}: fused F64 vs oracle {e:.2e} | vs REF64 {rel_l2(f64, r64):.2e} | hybrid vs F64 {rel_l2(hyb, f64):.2e}"
                  f" | fp32 vs F64 {rel_l2(f32, f64):.2e}", flush=True)
            ok &= e < 1e-10 and rel_l2(hyb, f64) < 1e-4
    # 10 s clip
    wav = synth_clip(0, 10.0)
    mel = O.to_mel(O.config(), wav)
    for seed in (5, 13):
        init = np.random.default_rng(seed).random(440576)
        f64 = run(mel, init, 32, True)
        r64 = run(mel, init, 32, "ref")
        line = f"[10s it32 seed {seed}] fused F64 vs REF64 {rel_l2(f64, r64):.2e} |"
        for lead in (0, 1, 2, 3, 4, 6, 8):
            line += f" lead{lead}: {rel_l2(run(mel, init, 32, False, lead=lead), f64):.1e}"
        print(line, flush=True)
        ok &= rel_l2(f64, r64) < 1e-10
    init = np.random.default_rng(5).random(440576)
    t0 = time.time()
    ref = O.from_mel(O.config(gl_iters=32), mel, init)
    print(f"[10s it32 seed 5] fused F64 vs oracle {rel_l2(run(mel, init, 32, True), ref):.2e}; hybrid vs oracle "
          f"{rel_l2(run(mel, init, 32), ref):.2e} (oracle {time.time() - t0:.0f} s)", flush=True)
    print("CHECK", "OK" if ok else "FAILED", flush=True)
    return ok


def speed(clips=256):
    ctx = _lib.Context(0)
    cfg = _lib.make_config(gl_iters=32)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    n = 441000
    _, frames, ola = _lib.frames(cfg, n)
    nb = 8
    wav = np.stack([synth_clip(c, 10.0) for c in range(nb)]).astype(np.float32)
    base = np.empty((nb, frames * 192 * 2), np.float32)
    ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), nb, n,
                                              base.ctypes.data_as(C.c_void_p), 8))
    mel = np.concatenate([base] * (clips // nb))
    d_mel = ctx.dev_malloc(mel.nbytes)
    d_out = ctx.dev_malloc(clips * ola * 4)
    ctx.h2d(d_mel, mel)
    res = {}
    for name, iters, flags, lead in (("f64_all_8it", 8, _lib.FLAG_F64, None), ("hybrid4_32it", 32, 0, 4), ("hybrid2_32it", 32, 0, 2),
                                     ("f32_32it", 32, 0, 0)):
        c = _lib.make_config(gl_iters=iters, flags=flags)
        if lead is not None:
            ctx.set_gl_precision(lead, -1)
        for rep in range(3):
            ctx.timer_start()
            ctx.check(ctx.lib.gomel_from_mel_dev(ctx.h, C.byref(c), d_mel, clips, frames, None, 7, ola, d_out))
            ms = ctx.timer_stop()
        lms, ln = ctx.last_lead_kernel_ms()
        hms, hn = ctx.last_hot_kernel_ms()
        fi = clips * frames
        res[name] = {"ms": ms, "lead_ms": lms, "lead_iters": ln, "f32_ms": hms, "f32_iters": hn,
                     "lead_frame_iter_per_s": fi * ln / (lms / 1e3) if ln else None,
                     "f32_frame_iter_per_s": fi * hn / (hms / 1e3) if hn else None,
                     "audio_s_per_s": clips * frames * 1280 / 44100 / (ms / 1e3)}
        print(name, json.dumps(res[name]), flush=True)
    ctx.set_gl_precision(16, 16)
    json.dump(res, open(os.path.join(OUT, f"gl_modes_speed_{clips}.json"), "w"), indent=1)


def sweep(n_clips=4, n_seeds=16):
    rows = []
    kinds = [("clip%d" % c, synth_clip(c, 10.0)) for c in range(n_clips)]
    kinds.append(("white_noise", np.random.default_rng(77).uniform(-1, 1, 441000)))
    kinds.append(("silence", np.zeros(441000)))
    leads = (0, 1, 2, 3, 4, 6, 8)
    for name, wav in kinds:
        mel = O.to_mel(O.config(), wav)
        for iters in (32, 100):
            for s in range(n_seeds):
                seed = 100 + s
                init = np.random.default_rng(seed).random(440576)
                f64 = run(mel, init, iters, True)
                row = {"clip": name, "iters": iters, "seed": seed}
                for lead in leads:
                    row["lead%d" % lead] = rel_l2(run(mel, init, iters, False, lead=lead), f64)
                if s == 0:
                    row["f64_vs_ref64"] = rel_l2(f64, run(mel, init, iters, "ref"))
                rows.append(row)
            sub = [r for r in rows if r["clip"] == name and r["iters"] == iters]
            print(name, iters, " ".join("lead%d max %.1e" % (l, max(r["lead%d" % l] for r in sub)) for l in leads), flush=True)
    summ = {}
    for iters in (32, 100):
        sub = [r for r in rows if r["iters"] == iters]
        summ[iters] = {"pairs": len(sub)}
        for lead in leads:
            v = np.array([r["lead%d" % lead] for r in sub])
            summ[iters]["lead%d" % lead] = {"max": float(v.max()), "median": float(np.median(v)), "pass_frac_1e-4": float(np.mean(v <= 1e-4))}
    json.dump({"rows": rows, "summary": summ}, open(os.path.join(OUT, "gl_parity_sweep.json"), "w"), indent=1)
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    if what == "check":
        sys.exit(0 if check() else 1)
    elif what == "speed":
        speed(int(sys.argv[2]) if len(sys.argv) > 2 else 256)
    elif what == "sweep":
        sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 4, int(sys.argv[3]) if len(sys.argv) > 3 else 16)
