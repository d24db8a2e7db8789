#!/usr/bin/env python3
"""bench.py -- headline benchmark of the gomel_b200 hot path (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W          # product arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...          # reference arm: the CPU float64 port of
                                                           # the Go path (oracle/), all host threads

Metric (BASELINE.json): mel->wav audio-seconds per second, Griffin-Lim 32 iterations, 192 mels,
Resolut 4096, Window 1280, on 1024 synthetic 10 s 44.1 kHz clips per GPU (configs[3]); clips are
sharded across ranks with no data-path collective (weak scaling: 1024 clips on every rank).
A step = one FromMel over the rank's whole batch: K3 (mel -> target magnitudes), 32 launches of
the Griffin-Lim iteration kernel, the tile-boundary fix-up.

  value     device-resident: mel spectrograms already in HBM when the timed region starts
  e2e       the same metric through the C ABI call gomel_from_mel_batch_host with PINNED HOST
            buffers: H2D of the mel batch and D2H of the waveforms inside the timed region
  roofline  the Griffin-Lim iteration kernel against the measured HBM copy bandwidth
  cpu_baseline  the oracle port timed on this box's host cores on a bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 44100
N_FFT, HOP, N_MELS, GL_ITERS = 4096, 1280, 192, 32
CLIP_SECONDS = 10.0
CLIPS_PER_GPU = 1024
BASE_CLIPS = 32                      # distinct spectrograms, tiled to CLIPS_PER_GPU
BYTES_PER_FRAME_ITER = 18436         # 2049*4 magnitudes + 1280*4 read + 1280*4 write
BYTES_PER_FRAME_ITER_F64 = 36872     # the float64 lead iterations move the same items as doubles
BYTES_PER_FRAME_ONCE = 9732          # K3: read mel 1536 + write magnitudes 8196
METRIC = "mel2wav_griffinlim32_audio_seconds_per_second"
UNIT = "audio-s/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:                      # region shorter than one sample: take the nearest rows
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_cpu_share(local_rank, local_world):
    """No NUMA information: give every local rank its own contiguous share of the allowed CPUs, so the ranks' copy
    threads and first-touch pinned pages do not migrate over each other."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = len(cpus) // max(local_world, 1)
        if per >= 1 and local_world > 1:
            os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]))
            return [cpus[local_rank * per], cpus[(local_rank + 1) * per - 1]]
    except Exception:
        pass
    return None


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host thread (and therefore its first-touch pinned buffers) to the NUMA node of its GPU,
    so that the end-to-end H2D/D2H traffic of 8 ranks does not cross the socket interconnect.  Plumbing only."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------- reference arm
def cpu_reference(n_threads, seconds_per_clip, steps, warmup):
    """Times the oracle port of mel.FromMel (GL 32) on n_threads clips in parallel (OpenMP)."""
    from oracle import oracle as O
    from util import synth_clip
    cfg = O.config(num_mels=N_MELS, window=HOP, resolut=N_FFT, gl_iters=GL_ITERS)
    base = O.to_mel(cfg, synth_clip(0, seconds_per_clip))
    frames = len(base) // N_MELS
    ola = N_FFT + (frames - 1) * HOP
    mel = np.stack([base] * n_threads)
    init = np.random.default_rng(5000).random((n_threads, ola))
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.from_mel_batch(cfg, mel, init, threads=n_threads)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    audio_s = n_threads * frames * HOP / SR
    return audio_s * len(times) / sum(times), frames, float(np.mean(times))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    # ~1.9 CPU-seconds per audio-second per core for GL-32: keep the whole run to a few minutes
    seconds = float(min(CLIP_SECONDS, max(1.0, 150.0 / max(total_steps, 1) / 1.9)))
    value, frames, step_s = cpu_reference(cores, seconds, args.steps, args.warmup)
    sample = (f"{cores} clips x {seconds:.2f} s ({frames} frames each), one clip per host thread (OpenMP), "
              f"float64 port of mel.FromMel GL-{GL_ITERS} incl. undomel; same per-frame work as the 10 s clips")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Go with un-vendored modules and no Go toolchain in the image: the CPU arm is the "
                "line-by-line float64 C port in oracle/ (kind=port)",
    }
    print(json.dumps(line))
    return 0


def workload_config(n_gpus):
    return {"workload": f"configs[3]: batched FromMel Griffin-Lim {GL_ITERS} it, {CLIPS_PER_GPU} synthetic 10 s 44.1 kHz "
                        f"clips per GPU (192 mels, Resolut 4096, Window 1280), sharded by clip, no collective",
            "clips_per_gpu": CLIPS_PER_GPU, "frames_per_clip": 342, "gl_iters": GL_ITERS,
            "distinct_spectrograms": BASE_CLIPS, "start_signal": "device U[0,1) per clip (seeded)",
            "l2": "inputs larger than L2 (2.9 GB magnitudes + 1.8 GB signal per buffer)",
            "parallelism": f"clip-sharded x{n_gpus}"}


# ------------------------------------------------------------------------------- product arm
def host_link_probe(local_rank, use_dist, mb=512, reps=4):
    """Ceiling of the end-to-end path: pinned H2D and D2H copies running concurrently on this GPU -- and, under
    torchrun, on every rank's GPU at the same time (barrier before, max time over ranks).  torch is plumbing here."""
    import torch
    dev = torch.device("cuda", local_rank)
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def one():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    one()
    torch.cuda.synchronize(dev)
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    world = 1
    if use_dist:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        world = dist.get_world_size()
    per_dir = n * reps / dt / 1e9
    del h_in, h_out, d_in, d_out
    torch.cuda.empty_cache()
    return {"what": f"concurrent pinned H2D + D2H of {mb} MiB per direction per GPU, all {world} GPU(s) at once, max over ranks",
            "gbs_per_direction_per_gpu": per_dir, "aggregate_gbs_both_directions": 2 * per_dir * world}


def timesplit_measure(ctx, cfg, rank, local_rank, world, use_dist, seconds, steps, warmup, ts_tile=0, ts_edge=8,
                      overlap=True, exchange_kind="native"):
    """configs[4]: ONE long clip, Griffin-Lim 32 it, frames split by time across the ranks, two 2816-element
    partials exchanged per boundary per iteration over NCCL.  A step = the whole FromMel of the clip: target
    magnitudes + start signal (gomel_ts_load from device-resident mel), 32 iterations under the precision policy
    (float64 lead iterations exchange doubles), timed with the host clock around a stream-synchronised region,
    max over ranks."""
    from gomel_b200 import _lib, timesplit
    from util import synth_clip
    if use_dist:
        import torch
        import torch.distributed as dist
    n_total = int(round(seconds * SR))
    _, frames_total, ola = _lib.frames(cfg, n_total)
    # spectrogram: a 60 s synthetic clip's mel (GPU ToMel) repeated along time
    base_n = 60 * SR
    wav = synth_clip(9000, 60.0).astype(np.float32)[None, :]
    _, base_frames, _ = _lib.frames(cfg, base_n)
    base_mel = np.empty((1, base_frames * N_MELS * 2), np.float32)
    ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), 1, base_n,
                                              base_mel.ctypes.data_as(C.c_void_p), 1))
    base_mel = base_mel.reshape(base_frames, N_MELS * 2)
    if ts_tile <= 0:
        # frames per interior tile: fill whole waves of 2 CTAs/SM (296 CTAs) with this rank's tiles; long tiles
        # amortise the per-tile prologue, the short boundary tiles (ts_edge) keep the exchange path short
        per_rank = (frames_total + world - 1) // world
        best = (0.0, 16)
        for T in range(16, 122, 2):
            tiles = (per_rank + T - 1) // T
            eff = tiles / (((tiles + 295) // 296) * 296.0) + 0.0005 * T
            if eff > best[0]:
                best = (eff, T)
        ts_tile = best[1]
    s = timesplit.Session(ctx, cfg, frames_total, rank, world, ts_tile, ts_edge)
    idx = (np.arange(s.frame_begin, s.frame_begin + s.n_frames) % base_frames)
    mel_local = np.ascontiguousarray(base_mel[idx].reshape(-1, 2))
    d_mel = ctx.dev_malloc(mel_local.nbytes)
    ctx.h2d(d_mel, mel_local)
    exchange = timesplit.NcclExchange(s) if (use_dist and exchange_kind != "native") else (lambda it: None)
    native = timesplit.NativeNccl(s) if exchange_kind == "native" else None

    def barrier():
        s.sync()
        if use_dist:
            torch.cuda.synchronize()
            dist.barrier()

    def one_step(k):
        s.load_dev(d_mel, None, seed=9001 + k)
        if native is not None:
            native.run(0, GL_ITERS, overlap=overlap)
            return
        for it in range(GL_ITERS):
            if overlap:
                s.iterate(it, 1); exchange(it); s.iterate(it, 2)
            else:
                s.iterate(it, 0); exchange(it)

    for k in range(warmup):
        one_step(k)
    barrier()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    for k in range(steps):
        one_step(100 + k)
    s.sync()
    ms = (time.perf_counter() - t0) * 1e3
    barrier()
    launches = ctx.launch_count() - launches0
    if use_dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    lead = s.lead_iters()
    s.close()
    ctx.dev_free(d_mel)
    audio_s = frames_total * HOP / SR
    peak, _ = peaks()
    fi = frames_total * GL_ITERS * steps / (ms / 1e3)
    bytes_iter = (min(lead, GL_ITERS) * BYTES_PER_FRAME_ITER_F64 + max(GL_ITERS - lead, 0) * BYTES_PER_FRAME_ITER) / GL_ITERS
    return {
        "workload": "timesplit", "metric": METRIC, "value": audio_s * steps / (ms / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "ms_per_iteration": ms / steps / GL_ITERS, "higher_is_better": True, "scaling": "strong",
        "dtype": f"f64 x{min(lead, GL_ITERS)} + f32 x{max(GL_ITERS - lead, 0)} iterations", "data": "synthetic",
        "config": {"workload": f"configs[4]: one {seconds:.0f} s 44.1 kHz clip ({frames_total} frames), Griffin-Lim "
                               f"{GL_ITERS} it, frames split by time over {world} GPU(s), NCCL exchange of two 2816-element "
                               f"partials per boundary per iteration ({exchange_kind} NCCL), tile {ts_tile} frames, boundary tiles {ts_edge}, overlap={bool(overlap)}",
                   "step": "mel -> magnitudes + start signal + all iterations (whole FromMel of the clip)",
                   "timing": "host wall clock around stream-synchronised region, max over ranks"},
        "frame_iterations_per_s": fi, "hbm_frac_whole_job": fi * bytes_iter / 1e9 / (peak * world),
        "gpu_launches": int(launches)}


def timesplit_phase_measure(ctx, rank, world, use_dist, seconds, steps, tile=64):
    """SURVEY 8(e) third case: phase.ISTFT of ONE long clip (NumFreqs 768), frames split by time over the ranks, one
    NCCL transfer of 2816 floats per rank boundary (gomel_ts_phase_run_nccl).  Device-resident spectrogram slices."""
    from gomel_b200 import _lib, timesplit
    if use_dist:
        import torch
        import torch.distributed as dist
    nfq = 768
    pcfg = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=0, n_freqs=nfq, gl_iters=0)
    n_total = int(round(seconds * SR))
    _, frames_total, ola = _lib.frames(pcfg, n_total)
    s = timesplit.Session(ctx, pcfg, frames_total, rank, world, tile, 0)
    rng = np.random.default_rng(77 + rank)
    base = rng.standard_normal((2048, nfq * 2)).astype(np.float32)
    spec = np.ascontiguousarray(base[np.arange(s.n_frames) % 2048])
    d_spec, d_out = ctx.dev_malloc(spec.nbytes), ctx.dev_malloc(s.n_samples * 4)
    ctx.h2d(d_spec, spec)
    timesplit.NativeNccl(s)

    def barrier():
        s.sync()
        if use_dist:
            torch.cuda.synchronize()
            dist.barrier()
    for _ in range(2):
        ctx.check(ctx.lib.gomel_ts_phase_run_nccl(s.h, d_spec, d_out))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.check(ctx.lib.gomel_ts_phase_run_nccl(s.h, d_spec, d_out))
    s.sync()
    ms = (time.perf_counter() - t0) * 1e3
    barrier()
    if use_dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    s.close()
    ctx.dev_free(d_spec)
    ctx.dev_free(d_out)
    peak, _ = peaks()
    fps = frames_total * steps / (ms / 1e3)
    return {"workload": f"phase.ISTFT of one {seconds:.0f} s clip ({frames_total} frames, NumFreqs {nfq}) split by time over {world} GPU(s), "
                        f"one NCCL transfer of 2816 floats per boundary, tile {tile} frames; host clock, max over ranks",
            "ms_per_step": ms / steps, "frames_per_s": fps, "audio_s_per_s": frames_total * HOP / SR * steps / (ms / 1e3),
            "hbm_frac_whole_job": fps * 11264 / 1e9 / (peak * world), "scaling": "strong"}


def parity_record(ctx, cfg, _lib, base_mel, frames, ola, n=4):
    """rel-L2 of n bench clips through the benchmarked call (gomel_from_mel_batch_host, default precision policy,
    injected start signals) against the all-float64 fused kernel on the same inputs -- the number the north-star
    tolerance (1e-4) is about; the full sweep is tests/test_gpu_parity.py::test_gl_precision_policy_sweep."""
    nb = min(n, len(base_mel))
    mel32 = np.ascontiguousarray(base_mel[:nb])
    init32 = np.random.default_rng(4242).random((nb, ola), dtype=np.float32)
    out = np.empty((nb, ola), np.float32)
    ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(cfg), mel32.ctypes.data_as(C.c_void_p), nb, frames,
                                                init32.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_void_p), nb))
    cfg64 = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=GL_ITERS, flags=_lib.FLAG_F64)
    prev = ctx.set_gl_precision(0, -1)
    out32 = np.empty((nb, ola), np.float32)
    try:
        ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(cfg), mel32.ctypes.data_as(C.c_void_p), nb, frames,
                                                    init32.ctypes.data_as(C.c_void_p), 0, out32.ctypes.data_as(C.c_void_p), nb))
    finally:
        ctx.set_gl_precision(*prev)
    errs, errs32 = [], []
    for c in range(nb):
        exact = ctx.from_mel(cfg64, mel32[c].reshape(-1, 2).astype(np.float64), init=init32[c].astype(np.float64))
        d = np.linalg.norm(exact)
        errs.append(float(np.linalg.norm(out[c] - exact) / d))
        errs32.append(float(np.linalg.norm(out32[c] - exact) / d))
    return {"what": f"{nb} bench clips, benchmarked call vs all-float64 fused kernel (GOMEL_FLAG_F64), same float32 inputs",
            "rel_l2": errs, "rel_l2_max": max(errs), "tolerance": 1e-4, "all_float32_rel_l2": errs32,
            "sweep": "profiles/r02_gl_parity_sweep.md, r02_gl_guard.md (44,352 pairs at 32 iterations under the default policy with its guard: one at 1.3e-4, one at 9.8e-5, all others <= 2.4e-5; 1,056 at 100 iterations: max 6.9e-6)"}


def run_product(args):
    rank, local_rank, world = dist_env()
    n_gpus = world                      # one process per GPU; --gpus is informational when launched by torchrun
    use_dist = world > 1
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gomel_b200 import _lib
    from util import synth_clip

    numa_node = bind_to_gpu_numa(local_rank) if args.numa else None
    cpu_share = bind_cpu_share(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if (args.numa and numa_node is None) else None
    ctx = _lib.Context(local_rank)
    cfg = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=GL_ITERS)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    if args.tile:
        ctx.set_tile_frames(args.tile)
    clips = args.clips
    n = int(round(CLIP_SECONDS * SR))
    npad, frames, ola = _lib.frames(cfg, n)
    mel_per = frames * N_MELS * 2

    # ---- synthetic spectrograms: BASE_CLIPS distinct clips -> GPU ToMel (the product path) -> tiled
    nb = min(BASE_CLIPS, clips)
    wav = np.stack([synth_clip(rank * 100000 + c, CLIP_SECONDS) for c in range(nb)]).astype(np.float32)
    base_mel = np.empty((nb, mel_per), np.float32)
    ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), nb, n,
                                              base_mel.ctypes.data_as(C.c_void_p), 16))
    h_mel, h_mel_owner = ctx.pinned_array((clips, mel_per), np.float32)
    for c in range(clips):
        h_mel[c] = base_mel[c % nb]
    h_out, h_out_owner = ctx.pinned_array((clips, ola), np.float32)
    d_mel = ctx.dev_malloc(h_mel.nbytes)
    d_out = ctx.dev_malloc(clips * ola * 4)
    ctx.h2d(d_mel, h_mel)

    def barrier():
        ctx.sync()
        if use_dist:
            torch.cuda.synchronize()
            dist.barrier()

    def reduce_max(vals):
        if not use_dist:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def step_device(seed, n_clips=clips, c=cfg):
        ctx.check(ctx.lib.gomel_from_mel_dev(ctx.h, C.byref(c), d_mel, n_clips, frames, None, seed, ola, d_out))

    def step_e2e(seed):
        ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(cfg), h_mel.ctypes.data_as(C.c_void_p), clips, frames,
                                                    None, seed, h_out.ctypes.data_as(C.c_void_p), args.chunk))

    # ---- device-resident timing
    for w in range(args.warmup):
        step_device(w)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = ctx.launch_count()
    hot_ms, hot_n, lead_ms, lead_n = 0.0, 0, 0.0, 0
    t_wall0 = time.time()
    ctx.timer_start()
    for s in range(args.steps):
        step_device(1000 + s)
        if args.per_kernel:
            ms, nl = ctx.last_hot_kernel_ms()     # waits for this step's last Griffin-Lim launch
            hot_ms += ms
            hot_n += nl
            ms, nl = ctx.last_lead_kernel_ms()
            lead_ms += ms
            lead_n += nl
    dev_ms = ctx.timer_stop()
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    g_n, g_rerun, g_max, _ = ctx.last_gl_guard()        # the last timed step's singular-bin guard record

    # ---- end-to-end timing through the host-buffer C ABI call
    for w in range(max(1, min(args.warmup, 2))):
        step_e2e(w)
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    for s in range(args.steps):
        step_e2e(2000 + s)
    e2e_ms_dev = ctx.timer_stop()
    ctx.sync()
    e2e_ms = max((time.perf_counter() - t0) * 1e3, e2e_ms_dev)
    barrier()
    checksum = float(np.abs(h_out[:: max(1, clips // 8), ::4099]).sum())

    # ---- the same end-to-end call with 16-bit PCM output (what cmd/towav writes): half the D2H bytes
    h_pcm, h_pcm_owner = ctx.pinned_array((clips, ola), np.int16)

    def step_pcm(seed):
        ctx.check(ctx.lib.gomel_from_mel_batch_host_pcm16(ctx.h, C.byref(cfg), h_mel.ctypes.data_as(C.c_void_p), clips, frames,
                                                          None, seed, h_pcm.ctypes.data_as(C.c_void_p), args.chunk))
    step_pcm(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        step_pcm(3000 + s)
    ctx.sync()
    pcm_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    dev_ms, e2e_ms, pcm_ms = reduce_max([dev_ms, e2e_ms, pcm_ms])

    # ---- the other two precision modes on the same batch, device-resident (one timed step each after one warm-up)
    modes = {}
    cfg_f64 = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=GL_ITERS, flags=_lib.FLAG_F64)
    for name, c, prec in ((("all_float32", cfg, (0, -1)), ("lead4_float64_then_float32", cfg, (4, -1)),
                           ("default_split_without_guard", cfg, None), ("all_float64", cfg_f64, None)) if args.modes else ()):
        prev = ctx.set_gl_precision(*prec) if prec else None
        prev_guard = ctx.set_gl_guard(0.0) if name != "all_float64" else None      # the bare splits
        step_device(1, c=c)
        barrier()
        ctx.timer_start()
        step_device(2, c=c)
        ms = ctx.timer_stop()
        if prev:
            ctx.set_gl_precision(*prev)
        if prev_guard is not None:
            ctx.set_gl_guard(prev_guard)
        barrier()
        ms = reduce_max([ms])[0]
        modes[name] = {"ms_per_step": ms, "value": world * clips * frames * HOP / SR / (ms / 1e3),
                       "misses_1e-4": {"all_float32": "3 % of 1,056 pairs (29 % at 100 iterations)",
                                       "lead4_float64_then_float32": "0.9 % of 1,056 pairs",
                                       "default_split_without_guard": "1 of 10,560 pairs (0 of the first 2,112)",
                                       "all_float64": "0"}[name]}

    # ---- strong scaling of configs[3]: 1024 clips in TOTAL, 1024 / N per rank (device-resident and end to end)
    strong = None
    if args.strong:
        sc = max(1, CLIPS_PER_GPU // world)
        if sc <= clips:
            for w in range(2):
                step_device(w, n_clips=sc)
            barrier()
            ctx.timer_start()
            for s in range(args.steps):
                step_device(4000 + s, n_clips=sc)
            s_ms = ctx.timer_stop()
            barrier()
            call = lambda sd: ctx.check(ctx.lib.gomel_from_mel_batch_host(
                ctx.h, C.byref(cfg), h_mel.ctypes.data_as(C.c_void_p), sc, frames, None, sd, h_out.ctypes.data_as(C.c_void_p),
                min(args.chunk, max(64, sc // 2))))
            call(0)
            barrier()
            t0 = time.perf_counter()
            for s in range(args.steps):
                call(5000 + s)
            ctx.sync()
            se_ms = (time.perf_counter() - t0) * 1e3
            barrier()
            s_ms, se_ms = reduce_max([s_ms, se_ms])
            tot = sc * world * frames * HOP / SR
            strong = {"workload": f"configs[3] strong scaling: {sc * world} clips in total, {sc} per GPU", "scaling": "strong",
                      "clips_per_gpu": sc, "value": tot * args.steps / (s_ms / 1e3), "ms_per_step": s_ms / args.steps,
                      "frame_iterations_per_s": sc * world * frames * GL_ITERS * args.steps / (s_ms / 1e3),
                      "e2e_value": tot * args.steps / (se_ms / 1e3), "e2e_ms_per_step": se_ms / args.steps, "unit": UNIT}

    # ---- ceiling of the end-to-end path: what the host link moves with every GPU copying at once
    link = None
    if args.link:
        try:
            link = host_link_probe(local_rank, use_dist)
        except Exception as e:       # noqa: BLE001  (diagnostic leg only)
            link = {"unavailable": repr(e)[:200]}

    # ---- configs[4]: one long clip split by time over the ranks (NCCL halo exchange), same process group
    ts = None
    if args.timesplit_seconds > 0:
        ts = timesplit_measure(ctx, cfg, rank, local_rank, world, use_dist, args.timesplit_seconds, max(1, min(args.steps, 3)), 1,
                               args.ts_tile, args.ts_edge, args.ts_overlap, args.ts_exchange)

        ts_phase = timesplit_phase_measure(ctx, rank, world, use_dist, args.timesplit_seconds, max(1, min(args.steps, 3)))
    else:
        ts_phase = None

    audio_s_per_step = world * clips * frames * HOP / SR          # seconds of audio covered by the frames
    value = audio_s_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = audio_s_per_step * args.steps / (e2e_ms / 1e3)
    frame_iters = clips * frames * GL_ITERS

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        roofline = None
        if hot_n:
            per_launch_s = hot_ms / 1e3 / hot_n
            achieved = BYTES_PER_FRAME_ITER * clips * frames / per_launch_s / 1e9
            traffic = lead_traffic = None
            tp = os.path.join(ROOT, "profiles", "gl_iter_traffic.json")
            if os.path.exists(tp):
                try:
                    tj = json.load(open(tp))
                    traffic = tj["dram_bytes_per_frame_iter"] * clips * frames
                    lead_traffic = tj["lead_kernel"]["dram_bytes_per_frame_iter"] * clips * frames
                except Exception:
                    pass
            f32_rec = {"bound": "hbm", "kernel": "k_gl_iter<5, 16, GUARD>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                       "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                       "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per frame-iteration of one --set full capture "
                                         "of this build (profiles/gl_iter_traffic.json) x the frame-iterations of one launch",
                       "algorithmic_bytes_per_launch": BYTES_PER_FRAME_ITER * clips * frames,
                       "avg_launch_ms": per_launch_s * 1e3, "launches_timed": hot_n,
                       "kernel_share_of_step": hot_ms / dev_ms,
                       "launch_note": "one timed launch = one Griffin-Lim iteration over the whole batch, issued as two "
                                      "concurrent half-batch launches of the kernel (clips split over two streams)",
                       "frame_iterations_per_s_per_gpu": clips * frames * hot_n / (hot_ms / 1e3)}
            roofline = f32_rec
            if lead_n:
                l_s = lead_ms / 1e3 / lead_n
                l_ach = BYTES_PER_FRAME_ITER_F64 * clips * frames / l_s / 1e9
                f64_rec = {
                    "bound": "hbm", "kernel": "k_gl_iter_f64<5>", "achieved": l_ach, "peak": peak, "unit": "GB/s", "frac": l_ach / peak,
                    "traffic": lead_traffic, "peak_source": peak_src, "traffic_source": f32_rec["traffic_source"],
                    "algorithmic_bytes_per_launch": BYTES_PER_FRAME_ITER_F64 * clips * frames,
                    "algorithmic_bytes_note": "36,872 B per frame-iteration: the items of SURVEY 8(d) (2049 magnitudes + 1280 samples in + "
                                              "1280 samples out) as float64, which is what these iterations must move; counted in the survey's "
                                              "float32 units (18,436 B) the fraction is half of `frac`",
                    "frac_in_float32_units": l_ach / peak / 2, "avg_launch_ms": l_s * 1e3,
                    "launches_timed": lead_n, "kernel_share_of_step": lead_ms / dev_ms,
                    "practical_bound": "FP64 pipe 57 % busy + LSU / shared-memory data pipe 53 % busy at 16 warps per SM (profiles/r02_gl_iter_f64_v3.md)",
                    "launch_note": f32_rec["launch_note"],
                    "frame_iterations_per_s_per_gpu": clips * frames * lead_n / (lead_ms / 1e3)}
                # `roofline` describes the kernel with the larger share of the step; the other one rides along
                if lead_ms >= hot_ms:
                    roofline = f64_rec
                    roofline["other_iteration_kernel"] = f32_rec
                else:
                    roofline["other_iteration_kernel"] = f64_rec
        cpu = None
        if world == 1:                              # rank 0 at N=1 only
            if not args.no_cpu:
                cores = os.cpu_count() or 1
                secs = 10.0 if cores >= 4 else 5.0
                v, fr, step_s = cpu_reference(cores, secs, 1, 0)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"{cores} clips x {secs:.0f} s ({fr} frames), GL-{GL_ITERS}, one clip per host thread, "
                                 f"{step_s:.1f} s wall; float64 C port of the Go path (no Go toolchain in the image)"}
        lead_it = lead_n // max(args.steps, 1) if args.per_kernel else None
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_mel.nbytes),
               "d2h_bytes_per_step": int(h_out.nbytes), "ms_per_step": e2e_ms / args.steps,
               "api": "gomel_from_mel_batch_host (pinned host float32 in/out, H2D / magnitudes / iterations / D2H on four streams, three chunk buffer sets)",
               "host_numa_node": numa_node, "cpu_share": cpu_share,
               "checksum": checksum,
               "pcm16": {"value": audio_s_per_step * args.steps / (pcm_ms / 1e3), "ms_per_step": pcm_ms / args.steps,
                         "d2h_bytes_per_step": int(h_pcm.nbytes),
                         "api": "gomel_from_mel_batch_host_pcm16 (int16 PCM out, the sample format of cmd/towav)"}}
        if link and "gbs_per_direction_per_gpu" in link:
            bw = link["gbs_per_direction_per_gpu"] * 1e9
            floor_ms = max(h_mel.nbytes, h_out.nbytes) / bw * 1e3
            e2e["link_floor_ms_per_step"] = floor_ms
            e2e["frac_of_link_ceiling"] = floor_ms / (e2e_ms / args.steps)
            e2e["compute_floor_ms_per_step"] = dev_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic", "config": workload_config(n_gpus),
            "precision": {"policy": "lead = max(16, iters - 16) Griffin-Lim iterations in float64 (k_gl_iter_f64), the rest in float32 (k_gl_iter); "
                                    "clips whose float32 iterations meet a near-singular bin (leverage > 5e4) have them re-run in float64 "
                                    "(profiles/r02_gl_guard.md: 44,351 of 44,352 pairs within 1e-4)",
                          "float64_iterations": lead_it, "float32_iterations": (GL_ITERS - lead_it) if lead_it is not None else None,
                          "guard": {"threshold": 5e4, "clips_seen_last_step": g_n, "clips_rerun_last_step": g_rerun, "max_leverage_last_step": g_max},
                          "other_modes_device_resident": modes},
            "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "host_link": link, "strong_1024": strong, "timesplit": ts, "timesplit_phase_istft": ts_phase,
            "stft_frames_per_s": frame_iters * world * args.steps / (dev_ms / 1e3),
            "stft_frames_per_s_note": "Griffin-Lim frame-iterations (one analysis STFT + one synthesis ISTFT each) per second, whole job",
        }
        if args.parity:
            line["parity"] = parity_record(ctx, cfg, _lib, base_mel, frames, ola)
    if args.stft and rank == 0 and line is not None:
        line["stft_side"] = bench_to_mel(ctx, cfg, _lib, args)
        # configs[3] also names 100 iterations: same batch, same call, GriffinLimIterations = 100
        cfg100 = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=100)
        call = lambda sd: ctx.check(ctx.lib.gomel_from_mel_dev(ctx.h, C.byref(cfg100), d_mel, clips, frames, None, sd, ola, d_out))
        call(1)
        ctx.sync()
        ctx.timer_start()
        call(2)
        ms100 = ctx.timer_stop()
        line["cufft_comparison"] = bench_cufft_comparison(clips, frames)
        # the drop-in single-clip call (configs[0] shape: one clip, host float64 buffers): 2 and 32 iterations
        mel1 = base_mel[0].reshape(-1, 2).astype(np.float64)
        init1 = np.random.default_rng(1).random(ola)
        single = {"workload": "gomel_from_mel, one 10 s clip, float64 host buffers, incl. copies (the call mel.FromMel / cmd/towav makes)"}
        for name, it_, fl in (("gl2", 2, 0), ("gl32", GL_ITERS, 0), ("gl32_all_float64", GL_ITERS, _lib.FLAG_F64)):
            c1 = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=it_, flags=fl)
            for _ in range(3):
                ctx.from_mel(c1, mel1, init=init1)
            t0 = time.perf_counter()
            for _ in range(5):
                ctx.from_mel(c1, mel1, init=init1)
            sec = (time.perf_counter() - t0) / 5
            lms, ln = ctx.last_lead_kernel_ms()
            hms, hn = ctx.last_hot_kernel_ms()
            single[name] = {"ms": sec * 1e3, "audio_s_per_s": frames * HOP / SR / sec,
                            "iteration_kernels_ms": lms + hms, "float64_iterations": ln, "float32_iterations": hn,
                            "note": "the rest of the call is the pageable float64 host copies (mel 1 MB + start signal 3.5 MB in, 3.5 MB out), "
                                    "magnitudes, conversions and one stream synchronisation"}
        # the same single clip through the pinned float32 call (start signal drawn on the device): what a caller that
        # keeps its own pinned buffers pays -- no pageable staging, 0.26 MB in, 1.75 MB out
        pinned = {"workload": "gomel_from_mel_batch_host, one 10 s clip, pinned float32 host buffers, start signal from the seed"}
        for name, it_ in (("gl2", 2), ("gl32", GL_ITERS)):
            c1 = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=it_)
            one = lambda sd: ctx.check(ctx.lib.gomel_from_mel_batch_host(ctx.h, C.byref(c1), h_mel.ctypes.data_as(C.c_void_p), 1, frames,
                                                                        None, sd, h_out.ctypes.data_as(C.c_void_p), 1))
            for w in range(3):
                one(w)
            t0 = time.perf_counter()
            for w in range(10):
                one(10 + w)
            sec = (time.perf_counter() - t0) / 10
            pinned[name] = {"ms": sec * 1e3, "audio_s_per_s": frames * HOP / SR / sec}
        single["pinned_float32_call"] = pinned
        line["single_clip_host_api"] = single
        line["gl100"] = {"workload": f"configs[3] with 100 iterations (84 float64 + 16 float32 under the policy), {clips} clips, device-resident",
                         "ms_per_step": ms100,
                         "audio_s_per_s": clips * frames * HOP / SR / (ms100 / 1e3),
                         "frame_iterations_per_s": clips * frames * 100 / (ms100 / 1e3)}
    ctx.dev_free(d_mel)
    ctx.dev_free(d_out)
    ctx.host_free(h_mel_owner)
    ctx.host_free(h_out_owner)
    ctx.host_free(h_pcm_owner)
    if rank == 0:
        print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0



# ------------------------------------------------------------------------------- config 5 side bench
def run_timesplit(args):
    """configs[4] on its own (`--workload timesplit`): prints the time-split record as its own JSON line."""
    rank, local_rank, world = dist_env()
    use_dist = world > 1
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gomel_b200 import _lib
    ctx = _lib.Context(local_rank)
    cfg = _lib.make_config(n_fft=N_FFT, hop=HOP, n_mels=N_MELS, n_freqs=768, gl_iters=GL_ITERS)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    rec = timesplit_measure(ctx, cfg, rank, local_rank, world, use_dist, args.seconds, args.steps, args.warmup,
                            args.ts_tile, args.ts_edge, args.ts_overlap, args.ts_exchange)
    if rank == 0:
        print(json.dumps(rec))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_cufft_comparison(clips, frames):
    """Comparison point only (BASELINE north_star): cuFFT batched R2C-4096 + C2R-4096 over the frames of ONE
    Griffin-Lim iteration (through torch.fft, which calls cuFFT), i.e. just the two transforms as separate
    library passes -- no framing, window, magnitude substitution or overlap-add.  Not on the product path."""
    try:
        import torch
        x = torch.rand((clips * frames, N_FFT), device="cuda", dtype=torch.float32)
        for _ in range(2):
            y = torch.fft.irfft(torch.fft.rfft(x, dim=1), n=N_FFT, dim=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            y = torch.fft.irfft(torch.fft.rfft(x, dim=1), n=N_FFT, dim=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        del x, y
        torch.cuda.empty_cache()
        return {"what": "cuFFT R2C-4096 + C2R-4096 (torch.fft.rfft/irfft), transforms only, separate passes",
                "frames": clips * frames, "ms": ms, "frame_pairs_of_transforms_per_s": clips * frames / (ms / 1e3)}
    except Exception as e:      # noqa: BLE001  (comparison point only)
        return {"unavailable": repr(e)[:200]}


def bench_to_mel(ctx, cfg, _lib, args):
    """configs[1] and configs[2]: batched ToMel, ToPhase, FromPhase on 256 synthetic 10 s clips,
    device-resident (CUDA events on the library stream), plus ToMel end to end from pinned host memory."""
    from util import synth_clip
    clips = 256
    n = int(round(CLIP_SECONDS * SR))
    npad, frames, ola = _lib.frames(cfg, n)
    stride = (npad + 3) & ~3
    nb = 16
    wav = np.zeros((clips, stride), np.float32)
    base = np.stack([synth_clip(500 + c, CLIP_SECONDS) for c in range(nb)]).astype(np.float32)
    for c in range(clips):
        wav[c, :n] = base[c % nb]
    nfq = 768
    d_sig = ctx.dev_malloc(wav.nbytes)
    d_mel = ctx.dev_malloc(clips * frames * N_MELS * 2 * 4)
    d_ph = ctx.dev_malloc(clips * frames * nfq * 2 * 4)
    d_wav = ctx.dev_malloc(clips * ola * 4)
    ctx.h2d(d_sig, wav)
    peak, _ = peaks()

    def timed(call, reps=20):
        for _ in range(3):
            call()
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            call()
        return ctx.timer_stop() / reps

    out = {"workload": "configs[1]/[2]: 256 x 10 s clips (87,552 frames), device-resident; smaller than L2, FP32-bound"}
    ms = timed(lambda: ctx.check(ctx.lib.gomel_to_mel_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_mel)))
    out["to_mel"] = {"frames_per_s": clips * frames / (ms / 1e3), "ms": ms,
                     "hbm_frac": 6656 * clips * frames / (ms / 1e3) / 1e9 / peak, "bytes_per_frame": 6656}
    ms = timed(lambda: ctx.check(ctx.lib.gomel_to_phase_dev(ctx.h, C.byref(cfg), d_sig, clips, stride, npad, frames, d_ph)))
    out["to_phase"] = {"frames_per_s": clips * frames / (ms / 1e3), "ms": ms,
                       "hbm_frac": 11264 * clips * frames / (ms / 1e3) / 1e9 / peak, "bytes_per_frame": 11264}
    ms = timed(lambda: ctx.check(ctx.lib.gomel_from_phase_dev(ctx.h, C.byref(cfg), d_ph, clips, frames, ola, d_wav)))
    out["from_phase"] = {"frames_per_s": clips * frames / (ms / 1e3), "ms": ms,
                         "hbm_frac": 11264 * clips * frames / (ms / 1e3) / 1e9 / peak, "bytes_per_frame": 11264}
    # end to end ToMel: pinned host waveforms in, mel out
    h_wav, own1 = ctx.pinned_array((clips, n), np.float32)
    h_wav[:] = wav[:, :n]
    h_mel, own2 = ctx.pinned_array((clips, frames * N_MELS * 2), np.float32)
    call = lambda: ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), h_wav.ctypes.data_as(C.c_void_p), clips, n,
                                                             h_mel.ctypes.data_as(C.c_void_p), 32))
    call()
    t0 = time.perf_counter()
    for _ in range(3):
        call()
    ms = (time.perf_counter() - t0) * 1e3 / 3
    out["to_mel_e2e"] = {"frames_per_s": clips * frames / (ms / 1e3), "ms": ms, "h2d_bytes": int(h_wav.nbytes),
                         "d2h_bytes": int(h_mel.nbytes), "api": "gomel_to_mel_batch_host (pinned float32)"}
    for p in (d_sig, d_mel, d_ph, d_wav):
        ctx.dev_free(p)
    ctx.host_free(own1)
    ctx.host_free(own2)
    out["newmel_defaults"] = bench_newmel_defaults(ctx, _lib, base, n, timed)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    return out


def bench_newmel_defaults(ctx, _lib, base, n, timed):
    """mel.NewMel() defaults (mel/mel.go:30-41): 160 mels, fmax 8000, Window 256, Resolut 2048, GL 2 iterations.
    Runs on the 4096-point core with zero-extended frames: 5x the frames per second of audio of the headline
    geometry, each costing a full core transform."""
    clips = 64
    dcfg = _lib.make_config(n_fft=2048, hop=256, n_mels=160, n_freqs=0, gl_iters=2)
    ctx.set_mel_tables(dcfg, 0.0, 8000.0)
    npad, frames, ola = _lib.frames(dcfg, n)
    stride = (npad + 3) & ~3
    wav = np.zeros((clips, stride), np.float32)
    for c in range(clips):
        wav[c, :n] = base[c % len(base)]
    d_sig = ctx.dev_malloc(wav.nbytes)
    d_mel = ctx.dev_malloc(clips * frames * 160 * 2 * 4)
    d_wav = ctx.dev_malloc(clips * ola * 4)
    ctx.h2d(d_sig, wav)
    res = {"workload": "%d x 10 s clips, %d frames each, device-resident" % (clips, frames)}
    ms = timed(lambda: ctx.check(ctx.lib.gomel_to_mel_dev(ctx.h, C.byref(dcfg), d_sig, clips, stride, npad, frames, d_mel)))
    res["to_mel"] = {"frames_per_s": clips * frames / (ms / 1e3), "audio_s_per_s": clips * CLIP_SECONDS / (ms / 1e3), "ms": ms}
    ms = timed(lambda: ctx.check(ctx.lib.gomel_from_mel_dev(ctx.h, C.byref(dcfg), d_mel, clips, frames, None, 1, ola, d_wav)),
               reps=10)
    res["from_mel_gl2"] = {"frame_iterations_per_s": 2 * clips * frames / (ms / 1e3),
                           "audio_s_per_s": clips * CLIP_SECONDS / (ms / 1e3), "ms": ms}
    for p in (d_sig, d_mel, d_wav):
        ctx.dev_free(p)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gomel_b200", choices=["gomel_b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU (default: the named workload)")
    ap.add_argument("--chunk", type=int, default=512, help="clips per pipeline chunk in the e2e call")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-stft", dest="stft", action="store_false", help="skip the ToMel side measurement")
    ap.add_argument("--no-per-kernel", dest="per_kernel", action="store_false")
    ap.add_argument("--no-numa", dest="numa", action="store_false", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--tile", type=int, default=0, help="frames per tile (0 = library heuristic)")
    ap.add_argument("--no-modes", dest="modes", action="store_false", help="skip the all-float32 / all-float64 comparison steps")
    ap.add_argument("--no-strong", dest="strong", action="store_false", help="skip the 1024-clips-in-total strong-scaling leg")
    ap.add_argument("--no-link", dest="link", action="store_false", help="skip the host-link ceiling probe")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the parity sub-record")
    ap.add_argument("--timesplit-seconds", type=float, default=3600.0,
                    help="length of the single clip of the configs[4] sub-record (0 = skip)")
    ap.add_argument("--workload", default="clips", choices=["clips", "timesplit"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="timesplit: clip length")
    ap.add_argument("--ts-tile", type=int, default=0, help="timesplit: frames per tile (0 = fill whole waves)")
    ap.add_argument("--ts-edge", type=int, default=8, help="timesplit: frames in the tiles next to a rank boundary (0 = uniform)")
    ap.add_argument("--no-ts-overlap", dest="ts_overlap", action="store_false")
    ap.add_argument("--ts-exchange", default="native", choices=["native", "torch"],
                    help="timesplit: NCCL called by the library (dlopen) or through torch.distributed P2P ops")
    args = ap.parse_args()
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)          # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "timesplit":
        return run_timesplit(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
