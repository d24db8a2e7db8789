#!/usr/bin/env python3
"""Multi-GPU check of the time-split path with the real NCCL exchange; launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_timesplit_nccl.py

Every rank runs its slice with NcclExchange; rank 0 gathers the owned samples, runs the same clip
unsplit on its own GPU with the same tile size and requires BIT-IDENTICAL output, then checks the
oracle tolerance."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gomel_b200 import _lib, timesplit
    from oracle import oracle as O
    from util import rel_l2, synth_clip
    iters, tile = 18, 8          # crosses the float64 -> float32 hand-over of the precision policy (16 lead iterations)
    ctx = _lib.Context(local)
    cfg = _lib.make_config(gl_iters=iters)
    ctx.set_mel_tables(cfg, 0.0, 16000.0)
    wav = synth_clip(123, 6.0)
    mel = O.to_mel(O.config(), wav).astype(np.float32)
    frames = len(mel) // 192
    ola = 4096 + (frames - 1) * 1280
    init = np.random.default_rng(9).random(ola).astype(np.float32)
    ok = True
    for overlap, native, edge in ((False, False, 0), (True, False, 0), (True, True, 0), (False, True, 0), (True, True, 4)):
        s = timesplit.Session(ctx, cfg, frames, rank, world, tile if not edge else 12, edge)
        s.load(mel[s.frame_begin * 192:(s.frame_begin + s.n_frames) * 192], init[s.sample_begin:s.sample_begin + s.n_samples])
        if native:          # library-owned NCCL communicator, whole loop enqueued by one C call
            timesplit.NativeNccl(s).run(0, iters, overlap=overlap)
        else:               # torch.distributed P2P ops on the library's communication stream
            timesplit.run(s, iters, timesplit.NcclExchange(s), overlap=overlap)
        mine = torch.from_numpy(np.ascontiguousarray(s.owned(s.finish(iters)))).cuda()
        lens = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(lens, torch.tensor([mine.numel()], dtype=torch.int64, device="cuda"))
        s.close()
        if rank == 0:
            pieces = [mine.cpu().numpy()]
            for r in range(1, world):
                buf = torch.zeros(int(lens[r]), dtype=torch.float32, device="cuda")
                dist.recv(buf, r)
                pieces.append(buf.cpu().numpy())
            split = np.concatenate(pieces)
            ctx.set_tile_frames(tile)
            whole = ctx.from_mel(cfg, mel.astype(np.float64), init=init.astype(np.float64))
            ctx.set_tile_frames(0)
            ref = O.from_mel(O.config(gl_iters=iters), mel.astype(np.float64), init.astype(np.float64))
            # uniform tiles share their boundaries with the unsplit run -> identical bits; short boundary tiles
            # change the order of the partial sums -> equal to rounding
            same = np.array_equal(split, whole.astype(np.float32)) if not edge else rel_l2(split, whole) < 2e-6
            err = rel_l2(split, ref)
            print(f"timesplit world={world} overlap={overlap} native_nccl={native} edge={edge}: bit-identical to unsplit = {same}, rel-L2 vs oracle = {err:.3e}")
            ok = ok and same and err < 1e-4 and len(split) == ola
        else:
            dist.send(mine, 0)
        dist.barrier()
    # phase.ISTFT of one clip split by time: one NCCL transfer per boundary (library-owned communicator)
    ocfg = O.config(num_freqs=768)
    pcfg = _lib.make_config(n_fft=4096, hop=1280, n_mels=0, n_freqs=768, gl_iters=0)
    spec = O.to_phase(ocfg, wav).astype(np.float32)
    s = timesplit.Session(ctx, pcfg, frames, rank, world, tile, 0)
    timesplit.NativeNccl(s)
    mine = torch.from_numpy(np.ascontiguousarray(s.owned(timesplit.phase_istft_nccl(
        s, spec[s.frame_begin * 768:(s.frame_begin + s.n_frames) * 768])))).cuda()
    lens = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(lens, torch.tensor([mine.numel()], dtype=torch.int64, device="cuda"))
    s.close()
    if rank == 0:
        pieces = [mine.cpu().numpy()]
        for r in range(1, world):
            buf = torch.zeros(int(lens[r]), dtype=torch.float32, device="cuda")
            dist.recv(buf, r)
            pieces.append(buf.cpu().numpy())
        split = np.concatenate(pieces)
        ctx.set_tile_frames(tile)
        whole = ctx.from_phase(pcfg, spec.astype(np.float64))
        ctx.set_tile_frames(0)
        same = np.array_equal(split, whole.astype(np.float32))
        err = rel_l2(split, O.from_phase(ocfg, spec.astype(np.float64)))
        print(f"timesplit phase.ISTFT world={world}: bit-identical to unsplit = {same}, rel-L2 vs oracle = {err:.3e}")
        ok = ok and same and err < 1e-5 and len(split) == ola
    else:
        dist.send(mine, 0)
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("TIMESPLIT_NCCL_OK" if ok else "TIMESPLIT_NCCL_FAIL")
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
