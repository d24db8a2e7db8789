"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: clip sharding (no collective on the
data path) and the time-split halo protocol (two 2816-float partials per boundary per iteration)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gomel_b200 import shard, timesplit
        from oracle import oracle_np as ONP
        from ts_protocol_np import HALO, RankModel
        from util import synth_clip
        # ---- clip sharding: every clip exactly once, balanced, no data-path collective needed
        n_clips = 1024 + 3
        mine = shard.clip_range(n_clips, rank, world)
        cnt = torch.tensor([mine.stop - mine.start], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, cnt)
        sizes = [int(g) for g in gathered]
        assert sum(sizes) == n_clips and max(sizes) - min(sizes) <= 1
        assert mine.start == sum(sizes[:rank])
        # ---- time-split protocol on a 2.5 s clip, 3 iterations
        iters = 3
        wav = synth_clip(9, 2.5)
        mel = ONP.to_mel(wav)
        M = ONP.gl_magnitudes(mel)
        frames = M.shape[0]
        ola = 4096 + (frames - 1) * 1280
        init = np.random.default_rng(4).random(ola)
        parts = timesplit.partition(frames, world, tile_frames=8)
        fb, nf = parts[rank]
        assert sum(p[1] for p in parts) == frames and parts[0][0] == 0
        assert all(p[0] % 8 == 0 for p in parts)
        n_samples = nf * 1280 + HALO
        rm = RankModel(M[fb:fb + nf], init[fb * 1280:fb * 1280 + n_samples], fb, nf, rank > 0, rank + 1 < world)
        for _ in range(iters):
            tail, head = rm.iterate()
            t_prev = h_next = None
            reqs = []
            if rank + 1 < world:
                h_next_t = torch.zeros(HALO, dtype=torch.float64)
                reqs.append(dist.isend(torch.from_numpy(tail), rank + 1))
                reqs.append(dist.irecv(h_next_t, rank + 1))
            if rank > 0:
                t_prev_t = torch.zeros(HALO, dtype=torch.float64)
                reqs.append(dist.isend(torch.from_numpy(head), rank - 1))
                reqs.append(dist.irecv(t_prev_t, rank - 1))
            for r in reqs:
                r.wait()
            if rank + 1 < world:
                h_next = h_next_t.numpy()
            if rank > 0:
                t_prev = t_prev_t.numpy()
            rm.absorb(t_prev, h_next)
        mine_sig = torch.from_numpy(np.ascontiguousarray(rm.owned()))
        lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(lens, torch.tensor([len(mine_sig)], dtype=torch.int64))
        if rank == 0:
            pieces = [mine_sig.numpy()]
            for r in range(1, world):
                buf = torch.zeros(int(lens[r]), dtype=torch.float64)
                dist.recv(buf, r)
                pieces.append(buf.numpy())
            stitched = np.concatenate(pieces)
            ref = ONP.griffin_lim(M, init, iters)
            err = float(np.linalg.norm(stitched - ref) / np.linalg.norm(ref))
            q.put(("ok", err, len(stitched), ola))
        else:
            dist.send(mine_sig, 0)
    except Exception as e:      # noqa: BLE001
        q.put(("fail", repr(e), 0, 0))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_halo_protocol():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, err, n, ola = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", err
    assert n == ola
    assert err < 1e-12          # the split is exact up to float64 rounding of two-term sums


def _phase_worker(rank, world, port, q):
    """phase.ISTFT of one clip split by time (third case) as a NumPy rank model over gloo: every rank
    inverts its own frames, ONE partial of 2816 samples per boundary travels tail -> next rank, the owner adds it and
    applies the window-sum gain indexed by the GLOBAL sample position (phase/phase.go:93-133)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gomel_b200 import timesplit
        from oracle import oracle_np as ONP
        from util import synth_clip
        N, H, HALO, nfq = 4096, 1280, 4096 - 1280, 768
        spec = ONP.to_phase(synth_clip(12, 2.2), nfq)
        frames = len(spec) // nfq
        ola = N + (frames - 1) * H
        fb, nf = timesplit.partition(frames, world, tile_frames=6)[rank]
        s = spec.reshape(frames, nfq, 2)[fb:fb + nf]
        X = np.zeros((nf, N // 2 + 1), np.complex128)
        X[:, 1:nfq + 1] = s[:, :, 1] + 1j * s[:, :, 0]
        X[:, nfq + 1:] = X[:, nfq:nfq + 1]
        X[:, N // 2] = X[:, N // 2].real
        w = np.hanning(N)
        y = np.fft.irfft(X, n=N, axis=1) * w
        local = np.zeros(nf * H + HALO)
        for f in range(nf):
            local[f * H:f * H + N] += y[f]
        reqs = []
        recv = torch.zeros(HALO, dtype=torch.float64)
        if rank + 1 < world:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(local[nf * H:])), rank + 1))
        if rank > 0:
            reqs.append(dist.irecv(recv, rank - 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            local[:HALO] += recv.numpy()
        # window-sum gain from the GLOBAL position: data independent, every rank rebuilds it from the frame count
        ws = np.zeros(ola)
        for f in range(frames):
            ws[f * H:f * H + N] += w * w
        thr = ws.max() * 0.5
        g = ws[fb * H:fb * H + len(local)]
        out = local.copy()
        hi, mid = g > thr, (g <= thr) & (g > 1e-21)
        out[hi] /= g[hi]
        out[mid] = out[mid] / g[mid] * (g[mid] / thr)
        own = out[:nf * H + (HALO if rank == world - 1 else 0)]
        mine = torch.from_numpy(np.ascontiguousarray(own))
        lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(lens, torch.tensor([len(mine)], dtype=torch.int64))
        if rank == 0:
            pieces = [mine.numpy()]
            for r in range(1, world):
                buf = torch.zeros(int(lens[r]), dtype=torch.float64)
                dist.recv(buf, r)
                pieces.append(buf.numpy())
            stitched = np.concatenate(pieces)
            ref = ONP.from_phase(spec, nfq)
            q.put(("ok", float(np.linalg.norm(stitched - ref) / np.linalg.norm(ref)), len(stitched), ola))
        else:
            dist.send(mine, 0)
    except Exception as e:      # noqa: BLE001
        q.put(("fail", repr(e), 0, 0))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_phase_istft_single_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_phase_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, err, n, ola = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", err
    assert n == ola and err < 1e-12


def test_partition_properties():
    from gomel_b200 import shard, timesplit
    for frames, world, T in ((124029, 8, 16), (342, 2, 8), (342, 8, 16), (1000, 3, 54)):
        parts = timesplit.partition(frames, world, T)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == frames
        for (b0, n0), (b1, _) in zip(parts, parts[1:]):
            assert b0 + n0 == b1 and n0 % 2 == 0 and n0 >= 4       # all but the last range hold whole pairs
    with pytest.raises(ValueError):
        timesplit.partition(20, 8, 16)
    for n, w in ((1024, 8), (1027, 4), (5, 8)):
        ranges = [shard.clip_range(n, r, w) for r in range(w)]
        assert ranges[0].start == 0 and ranges[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(ranges, ranges[1:]))
