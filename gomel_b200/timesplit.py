"""One long clip split by time across GPUs (BASELINE config 5).

Each rank owns a contiguous, tile-aligned range of the clip's frames and keeps its slice of the
signal and of the target magnitudes in HBM for all Griffin-Lim iterations.  Per iteration each rank
boundary exchanges two partial sums of Resolut-Window = 2816 floats (11,264 B each way): the
earlier rank's tail partial and the later rank's head partial; both sides add local + received.

The CUDA side is the `gomel_ts_*` part of the C ABI; this module only sequences it and performs the
exchange -- `NcclExchange` uses torch.distributed (NCCL over NVLink) send/recv on the library's
communication stream, `LocalExchange` copies device-to-device between several sessions of one
process (used to test the multi-rank path on a single GPU)."""
import ctypes as C

import numpy as np

from . import _lib

HALO = 4096 - 1280          # Resolut - Window


def partition(n_frames_total, world, tile_frames=16):
    """[(frame_begin, n_frames_local)] per rank -- the same arithmetic as gomel_ts_create."""
    T = tile_frames if tile_frames > 0 else 16
    T += T & 1
    T = max(T, 4)
    tiles_total = (n_frames_total + T - 1) // T
    if tiles_total < world:
        raise ValueError("fewer tiles than ranks: lower tile_frames")
    out = []
    for r in range(world):
        a, b = r * tiles_total // world, (r + 1) * tiles_total // world
        out.append((a * T, min(b * T, n_frames_total) - a * T))
    return out


class Session:
    """gomel_ts: this rank's slice of one clip."""

    def __init__(self, ctx, cfg, n_frames_total, rank, world, tile_frames=16, edge_frames=0):
        self.ctx, self.cfg, self.rank, self.world = ctx, cfg, rank, world
        h = C.c_void_p()
        ctx.check(ctx.lib.gomel_ts_create2(ctx.h, C.byref(cfg), n_frames_total, rank, world, tile_frames, edge_frames,
                                           C.byref(h)))
        self.h = h
        a, b, c, d = C.c_long(), C.c_long(), C.c_long(), C.c_long()
        ctx.check(ctx.lib.gomel_ts_range(h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.frame_begin, self.n_frames, self.sample_begin, self.n_samples = a.value, b.value, c.value, d.value
        self.comm_stream = ctx.lib.gomel_ts_comm_stream(h)

    def close(self):
        if self.h:
            self.ctx.lib.gomel_ts_destroy(self.h)
            self.h = None

    def load(self, mel_local, init_local=None, seed=0):
        """mel_local: (n_frames*n_mels, 2) float32 host array of THIS rank's frames; init_local:
        n_samples floats (global samples [sample_begin, sample_begin+n_samples)) or None."""
        ctx = self.ctx
        mel_local = np.ascontiguousarray(mel_local, np.float32)
        d_mel = ctx.dev_malloc(mel_local.nbytes)
        ctx.h2d(d_mel, mel_local)
        d_init = None
        if init_local is not None:
            init_local = np.ascontiguousarray(init_local, np.float32)
            assert len(init_local) == self.n_samples
            d_init = ctx.dev_malloc(init_local.nbytes)
            ctx.h2d(d_init, init_local)
        self.load_dev(d_mel, d_init, seed)
        ctx.dev_free(d_mel)
        if d_init is not None:
            ctx.dev_free(d_init)

    def load_dev(self, d_mel_local, d_init_local=None, seed=0):
        """as load(), from device-resident float32 buffers: computes this rank's target magnitudes (both precisions
        the policy needs) and the start signal -- the once-per-FromMel part of the work"""
        self.ctx.check(self.ctx.lib.gomel_ts_load(self.h, d_mel_local, d_init_local, seed))

    def lead_iters(self):
        return int(self.ctx.lib.gomel_ts_lead_iters(self.h))

    def iterate(self, it, part=0):
        self.ctx.check(self.ctx.lib.gomel_ts_iterate(self.h, it, part))

    def halo_ptrs(self, it):
        p = [C.c_void_p() for _ in range(4)]
        self.ctx.check(self.ctx.lib.gomel_ts_halo_ptrs(self.h, it, *[C.byref(x) for x in p]))
        return dict(send_tail=p[0].value, send_head=p[1].value, recv_tail=p[2].value, recv_head=p[3].value)

    def halo_elem_bytes(self, it):
        """8 while iteration `it` is one of the float64 lead iterations (the halo pointers then address doubles), else 4"""
        return int(self.ctx.lib.gomel_ts_halo_elem_bytes(self.h, it))

    def comm_begin(self, it):
        self.ctx.check(self.ctx.lib.gomel_ts_comm_begin(self.h, it))

    def comm_end(self, it):
        self.ctx.check(self.ctx.lib.gomel_ts_comm_end(self.h, it))

    def sync(self):
        self.ctx.check(self.ctx.lib.gomel_ts_sync(self.h))

    def finish(self, iters):
        """-> this rank's local signal (n_samples float32); samples [0, n_frames*Window) are owned by
        this rank (the last rank also owns the final 2816)."""
        ctx = self.ctx
        d_out = ctx.dev_malloc(self.n_samples * 4)
        ctx.check(ctx.lib.gomel_ts_finish(self.h, iters, d_out))
        out = np.empty(self.n_samples, np.float32)
        ctx.d2h(out, d_out)
        ctx.dev_free(d_out)
        return out

    def owned(self, local_signal):
        n = self.n_frames * self.cfg.hop + (HALO if self.rank == self.world - 1 else 0)
        return local_signal[:n]


class _DevArray:
    """zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n, elem_bytes=4):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8" if elem_bytes == 8 else "<f4",
                                         "data": (int(ptr), False), "version": 2}


class NcclExchange:
    """Halo exchange with torch.distributed P2P ops (NCCL over NVLink), issued on the library's
    communication stream so that no host synchronisation is needed between iterations."""

    def __init__(self, session, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.s, self.group = torch, dist, session, group
        self.stream = torch.cuda.ExternalStream(session.comm_stream, device=torch.device("cuda", session.ctx.device))
        self._cache = {}

    def _ops(self, it):
        """P2P op list for iteration `it`: the four halo pointers alternate between two buffer sets, and the float64
        lead iterations have buffer sets (and an element type) of their own"""
        eb = self.s.halo_elem_bytes(it)
        parity = (it & 1, eb)
        if parity not in self._cache:
            torch, dist, s = self.torch, self.dist, self.s
            p = s.halo_ptrs(it)
            dev = torch.device("cuda", s.ctx.device)
            t = lambda ptr: torch.as_tensor(_DevArray(ptr, HALO, eb), device=dev)
            ops = []
            if p["send_head"]:          # previous rank exists
                ops.append(dist.P2POp(dist.isend, t(p["send_head"]), s.rank - 1, self.group))
                ops.append(dist.P2POp(dist.irecv, t(p["recv_tail"]), s.rank - 1, self.group))
            if p["send_tail"]:          # next rank exists
                ops.append(dist.P2POp(dist.isend, t(p["send_tail"]), s.rank + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, t(p["recv_head"]), s.rank + 1, self.group))
            self._cache[parity] = ops
        return self._cache[parity]

    def __call__(self, it):
        s = self.s
        ops = self._ops(it)
        s.comm_begin(it)
        if ops:
            with self.torch.cuda.stream(self.stream):
                for r in self.dist.batch_isend_irecv(ops):
                    r.wait()            # stream-level wait: the comm stream waits for NCCL, the host does not
        s.comm_end(it)


class NativeNccl:
    """In-library exchange: the library dlopen()s NCCL, builds its own communicator from an id that is
    distributed here with torch.distributed (any backend), and runs whole iterations without returning to
    Python (gomel_ts_run_nccl)."""

    def __init__(self, session, group=None):
        import torch
        import torch.distributed as dist
        s = self.s = session
        ident = C.create_string_buffer(128)
        if s.world > 1:
            if s.rank == 0:
                s.ctx.check(s.ctx.lib.gomel_nccl_unique_id(s.ctx.h, ident))
            dev = torch.device("cuda", s.ctx.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.tensor(list(ident.raw), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0, group=group)
            ident = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
            s.ctx.check(s.ctx.lib.gomel_ts_nccl_init(s.h, ident))

    def run(self, first_iter, n_iters, overlap=True):
        self.s.ctx.check(self.s.ctx.lib.gomel_ts_run_nccl(self.s.h, first_iter, n_iters, int(overlap)))


# ---------------------------------------------------------------- phase.ISTFT of a long clip (third case)
def phase_istft_local(ctx, cfg, spec, world, tile_frames=16):
    """FromPhase of ONE clip with the frames split over `world` emulated ranks (sessions of one process, one GPU):
    every rank inverts its frames, the tail partial of rank r is copied into the head region of rank r+1 (what
    NCCL does between GPUs), the owner adds it and applies the window-sum gain.  Returns the stitched float32 signal."""
    spec = np.ascontiguousarray(spec, np.float32).reshape(-1, 2)
    n_frames = len(spec) // cfg.n_freqs
    sessions = [Session(ctx, cfg, n_frames, r, world, tile_frames, 0) for r in range(world)]
    bufs = []
    try:
        for s in sessions:
            part = np.ascontiguousarray(spec[s.frame_begin * cfg.n_freqs:(s.frame_begin + s.n_frames) * cfg.n_freqs])
            d = ctx.dev_malloc(part.nbytes)
            ctx.h2d(d, part)
            bufs.append(d)
            ctx.check(ctx.lib.gomel_ts_phase_istft(s.h, d))
        for s in sessions:
            s.sync()
        ptrs = []
        for s in sessions:
            a, b = C.c_void_p(), C.c_void_p()
            ctx.check(ctx.lib.gomel_ts_phase_halo_ptrs(s.h, C.byref(a), C.byref(b)))
            ptrs.append((a.value, b.value))
        for s in sessions:
            s.comm_begin(0)
        for r in range(world - 1):
            ctx.check(ctx.lib.gomel_copy_d2d(ctx.h, ptrs[r + 1][1], ptrs[r][0], HALO * 4, sessions[r + 1].comm_stream))
        for s in sessions:
            s.comm_end(0)
        outs = []
        for s in sessions:
            d_out = ctx.dev_malloc(s.n_samples * 4)
            ctx.check(ctx.lib.gomel_ts_phase_finish(s.h, d_out))
            o = np.empty(s.n_samples, np.float32)
            ctx.d2h(o, d_out)
            ctx.dev_free(d_out)
            outs.append(s.owned(o))
        return np.concatenate(outs)
    finally:
        for d in bufs:
            ctx.dev_free(d)
        for s in sessions:
            s.close()


def phase_istft_nccl(session, spec_local):
    """one rank of the real multi-GPU form: library-owned NCCL (NativeNccl(session) must have been built).
    spec_local: this rank's (n_frames*n_freqs, 2) float32 rows; returns this rank's local signal (see Session.owned)."""
    ctx = session.ctx
    spec_local = np.ascontiguousarray(spec_local, np.float32)
    d, d_out = ctx.dev_malloc(spec_local.nbytes), ctx.dev_malloc(session.n_samples * 4)
    ctx.h2d(d, spec_local)
    ctx.check(ctx.lib.gomel_ts_phase_run_nccl(session.h, d, d_out))
    out = np.empty(session.n_samples, np.float32)
    ctx.d2h(out, d_out)
    ctx.dev_free(d)
    ctx.dev_free(d_out)
    return out


def run(session, iters, exchange, overlap=True):
    """All Griffin-Lim iterations of one rank.  overlap=True launches the boundary tiles first, the
    exchange on the communication stream, and the interior tiles concurrently with it."""
    for it in range(iters):
        if overlap:
            session.iterate(it, 1)
            exchange(it)
            session.iterate(it, 2)
        else:
            session.iterate(it, 0)
            exchange(it)


def run_local(ctx, cfg, mel, init, iters, world, tile_frames=16, overlap=False, edge_frames=0):
    """`world` ranks emulated as `world` sessions of ONE process / ONE GPU; the exchange is a
    device-to-device copy.  Returns the stitched float32 signal.  (Multi-rank test on one GPU.)"""
    mel = np.ascontiguousarray(mel, np.float32).reshape(-1, 2)
    n_frames = len(mel) // cfg.n_mels
    sessions = [Session(ctx, cfg, n_frames, r, world, tile_frames, edge_frames) for r in range(world)]
    try:
        for s in sessions:
            m = mel[s.frame_begin * cfg.n_mels:(s.frame_begin + s.n_frames) * cfg.n_mels]
            s.load(m, None if init is None else init[s.sample_begin:s.sample_begin + s.n_samples])
        for it in range(iters):
            for s in sessions:
                s.iterate(it, 1 if overlap else 0)
            # every session's boundary tiles (their own st_edge stream) must have finished before ANOTHER session's
            # communication stream reads their partials: comm_begin only orders a session against its own tiles
            for s in sessions:
                s.sync()
            ptrs = [s.halo_ptrs(it) for s in sessions]
            nbytes = HALO * sessions[0].halo_elem_bytes(it)
            for s in sessions:
                s.comm_begin(it)
            for r in range(world - 1):      # boundary between rank r and r+1
                ctx.check(ctx.lib.gomel_copy_d2d(ctx.h, ptrs[r + 1]["recv_tail"], ptrs[r]["send_tail"], nbytes,
                                                 sessions[r + 1].comm_stream))
                ctx.check(ctx.lib.gomel_copy_d2d(ctx.h, ptrs[r]["recv_head"], ptrs[r + 1]["send_head"], nbytes,
                                                 sessions[r].comm_stream))
            for s in sessions:
                s.comm_end(it)
            if overlap:
                for s in sessions:
                    s.iterate(it, 2)
        out = np.concatenate([s.owned(s.finish(iters)) for s in sessions])
    finally:
        for s in sessions:
            s.close()
    return out
