"""NumPy model of ONE RANK of the time-split Griffin-Lim protocol (test infrastructure, built on the
oracle's NumPy restatement).  It mirrors what the CUDA session does at the level that matters for
the multi-rank logic: which frames a rank owns, which 2816-sample partial sums it sends/receives
per iteration, and how the final signal is stitched.  Used by the world_size-2 gloo test."""
import numpy as np

N, H = 4096, 1280
HALO = N - H


class RankModel:
    def __init__(self, M_local, init_local, frame_begin, n_frames, has_prev, has_next):
        self.M = np.abs(M_local)                 # (n_frames, 2049) target magnitudes of this rank's frames
        self.F = n_frames
        self.has_prev, self.has_next = has_prev, has_next
        self.n_samples = n_frames * H + HALO
        self.sig = np.array(init_local, np.float64)          # complete samples of the local range
        assert len(self.sig) == self.n_samples
        self.w = np.hanning(N)
        self.idx = np.arange(N)[None, :] + np.arange(n_frames)[:, None] * H

    def iterate(self):
        """one Jacobi pass over the local frames; returns (tail_partial, head_partial) to send"""
        X = np.fft.rfft(self.sig[self.idx] * self.w, axis=1)
        mag = np.abs(X)
        unit = np.where(mag > 0, X / np.where(mag > 0, mag, 1), 1.0)
        y = np.fft.irfft(self.M * unit, n=N, axis=1) * self.w
        new = np.zeros(self.n_samples)
        for f in range(self.F):
            new[f * H:f * H + N] += y[f]
        self.partial = new
        tail = new[self.F * H:].copy() if self.has_next else None        # -> next rank
        head = new[:HALO].copy() if self.has_prev else None              # -> previous rank
        return tail, head

    def absorb(self, tail_from_prev, head_from_next):
        """local + received for the two shared regions; afterwards self.sig is complete again"""
        new = self.partial
        if self.has_prev:
            new[:HALO] += tail_from_prev
        if self.has_next:
            new[self.F * H:] += head_from_next
        self.sig = new

    def owned(self):
        return self.sig[:self.F * H + (0 if self.has_next else HALO)]
