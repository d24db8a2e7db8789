"""Minimal native FLAC reader (and a small writer for fixtures) -- host-side file plumbing of the ingest path.

The reference decodes FLAC with mewkiz/flac (Go: mel/impl.go:266-296, phase/impl.go:351-381) and with
soundfile (Python: phase.py:572-589); neither is in the image, so the container is decoded here: STREAMINFO,
frame headers, CONSTANT / VERBATIM / FIXED / LPC subframes, Rice and Rice2 residuals with escape partitions,
wasted bits, left-side / right-side / mid-side decorrelation, 8..32 bits per sample.  CRCs are skipped.
The two callers' scalings and channel rules live in codec.py (`load_flac_go`, `load_flac_sf`).
"""
import struct

import numpy as np


class FlacError(ValueError):
    pass


class _Bits:
    """MSB-first bit reader over a bytes object"""

    def __init__(self, data, pos=0):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0      # acc holds n unread bits

    def _fill(self, need):
        while self.n < need:
            if self.p >= len(self.d):
                raise EOFError
            take = self.d[self.p:self.p + 8]
            self.acc = (self.acc << (8 * len(take))) | int.from_bytes(take, "big")
            self.n += 8 * len(take)
            self.p += len(take)

    def u(self, k):
        if k == 0:
            return 0
        self._fill(k)
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def s(self, k):
        v = self.u(k)
        return v - (1 << k) if k and (v >> (k - 1)) else v

    def unary(self):
        """number of 0 bits before the next 1 bit"""
        q = 0
        while True:
            if self.n == 0:
                self._fill(1)
            if self.acc == 0:
                q += self.n
                self.n = 0
                continue
            lead = self.n - self.acc.bit_length()
            q += lead
            self.n -= lead + 1
            self.acc &= (1 << self.n) - 1
            return q

    def align(self):
        self.n -= self.n % 8
        self.acc &= (1 << self.n) - 1

    def byte_pos(self):
        return self.p - self.n // 8


def _utf8_number(b):
    x = b.u(8)
    if x < 0x80:
        return x
    n = 0
    while x & (0x80 >> n):
        n += 1
    v = x & ((1 << (7 - n)) - 1)
    for _ in range(n - 1):
        v = (v << 6) | (b.u(8) & 0x3F)
    return v


_FIXED = ((), (1,), (2, -1), (3, -3, 1), (4, -6, 4, -1))


def _residual(b, blocksize, order, out):
    method = b.u(2)
    if method > 1:
        raise FlacError("reserved residual coding method")
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    porder = b.u(4)
    nparts = 1 << porder
    i = order
    for part in range(nparts):
        cnt = (blocksize >> porder) - (order if part == 0 else 0)
        k = b.u(pbits)
        if k == esc:
            w = b.u(5)
            for _ in range(cnt):
                out[i] = b.s(w)
                i += 1
        else:
            for _ in range(cnt):
                v = (b.unary() << k) | b.u(k)
                out[i] = (v >> 1) ^ -(v & 1)
                i += 1


def _subframe(b, blocksize, bps):
    if b.u(1):
        raise FlacError("subframe padding bit set")
    typ = b.u(6)
    wasted = 0
    if b.u(1):
        wasted = b.unary() + 1
        bps -= wasted
    if typ == 0:
        out = [b.s(bps)] * blocksize
    elif typ == 1:
        out = [b.s(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:
        order = typ - 8
        out = [0] * blocksize
        for i in range(order):
            out[i] = b.s(bps)
        _residual(b, blocksize, order, out)
        c = _FIXED[order]
        for i in range(order, blocksize):
            acc = out[i]
            for j, cj in enumerate(c):
                acc += cj * out[i - 1 - j]
            out[i] = acc
    elif typ >= 32:
        order = (typ & 31) + 1
        out = [0] * blocksize
        for i in range(order):
            out[i] = b.s(bps)
        prec = b.u(4) + 1
        if prec == 16:
            raise FlacError("invalid LPC precision")
        shift = b.s(5)
        coef = [b.s(prec) for _ in range(order)]
        _residual(b, blocksize, order, out)
        for i in range(order, blocksize):
            acc = 0
            for j in range(order):
                acc += coef[j] * out[i - 1 - j]
            out[i] += acc >> shift
    else:
        raise FlacError("reserved subframe type")
    if wasted:
        out = [v << wasted for v in out]
    return out


def decode(path):
    """-> (blocks, sample_rate, bits_per_sample, channels): `blocks` is a list of frames, each a list of per-channel
    int64 arrays (left/right restored) -- the shape mewkiz/flac hands to the reference (frame.Subframes[i].Samples)."""
    data = open(path, "rb").read()
    if data[:4] != b"fLaC":
        raise FlacError("not a FLAC stream")
    pos, info = 4, None
    while True:
        hdr = data[pos]
        ln = int.from_bytes(data[pos + 1:pos + 4], "big")
        body = data[pos + 4:pos + 4 + ln]
        if hdr & 0x7F == 0:
            x = int.from_bytes(body[10:18], "big")
            info = {"sr": x >> 44, "ch": ((x >> 41) & 7) + 1, "bps": ((x >> 36) & 31) + 1, "total": x & ((1 << 36) - 1),
                    "max_block": int.from_bytes(body[2:4], "big")}
        pos += 4 + ln
        if hdr & 0x80:
            break
    if info is None:
        raise FlacError("no STREAMINFO")
    blocks = []
    b = _Bits(data, pos)
    while True:
        try:
            sync = b.u(15)
        except EOFError:
            break
        if sync != 0x7FFC:
            raise FlacError("lost frame sync")
        b.u(1)                                   # blocking strategy
        bs_code, sr_code = b.u(4), b.u(4)
        ch_code, ss_code = b.u(4), b.u(3)
        b.u(1)
        _utf8_number(b)
        if bs_code == 1:
            bs = 192
        elif 2 <= bs_code <= 5:
            bs = 576 << (bs_code - 2)
        elif bs_code == 6:
            bs = b.u(8) + 1
        elif bs_code == 7:
            bs = b.u(16) + 1
        elif bs_code >= 8:
            bs = 256 << (bs_code - 8)
        else:
            raise FlacError("reserved block size")
        if sr_code == 12:
            b.u(8)
        elif sr_code in (13, 14):
            b.u(16)
        b.u(8)                                   # CRC-8
        bps = {0: info["bps"], 1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}.get(ss_code)
        if bps is None:
            raise FlacError("reserved sample size")
        if ch_code < 8:
            subs = [np.array(_subframe(b, bs, bps), np.int64) for _ in range(ch_code + 1)]
        elif ch_code == 8:                       # left, side
            l, s = np.array(_subframe(b, bs, bps), np.int64), np.array(_subframe(b, bs, bps + 1), np.int64)
            subs = [l, l - s]
        elif ch_code == 9:                       # side, right
            s, r = np.array(_subframe(b, bs, bps + 1), np.int64), np.array(_subframe(b, bs, bps), np.int64)
            subs = [s + r, r]
        elif ch_code == 10:                      # mid, side
            m, s = np.array(_subframe(b, bs, bps), np.int64), np.array(_subframe(b, bs, bps + 1), np.int64)
            m = (m << 1) | (s & 1)
            subs = [(m + s) >> 1, (m - s) >> 1]
        else:
            raise FlacError("reserved channel assignment")
        b.align()
        b.u(16)                                  # CRC-16
        blocks.append(subs)
    return blocks, info["sr"], info["bps"], info["ch"]


# ---------------------------------------------------------------- writer (fixtures, tools)
class _BitsOut:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def u(self, v, k):
        if k == 0:
            return
        self.acc = (self.acc << k) | (v & ((1 << k) - 1))
        self.n += k
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def align(self):
        if self.n:
            self.u(0, 8 - self.n)


def _crc8(data):
    c = 0
    for x in data:
        c ^= x
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def _crc16(data):
    c = 0
    for x in data:
        c ^= x << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def _put_subframe(o, x, bps, order, rice_k, lpc=False, escape=False):
    """FIXED predictor of `order` (0..4) -- or the same predictor written as an LPC subframe (14-bit coefficients,
    shift 10) -- with one Rice partition (or one escaped, raw partition), or VERBATIM when order < 0.  A common
    power-of-two factor of the block is stored as wasted bits; an all-equal block as CONSTANT."""
    n = len(x)
    x = [int(v) for v in x]
    wasted = 0
    if any(x):
        while wasted < bps - 1 and not any((v >> wasted) & 1 for v in x):
            wasted += 1
    if wasted:
        x = [v >> wasted for v in x]
        bps -= wasted

    def head(typ):
        o.u(0, 1); o.u(typ, 6)
        if wasted:
            o.u(1, 1); o.u(1, wasted)
        else:
            o.u(0, 1)
    if n > 1 and all(v == x[0] for v in x):
        head(0)
        o.u(x[0], bps)
        return
    if order < 0 or n <= order:
        head(1)
        for v in x:
            o.u(v, bps)
        return
    c = _FIXED[order]
    if lpc and order >= 1:
        head(32 + order - 1)
        for i in range(order):
            o.u(x[i], bps)
        o.u(14 - 1, 4); o.u(10, 5)
        for cj in c:
            o.u(cj << 10, 14)
    else:
        head(8 + order)
        for i in range(order):
            o.u(x[i], bps)
    res = []
    for i in range(order, n):
        pred = 0
        for j, cj in enumerate(c):
            pred += cj * x[i - 1 - j]
        res.append(x[i] - pred)
    o.u(0, 2); o.u(0, 4)
    if escape:
        w = max(max((abs(r) for r in res), default=0).bit_length() + 1, 1)
        o.u(15, 4); o.u(w, 5)
        for r in res:
            o.u(r, w)
        return
    o.u(rice_k, 4)
    for r in res:
        v = (r << 1) if r >= 0 else ((-r) << 1) - 1
        q = v >> rice_k
        while q >= 32:
            o.u(0, 32)
            q -= 32
        o.u(1, q + 1)
        o.u(v, rice_k)


def encode(path, pcm, sample_rate, bps=16, blocksize=4096, stereo_mode="independent", order=2, lpc=False, escape=False):
    """pcm: int array (n,) or (n, channels).  Writes a valid FLAC stream with FIXED-predictor subframes (order < 0:
    VERBATIM); stereo_mode in {"independent", "left_side", "right_side", "mid_side"}.  Fixture writer for the tests."""
    pcm = np.asarray(pcm, np.int64)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, ch = pcm.shape
    out = bytearray(b"fLaC")
    si = struct.pack(">HH", blocksize, blocksize) + b"\0\0\0" + b"\0\0\0"
    si += ((sample_rate << 44) | ((ch - 1) << 41) | ((bps - 1) << 36) | n).to_bytes(8, "big") + b"\0" * 16
    out += bytes([0x80]) + len(si).to_bytes(3, "big") + si
    sr_code = {44100: 9, 48000: 10, 22050: 6, 16000: 5, 32000: 8, 8000: 4, 24000: 7, 96000: 11}.get(sample_rate, 0)
    ss_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6}[bps]
    frame_no = 0
    for s0 in range(0, n, blocksize):
        blk = pcm[s0:s0 + blocksize]
        bs = len(blk)
        o = _BitsOut()
        o.u(0x7FFC, 15); o.u(0, 1)
        o.u(7, 4); o.u(sr_code, 4)
        if ch == 2 and stereo_mode != "independent":
            ch_code = {"left_side": 8, "right_side": 9, "mid_side": 10}[stereo_mode]
        else:
            ch_code = ch - 1
        o.u(ch_code, 4); o.u(ss_code, 3); o.u(0, 1)
        if frame_no < 0x80:
            o.u(frame_no, 8)
        else:
            o.u(0xC0 | (frame_no >> 6), 8); o.u(0x80 | (frame_no & 0x3F), 8)
        o.u(bs - 1, 16)
        o.u(_crc8(bytes(o.out)), 8)
        k = max(0, int(np.ceil(np.log2(np.abs(np.diff(blk[:, 0], n=max(order, 0))).mean() + 1.0))))
        k = min(k, 14)
        if ch_code < 8:
            for c in range(ch):
                _put_subframe(o, blk[:, c], bps, order, k, lpc, escape)
        else:
            l, r = blk[:, 0], blk[:, 1]
            side = l - r
            if ch_code == 8:
                _put_subframe(o, l, bps, order, k, lpc, escape); _put_subframe(o, side, bps + 1, order, min(k + 1, 14), lpc, escape)
            elif ch_code == 9:
                _put_subframe(o, side, bps + 1, order, min(k + 1, 14), lpc, escape); _put_subframe(o, r, bps, order, k, lpc, escape)
            else:
                _put_subframe(o, (l + r) >> 1, bps, order, k, lpc, escape); _put_subframe(o, side, bps + 1, order, min(k + 1, 14), lpc, escape)
        o.align()
        body = bytes(o.out)
        out += body + struct.pack(">H", _crc16(body))
        frame_no += 1
    with open(path, "wb") as f:
        f.write(out)
