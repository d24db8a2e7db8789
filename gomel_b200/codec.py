"""Host-side file codecs around the GPU path: WAV and FLAC readers in the reference's two flavours (Go: beep /
mewkiz-flac semantics; Python: soundfile semantics), PNG containers (8-bit via PIL, 16-bit via a small zlib
writer/reader), float16 metadata bytes, pad bookkeeping.

Only container work and metadata packing happen here; every per-pixel float operation of
dumpimage / loadpng runs on the GPU through gomel_quantise / gomel_dequantise.

Reference: mel/impl.go:46-264, phase/impl.go:45-349, phase.py:352-402, :557-852.
"""
import struct
import wave
import zlib

import numpy as np

from . import _lib


# ---------------------------------------------------------------- pad bookkeeping (host ints)
def pad_len(n, window):
    """mel/impl.go:429-455 -- number of zeros pad() appends"""
    mt = 15 * window
    if n >= mt:
        r = (n - mt) % window
        return window - r - 1 if r else 0
    return max(mt - n - 1, 0)


def pad(buf, window):
    p = pad_len(len(buf), window)
    buf = np.asarray(buf, np.float64)
    return np.concatenate([buf, np.zeros(p)]) if p > 0 else buf


def is_padded(original_len, padded_len, window):
    """mel/impl.go:457-479"""
    mt = 15 * window
    if original_len >= mt:
        r = (original_len - mt) % window
        if r:
            return padded_len == original_len + (window - r - 1)
        return padded_len == original_len
    return padded_len == original_len + (mt - original_len - 1)


# ---------------------------------------------------------------- float16 metadata
def pack_f16_go(v):
    """packFloat16ToBytes (mel/impl.go:120-125): float64 -> float32 -> float16 (x448/float16, RNE)"""
    return np.array([np.float32(v)], np.float32).astype(np.float16).tobytes()


def pack_f16_py(v):
    """pack_float16_to_bytes (phase.py:604-620): float64 -> float16 directly"""
    return struct.pack("<e", np.float16(v))


def unpack_f16(b):
    """unpackBytesToFloat64 (mel/impl.go:46-50)"""
    return float(np.frombuffer(bytes(b), np.float16)[0])


# ---------------------------------------------------------------- WAV / FLAC
def _wav_pcm(path):
    """-> (int64 samples (n, channels), sample width in bytes, sample rate, is_float) of a RIFF/WAVE file"""
    with wave.open(path, "rb") as w:
        nch, sw, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2:
        v = np.frombuffer(raw, "<i2").astype(np.int64)
    elif sw == 1:
        v = np.frombuffer(raw, np.uint8).astype(np.int64)            # unsigned in the container
    elif sw == 3:
        b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int64)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
    else:
        v = np.frombuffer(raw, "<i4").astype(np.int64)
    return v.reshape(-1, nch), sw, sr


def load_wav(path):
    """Go flavour -- loadwav (mel/impl.go:234-264, phase/impl.go:309-349) over faiface/beep v1.1.0 wav.Decode: LEFT
    channel only; 8-bit p/255*2-1, 16-bit v/(2^15-1), 24-bit v/(2^23-1); returns (float64 samples, sample rate)"""
    try:
        v, sw, sr = _wav_pcm(path)
    except (OSError, wave.Error, EOFError) as e:
        print(e)
        return np.zeros(0), 0.0
    left = v[:, 0].astype(np.float64)
    if sw == 2:
        a = left / 32767.0
    elif sw == 1:
        a = left / 255.0 * 2.0 - 1.0
    elif sw == 3:
        a = left / float((1 << 23) - 1)
    else:
        a = left / float((1 << 31) - 1)            # beyond beep v1.1.0 (8/16/24 only); same rule extended
    return np.ascontiguousarray(a), float(sr)


def load_wav_sf(path):
    """Python flavour -- load_wav_with_sr (phase.py:551-567) over soundfile.read(dtype='float64'): integer PCM
    scaled by 1/2^(bits-1) (8-bit: (p-128)/128), channels AVERAGED; returns (float64 samples, int sample rate)"""
    v, sw, sr = _wav_pcm(path)
    x = v.astype(np.float64)
    a = (x - 128.0) / 128.0 if sw == 1 else x / float(1 << (8 * sw - 1))
    a = a[:, 0] if a.shape[1] == 1 else np.mean(a, axis=1)
    return np.ascontiguousarray(a), int(sr)


def load_flac_go(path, scale):
    """Go flavour -- loadflac (mel/impl.go:266-296: scale 256*256; phase/impl.go:351-381: scale 256*128) over
    mewkiz/flac: every frame's subframes are appended one after the other -- block of channel 0, block of channel 1,
    next frame ... -- NOT interleaved and not mixed down; integer sample / scale.  Returns (samples, sample rate)."""
    from . import flac
    try:
        blocks, sr, _, _ = flac.decode(path)
    except (OSError, flac.FlacError, EOFError, IndexError) as e:
        print(e)
        return np.zeros(0), 0.0
    if not blocks:
        return np.zeros(0), float(sr)
    out = np.concatenate([ch for blk in blocks for ch in blk]).astype(np.float64) / float(scale)
    return out, float(sr)


def load_flac_sf(path):
    """Python flavour -- load_flac_with_sr (phase.py:570-586) over soundfile: sample / 2^(bits-1), channels averaged"""
    from . import flac
    blocks, sr, bps, nch = flac.decode(path)
    if not blocks:
        return np.zeros(0), int(sr)
    x = np.stack([np.concatenate([blk[c] for blk in blocks]) for c in range(nch)], axis=1).astype(np.float64)
    x /= float(1 << (bps - 1))
    a = x[:, 0] if nch == 1 else np.mean(x, axis=1)
    return np.ascontiguousarray(a), int(sr)


def save_wav(path, data, sr):
    """Go flavour -- dumpwav (mel/impl.go:195-232): mono 16-bit PCM; beep clamps to [-1,1], scales by 2^15-1 and
    TRUNCATES (int16(v * 32767))"""
    x = np.clip(np.asarray(data, np.float64), -1.0, 1.0)
    save_wav_pcm16(path, (x * 32767.0).astype("<i2"), sr)


def save_wav_sf(path, data, sr):
    """Python flavour -- save_wav (phase.py:589-601): clip to [-1,1], soundfile.write(subtype='PCM_16'), i.e.
    libsndfile's double -> short conversion lrint(x * 0x7FFF): ROUND to nearest even"""
    x = np.clip(np.asarray(data, np.float64), -1.0, 1.0)
    save_wav_pcm16(path, np.rint(x * 32767.0).astype("<i2"), sr)


def save_wav_pcm16(path, pcm, sr):
    """The container half of dumpwav for samples already quantised (gomel_from_mel_batch_host_pcm16)."""
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr) if sr else 44100)
        w.writeframes(np.ascontiguousarray(pcm, "<i2").tobytes())


# ---------------------------------------------------------------- PNG containers
def _png_chunk(tag, data):
    c = struct.pack(">I", len(data)) + tag + data
    return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path, px):
    """px: (h, w, 3|4) uint8 or uint16 -> PNG (8- or 16-bit, RGB/RGBA), filter 0."""
    px = np.ascontiguousarray(px)
    h, w, ch = px.shape
    depth = 16 if px.dtype == np.uint16 else 8
    ctype = 6 if ch == 4 else 2
    rows = px.astype(">u2").reshape(h, -1).view(np.uint8) if depth == 16 else px.reshape(h, -1)
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rows], axis=1).tobytes()
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(_png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)))
        f.write(_png_chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(_png_chunk(b"IEND", b""))


def read_png(path):
    """-> (h, w, 3) uint8 or uint16 RGB (alpha dropped; grey / palette expanded)."""
    from PIL import Image
    with open(path, "rb") as f:
        head = f.read(26)
    depth = head[24] if len(head) >= 26 else 8
    if depth == 16:
        return _read_png16(path)
    return np.array(Image.open(path).convert("RGB"), np.uint8)


def _read_png16(path):
    data = open(path, "rb").read()
    pos, idat, w = 8, b"", None
    while pos < len(data):
        ln, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + ln]
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if interlace:
                raise ValueError("interlaced 16-bit PNG not supported")
        elif tag == b"IDAT":
            idat += body
        pos += 12 + ln
    ch = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    bpp = ch * 2
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * bpp)
    out = np.zeros((h, w * bpp), np.uint8)
    prev = np.zeros((w, bpp), np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32).reshape(w, bpp)
        if ft == 0:
            cur = line
        elif ft == 2:                                   # Up
            cur = (line + prev) & 255
        elif ft == 1:                                   # Sub: running sum per byte lane
            cur = np.cumsum(line, axis=0) & 255
        else:                                           # Average / Paeth: sequential in x, all byte lanes at once
            cur = np.zeros((w, bpp), np.int32)
            a = np.zeros(bpp, np.int32)
            c = np.zeros(bpp, np.int32)
            for x in range(w):
                b = prev[x]
                if ft == 3:
                    pr = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
                    pr = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
                a = (line[x] + pr) & 255
                cur[x] = a
                c = b
        out[y] = cur.reshape(-1)
        prev = cur
    px = out.view(">u2").astype(np.uint16).reshape(h, w, ch)
    if ch == 1:
        px = np.repeat(px, 3, axis=2)
    elif ch == 2:
        px = np.repeat(px[:, :, :1], 3, axis=2)
    return np.ascontiguousarray(px[:, :, :3])


# ---------------------------------------------------------------- mel image (mel/impl.go:52-193)
def mel_dump_image(path, buf, mels, reverse, samples_in_mel, sr, device=0):
    """dumpimage (mel/impl.go:127-193): single min/max, 8-bit NRGBA, 8 metadata bytes in blue"""
    ctx = _lib.default_context(device)
    rgb, mm = ctx.quantise(buf, mels, _lib.Q_SINGLE_MINMAX)
    stride = len(rgb) // mels
    floats = pack_f16_go(mm[0]) + pack_f16_go(mm[2]) + pack_f16_go(samples_in_mel) + pack_f16_go(sr)
    img = np.zeros((mels, stride, 4), np.uint8)
    img[:, :, :3] = rgb.reshape(stride, mels, 3).transpose(1, 0, 2).astype(np.uint8)
    img[:, :, 2] = 0
    img[mels - len(floats):, 0, 2] = np.frombuffer(floats, np.uint8)
    img[:, :, 3] = 255
    if reverse:
        img = img[::-1]
    write_png(path, img)


def mel_load_png(path, reverse, device=0):
    """loadpng (mel/impl.go:52-118) -> (buf (w*h, 2), samples, samplerate)"""
    try:
        px = read_png(path)
    except (OSError, ValueError) as e:
        print(e)
        return np.zeros((0, 2)), 0.0, 0.0
    if px.dtype == np.uint16:
        px = (px >> 8).astype(np.uint8)                  # color.RGBA() >> 8
    if reverse:
        px = px[::-1]
    h, w = px.shape[:2]
    floats = bytes(px[h - 8:, 0, 2].tolist())
    mx, mn, sim, sr = (unpack_f16(floats[i:i + 2]) for i in (0, 2, 4, 6))
    if mx == sim:                                        # mel/impl.go:105-107
        sim = 0.0
    rg = px[:, :, :2].transpose(1, 0, 2).reshape(-1, 2).astype(np.uint16)     # index x*h + y
    buf = _lib.default_context(device).dequantise(rg, False, mx, mx, mn, mn)
    return buf, sim * float(w), sr


# ---------------------------------------------------------------- phase image (phase/impl.go:51-278, phase.py:643-852)
def phase_dump_image_go(path, buf, mels, reverse, samples_in_mel, sr, ihs_passes, hdr, device=0):
    """Go dumpimage (phase/impl.go:168-278): 16 metadata bytes, blue = wrap(-val0)"""
    ctx = _lib.default_context(device)
    flags = _lib.Q_BLUE_WRAP | (_lib.Q_HDR if hdr else 0)
    rgb, mm = ctx.quantise(buf, mels, flags, ihs_passes)
    stride = len(rgb) // mels
    floats = b"".join(pack_f16_go(v) for v in (mm[0], mm[1], 0.0, mm[2], mm[3], 0.0, samples_in_mel, sr))
    dt = np.uint16 if hdr else np.uint8
    img = np.zeros((mels, stride, 4), dt)
    img[:, :, :3] = rgb.reshape(stride, mels, 3).transpose(1, 0, 2).astype(dt)
    img[mels - len(floats):, 0, 2] = np.frombuffer(floats, np.uint8)
    img[:, :, 3] = 65535 if hdr else 255
    if reverse:
        img = img[::-1]
    write_png(path, img)


def phase_load_png_go(path, reverse, ihs_passes, hdr, device=0):
    """Go loadpng (phase/impl.go:51-153) -> (buf, samples, samplerate)"""
    try:
        px = read_png(path)
    except (OSError, ValueError) as e:
        print(e)
        return np.zeros((0, 2)), 0.0, 0.0
    if hdr and px.dtype == np.uint8:
        px = px.astype(np.uint16) * 0x101                # color.RGBA() of an 8-bit image
    if not hdr and px.dtype == np.uint16:
        px = (px >> 8).astype(np.uint8)
    if reverse:
        px = px[::-1]
    h, w = px.shape[:2]
    blue = px[h - 16:, 0, 2]
    floats = bytes((blue & 0xFF).astype(np.uint8).tolist())
    v = [unpack_f16(floats[i:i + 2]) for i in range(0, 16, 2)]
    rg = px[:, :, :2].transpose(1, 0, 2).reshape(-1, 2).astype(np.uint16)
    buf = _lib.default_context(device).dequantise(rg, hdr, v[0], v[1], v[3], v[4], ihs_passes)
    return buf, v[6] * float(w), v[7]


def phase_save_image_py(path, spectrogram, num_freqs, samples_in_mel, sample_rate, y_reverse=True, hdr=False,
                        ihs=0, device=0):
    """Python save_image (phase.py:643-752): 12 metadata bytes, blue = 0, clamp, zero range -> max//2"""
    ctx = _lib.default_context(device)
    rgb, mm = ctx.quantise(spectrogram, num_freqs, _lib.Q_HDR if hdr else 0, ihs)
    stride = len(rgb) // num_freqs
    max_val = 65535 if hdr else 255
    dt = np.uint16 if hdr else np.uint8
    img = np.zeros((num_freqs, stride, 3), dt)
    # int(max_val*val) clamped to [0, max_val] (phase.py:704): in-range values never wrap, so the
    # device's truncation is already the clamped value
    img[:, :, :2] = rgb.reshape(stride, num_freqs, 3).transpose(1, 0, 2)[:, :, :2].astype(dt)
    for ch in range(2):
        if not (mm[ch] - mm[2 + ch] > 0):                # phase.py:705-706
            img[:, :, ch] = max_val // 2
    floats = b"".join(pack_f16_py(v) for v in (mm[0], mm[1], mm[2], mm[3], samples_in_mel, sample_rate))
    img[num_freqs - len(floats):, 0, 2] = np.frombuffer(floats, np.uint8)
    if y_reverse:
        img = img[::-1]
    write_png(path, img)


def phase_load_image_py(path, y_reverse=True, hdr=False, ihs=0, device=0):
    """Python load_image (phase.py:755-852) -> (buf, samples, sample_rate, num_freqs)"""
    px = read_png(path)
    if hdr and px.dtype == np.uint8:
        px = px.astype(np.uint16)
    if not hdr and px.dtype == np.uint16:
        px = (px >> 8).astype(np.uint8)
    if y_reverse:
        px = px[::-1]
    h, w = px.shape[:2]
    floats = bytes((px[h - 12:, 0, 2] & 0xFF).astype(np.uint8).tolist())
    meta = [unpack_f16(floats[i:i + 2]) for i in range(0, 12, 2)]
    rg = px[:, :, :2].transpose(1, 0, 2).reshape(-1, 2).astype(np.uint16)
    buf = _lib.default_context(device).dequantise(rg, hdr, meta[0], meta[1], meta[2], meta[3], ihs)
    return buf, meta[4] * w, int(meta[5]), h
