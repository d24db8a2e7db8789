"""Directory-level batch tools on top of the batched C ABI: many files per GPU call.

    tomel_dir(in_dir, out_dir, mel)     every *.wav / *.flac  ->  <name>.png, identical to Mel.ToMelWav / ToMelFlac per file
    towav_dir(in_dir, out_dir, mel)     every *.png  ->  <name>.wav, identical to Mel.ToWavPng per file (same start signal)

ToMel frames depend only on local samples, so clips of different length share one batch: every clip is
zero-padded (pad() zeros, then more zeros) to the longest one and only its own frames are kept.  Griffin-Lim
couples neighbouring frames, so spectrograms are grouped by frame count.
"""
import ctypes as C
import glob
import os

import numpy as np

from . import _lib, codec


def tomel_dir(in_dir, out_dir, mel, chunk=64):
    """mel: a configured gomel_b200.Mel.  Returns the list of PNG paths written."""
    files = sorted(glob.glob(os.path.join(in_dir, "*.wav")) + glob.glob(os.path.join(in_dir, "*.flac")))
    os.makedirs(out_dir, exist_ok=True)
    cfg = mel._cfg()
    ctx = mel._ctx(cfg)
    written = []
    for c0 in range(0, len(files), chunk):
        part = files[c0:c0 + chunk]
        # per file exactly what the single-file methods decode: loadwav, or loadflac with the mel package's 1/65536
        clips = [codec.load_flac_go(f, 256 * 256) if f.endswith(".flac") else codec.load_wav(f) for f in part]
        clips = [(f, b, sr) for f, (b, sr) in zip(part, clips) if len(b) > 0]
        if not clips:
            continue
        n_max = max(len(b) for _, b, _ in clips)
        _, fr_max, _ = _lib.frames(cfg, n_max)
        wav = np.zeros((len(clips), n_max), np.float32)
        for i, (_, b, _) in enumerate(clips):
            wav[i, :len(b)] = b
        out = np.empty((len(clips), fr_max * cfg.n_mels, 2), np.float32)
        ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), len(clips), n_max,
                                                  out.ctypes.data_as(C.c_void_p), 0))
        for i, (f, b, sr) in enumerate(clips):
            _, fr, _ = _lib.frames(cfg, len(b))
            spec = out[i, :fr * cfg.n_mels].astype(np.float64)
            dst = os.path.join(out_dir, os.path.basename(f) + ".png")
            codec.mel_dump_image(dst, spec, mel.NumMels, mel.YReverse, float(len(b) * mel.NumMels) / float(len(spec)),
                                 float(sr), device=mel.Device)
            written.append(dst)
    return written


def towav_dir(in_dir, out_dir, mel, seed=0, init_signals=None):
    """Returns the list of WAV paths written.  init_signals: optional {basename: start signal} for parity runs;
    otherwise the device draws U[0,1) per clip from `seed`."""
    files = sorted(glob.glob(os.path.join(in_dir, "*.png")))
    os.makedirs(out_dir, exist_ok=True)
    cfg = mel._cfg()
    ctx = mel._ctx(cfg)
    loaded = {}
    for f in files:
        buf, samples, sr = codec.mel_load_png(f, mel.YReverse, device=mel.Device)
        if len(buf) == 0 or len(buf) % mel.NumMels:
            continue                                            # the reference prints / panics; the batch tool skips the file
        loaded.setdefault(len(buf) // mel.NumMels, []).append((f, buf + mel.VolumeBoost, samples, sr))
    written = []
    for frames, group in sorted(loaded.items()):
        ola = cfg.n_fft + (frames - 1) * cfg.hop
        spec = np.stack([g[1] for g in group]).astype(np.float32)
        init = None
        if init_signals is not None:
            init = np.stack([init_signals[os.path.basename(g[0])] for g in group]).astype(np.float32)
        # the waveforms come back as the 16-bit PCM samples dumpwav would write (quantised on the GPU)
        out = np.empty((len(group), ola), np.int16)
        ctx.check(ctx.lib.gomel_from_mel_batch_host_pcm16(
            ctx.h, C.byref(cfg), spec.ctypes.data_as(C.c_void_p), len(group), frames,
            init.ctypes.data_as(C.c_void_p) if init is not None else None, seed, out.ctypes.data_as(C.c_void_p), 0))
        for i, (f, _, samples, sr) in enumerate(group):
            w = out[i]
            if int(samples) > 0 and codec.is_padded(int(samples), len(w), mel.Window) and len(w) > int(samples):
                w = w[:int(samples)]
            rate = mel.SampleRate if mel.SampleRate else int(sr)
            dst = os.path.join(out_dir, os.path.basename(f) + ".wav")
            codec.save_wav_pcm16(dst, w, rate)
            written.append(dst)
    return written
