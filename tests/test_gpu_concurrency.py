"""Concurrent callers.  The reference's Mel / Phase methods are pure functions of the receiver and safe to call
from many goroutines (SURVEY 8b "Threading"); the drop-in keeps that: calls on one context are serialised by the
library, filterbank tables are looked up by the key carried in each call's config (no context-wide "current
tables"), and several contexts may run side by side.  ctypes releases the GIL, so these threads really overlap."""
import threading

import numpy as np
import pytest

from util import rel_l2, synth_clip

pytestmark = pytest.mark.gpu

CONFIGS = [  # NumMels, MelFmax, Window, Resolut
    (192, 16000.0, 1280, 4096),      # cmd/tomel
    (160, 8000.0, 256, 2048),        # mel.NewMel defaults
    (80, 16000.0, 1280, 4096),
    (192, 8000.0, 1280, 4096),       # same Resolut / NumMels as the first, different MelFmax -> different tables
]


def _mel(conf, iters=3):
    from gomel_b200 import NewMel
    m = NewMel()
    m.NumMels, m.MelFmax, m.Window, m.Resolut = conf
    m.GriffinLimIterations = iters
    return m


def _work(conf, clip):
    """one ToMel + FromMel round with a fixed start signal -> deterministic result"""
    m = _mel(conf)
    wav = synth_clip(clip, 0.5)
    mel = m.ToMel(wav)
    frames = len(mel) // conf[0]
    m.InitSignal = np.random.default_rng(clip).random(conf[3] + (frames - 1) * conf[2])
    return mel, m.FromMel(mel.copy())


def _run_threads(fn, n):
    out, errs = [None] * n, []

    def body(i):
        try:
            out[i] = fn(i)
        except Exception as e:      # surfaced below
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=body, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    return out


def test_threads_with_different_mel_configs_share_the_default_context(ctx, oracle):
    expect = [_work(CONFIGS[i % 4], 300 + i) for i in range(8)]            # sequential
    for _ in range(3):
        got = _run_threads(lambda i: _work(CONFIGS[i % 4], 300 + i), 8)    # concurrent, interleaved set + call
        for (m0, w0), (m1, w1) in zip(expect, got):
            assert np.array_equal(m0, m1) and np.array_equal(w0, w1)
    # and the sequential results are right
    conf = CONFIGS[3]
    ocfg = oracle.config(num_mels=conf[0], mel_fmax=conf[1], window=conf[2], resolut=conf[3], gl_iters=3)
    wav = synth_clip(303, 0.5)
    ref = oracle.to_mel(ocfg, wav)
    assert rel_l2(np.exp(expect[3][0]), np.exp(ref)) < 1e-5
    ctx.set_mel_tables(__import__("gomel_b200")._lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, 16000.0)


def test_one_context_per_thread(oracle):
    from gomel_b200 import _lib
    conf = CONFIGS[0]
    wavs = [synth_clip(320 + i, 0.4) for i in range(4)]
    ocfg = oracle.config(gl_iters=2)
    refs = [oracle.to_mel(ocfg, w) for w in wavs]

    def fn(i):
        c = _lib.Context(0)
        cfg = _lib.make_config(n_fft=4096, hop=1280, n_mels=192, gl_iters=2)
        c.use_mel_tables(cfg, 0.0, 16000.0)
        res = [c.to_mel(cfg, wavs[i]) for _ in range(3)]
        frames = len(res[0]) // conf[0]
        init = np.random.default_rng(i).random(4096 + (frames - 1) * 1280)
        wav = c.from_mel(cfg, refs[i], init=init)
        c.close()
        return res, wav, init

    for i, (res, wav, init) in enumerate(_run_threads(fn, 4)):
        assert all(np.array_equal(res[0], r) for r in res[1:])
        assert rel_l2(np.exp(res[0]), np.exp(refs[i])) < 1e-5
        assert rel_l2(wav, oracle.from_mel(ocfg, refs[i], init)) < 1e-4


def test_phase_and_mel_threads_interleave(ctx, oracle):
    from gomel_b200.phase import NewPhase
    wav = synth_clip(340, 0.6)
    p = NewPhase()
    spec0 = p.ToPhase(wav)
    back0 = p.FromPhase(spec0)
    mel0, gl0 = _work(CONFIGS[1], 341)

    def fn(i):
        if i % 2:
            return _work(CONFIGS[1], 341)
        q = NewPhase()
        s = q.ToPhase(wav)
        return s, q.FromPhase(s)

    for i, (a, b) in enumerate(_run_threads(fn, 6)):
        if i % 2:
            assert np.array_equal(a, mel0) and np.array_equal(b, gl0)
        else:
            assert np.array_equal(a, spec0) and np.array_equal(b, back0)
    ctx.set_mel_tables(__import__("gomel_b200")._lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, 16000.0)


def test_legacy_unkeyed_config_uses_the_most_recent_tables(ctx, oracle):
    """configs that leave mel_fmin = mel_fmax = 0 keep the round-1 behaviour: the last set_mel_tables wins"""
    from gomel_b200 import _lib
    wav = synth_clip(350, 0.3)
    cfg = _lib.make_config(n_fft=4096, hop=1280, n_mels=192)
    for fmax in (8000.0, 16000.0, 8000.0):
        ctx.set_mel_tables(_lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, fmax)
        ref = oracle.to_mel(oracle.config(mel_fmax=fmax), wav)
        assert rel_l2(np.exp(ctx.to_mel(cfg, wav)), np.exp(ref)) < 1e-5
    # a keyed config is not disturbed by what was set last
    kcfg = _lib.make_config(n_fft=4096, hop=1280, n_mels=192)
    ctx.use_mel_tables(kcfg, 0.0, 16000.0)
    ctx.set_mel_tables(_lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, 8000.0)
    ref = oracle.to_mel(oracle.config(mel_fmax=16000.0), wav)
    assert rel_l2(np.exp(ctx.to_mel(kcfg, wav)), np.exp(ref)) < 1e-5
    ctx.set_mel_tables(_lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, 16000.0)


def test_table_sets_are_evicted_least_recently_used_and_recovered(ctx, oracle):
    """more distinct keys than the library keeps (64): the Python mirror re-registers on GOMEL_E_STATE"""
    from gomel_b200 import _lib
    wav = synth_clip(351, 0.2)
    first = _lib.make_config(n_fft=4096, hop=1280, n_mels=192)
    ctx.use_mel_tables(first, 0.0, 16000.0)
    a = ctx.to_mel(first, wav)
    for i in range(70):
        c = _lib.make_config(n_fft=4096, hop=1280, n_mels=192)
        ctx.use_mel_tables(c, 0.0, 9000.0 + i)
    b = ctx.to_mel(first, wav)          # its set was evicted; recovered transparently
    assert np.array_equal(a, b)
    ctx.set_mel_tables(_lib.make_config(n_fft=4096, hop=1280, n_mels=192), 0.0, 16000.0)
