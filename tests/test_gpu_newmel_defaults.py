"""GPU parity tests for the mel.NewMel() DEFAULT geometry (mel/mel.go:30-41: NumMels 160, MelFmax 8000,
Window 256, Resolut 2048, GriffinLimIterations 2) -- a library user who never touches Window/Resolut gets
these; the cmd/* tools all set 1280/4096 (covered by test_gpu_parity.py).  Same tolerances."""
import ctypes as C

import numpy as np
import pytest

from util import rel_l2, synth_clip

pytestmark = pytest.mark.gpu

TOL_STFT = 1e-5
TOL_GL = 1e-4


@pytest.fixture(scope="module")
def lib():
    from gomel_b200 import _lib
    return _lib


@pytest.fixture()
def restore_tables(ctx, lib):
    yield
    ctx.set_mel_tables(lib.make_config(n_fft=4096, hop=1280, n_mels=192, n_freqs=768, gl_iters=2), 0.0, 16000.0)


def _ocfg(oracle, iters=2, mels=160, fmax=8000.0, **kw):
    return oracle.config(num_mels=mels, window=256, resolut=2048, mel_fmax=fmax, gl_iters=iters, **kw)


def _newmel(iters=2):
    from gomel_b200 import NewMel
    m = NewMel()                     # defaults untouched on purpose
    m.GriffinLimIterations = iters
    return m


@pytest.mark.parametrize("clip,seconds,tile", [(0, 1.0, 0), (1, 0.05, 0), (2, 2.3, 6), (3, 0.31, 4), (4, 5.0, 0)])
def test_to_mel_defaults(ctx, oracle, restore_tables, clip, seconds, tile):
    wav = synth_clip(clip, seconds)
    ctx.set_tile_frames(tile)
    got = _newmel().ToMel(wav)
    ctx.set_tile_frames(0)
    ref = oracle.to_mel(_ocfg(oracle), wav)
    assert got.shape == ref.shape
    assert rel_l2(np.exp(got), np.exp(ref)) < TOL_STFT, rel_l2(np.exp(got), np.exp(ref))
    assert np.max(np.abs(got - ref)) < 2e-3           # log domain, incl. the 1e-5 clamp floor


def _from_mel_case(ctx, oracle, clip, seconds, iters, tile, seed):
    wav = synth_clip(clip, seconds)
    ocfg = _ocfg(oracle, iters)
    mel = oracle.to_mel(ocfg, wav)
    frames = len(mel) // 160
    init = np.random.default_rng(seed).random(2048 + (frames - 1) * 256)
    m = _newmel(iters)
    m.InitSignal = init
    ctx.set_tile_frames(tile)
    got = m.FromMel(mel.copy())
    ctx.set_tile_frames(0)
    return got, oracle.from_mel(ocfg, mel, init)


# tile 4 and 6 are below the halo length (7 hops): the library must raise them to 8 on its own
@pytest.mark.parametrize("seconds,iters,tile", [(0.3, 0, 0), (0.3, 1, 0), (0.3, 2, 4), (1.0, 2, 0), (1.0, 3, 6),
                                                (1.0, 8, 8), (0.45, 2, 10), (3.0, 2, 0)])
def test_from_mel_defaults(ctx, oracle, restore_tables, seconds, iters, tile):
    got, ref = _from_mel_case(ctx, oracle, 21, seconds, iters, tile, 7000)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < TOL_GL, rel_l2(got, ref)


def test_from_mel_defaults_32_iterations(ctx, oracle, restore_tables):
    got, ref = _from_mel_case(ctx, oracle, 22, 0.8, 32, 8, 7001)
    assert rel_l2(got, ref) < TOL_GL, rel_l2(got, ref)


@pytest.mark.parametrize("seconds,iters,tile", [(0.3, 1, 0), (1.0, 3, 6), (1.0, 20, 8), (2.2, 32, 0)])
def test_from_mel_defaults_float64_matches_oracle(ctx, oracle, restore_tables, seconds, iters, tile):
    """the float64 kernel on the Resolut 2048 / Window 256 geometry (frame zero-extended through the 4096-point core):
    GOMEL_FLAG_F64 reproduces the oracle to 1e-10 here too, and the default policy (<= 16 iterations: all float64) with it"""
    wav = synth_clip(23, seconds)
    ocfg = _ocfg(oracle, iters)
    mel = oracle.to_mel(ocfg, wav)
    frames = len(mel) // 160
    init = np.random.default_rng(7002).random(2048 + (frames - 1) * 256)
    ref = oracle.from_mel(ocfg, mel, init)
    ctx.set_tile_frames(tile)
    for strict in (True, False):
        m = _newmel(iters)
        m.Strict = strict
        m.InitSignal = init
        got = m.FromMel(mel.copy())
        assert got.shape == ref.shape
        assert rel_l2(got, ref) < (1e-10 if (strict or iters <= 16) else TOL_GL), (strict, rel_l2(got, ref))
    ctx.set_tile_frames(0)


@pytest.mark.parametrize("tile", [0, 10])
def test_gl_guard_on_the_defaults_geometry(ctx, oracle, restore_tables, tile):
    """the float32 tail's singular-bin guard on Resolut 2048 / Window 256: with the threshold under the clip's leverage the
    float32 iterations are re-run in float64 and the call ends at the GOMEL_FLAG_F64 result; above it nothing changes"""
    wav = synth_clip(24, 1.3)
    ocfg = _ocfg(oracle, 20)
    mel = oracle.to_mel(ocfg, wav)
    frames = len(mel) // 160
    init = np.random.default_rng(7003).random(2048 + (frames - 1) * 256)

    def run(strict):
        m = _newmel(20)
        m.Strict = strict
        m.InitSignal = init
        return m.FromMel(mel.copy())

    ctx.set_tile_frames(tile)
    prev = ctx.set_gl_guard(0.0)
    try:
        exact, bare = run(True), run(False)
        assert ctx.last_gl_guard()[0] == 0
        ctx.set_gl_guard(1e30)
        assert np.array_equal(run(False), bare)
        n, rerun, lev, _ = ctx.last_gl_guard()
        assert n == 1 and rerun == 0 and lev > 0
        ctx.set_gl_guard(lev / 2 / np.sqrt(frames / 342.0))
        got = run(False)
        assert ctx.last_gl_guard()[1] == 1
        assert rel_l2(got, exact) < 2e-7, rel_l2(got, exact)
        assert 0 < rel_l2(bare, exact) < TOL_GL
    finally:
        ctx.set_gl_guard(prev)
        ctx.set_tile_frames(0)


@pytest.mark.parametrize("frames", [1, 2, 3, 7, 8, 9, 15, 16, 17])
def test_from_mel_defaults_tiny_frame_counts(ctx, oracle, restore_tables, frames):
    rng = np.random.default_rng(frames)
    mel = rng.uniform(-9.0, 3.0, (frames * 160, 2))
    init = rng.random(2048 + (frames - 1) * 256)
    m = _newmel(3)
    m.InitSignal = init
    got = m.FromMel(mel.copy())
    ref = oracle.from_mel(_ocfg(oracle, 3), mel, init)
    assert got.shape == ref.shape and rel_l2(got, ref) < TOL_GL, rel_l2(got, ref)


def test_roundtrip_defaults_through_both_directions(ctx, oracle, restore_tables):
    """ToMel -> FromMel on the GPU end to end, against the oracle doing the same"""
    wav = synth_clip(23, 0.7)
    m = _newmel(4)
    mel = m.ToMel(wav)
    frames = len(mel) // 160
    init = np.random.default_rng(9).random(2048 + (frames - 1) * 256)
    m.InitSignal = init
    got = m.FromMel(mel.copy())
    ref = oracle.from_mel(_ocfg(oracle, 4), mel, init)
    assert rel_l2(got, ref) < TOL_GL


def test_other_mel_counts_defaults_geometry(ctx, oracle, restore_tables):
    wav = synth_clip(24, 0.4)
    for mels, fmax in ((80, 8000.0), (192, 16000.0)):
        m = _newmel(2)
        m.NumMels, m.MelFmax = mels, fmax
        ocfg = _ocfg(oracle, 2, mels, fmax)
        ref = oracle.to_mel(ocfg, wav)
        got = m.ToMel(wav)
        assert got.shape == ref.shape and rel_l2(np.exp(got), np.exp(ref)) < TOL_STFT
        frames = len(ref) // mels
        init = np.random.default_rng(mels).random(2048 + (frames - 1) * 256)
        m.InitSignal = init
        assert rel_l2(m.FromMel(ref.copy()), oracle.from_mel(ocfg, ref, init)) < TOL_GL


def test_batch_apis_defaults_geometry(ctx, lib, oracle, restore_tables):
    cfg = lib.make_config(n_fft=2048, hop=256, n_mels=160, n_freqs=0, gl_iters=3)
    ctx.set_mel_tables(cfg, 0.0, 8000.0)
    n_clips, n = 5, 9000
    wav = np.stack([synth_clip(50 + c, n=n) for c in range(n_clips)]).astype(np.float32)
    _, fr, ola = lib.frames(cfg, n)
    mel = np.empty((n_clips, fr * 160, 2), np.float32)
    ctx.check(ctx.lib.gomel_to_mel_batch_host(ctx.h, C.byref(cfg), wav.ctypes.data_as(C.c_void_p), n_clips, n,
                                              mel.ctypes.data_as(C.c_void_p), 2))
    ocfg = _ocfg(oracle, 3)
    for c in range(n_clips):
        ref = oracle.to_mel(ocfg, wav[c].astype(np.float64))
        assert rel_l2(np.exp(mel[c]), np.exp(ref)) < TOL_STFT
    init = np.stack([np.random.default_rng(70 + c).random(ola) for c in range(n_clips)]).astype(np.float32)
    out = np.empty((n_clips, ola), np.float32)
    ctx.check(ctx.lib.gomel_from_mel_batch_host(
        ctx.h, C.byref(cfg), mel.ctypes.data_as(C.c_void_p), n_clips, fr,
        init.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_void_p), 2))
    for c in range(n_clips):
        ref = oracle.from_mel(ocfg, mel[c].astype(np.float64), init[c].astype(np.float64))
        assert rel_l2(out[c], ref) < TOL_GL


def test_geometries_outside_the_two_supported_fail_loudly(ctx, lib, restore_tables):
    from gomel_b200 import NewMel
    from gomel_b200.phase import NewPhase
    m = NewMel()
    m.Window, m.Resolut = 512, 2048
    with pytest.raises(lib.GomelError) as e:
        m.ToMel(np.zeros(10000))
    assert e.value.code == lib.E_UNSUPPORTED
    p = NewPhase()
    p.window, p.resolut = 256, 2048                     # the phase package never defaults to this (phase/phase.go:20-27)
    with pytest.raises(lib.GomelError) as e:
        p.ToPhase(np.zeros(10000))
    assert e.value.code == lib.E_UNSUPPORTED
    s = NewMel()
    s.Strict = "ref"                                    # the round-1 strict test instrument is native-geometry only
    with pytest.raises(lib.GomelError) as e:
        s.FromMel(np.zeros((160 * 4, 2)))
    assert e.value.code == lib.E_UNSUPPORTED
