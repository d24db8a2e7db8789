"""Host-side mirror of the reference's Go package `mel` (mel/mel.go) on top of the C ABI.

Same exported names, field names, argument meaning and error behaviour as the Go API, so the
parity tests read like tests of the reference:

    m = NewMel(); m.NumMels = 192; m.MelFmax = 16000; m.Window = 1280; m.Resolut = 4096
    spec = m.ToMel(wav)            # mel/mel.go:46    -> (frames*NumMels, 2) float64, natural log
    wav  = m.FromMel(spec)         # mel/mel.go:142   Griffin-Lim, GriffinLimIterations passes
    img  = m.Image(spec)           # mel/mel.go:171   uint16 per entry

All arithmetic runs in libgomelcuda.so on the GPU; this file only marshals.
"""
import numpy as np

from . import _lib
from . import codec


class ErrFileNotLoaded(Exception):
    """mel.ErrFileNotLoaded = errors.New("wavNotLoaded")  (mel/mel.go:43)"""

    def __init__(self):
        super().__init__("wavNotLoaded")


class Mel:
    """mel.Mel (mel/mel.go:10-27)"""

    def __init__(self):
        # NewMel defaults (mel/mel.go:30-41)
        self.NumMels = 160
        self.MelFmin = 0.0
        self.MelFmax = 8000.0
        self.TuneMul = 1.0
        self.TuneAdd = 0.0
        self.Window = 256
        self.Resolut = 2048
        self.YReverse = False
        self.GriffinLimIterations = 2
        self.VolumeBoost = 0.0
        self.SampleRate = 0
        # extensions (not in the reference): device and Griffin-Lim start-signal injection
        self.Device = 0
        self.InitSignal = None      # ola_len float64 replacing rand.Float64() (mel/mel.go:80-83)
        self.Seed = None
        # Griffin-Lim precision.  False (default): the library's float64 lead iterations, then float32;
        # True: every iteration in float64 on the fused kernel (GOMEL_FLAG_F64); "ref": the round-1 strict
        # float64 path (GOMEL_FLAG_F64_REF), a slow test instrument
        self.Strict = False

    # ---- plumbing
    def _cfg(self):
        return _lib.make_config(n_fft=self.Resolut, hop=self.Window, n_mels=self.NumMels, n_freqs=0,
                                gl_iters=self.GriffinLimIterations, tune_mul=self.TuneMul,
                                tune_add=self.TuneAdd, volume_boost=self.VolumeBoost,
                                flags=(_lib.FLAG_F64_REF if self.Strict == "ref" else _lib.FLAG_F64) if self.Strict else 0)

    def _ctx(self, cfg):
        ctx = _lib.default_context(self.Device)
        ctx.use_mel_tables(cfg, self.MelFmin, self.MelFmax)     # keyed: safe when Mel objects share the context
        return ctx

    # ---- buffer API
    def ToMel(self, buf):
        """mel.ToMel (mel/mel.go:46-74)"""
        cfg = self._cfg()
        return self._ctx(cfg).to_mel(cfg, buf)

    def FromMel(self, ospectrum):
        """mel.FromMel (mel/mel.go:142-152).  Like the reference it exp()s `ospectrum` IN PLACE
        (spectral_denormalize, mel/impl.go:421-427) when given a float64 ndarray."""
        cfg = self._cfg()
        ctx = self._ctx(cfg)
        spec = np.asarray(ospectrum)
        frames = spec.reshape(-1, 2).shape[0] // max(self.NumMels, 1)
        ola = self.Resolut + (frames - 1) * self.Window
        init = self.InitSignal
        if init is None:
            rng = np.random.default_rng(self.Seed)
            init = rng.random(max(ola, 0))               # rand.Float64() per sample, uniform [0,1)
        out = ctx.from_mel(cfg, spec, init=init)
        if isinstance(ospectrum, np.ndarray) and ospectrum.dtype == np.float64:
            np.exp(ospectrum, out=ospectrum)
        return out

    def Image(self, buf):
        """Mel.Image (mel/mel.go:171-173) -> dumpbuffer (mel/impl.go:16-44)"""
        return _lib.default_context(self.Device).image(buf, self.NumMels)

    # ---- file API (mel/mel.go:176-238); codecs on the host, arithmetic on the GPU
    def ToMelWav(self, inputFile, outputFile):
        buf, sr = codec.load_wav(inputFile)
        if len(buf) == 0:
            raise ErrFileNotLoaded()
        ospectrum = self.ToMel(buf)
        codec.mel_dump_image(outputFile, ospectrum, self.NumMels, self.YReverse,
                             float(len(buf) * self.NumMels) / float(len(ospectrum)), float(sr), device=self.Device)

    def ToMelFlac(self, inputFile, outputFile):
        """mel.ToMelFlac (mel/mel.go:176-192): loadflac scales samples by 1/65536 here (mel/impl.go:290) -- half the
        amplitude of the phase package's loader; kept"""
        buf, sr = codec.load_flac_go(inputFile, 256 * 256)
        if len(buf) == 0:
            raise ErrFileNotLoaded()
        ospectrum = self.ToMel(buf)
        codec.mel_dump_image(outputFile, ospectrum, self.NumMels, self.YReverse,
                             float(len(buf) * self.NumMels) / float(len(ospectrum)), float(sr), device=self.Device)

    def ToWavPng(self, inputFile, outputFile):
        buf, samples, samplerate = codec.mel_load_png(inputFile, self.YReverse, device=self.Device)
        if len(buf) == 0:
            raise ErrFileNotLoaded()
        buf += self.VolumeBoost                                   # mel/mel.go:218-221 (additive, log domain)
        owave = self.FromMel(buf)
        if int(samples) > 0 and codec.is_padded(int(samples), len(owave), self.Window) and len(owave) > int(samples):
            owave = owave[:int(samples)]
        if samplerate != 0 and self.SampleRate == 0:
            self.SampleRate = int(samplerate)
        codec.save_wav(outputFile, owave, self.SampleRate)


def NewMel():
    """mel.NewMel (mel/mel.go:30-41)"""
    return Mel()


def LoadFlac(inputFile):
    """mel.LoadFlac (mel/mel.go:155-158)"""
    return codec.load_flac_go(inputFile, 256 * 256)[0]


def LoadWav(inputFile):
    """mel.LoadWav (mel/mel.go:161-164)"""
    return codec.load_wav(inputFile)[0]


def SaveWav(outputFile, vec, sr):
    """mel.SaveWav (mel/mel.go:167-169)"""
    return codec.save_wav(outputFile, vec, sr)
