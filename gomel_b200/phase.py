"""Drop-in for the reference's Python module phase.py (class Phase, phase.py:16-349) and mirror
of the Go package `phase` (phase/phase.go): same constructor, attributes, method names, shapes
and error behaviour -- but to_phase / from_phase run on the GPU through the C ABI of
libgomelcuda.so (ctypes), not in NumPy.

    ph = Phase(sample_rate=48000)
    spec = ph.to_phase(audio)        # phase.py:113  -> (frames*num_freqs, 2) float64 (Im, Re)
    audio2 = ph.from_phase(spec)     # phase.py:144  -> float64 waveform

Go-style aliases (NumFreqs, ToPhase, FromPhase, Image ...) are provided for the Go package's
callers.  Python-vs-Go behavioural differences listed in SURVEY.md Appendix C are kept per
language (e.g. volume boost applied iff > 0 here, iff != 0 through the Go names).
"""
import numpy as np

from . import _lib
from . import codec
from .codec import is_padded, pad  # noqa: F401  (module-level helpers of the reference)


class Phase:
    """Phase-preserving spectrogram encoder/decoder (phase.py:16-349)."""

    def __init__(self, sample_rate=None, num_freqs=None, window=1280, resolut=4096, y_reverse=True,
                 volume_boost=0.0, HDR=False, IHS=False, device=0):
        self.sample_rate = sample_rate
        self.window = window
        self.resolut = resolut
        self.y_reverse = y_reverse
        self.volume_boost = volume_boost
        self.HDR = HDR
        self.IHS = 0 if HDR else 2 if IHS else 0          # phase.py:41
        self.num_freqs = 0                                # "bad defaults", phase.py:43
        self.family = None
        self.device = device
        if sample_rate is not None:
            self.reconfigure_sr(sample_rate)
        if num_freqs is not None and sample_rate is None:
            self.num_freqs = num_freqs                    # extension: explicit NumFreqs like the Go struct

    # ---- configuration (phase.py:49-111)
    def reconfigure_sr(self, sample_rate):
        if sample_rate in [8000, 16000, 24000, 32000, 48000]:
            self.num_freqs = 768 * 2 if self.HDR else 768
            self.family = True
        elif sample_rate in [11025, 22050, 44100]:
            self.num_freqs = 836 * 2 if self.HDR else 836
            self.family = False
        else:
            raise ValueError(
                f"Unsupported sample rate: {sample_rate}. "
                f"Supported rates are: 8000, 16000, 24000, 32000, 48000, 11025, 22050, 44100")

    def pad_shift(self, sample_rate):
        table = ({48000: (0, 0), 32000: (2, 1), 24000: (1, 1), 16000: (1, 2), 8000: (1, 5)} if self.family
                 else {44100: (0, 0), 22050: (1, 1), 11025: (1, 3)})
        if sample_rate in table:
            return table[sample_rate]
        raise ValueError("Unsupported sample_ratePlease configure sample_rate to Phase")

    def zero_pad(self, sr):
        return self.pad_shift(sr)[0]

    def zero_shift(self, sr):
        return self.pad_shift(sr)[1]

    # ---- GPU transforms
    def _cfg(self, boost):
        return _lib.make_config(n_fft=self.resolut, hop=self.window, n_mels=0, n_freqs=self.num_freqs,
                                gl_iters=0, volume_boost=boost)

    def to_phase(self, audio_buffer):
        """phase.py:113-142 / phase.ToPhase (phase/phase.go:41-70)"""
        return _lib.default_context(self.device).to_phase(self._cfg(0.0), audio_buffer)

    def from_phase(self, spectrogram):
        """phase.py:144-220: volume boost applied iff > 0 (phase.py:216)"""
        boost = self.volume_boost if self.volume_boost > 0 else 0.0
        return _lib.default_context(self.device).from_phase(self._cfg(boost), spectrogram)

    # ---- file API (phase.py:222-349)
    def _prepare(self, audio, sample_rate, update_rate):
        self.reconfigure_sr(sample_rate=sample_rate)
        zp, zs = self.zero_pad(sample_rate), self.zero_shift(sample_rate)
        if zp > 0:
            original_len = len(audio)
            audio = zero_stuff_upsample(audio, zp, zs)
            if update_rate:
                sample_rate = int(sample_rate * len(audio) / original_len)
        return audio, sample_rate

    def to_phase_wav(self, input_file, output_file):
        audio, sample_rate = load_wav_with_sr(input_file)
        audio, sample_rate = self._prepare(audio, sample_rate, update_rate=False)      # phase.py:234-240
        original_length = len(audio)
        spectrogram = self.to_phase(audio)
        samples_in_mel = float(original_length * self.num_freqs) / float(len(spectrogram))
        save_image(output_file, spectrogram, self.num_freqs, samples_in_mel, sample_rate, self.y_reverse,
                   self.HDR, self.IHS, device=self.device)

    def _flac_spectrogram(self, input_file):
        audio, sample_rate = load_flac_with_sr(input_file)
        audio, sample_rate = self._prepare(audio, sample_rate, update_rate=True)       # phase.py:267-276, :303-312
        return audio, sample_rate, self.to_phase(audio)

    def to_phase_flac(self, input_file, output_file):
        """phase.py:255-288: like to_phase_wav, but the embedded sample rate follows the zero-stuffing ratio"""
        audio, sample_rate, spectrogram = self._flac_spectrogram(input_file)
        samples_in_mel = float(len(audio) * self.num_freqs) / float(len(spectrogram))
        save_image(output_file, spectrogram, self.num_freqs, samples_in_mel, sample_rate, self.y_reverse,
                   self.HDR, self.IHS, device=self.device)

    def to_tensor_flac(self, input_file):
        """phase.py:291-318: FLAC file -> spectrogram array, nothing written"""
        return self._flac_spectrogram(input_file)[2]

    def to_wav_png(self, input_file, output_file):
        spectrogram, samples, embedded_sample_rate, self.num_freqs = load_image(
            input_file, self.y_reverse, self.HDR, self.IHS, device=self.device)
        audio = self.from_phase(spectrogram)
        main_rate = 48000 if self.num_freqs in [768, 768 * 2] else 44100
        standard_rates = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000]
        sample_rate = min(standard_rates, key=lambda x: abs(x - embedded_sample_rate))
        original_length = int(samples)
        if len(audio) > original_length > 0:
            audio = audio[:original_length]
        save_wav(output_file, audio, main_rate)
        return sample_rate

    # ---- Go package `phase` names (phase/phase.go)
    @property
    def NumFreqs(self):
        return self.num_freqs

    @NumFreqs.setter
    def NumFreqs(self, v):
        self.num_freqs = v

    def ToPhase(self, buf):
        """phase.ToPhase (phase/phase.go:41-70)"""
        return self.to_phase(buf)

    def _go_to_png(self, buf, sr, output_file):
        """body shared by phase.ToPhaseFlac / ToPhaseWav (phase/phase.go:195-244): the Go package keeps the length
        BEFORE zero stuffing for the metadata, pads / shifts by the Go table (phase/impl.go:476-507)"""
        if len(buf) == 0:
            raise ErrFileNotLoaded()
        original = len(buf)
        zp, zs = GO_PAD_SHIFT.get(int(sr), (0, 0))            # padShift: unknown rates -> no stuffing, no error
        if zp > 0:
            buf = zero_stuff_upsample(buf, zp, zs)
        spec = self.to_phase(buf)
        codec.phase_dump_image_go(output_file, spec, self.num_freqs, self.y_reverse,
                                  float(original * self.num_freqs) / float(len(spec)), float(sr), self.ihsPasses(),
                                  self.HDR, device=self.device)

    def ToPhaseFlac(self, inputFile, outputFile):
        """phase.ToPhaseFlac (phase/phase.go:195-219): loadflac scales by 1/32768 (phase/impl.go:375)"""
        buf, sr = codec.load_flac_go(inputFile, 256 * 128)
        self._go_to_png(buf, sr, outputFile)

    def ToPhaseWav(self, inputFile, outputFile):
        """phase.ToPhaseWav (phase/phase.go:221-244)"""
        buf, sr = codec.load_wav(inputFile)
        self._go_to_png(buf, sr, outputFile)

    def ToWavPng(self, inputFile, outputFile):
        """phase.ToWavPng (phase/phase.go:246-275): Go PNG flavour (16 metadata bytes, blue-channel wrap), trim only if
        isPadded, SampleRate defaults to the family's main rate, dumpwav (truncating 16-bit PCM)"""
        buf, samples, samplerate = codec.phase_load_png_go(inputFile, self.y_reverse, self.ihsPasses(), self.HDR,
                                                           device=self.device)
        if len(buf) == 0:
            raise ErrFileNotLoaded()
        owave = self.FromPhase(buf)
        if int(samples) > 0 and is_padded(int(samples), len(owave), self.window) and len(owave) > int(samples):
            owave = owave[:int(samples)]
        main_rate = 44100 if self.num_freqs in (836, 836 * 2) else 48000
        if samplerate != 0 and not self.sample_rate:
            self.sample_rate = main_rate
        codec.save_wav(outputFile, owave, self.sample_rate or 0)

    def FromPhase(self, ospectrum):
        """phase.FromPhase (phase/phase.go:136-153): boost applied iff != 0 (phase/phase.go:146)"""
        return _lib.default_context(self.device).from_phase(self._cfg(self.volume_boost), ospectrum)

    def Image(self, buf):
        """Phase.Image (phase/phase.go:190-192) -> dumpbuffer (phase/impl.go:15-43)"""
        return _lib.default_context(self.device).image(buf, self.num_freqs)

    def ihsPasses(self):
        """phase/phase.go:31-36"""
        return 2 if (self.IHS and not self.HDR) else 0


# padShift (phase/impl.go:476-507)
GO_PAD_SHIFT = {48000: (0, 0), 32000: (2, 1), 24000: (1, 1), 16000: (1, 2), 8000: (1, 5),
                44100: (0, 0), 22050: (1, 1), 11025: (1, 3)}


def NewPhase():
    """phase.NewPhase (phase/phase.go:21-28): NumFreqs 768, Window 1280, Resolut 4096, VolumeBoost 0"""
    p = Phase(num_freqs=768, y_reverse=False)
    return p


# ---------------------------------------------------------------- module helpers (phase.py:352-852)
def shrink(spectrogram, resolut, num_freqs):
    """phase.py:430-435 / shrink (phase/impl.go:383-391)"""
    original_bins = resolut // 2
    spectrogram = np.asarray(spectrogram)
    time_frames = len(spectrogram) // original_bins
    return spectrogram.reshape(time_frames, original_bins, 2)[:, :num_freqs, :].reshape(-1, 2)


def grow(spectrogram, resolut, num_freqs):
    """phase.py:438-466 / grow (phase/impl.go:392-403): replicate the last kept bin upward"""
    target_bins = resolut // 2
    s = np.asarray(spectrogram).reshape(-1, num_freqs, 2)
    rep = np.repeat(s[:, -1:, :], target_bins - num_freqs, axis=1)
    return np.concatenate([s, rep], axis=1).reshape(-1, 2)


def zero_stuff_upsample(audio, zero_pad, zero_shift):
    """phase.py:503-549 / zeroStuffUpsample (phase/impl.go:509-529)"""
    if zero_pad == 0:
        return audio
    audio = np.asarray(audio)
    n = len(audio)
    num_groups = (n + zero_pad - 1) // zero_pad
    output = np.zeros(n + num_groups * zero_shift, dtype=audio.dtype)
    i = np.arange(n)
    output[i + (i // zero_pad) * zero_shift] = audio * (1 + zero_shift)
    return output


class ErrFileNotLoaded(Exception):
    """phase.ErrFileNotLoaded (phase/phase.go:38)"""

    def __init__(self):
        super().__init__("wavNotLoaded")


def load_wav_with_sr(file_path):
    """phase.py:551-567 (soundfile semantics: /2^(bits-1), channels averaged)"""
    return codec.load_wav_sf(file_path)


def load_flac_with_sr(file_path):
    """phase.py:570-586 (soundfile semantics; the FLAC container is decoded by gomel_b200/flac.py)"""
    return codec.load_flac_sf(file_path)


def load_wav(file_path):
    return codec.load_wav_sf(file_path)[0]


def load_flac(file_path):
    return codec.load_flac_sf(file_path)[0]


def save_wav(file_path, audio_buffer, sample_rate):
    """phase.py:589-601: clip to [-1,1], 16-bit PCM mono through libsndfile's rounding conversion"""
    codec.save_wav_sf(file_path, audio_buffer, sample_rate)


# ---- exported helpers of the Go package `phase` (phase/phase.go:155-188), Go codec flavour
def LoadFlac(inputFile):
    """phase.LoadFlac: samples / 32768, subframes appended block by block (phase/impl.go:351-381)"""
    return codec.load_flac_go(inputFile, 256 * 128)[0]


def LoadWav(inputFile):
    """phase.LoadWav: left channel through beep's decoder (phase/impl.go:309-349)"""
    return codec.load_wav(inputFile)[0]


def LoadFlacSampleRate(inputFile):
    """phase.LoadFlacSampleRate: (samples, sample rate) or ErrFileNotLoaded"""
    mono, sr = codec.load_flac_go(inputFile, 256 * 128)
    if len(mono) == 0 or sr == 0:
        raise ErrFileNotLoaded()
    return mono, int(sr)


def LoadWavSampleRate(inputFile):
    """phase.LoadWavSampleRate: (samples, sample rate) or ErrFileNotLoaded"""
    mono, sr = codec.load_wav(inputFile)
    if len(mono) == 0 or sr == 0:
        raise ErrFileNotLoaded()
    return mono, int(sr)


def SaveWav(outputFile, vec, sr):
    """phase.SaveWav -> dumpwav (phase/impl.go:280-307): truncating 16-bit PCM"""
    return codec.save_wav(outputFile, vec, sr)


pack_float16_to_bytes = codec.pack_f16_py
unpack_bytes_to_float64 = codec.unpack_f16
save_image = codec.phase_save_image_py
load_image = codec.phase_load_image_py
