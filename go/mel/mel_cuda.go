// mel_cuda.go -- drop-in bodies for the hot-path methods of package mel (reference mel/mel.go).
// A maintainer adds this file to the reference's mel/ directory and deletes the bodies of ToMel,
// FromMel and Image from mel.go (INTEGRATION.md); struct, NewMel, file APIs, codecs stay as they are.
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain).
package mel

import (
	"math"
	"math/rand"

	"github.com/neurlang/gomel/internal/gomelcuda"
)

// CudaConfig / LoadWavRate / LoadFlacRate / LoadPng / DumpImage / IsPadded export what cmd/gomelbatch needs of the
// package's private helpers (mel/impl.go); DumpWavPCM16 is dumpwav (mel/impl.go:195-232) for samples the GPU has
// already quantised (gomel_from_mel_batch_host_pcm16).
func (m *Mel) CudaConfig() gomelcuda.Config                 { return m.cudaConfig() }
func LoadWavRate(name string) ([]float64, float64)          { return loadwav(name) }
func LoadFlacRate(name string) ([]float64, float64)         { return loadflac(name) }
func LoadPng(name string, reverse bool) ([][2]float64, float64, float64) { return loadpng(name, reverse) }
func IsPadded(originalLen, paddedLen, filter int) bool      { return isPadded(originalLen, paddedLen, filter) }
func DumpImage(name string, buf [][2]float64, mels int, reverse bool, samplesInMel, sr float64) error {
	return dumpimage(name, buf, mels, reverse, samplesInMel, sr)
}
func DumpWavPCM16(name string, pcm []int16, sr int) error {
	data := make([]float64, len(pcm))
	for i, v := range pcm {
		data[i] = float64(v) / 32767.0 // beep re-quantises with int16(v * 32767): exact for every int16
	}
	return dumpwav(name, data, sr)
}

func (m *Mel) cudaConfig() gomelcuda.Config {
	return gomelcuda.Config{NFFT: m.Resolut, Hop: m.Window, NMels: m.NumMels, GLIters: m.GriffinLimIterations,
		TuneMul: m.TuneMul, TuneAdd: m.TuneAdd, VolumeBoost: m.VolumeBoost, MelFmin: m.MelFmin, MelFmax: m.MelFmax}
}

// ToMel replaces mel/mel.go:46-74.
func (m *Mel) ToMel(buf []float64) ([][2]float64, error) {
	ctx, err := gomelcuda.Default()
	if err != nil {
		return nil, err
	}
	cfg := m.cudaConfig()
	if err := ctx.SetMelTables(cfg, m.MelFmin, m.MelFmax); err != nil {
		return nil, err
	}
	return ctx.ToMel(cfg, buf)
}

// FromMel replaces mel/mel.go:142-152.  It keeps the reference's observable behaviour: the start
// signal is drawn from the global math/rand source exactly like mel/mel.go:80-83, and the caller's
// slice is exp()ed in place like spectral_denormalize (mel/impl.go:421-427).
func (m *Mel) FromMel(ospectrum [][2]float64) ([]float64, error) {
	ctx, err := gomelcuda.Default()
	if err != nil {
		return nil, err
	}
	cfg := m.cudaConfig()
	if err := ctx.SetMelTables(cfg, m.MelFmin, m.MelFmax); err != nil {
		return nil, err
	}
	frames := len(ospectrum) / m.NumMels
	init := make([]float64, m.Resolut+(frames-1)*m.Window)
	for i := range init {
		init[i] = rand.Float64()
	}
	out, err := ctx.FromMel(cfg, ospectrum, init)
	for l := 0; l < 2; l++ {
		for i := range ospectrum {
			ospectrum[i][l] = math.Exp(ospectrum[i][l])
		}
	}
	return out, err
}

// Image replaces mel/mel.go:171-173.
func (m *Mel) Image(buf [][2]float64) []uint16 {
	ctx, err := gomelcuda.Default()
	if err != nil {
		panic(err)
	}
	out, err := ctx.Image(buf, m.NumMels)
	if err != nil {
		panic(err)
	}
	return out
}
