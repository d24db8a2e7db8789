// phase_cuda.go -- drop-in bodies for the hot-path methods of package phase (reference
// phase/phase.go).  Add to the reference's phase/ directory and delete the bodies of ToPhase,
// FromPhase and Image from phase.go (INTEGRATION.md).  NOT COMPILED IN THE BUILD IMAGE.
package phase

import "github.com/neurlang/gomel/internal/gomelcuda"

func (m *Phase) cudaConfig() gomelcuda.Config {
	return gomelcuda.Config{NFFT: m.Resolut, Hop: m.Window, NFreqs: m.NumFreqs, TuneMul: 1, VolumeBoost: m.VolumeBoost}
}

// ToPhase replaces phase/phase.go:41-70.
func (m *Phase) ToPhase(buf []float64) ([][2]float64, error) {
	ctx, err := gomelcuda.Default()
	if err != nil {
		return nil, err
	}
	return ctx.ToPhase(m.cudaConfig(), buf)
}

// FromPhase replaces phase/phase.go:136-153 (VolumeBoost applied iff != 0, by the library).
func (m *Phase) FromPhase(ospectrum [][2]float64) ([]float64, error) {
	ctx, err := gomelcuda.Default()
	if err != nil {
		return nil, err
	}
	return ctx.FromPhase(m.cudaConfig(), ospectrum)
}

// Image replaces phase/phase.go:190-192.
func (m *Phase) Image(buf [][2]float64) []uint16 {
	ctx, err := gomelcuda.Default()
	if err != nil {
		panic(err)
	}
	out, err := ctx.Image(buf, m.NumFreqs)
	if err != nil {
		panic(err)
	}
	return out
}
