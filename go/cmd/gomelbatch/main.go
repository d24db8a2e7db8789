// Command gomelbatch is the directory-level companion of the reference's cmd/tomel and cmd/towav
// (SURVEY 8(f) row f4): many files per GPU call through the batched C ABI.
//
//	gomelbatch tomel  <in_dir> <out_dir>     every *.wav / *.flac -> <name>.png   (mel.ToMelWav / ToMelFlac per file)
//	gomelbatch towav  <in_dir> <out_dir>     every *.png          -> <name>.wav   (mel.ToWavPng per file)
//
// The four single-file tools of the reference (cmd/tomel, cmd/towav, cmd/tophase, cmd/fromphase) need no change:
// they call mel.NewMel / phase.NewPhase and the file methods, whose hot-path bodies are the cgo bodies of
// mel_cuda.go / phase_cuda.go once those files are dropped in (INTEGRATION.md).
//
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain there) -- written against include/gomel_cuda.h by inspection.
package main

import (
	"fmt"
	"os"
	"path/filepath"
	"sort"
	"strings"

	"github.com/neurlang/gomel/internal/gomelcuda"
	"github.com/neurlang/gomel/mel"
)

// the configuration cmd/tomel and cmd/towav hard-code (cmd/tomel/main.go:24-31, cmd/towav/main.go:30-39)
func newMel() *mel.Mel {
	m := mel.NewMel()
	m.MelFmin, m.MelFmax, m.YReverse = 0, 16000, true
	m.Window, m.NumMels, m.Resolut = 1280, 192, 4096
	m.GriffinLimIterations, m.VolumeBoost = 2, 0.0
	return m
}

func list(dir string, exts ...string) []string {
	var out []string
	for _, e := range exts {
		g, _ := filepath.Glob(filepath.Join(dir, "*"+e))
		out = append(out, g...)
	}
	sort.Strings(out)
	return out
}

func main() {
	if len(os.Args) < 4 {
		fmt.Println("Usage: gomelbatch tomel|towav <in_dir> <out_dir>")
		os.Exit(1)
	}
	tool, in, out := os.Args[1], os.Args[2], os.Args[3]
	if err := os.MkdirAll(out, 0o755); err != nil {
		fmt.Println(err)
		os.Exit(1)
	}
	m := newMel()
	switch tool {
	case "tomel":
		// decode on the host, one batched ToMel per chunk of clips (gomel_to_mel_batch_host), PNG encode per file
		files := list(in, ".wav", ".flac")
		const chunk = 64
		for c0 := 0; c0 < len(files); c0 += chunk {
			c1 := c0 + chunk
			if c1 > len(files) {
				c1 = len(files)
			}
			var clips [][]float64
			var names []string
			var rates []float64
			for _, f := range files[c0:c1] {
				var buf []float64
				var sr float64
				if strings.HasSuffix(f, ".flac") {
					buf, sr = mel.LoadFlacRate(f)
				} else {
					buf, sr = mel.LoadWavRate(f)
				}
				if len(buf) == 0 {
					fmt.Printf("skipping %s: %v\n", f, mel.ErrFileNotLoaded)
					continue
				}
				clips, names, rates = append(clips, buf), append(names, f), append(rates, sr)
			}
			specs, err := gomelcuda.ToMelBatch(m.CudaConfig(), m.MelFmin, m.MelFmax, clips)
			if err != nil {
				fmt.Printf("Error generating mel spectrograms: %v\n", err)
				os.Exit(1)
			}
			for i, spec := range specs {
				dst := filepath.Join(out, filepath.Base(names[i])+".png")
				mel.DumpImage(dst, spec, m.NumMels, m.YReverse, float64(len(clips[i])*m.NumMels)/float64(len(spec)), rates[i])
			}
		}
	case "towav":
		// PNGs grouped by frame count (Griffin-Lim couples neighbouring frames), one batched FromMel per group
		// (gomel_from_mel_batch_host_pcm16: the waveforms come back as the 16-bit samples dumpwav writes)
		groups := map[int][]string{}
		bufs := map[string][][2]float64{}
		meta := map[string][2]float64{}
		for _, f := range list(in, ".png") {
			buf, samples, sr := mel.LoadPng(f, m.YReverse)
			if len(buf) == 0 || len(buf)%m.NumMels != 0 {
				fmt.Printf("skipping %s\n", f)
				continue
			}
			for i := range buf {
				buf[i][0] += m.VolumeBoost
				buf[i][1] += m.VolumeBoost
			}
			fr := len(buf) / m.NumMels
			groups[fr] = append(groups[fr], f)
			bufs[f], meta[f] = buf, [2]float64{samples, sr}
		}
		for fr, names := range groups {
			specs := make([][][2]float64, len(names))
			for i, f := range names {
				specs[i] = bufs[f]
			}
			pcm, err := gomelcuda.FromMelBatchPCM16(m.CudaConfig(), m.MelFmin, m.MelFmax, specs, fr)
			if err != nil {
				fmt.Printf("Error generating waves: %v\n", err)
				os.Exit(1)
			}
			for i, f := range names {
				w := pcm[i]
				if s := int(meta[f][0]); s > 0 && mel.IsPadded(s, len(w), m.Window) && len(w) > s {
					w = w[:s]
				}
				rate := m.SampleRate
				if rate == 0 {
					rate = int(meta[f][1])
				}
				mel.DumpWavPCM16(filepath.Join(out, filepath.Base(f)+".wav"), w, rate)
			}
		}
	default:
		fmt.Println("Usage: gomelbatch tomel|towav <in_dir> <out_dir>")
		os.Exit(1)
	}
}
