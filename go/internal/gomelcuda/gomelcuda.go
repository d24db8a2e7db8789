// Package gomelcuda is the thin cgo layer over libgomelcuda.so (include/gomel_cuda.h).
// NOT COMPILED IN THE BUILD IMAGE (no Go toolchain there) -- written against the C header by
// inspection; see INTEGRATION.md.  One cgo crossing per API call, never per frame.
package gomelcuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../gomel_b200 -lgomelcuda -Wl,-rpath,${SRCDIR}/../../../gomel_b200
#include <stdlib.h>
#include "gomel_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math"
	"sync"
	"unsafe"
)

// Config mirrors gomel_config.
type Config struct {
	NFFT, Hop, NMels, NFreqs, GLIters int
	TuneMul, TuneAdd, VolumeBoost     float64
	MelFmin, MelFmax                  float64 // key of the filterbank tables (with NFFT, NMels)
	Flags                             int     // FlagF64: every Griffin-Lim iteration in float64 (GOMEL_FLAG_F64)
}

const (
	FlagF64    = 1 // GOMEL_FLAG_F64
	FlagF64Ref = 2 // GOMEL_FLAG_F64_REF (test instrument)
)

// oneZero stands in for an empty input: the reference's pad() grows an empty buffer to 15*Window-1 zeros
// (mel/impl.go:429-455) and one zero sample pads to exactly the same signal, while &wav[0] of an empty slice
// would panic before the cgo call.
var oneZero = []float64{0}

func (c Config) c() C.gomel_config {
	return C.gomel_config{n_fft: C.int(c.NFFT), hop: C.int(c.Hop), n_mels: C.int(c.NMels), n_freqs: C.int(c.NFreqs),
		gl_iters: C.int(c.GLIters), tune_mul: C.double(c.TuneMul), tune_add: C.double(c.TuneAdd),
		volume_boost: C.double(c.VolumeBoost), flags: C.int(c.Flags), mel_fmin: C.double(c.MelFmin), mel_fmax: C.double(c.MelFmax)}
}

// Ctx owns one gomel_ctx.  Methods are safe for concurrent use: the library serialises calls per context
// and keeps the mel filterbank tables by key (NFFT, NMels, MelFmin, MelFmax), so goroutines with different
// Mel configurations can share one Ctx.  Use several Ctx for GPU-side concurrency.
type Ctx struct {
	h          *C.gomel_ctx
	mu         sync.Mutex
	registered map[[4]float64]bool
}

var (
	defOnce sync.Once
	defCtx  *Ctx
	defErr  error
)

// Default returns the process-wide context on device 0.  There is no CPU fallback: without a GPU
// this returns an error and every transform fails.
func Default() (*Ctx, error) {
	defOnce.Do(func() { defCtx, defErr = New(0) })
	return defCtx, defErr
}

func New(device int) (*Ctx, error) {
	var h *C.gomel_ctx
	if rc := C.gomel_ctx_create(C.int(device), &h); rc != 0 {
		return nil, fmt.Errorf("gomel_ctx_create: %d (no CUDA device?)", int(rc))
	}
	return &Ctx{h: h, registered: map[[4]float64]bool{}}, nil
}

func (x *Ctx) Close() { C.gomel_ctx_destroy(x.h); x.h = nil }

func (x *Ctx) err(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.gomel_last_error(x.h)))
}

// Frames = pad (mel/impl.go:429-455) + gossp NumFrames + ISTFT length.
func Frames(cfg Config, n int) (padded, frames, ola int, err error) {
	cc := cfg.c()
	var a, b, c C.long
	if rc := C.gomel_frames(&cc, C.long(n), &a, &b, &c); rc != 0 {
		return 0, 0, 0, errors.New("gomel_frames: bad length")
	}
	return int(a), int(b), int(c), nil
}

func hzToMel(v float64) float64 { return 1127.0 * math.Log(1.0+(v/700.0)) }   // mel/impl.go:304-308
func melToHz(v float64) float64 { return 700.0 * (math.Exp(v/1127.0) - 1.0) } // mel/impl.go:298-302

// SetMelTables computes the (int(inlo), int(inhi), modlo) triples of domel / undomel with Go's own
// math.Exp / math.Log -- exactly the expressions of mel/impl.go:313-323 and :350-360 -- and hands
// them to the library, so the two 1-ulp-fragile band edges fall where the reference puts them.
// Registered once per key; cfg.MelFmin / cfg.MelFmax of the compute calls select the set.
func (x *Ctx) SetMelTables(cfg Config, fmin, fmax float64) error {
	return x.setMelTables(cfg, fmin, fmax, false)
}

// force: register again even if this Ctx believes the key is registered (the library evicts the least
// recently used of its 64 table sets; a mel call then returns GOMEL_E_STATE and is retried once).
func (x *Ctx) setMelTables(cfg Config, fmin, fmax float64, force bool) error {
	key := [4]float64{float64(cfg.NFFT), float64(cfg.NMels), fmin, fmax}
	x.mu.Lock()
	defer x.mu.Unlock()
	if x.registered[key] && !force {
		return nil
	}
	cfg.MelFmin, cfg.MelFmax = fmin, fmax
	fs, mels := cfg.NFFT/2, cfg.NMels
	flo, fhi, fmod := make([]C.int, mels), make([]C.int, mels), make([]C.double, mels)
	melbin := hzToMel(fmax) / float64(mels)
	for i := 0; i < mels; i++ {
		vallo := float64(fs) * (fmin + melToHz(melbin*float64(i))) / (fmax + fmin)
		valhi := float64(fs) * (fmin + melToHz(melbin*float64(i+1))) / (fmax + fmin)
		inlo, modlo := math.Modf(vallo)
		inhi := math.Floor(valhi)
		if inlo < 0 {
			inlo, modlo, inhi = 0, 0, 0
		}
		flo[i], fhi[i], fmod[i] = C.int(int(inlo)), C.int(int(inhi)), C.double(modlo)
	}
	ilo, ihi, imod := make([]C.int, fs), make([]C.int, fs), make([]C.double, fs)
	filterbin := hzToMel(fmax) / float64(mels)
	for i := 0; i < fs; i++ {
		vallo := hzToMel((float64(i)*(fmax+fmin)/float64(fs))-fmin) / filterbin
		valhi := hzToMel((float64(i+1)*(fmax+fmin)/float64(fs))-fmin) / filterbin
		inlo, modlo := math.Modf(vallo)
		inhi := math.Floor(valhi)
		if inlo < 0 {
			inlo, modlo, inhi = 0, 0, 0
		}
		ilo[i], ihi[i], imod[i] = C.int(int(inlo)), C.int(int(inhi)), C.double(modlo)
	}
	cc := cfg.c()
	if e := x.err(C.gomel_set_mel_tables(x.h, &cc, &flo[0], &fhi[0], &fmod[0], &ilo[0], &ihi[0], &imod[0])); e != nil {
		return e
	}
	if len(x.registered) >= 48 { // the library keeps 64 sets (LRU); stay below that
		x.registered = map[[4]float64]bool{}
	}
	x.registered[key] = true
	return nil
}

// ToMel: [][2]float64 is a contiguous double[2n], passed as &out[0] (no Go pointer to Go pointer).
func (x *Ctx) ToMel(cfg Config, wav []float64) ([][2]float64, error) {
	if len(wav) == 0 {
		wav = oneZero
	}
	_, frames, _, err := Frames(cfg, len(wav))
	if err != nil {
		return nil, err
	}
	out := make([][2]float64, frames*cfg.NMels)
	cc := cfg.c()
	call := func() C.int {
		return C.gomel_to_mel(x.h, &cc, (*C.double)(unsafe.Pointer(&wav[0])), C.long(len(wav)),
			(*C.double)(unsafe.Pointer(&out[0])))
	}
	rc := call()
	if rc == C.GOMEL_E_STATE && x.setMelTables(cfg, cfg.MelFmin, cfg.MelFmax, true) == nil {
		rc = call()
	}
	return out, x.err(rc)
}

func (x *Ctx) FromMel(cfg Config, mel [][2]float64, init []float64) ([]float64, error) {
	if cfg.NMels <= 0 || len(mel) == 0 || len(mel)%cfg.NMels != 0 {
		// the reference strides by NumMels without a length check and panics (mel/impl.go:366-372)
		return nil, errors.New("gomel: len(ospectrum) is not a positive multiple of NumMels")
	}
	frames := len(mel) / cfg.NMels
	out := make([]float64, cfg.NFFT+(frames-1)*cfg.Hop)
	if len(init) != 0 && len(init) != len(out) {
		return nil, errors.New("gomel: start signal length != Resolut + (frames-1)*Window")
	}
	cc := cfg.c()
	var ip *C.double
	if len(init) > 0 {
		ip = (*C.double)(unsafe.Pointer(&init[0]))
	}
	call := func() C.int {
		return C.gomel_from_mel(x.h, &cc, (*C.double)(unsafe.Pointer(&mel[0])), C.long(frames), ip, 0,
			(*C.double)(unsafe.Pointer(&out[0])))
	}
	rc := call()
	if rc == C.GOMEL_E_STATE && x.setMelTables(cfg, cfg.MelFmin, cfg.MelFmax, true) == nil {
		rc = call()
	}
	return out, x.err(rc)
}

func (x *Ctx) ToPhase(cfg Config, wav []float64) ([][2]float64, error) {
	if len(wav) == 0 {
		wav = oneZero
	}
	_, frames, _, err := Frames(cfg, len(wav))
	if err != nil {
		return nil, err
	}
	out := make([][2]float64, frames*cfg.NFreqs)
	cc := cfg.c()
	rc := C.gomel_to_phase(x.h, &cc, (*C.double)(unsafe.Pointer(&wav[0])), C.long(len(wav)),
		(*C.double)(unsafe.Pointer(&out[0])))
	return out, x.err(rc)
}

func (x *Ctx) FromPhase(cfg Config, spec [][2]float64) ([]float64, error) {
	if cfg.NFreqs <= 0 || len(spec) == 0 || len(spec)%cfg.NFreqs != 0 {
		return nil, errors.New("gomel: len(ospectrum) is not a positive multiple of NumFreqs")
	}
	frames := len(spec) / cfg.NFreqs
	out := make([]float64, cfg.NFFT+(frames-1)*cfg.Hop)
	cc := cfg.c()
	rc := C.gomel_from_phase(x.h, &cc, (*C.double)(unsafe.Pointer(&spec[0])), C.long(frames),
		(*C.double)(unsafe.Pointer(&out[0])))
	return out, x.err(rc)
}

// SetGLPrecision: at least `lead` float64 Griffin-Lim iterations first, at most `tail` float32 ones at the end
// (tail < 0: unlimited).  Library defaults 16 and 16; (0, -1) is the all-float32 loop.
func (x *Ctx) SetGLPrecision(lead, tail int) {
	C.gomel_set_lead_f64(x.h, C.int(lead))
	C.gomel_set_f32_tail(x.h, C.int(tail))
}

// SetGLGuard sets the leverage threshold of the float32 tail's singular-bin guard (library default 5e4; 0 disables):
// clips whose float32 iterations meet a bin with M/|X| that large have those iterations re-run in float64.
func (x *Ctx) SetGLGuard(threshold float32) error {
	return x.err(C.gomel_set_gl_guard(x.h, C.float(threshold), nil))
}

// LastGLGuard reports the guard's record of the last Griffin-Lim call: clips seen, clips re-run, largest leverage.
func (x *Ctx) LastGLGuard() (seen, rerun int, maxLeverage float32, err error) {
	var n, r C.int
	var m C.float
	rc := C.gomel_last_gl_guard(x.h, &n, &r, &m, nil, 0)
	return int(n), int(r), float32(m), x.err(rc)
}

func (x *Ctx) Image(buf [][2]float64, mels int) ([]uint16, error) {
	if mels <= 0 || len(buf) < mels {
		return nil, errors.New("gomel: fewer entries than one column")
	}
	out := make([]uint16, (len(buf)/mels)*mels)
	rc := C.gomel_image(x.h, (*C.double)(unsafe.Pointer(&buf[0])), C.long(len(buf)), C.int(mels),
		(*C.ushort)(unsafe.Pointer(&out[0])), nil)
	return out, x.err(rc)
}

// ToMelBatch runs gomel_to_mel_batch_host on clips of any lengths: ToMel frames depend only on local samples, so
// every clip is zero-extended to the longest one and only its own frames are kept (float32 across the boundary).
func ToMelBatch(cfg Config, fmin, fmax float64, clips [][]float64) ([][][2]float64, error) {
	x, err := Default()
	if err != nil {
		return nil, err
	}
	if err := x.SetMelTables(cfg, fmin, fmax); err != nil {
		return nil, err
	}
	cfg.MelFmin, cfg.MelFmax = fmin, fmax
	nmax := 0
	for _, c := range clips {
		if len(c) > nmax {
			nmax = len(c)
		}
	}
	if nmax == 0 {
		return nil, errors.New("gomel: no samples")
	}
	_, frMax, _, err := Frames(cfg, nmax)
	if err != nil {
		return nil, err
	}
	wav := make([]float32, len(clips)*nmax)
	for i, c := range clips {
		for j, v := range c {
			wav[i*nmax+j] = float32(v)
		}
	}
	per := frMax * cfg.NMels * 2
	out := make([]float32, len(clips)*per)
	cc := cfg.c()
	if e := x.err(C.gomel_to_mel_batch_host(x.h, &cc, (*C.float)(unsafe.Pointer(&wav[0])), C.int(len(clips)), C.long(nmax),
		(*C.float)(unsafe.Pointer(&out[0])), 0)); e != nil {
		return nil, e
	}
	res := make([][][2]float64, len(clips))
	for i, c := range clips {
		_, fr, _, _ := Frames(cfg, len(c))
		spec := make([][2]float64, fr*cfg.NMels)
		for k := range spec {
			spec[k] = [2]float64{float64(out[i*per+2*k]), float64(out[i*per+2*k+1])}
		}
		res[i] = spec
	}
	return res, nil
}

// FromMelBatchPCM16 runs gomel_from_mel_batch_host_pcm16 on spectrograms that all have `frames` frames; the start
// signals are drawn on the device (U[0,1), seeded).  Returns the 16-bit samples dumpwav would write.
func FromMelBatchPCM16(cfg Config, fmin, fmax float64, specs [][][2]float64, frames int) ([][]int16, error) {
	x, err := Default()
	if err != nil {
		return nil, err
	}
	if err := x.SetMelTables(cfg, fmin, fmax); err != nil {
		return nil, err
	}
	cfg.MelFmin, cfg.MelFmax = fmin, fmax
	per := frames * cfg.NMels * 2
	mel := make([]float32, len(specs)*per)
	for i, s := range specs {
		if len(s) != frames*cfg.NMels {
			return nil, errors.New("gomel: spectrogram length != frames * NumMels")
		}
		for k, e := range s {
			mel[i*per+2*k], mel[i*per+2*k+1] = float32(e[0]), float32(e[1])
		}
	}
	ola := cfg.NFFT + (frames-1)*cfg.Hop
	pcm := make([]int16, len(specs)*ola)
	cc := cfg.c()
	if e := x.err(C.gomel_from_mel_batch_host_pcm16(x.h, &cc, (*C.float)(unsafe.Pointer(&mel[0])), C.int(len(specs)),
		C.long(frames), nil, 0, (*C.short)(unsafe.Pointer(&pcm[0])), 0)); e != nil {
		return nil, e
	}
	res := make([][]int16, len(specs))
	for i := range specs {
		res[i] = pcm[i*ola : (i+1)*ola]
	}
	return res, nil
}
