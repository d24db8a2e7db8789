// gomel.hpp -- C++ host-side mirror of the reference's Go packages `mel` and `phase`
// (mel/mel.go, phase/phase.go) on top of the C ABI (include/gomel_cuda.h).
//
// The reference's host language is Go; this image has no Go toolchain, so the compiled-language
// host side lives here with the same type names, field names, method names, argument meaning and
// error behaviour (a Go `error` is an `Error` string, empty == nil).  The cgo files under go/ are
// the same call sequence written in Go.  All arithmetic runs in libgomelcuda.so on the GPU.
#pragma once
#include <array>
#include <cmath>
#include <mutex>
#include <random>
#include <set>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/gomel_cuda.h"

namespace gomel {

using Pair = std::array<double, 2>;      // Go [2]float64
using Error = std::string;               // Go error; "" == nil

namespace detail {
inline gomel_ctx* default_ctx(Error* err)
{
    static gomel_ctx* ctx = nullptr;
    static int rc = gomel_ctx_create(0, &ctx);
    if (rc != 0 && err) *err = "gomel_ctx_create failed (no CUDA device; there is no CPU fallback)";
    return rc == 0 ? ctx : nullptr;
}
inline double hz_to_mel(double v) { return 1127.0 * std::log(1.0 + (v / 700.0)); }    // mel/impl.go:304-308
inline double mel_to_hz(double v) { return 700.0 * (std::exp(v / 1127.0) - 1.0); }    // mel/impl.go:298-302
}  // namespace detail

namespace mel {

const Error ErrFileNotLoaded = "wavNotLoaded";       // mel/mel.go:43

struct Mel {                                         // mel/mel.go:10-27
    int NumMels = 160;
    double MelFmin = 0, MelFmax = 8000, TuneMul = 1, TuneAdd = 0;
    int Window = 256, Resolut = 2048;
    bool YReverse = false;
    int GriffinLimIterations = 2;
    double VolumeBoost = 0;
    int SampleRate = 0;
    // extension: start-signal injection for parity runs (empty = draw U[0,1) like math/rand)
    std::vector<double> InitSignal;
    // extension: GOMEL_FLAG_F64 -- every Griffin-Lim iteration in float64 (default: the library's precision policy,
    // max(16, iterations - 16) float64 lead iterations, then float32)
    bool Float64 = false;

    gomel_config config() const
    {
        gomel_config c{};
        c.n_fft = Resolut; c.hop = Window; c.n_mels = NumMels; c.gl_iters = GriffinLimIterations;
        c.tune_mul = TuneMul; c.tune_add = TuneAdd; c.volume_boost = VolumeBoost;
        c.mel_fmin = MelFmin; c.mel_fmax = MelFmax;        // key of this object's filterbank tables
        c.flags = Float64 ? GOMEL_FLAG_F64 : 0;
        return c;
    }

    // Registers this configuration's tables once per process (the library keeps them by key, so Mel objects with
    // different configurations can share the context from different threads); `force` after GOMEL_E_STATE.
    Error use_tables(gomel_ctx* ctx, const gomel_config& cfg, bool force = false) const
    {
        static std::mutex mu;
        static std::set<std::tuple<int, int, double, double>> done;
        const auto key = std::make_tuple(Resolut, NumMels, MelFmin, MelFmax);
        std::lock_guard<std::mutex> g(mu);
        if (!force && done.count(key)) return "";
        Error e = set_tables(ctx, cfg);
        if (e.empty()) { if (done.size() >= 48) done.clear(); done.insert(key); }
        return e;
    }

    // the (int(inlo), int(inhi), modlo) triples of domel / undomel, mel/impl.go:313-323, :350-360
    Error set_tables(gomel_ctx* ctx, const gomel_config& cfg) const
    {
        const int fs = Resolut / 2, mels = NumMels;
        std::vector<int> flo(mels), fhi(mels), ilo(fs), ihi(fs);
        std::vector<double> fmod(mels), imod(fs);
        const double melbin = detail::hz_to_mel(MelFmax) / double(mels);
        for (int i = 0; i < mels; i++) {
            const double vallo = double(fs) * (MelFmin + detail::mel_to_hz(melbin * double(i))) / (MelFmax + MelFmin);
            const double valhi = double(fs) * (MelFmin + detail::mel_to_hz(melbin * double(i + 1))) / (MelFmax + MelFmin);
            double inlo, modlo = std::modf(vallo, &inlo), inhi = std::floor(valhi);
            if (inlo < 0) { inlo = 0; modlo = 0; inhi = 0; }
            flo[i] = int(inlo); fhi[i] = int(inhi); fmod[i] = modlo;
        }
        for (int i = 0; i < fs; i++) {
            const double vallo = detail::hz_to_mel((double(i) * (MelFmax + MelFmin) / double(fs)) - MelFmin) / melbin;
            const double valhi = detail::hz_to_mel((double(i + 1) * (MelFmax + MelFmin) / double(fs)) - MelFmin) / melbin;
            double inlo, modlo = std::modf(vallo, &inlo), inhi = std::floor(valhi);
            if (inlo < 0) { inlo = 0; modlo = 0; inhi = 0; }
            ilo[i] = int(inlo); ihi[i] = int(inhi); imod[i] = modlo;
        }
        if (gomel_set_mel_tables(ctx, &cfg, flo.data(), fhi.data(), fmod.data(), ilo.data(), ihi.data(), imod.data()))
            return gomel_last_error(ctx);
        return "";
    }

    // func (m *Mel) ToMel(buf []float64) ([][2]float64, error)          mel/mel.go:46
    std::pair<std::vector<Pair>, Error> ToMel(const std::vector<double>& buf) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        if (!ctx) return { {}, err };
        const gomel_config cfg = config();
        long np = 0, frames = 0, ola = 0;
        if (gomel_frames(&cfg, (long)buf.size(), &np, &frames, &ola)) return { {}, "bad length" };
        if (!(err = use_tables(ctx, cfg)).empty()) return { {}, err };
        std::vector<Pair> out((size_t)frames * NumMels);
        int rc = gomel_to_mel(ctx, &cfg, buf.data(), (long)buf.size(), &out[0][0]);
        if (rc == GOMEL_E_STATE && use_tables(ctx, cfg, true).empty())      // set evicted meanwhile
            rc = gomel_to_mel(ctx, &cfg, buf.data(), (long)buf.size(), &out[0][0]);
        if (rc) return { {}, gomel_last_error(ctx) };
        return { std::move(out), "" };
    }

    // func (m *Mel) FromMel(ospectrum [][2]float64) ([]float64, error)  mel/mel.go:142
    // exp()s `ospectrum` in place like spectral_denormalize (mel/impl.go:421-427)
    std::pair<std::vector<double>, Error> FromMel(std::vector<Pair>& ospectrum) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        if (!ctx) return { {}, err };
        const gomel_config cfg = config();
        if (NumMels <= 0 || ospectrum.empty() || ospectrum.size() % (size_t)NumMels)
            return { {}, "len(ospectrum) is not a multiple of NumMels (the Go reference panics)" };
        if (!(err = use_tables(ctx, cfg)).empty()) return { {}, err };
        const long frames = (long)(ospectrum.size() / NumMels);
        const long ola = Resolut + (frames - 1) * (long)Window;
        std::vector<double> init = InitSignal;
        if (init.empty()) {                                       // rand.Float64() per sample, mel/mel.go:80-83
            static std::mt19937_64 gen{ std::random_device{}() };
            std::uniform_real_distribution<double> u(0.0, 1.0);
            init.resize((size_t)ola);
            for (auto& v : init) v = u(gen);
        }
        if ((long)init.size() != ola) return { {}, "InitSignal length != ola_len" };
        std::vector<double> out((size_t)ola);
        int rc = gomel_from_mel(ctx, &cfg, &ospectrum[0][0], frames, init.data(), 0, out.data());
        if (rc == GOMEL_E_STATE && use_tables(ctx, cfg, true).empty())
            rc = gomel_from_mel(ctx, &cfg, &ospectrum[0][0], frames, init.data(), 0, out.data());
        for (auto& p : ospectrum) { p[0] = std::exp(p[0]); p[1] = std::exp(p[1]); }
        if (rc) return { {}, gomel_last_error(ctx) };
        return { std::move(out), "" };
    }

    // func (m *Mel) Image(buf [][2]float64) []uint16                     mel/mel.go:171
    std::vector<unsigned short> Image(const std::vector<Pair>& buf) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        std::vector<unsigned short> out(buf.size() / (size_t)NumMels * (size_t)NumMels);
        if (!ctx || gomel_image(ctx, &buf[0][0], (long)buf.size(), NumMels, out.data(), nullptr)) out.clear();
        return out;
    }
};

inline Mel* NewMel() { return new Mel(); }           // mel/mel.go:30-41

}  // namespace mel

namespace phase {

struct Phase {                                       // phase/phase.go:8-18
    int NumFreqs = 768, Window = 1280, Resolut = 4096;
    bool YReverse = false;
    int SampleRate = 0;
    double VolumeBoost = 0;
    bool IHS = false, HDR = false;

    int ihsPasses() const { return (IHS && !HDR) ? 2 : 0; }       // phase/phase.go:31-36

    gomel_config config() const
    {
        gomel_config c{};
        c.n_fft = Resolut; c.hop = Window; c.n_freqs = NumFreqs; c.tune_mul = 1; c.volume_boost = VolumeBoost;
        return c;
    }

    // func (m *Phase) ToPhase(buf []float64) ([][2]float64, error)       phase/phase.go:41
    std::pair<std::vector<Pair>, Error> ToPhase(const std::vector<double>& buf) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        if (!ctx) return { {}, err };
        const gomel_config cfg = config();
        long np = 0, frames = 0, ola = 0;
        if (gomel_frames(&cfg, (long)buf.size(), &np, &frames, &ola)) return { {}, "bad length" };
        std::vector<Pair> out((size_t)frames * NumFreqs);
        if (gomel_to_phase(ctx, &cfg, buf.data(), (long)buf.size(), &out[0][0])) return { {}, gomel_last_error(ctx) };
        return { std::move(out), "" };
    }

    // func (m *Phase) FromPhase(ospectrum [][2]float64) ([]float64, error)   phase/phase.go:136
    std::pair<std::vector<double>, Error> FromPhase(const std::vector<Pair>& ospectrum) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        if (!ctx) return { {}, err };
        const gomel_config cfg = config();
        if (NumFreqs <= 0 || ospectrum.empty() || ospectrum.size() % (size_t)NumFreqs)
            return { {}, "len(ospectrum) is not a multiple of NumFreqs" };
        const long frames = (long)(ospectrum.size() / NumFreqs);
        std::vector<double> out((size_t)(Resolut + (frames - 1) * (long)Window));
        if (gomel_from_phase(ctx, &cfg, &ospectrum[0][0], frames, out.data())) return { {}, gomel_last_error(ctx) };
        return { std::move(out), "" };
    }

    // func (m *Phase) Image(buf [][2]float64) []uint16                    phase/phase.go:190
    std::vector<unsigned short> Image(const std::vector<Pair>& buf) const
    {
        Error err;
        gomel_ctx* ctx = detail::default_ctx(&err);
        std::vector<unsigned short> out(buf.size() / (size_t)NumFreqs * (size_t)NumFreqs);
        if (!ctx || gomel_image(ctx, &buf[0][0], (long)buf.size(), NumFreqs, out.data(), nullptr)) out.clear();
        return out;
    }
};

inline Phase* NewPhase() { return new Phase(); }     // phase/phase.go:21-28

}  // namespace phase
}  // namespace gomel
