// abi.cu -- C ABI of libgomelcuda.so (include/gomel_cuda.h): context, tables, launch logic.
// No PyTorch, no cuFFT, no CPU fallback: every transform below is one of the kernels in
// kernels.cuh.  Host code here only sizes, stages and launches.
#include "../../include/gomel_cuda.h"
#include "kernels.cuh"
#include "kernels_f64.cuh"
#include "gl_f64.cuh"

#include <cmath>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace gomel;

namespace {

constexpr int kHS = 5;                       // hop slots: Window / 256 = 1280 / 256
constexpr int kHop = 256 * kHS;
constexpr int kHalo = (16 - kHS) * 256;      // Resolut - Window = 2816 samples shared by adjacent tiles

// The frame geometry of one call.  Native: Resolut 4096 / Window 1280 (every cmd/* tool of the reference).
// alt: Resolut 2048 / Window 256, the defaults of mel.NewMel (mel/mel.go:37-38), on the mel paths only; it
// runs on the same 4096-point core with the frame zero-extended (even core bins = the 2048-point spectrum).
struct Geo {
    int n_fft = kN, hop = kHop, halo = kHalo;
    bool alt = false;
    long ola(long n_frames) const { return n_fft + (n_frames - 1) * (long)hop; }
};
constexpr int kAltHS = 1, kAltFS = 8;
constexpr int kGlCtasPerSm = 2;

// One registered set of domel / undomel tables.  Keyed by (Resolut, NumMels, MelFmin, MelFmax) so that callers
// with different Mel configurations can share a context concurrently (set + compute are two calls).
struct MelTables {
    int n_fft = 0, n_mels = 0;
    double fmin = 0, fmax = 0;
    int *fwd_lo = nullptr, *fwd_hi = nullptr, *inv_lo = nullptr, *inv_hi = nullptr;
    float* fwd_mod = nullptr;
    double* inv_mod = nullptr;
    void release()
    {
        cudaFree(fwd_lo); cudaFree(fwd_hi); cudaFree(fwd_mod); cudaFree(inv_lo); cudaFree(inv_hi); cudaFree(inv_mod);
        fwd_lo = fwd_hi = inv_lo = inv_hi = nullptr; fwd_mod = nullptr; inv_mod = nullptr;
    }
};
constexpr size_t kMaxMelTables = 64;
static_assert(sizeof(gomel_config) == 72, "gomel_config layout is part of the ABI (ctypes / cgo mirror it)");

enum Scratch { S_F64IN = 0, S_SIG64A, S_SIG64B, S_Y64, S_MAGS64, S_F64IN2, S_F64OUT, S_F32A, S_F32B, S_SIGTMP, S_INIT, S_HB0, S_HB1, S_MAGS,
               S_MISC, S_LSIGA, S_LSIGB, S_LHB0, S_LHB1, S_LMAGS, S_GUARD, S_GHB0, S_GHB1, S_PIPE, S_COUNT = S_PIPE + 3 * 8 };
// chunk buffers of the pipelined host batches: kPipeSets sets of (mel | signal in, out, init, pcm, mags32, mags64)
constexpr int kPipeSets = 3;
enum PipeKind { P_IN = 0, P_OUT, P_INIT, P_PCM, P_MAGS32, P_MAGS64 };
constexpr int pipe_slot(int set, int kind) { return S_PIPE + set * 8 + kind; }
// Griffin-Lim precision policy (profiles/r02_gl_parity_sweep.md): at least kDefaultLeadF64 float64 iterations
// first, and at most kDefaultF32Tail float32 iterations at the end -> lead = max(16, iters - 16)
constexpr int kDefaultLeadF64 = 16;
constexpr int kDefaultF32Tail = 16;
// Singular-bin guard of the float32 tail (profiles/r02_gl_guard.md): clips whose statistic exceeds this many clip-rms
// units have their tail re-run in float64.  0 disables.
constexpr float kDefaultGlGuard = 5.0e4f;
constexpr int kGuardSyncClips = 4;
constexpr int kGuardRefFrames = 342;

}  // namespace

struct gomel_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr, st_h2d = nullptr, st_d2h = nullptr;
    // Griffin-Lim batches are split by clip over st + these streams: the iterations of one group of clips only
    // depend on that group, so the launch tail of one group is covered by the other groups' launches
    cudaStream_t st_gl[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t ev_fork = nullptr, ev_join[3] = { nullptr, nullptr, nullptr };
    cudaStream_t st_pre = nullptr;    // batch pipeline: magnitudes / start signal of the next chunk, beside the iterations
    int gl_streams = 2;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_l0 = nullptr, ev_l1 = nullptr;
    int hot_launches = 0;     // launches of the dominant kernel bracketed by ev_k0/ev_k1
    int lead_launches = 0;    // float64 lead iterations of the last Griffin-Lim, bracketed by ev_l0/ev_l1
    int lead_f64 = kDefaultLeadF64;   // gomel_set_lead_f64 / GOMEL_LEAD_F64
    int f32_tail = kDefaultF32Tail;   // gomel_set_f32_tail / GOMEL_F32_TAIL; < 0: unlimited
    float gl_guard = kDefaultGlGuard; // gomel_set_gl_guard / GOMEL_GL_GUARD
    int* h_guard_count = nullptr;     // pinned: the selection count of a small call (read back instead of launching blind)
    int guard_clips = 0;              // clips of the last guarded Griffin-Lim run (statistics in scratch[S_GUARD]); 0: none
    float guard_thr_units = 0;        // the threshold that run used, in statistic units
    float guard_to_leverage = 0;      // statistic / clip scale -> leverage, for that run
    double* d_tables_d64 = nullptr;   // gl_f64.cuh tables (built on first use)
    double* d_tables_d64_alt = nullptr;   // same twiddles, Hann window of the 2048-sample frame
    float4* d_tables = nullptr;
    float4* d_tables_alt = nullptr;   // same twiddles, Hann window of the 2048-sample frame
    double* d_tables64 = nullptr;     // strict float64 path (built on first use)
    // mel tables
    std::vector<MelTables> mel_tabs;      // registered filterbank table sets, most recently used last
    // phase gain tables
    float *d_gain_head = nullptr, *d_gain_mid = nullptr, *d_gain_tail = nullptr;
    long gain_frames_key = -1; double gain_boost_key = 0; int gain_head_len = 0, gain_tail_len = 0;
    void* scratch[S_COUNT] = {};
    size_t scratch_cap[S_COUNT] = {};
    unsigned long long launches = 0;
    int tile_override = 0;
    int gl_tile_waves = 4;            // Griffin-Lim: CTA waves per iteration the automatic tiling aims at
    std::string err;
    std::mutex mu;
};

namespace {

int fail(gomel_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? GOMEL_E_NOMEM : GOMEL_E_CUDA,         \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    } while (0)

// The table set of a config: exact (Resolut, NumMels, MelFmin, MelFmax) match; a config that leaves
// mel_fmin = mel_fmax = 0 selects the most recently registered set for its (Resolut, NumMels).
const MelTables* find_mel_tables(gomel_ctx* ctx, const gomel_config* cfg)
{
    const bool any = cfg->mel_fmin == 0 && cfg->mel_fmax == 0;
    for (size_t i = ctx->mel_tabs.size(); i-- > 0;) {
        const MelTables& t = ctx->mel_tabs[i];
        if (t.n_fft == cfg->n_fft && t.n_mels == cfg->n_mels && (any || (t.fmin == cfg->mel_fmin && t.fmax == cfg->mel_fmax))) {
            if (i + 1 != ctx->mel_tabs.size()) {            // most recently used last
                MelTables m = t;
                ctx->mel_tabs.erase(ctx->mel_tabs.begin() + (long)i);
                ctx->mel_tabs.push_back(m);
            }
            return &ctx->mel_tabs.back();
        }
    }
    fail(ctx, GOMEL_E_STATE, "mel tables not set for this Resolut / NumMels / MelFmin / MelFmax (gomel_set_mel_tables)");
    return nullptr;
}

int ensure(gomel_ctx* ctx, int slot, size_t bytes, void** out)
{
    if (ctx->scratch_cap[slot] < bytes) {
        if (ctx->scratch[slot]) { CU(cudaStreamSynchronize(ctx->st)); CU(cudaFree(ctx->scratch[slot])); }
        ctx->scratch[slot] = nullptr; ctx->scratch_cap[slot] = 0;
        size_t cap = bytes + (bytes >> 3) + 256;
        CU(cudaMalloc(&ctx->scratch[slot], cap));
        ctx->scratch_cap[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return 0;
}

int check_cfg(gomel_ctx* ctx, const gomel_config* cfg, Geo* geo = nullptr)
{
    if (!cfg) return fail(ctx, GOMEL_E_ARG, "config is NULL");
    if (cfg->n_fft == kN && cfg->hop == kHop) { if (geo) *geo = Geo(); return 0; }
    if (geo && cfg->n_fft == 256 * kAltFS && cfg->hop == 256 * kAltHS) {
        geo->n_fft = cfg->n_fft; geo->hop = cfg->hop; geo->halo = cfg->n_fft - cfg->hop; geo->alt = true;
        return 0;
    }
    return fail(ctx, GOMEL_E_UNSUPPORTED, geo ? "this build supports Resolut/Window = 4096/1280 and 2048/256 only"
                                              : "this entry point supports Resolut=4096, Window=1280 only");
}

long pad_len(long n, int filter)            // mel/impl.go:429-455
{
    const long min_target = 15L * filter;
    long pad = 0;
    if (n >= min_target) { const long rem = (n - min_target) % filter; if (rem) pad = filter - rem - 1; }
    else pad = min_target - n - 1;
    return pad > 0 ? pad : 0;
}

// t_floor: a tile must be at least as long as the halo (samples are shared by at most two tiles)
// waves: CTA waves per launch the automatic tiling aims at.  Single launches (forward / phase kernels) want
// many short waves (small tail); Griffin-Lim iterations run as two interleaved groups that cover each other's
// tails, so they take longer tiles (less per-tile prologue)
Tiling make_tiling(gomel_ctx* ctx, int n_clips, long n_frames, long sig_stride, long sig_len, int t_floor = 4, int waves = 8,
                   int ctas_per_sm = 2)
{
    Tiling tl;
    tl.n_frames = (int)n_frames;
    int T;
    if (ctx->tile_override > 0) T = ctx->tile_override;
    else {
        const long slots = (long)ctas_per_sm * ctx->sm_count;
        long tiles_wanted = (waves * slots + n_clips - 1) / n_clips;
        if (tiles_wanted < 1) tiles_wanted = 1;
        T = (int)((n_frames + tiles_wanted - 1) / tiles_wanted);
        // small batches are latency bound: allow tiles down to 4 frames until every CTA slot has a tile;
        // large batches keep tiles >= 16 frames so the per-tile prologue stays amortised
        const long tiles16 = (long)n_clips * ((n_frames + 15) / 16);
        const int t_min = tiles16 >= slots ? 16 : (tiles16 * 2 >= slots ? 8 : 4);
        if (T < t_min) T = t_min;
    }
    if (T & 1) T++;
    if (T < t_floor) T = t_floor;
    const long fr_even = n_frames + (n_frames & 1);
    if (T > fr_even) T = (int)fr_even;
    if (T < t_floor) T = t_floor;
    tl.tile_frames = T;
    tl.n_tiles = (int)((n_frames + T - 1) / T);
    tl.edge_first = tl.edge_last = 0;
    tl.sig_stride = sig_stride;
    tl.sig_len = sig_len;
    return tl;
}

int grid_1d(long n, int block) { long g = (n + block - 1) / block; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1; return (int)g; }

void build_fft_tables(std::vector<float>& blob, int n_win = kN)
{
    blob.assign(kTableBytes / 4, 0.0f);
    float* T1 = blob.data();
    float* T2 = T1 + kT1Cells * 2;
    float* win = T2 + kT2Cells * 2;
    const double two_pi = 6.283185307179586476925286766559;
    auto put = [](float* T, int cell, double a) { T[cell * 2 + 0] = (float)std::cos(a); T[cell * 2 + 1] = (float)(-std::sin(a)); };
    // cell index of power-slot i for a lane: float4 row i>>1, half i&1
    auto fill = [&](float* T, int mode, int lanes, int period) {
        for (int i = 0; i < mode; i++) {
            // mode 4: powers 1,2,4,8 ; mode 8: powers 1..8 ; mode 16: powers 0..15
            const int pw = mode == 4 ? (1 << i) : (mode == 8 ? i + 1 : i);
            for (int l = 0; l < lanes; l++)
                put(T, ((i >> 1) * lanes + l) * 2 + (i & 1), two_pi * (double)((l * pw) % period) / (double)period);
        }
    };
    fill(T1, kT1Mode, 256, 4096);
    fill(T2, kT2Mode, 16, 256);
    // symmetric Hann of gossp/go-dsp: 0.5*(1-cos(2 pi n/(N-1)))  (phase.py:122 np.hanning)
    for (int m = 0; m < n_win / 512; m++)            // first half only: w[n] = w[N-1-n]
        for (int t = 0; t < 256; t++) {
            const int n = t + 256 * m;
            win[m * 256 + t] = (float)(0.5 * (1.0 - std::cos(two_pi * (double)n / (double)(n_win - 1))));
        }
}

template <typename K>
int set_smem_attr(gomel_ctx* ctx, K kernel)
{
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    return 0;
}

// ---- window-sum gain tables for phase.ISTFT (phase/phase.go:114-130), float64 on the host
int prepare_gain(gomel_ctx* ctx, long n_frames, double boost)
{
    const long key = n_frames <= 16 ? n_frames : 17;
    if (ctx->gain_frames_key == key && ctx->gain_boost_key == boost && ctx->d_gain_head) return 0;
    const long F = n_frames <= 16 ? n_frames : 16;
    const long ola = kN + (F - 1) * kHop;
    std::vector<double> w(kN), ws(ola, 0.0);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < kN; n++) w[n] = 0.5 * (1.0 - std::cos(two_pi * (double)n / (double)(kN - 1)));
    for (long f = 0; f < F; f++)
        for (int j = 0; j < kN; j++) ws[f * kHop + j] += w[j] * w[j];
    double mx = 0.0;
    for (long i = 0; i < ola; i++) if (ws[i] > mx) mx = ws[i];
    const double thr = mx * 0.5;
    auto gain = [&](double v) -> float {
        double g = 1.0;
        if (v > thr) g = 1.0 / v;
        else if (v > 1e-21) g = 1.0 / thr;
        if (boost != 0) g *= boost;
        return (float)(g / (double)kN);         // includes the 1/N of the inverse transform (fft.IFFT divides by N)
    };
    std::vector<float> head, mid(kHop, 1.0f), tail;
    if (n_frames <= 16) {
        head.resize(ola);
        for (long i = 0; i < ola; i++) head[i] = gain(ws[i]);
        ctx->gain_head_len = (int)ola; ctx->gain_tail_len = 0;
        tail.assign(1, 1.0f);
    } else {
        head.resize(kN); tail.resize(kN);
        for (int i = 0; i < kN; i++) { head[i] = gain(ws[i]); tail[i] = gain(ws[ola - kN + i]); }
        for (long s = kN; s < kN + kHop; s++) mid[s % kHop] = gain(ws[s]);
        ctx->gain_head_len = kN; ctx->gain_tail_len = kN;
    }
    if (ctx->d_gain_head) { CU(cudaStreamSynchronize(ctx->st)); cudaFree(ctx->d_gain_head); cudaFree(ctx->d_gain_mid); cudaFree(ctx->d_gain_tail); }
    ctx->d_gain_head = ctx->d_gain_mid = ctx->d_gain_tail = nullptr;
    CU(cudaMalloc(&ctx->d_gain_head, head.size() * 4));
    CU(cudaMalloc(&ctx->d_gain_mid, mid.size() * 4));
    CU(cudaMalloc(&ctx->d_gain_tail, tail.size() * 4));
    CU(cudaMemcpyAsync(ctx->d_gain_head, head.data(), head.size() * 4, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_gain_mid, mid.data(), mid.size() * 4, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->d_gain_tail, tail.data(), tail.size() * 4, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->gain_frames_key = key; ctx->gain_boost_key = boost;
    return 0;
}

// ---- unlocked device-level implementations ---------------------------------------------
int fwd_dev(gomel_ctx* ctx, const gomel_config* cfg, int mode, const float* d_sig, int n_clips, long sig_stride,
            long sig_len, long n_frames, float* d_out)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, mode == MODE_MEL ? &geo : nullptr)) return rc;
    if (!d_sig || !d_out || n_clips <= 0 || n_frames <= 0 || sig_stride < sig_len)
        return fail(ctx, GOMEL_E_ARG, "bad argument to forward transform");
    if (sig_len < geo.ola(n_frames))
        return fail(ctx, GOMEL_E_ARG, "signal shorter than the frames requested");
    FwdParams p = {};
    p.sig = d_sig; p.tables = geo.alt ? ctx->d_tables_alt : ctx->d_tables;
    p.tl = make_tiling(ctx, n_clips, n_frames, sig_stride, sig_len);
    if (mode == MODE_MEL) {
        const MelTables* mt = find_mel_tables(ctx, cfg);
        if (!mt) return GOMEL_E_STATE;
        p.fwd_lo = mt->fwd_lo; p.fwd_hi = mt->fwd_hi; p.fwd_mod = mt->fwd_mod; p.n_mels = cfg->n_mels;
        p.mel_out = d_out;
    } else if (mode == MODE_PHASE) {
        if (cfg->n_freqs <= 0 || cfg->n_freqs > kN / 2) return fail(ctx, GOMEL_E_ARG, "NumFreqs out of range");
        p.n_freqs = cfg->n_freqs; p.phase_out = reinterpret_cast<float2*>(d_out);
    } else {
        p.spec_out = reinterpret_cast<float2*>(d_out);
    }
    const long grid = (long)n_clips * p.tl.n_tiles;
    if (grid > 0x7fffffffL) return fail(ctx, GOMEL_E_ARG, "too many tiles");
    CU(cudaEventRecord(ctx->ev_k0, ctx->st));
    if (mode == MODE_MEL && geo.alt) k_stft_fwd<kAltHS, MODE_MEL, kAltFS><<<(unsigned)grid, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    else if (mode == MODE_MEL) k_stft_fwd<kHS, MODE_MEL><<<(unsigned)grid, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    else if (mode == MODE_PHASE) k_stft_fwd<kHS, MODE_PHASE><<<(unsigned)grid, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    else k_stft_fwd<kHS, MODE_SPEC><<<(unsigned)grid, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    CU(cudaEventRecord(ctx->ev_k1, ctx->st));
    ctx->hot_launches = 1;
    ctx->launches++;
    CU(cudaGetLastError());
    return 0;
}

// ---- float64 tables of gl_f64.cuh (root powers 1,2,4,8 for both twiddle stages, half Hann window), built once
int ensure_tables_d64(gomel_ctx* ctx)
{
    namespace D = gomel::d64;
    if (ctx->d_tables_d64) return 0;
    std::vector<double> blob(D::kTableBytes / 8, 0.0);
    double* T1 = blob.data();
    double* T2 = T1 + D::kT1Cells * 2;
    double* win = T2 + D::kT2Cells * 2;
    const double two_pi = 6.283185307179586476925286766559;
    const int pw[4] = { 1, 2, 4, 8 };
    for (int i = 0; i < 4; i++) {
        for (int t = 0; t < 256; t++) {
            const double a = two_pi * (double)((t * pw[i]) % 4096) / 4096.0;
            T1[(i * 256 + t) * 2] = std::cos(a); T1[(i * 256 + t) * 2 + 1] = -std::sin(a);
        }
    }
    for (int i = 0; i < 4; i++)                      // stage 2: powers 1,2,4,8 of W256^n0
        for (int n0 = 0; n0 < 16; n0++) {
            const double a = two_pi * (double)((n0 * pw[i]) % 256) / 256.0;
            T2[(i * 16 + n0) * 2] = std::cos(a); T2[(i * 16 + n0) * 2 + 1] = -std::sin(a);
        }
    for (int n = 0; n < kN / 2; n++) win[n] = 0.5 * (1.0 - std::cos(two_pi * (double)n / (double)(kN - 1)));
    CU(cudaMalloc(&ctx->d_tables_d64, D::kTableBytes));
    CU(cudaMemcpy(ctx->d_tables_d64, blob.data(), D::kTableBytes, cudaMemcpyHostToDevice));
    const int n_alt = 256 * kAltFS;
    for (int n = 0; n < kN / 2; n++) win[n] = n < n_alt / 2 ? 0.5 * (1.0 - std::cos(two_pi * (double)n / (double)(n_alt - 1))) : 0.0;
    CU(cudaMalloc(&ctx->d_tables_d64_alt, D::kTableBytes));
    CU(cudaMemcpy(ctx->d_tables_d64_alt, blob.data(), D::kTableBytes, cudaMemcpyHostToDevice));
    CU(cudaFuncSetAttribute(D::k_gl_iter_f64<kHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::kSmemBytes));
    CU(cudaFuncSetAttribute(D::k_gl_iter_f64<kAltHS, kAltFS>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::kSmemBytes));
    CU((cudaFuncSetAttribute(D::k_gl_iter_f64<kHS, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::kSmemBytes)));
    CU((cudaFuncSetAttribute(D::k_gl_iter_f64<kAltHS, kAltFS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::kSmemBytes)));
    return 0;
}

// How many of `iters` Griffin-Lim iterations run in float64 (gl_f64.cuh) before the float32 kernel takes over.
// GOMEL_FLAG_F64: all of them.
int lead_iters(const gomel_ctx* ctx, const gomel_config* cfg, const Geo& geo)
{
    (void)geo;
    const int iters = cfg->gl_iters < 0 ? 0 : cfg->gl_iters;
    if (cfg->flags & GOMEL_FLAG_F64) return iters;
    int lead = ctx->lead_f64;
    if (ctx->f32_tail >= 0 && iters - ctx->f32_tail > lead) lead = iters - ctx->f32_tail;
    return lead < iters ? lead : iters;
}

// Buffers of one Griffin-Lim run.  Exactly one of out32 / out64 is set (out64 only when every iteration is
// float64); at most one of init32 / init64 (neither: U[0,1) from the seed).
struct GlIO {
    const float* mags32 = nullptr; const double* mags64 = nullptr;
    const float* init32 = nullptr; const double* init64 = nullptr;
    float* out32 = nullptr; double* out64 = nullptr;
};

// tiling of the guard's float64 re-run: few clips, so short tiles (8 frames) keep one iteration at a few pair times
Tiling redo_tiling(const Tiling& tl, const Geo& geo)
{
    (void)geo;
    Tiling rt = tl;
    int T = 8;
    const int fr_even = tl.n_frames + (tl.n_frames & 1);
    if (T > fr_even) T = fr_even;
    if (T < 4) T = 4;
    rt.tile_frames = T;
    rt.n_tiles = (tl.n_frames + T - 1) / T;
    rt.edge_first = rt.edge_last = 0;
    return rt;
}

// grow-only scratch of gl_dev for a batch of this size; called up front by the chunked pipelines so that no
// (device-synchronising) reallocation happens between chunks
int gl_reserve(gomel_ctx* ctx, const gomel_config* cfg, int n_clips, long n_frames, long sig_stride, bool need_init)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    const int iters = cfg->gl_iters, lead = lead_iters(ctx, cfg, geo);
    const long ola = geo.ola(n_frames);
    void* b;
    const size_t sig_bytes = (size_t)n_clips * sig_stride * 4;
    const Tiling tl = make_tiling(ctx, n_clips, n_frames, sig_stride, ola, geo.alt ? 8 : 4, ctx->gl_tile_waves);
    const size_t hb_elems = (size_t)n_clips * (tl.n_tiles + 1) * geo.halo + 4;
    if (need_init) { if (int rc = ensure(ctx, S_INIT, sig_bytes, &b)) return rc; }
    if (iters - lead > 0) {
        if (iters - lead > 1 || lead > 0) { if (int rc = ensure(ctx, S_SIGTMP, sig_bytes, &b)) return rc; }
        if (int rc = ensure(ctx, S_HB0, hb_elems * 4, &b)) return rc;
        if (int rc = ensure(ctx, S_HB1, hb_elems * 4, &b)) return rc;
    }
    if (lead > 0) {
        if (int rc = ensure(ctx, S_LSIGA, sig_bytes * 2, &b)) return rc;
        if (int rc = ensure(ctx, S_LSIGB, sig_bytes * 2, &b)) return rc;
        if (int rc = ensure(ctx, S_LHB0, hb_elems * 8, &b)) return rc;
        if (int rc = ensure(ctx, S_LHB1, hb_elems * 8, &b)) return rc;
        if (int rc = ensure_tables_d64(ctx)) return rc;
        if (iters - lead > 0 && ctx->gl_guard > 0.0f) {
            // the guard's statistic | clip scale | selection list | count, and the re-run's own head-partial buffers
            // (its tiles are short: it is latency bound, and indexed by list slot)
            if (int rc = ensure(ctx, S_GUARD, (size_t)n_clips * 12 + 16, &b)) return rc;
            const Tiling rt = redo_tiling(tl, geo);
            const size_t rhb = (size_t)n_clips * (rt.n_tiles + 1) * geo.halo * 8 + 32;
            if (int rc = ensure(ctx, S_GHB0, rhb, &b)) return rc;
            if (int rc = ensure(ctx, S_GHB1, rhb, &b)) return rc;
        }
    }
    return 0;
}

// the float64 state at the hand-over, kept for the guard's re-run of selected clips
struct RedoState { double* fin = nullptr; double* other = nullptr; double* hb[2] = { nullptr, nullptr }; const double* mags64 = nullptr; };
int gl_dev_f32(gomel_ctx* ctx, const gomel_config* cfg, const Geo& geo, const GlIO& io, const Tiling& tl,
               const float* cur32, float* tmp, int n_clips, int iters, int lead, int ns, cudaStream_t* gs, const RedoState* redo);

int gl_dev(gomel_ctx* ctx, const gomel_config* cfg, const GlIO& io, int n_clips, long n_frames,
           unsigned long long seed, long sig_stride)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    const long ola = geo.ola(n_frames);
    if (sig_stride < ola) return fail(ctx, GOMEL_E_ARG, "sig_stride < ola_len");
    if (io.init32 && io.init32 == io.out32) return fail(ctx, GOMEL_E_ARG, "d_init and d_out must not alias");
    const int iters = cfg->gl_iters;
    if (iters < 0) return fail(ctx, GOMEL_E_ARG, "GriffinLimIterations < 0");
    const int lead = lead_iters(ctx, cfg, geo);
    if (io.out64 && lead != iters) return fail(ctx, GOMEL_E_ARG, "float64 output needs GOMEL_FLAG_F64");
    if ((lead > 0 && !io.mags64) || (iters - lead > 0 && !io.mags32)) return fail(ctx, GOMEL_E_ARG, "magnitudes missing");
    const long n_sig = (long)n_clips * sig_stride;
    const size_t sig_bytes = (size_t)n_sig * 4;
    const bool fill = !io.init32 && !io.init64;
    if (int rc = gl_reserve(ctx, cfg, n_clips, n_frames, sig_stride, fill && lead_iters(ctx, cfg, geo) == 0)) return rc;
    const float* init32 = io.init32;
    const double* init64 = io.init64;
    if (fill && lead > 0) {         // the float64 lead iterations read their start signal as doubles: draw it as such
        double* b = (double*)ctx->scratch[S_LSIGA];
        k_fill_uniform<double><<<grid_1d(n_sig, 256), 256, 0, ctx->st>>>(b, n_sig, seed);
        ctx->launches++;
        init64 = b;
    } else if (fill) {
        float* b = (float*)ctx->scratch[S_INIT];
        k_fill_uniform<float><<<grid_1d(n_sig, 256), 256, 0, ctx->st>>>(b, n_sig, seed);
        ctx->launches++;
        init32 = b;
    }
    ctx->hot_launches = 0; ctx->lead_launches = 0; ctx->guard_clips = 0;
    if (iters == 0) {       // mel/mel.go:85: zero iterations return the start signal
        if (io.out64) {
            if (init64) CU(cudaMemcpyAsync(io.out64, init64, (size_t)n_sig * 8, cudaMemcpyDeviceToDevice, ctx->st));
            else { k_f32_to_f64<<<grid_1d(n_sig, 256), 256, 0, ctx->st>>>(init32, io.out64, n_sig, 1.0); ctx->launches++; }
        } else if (init64) { d64::k_f64_to_f32<<<grid_1d(n_sig, 256), 256, 0, ctx->st>>>(init64, io.out32, n_sig); ctx->launches++; }
        else CU(cudaMemcpyAsync(io.out32, init32, sig_bytes, cudaMemcpyDeviceToDevice, ctx->st));
        return 0;
    }
    const Tiling tl = make_tiling(ctx, n_clips, n_frames, sig_stride, ola, geo.alt ? 8 : 4, ctx->gl_tile_waves);
    const int hb_tiles = tl.n_tiles + 1;
    const long grid = (long)n_clips * tl.n_tiles;
    if (grid > 0x7fffffffL) return fail(ctx, GOMEL_E_ARG, "too many tiles");
    // groups of clips on concurrent streams (group g: clips [g*n/ns, (g+1)*n/ns)); every group needs whole waves
    int ns = ctx->gl_streams;
    while (ns > 1 && grid / ns < (long)kGlCtasPerSm * ctx->sm_count) ns--;
    if (ns > n_clips) ns = n_clips;
    cudaStream_t gs[4] = { ctx->st, ctx->st_gl[0], ctx->st_gl[1], ctx->st_gl[2] };
    auto c_lo = [&](int g) { return (int)((long)n_clips * g / ns); };
    auto fork = [&]() -> int {
        if (ns > 1) {
            CU(cudaEventRecord(ctx->ev_fork, ctx->st));
            for (int g = 1; g < ns; g++) CU(cudaStreamWaitEvent(gs[g], ctx->ev_fork, 0));
        }
        return 0;
    };
    auto join = [&]() -> int {
        for (int g = 1; g < ns; g++) {
            CU(cudaEventRecord(ctx->ev_join[g - 1], gs[g]));
            CU(cudaStreamWaitEvent(ctx->st, ctx->ev_join[g - 1], 0));
        }
        return 0;
    };

    float* tmp = (float*)ctx->scratch[S_SIGTMP];
    const float* cur32 = init32;
    RedoState redo;
    // ---------------- float64 lead iterations
    if (lead > 0) {
        double* sg[2] = { (double*)ctx->scratch[S_LSIGA], (double*)ctx->scratch[S_LSIGB] };
        double* hb[2] = { (double*)ctx->scratch[S_LHB0], (double*)ctx->scratch[S_LHB1] };
        const double* cur = init64;
        int w = (cur == sg[0]) ? 1 : 0;             // next buffer to write (a drawn start signal sits in sg[0])
        if (!cur) {
            k_f32_to_f64<<<grid_1d(n_sig, 256), 256, 0, ctx->st>>>(init32, sg[0], n_sig, 1.0);
            ctx->launches++;
            cur = sg[0]; w = 1;
        }
        CU(cudaEventRecord(ctx->ev_l0, ctx->st));
        if (int rc = fork()) return rc;
        d64::GLParams p = {};
        p.tables = geo.alt ? ctx->d_tables_d64_alt : ctx->d_tables_d64; p.tl = tl; p.mags = io.mags64;
        p.hb_tiles = hb_tiles; p.tile_lo = 0; p.tiles_in_launch = tl.n_tiles;
        for (int i = 0; i < lead; i++) {
            p.sig_in = cur; p.sig_out = sg[w];
            p.hb_in = (i == 0) ? nullptr : hb[(i - 1) & 1];
            p.hb_out = hb[i & 1];
            for (int g = 0; g < ns; g++) {
                p.clip0 = c_lo(g);
                const unsigned gg = (unsigned)((long)(c_lo(g + 1) - c_lo(g)) * tl.n_tiles);
                if (geo.alt) d64::k_gl_iter_f64<kAltHS, kAltFS><<<gg, kThreads, d64::kSmemBytes, gs[g]>>>(p);
                else d64::k_gl_iter_f64<kHS><<<gg, kThreads, d64::kSmemBytes, gs[g]>>>(p);
                ctx->launches++;
            }
            cur = sg[w]; w ^= 1;
        }
        double* fin = const_cast<double*>(cur);     // one of sg[]: lead >= 1
        // hand-over: fold the head partials in, then narrow to float32 (or deliver float64) -- per group, on its stream
        float* conv = nullptr;
        if (iters > lead) conv = ((((iters - 1 - lead) & 1) == 0) ? tmp : io.out32);     // not the first float32 iteration's output
        else if (io.out32) conv = io.out32;
        for (int g = 0; g < ns; g++) {
            const long c0 = c_lo(g), nc = c_lo(g + 1) - c0;
            if (tl.n_tiles > 1) {
                d64::k_halo_fix_f64<<<(unsigned)(nc * (tl.n_tiles - 1)), 256, 0, gs[g]>>>(
                    fin + c0 * sig_stride, hb[(lead - 1) & 1] + c0 * hb_tiles * geo.halo, tl, geo.hop, geo.halo, 1, hb_tiles, tl.n_tiles);
                ctx->launches++;
            }
            if (conv) {
                d64::k_f64_to_f32_rows<<<grid_1d(nc * ola, 256), 256, 0, gs[g]>>>(fin + c0 * sig_stride, conv + c0 * sig_stride, nc, ola, sig_stride);
                ctx->launches++;
            } else {
                CU(cudaMemcpy2DAsync(io.out64 + c0 * sig_stride, (size_t)sig_stride * 8, fin + c0 * sig_stride, (size_t)sig_stride * 8,
                                     (size_t)ola * 8, (size_t)nc, cudaMemcpyDeviceToDevice, gs[g]));
            }
        }
        if (int rc = join()) return rc;
        CU(cudaEventRecord(ctx->ev_l1, ctx->st));
        ctx->lead_launches = lead;
        cur32 = conv;
        if (iters == lead) { CU(cudaGetLastError()); return 0; }
        redo.fin = fin; redo.other = (fin == sg[0]) ? sg[1] : sg[0]; redo.hb[0] = hb[0]; redo.hb[1] = hb[1]; redo.mags64 = io.mags64;
    }
    // ---------------- float32 iterations
    return gl_dev_f32(ctx, cfg, geo, io, tl, cur32, tmp, n_clips, iters, lead, ns, gs,
                      (redo.fin && ctx->gl_guard > 0.0f && io.out32) ? &redo : nullptr);
}

int gl_dev_f32(gomel_ctx* ctx, const gomel_config* cfg, const Geo& geo, const GlIO& io, const Tiling& tl,
               const float* cur32, float* tmp, int n_clips, int iters, int lead, int ns, cudaStream_t* gs, const RedoState* redo)
{
    (void)cfg;
    const int hb_tiles = tl.n_tiles + 1;
    const long grid = (long)n_clips * tl.n_tiles;
    auto c_lo = [&](int g) { return (int)((long)n_clips * g / ns); };
    auto fork = [&]() -> int {
        if (ns > 1) {
            CU(cudaEventRecord(ctx->ev_fork, ctx->st));
            for (int g = 1; g < ns; g++) CU(cudaStreamWaitEvent(gs[g], ctx->ev_fork, 0));
        }
        return 0;
    };
    auto join = [&]() -> int {
        for (int g = 1; g < ns; g++) {
            CU(cudaEventRecord(ctx->ev_join[g - 1], gs[g]));
            CU(cudaStreamWaitEvent(ctx->st, ctx->ev_join[g - 1], 0));
        }
        return 0;
    };
    SynParams p = {};
    p.tables = geo.alt ? ctx->d_tables_alt : ctx->d_tables;
    p.tl = tl;
    p.mags = io.mags32;
    p.hb_tiles = hb_tiles; p.tile_lo = 0; p.tiles_in_launch = tl.n_tiles;
    float* hb[2] = { (float*)ctx->scratch[S_HB0], (float*)ctx->scratch[S_HB1] };
    // singular-bin guard: [n_clips] statistic (float bits) | [n_clips] clip scale | [n_clips] selection list | count
    unsigned int* g_stat = nullptr; float* g_scale = nullptr; int* g_list = nullptr; int* g_count = nullptr;
    float g_thr = 0.0f;
    if (redo) {
        g_stat = (unsigned int*)ctx->scratch[S_GUARD];
        g_scale = (float*)(g_stat + n_clips);
        g_list = (int*)(g_scale + n_clips);
        g_count = g_list + n_clips;
        CU(cudaMemsetAsync(g_stat, 0, (size_t)n_clips * 4, ctx->st));
        k_clip_scale<<<(unsigned)n_clips, 128, 0, ctx->st>>>(io.mags32, g_scale, tl.n_frames);
        ctx->launches++;
        p.guard_stat = g_stat;
        // statistic = M/|X| * rms_frame(M) with M pre-scaled by 1/N and |X| not: N * statistic / clip scale is the
        // leverage in the units of profiles/r02_gl_guard.md
        // The threshold is stated for a clip of kGuardRefFrames frames (the 10 s clips it was calibrated on): one bin's
        // share of a clip's norm falls with the square root of the frame count.
        g_thr = ctx->gl_guard * std::sqrt((float)tl.n_frames / (float)kGuardRefFrames) / (float)geo.n_fft;
    }
    CU(cudaEventRecord(ctx->ev_k0, ctx->st));
    if (int rc = fork()) return rc;
    for (int i = lead; i < iters; i++) {
        float* dst = (((iters - 1 - i) & 1) == 0) ? io.out32 : tmp;
        p.sig_in = cur32; p.sig_out = dst;
        p.hb_in = (i == lead) ? nullptr : (const float*)hb[(i - 1) & 1];
        p.hb_out = hb[i & 1];
        for (int g = 0; g < ns; g++) {
            p.clip0 = c_lo(g);
            const unsigned gg = (unsigned)((long)(c_lo(g + 1) - c_lo(g)) * tl.n_tiles);
            if (redo) {
                if (geo.alt) k_gl_iter<kAltHS, kAltFS, true><<<gg, kThreads, kGlSmemBytes, gs[g]>>>(p);
                else k_gl_iter<kHS, 16, true><<<gg, kThreads, kGlSmemBytes, gs[g]>>>(p);
            } else {
                if (geo.alt) k_gl_iter<kAltHS, kAltFS><<<gg, kThreads, kGlSmemBytes, gs[g]>>>(p);
                else k_gl_iter<kHS><<<gg, kThreads, kGlSmemBytes, gs[g]>>>(p);
            }
            ctx->launches++;
        }
        cur32 = dst;
    }
    p.clip0 = 0;
    if (int rc = join()) return rc;
    CU(cudaEventRecord(ctx->ev_k1, ctx->st));
    ctx->hot_launches = iters - lead;   // one "launch" = one iteration over the whole batch, whatever the grouping
    if (tl.n_tiles > 1) {
        k_halo_fix<<<(unsigned)(grid - n_clips), 256, 0, ctx->st>>>(io.out32, (const float*)hb[(iters - 1) & 1], tl, geo.hop,
                                                                  geo.halo, 0, 1, hb_tiles, p);
        ctx->launches++;
    }
    if (redo) {
        // The clips the guard selected run the same iterations again in float64, from the float64 signal of the
        // hand-over, and overwrite their float32 result: they end where GOMEL_FLAG_F64 would.  The selection is made on
        // the device; the re-run kernels are small fixed grids walking (list slot, tile) items over short tiles, so a
        // run without selected clips costs a few microseconds per launch and one with a few clips a few pair times
        // per iteration.
        d64::k_guard_select<<<1, 1024, 0, ctx->st>>>(g_stat, g_scale, g_thr, n_clips, g_list, g_count);
        ctx->launches++;
        ctx->guard_clips = n_clips; ctx->guard_thr_units = g_thr; ctx->guard_to_leverage = (float)geo.n_fft;
        // A call on a few clips is latency bound (the single-clip drop-in call): reading the count back costs one
        // stream synchronisation, the 18 blind launches cost more.  Large batches stay asynchronous.
        if (n_clips <= kGuardSyncClips) {
            CU(cudaMemcpyAsync(ctx->h_guard_count, g_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
            if (*ctx->h_guard_count == 0) { CU(cudaGetLastError()); return 0; }
        }
        const Tiling rt = redo_tiling(tl, geo);
        const int r_hb_tiles = rt.n_tiles + 1;
        double* rhb[2] = { (double*)ctx->scratch[S_GHB0], (double*)ctx->scratch[S_GHB1] };
        d64::GLParams q = {};
        q.tables = geo.alt ? ctx->d_tables_d64_alt : ctx->d_tables_d64; q.tl = rt; q.mags = redo->mags64;
        q.hb_tiles = r_hb_tiles; q.tile_lo = 0; q.tiles_in_launch = rt.n_tiles;
        q.sel.clips = g_list; q.sel.count = g_count;
        const long items = (long)n_clips * rt.n_tiles, slots = (long)kGlCtasPerSm * ctx->sm_count;
        const unsigned rgrid = (unsigned)(items < slots ? items : slots);
        const double* cur = redo->fin;
        double* dst64[2] = { redo->other, redo->fin };
        for (int j = 0; j < iters - lead; j++) {
            q.sig_in = cur; q.sig_out = dst64[j & 1];
            q.hb_in = (j == 0) ? nullptr : rhb[(j - 1) & 1];
            q.hb_out = rhb[j & 1];
            if (geo.alt) d64::k_gl_iter_f64<kAltHS, kAltFS, true><<<rgrid, kThreads, d64::kSmemBytes, ctx->st>>>(q);
            else d64::k_gl_iter_f64<kHS, 16, true><<<rgrid, kThreads, d64::kSmemBytes, ctx->st>>>(q);
            ctx->launches++;
            cur = dst64[j & 1];
        }
        double* fin = const_cast<double*>(cur);
        if (rt.n_tiles > 1) {
            d64::k_halo_fix_f64<<<rgrid, 256, 0, ctx->st>>>(fin, rhb[(iters - lead - 1) & 1], rt, geo.hop, geo.halo, 1, r_hb_tiles,
                                                           rt.n_tiles, q.sel);
            ctx->launches++;
        }
        const unsigned gy = (unsigned)(n_clips < 64 ? n_clips : 64);
        d64::k_f64_to_f32_selected<<<dim3(16, gy), 256, 0, ctx->st>>>(fin, io.out32, geo.ola(tl.n_frames), tl.sig_stride, q.sel);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

template <typename T, typename OUT>
int mags_dev(gomel_ctx* ctx, const gomel_config* cfg, const T* d_mel, long n_rows, OUT* d_mags, cudaStream_t stream = nullptr,
             float* d_mags_f32 = nullptr)         // OUT = double: also the float32 rows, in the same pass
{
    if (!stream) stream = ctx->st;
    const MelTables* mt = find_mel_tables(ctx, cfg);
    if (!mt) return GOMEL_E_STATE;
    if (cfg->tune_mul == 0) return fail(ctx, GOMEL_E_ARG, "TuneMul == 0");
    long g = (n_rows + kMagsRowsPerPass - 1) / kMagsRowsPerPass;
    if (g > 148L * 8) g = 148L * 8;
    const size_t sm = (size_t)kMagsRowsPerPass * cfg->n_mels * 2 * sizeof(double);
    if (cfg->n_fft == 256 * kAltFS)
        k_mags_from_mel<T, kAltFS, OUT><<<(unsigned)g, 256, sm, stream>>>(
            d_mel, d_mags, mt->inv_lo, mt->inv_hi, mt->inv_mod, cfg->n_mels, cfg->tune_add, cfg->tune_mul, n_rows, d_mags_f32);
    else
        k_mags_from_mel<T, 16, OUT><<<(unsigned)g, 256, sm, stream>>>(
            d_mel, d_mags, mt->inv_lo, mt->inv_hi, mt->inv_mod, cfg->n_mels, cfg->tune_add, cfg->tune_mul, n_rows, d_mags_f32);
    ctx->launches++;
    CU(cudaGetLastError());
    return 0;
}

// both precisions of the target magnitudes a Griffin-Lim run of this config needs
template <typename T>
int mags_for_gl(gomel_ctx* ctx, const gomel_config* cfg, const T* d_mel, long n_rows, int slot32, int slot64, GlIO* io,
                cudaStream_t stream = nullptr)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    const int lead = lead_iters(ctx, cfg, geo);
    void *m32 = nullptr, *m64 = nullptr;
    const bool need32 = cfg->gl_iters - lead > 0, need64 = lead > 0;
    if (need32) { if (int rc = ensure(ctx, slot32, (size_t)n_rows * kMagStride * 4, &m32)) return rc; }
    if (need64) { if (int rc = ensure(ctx, slot64, (size_t)n_rows * kMagStride * 8, &m64)) return rc; }
    if (need64) {         // one pass writes both precisions when both are needed
        if (int rc = mags_dev<T, double>(ctx, cfg, d_mel, n_rows, (double*)m64, stream, (float*)m32)) return rc;
    } else if (need32) {
        if (int rc = mags_dev<T, float>(ctx, cfg, d_mel, n_rows, (float*)m32, stream)) return rc;
    }
    io->mags32 = (const float*)m32;
    io->mags64 = (const double*)m64;
    return 0;
}

template <typename T>
int from_mel_dev_impl(gomel_ctx* ctx, const gomel_config* cfg, const T* d_mel, int n_clips, long n_frames,
                      GlIO io, unsigned long long seed, long sig_stride)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    if (!d_mel || (!io.out32 && !io.out64) || n_clips <= 0 || n_frames <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument to from_mel");
    if (int rc = mags_for_gl<T>(ctx, cfg, d_mel, (long)n_clips * n_frames, S_MAGS, S_LMAGS, &io)) return rc;
    return gl_dev(ctx, cfg, io, n_clips, n_frames, seed, sig_stride);
}

int from_phase_dev_impl(gomel_ctx* ctx, const gomel_config* cfg, const float* d_spec, int n_clips, long n_frames,
                        long sig_stride, float* d_out)
{
    if (int rc = check_cfg(ctx, cfg)) return rc;
    if (!d_spec || !d_out || n_clips <= 0 || n_frames <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument to from_phase");
    if (cfg->n_freqs <= 0 || cfg->n_freqs > kN / 2) return fail(ctx, GOMEL_E_ARG, "NumFreqs out of range");
    const long ola = kN + (n_frames - 1) * (long)kHop;
    if (sig_stride < ola) return fail(ctx, GOMEL_E_ARG, "sig_stride < ola_len");
    if (int rc = prepare_gain(ctx, n_frames, cfg->volume_boost)) return rc;
    SynParams p = {};
    p.tables = ctx->d_tables;
    p.tl = make_tiling(ctx, n_clips, n_frames, sig_stride, ola);
    p.sig_out = d_out;
    p.spec = reinterpret_cast<const float2*>(d_spec); p.n_freqs = cfg->n_freqs;
    p.gain_head = ctx->d_gain_head; p.gain_mid = ctx->d_gain_mid; p.gain_tail = ctx->d_gain_tail;
    p.head_len = ctx->gain_head_len; p.tail_len = ctx->gain_tail_len;
    p.gain_off = 0; p.total_len = ola;
    void* hb;
    p.hb_tiles = p.tl.n_tiles + 1; p.tile_lo = 0; p.tiles_in_launch = p.tl.n_tiles;
    if (int rc = ensure(ctx, S_HB0, (size_t)n_clips * p.hb_tiles * kHalo * 4 + 16, &hb)) return rc;
    p.hb_out = (float*)hb;
    const long grid = (long)n_clips * p.tl.n_tiles;
    if (grid > 0x7fffffffL) return fail(ctx, GOMEL_E_ARG, "too many tiles");
    k_istft_phase<kHS><<<(unsigned)grid, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    ctx->launches++;
    if (p.tl.n_tiles > 1) {
        k_halo_fix<<<(unsigned)(grid - n_clips), 256, 0, ctx->st>>>(d_out, (const float*)hb, p.tl, kHop, kHalo, 1, 1,
                                                                  p.hb_tiles, p);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}

// ---- STRICT float64 Griffin-Lim (GOMEL_FLAG_F64): same algorithm, float64 end to end ----------
int from_mel_f64(gomel_ctx* ctx, const gomel_config* cfg, const double* d_mel, long n_frames, const double* h_init,
                 unsigned long long seed, double* h_out)
{
    using namespace gomel::f64;
    const MelTables* mt = find_mel_tables(ctx, cfg);
    if (!mt) return GOMEL_E_STATE;
    if (cfg->tune_mul == 0) return fail(ctx, GOMEL_E_ARG, "TuneMul == 0");
    const long ola = kN + (n_frames - 1) * (long)kHop;
    if (!ctx->d_tables64) {
        std::vector<double> blob(kTableBytes64 / 8, 0.0);
        double* T1 = blob.data();
        double* T2 = T1 + kT1Cells64 * 2;
        double* win = T2 + kT2Cells64 * 2;
        const double two_pi = 6.283185307179586476925286766559;
        const int pw[4] = { 1, 2, 4, 8 };
        for (int i = 0; i < 4; i++) {
            for (int t = 0; t < 256; t++) {
                const double a = two_pi * (double)((t * pw[i]) % 4096) / 4096.0;
                T1[(i * 256 + t) * 2] = std::cos(a); T1[(i * 256 + t) * 2 + 1] = -std::sin(a);
            }
            for (int n0 = 0; n0 < 16; n0++) {
                const double a = two_pi * (double)((n0 * pw[i]) % 256) / 256.0;
                T2[(i * 16 + n0) * 2] = std::cos(a); T2[(i * 16 + n0) * 2 + 1] = -std::sin(a);
            }
        }
        for (int n = 0; n < kN; n++) win[n] = 0.5 * (1.0 - std::cos(two_pi * (double)n / (double)(kN - 1)));
        CU(cudaMalloc(&ctx->d_tables64, kTableBytes64));
        CU(cudaMemcpy(ctx->d_tables64, blob.data(), kTableBytes64, cudaMemcpyHostToDevice));
        CU(cudaFuncSetAttribute(k_gl_pair_f64<kHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes64));
    }
    void *sigA, *sigB, *Y, *mags;
    const long n_pairs = (n_frames + 1) / 2;
    if (int rc = ensure(ctx, S_SIG64A, (size_t)ola * 8, &sigA)) return rc;
    if (int rc = ensure(ctx, S_SIG64B, (size_t)ola * 8, &sigB)) return rc;
    if (int rc = ensure(ctx, S_Y64, (size_t)n_pairs * 2 * kN * 8, &Y)) return rc;
    if (int rc = ensure(ctx, S_MAGS64, (size_t)n_frames * 2049 * 8, &mags)) return rc;
    long g = n_frames < 148L * 8 ? n_frames : 148L * 8;
    k_mags_from_mel_f64<<<(unsigned)g, 256, (size_t)cfg->n_mels * 2 * sizeof(double), ctx->st>>>(
        d_mel, (double*)mags, mt->inv_lo, mt->inv_hi, mt->inv_mod, cfg->n_mels, cfg->tune_add, cfg->tune_mul, n_frames);
    ctx->launches++;
    if (h_init) CU(cudaMemcpyAsync(sigA, h_init, (size_t)ola * 8, cudaMemcpyHostToDevice, ctx->st));
    else {
        void* tmp;
        if (int rc = ensure(ctx, S_F32A, (size_t)ola * 4, &tmp)) return rc;
        k_fill_uniform<float><<<grid_1d(ola, 256), 256, 0, ctx->st>>>((float*)tmp, ola, seed);
        k_f32_to_f64<<<grid_1d(ola, 256), 256, 0, ctx->st>>>((const float*)tmp, (double*)sigA, ola, 1.0);
        ctx->launches += 2;
    }
    double *cur = (double*)sigA, *nxt = (double*)sigB;
    GL64Params p = {};
    p.tables = ctx->d_tables64; p.mags = (const double*)mags; p.Y = (double*)Y; p.n_frames = (int)n_frames; p.ola = ola;
    for (int i = 0; i < cfg->gl_iters; i++) {
        p.sig = cur;
        k_gl_pair_f64<kHS><<<(unsigned)n_pairs, kThreads, kSmemBytes64, ctx->st>>>(p);
        k_ola_f64<<<grid_1d(ola, 256), 256, 0, ctx->st>>>((const double*)Y, nxt, (int)n_frames, kHop, ola);
        ctx->launches += 2;
        double* t = cur; cur = nxt; nxt = t;
    }
    CU(cudaMemcpyAsync(h_out, cur, (size_t)ola * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return 0;
}

struct Guard {
    gomel_ctx* c;
    explicit Guard(gomel_ctx* ctx) : c(ctx) { c->mu.lock(); cudaSetDevice(c->device); }
    ~Guard() { c->mu.unlock(); }
};

}  // namespace

// =================================================================== exported C ABI
extern "C" {

const char* gomel_version(void) { return "gomel_b200 0.1 (sm_100a, Resolut/Window 4096/1280; mel paths also 2048/256)"; }

int gomel_ctx_create(int device, gomel_ctx** out)
{
    if (!out) return GOMEL_E_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) return GOMEL_E_CUDA;
    gomel_ctx* ctx = new gomel_ctx();
    ctx->device = device;
    auto boot = [&]() -> int {
        CU(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 3; i++) {
            CU(cudaStreamCreateWithFlags(&ctx->st_gl[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&ctx->st_pre, cudaStreamNonBlocking));
        if (const char* e = getenv("GOMEL_TILE_WAVES")) {           // tuning knob, 1..64
            const int v = atoi(e);
            if (v >= 1 && v <= 64) ctx->gl_tile_waves = v;
        }
        if (const char* e = getenv("GOMEL_GL_STREAMS")) {           // tuning knob, 1..4
            const int v = atoi(e);
            if (v >= 1 && v <= 4) ctx->gl_streams = v;
        }
        CU(cudaEventCreate(&ctx->ev0));
        CU(cudaEventCreate(&ctx->ev1));
        CU(cudaEventCreate(&ctx->ev_k0));
        CU(cudaEventCreate(&ctx->ev_k1));
        CU(cudaEventCreate(&ctx->ev_l0));
        CU(cudaEventCreate(&ctx->ev_l1));
        CU(cudaMallocHost(&ctx->h_guard_count, sizeof(int)));
        if (const char* e = getenv("GOMEL_LEAD_F64")) {             // float64 lead iterations, >= 0
            const int v = atoi(e);
            if (v >= 0) ctx->lead_f64 = v;
        }
        if (const char* e = getenv("GOMEL_F32_TAIL")) ctx->f32_tail = atoi(e) < 0 ? -1 : atoi(e);   // trailing float32 iterations, < 0 unlimited
        if (const char* e = getenv("GOMEL_GL_GUARD")) { const double v = atof(e); if (v >= 0) ctx->gl_guard = (float)v; }
        std::vector<float> blob;
        build_fft_tables(blob);
        CU(cudaMalloc(&ctx->d_tables, kTableBytes));
        CU(cudaMemcpy(ctx->d_tables, blob.data(), kTableBytes, cudaMemcpyHostToDevice));
        build_fft_tables(blob, 256 * kAltFS);
        CU(cudaMalloc(&ctx->d_tables_alt, kTableBytes));
        CU(cudaMemcpy(ctx->d_tables_alt, blob.data(), kTableBytes, cudaMemcpyHostToDevice));
        if (int rc = set_smem_attr(ctx, k_stft_fwd<kAltHS, MODE_MEL, kAltFS>)) return rc;
        CU(cudaFuncSetAttribute(k_gl_iter<kAltHS, kAltFS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlSmemBytes));
        CU(cudaFuncSetAttribute(k_gl_iter<kAltHS, kAltFS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlSmemBytes));
        CU(cudaFuncSetAttribute(k_gl_iter<kHS, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlSmemBytes));
        if (int rc = set_smem_attr(ctx, k_stft_fwd<kHS, MODE_MEL>)) return rc;
        if (int rc = set_smem_attr(ctx, k_stft_fwd<kHS, MODE_PHASE>)) return rc;
        if (int rc = set_smem_attr(ctx, k_stft_fwd<kHS, MODE_SPEC>)) return rc;
        CU(cudaFuncSetAttribute(k_gl_iter<kHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlSmemBytes));
        if (int rc = set_smem_attr(ctx, k_istft_phase<kHS>)) return rc;
        return 0;
    };
    const int rc = boot();
    if (rc) { fprintf(stderr, "gomel_ctx_create: %s\n", ctx->err.c_str()); delete ctx; return rc; }
    *out = ctx;
    return 0;
}

void gomel_ctx_destroy(gomel_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    for (int i = 0; i < S_COUNT; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_tables_alt);
    cudaFree(ctx->d_tables64);
    cudaFree(ctx->d_tables_d64); cudaFree(ctx->d_tables_d64_alt);
    cudaEventDestroy(ctx->ev_l0); cudaEventDestroy(ctx->ev_l1);
    cudaFreeHost(ctx->h_guard_count);
    for (MelTables& t : ctx->mel_tabs) t.release();
    cudaFree(ctx->d_gain_head); cudaFree(ctx->d_gain_mid); cudaFree(ctx->d_gain_tail);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev_k0); cudaEventDestroy(ctx->ev_k1);
    cudaStreamDestroy(ctx->st); cudaStreamDestroy(ctx->st_h2d); cudaStreamDestroy(ctx->st_d2h);
    for (int i = 0; i < 3; i++) { cudaStreamDestroy(ctx->st_gl[i]); cudaEventDestroy(ctx->ev_join[i]); }
    cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->st_pre);
    delete ctx;
}

const char* gomel_last_error(gomel_ctx* ctx) { return ctx ? ctx->err.c_str() : "NULL context"; }
unsigned long long gomel_launch_count(gomel_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gomel_set_tile_frames(gomel_ctx* ctx, int tile_frames)
{
    if (!ctx || tile_frames < 0) return GOMEL_E_ARG;
    ctx->tile_override = tile_frames;
    return 0;
}

int gomel_set_lead_f64(gomel_ctx* ctx, int lead)
{
    if (!ctx || lead < 0) return GOMEL_E_ARG;
    Guard g(ctx);
    const int prev = ctx->lead_f64;
    ctx->lead_f64 = lead;
    return prev;
}

int gomel_set_gl_guard(gomel_ctx* ctx, float threshold, float* previous)
{
    if (!ctx || !(threshold >= 0.0f)) return GOMEL_E_ARG;
    Guard g(ctx);
    if (previous) *previous = ctx->gl_guard;
    ctx->gl_guard = threshold;
    return 0;
}

int gomel_last_gl_guard(gomel_ctx* ctx, int* n_clips, int* n_rerun, float* max_leverage, float* leverage, int cap)
{
    if (!ctx || !n_clips || !n_rerun || !max_leverage || (cap > 0 && !leverage)) return GOMEL_E_ARG;
    Guard g(ctx);
    *n_clips = ctx->guard_clips; *n_rerun = 0; *max_leverage = 0.0f;
    if (ctx->guard_clips <= 0) return 0;
    const int n = ctx->guard_clips;
    std::vector<float> h((size_t)n * 2);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaMemcpy(h.data(), ctx->scratch[S_GUARD], (size_t)n * 8, cudaMemcpyDeviceToHost));
    const float to_lev = ctx->guard_to_leverage;     // statistic / clip scale -> leverage (the transform length)
    for (int c = 0; c < n; c++) {
        const float stat = h[c], scale = h[(size_t)n + c];
        const float lev = scale > 0 ? stat / scale * to_lev : 0.0f;
        if (stat > ctx->guard_thr_units * scale) ++*n_rerun;
        if (lev > *max_leverage) *max_leverage = lev;
        if (c < cap) leverage[c] = lev;
    }
    return 0;
}

int gomel_set_f32_tail(gomel_ctx* ctx, int tail)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    const int prev = ctx->f32_tail;
    ctx->f32_tail = tail < 0 ? -1 : tail;
    return prev < 0 ? 0x7fffffff : prev;
}

int gomel_frames(const gomel_config* cfg, long n_samples, long* n_padded, long* n_frames, long* ola_len)
{
    if (!cfg || n_samples <= 0 || cfg->hop <= 0 || cfg->n_fft <= 0) return GOMEL_E_ARG;
    const long np = n_samples + pad_len(n_samples, cfg->hop);
    if (np < cfg->n_fft) return GOMEL_E_ARG;
    const long fr = (long)((double)(np - cfg->n_fft) / (double)cfg->hop) + 1;   // gossp NumFrames
    if (n_padded) *n_padded = np;
    if (n_frames) *n_frames = fr;
    if (ola_len) *ola_len = cfg->n_fft + (fr - 1) * (long)cfg->hop;
    return 0;
}

long gomel_ola_len(const gomel_config* cfg, long n_frames)
{
    if (!cfg || n_frames <= 0) return GOMEL_E_ARG;
    return cfg->n_fft + (n_frames - 1) * (long)cfg->hop;
}

int gomel_set_mel_tables(gomel_ctx* ctx, const gomel_config* cfg, const int* fwd_lo, const int* fwd_hi,
                         const double* fwd_mod, const int* inv_lo, const int* inv_hi, const double* inv_mod)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    const int mels = cfg->n_mels, B = geo.n_fft / 2;
    if (mels <= 0 || mels > 4096 || !fwd_lo || !fwd_hi || !fwd_mod || !inv_lo || !inv_hi || !inv_mod)
        return fail(ctx, GOMEL_E_ARG, "bad mel table arguments");
    // index ranges the reference would panic on (mel/impl.go:331-338, :366-377) are refused here
    for (int i = 0; i < mels; i++) {
        const bool lerp = fwd_lo[i] + 1 == fwd_hi[i];
        if (fwd_lo[i] < 0 || (lerp && fwd_hi[i] > B - 1) || (!lerp && fwd_hi[i] > B))
            return fail(ctx, GOMEL_E_ARG, "forward mel table indexes outside the spectrum (the Go reference panics)");
    }
    for (int i = 0; i < B; i++) {
        const bool copy = inv_lo[i] == inv_hi[i], lerp = inv_lo[i] + 1 == inv_hi[i] && inv_hi[i] < mels;
        if (inv_lo[i] < 0 || (copy && inv_lo[i] >= mels) || (!copy && !lerp && inv_hi[i] > mels))
            return fail(ctx, GOMEL_E_ARG, "inverse mel table indexes outside the mel axis (the Go reference panics)");
    }
    // kernels already enqueued may still read a set that is replaced or evicted below
    CU(cudaStreamSynchronize(ctx->st));
    for (size_t i = 0; i < ctx->mel_tabs.size(); i++) {          // same key: replace
        MelTables& t = ctx->mel_tabs[i];
        if (t.n_fft == cfg->n_fft && t.n_mels == mels && t.fmin == cfg->mel_fmin && t.fmax == cfg->mel_fmax) {
            t.release();
            ctx->mel_tabs.erase(ctx->mel_tabs.begin() + (long)i);
            break;
        }
    }
    if (ctx->mel_tabs.size() >= kMaxMelTables) {                 // evict the least recently used set
        ctx->mel_tabs.front().release();
        ctx->mel_tabs.erase(ctx->mel_tabs.begin());
    }
    std::vector<float> fm(mels);
    for (int i = 0; i < mels; i++) fm[i] = (float)fwd_mod[i];
    MelTables t;
    t.n_fft = cfg->n_fft; t.n_mels = mels; t.fmin = cfg->mel_fmin; t.fmax = cfg->mel_fmax;
    auto upload = [&]() -> int {
        CU(cudaMalloc(&t.fwd_lo, mels * 4)); CU(cudaMalloc(&t.fwd_hi, mels * 4)); CU(cudaMalloc(&t.fwd_mod, mels * 4));
        CU(cudaMalloc(&t.inv_lo, B * 4)); CU(cudaMalloc(&t.inv_hi, B * 4)); CU(cudaMalloc(&t.inv_mod, B * 8));
        CU(cudaMemcpy(t.fwd_lo, fwd_lo, mels * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.fwd_hi, fwd_hi, mels * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.fwd_mod, fm.data(), mels * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.inv_lo, inv_lo, B * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.inv_hi, inv_hi, B * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.inv_mod, inv_mod, B * 8, cudaMemcpyHostToDevice));
        return 0;
    };
    if (int rc = upload()) { t.release(); return rc; }
    ctx->mel_tabs.push_back(t);
    return 0;
}

// ------------------------------------------------------------------- host-buffer API
static int host_forward(gomel_ctx* ctx, const gomel_config* cfg, int mode, const double* wav, long n, double* out)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, mode == MODE_MEL ? &geo : nullptr)) return rc;
    if (!wav || !out || n <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    long np, fr, ola;
    if (gomel_frames(cfg, n, &np, &fr, &ola)) return fail(ctx, GOMEL_E_ARG, "bad length");
    const long per_frame = (mode == MODE_MEL) ? 2L * cfg->n_mels : 2L * cfg->n_freqs;
    if (per_frame <= 0) return fail(ctx, GOMEL_E_ARG, "NumMels / NumFreqs must be positive");
    void *d64, *dsig, *dout, *dout64;
    if (int rc = ensure(ctx, S_F64IN, (size_t)n * 8, &d64)) return rc;
    if (int rc = ensure(ctx, S_F32A, (size_t)np * 4, &dsig)) return rc;
    if (int rc = ensure(ctx, S_F32B, (size_t)fr * per_frame * 4, &dout)) return rc;
    if (int rc = ensure(ctx, S_F64OUT, (size_t)fr * per_frame * 8, &dout64)) return rc;
    CU(cudaMemcpyAsync(d64, wav, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st));
    k_f64_to_f32_pad<<<grid_1d(np, 256), 256, 0, ctx->st>>>((const double*)d64, n, (float*)dsig, np);
    ctx->launches++;
    if (int rc = fwd_dev(ctx, cfg, mode, (const float*)dsig, 1, np, np, fr, (float*)dout)) return rc;
    k_f32_to_f64<<<grid_1d(fr * per_frame, 256), 256, 0, ctx->st>>>((const float*)dout, (double*)dout64, fr * per_frame, 1.0);
    ctx->launches++;
    CU(cudaMemcpyAsync(out, dout64, (size_t)fr * per_frame * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

int gomel_to_mel(gomel_ctx* ctx, const gomel_config* cfg, const double* wav, long n, double* mel_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return host_forward(ctx, cfg, MODE_MEL, wav, n, mel_out);
}

int gomel_to_phase(gomel_ctx* ctx, const gomel_config* cfg, const double* wav, long n, double* out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return host_forward(ctx, cfg, MODE_PHASE, wav, n, out);
}

int gomel_from_mel(gomel_ctx* ctx, const gomel_config* cfg, const double* mel, long n_frames,
                   const double* init_signal, unsigned long long seed, double* wav_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    Geo geo;
    const bool ref64 = cfg && (cfg->flags & GOMEL_FLAG_F64_REF);
    if (int rc = check_cfg(ctx, cfg, ref64 ? nullptr : &geo)) return rc;
    if (!mel || !wav_out || n_frames <= 0 || cfg->n_mels <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    if (cfg->gl_iters < 0) return fail(ctx, GOMEL_E_ARG, "GriffinLimIterations < 0");
    const long ola = geo.ola(n_frames);
    const long n_mel = n_frames * 2L * cfg->n_mels;
    void *dmel, *dinit64 = nullptr, *dinit = nullptr, *dout, *dout64;
    if (int rc = ensure(ctx, S_F64IN, (size_t)n_mel * 8, &dmel)) return rc;
    if (int rc = ensure(ctx, S_F32B, (size_t)ola * 4, &dout)) return rc;
    if (int rc = ensure(ctx, S_F64OUT, (size_t)ola * 8, &dout64)) return rc;
    CU(cudaMemcpyAsync(dmel, mel, (size_t)n_mel * 8, cudaMemcpyHostToDevice, ctx->st));
    if (ref64) return from_mel_f64(ctx, cfg, (const double*)dmel, n_frames, init_signal, seed, wav_out);
    const int lead = lead_iters(ctx, cfg, geo);
    GlIO io;
    if (init_signal) {
        if (int rc = ensure(ctx, S_F64IN2, (size_t)ola * 8, &dinit64)) return rc;
        CU(cudaMemcpyAsync(dinit64, init_signal, (size_t)ola * 8, cudaMemcpyHostToDevice, ctx->st));
        if (lead > 0 || cfg->gl_iters == 0) io.init64 = (const double*)dinit64;     // the float64 iterations start from the caller's exact signal
        else {
            if (int rc = ensure(ctx, S_F32A, (size_t)ola * 4, &dinit)) return rc;
            k_f64_to_f32_pad<<<grid_1d(ola, 256), 256, 0, ctx->st>>>((const double*)dinit64, ola, (float*)dinit, ola);
            ctx->launches++;
            io.init32 = (const float*)dinit;
        }
    }
    const bool all64 = lead == cfg->gl_iters;
    if (all64) io.out64 = (double*)dout64; else io.out32 = (float*)dout;
    if (int rc = from_mel_dev_impl<double>(ctx, cfg, (const double*)dmel, 1, n_frames, io, seed, ola)) return rc;
    if (!all64) {
        k_f32_to_f64<<<grid_1d(ola, 256), 256, 0, ctx->st>>>((const float*)dout, (double*)dout64, ola, 1.0);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(wav_out, dout64, (size_t)ola * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

int gomel_from_phase(gomel_ctx* ctx, const gomel_config* cfg, const double* spec, long n_frames, double* wav_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    if (int rc = check_cfg(ctx, cfg)) return rc;
    if (!spec || !wav_out || n_frames <= 0 || cfg->n_freqs <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    const long ola = kN + (n_frames - 1) * (long)kHop;
    const long n_spec = n_frames * 2L * cfg->n_freqs;
    void *d64, *d32, *dout, *dout64;
    if (int rc = ensure(ctx, S_F64IN, (size_t)n_spec * 8, &d64)) return rc;
    if (int rc = ensure(ctx, S_F32A, (size_t)n_spec * 4, &d32)) return rc;
    if (int rc = ensure(ctx, S_F32B, (size_t)ola * 4, &dout)) return rc;
    if (int rc = ensure(ctx, S_F64OUT, (size_t)ola * 8, &dout64)) return rc;
    CU(cudaMemcpyAsync(d64, spec, (size_t)n_spec * 8, cudaMemcpyHostToDevice, ctx->st));
    k_f64_to_f32_pad<<<grid_1d(n_spec, 256), 256, 0, ctx->st>>>((const double*)d64, n_spec, (float*)d32, n_spec);
    ctx->launches++;
    if (int rc = from_phase_dev_impl(ctx, cfg, (const float*)d32, 1, n_frames, ola, (float*)dout)) return rc;
    k_f32_to_f64<<<grid_1d(ola, 256), 256, 0, ctx->st>>>((const float*)dout, (double*)dout64, ola, 1.0);
    ctx->launches++;
    CU(cudaMemcpyAsync(wav_out, dout64, (size_t)ola * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

int gomel_image(gomel_ctx* ctx, const double* buf, long n_entries, int mels, unsigned short* out, double* minmax_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    if (!buf || !out || n_entries <= 0 || mels <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    // dumpbuffer covers stride*mels entries, stride = len/mels (mel/impl.go:17)
    const long used = (n_entries / mels) * mels;
    if (used <= 0) return fail(ctx, GOMEL_E_ARG, "fewer entries than one column");
    void *d64, *dpart, *dout;
    const int blocks = grid_1d(used, 256);
    if (int rc = ensure(ctx, S_F64IN, (size_t)used * 16, &d64)) return rc;
    if (int rc = ensure(ctx, S_MISC, (size_t)(blocks + 1) * 4 * 8, &dpart)) return rc;
    if (int rc = ensure(ctx, S_F32B, (size_t)used * 2, &dout)) return rc;
    double* dmm = (double*)dpart + (size_t)blocks * 4;
    CU(cudaMemcpyAsync(d64, buf, (size_t)used * 16, cudaMemcpyHostToDevice, ctx->st));
    k_minmax_f64<<<blocks, 256, 0, ctx->st>>>((const double*)d64, used, -99999999., 9999999., (double*)dpart);
    k_minmax_final<<<1, 32, 0, ctx->st>>>((const double*)dpart, blocks, -99999999., 9999999., dmm);
    k_quantise_u16<<<blocks, 256, 0, ctx->st>>>((const double*)d64, used, dmm, (unsigned short*)dout);
    ctx->launches += 3;
    CU(cudaMemcpyAsync(out, dout, (size_t)used * 2, cudaMemcpyDeviceToHost, ctx->st));
    if (minmax_out) CU(cudaMemcpyAsync(minmax_out, dmm, 32, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

int gomel_quantise(gomel_ctx* ctx, const double* buf, long n_entries, int mels, int flags, int ihs_passes,
                   unsigned short* rgb_out, double* minmax_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    if (!buf || !rgb_out || n_entries <= 0 || mels <= 0 || ihs_passes < 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    const long used = (n_entries / mels) * mels;
    if (used <= 0) return fail(ctx, GOMEL_E_ARG, "fewer entries than one column");
    void *d64, *dpart, *dout;
    const int blocks = grid_1d(used, 256);
    if (int rc = ensure(ctx, S_F64IN, (size_t)used * 16, &d64)) return rc;
    if (int rc = ensure(ctx, S_MISC, (size_t)(blocks + 1) * 4 * 8, &dpart)) return rc;
    if (int rc = ensure(ctx, S_F32B, (size_t)used * 6, &dout)) return rc;
    double* dmm = (double*)dpart + (size_t)blocks * 4;
    const double big = 1.79769313486231570814527423731704357e+308;      // math.MaxFloat64
    CU(cudaMemcpyAsync(d64, buf, (size_t)used * 16, cudaMemcpyHostToDevice, ctx->st));
    if (ihs_passes > 0) { k_asinh_passes<<<blocks, 256, 0, ctx->st>>>((double*)d64, used * 2, ihs_passes, 0); ctx->launches++; }
    k_minmax_f64<<<blocks, 256, 0, ctx->st>>>((const double*)d64, used, -big, big, (double*)dpart);
    k_minmax_final<<<1, 32, 0, ctx->st>>>((const double*)dpart, blocks, -big, big, dmm);
    ctx->launches += 2;
    double mm[4];
    CU(cudaMemcpyAsync(mm, dmm, 32, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (flags & GOMEL_Q_SINGLE_MINMAX) {
        const double mx = mm[0] > mm[1] ? mm[0] : mm[1], mn = mm[2] < mm[3] ? mm[2] : mm[3];
        mm[0] = mm[1] = mx; mm[2] = mm[3] = mn;
        CU(cudaMemcpyAsync(dmm, mm, 32, cudaMemcpyHostToDevice, ctx->st));
    }
    k_quantise_rgb<<<blocks, 256, 0, ctx->st>>>((const double*)d64, used, dmm, (flags & GOMEL_Q_HDR) ? 65535 : 255,
                                               (flags & GOMEL_Q_BLUE_WRAP) ? 1 : 0, (unsigned short*)dout);
    ctx->launches++;
    CU(cudaMemcpyAsync(rgb_out, dout, (size_t)used * 6, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (minmax_out) memcpy(minmax_out, mm, 32);
    return 0;
}

int gomel_dequantise(gomel_ctx* ctx, const unsigned short* rg, long n_entries, int hdr, double max0, double max1,
                     double min0, double min1, int ihs_passes, double* out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    if (!rg || !out || n_entries <= 0 || ihs_passes < 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    void *din, *dout;
    if (int rc = ensure(ctx, S_F32A, (size_t)n_entries * 4, &din)) return rc;
    if (int rc = ensure(ctx, S_F64OUT, (size_t)n_entries * 16, &dout)) return rc;
    CU(cudaMemcpyAsync(din, rg, (size_t)n_entries * 4, cudaMemcpyHostToDevice, ctx->st));
    k_dequantise<<<grid_1d(n_entries, 256), 256, 0, ctx->st>>>((const unsigned short*)din, n_entries, hdr ? 65535.0 : 255.0,
                                                              max0, max1, min0, min1, ihs_passes, (double*)dout);
    ctx->launches++;
    CU(cudaMemcpyAsync(out, dout, (size_t)n_entries * 16, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

// ------------------------------------------------------------------- device-resident API
int gomel_dev_malloc(gomel_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaMalloc(out, bytes ? bytes : 16));
    return 0;
}
int gomel_dev_free(gomel_ctx* ctx, void* p)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaFree(p));
    return 0;
}
int gomel_host_malloc(gomel_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaMallocHost(out, bytes ? bytes : 16));
    return 0;
}
int gomel_host_free(gomel_ctx* ctx, void* p)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaFreeHost(p));
    return 0;
}
int gomel_copy_h2d(gomel_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->st));
    return 0;
}
int gomel_copy_d2h(gomel_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->st));
    return 0;
}
int gomel_sync(gomel_ctx* ctx)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaStreamSynchronize(ctx->st_h2d));
    CU(cudaStreamSynchronize(ctx->st_d2h));
    return 0;
}
int gomel_timer_start(gomel_ctx* ctx)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaEventRecord(ctx->ev0, ctx->st));
    return 0;
}
int gomel_timer_stop(gomel_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaEventRecord(ctx->ev1, ctx->st));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return 0;
}

int gomel_last_hot_kernel_ms(gomel_ctx* ctx, float* ms, int* launches)
{
    if (!ctx || !ms || !launches) return GOMEL_E_ARG;
    Guard g(ctx);
    *ms = 0.f; *launches = 0;
    if (ctx->hot_launches <= 0) return ctx->lead_launches > 0 ? 0 : fail(ctx, GOMEL_E_STATE, "no transform has run on this context yet");
    CU(cudaEventSynchronize(ctx->ev_k1));
    CU(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
    *launches = ctx->hot_launches;
    return 0;
}

int gomel_last_lead_kernel_ms(gomel_ctx* ctx, float* ms, int* launches)
{
    if (!ctx || !ms || !launches) return GOMEL_E_ARG;
    Guard g(ctx);
    *ms = 0.f; *launches = 0;
    if (ctx->lead_launches <= 0) return 0;
    CU(cudaEventSynchronize(ctx->ev_l1));
    CU(cudaEventElapsedTime(ms, ctx->ev_l0, ctx->ev_l1));
    *launches = ctx->lead_launches;
    return 0;
}

int gomel_to_mel_dev(gomel_ctx* ctx, const gomel_config* cfg, const float* d_sig, int n_clips, long sig_stride,
                     long sig_len, long n_frames, float* d_mel)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return fwd_dev(ctx, cfg, MODE_MEL, d_sig, n_clips, sig_stride, sig_len, n_frames, d_mel);
}
int gomel_to_phase_dev(gomel_ctx* ctx, const gomel_config* cfg, const float* d_sig, int n_clips, long sig_stride,
                       long sig_len, long n_frames, float* d_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return fwd_dev(ctx, cfg, MODE_PHASE, d_sig, n_clips, sig_stride, sig_len, n_frames, d_out);
}
int gomel_stft_dev(gomel_ctx* ctx, const gomel_config* cfg, const float* d_sig, int n_clips, long sig_stride,
                   long sig_len, long n_frames, float* d_spec)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return fwd_dev(ctx, cfg, MODE_SPEC, d_sig, n_clips, sig_stride, sig_len, n_frames, d_spec);
}
int gomel_from_mel_dev(gomel_ctx* ctx, const gomel_config* cfg, const float* d_mel, int n_clips, long n_frames,
                       const float* d_init, unsigned long long seed, long sig_stride, float* d_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    GlIO io;
    io.init32 = d_init; io.out32 = d_out;
    if (cfg && (cfg->flags & GOMEL_FLAG_F64_REF)) return fail(ctx, GOMEL_E_UNSUPPORTED, "GOMEL_FLAG_F64_REF is a host-buffer test path");
    return from_mel_dev_impl<float>(ctx, cfg, d_mel, n_clips, n_frames, io, seed, sig_stride);
}
int gomel_from_phase_dev(gomel_ctx* ctx, const gomel_config* cfg, const float* d_spec, int n_clips, long n_frames,
                         long sig_stride, float* d_out)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return from_phase_dev_impl(ctx, cfg, d_spec, n_clips, n_frames, sig_stride, d_out);
}

}  // extern "C"

// ------------------------------------------------------------------- pipelined host batches
// chunk c uses buffer set c % 3; H2D on st_h2d, magnitudes / start signals on st_pre, iterations on st (+ the
// group streams), D2H on st_d2h, ordered by events.  Three sets: the copy-out of chunk c may still run while chunk
// c+2 computes, so compute never waits for the host link unless the link is the bottleneck over a whole chunk.
namespace {
struct PipeEvents {
    cudaEvent_t up[kPipeSets] = {}, done[kPipeSets] = {}, down[kPipeSets] = {}, pre[kPipeSets] = {};
    cudaError_t create()
    {
        for (int b = 0; b < kPipeSets; b++)
            for (cudaEvent_t* e : { &up[b], &done[b], &down[b], &pre[b] }) {
                const cudaError_t rc = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
                if (rc != cudaSuccess) return rc;
            }
        return cudaSuccess;
    }
    ~PipeEvents()
    {
        for (int b = 0; b < kPipeSets; b++)
            for (cudaEvent_t e : { up[b], done[b], down[b], pre[b] }) if (e) cudaEventDestroy(e);
    }
};
// first CUDA error of a pipelined loop; later calls are skipped
struct CudaChain {
    cudaError_t err = cudaSuccess; const char* what = "";
    bool operator()(cudaError_t e, const char* w) { if (err == cudaSuccess && e != cudaSuccess) { err = e; what = w; } return err == cudaSuccess; }
    bool ok() const { return err == cudaSuccess; }
};
#define CH(call) chain((call), #call)
}  // namespace

static int from_mel_batch_host_impl(gomel_ctx* ctx, const gomel_config* cfg, const float* mel, int n_clips, long n_frames,
                                    const float* init, unsigned long long seed, void* out, int clips_per_chunk, bool pcm16)
{
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    if (!mel || !out || n_clips <= 0 || n_frames <= 0 || cfg->n_mels <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    if (cfg->flags & GOMEL_FLAG_F64_REF) return fail(ctx, GOMEL_E_UNSUPPORTED, "GOMEL_FLAG_F64_REF is a host-buffer test path");
    const long ola = geo.ola(n_frames);
    const long mel_per = n_frames * 2L * cfg->n_mels;
    int cpc = clips_per_chunk > 0 ? clips_per_chunk : 64;
    if (cpc > n_clips) cpc = n_clips;
    if (cfg->gl_iters < 0) return fail(ctx, GOMEL_E_ARG, "GriffinLimIterations < 0");
    const int lead = lead_iters(ctx, cfg, geo);
    const bool need32 = cfg->gl_iters - lead > 0, need64 = lead > 0;
    void *dmel[kPipeSets], *dinit[kPipeSets], *dout[kPipeSets], *dpcm[kPipeSets] = {}, *dm32[kPipeSets] = {}, *dm64[kPipeSets] = {};
    for (int b = 0; b < kPipeSets; b++) {
        // per-chunk magnitudes and start signals have buffers of their own, so the next chunk's are produced on
        // st_pre while this chunk iterates: HBM-bound work under FP32-bound work
        if (need32) { if (int rc = ensure(ctx, pipe_slot(b, P_MAGS32), (size_t)cpc * n_frames * kMagStride * 4, &dm32[b])) return rc; }
        if (need64) { if (int rc = ensure(ctx, pipe_slot(b, P_MAGS64), (size_t)cpc * n_frames * kMagStride * 8, &dm64[b])) return rc; }
        if (int rc = ensure(ctx, pipe_slot(b, P_INIT), (size_t)cpc * ola * 4, &dinit[b])) return rc;
        if (int rc = ensure(ctx, pipe_slot(b, P_IN), (size_t)cpc * mel_per * 4, &dmel[b])) return rc;
        if (int rc = ensure(ctx, pipe_slot(b, P_OUT), (size_t)cpc * ola * 4, &dout[b])) return rc;
        if (pcm16) { if (int rc = ensure(ctx, pipe_slot(b, P_PCM), (size_t)cpc * ola * 2, &dpcm[b])) return rc; }
    }
    // the iteration scratch is sized for a full chunk up front: no reallocation (= device synchronisation) mid-pipeline
    if (int rc = gl_reserve(ctx, cfg, cpc, n_frames, ola, false)) return rc;
    PipeEvents ev;
    CU(ev.create());
    // chunk schedule: a small first chunk (short pipeline fill), full chunks, then a taper down to 64 clips (GOMEL_CHUNK_TAPER_MIN)
    // (short drain: the last D2H is the only copy that cannot overlap compute)
    std::vector<int> sizes;
    {
        int left = n_clips;
        const char* ef = getenv("GOMEL_CHUNK_TAPER_MIN");           // smallest tapered chunk (tuning knob)
        int floor_ = ef ? atoi(ef) : 64;
        if (floor_ < 1) floor_ = 1;
        // first chunk: its H2D is the pipeline fill, so it is small -- but not so small that the second chunk's
        // H2D outlasts its compute
        int first = cpc >= 128 ? (cpc / 8 > 64 ? cpc / 8 : 64) : cpc;
        if (const char* e1 = getenv("GOMEL_CHUNK_FIRST")) {        // clips in the first chunk (tuning knob)
            const int v = atoi(e1);
            if (v >= 1 && v <= cpc) first = v;
        }
        auto push = [&](int n) { n = n < left ? n : left; if (n > 0) { sizes.push_back(n); left -= n; } };
        push(first);
        int taper = 0;
        for (int t = cpc / 2; t >= floor_; t /= 2) taper += t;
        while (left > taper + cpc) push(cpc);
        if (left > taper) push(left - taper);
        for (int t = cpc / 2; t >= floor_ && left > 0; t /= 2) push(t);
        push(left);
    }
    const int n_chunks = (int)sizes.size();
    int rc = 0, c0 = 0;
    CudaChain chain;
    for (int c = 0; c < n_chunks && !rc && chain.ok(); c0 += sizes[c], c++) {
        const int b = c % kPipeSets, nc = sizes[c];
        if (c >= kPipeSets) CH(cudaStreamWaitEvent(ctx->st_h2d, ev.done[b], 0));   // inputs of chunk c-3 consumed
        CH(cudaMemcpyAsync(dmel[b], mel + (size_t)c0 * mel_per, (size_t)nc * mel_per * 4, cudaMemcpyHostToDevice, ctx->st_h2d));
        if (init) CH(cudaMemcpyAsync(dinit[b], init + (size_t)c0 * ola, (size_t)nc * ola * 4, cudaMemcpyHostToDevice, ctx->st_h2d));
        CH(cudaEventRecord(ev.up[b], ctx->st_h2d));
        CH(cudaStreamWaitEvent(ctx->st_pre, ev.up[b], 0));
        if (c >= kPipeSets) CH(cudaStreamWaitEvent(ctx->st_pre, ev.done[b], 0));    // chunk c-3 no longer reads dmags[b] / dinit[b]
        if (!chain.ok()) break;
        GlIO io;
        if (need64) {         // one pass writes both precisions when both are needed
            rc = mags_dev<float, double>(ctx, cfg, (const float*)dmel[b], (long)nc * n_frames, (double*)dm64[b], ctx->st_pre,
                                         need32 ? (float*)dm32[b] : nullptr);
            io.mags64 = (const double*)dm64[b];
        } else if (need32) {
            rc = mags_dev<float, float>(ctx, cfg, (const float*)dmel[b], (long)nc * n_frames, (float*)dm32[b], ctx->st_pre);
        }
        if (need32) io.mags32 = (const float*)dm32[b];
        if (!rc && !init) {
            // indexed by the sample's position in the whole batch: the start signals do not depend on the chunking
            // and equal those of gomel_from_mel_dev(seed) on the same batch
            k_fill_uniform<float><<<grid_1d((long)nc * ola, 256), 256, 0, ctx->st_pre>>>((float*)dinit[b], (long)nc * ola, seed,
                                                                                 (long)c0 * ola);
            ctx->launches++;
        }
        if (rc) break;
        CH(cudaEventRecord(ev.pre[b], ctx->st_pre));
        CH(cudaStreamWaitEvent(ctx->st, ev.pre[b], 0));
        if (c >= kPipeSets) CH(cudaStreamWaitEvent(ctx->st, ev.down[b], 0));        // output buffer of chunk c-3 drained
        if (!chain.ok()) break;
        io.init32 = (const float*)dinit[b]; io.out32 = (float*)dout[b];
        rc = gl_dev(ctx, cfg, io, nc, n_frames, seed + (unsigned long long)c0, ola);
        if (rc) break;
        if (pcm16) {
            k_f32_to_pcm16<<<grid_1d((long)nc * ola, 256), 256, 0, ctx->st>>>((const float*)dout[b], (short*)dpcm[b], (long)nc * ola);
            ctx->launches++;
        }
        CH(cudaEventRecord(ev.done[b], ctx->st));
        CH(cudaStreamWaitEvent(ctx->st_d2h, ev.done[b], 0));
        if (pcm16)
            CH(cudaMemcpyAsync((short*)out + (size_t)c0 * ola, dpcm[b], (size_t)nc * ola * 2, cudaMemcpyDeviceToHost, ctx->st_d2h));
        else
            CH(cudaMemcpyAsync((float*)out + (size_t)c0 * ola, dout[b], (size_t)nc * ola * 4, cudaMemcpyDeviceToHost, ctx->st_d2h));
        CH(cudaEventRecord(ev.down[b], ctx->st_d2h));
    }
    // one exit path: drain every stream whatever happened, then report the first failure
    CH(cudaStreamSynchronize(ctx->st_h2d)); CH(cudaStreamSynchronize(ctx->st_pre)); CH(cudaStreamSynchronize(ctx->st));
    CH(cudaStreamSynchronize(ctx->st_d2h));
    if (rc) return rc;
    if (!chain.ok()) return fail(ctx, GOMEL_E_CUDA, std::string(chain.what) + ": " + cudaGetErrorString(chain.err));
    CU(cudaGetLastError());
    return 0;
}

extern "C" {

int gomel_from_mel_batch_host(gomel_ctx* ctx, const gomel_config* cfg, const float* mel, int n_clips, long n_frames,
                              const float* init, unsigned long long seed, float* out, int clips_per_chunk)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return from_mel_batch_host_impl(ctx, cfg, mel, n_clips, n_frames, init, seed, out, clips_per_chunk, false);
}

int gomel_from_mel_batch_host_pcm16(gomel_ctx* ctx, const gomel_config* cfg, const float* mel, int n_clips, long n_frames,
                                    const float* init, unsigned long long seed, short* out, int clips_per_chunk)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    return from_mel_batch_host_impl(ctx, cfg, mel, n_clips, n_frames, init, seed, out, clips_per_chunk, true);
}

int gomel_to_mel_batch_host(gomel_ctx* ctx, const gomel_config* cfg, const float* wav, int n_clips, long n_samples,
                            float* mel_out, int clips_per_chunk)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    Geo geo;
    if (int rc = check_cfg(ctx, cfg, &geo)) return rc;
    if (!wav || !mel_out || n_clips <= 0 || n_samples <= 0 || cfg->n_mels <= 0) return fail(ctx, GOMEL_E_ARG, "bad argument");
    long np, fr, ola;
    if (gomel_frames(cfg, n_samples, &np, &fr, &ola)) return fail(ctx, GOMEL_E_ARG, "bad length");
    const long stride = (np + 3) & ~3L;
    const long mel_per = fr * 2L * cfg->n_mels;
    int cpc = clips_per_chunk > 0 ? clips_per_chunk : 64;
    if (cpc > n_clips) cpc = n_clips;
    void *dsig[kPipeSets], *dout[kPipeSets];
    for (int b = 0; b < kPipeSets; b++) {
        if (int rc = ensure(ctx, pipe_slot(b, P_IN), (size_t)cpc * stride * 4, &dsig[b])) return rc;
        if (int rc = ensure(ctx, pipe_slot(b, P_OUT), (size_t)cpc * mel_per * 4, &dout[b])) return rc;
        CU(cudaMemsetAsync(dsig[b], 0, (size_t)cpc * stride * 4, ctx->st_h2d));   // pad() zeros, written once
    }
    PipeEvents ev;
    CU(ev.create());
    int rc = 0;
    CudaChain chain;
    const int n_chunks = (n_clips + cpc - 1) / cpc;
    for (int c = 0; c < n_chunks && !rc && chain.ok(); c++) {
        const int b = c % kPipeSets, c0 = c * cpc, nc = (n_clips - c0 < cpc) ? n_clips - c0 : cpc;
        if (c >= kPipeSets) CH(cudaStreamWaitEvent(ctx->st_h2d, ev.done[b], 0));
        CH(cudaMemcpy2DAsync(dsig[b], (size_t)stride * 4, wav + (size_t)c0 * n_samples, (size_t)n_samples * 4,
                             (size_t)n_samples * 4, nc, cudaMemcpyHostToDevice, ctx->st_h2d));
        CH(cudaEventRecord(ev.up[b], ctx->st_h2d));
        CH(cudaStreamWaitEvent(ctx->st, ev.up[b], 0));
        if (c >= kPipeSets) CH(cudaStreamWaitEvent(ctx->st, ev.down[b], 0));
        if (!chain.ok()) break;
        rc = fwd_dev(ctx, cfg, MODE_MEL, (const float*)dsig[b], nc, stride, np, fr, (float*)dout[b]);
        if (rc) break;
        CH(cudaEventRecord(ev.done[b], ctx->st));
        CH(cudaStreamWaitEvent(ctx->st_d2h, ev.done[b], 0));
        CH(cudaMemcpyAsync(mel_out + (size_t)c0 * mel_per, dout[b], (size_t)nc * mel_per * 4, cudaMemcpyDeviceToHost, ctx->st_d2h));
        CH(cudaEventRecord(ev.down[b], ctx->st_d2h));
    }
    CH(cudaStreamSynchronize(ctx->st_h2d)); CH(cudaStreamSynchronize(ctx->st)); CH(cudaStreamSynchronize(ctx->st_d2h));
    if (rc) return rc;
    if (!chain.ok()) return fail(ctx, GOMEL_E_CUDA, std::string(chain.what) + ": " + cudaGetErrorString(chain.err));
    CU(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------- time-split Griffin-Lim (config 5)
static int (*g_nccl_destroy)(void*) = nullptr;      // set once NCCL has been loaded
struct gomel_ts {
    gomel_ctx* ctx = nullptr;
    gomel_config cfg;
    int rank = 0, world = 1;
    long f_begin = 0, n_local = 0, sample_begin = 0, n_samples = 0, n_frames_total = 0;
    Tiling tl;
    int ext_prev = 0, ext_next = 0;
    float *sig[2] = { nullptr, nullptr }, *hb[2] = { nullptr, nullptr }, *mags = nullptr;
    // float64 lead iterations (same policy as the batch path: iterations [0, lead) run on k_gl_iter_f64)
    int lead = 0;
    double *sig64[2] = { nullptr, nullptr }, *hb64[2] = { nullptr, nullptr }, *mags64 = nullptr;
    cudaStream_t st_edge = nullptr, st_comm = nullptr;
    cudaEvent_t ev_edge = nullptr, ev_int = nullptr, ev_comm = nullptr;
    bool have_edge = false, have_int = false, have_comm = false;
    void* nccl_comm = nullptr;
};

extern "C" {

int gomel_ts_create2(gomel_ctx* ctx, const gomel_config* cfg, long n_frames_total, int rank, int world,
                     int tile_frames, int edge_frames, gomel_ts** out)
{
    if (!ctx || !out) return GOMEL_E_ARG;
    Guard g(ctx);
    *out = nullptr;
    if (int rc = check_cfg(ctx, cfg)) return rc;
    if (world < 1 || rank < 0 || rank >= world || n_frames_total <= 0 || edge_frames < 0)
        return fail(ctx, GOMEL_E_ARG, "bad rank/world/frames");
    int T = tile_frames > 0 ? tile_frames : 16;
    if (T & 1) T++;
    if (T < 4) T = 4;
    const long tiles_total = (n_frames_total + T - 1) / T;
    if (tiles_total < world) return fail(ctx, GOMEL_E_ARG, "fewer tiles than ranks: lower tile_frames");
    const long a = rank * tiles_total / world, b = (rank + 1) * tiles_total / world;
    gomel_ts* ts = new gomel_ts();
    ts->ctx = ctx; ts->cfg = *cfg; ts->rank = rank; ts->world = world; ts->n_frames_total = n_frames_total;
    ts->f_begin = a * T;
    const long f_end = (b * T < n_frames_total) ? b * T : n_frames_total;
    ts->n_local = f_end - ts->f_begin;
    ts->sample_begin = ts->f_begin * kHop;
    ts->n_samples = ts->n_local * kHop + kHalo;
    ts->ext_prev = rank > 0; ts->ext_next = rank + 1 < world;
    ts->tl.n_frames = (int)ts->n_local; ts->tl.tile_frames = T;
    ts->tl.sig_stride = ts->n_samples; ts->tl.sig_len = ts->n_samples;
    // short tiles next to a rank boundary (the ones the halo exchange waits for); every tile but the clip's last
    // must hold whole pairs and be at least 4 frames long (>= Resolut - Window samples)
    int e = edge_frames & ~1;
    if (e > 0 && e < 4) e = 4;
    if (e >= T) e = 0;
    int ef = ts->ext_prev ? e : 0, el = ts->ext_next ? e : 0;
    auto rem_ok = [&](int ef_, int el_) {
        const long mid = ts->n_local - ef_ - el_;
        if (mid < 4) return false;
        const long r = mid % T;
        return r == 0 || r >= 4 || !ts->ext_next;      // the remainder tile precedes the short last tile
    };
    while ((ef || el) && !rem_ok(ef, el)) { if (el) el = el > 4 ? el - 2 : 0; else ef = ef > 4 ? ef - 2 : 0; }
    ts->tl.edge_first = ef; ts->tl.edge_last = el;
    const long mid = ts->n_local - ef - el;
    ts->tl.n_tiles = (ef > 0) + (int)((mid + T - 1) / T) + (el > 0);
    ts->lead = (cfg->flags & GOMEL_FLAG_F64) ? 0x7fffffff : lead_iters(ctx, cfg, Geo());
    auto boot = [&]() -> int {
        const size_t hb_bytes = (size_t)(ts->tl.n_tiles + 1) * kHalo * 4;
        for (int i = 0; i < 2; i++) {
            CU(cudaMalloc(&ts->sig[i], (size_t)ts->n_samples * 4));
            CU(cudaMalloc(&ts->hb[i], hb_bytes));
            CU(cudaMemsetAsync(ts->hb[i], 0, hb_bytes, ctx->st));
            if (ts->lead > 0) {
                CU(cudaMalloc(&ts->sig64[i], (size_t)ts->n_samples * 8));
                CU(cudaMalloc(&ts->hb64[i], hb_bytes * 2));
                CU(cudaMemsetAsync(ts->hb64[i], 0, hb_bytes * 2, ctx->st));
            }
        }
        CU(cudaMalloc(&ts->mags, (size_t)ts->n_local * kMagStride * 4));
        if (ts->lead > 0) {
            CU(cudaMalloc(&ts->mags64, (size_t)ts->n_local * kMagStride * 8));
            if (int rc = ensure_tables_d64(ctx)) return rc;
        }
        CU(cudaStreamCreateWithFlags(&ts->st_edge, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ts->st_comm, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ts->ev_edge, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ts->ev_int, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ts->ev_comm, cudaEventDisableTiming));
        return 0;
    };
    if (int rc = boot()) { gomel_ts_destroy(ts); return rc; }
    *out = ts;
    return 0;
}

void gomel_ts_destroy(gomel_ts* ts)
{
    if (!ts) return;
    cudaSetDevice(ts->ctx->device);
    cudaStreamSynchronize(ts->ctx->st);
    if (ts->st_edge) cudaStreamSynchronize(ts->st_edge);
    if (ts->st_comm) cudaStreamSynchronize(ts->st_comm);
    if (ts->nccl_comm && g_nccl_destroy) g_nccl_destroy(ts->nccl_comm);
    for (int i = 0; i < 2; i++) { cudaFree(ts->sig[i]); cudaFree(ts->hb[i]); cudaFree(ts->sig64[i]); cudaFree(ts->hb64[i]); }
    cudaFree(ts->mags); cudaFree(ts->mags64);
    if (ts->ev_edge) cudaEventDestroy(ts->ev_edge);
    if (ts->ev_int) cudaEventDestroy(ts->ev_int);
    if (ts->ev_comm) cudaEventDestroy(ts->ev_comm);
    if (ts->st_edge) cudaStreamDestroy(ts->st_edge);
    if (ts->st_comm) cudaStreamDestroy(ts->st_comm);
    delete ts;
}

int gomel_ts_range(gomel_ts* ts, long* frame_begin, long* n_frames_local, long* sample_begin, long* n_samples_local)
{
    if (!ts) return GOMEL_E_ARG;
    if (frame_begin) *frame_begin = ts->f_begin;
    if (n_frames_local) *n_frames_local = ts->n_local;
    if (sample_begin) *sample_begin = ts->sample_begin;
    if (n_samples_local) *n_samples_local = ts->n_samples;
    return 0;
}

int gomel_ts_load(gomel_ts* ts, const float* d_mel_local, const float* d_init_local, unsigned long long seed)
{
    if (!ts || !d_mel_local) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    if (int rc = mags_dev<float, float>(ctx, &ts->cfg, d_mel_local, ts->n_local, ts->mags)) return rc;
    if (ts->lead > 0) { if (int rc = mags_dev<float, double>(ctx, &ts->cfg, d_mel_local, ts->n_local, ts->mags64)) return rc; }
    if (d_init_local) CU(cudaMemcpyAsync(ts->sig[0], d_init_local, (size_t)ts->n_samples * 4, cudaMemcpyDeviceToDevice, ctx->st));
    else {
        k_fill_uniform<float><<<grid_1d(ts->n_samples, 256), 256, 0, ctx->st>>>(ts->sig[0], ts->n_samples, seed, ts->sample_begin);
        ctx->launches++;
    }
    if (ts->lead > 0) {       // the float64 iterations start from the exact float32 start signal
        k_f32_to_f64<<<grid_1d(ts->n_samples, 256), 256, 0, ctx->st>>>(ts->sig[0], ts->sig64[0], ts->n_samples, 1.0);
        ctx->launches++;
    }
    ts->have_edge = ts->have_int = ts->have_comm = false;
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

// hand-over from the float64 lead iterations to the float32 ones (and the float64 form of gomel_ts_finish): fold
// every halo partial of iteration `last_iter` in -- own tiles' heads, the previous rank's tail (already in sig[0..halo))
// and the next rank's head partial over this rank's tail region -- then narrow to float32.  Runs on ctx->st.
static int ts_fold_f64(gomel_ts* ts, int last_iter, float* dst32)
{
    gomel_ctx* ctx = ts->ctx;
    if (ts->have_comm) CU(cudaStreamWaitEvent(ctx->st, ts->ev_comm, 0));
    if (ts->have_edge) CU(cudaStreamWaitEvent(ctx->st, ts->ev_edge, 0));
    double* fin = ts->sig64[(last_iter + 1) & 1];
    const int t_first = ts->ext_prev ? 0 : 1, t_end = ts->tl.n_tiles + (ts->ext_next ? 1 : 0);
    if (t_end > t_first) {
        d64::k_halo_fix_f64<<<(unsigned)(t_end - t_first), 256, 0, ctx->st>>>(fin, ts->hb64[last_iter & 1], ts->tl, kHop, kHalo,
                                                                            t_first, ts->tl.n_tiles + 1, t_end);
        ctx->launches++;
    }
    d64::k_f64_to_f32<<<grid_1d(ts->n_samples, 256), 256, 0, ctx->st>>>(fin, dst32, ts->n_samples);
    ctx->launches++;
    return 0;
}

int gomel_ts_iterate(gomel_ts* ts, int iter, int part)
{
    if (!ts || iter < 0 || part < 0 || part > 2) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    const bool f64 = iter < ts->lead;
    const int nt = ts->tl.n_tiles;
    if (!f64 && iter == ts->lead && iter > 0 && part != 2) {
        // first float32 iteration: its input is the folded, narrowed result of the last float64 iteration
        if (int rc = ts_fold_f64(ts, iter - 1, ts->sig[iter & 1])) return rc;
        CU(cudaEventRecord(ts->ev_int, ctx->st)); ts->have_int = true;
    }
    SynParams p = {};
    d64::GLParams q = {};
    if (f64) {
        q.tables = ctx->d_tables_d64; q.tl = ts->tl; q.mags = ts->mags64;
        q.sig_in = ts->sig64[iter & 1]; q.sig_out = ts->sig64[(iter + 1) & 1];
        q.hb_in = iter ? ts->hb64[(iter + 1) & 1] : nullptr; q.hb_out = ts->hb64[iter & 1];
        q.hb_tiles = nt + 1; q.ext_prev = ts->ext_prev; q.ext_next = ts->ext_next;
    } else {
        p.tables = ctx->d_tables; p.tl = ts->tl; p.mags = ts->mags;
        p.sig_in = ts->sig[iter & 1]; p.sig_out = ts->sig[(iter + 1) & 1];
        p.hb_in = (iter && iter != ts->lead) ? ts->hb[(iter + 1) & 1] : nullptr; p.hb_out = ts->hb[iter & 1];
        p.hb_tiles = nt + 1; p.ext_prev = ts->ext_prev; p.ext_next = ts->ext_next;
    }
    auto launch = [&](int grid, cudaStream_t st, int tile_lo, int tiles, int edge_mode, int e0, int e1) {
        if (f64) {
            q.tile_lo = tile_lo; q.tiles_in_launch = tiles; q.edge_mode = edge_mode; q.edge_tile0 = e0; q.edge_tile1 = e1;
            d64::k_gl_iter_f64<kHS><<<grid, kThreads, d64::kSmemBytes, st>>>(q);
        } else {
            p.tile_lo = tile_lo; p.tiles_in_launch = tiles; p.edge_mode = edge_mode; p.edge_tile0 = e0; p.edge_tile1 = e1;
            k_gl_iter<kHS><<<grid, kThreads, kGlSmemBytes, st>>>(p);
        }
        ctx->launches++;
    };
    // boundary tiles: tile 0 if a previous rank exists, tile nt-1 if a next rank exists
    int edges[2], n_edge = 0;
    if (ts->ext_prev) edges[n_edge++] = 0;
    if (ts->ext_next && !(n_edge && nt == 1)) edges[n_edge++] = nt - 1;
    const int lo = ts->ext_prev ? 1 : 0, hi = ts->ext_next ? nt - 1 : nt;      // interior [lo, hi)
    if (part == 0) {
        if (ts->have_comm) CU(cudaStreamWaitEvent(ctx->st, ts->ev_comm, 0));
        if (ts->have_edge) CU(cudaStreamWaitEvent(ctx->st, ts->ev_edge, 0));
        launch(nt, ctx->st, 0, nt, 0, 0, 0);
        CU(cudaEventRecord(ts->ev_edge, ctx->st)); ts->have_edge = true;
        CU(cudaEventRecord(ts->ev_int, ctx->st)); ts->have_int = true;
    } else if (part == 1) {
        if (ts->have_int) CU(cudaStreamWaitEvent(ts->st_edge, ts->ev_int, 0));
        if (ts->have_comm) CU(cudaStreamWaitEvent(ts->st_edge, ts->ev_comm, 0));
        if (n_edge > 0) launch(n_edge, ts->st_edge, 0, nt, 1, edges[0], edges[n_edge - 1]);
        CU(cudaEventRecord(ts->ev_edge, ts->st_edge)); ts->have_edge = true;
    } else {
        // interior of iteration `iter` needs the boundary tiles of iteration iter-1 (recorded before
        // this iteration's part-1 call re-records ev_edge; callers issue part 2 of iteration i-1
        // before part 1 of iteration i, so the wait below is enqueued by the previous part-1 call)
        if (hi > lo) launch(hi - lo, ctx->st, lo, hi - lo, 0, 0, 0);
        CU(cudaEventRecord(ts->ev_int, ctx->st)); ts->have_int = true;
        // the NEXT iteration's interior must not start before this iteration's boundary tiles finished
        if (ts->have_edge) CU(cudaStreamWaitEvent(ctx->st, ts->ev_edge, 0));
    }
    CU(cudaGetLastError());
    return 0;
}

int gomel_ts_halo_ptrs(gomel_ts* ts, int iter, float** send_tail, float** send_head, float** recv_tail, float** recv_head)
{
    if (!ts || iter < 0) return GOMEL_E_ARG;
    if (iter < ts->lead) {          // float64 iteration: the four pointers address 2816 DOUBLES each (gomel_ts_halo_elem_bytes)
        double* out = ts->sig64[(iter + 1) & 1];
        double* hbo = ts->hb64[iter & 1];
        if (send_tail) *send_tail = ts->ext_next ? (float*)(out + ts->n_local * kHop) : nullptr;
        if (recv_head) *recv_head = ts->ext_next ? (float*)(hbo + (long)ts->tl.n_tiles * kHalo) : nullptr;
        if (send_head) *send_head = ts->ext_prev ? (float*)hbo : nullptr;
        if (recv_tail) *recv_tail = ts->ext_prev ? (float*)out : nullptr;
        return 0;
    }
    float* out = ts->sig[(iter + 1) & 1];
    float* hbo = ts->hb[iter & 1];
    if (send_tail) *send_tail = ts->ext_next ? out + ts->n_local * kHop : nullptr;
    if (recv_head) *recv_head = ts->ext_next ? hbo + (long)ts->tl.n_tiles * kHalo : nullptr;
    if (send_head) *send_head = ts->ext_prev ? hbo : nullptr;
    if (recv_tail) *recv_tail = ts->ext_prev ? out : nullptr;
    return 0;
}

int gomel_ts_halo_elem_bytes(gomel_ts* ts, int iter)
{
    if (!ts || iter < 0) return GOMEL_E_ARG;
    return iter < ts->lead ? 8 : 4;
}

int gomel_ts_lead_iters(gomel_ts* ts) { return ts ? ts->lead : GOMEL_E_ARG; }

void* gomel_ts_comm_stream(gomel_ts* ts) { return ts ? (void*)ts->st_comm : nullptr; }

int gomel_ts_comm_begin(gomel_ts* ts, int iter)
{
    (void)iter;
    if (!ts) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    if (ts->have_edge) CU(cudaStreamWaitEvent(ts->st_comm, ts->ev_edge, 0));
    return 0;
}

int gomel_ts_comm_end(gomel_ts* ts, int iter)
{
    (void)iter;
    if (!ts) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    CU(cudaEventRecord(ts->ev_comm, ts->st_comm)); ts->have_comm = true;
    return 0;
}

int gomel_ts_finish(gomel_ts* ts, int iters, float* d_out_local)
{
    if (!ts || !d_out_local || iters < 0) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    CU(cudaStreamSynchronize(ts->st_edge));
    CU(cudaStreamSynchronize(ts->st_comm));
    CU(cudaStreamSynchronize(ctx->st));
    if (iters > 0 && iters <= ts->lead) {       // the last iteration was a float64 one
        if (int rc = ts_fold_f64(ts, iters - 1, d_out_local)) return rc;
        CU(cudaStreamSynchronize(ctx->st));
        return 0;
    }
    float* fin = ts->sig[iters & 1];
    const int t_first = ts->ext_prev ? 0 : 1;
    if (iters > 0 && ts->tl.n_tiles - t_first > 0) {
        SynParams p = {};
        k_halo_fix<<<(unsigned)(ts->tl.n_tiles - t_first), 256, 0, ctx->st>>>(fin, ts->hb[(iters - 1) & 1], ts->tl, kHop, kHalo,
                                                                            0, t_first, ts->tl.n_tiles + 1, p);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(d_out_local, fin, (size_t)ts->n_samples * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

// ---- phase.ISTFT of one long clip split by time (SURVEY 8(e) third case; phase/phase.go:93-133) -------------
// Each rank inverts its own frames; a sample next to a rank boundary is the sum of the earlier rank's tail partial
// and the later rank's head partial, and it belongs to the later rank: ONE transfer of 2816 floats per boundary
// (tail -> next rank), then the window-sum gain -- a function of the global sample index, max(sum w^2) in closed
// form from the host tables of prepare_gain -- is applied by the owner.
static SynParams ts_phase_params(gomel_ts* ts)
{
    gomel_ctx* ctx = ts->ctx;
    SynParams p = {};
    p.tables = ctx->d_tables; p.tl = ts->tl;
    p.sig_out = ts->sig[0]; p.hb_out = ts->hb[0];
    p.n_freqs = ts->cfg.n_freqs;
    p.gain_head = ctx->d_gain_head; p.gain_mid = ctx->d_gain_mid; p.gain_tail = ctx->d_gain_tail;
    p.head_len = ctx->gain_head_len; p.tail_len = ctx->gain_tail_len;
    p.hb_tiles = ts->tl.n_tiles + 1; p.tile_lo = 0; p.tiles_in_launch = ts->tl.n_tiles;
    p.ext_prev = ts->ext_prev; p.ext_next = ts->ext_next;
    p.gain_off = ts->sample_begin; p.total_len = kN + (ts->n_frames_total - 1) * (long)kHop;
    return p;
}

int gomel_ts_phase_istft(gomel_ts* ts, const float* d_spec_local)
{
    if (!ts || !d_spec_local) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    if (ts->cfg.n_freqs <= 0 || ts->cfg.n_freqs > kN / 2) return fail(ctx, GOMEL_E_ARG, "NumFreqs out of range");
    if (ts->tl.edge_first || ts->tl.edge_last) return fail(ctx, GOMEL_E_ARG, "phase ISTFT sessions use uniform tiles (edge_frames = 0)");
    if (int rc = prepare_gain(ctx, ts->n_frames_total, ts->cfg.volume_boost)) return rc;
    SynParams p = ts_phase_params(ts);
    p.spec = reinterpret_cast<const float2*>(d_spec_local);
    k_istft_phase<kHS><<<(unsigned)ts->tl.n_tiles, kThreads, kFwdSmemBytes, ctx->st>>>(p);
    ctx->launches++;
    CU(cudaEventRecord(ts->ev_edge, ctx->st)); ts->have_edge = true;      // the exchange waits for this (gomel_ts_comm_begin)
    ts->have_comm = false;
    CU(cudaGetLastError());
    return 0;
}

int gomel_ts_phase_halo_ptrs(gomel_ts* ts, float** send_tail, float** recv_tail)
{
    if (!ts) return GOMEL_E_ARG;
    if (send_tail) *send_tail = ts->ext_next ? ts->sig[0] + ts->n_local * kHop : nullptr;
    if (recv_tail) *recv_tail = ts->ext_prev ? ts->sig[0] : nullptr;
    return 0;
}

int gomel_ts_phase_finish(gomel_ts* ts, float* d_out_local)
{
    if (!ts || !d_out_local) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    if (ts->have_comm) CU(cudaStreamWaitEvent(ctx->st, ts->ev_comm, 0));
    const int t_first = ts->ext_prev ? 0 : 1;
    if (ts->tl.n_tiles - t_first > 0) {
        SynParams p = ts_phase_params(ts);
        k_halo_fix<<<(unsigned)(ts->tl.n_tiles - t_first), 256, 0, ctx->st>>>(ts->sig[0], ts->hb[0], ts->tl, kHop, kHalo, 1, t_first,
                                                                            ts->tl.n_tiles + 1, p);
        ctx->launches++;
    }
    CU(cudaMemcpyAsync(d_out_local, ts->sig[0], (size_t)ts->n_samples * 4, cudaMemcpyDeviceToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return 0;
}

int gomel_ts_create(gomel_ctx* ctx, const gomel_config* cfg, long n_frames_total, int rank, int world,
                    int tile_frames, gomel_ts** out)
{
    return gomel_ts_create2(ctx, cfg, n_frames_total, rank, world, tile_frames, 0, out);
}

int gomel_ts_sync(gomel_ts* ts)
{
    if (!ts) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    CU(cudaStreamSynchronize(ts->st_edge));
    CU(cudaStreamSynchronize(ts->st_comm));
    CU(cudaStreamSynchronize(ctx->st));
    return 0;
}

int gomel_copy_d2d(gomel_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream)
{
    if (!ctx) return GOMEL_E_ARG;
    Guard g(ctx);
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream ? (cudaStream_t)stream : ctx->st));
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------- in-library NCCL exchange (dlopen)
namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, char[128], int) = nullptr;      // ncclUniqueId is passed by value (128 bytes)
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
struct NcclId { char b[128]; };
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool load_nccl(std::string& err)
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok) return true;
    const char* env = getenv("GOMEL_NCCL_LIB");
    const char* names[3] = { env, "libnccl.so.2", "libnccl.so" };
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { err = "cannot dlopen libnccl.so.2 (set GOMEL_NCCL_LIB)"; return false; }
    auto sym = [&](const char* n) { return dlsym(g_nccl.handle, n); };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, char[128], int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        err = "libnccl is missing a required symbol";
        return false;
    }
    g_nccl_destroy = g_nccl.CommDestroy;
    g_nccl.ok = true;
    return true;
}
// ncclCommInitRank(ncclComm_t*, int nranks, ncclUniqueId commId /* by value */, int rank)
typedef int (*comm_init_fn)(void**, int, NcclId, int);
}  // namespace

extern "C" {

int gomel_nccl_unique_id(gomel_ctx* ctx, char id_out[128])
{
    if (!ctx || !id_out) return GOMEL_E_ARG;
    Guard g(ctx);
    std::string err;
    if (!load_nccl(err)) return fail(ctx, GOMEL_E_STATE, err);
    NcclId id;
    memset(&id, 0, sizeof id);
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail(ctx, GOMEL_E_CUDA, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    memcpy(id_out, id.b, 128);
    return 0;
}

int gomel_ts_nccl_init(gomel_ts* ts, const char id[128])
{
    if (!ts || !id) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    Guard g(ctx);
    std::string err;
    if (!load_nccl(err)) return fail(ctx, GOMEL_E_STATE, err);
    if (ts->nccl_comm) return 0;
    NcclId nid;
    memcpy(nid.b, id, 128);
    const int rc = ((comm_init_fn)(void*)g_nccl.CommInitRank)(&ts->nccl_comm, ts->world, nid, ts->rank);
    if (rc) return fail(ctx, GOMEL_E_CUDA, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return 0;
}

int gomel_ts_run_nccl(gomel_ts* ts, int first_iter, int n_iters, int overlap)
{
    if (!ts || first_iter < 0 || n_iters < 0) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    if (ts->world > 1 && !ts->nccl_comm) { Guard g(ctx); return fail(ctx, GOMEL_E_STATE, "gomel_ts_nccl_init has not been called"); }
    for (int it = first_iter; it < first_iter + n_iters; it++) {
        if (int rc = gomel_ts_iterate(ts, it, overlap ? 1 : 0)) return rc;
        {
            Guard g(ctx);
            if (ts->have_edge) CU(cudaStreamWaitEvent(ts->st_comm, ts->ev_edge, 0));
            float *send_tail, *send_head, *recv_tail, *recv_head;
            gomel_ts_halo_ptrs(ts, it, &send_tail, &send_head, &recv_tail, &recv_head);
            if (ts->world > 1) {
                const int dt = it < ts->lead ? 8 /* ncclDouble */ : 7 /* ncclFloat */;
                int rc = g_nccl.GroupStart();
                if (!rc && send_head) rc = g_nccl.Send(send_head, kHalo, dt, ts->rank - 1, ts->nccl_comm, ts->st_comm);
                if (!rc && recv_tail) rc = g_nccl.Recv(recv_tail, kHalo, dt, ts->rank - 1, ts->nccl_comm, ts->st_comm);
                if (!rc && send_tail) rc = g_nccl.Send(send_tail, kHalo, dt, ts->rank + 1, ts->nccl_comm, ts->st_comm);
                if (!rc && recv_head) rc = g_nccl.Recv(recv_head, kHalo, dt, ts->rank + 1, ts->nccl_comm, ts->st_comm);
                const int rc2 = g_nccl.GroupEnd();
                if (rc || rc2) return fail(ctx, GOMEL_E_CUDA, std::string("NCCL halo exchange: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc ? rc : rc2) : "error"));
            }
            CU(cudaEventRecord(ts->ev_comm, ts->st_comm)); ts->have_comm = true;
        }
        if (overlap) { if (int rc = gomel_ts_iterate(ts, it, 2)) return rc; }
    }
    return 0;
}

int gomel_ts_phase_run_nccl(gomel_ts* ts, const float* d_spec_local, float* d_out_local)
{
    if (!ts) return GOMEL_E_ARG;
    gomel_ctx* ctx = ts->ctx;
    if (ts->world > 1 && !ts->nccl_comm) { Guard g(ctx); return fail(ctx, GOMEL_E_STATE, "gomel_ts_nccl_init has not been called"); }
    if (int rc = gomel_ts_phase_istft(ts, d_spec_local)) return rc;
    {
        Guard g(ctx);
        CU(cudaStreamWaitEvent(ts->st_comm, ts->ev_edge, 0));
        float *send_tail, *recv_tail;
        gomel_ts_phase_halo_ptrs(ts, &send_tail, &recv_tail);
        if (ts->world > 1) {
            int rc = g_nccl.GroupStart();
            if (!rc && recv_tail) rc = g_nccl.Recv(recv_tail, kHalo, 7 /* ncclFloat */, ts->rank - 1, ts->nccl_comm, ts->st_comm);
            if (!rc && send_tail) rc = g_nccl.Send(send_tail, kHalo, 7, ts->rank + 1, ts->nccl_comm, ts->st_comm);
            const int rc2 = g_nccl.GroupEnd();
            if (rc || rc2) return fail(ctx, GOMEL_E_CUDA, std::string("NCCL halo exchange: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc ? rc : rc2) : "error"));
        }
        CU(cudaEventRecord(ts->ev_comm, ts->st_comm)); ts->have_comm = true;
    }
    return gomel_ts_phase_finish(ts, d_out_local);
}

}  // extern "C"
