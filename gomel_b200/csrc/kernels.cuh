// kernels.cuh -- sm_100a kernels of the gomel spectrogram hot path.
//
// All transforms process frames in PAIRS (frame A = 2p, frame B = 2p+1 ride the real and
// imaginary lane of one complex FFT-4096, fft4096.cuh).  A CTA walks a TILE of consecutive
// frames of one clip.  Because the hop (Window, 1280) is a multiple of the per-thread sample
// stride (256), the samples a thread needs for the next pair are the ones it already holds,
// shifted by 2*hop/256 slots: the analysis input window and the overlap-add accumulator both
// live in registers and slide; per pair a thread reads 2*hop/256 new samples and writes
// 2*hop/256 finished ones -- exactly the algorithmic bytes of SURVEY.md 8(d).
#pragma once
#include "fft4096.cuh"
#include <type_traits>

namespace gomel {

constexpr int kMagStride = 2052;     // floats per magnitude row: 2049 bins, 16-byte aligned rows
constexpr int kMagRmsCell = 2049;    // first pad cell of a row: the frame's rms magnitude (k_mags_from_mel)

// position of bin k = k0 + 16*k1 + 256*k2 inside a magnitude row: k2*256 + rho(k0)*16 + k1 with
// rho = rotate-left-by-1 of the 4-bit k0.  The digit-reversed spectrum layout of fft4096_fwd then
// reads a row with unit stride across each half-warp (coalesced from HBM), and the two half-warps
// of a warp (k0 and 16-k0, which differ in bit 3) land in different halves of the 32 smem banks.
__host__ __device__ __forceinline__ int mag_pos(int k)
{
    if (k >= 2048) return 2048;
    const int k0 = k & 15, k1 = (k >> 4) & 15;
    const int rho = ((k0 & 7) << 1) | (k0 >> 3);
    return (k & ~255) | (rho << 4) | k1;
}
// inverse of mag_pos for positions < 2048
__host__ __device__ __forceinline__ int mag_unpos(int pos)
{
    const int rho = (pos >> 4) & 15, k1 = pos & 15;
    const int k0 = (rho >> 1) | ((rho & 1) << 3);
    return (pos & ~255) | (k1 << 4) | k0;
}

// ------------------------------------------------------------------ async bulk copy (TMA unit) helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar`.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// ------------------------------------------------------------------ small conversion kernels
__global__ void k_f64_to_f32_pad(const double* __restrict__ in, long n_in, float* __restrict__ out, long n_out)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n_out; i += step) out[i] = (i < n_in) ? (float)in[i] : 0.0f;
}
__global__ void k_f32_to_f64(const float* __restrict__ in, double* __restrict__ out, long n, double scale)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = (double)in[i] * scale;
}
// 16-bit PCM samples as dumpwav writes them (mel/impl.go:195-232 -> beep wav.Encode, Precision 2): clamp to
// [-1, 1], scale by 2^15 - 1 in float64, convert like Go's int16() (truncation; NaN -> 0)
__global__ void k_f32_to_pcm16(const float* __restrict__ in, short* __restrict__ out, long n)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        double v = (double)in[i];
        if (v < -1.0) v = -1.0;
        if (v > 1.0) v = 1.0;
        out[i] = (v == v) ? (short)__double2int_rz(__dmul_rn(v, 32767.0)) : (short)0;
    }
}
// counter-based U[0,1) fill for the Griffin-Lim start signal when the caller injects none
// (mel/mel.go:80-83 draws rand.Float64(); same distribution, not bit-compatible with math/rand)
template <typename T>
__global__ void k_fill_uniform(T* __restrict__ out, long n, unsigned long long seed, long index_offset = 0)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + index_offset + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        out[i] = (T)((float)(z >> 40) * (1.0f / 16777216.0f));       // 24-bit grid: the same value as float or double
    }
}

// ------------------------------------------------------------------ tile geometry
struct Tiling {
    int n_frames;        // frames per clip
    int tile_frames;     // T, even: frames per (interior) tile
    int n_tiles;         // tiles per clip
    long sig_stride;     // floats between clips in signal buffers
    long sig_len;        // valid samples per clip (loads beyond read as 0, stores beyond dropped)
    // time-split sessions only: a short first / last tile (even, >= 4 frames; 0 = none) so that the tiles that
    // wait for a neighbouring rank are cheap while the interior tiles stay long
    int edge_first, edge_last;
};
// first frame of tile `tile` (tile == n_tiles gives n_frames)
__host__ __device__ __forceinline__ int tile_begin(const Tiling& tl, int tile)
{
    if (tile >= tl.n_tiles) return tl.n_frames;
    if (tl.edge_first > 0) { if (tile == 0) return 0; tile--; }
    const int b = tl.edge_first + tile * tl.tile_frames;
    const int last0 = tl.n_frames - tl.edge_last;            // start of the short last tile (== n_frames if none)
    return b < last0 ? b : last0;
}

__device__ __forceinline__ float sqrt_fast(float x)        // one MUFU.SQRT (max relative error 2^-23): plenty for 1e-5 parity
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// the two real spectra riding one complex transform, up to a common factor 2 (P = Z[N-k]):
//   2*XA = Z + conj P ,  2*XB = (Z - conj P)/i
__device__ __forceinline__ float2 split_a(float2 z, float2 P) { return __fadd2_rn(z, make_float2(P.x, -P.y)); }
__device__ __forceinline__ float2 split_b(float2 z, float2 P) { return __fadd2_rn(make_float2(z.y, -z.x), make_float2(P.y, P.x)); }
// Z'[k] = YA + i*YB ,  Z'[N-k] = conj(YA) + i*conj(YB)
__device__ __forceinline__ float2 join_lo(float2 ya, float2 yb) { return __fadd2_rn(ya, make_float2(-yb.y, yb.x)); }
__device__ __forceinline__ float2 join_hi(float2 ya, float2 yb) { return __fadd2_rn(make_float2(ya.x, -ya.y), make_float2(yb.y, yb.x)); }

// ------------------------------------------------------------------ K1+K2 / K1+K4: STFT forward
// Replaces gossp STFT.STFT + the magnitude loop + domel + spectral_normalize of mel.ToMel
// (mel/mel.go:50-70, mel/impl.go:310-345, :410-419)  [MODE_MEL]
// and gossp STFT.STFT + the (Im,Re) gather + shrink of phase.ToPhase
// (phase/phase.go:45-66, phase/impl.go:383-391)      [MODE_PHASE]
enum { MODE_MEL = 0, MODE_PHASE = 1, MODE_SPEC = 2 };
// The forward / phase-ISTFT kernels stage their spectra through a buffer of their own (not the FFT exchange
// buffer): the only CTA barriers per frame pair are the transform's one and the staging hand-over.
constexpr int kStageBytes = 2 * 2176 * 8;                  // two frames x 2048 padded float2 cells
constexpr int kFwdSmemBytes = kSmemBytes + kStageBytes;

struct FwdParams {
    const float* sig;        // [clips][sig_stride], zero padded per pad()
    const float4* tables;
    Tiling tl;
    // MODE_MEL
    const int* fwd_lo; const int* fwd_hi; const float* fwd_mod; int n_mels;
    float* mel_out;          // [clips][frames][n_mels][2]  ln(max(v,1e-5))
    // MODE_PHASE
    int n_freqs;
    float2* phase_out;       // [clips][frames][n_freqs] (Im X[j+1], Re X[j+1])
    // MODE_SPEC (tests): half spectrum [clips][frames][2049] natural order
    float2* spec_out;
};

// FS = frame slots (Resolut / 256).  FS = 16 is the native frame.  FS = 8 (Resolut 2048, the mel.NewMel default)
// runs the same 4096-point core on the frame zero-extended to 4096 samples: X2048[k] = X4096[2k].
template <int HS, int MODE, int FS = 16>
__global__ void __launch_bounds__(kThreads, 2) k_stft_fwd(const FwdParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    const Lanes L = make_lanes();
    load_tables(s, p.tables, L.t);
    static_assert(FS == 16 || (FS == 8 && MODE == MODE_MEL), "only the mel forward path has the Resolut 2048 variant");
    constexpr int NR = FS + HS, SH = 2 * HS, KEEP = FS - HS, H = 256 * HS;
    const int tile = blockIdx.x % p.tl.n_tiles, clip = blockIdx.x / p.tl.n_tiles;
    const int f0 = tile_begin(p.tl, tile);
    const int nf = tile_begin(p.tl, tile + 1) - f0;
    const int npairs = (nf + 1) >> 1;
    const float* __restrict__ sig = p.sig + (long)clip * p.tl.sig_stride + (long)f0 * H;
    const long lim = p.tl.sig_len - (long)f0 * H;
    const int t = L.t;
    if (MODE == MODE_MEL) {          // the pad cells of the magnitude staging rows stay zero for the whole tile
        float* z = reinterpret_cast<float*>(smem_raw + kSmemBytes);
        for (int i = t; i < kStageBytes / 4; i += kThreads) z[i] = 0.0f;
    }

    float raw[NR], nxt[SH];
#pragma unroll
    for (int j = 0; j < KEEP; j++) { const long o = j * 256 + t; raw[j] = (o < lim) ? __ldg(sig + o) : 0.0f; }
#pragma unroll
    for (int j = 0; j < SH; j++) { const long o = (KEEP + j) * 256 + t; nxt[j] = (o < lim) ? __ldg(sig + o) : 0.0f; }
    __syncthreads();

    // MODE_MEL: a thread's mel work items (frame of the pair, mel band) are the same for every pair of the tile;
    // their band descriptors stay in registers.  Items are ordered widest band first and dealt boustrophedon so
    // every thread gets about the same number of taps.
    constexpr int kHoist = 2;                      // rounds covered by registers: NumMels <= 256
    const int n_items = (MODE == MODE_MEL) ? 2 * p.n_mels : 0;
    int h_lo[kHoist], h_hi[kHoist], h_out[kHoist];
    float h_w[kHoist];
    if (MODE == MODE_MEL) {
#pragma unroll
        for (int q = 0; q < kHoist; q++) {
            const int i = q * kThreads + ((q & 1) ? kThreads - 1 - t : t);
            h_out[q] = -1; h_lo[q] = h_hi[q] = 0; h_w[q] = 0.0f;
            if (i < n_items) {
                const int mel = p.n_mels - 1 - (i >> 1);
                h_lo[q] = __ldg(p.fwd_lo + mel); h_hi[q] = __ldg(p.fwd_hi + mel);
                h_w[q] = (h_lo[q] + 1 == h_hi[q]) ? __ldg(p.fwd_mod + mel) : 1.0f / (float)(h_hi[q] - h_lo[q] + 1);
                h_out[q] = (i & 1) * p.n_mels + mel;
            }
        }
    }

    for (int pr = 0; pr < npairs; pr++) {
        const long off0 = (long)pr * 2 * H;
#pragma unroll
        for (int j = 0; j < SH; j++) raw[KEEP + j] = nxt[j];
        if (pr + 1 < npairs) {          // prefetch the next pair's new rows; consumed at the top of the next pass
            const long r0 = off0 + 2 * H + KEEP * 256;
            if (r0 + SH * 256 <= lim) {
#pragma unroll
                for (int j = 0; j < SH; j++) nxt[j] = __ldg(sig + r0 + j * 256 + t);
            } else {
#pragma unroll
                for (int j = 0; j < SH; j++) { const long o = r0 + j * 256 + t; nxt[j] = (o < lim) ? __ldg(sig + o) : 0.0f; }
            }
        }
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) {
            if (m < FS) { const float w = win_at<FS>(s.win, m, t); v[m] = make_float2(raw[m] * w, raw[m + HS] * w); }
            else v[m] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < KEEP; j++) raw[j] = raw[j + SH];

        fft4096_fwd(v, s, L);

        const int fA = f0 + 2 * pr;
        const bool validB = (fA + 1) < p.tl.n_frames;

        // split the two real spectra: XA[k] = (Z[k] + conj Z[N-k])/2, XB[k] = (Z[k] - conj Z[N-k])/(2i)
        // only slots k2 < 8 (k < 2048) and the Nyquist bin carry new information
        float2 xa[9], xb[9];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 P = shfl2(v[15 - j], L.src);                     // Z[N-k], unconditional shuffle
            xa[j] = __fmul2_rn(split_a(v[j], P), make_float2(0.5f, 0.5f));
            xb[j] = __fmul2_rn(split_b(v[j], P), make_float2(0.5f, 0.5f));
        }
        xa[8] = xb[8] = make_float2(0.f, 0.f);
        if (L.special) {            // klow == 0: bins 256*j pair with 256*(16-j) inside this thread; j = 8 is the Nyquist bin
#pragma unroll
            for (int j = 0; j <= 8; j++) {
                const float2 P = v[(16 - j) & 15];
                xa[j] = __fmul2_rn(split_a(v[j], P), make_float2(0.5f, 0.5f));
                xb[j] = __fmul2_rn(split_b(v[j], P), make_float2(0.5f, 0.5f));
            }
        }

        float2* const stg = reinterpret_cast<float2*>(smem_raw + kSmemBytes);
        if (MODE == MODE_MEL) {
            float* SA = reinterpret_cast<float*>(stg);
            float* SB = SA + 2184;
            // staged by bin of the configured transform: FS = 8 keeps the even bins of the 4096-point core
            if (FS == 16 || (L.klow & 1) == 0) {
                // padded cell of bin klow + 256 j: (k + (k >> 4)) is linear in j (256 j is a multiple of 16)
                const int k0 = L.klow >> (FS == 16 ? 0 : 1), q0 = k0 + (k0 >> 4);
                constexpr int qstep = (FS == 16) ? 272 : 136;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    SA[q0 + qstep * j] = sqrt_fast(fmaf(xa[j].x, xa[j].x, xa[j].y * xa[j].y));
                    SB[q0 + qstep * j] = sqrt_fast(fmaf(xb[j].x, xb[j].x, xb[j].y * xb[j].y));
                }
            }
            if (L.special) {
                constexpr int qn = FS * 128 + FS * 8;        // Nyquist bin FS*128, padded
                SA[qn] = sqrt_fast(fmaf(xa[8].x, xa[8].x, xa[8].y * xa[8].y));
                SB[qn] = sqrt_fast(fmaf(xb[8].x, xb[8].x, xb[8].y * xb[8].y));
            }
            __syncthreads();
            // domel (mel/impl.go:310-345): one work item = (frame, mel), both channels in one pass over the band:
            // ch0 sums |X[k]| for k in [lo,hi), ch1 sums |X[N-1-k]| = |X[k+1]| for the same k.  Items are ordered
            // widest band first and dealt boustrophedon so every thread gets about the same number of taps.
            float2* outp = reinterpret_cast<float2*>(p.mel_out) + ((long)clip * p.tl.n_frames + fA) * p.n_mels;
            auto do_item = [&](int o, int lo, int hi, float w) {       // o = fr * n_mels + mel
                const int fr = o >= p.n_mels;
                if (fr == 1 && !validB) return;
                const float* S = fr ? SB : SA;
                float t0 = 0.0f, t1 = 0.0f;
                if (lo + 1 == hi) {                                  // w = modlo
                    const float s0 = S[lo + (lo >> 4)], s1 = S[hi + (hi >> 4)], s2 = S[hi + 1 + ((hi + 1) >> 4)];
                    t0 = s0 * (1.0f - w); t0 += s1 * w;
                    t1 = s1 * (1.0f - w); t1 += s2 * w;
                } else if (hi > lo) {                                // w = 1 / (count + 1)
                    // The band is a contiguous run of the padded staging row; the pad cells inside it hold zeros
                    // (cleared once per CTA), so the run is summed without any index arithmetic.  ch1 covers the
                    // same bins shifted by one: the run minus its first bin plus the bin after its last.
                    int q = lo + (lo >> 4);
                    const int qe = (hi - 1) + ((hi - 1) >> 4);
                    const float first = S[q];
                    float sum = first;
#pragma unroll 4
                    for (q++; q <= qe; q++) sum += S[q];
                    t0 = sum;
                    t1 = (sum - first) + S[hi + (hi >> 4)];
                    t0 *= w; t1 *= w;
                }
                t0 = (t0 < 1e-5f) ? 1e-5f : t0;
                t1 = (t1 < 1e-5f) ? 1e-5f : t1;
                outp[o] = make_float2(logf(t0), logf(t1));
            };
#pragma unroll
            for (int q = 0; q < kHoist; q++)
                if (h_out[q] >= 0) do_item(h_out[q], h_lo[q], h_hi[q], h_w[q]);
            for (int q = kHoist; q * kThreads < n_items; q++) {      // NumMels > 256 only
                const int i = q * kThreads + ((q & 1) ? kThreads - 1 - t : t);
                if (i >= n_items) continue;
                const int mel = p.n_mels - 1 - (i >> 1);
                const int lo = p.fwd_lo[mel], hi = p.fwd_hi[mel];
                do_item((i & 1) * p.n_mels + mel, lo, hi, (lo + 1 == hi) ? p.fwd_mod[mel] : 1.0f / (float)(hi - lo + 1));
            }
        } else if (MODE == MODE_PHASE) {
            float2* SA = stg;
            float2* SB = stg + 2176;
            {
                // bin k = klow + 256 j -> entry e = k - 1 -> padded cell e + (e >> 4) = q0 + 272 j (arithmetic shift:
                // klow = 0 gives q0 = -2, whose j = 0 cell -- bin 0, dropped by ToPhase -- is skipped)
                const int e0 = L.klow - 1, q0 = e0 + (e0 >> 4);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (j > 0 || e0 >= 0) {
                        SA[q0 + 272 * j] = make_float2(xa[j].y, xa[j].x);
                        SB[q0 + 272 * j] = make_float2(xb[j].y, xb[j].x);
                    }
                }
            }
            if (L.special) {
                const int e = 2047, q = e + (e >> 4);
                SA[q] = make_float2(xa[8].y, xa[8].x);
                SB[q] = make_float2(xb[8].y, xb[8].x);
            }
            __syncthreads();
            float2* outp = p.phase_out + ((long)clip * p.tl.n_frames + fA) * p.n_freqs;
            const int nfq = p.n_freqs;
            {
                const int qt = t + (t >> 4);                  // entry t + 256 i sits in padded cell qt + 272 i
                for (int i = 0, e = t; e < nfq; i++, e += kThreads) outp[e] = SA[qt + 272 * i];
                if (validB)
                    for (int i = 0, e = t; e < nfq; i++, e += kThreads) outp[nfq + e] = SB[qt + 272 * i];
            }
        } else {
            float2* SA = stg;
            float2* SB = stg + 2176;       // 2049 + 128 pad = 2177 cells > 2176: bin 2048 goes straight out
            float2* outp = p.spec_out + ((long)clip * p.tl.n_frames + fA) * 2049;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = L.klow + 256 * j, q = k + (k >> 4);
                SA[q] = xa[j]; SB[q] = xb[j];
            }
            if (L.special) { outp[2048] = xa[8]; if (validB) outp[2049 + 2048] = xb[8]; }
            __syncthreads();
            for (int o = t; o < 2 * 2048; o += kThreads) {
                const int fr = (o >= 2048), e = o - fr * 2048;
                if (fr == 1 && !validB) continue;
                outp[fr * 2049 + e] = (fr ? SB : SA)[e + (e >> 4)];
            }
        }
        // no barrier here: the staging buffer is rewritten only after the next pair's transform barrier, and the
        // staging hand-over barrier above already ordered every stage-3 read before the next pattern-(a) write
    }
}

// ------------------------------------------------------------------ K3: mel -> GL target magnitudes
// Replaces spectral_denormalize + undomel + Mel.undospectrum (mel/impl.go:421-427, :347-384,
// :386-408) reduced to the 2049 magnitudes mel.ISTFT actually uses: |X[k]| = ch0[k] for k < 2048
// and ch1[2047] for k = 2048 (cmplx.Abs at mel/mel.go:99).  Output rows are pre-scaled by 1/N
// (the IFFT normalisation) and stored in mag_pos() order.  Arithmetic in float64.
constexpr int kMagsRowsPerPass = 4;       // frames handled together by one CTA pass (amortises the barriers)

// FS = 8 (Resolut 2048): bin k' of the 2048-point transform sits where bin 2k' of the 4096-point core is read,
// the odd core bins are zero, and the scale is 1/2048 (the core's unnormalised inverse of the even-bin spectrum
// is 2048 times the reference's IFFT).
// OUT = float: the float32 iteration kernel's rows.  OUT = double: the float64 lead iterations' rows (gl_f64.cuh):
// the reference's own division by TuneMul, then the exact power-of-two scale.
template <typename T, int FS = 16, typename OUT = float>
__global__ void __launch_bounds__(256) k_mags_from_mel(const T* __restrict__ mel, OUT* __restrict__ mags,
                                                       const int* __restrict__ inv_lo, const int* __restrict__ inv_hi,
                                                       const double* __restrict__ inv_mod, int n_mels,
                                                       double tune_add, double tune_mul, long n_rows,
                                                       float* __restrict__ mags_f32 = nullptr)
{
    // mags_f32 (OUT = double only): the float32 rows of the same frames, written in the same pass -- a Griffin-Lim run
    // under the precision policy needs both, and the second pass over the mel rows was 0.5 % of the bench step
    const bool dual = std::is_same<OUT, double>::value && mags_f32 != nullptr;
    extern __shared__ double e[];     // [kMagsRowsPerPass][n_mels][2]
    __shared__ double red[8][kMagsRowsPerPass];
    const int per_row = 2 * n_mels;
    // (v - TuneAdd) / TuneMul / N as one multiplication: exact for the default TuneMul = 1, otherwise within one
    // float64 ulp of the reference's division, far below the float32 rounding of the stored magnitude
    const double scale = (1.0 / tune_mul) * (1.0 / (256.0 * FS));
    for (long row0 = (long)blockIdx.x * kMagsRowsPerPass; row0 < n_rows; row0 += (long)gridDim.x * kMagsRowsPerPass) {
        const int nr = (int)((n_rows - row0 < kMagsRowsPerPass) ? n_rows - row0 : kMagsRowsPerPass);
        const T* m = mel + row0 * per_row;
        __syncthreads();
        for (int i = threadIdx.x; i < nr * per_row; i += blockDim.x) e[i] = exp((double)m[i]);
        __syncthreads();
        double ss[kMagsRowsPerPass];
#pragma unroll
        for (int r = 0; r < kMagsRowsPerPass; r++) ss[r] = 0.0;
        for (int pos = threadIdx.x; pos < 2048; pos += blockDim.x) {      // unit-stride writes in mag_pos order
            const int kc = mag_unpos(pos);
            if (FS == 8 && (kc & 1)) {
                for (int r = 0; r < nr; r++) {
                    mags[(row0 + r) * kMagStride + pos] = (OUT)0;
                    if (dual) mags_f32[(row0 + r) * kMagStride + pos] = 0.0f;
                }
                continue;
            }
            const int i = kc >> (FS == 16 ? 0 : 1);
            const int lo = inv_lo[i], hi = inv_hi[i];
            const bool copy = lo == hi, lerp = (lo + 1 == hi) && hi < n_mels;
            const double md = lerp ? inv_mod[i] : 0.0;
#pragma unroll
            for (int r = 0; r < kMagsRowsPerPass; r++) {
                if (r >= nr) break;
                const double* er = e + r * per_row;
                OUT* out = mags + (row0 + r) * kMagStride;
                const int nch = (i == FS * 128 - 1) ? 2 : 1;
                for (int l = 0; l < nch; l++) {
                    double total = 0.0;
                    if (copy) total = er[2 * lo + l];
                    else if (lerp) {
                        total = er[2 * lo + l] * (1.0 - md);
                        total += er[2 * hi + l] * md;
                    } else {
                        for (int k = lo; k < hi; k++) total += er[2 * k + l];
                        total /= (double)(hi - lo + 1);
                    }
                    double v = std::is_same<OUT, double>::value
                                   ? fabs((total - tune_add) / tune_mul) * (1.0 / (256.0 * FS))
                                   : fabs((total - tune_add) * scale);
                    out[l ? 2048 : pos] = (OUT)v;
                    if (dual) {          // exactly the value the OUT = float instantiation stores
                        v = fabs((total - tune_add) * scale);
                        mags_f32[(row0 + r) * kMagStride + (l ? 2048 : pos)] = (float)v;
                    }
                    ss[r] += v * v;
                }
            }
        }
        // the frame's rms target magnitude goes to the row's first pad cell (kMagRmsCell): the scale of the
        // float32 iterations' singular-bin guard (k_gl_iter<.., GUARD>); fixed reduction order, deterministic
#pragma unroll
        for (int r = 0; r < kMagsRowsPerPass; r++) {
            double v = ss[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][r] = v;
        }
        __syncthreads();
        if (threadIdx.x < nr) {
            double v = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[w][threadIdx.x];
            mags[(row0 + threadIdx.x) * kMagStride + kMagRmsCell] = (OUT)sqrt(v * (1.0 / 2049.0));
            if (dual) mags_f32[(row0 + threadIdx.x) * kMagStride + kMagRmsCell] = (float)sqrt(v * (1.0 / 2049.0));
        }
    }
}

// rms of a clip's frame rms values (fixed order): the clip-level scale of the singular-bin guard
__global__ void __launch_bounds__(128) k_clip_scale(const float* __restrict__ mags, float* __restrict__ scale, long n_frames)
{
    __shared__ float red[4];
    const float* m = mags + (long)blockIdx.x * n_frames * kMagStride + kMagRmsCell;
    float v = 0.0f;
    for (long f = threadIdx.x; f < n_frames; f += blockDim.x) { const float x = m[f * kMagStride]; v = fmaf(x, x, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) scale[blockIdx.x] = sqrtf((red[0] + red[1] + red[2] + red[3]) / (float)n_frames);
}

// ------------------------------------------------------------------ shared pieces of the synthesis kernels

struct SynParams {
    const float4* tables;
    Tiling tl;
    const float* sig_in;     // GL: previous signal [clips][sig_stride]
    float* sig_out;          // new signal
    const float* hb_in;      // head partials of the previous iteration [clips][n_tiles][halo], or null
    float* hb_out;           // head partials of this iteration
    const float* mags;       // GL: [clips][frames][kMagStride]
    // phase ISTFT
    const float2* spec;      // [clips][frames][n_freqs] (Im, Re)
    int n_freqs;
    const float* gain_head; const float* gain_mid; const float* gain_tail;   // window-sum normalisation
    int head_len, tail_len;
    // launch subset + time-split neighbours (config 5): defaults 0 = whole clip on this GPU
    int hb_tiles;            // tiles per clip in hb buffers (n_tiles + 1: last slot = head partial of the next rank)
    int tile_lo, tiles_in_launch;     // this launch covers tiles [tile_lo, tile_lo + tiles_in_launch)
    int edge_mode, edge_tile0, edge_tile1;   // edge_mode: grid = 1 or 2 CTAs mapped to these tiles
    int ext_prev, ext_next;  // a previous / next rank continues the clip beyond this buffer
    int clip0;               // first clip of this launch (a batch is split over concurrent streams by clip)
    // phase ISTFT of one rank's slice of a long clip: the window-sum gain is a function of the GLOBAL sample index
    long gain_off;           // global index of local sample 0 (a multiple of the hop); 0 = whole clip here
    long total_len;          // length of the whole clip's signal
    // k_gl_iter<.., GUARD = true>: per-clip running maximum of the singular-bin statistic (float bits, atomicMax)
    unsigned int* guard_stat;
};

__device__ __forceinline__ float rsqrt_fast(float x)       // one MUFU.RSQ; callers guarantee x >= 1e-36 (no denormal path)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c)     // one FMNMX3
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float2 subst_phase(float2 X, float M, float* ratio = nullptr)
{
    // cmplx.Rect(M, cmplx.Phase(X)) = M * X/|X|, and Phase(0) = 0 -> (M, 0)   (mel/mel.go:98-102).
    // Branch-free: Y = (X + (d,0)) * M / sqrt(|X|^2 + d^2) with d = 1e-18.  For X = 0 this is exactly the
    // reference's (M, 0); for any X the transforms can produce (|X| >> 1e-10) d is far below one ulp.
    const float n = fmaf(X.x, X.x, fmaf(X.y, X.y, 1e-36f));
    const float r = M * rsqrt_fast(n);
    if (ratio) *ratio = r;                         // M / |X| for the guard
    float2 y = __fmul2_rn(X, make_float2(r, r));
    y.x = fmaf(1e-18f, r, y.x);
    return y;
}
// ------------------------------------------------------------------ K5: one Griffin-Lim iteration
// Replaces one pass of the loop body of mel.ISTFT (mel/mel.go:85-136): frame gather x Hann ->
// FFTReal -> Rect(|S|, Phase(F)) -> conj symmetry -> IFFT -> x Hann -> overlap-add, NO window-sum
// normalisation (commented out in the reference, mel/mel.go:113,122,127-132).  Jacobi update:
// reads sig_in only, writes sig_out only.
// Tile edges: samples whose contributing frames straddle two tiles are produced as two partial
// sums -- the earlier tile's into sig_out (its tail), the later tile's into hb_out (its head) --
// and summed on load (a+b is commutative, so both readers see the same value).
constexpr int kGlMagBytes = 2 * kMagStride * 4;                 // two magnitude rows (frames A and B)
constexpr int kGlSmemBytes = kSmemBytes + kGlMagBytes + 16 + 256;   // + one mbarrier + the special coset's scratch

// GUARD: the singular-bin guard.  A float32 iteration decides the phase of a bin from a value with an absolute error
// of ~1e-7 of the frame's rms bin; where |X[k]| is that small while the target M[k] is not, the float32 and float64
// trajectories take different branches (profiles/r02_gl_guard.md).  The kernel records, per clip, the maximum over
// bins and frames of  M[k]/|X[k]| * rms(M of the frame)  -- the error a unit absolute perturbation of that bin
// injects, up to the clip's scale; the host side re-runs the float32 tail of the clips above a threshold in float64.
// Cost: one FMNMX per bin on the ALU pipe.
template <int HS, int FS = 16, bool GUARD = false>
__global__ void __launch_bounds__(kThreads, 2) k_gl_iter(const SynParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    float* const smag = reinterpret_cast<float*>(smem_raw + kSmemBytes);                 // [2][kMagStride]
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(smem_raw + kSmemBytes + kGlMagBytes);
    float2* const zsc = reinterpret_cast<float2*>(smem_raw + kSmemBytes + kGlMagBytes + 16);   // [2][16]
    const Lanes L = make_lanes();
    load_tables(s, p.tables, L.t);
    constexpr int NR = FS + HS, SH = 2 * HS, KEEP = FS - HS, H = 256 * HS, HALO = KEEP * 256;
    int tile, clip;
    if (p.edge_mode) { clip = 0; tile = blockIdx.x == 0 ? p.edge_tile0 : p.edge_tile1; }
    else { tile = p.tile_lo + blockIdx.x % p.tiles_in_launch; clip = p.clip0 + blockIdx.x / p.tiles_in_launch; }
    const int f0 = tile_begin(p.tl, tile);
    const int nf = tile_begin(p.tl, tile + 1) - f0;
    const int npairs = (nf + 1) >> 1;
    const int tile_len = nf * H;           // this tile's own length (every tile but a clip's last holds whole pairs)
    const long sbase = (long)f0 * H;
    const float* __restrict__ sin_ = p.sig_in + (long)clip * p.tl.sig_stride + sbase;
    float* __restrict__ sout = p.sig_out + (long)clip * p.tl.sig_stride + sbase;
    const long lim_l = p.tl.sig_len - sbase;
    const int lim = (int)(lim_l < 0x7fffff00L ? lim_l : 0x7fffff00L);     // tile-relative, fits int
    const bool has_prev = tile > 0 || p.ext_prev, has_next = (tile + 1) < p.tl.n_tiles || p.ext_next;
    const float* __restrict__ hin_own = p.hb_in ? p.hb_in + ((long)clip * p.hb_tiles + tile) * HALO : nullptr;
    const float* __restrict__ hin_next = hin_own ? hin_own + HALO : nullptr;
    float* __restrict__ hout = p.hb_out + ((long)clip * p.hb_tiles + tile) * HALO;
    const int t = L.t;

    // generic (tile edge) load/store of one 256-sample row at tile-relative offset `row`
    auto ld = [&](int row) -> float {
        const int o = row + t;
        float x = (o < lim) ? __ldg(sin_ + o) : 0.0f;
        if (hin_own) {
            if (has_prev && row < HALO) x += __ldg(hin_own + o);
            else if (has_next && row >= tile_len) x += __ldg(hin_next + (o - tile_len));
        }
        return x;
    };
    auto st = [&](int row, float val) {
        const int o = row + t;
        if (o >= lim) return;
        if (has_prev && row < HALO) hout[o] = val;
        else sout[o] = val;
    };
    // rows [row0, row0 + n*256) need no edge handling: inside the signal, outside both halo zones
    auto plain_rows = [&](int row0, int n) -> bool {
        const int end = row0 + n * 256;
        return end <= lim && (!has_prev || row0 >= HALO) && (!(has_next && hin_own) || end <= tile_len);
    };

    if (t == 0) mbar_init(bar, 1);
    float raw[NR], acc[NR], nxt[SH], win[FS];
#pragma unroll
    for (int j = 0; j < KEEP; j++) { raw[j] = ld(j * 256); acc[j] = 0.0f; }
#pragma unroll
    for (int j = 0; j < SH; j++) nxt[j] = ld((KEEP + j) * 256);

    const int idx_lo = mag_pos(L.klow);
    const float* __restrict__ mrow = p.mags + ((long)clip * p.tl.n_frames + f0) * kMagStride;
    __syncthreads();                // tables + mbarrier init visible
#pragma unroll
    for (int m = 0; m < FS; m++) win[m] = win_at<FS>(s.win, m, t);

    float gmax = 0.0f;
    for (int pr = 0; pr < npairs; pr++) {
        const int off0 = pr * 2 * H;
        const bool validB = (f0 + 2 * pr + 1) < p.tl.n_frames;
        // stage this pair's target magnitudes with one TMA bulk copy; it lands during the forward FFT.
        // (the buffer was last read in the previous pair's substitution stage, >= 2 barriers ago)
        if (t == 0) {
            const unsigned bytes = validB ? (unsigned)kGlMagBytes : (unsigned)(kMagStride * 4);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smag, mrow + (long)(2 * pr) * kMagStride, bytes, bar);
        }
#pragma unroll
        for (int j = 0; j < SH; j++) { raw[KEEP + j] = nxt[j]; acc[KEEP + j] = 0.0f; }
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; m++) {
            if (m < FS) v[m] = make_float2(raw[m] * win[m], raw[m + HS] * win[m]);
            else v[m] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < KEEP; j++) raw[j] = raw[j + SH];

        fft4096_fwd(v, s, L);

        // prefetch the next pair's new signal rows into registers; consumed at the top of the next pass
        if (pr + 1 < npairs) {
            const int r0 = off0 + 2 * H + KEEP * 256;
            if (plain_rows(r0, SH)) {
#pragma unroll
                for (int j = 0; j < SH; j++) nxt[j] = __ldg(sin_ + r0 + j * 256 + t);
            } else {
#pragma unroll
                for (int j = 0; j < SH; j++) nxt[j] = ld(r0 + j * 256);
            }
        }

        // magnitude substitution on both frames at once.  With P = Z[N-k]:
        //   2*XA[k] = Z + conj P,  2*XB[k] = (Z - conj P)/i ; Y = M * X/|X| ;
        //   Z'[k] = YA + i*YB ,  Z'[N-k] = conj(YA) + i*conj(YB)
        // Each thread does this for its LOWER slots (k < 2048) only and hands Z'[N-k] to the partner lane
        // that owns bin N-k (its slot 15-j): no bin is substituted twice.
        mbar_wait(bar, (unsigned)(pr & 1));
        {
            const float* __restrict__ mA = smag + idx_lo;
            const float* __restrict__ mB = smag + kMagStride + idx_lo;
            // The one thread with klow == 0 holds bins 256*s, whose Hermitian partners 256*(16-s) sit in ITS OWN
            // slots (16-s)&15 -- off the generic lane pattern.  It parks its 16 values in a 256-byte scratch; lanes
            // 1..9 of warp 0 each substitute one of the nine special pairs {s, 16-s} after the generic loop and
            // thread 0 picks the results up again.  Everything stays inside warp 0 (__syncwarp), no code variant.
            const bool w0 = (t >> 5) == 0;
            if (w0) {
                if (L.special) {
#pragma unroll
                    for (int i = 0; i < 16; i++) zsc[i] = v[i];
                }
                __syncwarp();
            }
            float2 nlo[8], nhi[8];               // results go to fresh registers: no in-place ordering constraints
            float ra = 0.0f, rb = 0.0f, qa[2], qb[2];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float ma = mA[j * 256];
                const float mb = validB ? mB[j * 256] : 0.0f;
                const float2 P = shfl2(v[15 - j], L.src);
                const float2 z = v[j];
                const float2 ya = subst_phase(split_a(z, P), ma, GUARD ? &qa[j & 1] : nullptr);
                const float2 yb = subst_phase(split_b(z, P), mb, GUARD ? &qb[j & 1] : nullptr);
                if (GUARD && (j & 1)) { ra = max3(ra, qa[0], qa[1]); rb = max3(rb, qb[0], qb[1]); }
                nlo[j] = join_lo(ya, yb);
                nhi[j] = shfl2(join_hi(ya, yb), L.src);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j] = nlo[j]; v[15 - j] = nhi[j]; }
            if (w0) {
                const int j = t - 1;                                   // lanes 1..9 -> special pair j = 0..8
                if (j >= 0 && j <= 8) {
                    const int jp = (16 - j) & 15, mi = (j == 8) ? 2048 : j * 256;
                    const float ma = smag[mi];
                    const float mb = validB ? smag[kMagStride + mi] : 0.0f;
                    const float2 z = zsc[j], P = zsc[jp];
                    const float2 ya = subst_phase(split_a(z, P), ma, GUARD ? &qa[0] : nullptr);
                    const float2 yb = subst_phase(split_b(z, P), mb, GUARD ? &qb[0] : nullptr);
                    if (GUARD) { ra = fmaxf(ra, qa[0]); rb = fmaxf(rb, qb[0]); }
                    zsc[16 + j] = join_lo(ya, yb);
                    if (jp != j) zsc[16 + jp] = join_hi(ya, yb);
                }
                __syncwarp();
                if (L.special) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = zsc[16 + i];
                }
            }
            if (GUARD) gmax = fmaxf(gmax, fmaxf(ra * smag[kMagRmsCell], rb * smag[kMagStride + kMagRmsCell]));
        }

        fft4096_inv(v, s, L);

        // the Hann coefficients are read once per pair: here for the synthesis window, then kept in
        // registers for the next pair's analysis window (not live across the transforms)
#pragma unroll
        for (int m = 0; m < FS; m++) {
            win[m] = win_at<FS>(s.win, m, t);
            acc[m] = fmaf(v[m].x, win[m], acc[m]);
            acc[m + HS] = fmaf(v[m].y, win[m], acc[m + HS]);
        }
        if (off0 + SH * 256 <= lim && (!has_prev || off0 >= HALO)) {
#pragma unroll
            for (int j = 0; j < SH; j++) sout[off0 + j * 256 + t] = acc[j];
        } else {
#pragma unroll
            for (int j = 0; j < SH; j++) st(off0 + j * 256, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < KEEP; j++) acc[j] = acc[j + SH];
    }
#pragma unroll
    for (int j = 0; j < KEEP; j++) st(npairs * 2 * H + j * 256, acc[j]);
    if (GUARD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        if ((t & 31) == 0 && gmax > 0.0f) atomicMax(p.guard_stat + clip, __float_as_uint(gmax));     // non-negative floats order like their bits
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------ K4+K6: phase ISTFT
// Replaces grow + Phase.undospectrum + phase.ISTFT + the VolumeBoost loop of phase.FromPhase
// (phase/impl.go:392-403, phase/phase.go:72-91, :93-133, :146-150).  The window-sum
// normalisation is data independent: gain tables (1/ws, 1/thr or 1, times VolumeBoost) are built
// on the host in float64.  Samples shared by two tiles are left un-normalised (partials in
// sig_out / hb_out) and finished by k_halo_fix.
// mid_idx = s_abs mod hop, supplied by the caller (cheap where hop is a compile-time constant)
__device__ __forceinline__ float gain_at(const SynParams& p, long s_abs, int mid_idx)
{
    s_abs += p.gain_off;
    if (s_abs < p.head_len) return __ldg(p.gain_head + s_abs);
    const long tail0 = p.total_len - p.tail_len;
    if (s_abs >= tail0) return __ldg(p.gain_tail + (s_abs - tail0));
    return __ldg(p.gain_mid + mid_idx);
}

template <int HS>
__global__ void __launch_bounds__(kThreads, 2) k_istft_phase(const SynParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve_smem(smem_raw);
    const Lanes L = make_lanes();
    load_tables(s, p.tables, L.t);
    constexpr int NR = 16 + HS, SH = 2 * HS, KEEP = 16 - HS, H = 256 * HS, HALO = KEEP * 256;
    const int tile = blockIdx.x % p.tl.n_tiles, clip = blockIdx.x / p.tl.n_tiles;
    const int f0 = tile_begin(p.tl, tile);
    const int nf = tile_begin(p.tl, tile + 1) - f0;
    const int npairs = (nf + 1) >> 1;
    const int tile_len = nf * H;
    const long sbase = (long)f0 * H;
    float* __restrict__ sout = p.sig_out + (long)clip * p.tl.sig_stride + sbase;
    const long lim = p.tl.sig_len - sbase;
    const bool has_prev = tile > 0 || p.ext_prev, has_next = (tile + 1) < p.tl.n_tiles || p.ext_next;
    float* __restrict__ hout = p.hb_out + ((long)clip * p.hb_tiles + tile) * HALO;
    const int t = L.t, nfq = p.n_freqs;

    auto st = [&](int row, float val) {
        const int o = row + t;
        if (o >= lim) return;
        if (has_prev && row < HALO) hout[o] = val;                    // head partial
        else if (has_next && row >= tile_len) sout[o] = val;          // tail partial
        else sout[o] = val * gain_at(p, sbase + o, o % H);      // sbase is a multiple of H
    };

    float acc[NR];
#pragma unroll
    for (int j = 0; j < KEEP; j++) acc[j] = 0.0f;
    // interior pairs: a pair's 2*hop finished samples start at a multiple of 2*hop, so the periodic part of the
    // window-sum gain a thread needs is the same ten values for every pair
    float gm[SH];
#pragma unroll
    for (int j = 0; j < SH; j++) gm[j] = __ldg(p.gain_mid + (j * 256 + t) % H);
    const long mid_lo = p.head_len - p.gain_off, mid_hi = p.total_len - p.tail_len - p.gain_off;     // local coordinates
    // the first kPre*256 entries of a pair's two spectrogram rows are prefetched into registers one pair ahead
    constexpr int kPre = 6;                                   // covers NumFreqs <= 768
    float2 pre[kPre];
    auto fetch = [&](int pr_) {
        const int fA_ = f0 + 2 * pr_;
        const bool vB_ = (fA_ + 1) < p.tl.n_frames;
        const float2* __restrict__ src_ = p.spec + ((long)clip * p.tl.n_frames + fA_) * nfq;
#pragma unroll
        for (int i = 0; i < kPre; i++) {
            const int o = t + i * kThreads;
            pre[i] = (o < 2 * nfq && (o < nfq || vB_)) ? __ldg(src_ + o) : make_float2(0.f, 0.f);
        }
    };
    fetch(0);
    __syncthreads();

    for (int pr = 0; pr < npairs; pr++) {
        const int off0 = pr * 2 * H;
        const int fA = f0 + 2 * pr;
        const bool validB = (fA + 1) < p.tl.n_frames;
#pragma unroll
        for (int j = KEEP; j < NR; j++) acc[j] = 0.0f;

        // stage the two spectrogram rows (contiguous in memory) into the exchange buffer
        const float2* __restrict__ src = p.spec + ((long)clip * p.tl.n_frames + fA) * nfq;
        float2* SA = reinterpret_cast<float2*>(smem_raw + kSmemBytes);
        float2* SB = SA + 2176;
#pragma unroll
        for (int i = 0; i < kPre; i++) {
            const int o = t + i * kThreads;
            if (o < 2 * nfq) { const int fr = (o >= nfq), e = o - fr * nfq; (fr ? SB : SA)[e + (e >> 4)] = pre[i]; }
        }
        for (int o = t + kPre * kThreads; o < 2 * nfq; o += kThreads) {
            const int fr = (o >= nfq), e = o - fr * nfq;
            const float2 x = (fr == 0 || validB) ? __ldg(src + o) : make_float2(0.f, 0.f);
            (fr ? SB : SA)[e + (e >> 4)] = x;
        }
        if (pr + 1 < npairs) fetch(pr + 1);
        __syncthreads();
        // X[j+1] = complex(realm0, realn1) = (entry.y, entry.x); entries >= n_freqs replicate the last kept one
        // (grow); X[0] = 0; X[2048] keeps only its real part.  Lower slots (k < 2048) hold Z'[k] = XA + i*XB, upper
        // slots Z'[k] = conj(XA[N-k]) + i*conj(XB[N-k]).  The 1/N of the inverse transform lives in the gain tables.
        float2 v[16];
        {
            const int top = nfq - 1, qtop = top + (top >> 4);
            const int elo = L.klow - 1;                 // entry of bin klow + 256 j          (lower slot j)
            const int ehi = 255 - L.klow;               // entry of bin 4096 - (klow + 256 j)  (upper slot j: + 256 (15 - j))
            // padded cell of entry e0 + 256 i is q(e0) + 272 i; entries beyond the kept ones read the last kept one (grow)
            // (arithmetic shift: elo = -1, the klow = 0 thread, gives 272 j - 2; its j = 0 value is replaced below)
            const int qlo = elo + (elo >> 4), qhi = ehi + (ehi >> 4);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int q = qlo + 272 * j;
                if (j == 0) q = max(q, 0);
                if (elo + 256 * j > top) q = qtop;
                const float2 ea = SA[q], eb = SB[q];
                v[j] = join_lo(make_float2(ea.y, ea.x), make_float2(eb.y, eb.x));
            }
#pragma unroll
            for (int j = 8; j < 16; j++) {
                const int q = (ehi + 256 * (15 - j) <= top) ? qhi + 272 * (15 - j) : qtop;
                const float2 ea = SA[q], eb = SB[q];
                v[j] = join_hi(make_float2(ea.y, ea.x), make_float2(eb.y, eb.x));
            }
            if (L.special) {                            // klow == 0: bin 0 is zero, bin 2048 (slot 8) is real
                v[0] = make_float2(0.f, 0.f);
                const int e = min(2047, top), q = e + (e >> 4);
                v[8] = make_float2(SA[q].y, SB[q].y);   // (Re XA, Re XB): Z'[2048] = Re XA + i Re XB
            }
        }
        fft4096_inv(v, s, L);   // (the staging hand-over barrier above also orders the previous pair's pattern-(a) reads
                                // before this pair's first exchange write)

#pragma unroll
        for (int m = 0; m < 16; m++) {
            const float w = win_at(s.win, m, t);
            acc[m] = fmaf(v[m].x, w, acc[m]);
            acc[m + HS] = fmaf(v[m].y, w, acc[m + HS]);
        }
        if (sbase + off0 >= mid_lo && sbase + off0 + SH * 256 <= mid_hi && (!has_prev || off0 >= HALO)) {
#pragma unroll
            for (int j = 0; j < SH; j++) sout[off0 + j * 256 + t] = acc[j] * gm[j];
        } else {
#pragma unroll
            for (int j = 0; j < SH; j++) st(off0 + j * 256, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < KEEP; j++) acc[j] = acc[j + SH];
    }
#pragma unroll
    for (int j = 0; j < KEEP; j++) st(npairs * 2 * H + j * 256, acc[j]);
}

// finishes the samples shared by two tiles: sig[s] = (sig[s] + hb[s]) * gain, for the head regions of
// tiles t_first .. n_tiles-1 (t_first = 0 when a previous rank's tail partial sits in sig[0..halo)).
// One CTA per (clip, tile) head region: no per-element index arithmetic.
__global__ void k_halo_fix(float* __restrict__ sig, const float* __restrict__ hb, Tiling tl, int hop, int halo,
                           int use_gain, int t_first, int hb_tiles, SynParams gp)
{
    const int nt = tl.n_tiles - t_first;
    const int clip = blockIdx.x / nt, tile = blockIdx.x % nt + t_first;
    const long s0 = (long)tile_begin(tl, tile) * hop;            // a multiple of hop
    float* __restrict__ d = sig + (long)clip * tl.sig_stride + s0;
    const float* __restrict__ h = hb + ((long)clip * hb_tiles + tile) * halo;
    const long room = tl.sig_len - s0;
    const int n = (int)(room < halo ? (room < 0 ? 0 : room) : halo);
    // interior regions (the usual case): whole region present, 16-byte aligned, gain from the periodic table only
    const long g0 = s0 + gp.gain_off;
    const bool mid_only = !use_gain || (g0 >= gp.head_len && g0 + halo <= gp.total_len - gp.tail_len);
    if (n == halo && mid_only && (halo & 3) == 0 && (hop & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(h)) & 15) == 0) {
        float4* d4 = reinterpret_cast<float4*>(d);
        const float4* h4 = reinterpret_cast<const float4*>(h);
        for (int o4 = threadIdx.x; o4 < halo / 4; o4 += blockDim.x) {
            float4 x = d4[o4];
            const float4 y = __ldg(h4 + o4);
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
            if (use_gain) {
                int m = o4 * 4;                                   // position inside the hop period (s0 is a multiple of hop)
                while (m >= hop) m -= hop;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gp.gain_mid + m));
                x.x *= g.x; x.y *= g.y; x.z *= g.z; x.w *= g.w;
            }
            d4[o4] = x;
        }
        return;
    }
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        float x = d[o] + h[o];
        if (use_gain) x *= gain_at(gp, s0 + o, o % hop);
        d[o] = x;
    }
}

// ------------------------------------------------------------------ K7: Image (dumpbuffer)
// Replaces mel dumpbuffer (mel/impl.go:16-44) and phase dumpbuffer (phase/impl.go:15-43):
// per-channel min/max with the reference's sentinels, truncating 8-bit quantisation, float64.
__global__ void k_minmax_f64(const double* __restrict__ buf, long n_entries, double init_max, double init_min,
                             double* __restrict__ partial /* [grid][4] */)
{
    __shared__ double sm[4][256];
    double mx0 = init_max, mx1 = init_max, mn0 = init_min, mn1 = init_min;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n_entries; i += step) {
        const double a = buf[2 * i], b = buf[2 * i + 1];
        if (a > mx0) mx0 = a;
        if (a < mn0) mn0 = a;
        if (b > mx1) mx1 = b;
        if (b < mn1) mn1 = b;
    }
    sm[0][threadIdx.x] = mx0; sm[1][threadIdx.x] = mx1; sm[2][threadIdx.x] = mn0; sm[3][threadIdx.x] = mn1;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            if (sm[0][threadIdx.x + w] > sm[0][threadIdx.x]) sm[0][threadIdx.x] = sm[0][threadIdx.x + w];
            if (sm[1][threadIdx.x + w] > sm[1][threadIdx.x]) sm[1][threadIdx.x] = sm[1][threadIdx.x + w];
            if (sm[2][threadIdx.x + w] < sm[2][threadIdx.x]) sm[2][threadIdx.x] = sm[2][threadIdx.x + w];
            if (sm[3][threadIdx.x + w] < sm[3][threadIdx.x]) sm[3][threadIdx.x] = sm[3][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int c = 0; c < 4; c++) partial[blockIdx.x * 4 + c] = sm[c][0];
}
__global__ void k_minmax_final(const double* __restrict__ partial, int n, double init_max, double init_min,
                               double* __restrict__ out4)
{
    if (threadIdx.x || blockIdx.x) return;
    double mx0 = init_max, mx1 = init_max, mn0 = init_min, mn1 = init_min;
    for (int i = 0; i < n; i++) {
        if (partial[4 * i] > mx0) mx0 = partial[4 * i];
        if (partial[4 * i + 1] > mx1) mx1 = partial[4 * i + 1];
        if (partial[4 * i + 2] < mn0) mn0 = partial[4 * i + 2];
        if (partial[4 * i + 3] < mn1) mn1 = partial[4 * i + 3];
    }
    out4[0] = mx0; out4[1] = mx1; out4[2] = mn0; out4[3] = mn1;
}
// Go int(f) on amd64 then uint16(): truncation, NaN/overflow -> 0x8000000000000000 -> low bits 0
__device__ __forceinline__ unsigned go_trunc_u16(double f)
{
    if (!(f == f) || f >= 9223372036854775808.0 || f < -9223372036854775808.0) return 0u;
    return (unsigned)((unsigned long long)(long long)f & 0xffffull);
}
__global__ void k_quantise_u16(const double* __restrict__ buf, long n_entries, const double* __restrict__ mm,
                               unsigned short* __restrict__ out)
{
    const double mx0 = mm[0], mx1 = mm[1], mn0 = mm[2], mn1 = mm[3];
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n_entries; i += step) {
        const double v0 = (buf[2 * i] - mn0) / (mx0 - mn0);
        const double v1 = (buf[2 * i + 1] - mn1) / (mx1 - mn1);
        out[i] = (unsigned short)((go_trunc_u16(255 * v0) | (go_trunc_u16(255 * v1) << 8)) & 0xffffu);
    }
}

// ------------------------------------------------------------------ K7b: PNG pixel arithmetic
// Replaces the quantisation loops of mel dumpimage (mel/impl.go:138-181) and phase dumpimage
// (phase/impl.go:170-266) and the de-quantisation of the two loadpng (mel/impl.go:92-112,
// phase/impl.go:98-147).  float64 with explicitly un-fused IEEE operations so pixel bytes match
// the Go float64 arithmetic bit for bit; zlib/PNG container work stays on the host.
__global__ void k_asinh_passes(double* __restrict__ buf, long n, int passes, int inverse)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        double v = buf[i];
        for (int p = 0; p < passes; p++) v = inverse ? sinh(v) : asinh(v);
        buf[i] = v;
    }
}
__device__ __forceinline__ unsigned go_trunc_mask(double f, unsigned long long mask)
{
    if (!(f == f) || f >= 9223372036854775808.0 || f < -9223372036854775808.0) return 0u;
    return (unsigned)((unsigned long long)(long long)f & mask);
}
// out3[i] = (R, G, B) for entry i (buffer order x*mels+y); mm = max0,max1,min0,min1
__global__ void k_quantise_rgb(const double* __restrict__ buf, long n_entries, const double* __restrict__ mm,
                               int maxval, int blue_wrap, unsigned short* __restrict__ out3)
{
    const double mx0 = mm[0], mx1 = mm[1], mn0 = mm[2], mn1 = mm[3], mv = (double)maxval;
    const unsigned long long mask = (unsigned long long)maxval;      // 255 or 65535: uint8()/uint16() wrap
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n_entries; i += step) {
        const double v0 = __ddiv_rn(__dsub_rn(buf[2 * i], mn0), __dsub_rn(mx0, mn0));
        const double v1 = __ddiv_rn(__dsub_rn(buf[2 * i + 1], mn1), __dsub_rn(mx1, mn1));
        out3[3 * i + 0] = (unsigned short)go_trunc_mask(__dmul_rn(mv, v0), mask);
        out3[3 * i + 1] = (unsigned short)go_trunc_mask(__dmul_rn(mv, v1), mask);
        out3[3 * i + 2] = blue_wrap ? (unsigned short)go_trunc_mask(__dmul_rn(mv, -v0), mask) : (unsigned short)0;
    }
}
// rg[i] = (R, G) pixel values (already >>8 for 8-bit images); out[i] = px/maxval*(max-min)+min, then sinh passes
__global__ void k_dequantise(const unsigned short* __restrict__ rg, long n_entries, double maxval, double mx0,
                             double mx1, double mn0, double mn1, int passes, double* __restrict__ out)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n_entries; i += step) {
        double a = __dadd_rn(__dmul_rn(__ddiv_rn((double)rg[2 * i], maxval), __dsub_rn(mx0, mn0)), mn0);
        double b = __dadd_rn(__dmul_rn(__ddiv_rn((double)rg[2 * i + 1], maxval), __dsub_rn(mx1, mn1)), mn1);
        for (int p = 0; p < passes; p++) { a = sinh(a); b = sinh(b); }
        out[2 * i] = a; out[2 * i + 1] = b;
    }
}

}  // namespace gomel
