// kernels_f64.cuh -- STRICT (float64) Griffin-Lim path, selected with GOMEL_FLAG_F64.
//
// Why it exists: Griffin-Lim (mel.ISTFT, mel/mel.go:76-139) is ill-conditioned -- a 6e-8 relative
// perturbation of the start signal alone moves the float64 result by 1e-6..2e-5 after 32 iterations
// on a 10 s clip, so a float32 pipeline lands anywhere between 7e-6 and 3e-4 of the float64 reference
// depending on the start signal (tests/tools/gl_seed_scan.py).  This path runs the same algorithm in
// float64 end to end and reproduces the reference to ~1e-12 for any start signal; it is the parity
// instrument, the float32 kernels (kernels.cuh) are the throughput path.
//
// Same FFT-4096 index algebra as fft4096.cuh (two real frames per complex transform, DIF forward /
// DIT inverse, in-place padded exchange, Hermitian partner by warp shuffle); no sliding registers:
// every CTA handles one frame pair, writes its two windowed synthesis frames to Y[frame][4096], and
// k_ola_f64 adds the overlapping frames in ascending frame order -- the reference's own summation
// order (mel/mel.go:115-125).
#pragma once
#include "fft4096.cuh"

namespace gomel {
namespace f64 {

constexpr int kXchgBytes64 = kXchgCells * 16;                    // double2 cells, 73,728 B
constexpr int kT1Cells64 = 4 * 256, kT2Cells64 = 4 * 16;         // powers 1,2,4,8 of each lane's root
constexpr int kTableBytes64 = (kT1Cells64 + kT2Cells64) * 16 + 4096 * 8;
constexpr int kSmemBytes64 = kTableBytes64 + kXchgBytes64;      // 123,904 B -> one CTA per SM

struct Smem64 { const double2* T1; const double2* T2; const double* win; double2* xb; };

__device__ __forceinline__ Smem64 carve(unsigned char* base)
{
    Smem64 s;
    s.T1 = reinterpret_cast<double2*>(base);
    s.T2 = s.T1 + kT1Cells64;
    s.win = reinterpret_cast<const double*>(s.T2 + kT2Cells64);
    s.xb = reinterpret_cast<double2*>(const_cast<double*>(s.win) + 4096);
    return s;
}

using c64 = double2;
__device__ __forceinline__ c64 mk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ c64 cadd(c64 a, c64 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ c64 csub(c64 a, c64 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ c64 cadd_i(c64 a, c64 b) { return mk(a.x - b.y, a.y + b.x); }
__device__ __forceinline__ c64 csub_i(c64 a, c64 b) { return mk(a.x + b.y, a.y - b.x); }
__device__ __forceinline__ c64 cmul(c64 a, double wr, double wi) { return mk(fma(-a.y, wi, a.x * wr), fma(a.y, wr, a.x * wi)); }
template <bool INV> __device__ __forceinline__ c64 cmul_tw(c64 a, c64 w) { return INV ? cmul(a, w.x, -w.y) : cmul(a, w.x, w.y); }

template <bool INV>
__device__ __forceinline__ void radix4(c64& a0, c64& a1, c64& a2, c64& a3)
{
    const c64 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    if (!INV) { a1 = csub_i(t1, t3); a3 = cadd_i(t1, t3); }
    else      { a1 = cadd_i(t1, t3); a3 = csub_i(t1, t3); }
}

template <bool INV>
__device__ __forceinline__ void radix16(c64 (&v)[16])
{
    constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
#pragma unroll
    for (int m0 = 0; m0 < 4; m0++) radix4<INV>(v[m0], v[m0 + 4], v[m0 + 8], v[m0 + 12]);
    const double sg = INV ? 1.0 : -1.0;                 // forward twiddles e^{-i...}
    v[5] = cmul(v[5], c1, sg * s1);    v[9] = cmul(v[9], h, sg * h);     v[13] = cmul(v[13], s1, sg * c1);
    v[6] = cmul(v[6], h, sg * h);      v[10] = cmul(v[10], 0.0, sg);     v[14] = cmul(v[14], -h, sg * h);
    v[7] = cmul(v[7], s1, sg * c1);    v[11] = cmul(v[11], -h, sg * h);  v[15] = cmul(v[15], -c1, -sg * s1);
    c64 o[16];
#pragma unroll
    for (int ka = 0; ka < 4; ka++) {
        c64 b0 = v[4 * ka], b1 = v[4 * ka + 1], b2 = v[4 * ka + 2], b3 = v[4 * ka + 3];
        radix4<INV>(b0, b1, b2, b3);
        o[ka] = b0; o[ka + 4] = b1; o[ka + 8] = b2; o[ka + 12] = b3;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = o[i];
}

template <bool INV>
__device__ __forceinline__ void apply_twiddles(c64 (&v)[16], const double2* T, int rowlen, int lane)
{
    // T[(i)][lane], i = 0..3 -> w^1, w^2, w^4, w^8
    const c64 w1 = T[lane], w2 = T[rowlen + lane], w4 = T[2 * rowlen + lane], w8 = T[3 * rowlen + lane];
    auto mul = [](c64 x, c64 y) { return cmul(x, y.x, y.y); };
    const c64 w3 = mul(w2, w1), w5 = mul(w4, w1), w6 = mul(w4, w2), w7 = mul(w4, w3);
    v[1] = cmul_tw<INV>(v[1], w1); v[2] = cmul_tw<INV>(v[2], w2); v[3] = cmul_tw<INV>(v[3], w3); v[4] = cmul_tw<INV>(v[4], w4);
    v[5] = cmul_tw<INV>(v[5], w5); v[6] = cmul_tw<INV>(v[6], w6); v[7] = cmul_tw<INV>(v[7], w7); v[8] = cmul_tw<INV>(v[8], w8);
    v[9] = cmul_tw<INV>(v[9], mul(w8, w1));   v[10] = cmul_tw<INV>(v[10], mul(w8, w2)); v[11] = cmul_tw<INV>(v[11], mul(w8, w3));
    v[12] = cmul_tw<INV>(v[12], mul(w8, w4)); v[13] = cmul_tw<INV>(v[13], mul(w8, w5)); v[14] = cmul_tw<INV>(v[14], mul(w8, w6));
    v[15] = cmul_tw<INV>(v[15], mul(w8, w7));
}

__device__ __forceinline__ void fft_fwd(c64 (&v)[16], const Smem64& s, const Lanes& L)
{
    radix16<false>(v);
    apply_twiddles<false>(v, s.T1, 256, L.t);
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) s.xb[k0 * kPlane + L.base_a] = v[k0];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = s.xb[L.base_b + r * kRow];
    radix16<false>(v);
    apply_twiddles<false>(v, s.T2, 16, L.n0b);
#pragma unroll
    for (int r = 0; r < 16; r++) s.xb[L.base_b + r * kRow] = v[r];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; c++) v[c] = s.xb[L.base_c + c];
    radix16<false>(v);
}

__device__ __forceinline__ void fft_inv(c64 (&v)[16], const Smem64& s, const Lanes& L)
{
    radix16<true>(v);
    apply_twiddles<true>(v, s.T2, 16, L.k1c);
#pragma unroll
    for (int c = 0; c < 16; c++) s.xb[L.base_c + c] = v[c];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = s.xb[L.base_b + r * kRow];
    radix16<true>(v);
#pragma unroll
    for (int r = 0; r < 16; r++) s.xb[L.base_b + r * kRow] = v[r];
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) v[k0] = s.xb[k0 * kPlane + L.base_a];
    apply_twiddles<true>(v, s.T1, 256, L.t);
    radix16<true>(v);
}

__device__ __forceinline__ c64 shfl2(c64 a, int src)
{
    return mk(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
}
// cmplx.Rect(M, cmplx.Phase(X)) (mel/mel.go:98-102): M * X/|X|, Phase(0) = 0 -> (M, 0)
__device__ __forceinline__ c64 subst(c64 X, double M)
{
    const double n = hypot(X.x, X.y);
    if (n > 0.0) { const double r = M / n; return mk(X.x * r, X.y * r); }
    return mk(M, 0.0);
}

// magnitudes, float64, natural bin order [frame][2049], NOT pre-scaled (the 1/N is applied after the IFFT
// like go-dsp does)
__global__ void __launch_bounds__(256) k_mags_from_mel_f64(const double* __restrict__ mel, double* __restrict__ mags,
                                                           const int* __restrict__ inv_lo, const int* __restrict__ inv_hi,
                                                           const double* __restrict__ inv_mod, int n_mels,
                                                           double tune_add, double tune_mul, long n_rows)
{
    extern __shared__ double e[];
    for (long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const double* m = mel + row * 2 * n_mels;
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n_mels; i += blockDim.x) e[i] = exp(m[i]);
        __syncthreads();
        double* out = mags + row * 2049;
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {
            const int lo = inv_lo[i], hi = inv_hi[i];
            const int nch = (i == 2047) ? 2 : 1;
            for (int l = 0; l < nch; l++) {
                double total = 0.0;
                if (lo == hi) total = e[2 * lo + l];
                else if (lo + 1 == hi && hi < n_mels) {
                    const double md = inv_mod[i];
                    total = e[2 * lo + l] * (1.0 - md);
                    total += e[2 * hi + l] * md;
                } else {
                    for (int k = lo; k < hi; k++) total += e[2 * k + l];
                    total /= (double)(hi - lo + 1);
                }
                out[l ? 2048 : i] = fabs((total - tune_add) / tune_mul);
            }
        }
    }
}

struct GL64Params {
    const double* tables;    // T1 | T2 | win (kTableBytes64)
    const double* sig;       // [ola]
    const double* mags;      // [frames][2049]
    double* Y;               // [frames][4096] windowed synthesis frames
    int n_frames; long ola;
};

template <int HS>
__global__ void __launch_bounds__(kThreads, 1) k_gl_pair_f64(const GL64Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem64 s = carve(smem_raw);
    const Lanes L = make_lanes();
    {
        double2* d = reinterpret_cast<double2*>(smem_raw);
        const double2* g = reinterpret_cast<const double2*>(p.tables);
        for (int i = L.t; i < kTableBytes64 / 16; i += kThreads) d[i] = g[i];
    }
    __syncthreads();
    constexpr int H = 256 * HS;
    const int t = L.t, fA = 2 * blockIdx.x, fB = fA + 1;
    const bool validB = fB < p.n_frames;
    c64 v[16];
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const int n = t + 256 * m;
        const long sa = (long)fA * H + n, sb = (long)fB * H + n;
        const double w = s.win[n];
        const double xa = sa < p.ola ? p.sig[sa] : 0.0;
        const double xb = (validB && sb < p.ola) ? p.sig[sb] : 0.0;
        v[m] = mk(xa * w, xb * w);
    }
    fft_fwd(v, s, L);
    {
        const double* mA = p.mags + (long)fA * 2049;
        const double* mB = mA + 2049;
        c64 zs[16];
#pragma unroll
        for (int i = 0; i < 16; i++) zs[i] = v[i];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = L.klow + 256 * j;
            const double ma = mA[k], mb = validB ? mB[k] : 0.0;
            const c64 P = shfl2(v[15 - j], L.src), z = v[j];
            const c64 ya = subst(mk(z.x + P.x, z.y - P.y), ma);
            const c64 yb = subst(mk(z.y + P.y, P.x - z.x), mb);
            v[j] = mk(ya.x - yb.y, ya.y + yb.x);
            v[15 - j] = shfl2(mk(ya.x + yb.y, yb.x - ya.y), L.src);
        }
        if (L.special) {
#pragma unroll
            for (int j = 0; j <= 8; j++) {
                const int jp = (16 - j) & 15, k = 256 * j;
                const double ma = mA[k], mb = validB ? mB[k] : 0.0;
                const c64 z = zs[j], P = zs[jp];
                const c64 ya = subst(mk(z.x + P.x, z.y - P.y), ma);
                const c64 yb = subst(mk(z.y + P.y, P.x - z.x), mb);
                v[j] = mk(ya.x - yb.y, ya.y + yb.x);
                if (jp != j) v[jp] = mk(ya.x + yb.y, yb.x - ya.y);
            }
        }
    }
    fft_inv(v, s, L);
    double* YA = p.Y + (long)fA * 4096;
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const int n = t + 256 * m;
        const double w = s.win[n];
        YA[n] = (v[m].x / 4096.0) * w;                 // fft.IFFT divides by N, then x window (mel/mel.go:116-121)
        if (validB) YA[4096 + n] = (v[m].y / 4096.0) * w;
    }
}

// newReconstructed[pos] += val, frames in ascending order (mel/mel.go:115-125)
__global__ void k_ola_f64(const double* __restrict__ Y, double* __restrict__ out, int n_frames, int hop, long ola)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < ola; i += step) {
        long f0 = (i - 4096 + hop) / hop;            // ceil((i - 4095) / hop)
        if (i < 4096) f0 = 0;
        long f1 = i / hop;
        if (f1 > n_frames - 1) f1 = n_frames - 1;
        double acc = 0.0;
        for (long f = f0; f <= f1; f++) acc += Y[f * 4096 + (i - f * hop)];
        out[i] = acc;
    }
}

}  // namespace f64
}  // namespace gomel
