// gl_f64.cuh -- fused float64 Griffin-Lim iteration kernel, sm_100a.
//
// Why it exists.  Griffin-Lim (mel.ISTFT, mel/mel.go:76-139) is ill-conditioned in its FIRST iterations: a
// rounding error injected in iteration 0 or 1 ends ~300x larger after 32 iterations, one injected in
// iteration 3 ~15x, one injected after iteration 8 not at all (profiles/r02_gl_parity_sweep.md).  A float32
// pipeline therefore misses the 1e-4 tolerance for some start signals, while float64 LEAD iterations followed
// by float32 iterations stay two orders of magnitude inside it.  This kernel runs those lead iterations (and
// the whole loop under GOMEL_FLAG_F64) in float64 end to end: signal, window, transforms, magnitudes,
// overlap-add.
//
// Same algorithm and index algebra as k_gl_iter (kernels.cuh): two real frames per complex FFT-4096, three
// radix-16 register butterflies per thread, DIF forward / DIT inverse with the spectrum left in digit-reversed
// order, in-place padded exchange buffer, Hermitian partner by warp shuffle.  What differs, because a complex
// value now costs four registers and B200's FP64 units run at half the FP32 lane rate:
//  * no sliding registers: a pair's 21 input rows are (re)loaded per pair (the 2.1x re-read hits L2; the next
//    pair's new rows and magnitude lines are prefetched into L2 one pair ahead), and the overlap-add is carried
//    through the output buffer itself -- first touch of a row is a plain store, later contributions are
//    fire-and-forget RED.ADD.F64 (each address is touched by ONE thread of ONE CTA, in ascending frame order:
//    the reference's own summation order, mel/mel.go:115-125, and deterministic);
//  * exchange rows are padded to 17 cells (16-byte cells): all three access patterns stay conflict free per
//    quarter-warp; 103.9 KB shared memory and <= 128 registers give 2 CTAs per SM.
#pragma once
#include "kernels.cuh"

namespace gomel {
namespace d64 {

constexpr int kRow = 17;                                   // double2 cells per 16-cell row (1 pad)
constexpr int kPlane = 16 * kRow;                          // 272
constexpr int kXchgCells = 16 * kPlane;                    // 4352
constexpr int kXchgBytes = kXchgCells * 16;                // 69,632 B
constexpr int kT1Cells = 4 * 256;                          // stage 1: powers 1,2,4,8 of each lane's root W4096^t
constexpr int kT2Cells = 4 * 16;                           // stage 2: powers 1,2,4,8 of W256^n0
constexpr int kWinCells = 2048;                            // first half of the symmetric Hann window, [m][t]
constexpr int kTableBytes = (kT1Cells + kT2Cells) * 16 + kWinCells * 8;      // 33,792 B
constexpr int kScratchBytes = 32 * 16;                     // the special coset's hand-over (warp 0)
constexpr int kSmemBytes = kTableBytes + kXchgBytes + kScratchBytes;         // 103,936 B -> 2 CTAs / SM

using c64 = double2;

struct Smem { const c64* T1; const c64* T2; const double* win; c64* xb; c64* zsc; };

__device__ __forceinline__ Smem carve(unsigned char* base)
{
    Smem s;
    s.T1 = reinterpret_cast<const c64*>(base);
    s.T2 = s.T1 + kT1Cells;
    s.win = reinterpret_cast<const double*>(s.T2 + kT2Cells);
    s.xb = reinterpret_cast<c64*>(base + kTableBytes);
    s.zsc = reinterpret_cast<c64*>(base + kTableBytes + kXchgBytes);
    return s;
}

// Hann coefficient of sample n = t + 256*m of an FS*256-sample frame (symmetric window, first half stored)
template <int FS>
__device__ __forceinline__ double win_at(const double* win, int m, int t)
{
    return (m < FS / 2) ? win[m * 256 + t] : win[(FS - 1 - m) * 256 + (255 - t)];
}

// the thread <-> digit mapping of fft4096.cuh with this file's row padding
__device__ __forceinline__ Lanes make_lanes64()
{
    Lanes L = make_lanes();
    L.base_a = L.t + (L.t >> 4);
    L.base_b = L.k0c * kPlane + L.n0b;
    L.base_c = L.k0c * kPlane + L.k1c * kRow;
    return L;
}

__device__ __forceinline__ c64 mk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ c64 cadd(c64 a, c64 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ c64 csub(c64 a, c64 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ c64 cadd_i(c64 a, c64 b) { return mk(a.x - b.y, a.y + b.x); }      // a + i*b
__device__ __forceinline__ c64 csub_i(c64 a, c64 b) { return mk(a.x + b.y, a.y - b.x); }      // a - i*b
__device__ __forceinline__ c64 cfma_s(c64 b, double s, c64 a) { return mk(fma(b.x, s, a.x), fma(b.y, s, a.y)); }   // a + s*b
__device__ __forceinline__ c64 cmul(c64 a, double wr, double wi) { return mk(fma(-a.y, wi, a.x * wr), fma(a.x, wi, a.y * wr)); }
template <bool INV> __device__ __forceinline__ c64 cmul_tw(c64 a, c64 w) { return INV ? cmul(a, w.x, -w.y) : cmul(a, w.x, w.y); }

template <bool INV>
__device__ __forceinline__ void radix4(c64& a0, c64& a1, c64& a2, c64& a3)
{
    const c64 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    if (!INV) { a1 = csub_i(t1, t3); a3 = cadd_i(t1, t3); }
    else      { a1 = cadd_i(t1, t3); a3 = csub_i(t1, t3); }
}

// multiply by W16^e (forward) or its conjugate (INV); e in {1,3,9}
template <bool INV, int E>
__device__ __forceinline__ c64 mul_w16(c64 a)
{
    constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
    if (E == 1) return INV ? cmul(a, c1, s1) : cmul(a, c1, -s1);
    if (E == 3) return INV ? cmul(a, s1, c1) : cmul(a, s1, -c1);
    /* E == 9 */ return INV ? cmul(a, -c1, -s1) : cmul(a, -c1, s1);
}

// 4-point DFT over (a0, h*u1, W16^4*r2, h*u3): u1, u3 carry a pending factor h = sqrt(1/2), r2 the pending
// rotation W16^4 = -i (forward) / +i (inverse); both are folded into the adds (cf. fft4096.cuh)
template <bool INV>
__device__ __forceinline__ void radix4_h13(c64& a0, c64& u1, c64& r2, c64& u3)
{
    constexpr double h = 0.70710678118654752440;
    const c64 t0 = INV ? cadd_i(a0, r2) : csub_i(a0, r2);
    const c64 t1 = INV ? csub_i(a0, r2) : cadd_i(a0, r2);
    const c64 s2 = cadd(u1, u3), s3 = csub(u1, u3);
    a0 = cfma_s(s2, h, t0);
    r2 = cfma_s(s2, -h, t0);
    if (!INV) { u1 = mk(fma(s3.y, h, t1.x), fma(s3.x, -h, t1.y)); u3 = mk(fma(s3.y, -h, t1.x), fma(s3.x, h, t1.y)); }   // t1 -+ i*h*s3
    else      { u1 = mk(fma(s3.y, -h, t1.x), fma(s3.x, h, t1.y)); u3 = mk(fma(s3.y, h, t1.x), fma(s3.x, -h, t1.y)); }
}
// 4-point DFT whose input u2 carries a pending factor h
template <bool INV>
__device__ __forceinline__ void radix4_h2(c64& a0, c64& a1, c64& u2, c64& a3)
{
    constexpr double h = 0.70710678118654752440;
    const c64 t0 = cfma_s(u2, h, a0), t1 = cfma_s(u2, -h, a0);
    const c64 t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    u2 = csub(t0, t2);
    if (!INV) { a1 = csub_i(t1, t3); a3 = cadd_i(t1, t3); }
    else      { a1 = cadd_i(t1, t3); a3 = csub_i(t1, t3); }
}

// 16-point DFT in registers, natural order in and out: v[k] <- sum_m v[m] W16^{mk}
template <bool INV>
__device__ __forceinline__ void radix16(c64 (&v)[16])
{
#pragma unroll
    for (int m0 = 0; m0 < 4; m0++) radix4<INV>(v[m0], v[m0 + 4], v[m0 + 8], v[m0 + 12]);
    // inner twiddles W16^{m0*ka}: W16^2 = h(1 -+ i) and W16^6 = h(-1 -+ i) keep their factor h pending,
    // W16^4 = -+i stays pending entirely (folded into the second step)
    v[5]  = mul_w16<INV, 1>(v[5]);   v[13] = mul_w16<INV, 3>(v[13]);
    v[7]  = mul_w16<INV, 3>(v[7]);   v[15] = mul_w16<INV, 9>(v[15]);
    auto w2 = [](c64 a) { return INV ? cadd_i(a, a) : csub_i(a, a); };                               // a(1 -+ i)
    auto w6 = [](c64 a) { return INV ? mk(-a.x - a.y, a.x - a.y) : mk(a.y - a.x, -a.x - a.y); };     // a(-1 -+ i)
    v[6] = w2(v[6]); v[9] = w2(v[9]); v[11] = w6(v[11]); v[14] = w6(v[14]);
    c64 o[16];
    {
        c64 b0 = v[0], b1 = v[1], b2 = v[2], b3 = v[3];
        radix4<INV>(b0, b1, b2, b3);
        o[0] = b0; o[4] = b1; o[8] = b2; o[12] = b3;
    }
    {
        c64 b0 = v[4], b1 = v[5], b2 = v[6], b3 = v[7];
        radix4_h2<INV>(b0, b1, b2, b3);
        o[1] = b0; o[5] = b1; o[9] = b2; o[13] = b3;
    }
    {
        c64 b0 = v[8], b1 = v[9], b2 = v[10], b3 = v[11];
        radix4_h13<INV>(b0, b1, b2, b3);
        o[2] = b0; o[6] = b1; o[10] = b2; o[14] = b3;
    }
    {
        c64 b0 = v[12], b1 = v[13], b2 = v[14], b3 = v[15];
        radix4_h2<INV>(b0, b1, b2, b3);
        o[3] = b0; o[7] = b1; o[11] = b2; o[15] = b3;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = o[i];
}

// external twiddles v[k] *= w^k (forward) or conj(w)^k (inverse); the table holds w^1, w^2, w^4, w^8 and the other
// eleven powers are products.  Measured on both stages against full tables: the kernel's shared-memory pipe is as
// busy as its FP64 pipe (54 % / 51 %) and a table load costs more than the 4-instruction product it saves (a full
// stage-2 table: +88 shared-memory wavefronts, -88 FP64 instructions per pair, 5 % slower).
template <bool INV>
__device__ __forceinline__ void apply_twiddles(c64 (&v)[16], const c64* T, int rowlen, int lane)
{
    const c64 w1 = T[lane], w2 = T[rowlen + lane], w4 = T[2 * rowlen + lane], w8 = T[3 * rowlen + lane];
    auto mul = [](c64 x, c64 y) { return cmul(x, y.x, y.y); };
    v[1] = cmul_tw<INV>(v[1], w1); v[2] = cmul_tw<INV>(v[2], w2); v[4] = cmul_tw<INV>(v[4], w4); v[8] = cmul_tw<INV>(v[8], w8);
    const c64 w3 = mul(w2, w1);
    v[3] = cmul_tw<INV>(v[3], w3);
    { const c64 w5 = mul(w4, w1); v[5] = cmul_tw<INV>(v[5], w5); v[13] = cmul_tw<INV>(v[13], mul(w8, w5)); }
    { const c64 w6 = mul(w4, w2); v[6] = cmul_tw<INV>(v[6], w6); v[14] = cmul_tw<INV>(v[14], mul(w8, w6)); }
    { const c64 w7 = mul(w4, w3); v[7] = cmul_tw<INV>(v[7], w7); v[15] = cmul_tw<INV>(v[15], mul(w8, w7)); }
    v[9] = cmul_tw<INV>(v[9], mul(w8, w1));
    v[10] = cmul_tw<INV>(v[10], mul(w8, w2));
    v[11] = cmul_tw<INV>(v[11], mul(w8, w3));
    v[12] = cmul_tw<INV>(v[12], mul(w8, w4));
}

// in : v[m]  = z[t + 256*m] ; out: v[k2] = Z[klow + 256*k2]  (un-normalised, kernel e^{-2 pi i nk/N})
__device__ __forceinline__ void fft_fwd(c64 (&v)[16], const Smem& s, const Lanes& L)
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
    __syncwarp();                                   // plane k0 is private to this half-warp
#pragma unroll
    for (int c = 0; c < 16; c++) v[c] = s.xb[L.base_c + c];
    radix16<false>(v);
}

// in : v[k2] = Z[klow + 256*k2] ; out: v[m] = sum_k Z[k] e^{+2 pi i nk/N}, n = t + 256*m  (no 1/N)
__device__ __forceinline__ void fft_inv(c64 (&v)[16], const Smem& s, const Lanes& L)
{
    radix16<true>(v);
    apply_twiddles<true>(v, s.T2, 16, L.k1c);
#pragma unroll
    for (int c = 0; c < 16; c++) s.xb[L.base_c + c] = v[c];
    __syncwarp();
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
// 1/sqrt(n) to float64 rounding, branch free: float32 MUFU seed y0 (relative error e0 < 2^-22), then ONE third-order
// (Halley) step  y = y0 (1 + e/2 + 3 e^2/8),  e = 1 - n y0^2  -- remaining error O(e0^3) ~ 1e-20, five float64
// instructions.  n = |X|^2 of a spectrum of O(1) signals: far inside the float32 range; n = 0 gives inf * 0 = NaN,
// which the caller's n > 0 select discards.
__device__ __forceinline__ double rsqrt_nr(double n)
{
    const double y0 = (double)rsqrtf((float)n);
    const double e = fma(-(n * y0), y0, 1.0);
    return fma(y0, e * fma(0.375, e, 0.5), y0);
}
// cmplx.Rect(M, cmplx.Phase(X)) (mel/mel.go:98-102): M * X/|X|, Phase(0) = 0 -> (M, 0)
__device__ __forceinline__ c64 subst(c64 X, double M)
{
    const double n = fma(X.x, X.x, X.y * X.y);
    const double r = M * rsqrt_nr(n);
    return (n > 0.0) ? mk(X.x * r, X.y * r) : mk(M, 0.0);
}
// the two real spectra riding one complex transform, up to a common factor 2 (P = Z[N-k])
__device__ __forceinline__ c64 split_a(c64 z, c64 P) { return mk(z.x + P.x, z.y - P.y); }
__device__ __forceinline__ c64 split_b(c64 z, c64 P) { return mk(z.y + P.y, P.x - z.x); }
__device__ __forceinline__ c64 join_lo(c64 ya, c64 yb) { return mk(ya.x - yb.y, ya.y + yb.x); }
__device__ __forceinline__ c64 join_hi(c64 ya, c64 yb) { return mk(ya.x + yb.y, yb.x - ya.y); }

using gomel::prefetch_l2;

// The clips whose float32 tail is re-run in float64 (see k_gl_iter<.., GUARD>), as a dense list built on the device:
// the re-run kernels are launched with a small fixed grid and walk (list slot, tile) items, so a run in which the
// guard selects nothing costs a few microseconds per launch and the host never has to read the count.
struct Selection {
    const int* clips;        // [count] clip indices, ascending
    const int* count;
};
// stat = the float32 iterations' per-clip maximum of M/|X| * rms_frame(M) (float bits), scale = the clip's rms
// magnitude, thr = the threshold in units of the clip scale.  One block; ordered, deterministic.
__global__ void __launch_bounds__(1024) k_guard_select(const unsigned int* __restrict__ stat, const float* __restrict__ scale,
                                                       float thr, int n_clips, int* __restrict__ clips, int* __restrict__ count)
{
    __shared__ int warp_n[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_clips; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const bool sel = c < n_clips && __uint_as_float(stat[c]) > thr * scale[c];
        const unsigned m = __ballot_sync(0xffffffffu, sel);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_n[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int i = 0; i < w; i++) off += warp_n[i];
        if (sel) clips[off + __popc(m & ((1u << lane) - 1u))] = c;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; i++) t += warp_n[i]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

struct GLParams {
    const double* tables;    // T1 | T2 | win  (kTableBytes)
    Tiling tl;
    const double* sig_in;    // [clips][sig_stride]
    double* sig_out;
    const double* hb_in;     // head partials of the previous iteration [clips][hb_tiles][halo], or null
    double* hb_out;
    const double* mags;      // [clips][frames][kMagStride], mag_pos order, pre-scaled by 1/4096 (exact)
    int hb_tiles, tile_lo, tiles_in_launch;
    int edge_mode, edge_tile0, edge_tile1;
    int ext_prev, ext_next, clip0;
    Selection sel;           // clips != null: grid-stride walk over (list slot, tile) items; hb buffers are indexed by slot
};

// One Griffin-Lim iteration (one pass of the loop body of mel.ISTFT, mel/mel.go:85-136) in float64.
// Tile edges exactly as in k_gl_iter: the earlier tile's partial sum of a shared region goes to sig_out
// (its tail), the later tile's to hb_out (its head); every reader adds the two.
// FS = frame slots (Resolut / 256): 16 is the native frame; 8 (Resolut 2048, the mel.NewMel default) runs the frame
// zero-extended through the same 4096-point core, its spectrum on the even core bins (cf. k_gl_iter).
template <int HS, int FS = 16, bool LISTED = false>
__global__ void __launch_bounds__(kThreads, 2) k_gl_iter_f64(const GLParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool listed = LISTED;
    int item = blockIdx.x, n_items = 0;
    if (listed) {                    // re-run of the float32 tail of selected clips: usually nothing to do
        n_items = *p.sel.count * p.tiles_in_launch;
        if (item >= n_items) return;
    }
    const Smem s = carve(smem_raw);
    const Lanes L = make_lanes64();
    {
        double2* d = reinterpret_cast<double2*>(smem_raw);
        const double2* g = reinterpret_cast<const double2*>(p.tables);
        for (int i = L.t; i < kTableBytes / 16; i += kThreads) d[i] = __ldg(g + i);
    }
    constexpr int NR = FS + HS, KEEP = FS - HS, H = 256 * HS, HALO = KEEP * 256;
  for (;;) {                         // one pass unless listed
    int tile, clip, hslot;
    if (p.edge_mode) { clip = 0; tile = blockIdx.x == 0 ? p.edge_tile0 : p.edge_tile1; hslot = 0; }
    else if (listed) { hslot = item / p.tiles_in_launch; tile = p.tile_lo + item % p.tiles_in_launch; clip = p.sel.clips[hslot]; }
    else { tile = p.tile_lo + blockIdx.x % p.tiles_in_launch; clip = p.clip0 + blockIdx.x / p.tiles_in_launch; hslot = clip; }
    const int f0 = tile_begin(p.tl, tile);
    const int nf = tile_begin(p.tl, tile + 1) - f0;
    const int npairs = (nf + 1) >> 1;
    const int tile_len = nf * H;
    const long sbase = (long)f0 * H;
    const double* __restrict__ sin_ = p.sig_in + (long)clip * p.tl.sig_stride + sbase;
    double* __restrict__ sout = p.sig_out + (long)clip * p.tl.sig_stride + sbase;
    const long lim_l = p.tl.sig_len - sbase;
    const int lim = (int)(lim_l < 0x7fffff00L ? lim_l : 0x7fffff00L);
    const bool has_prev = tile > 0 || p.ext_prev, has_next = (tile + 1) < p.tl.n_tiles || p.ext_next;
    const double* __restrict__ hin_own = p.hb_in ? p.hb_in + ((long)hslot * p.hb_tiles + tile) * HALO : nullptr;
    const double* __restrict__ hin_next = hin_own ? hin_own + HALO : nullptr;
    double* __restrict__ hout = p.hb_out + ((long)hslot * p.hb_tiles + tile) * HALO;
    const int t = L.t;

    auto ld = [&](int row) -> double {
        const int o = row + t;
        double x = (o < lim) ? sin_[o] : 0.0;
        if (hin_own) {
            if (has_prev && row < HALO) x += hin_own[o];
            else if (has_next && row >= tile_len) x += hin_next[o - tile_len];
        }
        return x;
    };
    auto plain_rows = [&](int row0, int n) -> bool {
        const int end = row0 + n * 256;
        return end <= lim && (!has_prev || row0 >= HALO) && (!(has_next && hin_own) || end <= tile_len);
    };

    const int idx_lo = mag_pos(L.klow);
    const double* __restrict__ mrow = p.mags + ((long)clip * p.tl.n_frames + f0) * kMagStride;
    __syncthreads();                // tables visible

    for (int pr = 0; pr < npairs; pr++) {
        const int off0 = pr * 2 * H;
        const bool validB = (f0 + 2 * pr + 1) < p.tl.n_frames;
        const double* __restrict__ mA = mrow + (long)(2 * pr) * kMagStride + idx_lo;
        const double* __restrict__ mB = mA + kMagStride;

        c64 v[16];
        {
            double raw[NR];
            if (plain_rows(off0, NR)) {
#pragma unroll
                for (int j = 0; j < NR; j++) raw[j] = sin_[off0 + j * 256 + t];
            } else {
#pragma unroll
                for (int j = 0; j < NR; j++) raw[j] = ld(off0 + j * 256);
            }
#pragma unroll
            for (int m = 0; m < 16; m++) {
                if (m < FS) { const double w = win_at<FS>(s.win, m, t); v[m] = mk(raw[m] * w, raw[m + HS] * w); }
                else v[m] = mk(0.0, 0.0);
            }
        }
        // pull the next pair's new signal rows and both of its magnitude lines towards L2 while this pair computes:
        // two contiguous spans, one 128-byte line per lane (3 instructions per thread instead of 26 -- the LSU pipe is
        // this kernel's co-bottleneck)
        if (pr + 1 < npairs) {
            const int r0 = off0 + 2 * H + KEEP * 256;
            int ns = lim - r0; ns = ns < 2 * H ? ns : 2 * H;                       // doubles
            const char* sp = reinterpret_cast<const char*>(sin_ + r0);
            if (t * 16 < ns) prefetch_l2(sp + t * 128);
            const bool twoM = (f0 + 2 * pr + 3) < p.tl.n_frames;
            const int nm = (twoM ? 2 : 1) * kMagStride;                            // doubles
            const char* mp = reinterpret_cast<const char*>(mrow + (long)(2 * pr + 2) * kMagStride);
            if (t * 16 < nm) prefetch_l2(mp + t * 128);
            if ((t + 256) * 16 < nm) prefetch_l2(mp + (t + 256) * 128);
        }

        fft_fwd(v, s, L);

        // magnitude substitution on both frames at once (see k_gl_iter): each thread handles its LOWER slots
        // (k < 2048) and hands Z'[N-k] to the partner lane that owns bin N-k (its slot 15-j)
        {
            const bool w0 = (t >> 5) == 0;
            if (w0) {
                if (L.special) {
#pragma unroll
                    for (int i = 0; i < 16; i++) s.zsc[i] = v[i];
                }
                __syncwarp();
            }
            // all sixteen target magnitudes are requested before the first one is used: one L2 round trip per pair
            // instead of eight
            double mav[8], mbv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { mav[j] = __ldg(mA + j * 256); mbv[j] = validB ? __ldg(mB + j * 256) : 0.0; }
            c64 nlo[8], nhi[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double ma = mav[j];
                const double mb = mbv[j];
                const c64 P = shfl2(v[15 - j], L.src);
                const c64 z = v[j];
                const c64 ya = subst(split_a(z, P), ma);
                const c64 yb = subst(split_b(z, P), mb);
                nlo[j] = join_lo(ya, yb);
                nhi[j] = shfl2(join_hi(ya, yb), L.src);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j] = nlo[j]; v[15 - j] = nhi[j]; }
            if (w0) {
                // the thread with klow == 0 holds bins 256*s whose partners 256*(16-s) sit in its own slots:
                // lanes 1..9 of warp 0 each substitute one of the nine special pairs {s, 16-s}
                const int j = t - 1;
                if (j >= 0 && j <= 8) {
                    const int jp = (16 - j) & 15, mi = (j == 8) ? 2048 : j * 256;
                    const double* m0 = mrow + (long)(2 * pr) * kMagStride;
                    const double ma = __ldg(m0 + mi);
                    const double mb = validB ? __ldg(m0 + kMagStride + mi) : 0.0;
                    const c64 z = s.zsc[j], P = s.zsc[jp];
                    const c64 ya = subst(split_a(z, P), ma);
                    const c64 yb = subst(split_b(z, P), mb);
                    s.zsc[16 + j] = join_lo(ya, yb);
                    if (jp != j) s.zsc[16 + jp] = join_hi(ya, yb);
                }
                __syncwarp();
                if (L.special) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = s.zsc[16 + i];
                }
            }
        }

        fft_inv(v, s, L);

        // windowed overlap-add through the output buffer.  Row r of the pair's 21-row span receives frame A's
        // sample r (r < 16) and frame B's sample r - HS (r >= HS).  Rows 0..KEEP-1 were first written by the
        // previous pair of this tile (same thread), so they are accumulated; the others are first touches.
        {
            double o[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) o[r] = 0.0;
#pragma unroll
            for (int m = 0; m < FS; m++) {
                const double w = win_at<FS>(s.win, m, t);
                o[m] = v[m].x * w;                        // A first, then B: ascending frame order
            }
#pragma unroll
            for (int m = 0; m < FS; m++) {
                const double w = win_at<FS>(s.win, m, t);
                o[m + HS] = (m + HS < FS) ? fma(v[m].y, w, o[m + HS]) : v[m].y * w;
            }
            const bool fast = off0 + NR * 256 <= lim && (!has_prev || off0 >= HALO);
            if (fast && pr > 0) {
#pragma unroll
                for (int r = 0; r < KEEP; r++) atomicAdd(sout + off0 + r * 256 + t, o[r]);
#pragma unroll
                for (int r = KEEP; r < NR; r++) sout[off0 + r * 256 + t] = o[r];
            } else {
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const int row = off0 + r * 256, ofs = row + t;
                    if (ofs >= lim) continue;
                    double* dst = (has_prev && row < HALO) ? hout + ofs : sout + ofs;
                    if (pr == 0 || r >= KEEP) *dst = o[r];
                    else atomicAdd(dst, o[r]);
                }
            }
        }
    }
    if (!listed) break;
    item += gridDim.x;
    if (item >= n_items) break;
  }
}

// folds the head partials in: sig[s] += hb[s] over the head regions of tiles t_first .. t_end-1 (t_end = n_tiles, or
// n_tiles + 1 when slot n_tiles holds the next rank's head partial over this rank's tail region)
__global__ void k_halo_fix_f64(double* __restrict__ sig, const double* __restrict__ hb, Tiling tl, int hop, int halo,
                               int t_first, int hb_tiles, int t_end, Selection sel = Selection{ nullptr, nullptr })
{
    const int nt = t_end - t_first;
    const int n_items = sel.clips ? *sel.count * nt : (int)gridDim.x;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int slot = item / nt, tile = item % nt + t_first;
        const int clip = sel.clips ? sel.clips[slot] : slot;
        const long s0 = (long)tile_begin(tl, tile) * hop;
        double* __restrict__ d = sig + (long)clip * tl.sig_stride + s0;
        const double* __restrict__ h = hb + ((long)slot * hb_tiles + tile) * halo;
        const long room = tl.sig_len - s0;
        const int n = (int)(room < halo ? (room < 0 ? 0 : room) : halo);
        for (int o = threadIdx.x; o < n; o += blockDim.x) d[o] += h[o];
    }
}

// rows of `len` valid samples, `stride` apart in both buffers: the gap after each row is not touched
__global__ void k_f64_to_f32_rows(const double* __restrict__ in, float* __restrict__ out, long n_rows, long len, long stride)
{
    const long n = n_rows * len;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const long r = i / len, o = r * stride + (i - r * len);
        out[o] = (float)in[o];
    }
}

// the same for the guard's selected clips only: grid = (chunks, list slots)
__global__ void k_f64_to_f32_selected(const double* __restrict__ in, float* __restrict__ out, long len, long stride, Selection sel)
{
    const int count = *sel.count;
    for (int slot = blockIdx.y; slot < count; slot += gridDim.y) {
        const int clip = sel.clips[slot];
        const double* a = in + (long)clip * stride;
        float* b = out + (long)clip * stride;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (long)gridDim.x * blockDim.x) b[i] = (float)a[i];
    }
}

__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, long n)
{
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long step = (long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = (float)in[i];
}

}  // namespace d64
}  // namespace gomel
