// fft4096.cuh -- complex FFT-4096 for one CTA of 256 threads, sm_100a.
//
// Replaces, for the gomel hot path, the per-frame calls into go-dsp's radix-2 FFT
// (fft.FFTReal at mel/mel.go:95, fft.IFFT at mel/mel.go:116 and phase/phase.go:103, and the FFT
// inside gossp STFT.STFT at mel/mel.go:52, phase/phase.go:47).
//
// Design (see DESIGN.md "FFT core"):
//  * TWO real frames ride one complex transform: z = w*(xA + i*xB).  4096 = 16*16*16, so a
//    transform is three radix-16 register butterflies per thread with two shared-memory
//    exchanges between them.  Thread t holds points n = t + 256*m, m = 0..15.
//  * forward = decimation in frequency (natural in, digit-reversed out), inverse = decimation in
//    time (digit-reversed in, natural out): the spectrum never has to be re-ordered, the
//    point-wise stage between them works on the digit-reversed layout
//        thread (k0,k1), slot k2  <->  bin k = k0 + 16*k1 + 256*k2.
//  * the exchange buffer is used IN PLACE: every thread always writes exactly the logical cells
//    it read last, so one barrier per exchange (RAW only) is enough -- and the exchange between
//    stages 2 and 3 stays inside a half-warp, so it needs only __syncwarp: two CTA barriers per
//    transform pair instead of four.  Physical address of
//    logical cell L is L + 2*(L>>4) (two pads per 16), which makes all three access patterns
//    bank-conflict free with compile-time slot offsets -- pattern (c) with 128-bit accesses.
//  * bins k and 4096-k (needed together to split the two real spectra) are mapped to lanes of
//    the same warp, so that split is done with warp shuffles, not another smem pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gomel {

constexpr int kN = 4096;          // Resolut
constexpr int kThreads = 256;     // threads per CTA = points / 16
constexpr int kRow = 18;                                   // float2 cells per 16-cell row (2 pad)
constexpr int kPlane = 16 * kRow;                          // cells per k0 plane (288)
constexpr int kXchgCells = 16 * kPlane;                    // padded float2 cells
constexpr int kXchgBytes = kXchgCells * 8;                 // 36,864 B
// External twiddles are powers of one per-thread root: W4096^(t*k0) = (W4096^t)^k0, W256^(n0*k1) = (W256^n0)^k1.
// Table modes (per stage): 4 = powers 1,2,4,8 stored, eleven formed by <=3 packed complex multiplies;
// 8 = powers 1..8 stored, seven formed by one multiply (w^(8+i) = w^8 * w^i); 16 = full table.
// Fewer stored powers trade shared-memory traffic (and capacity) for packed FP32 instructions.
#ifndef GOMEL_T1_MODE
#define GOMEL_T1_MODE 4
#endif
#ifndef GOMEL_T2_MODE
#define GOMEL_T2_MODE 4
#endif
constexpr int kT1Mode = GOMEL_T1_MODE, kT2Mode = GOMEL_T2_MODE;
constexpr int kT1Cells = kT1Mode * 256;                    // [mode/2][t] float4
constexpr int kT2Cells = kT2Mode * 16;
constexpr int kWinCells = 2048;                            // first half of the symmetric Hann window, [m][t], m < 8
constexpr int kTableBytes = kT1Cells * 8 + kT2Cells * 8 + kWinCells * 4;   // 16,896 B in the default 4/4 mode
constexpr int kSmemBytes = kTableBytes + kXchgBytes;       // 53,760 B

struct Smem {
    float2* T1;     // stage-1 roots W4096^t: [kT1Mode/2][256] float4 (two powers per cell)
    float2* T2;     // stage-2 roots W256^n0: [kT2Mode/2][16] float4
    float*  win;    // first half of the Hann window, [8][256]
    float2* xb;     // exchange buffer, kXchgCells
};

__device__ __forceinline__ Smem carve_smem(unsigned char* base)
{
    Smem s;
    s.T1 = reinterpret_cast<float2*>(base);
    s.T2 = s.T1 + kT1Cells;
    s.win = reinterpret_cast<float*>(s.T2 + kT2Cells);
    s.xb = reinterpret_cast<float2*>(s.win + kWinCells);
    return s;
}

// Hann coefficient of sample n = t + 256*m of an FS*256-sample frame.  The window is symmetric
// (w[n] = w[N-1-n]), only its first half is kept in shared memory; both access patterns are unit-stride
// across a warp.
template <int FS = 16>
__device__ __forceinline__ float win_at(const float* win, int m, int t)
{
    return (m < FS / 2) ? win[m * 256 + t] : win[(FS - 1 - m) * 256 + (255 - t)];
}

// global table blob layout == smem table layout (T1 | T2 | win), kTableBytes long
__device__ __forceinline__ void load_tables(const Smem& s, const float4* __restrict__ g, int t)
{
    float4* d = reinterpret_cast<float4*>(s.T1);
#pragma unroll 4
    for (int i = t; i < kTableBytes / 16; i += kThreads) d[i] = __ldg(g + i);
}

// per-thread indices for the three exchange patterns and the spectrum layout
struct Lanes {
    int t;         // thread id
    int base_a;    // t + 2*(t>>4); slot k0: + k0*kPlane
    int base_b;    // pattern (b): thread (k0b,n0b), slot r: base_b + r*kRow.  k0b == k0c: the 16 threads that share a
                   // k0 plane form the same half-warp in stages 2 and 3, so the exchange between those stages is warp-local
    int n0b;       // n0 digit owned in stage 2 (lane & 15)
    int base_c;    // pattern (c): thread (k0c,k1c), slot n0: base_c + n0   (even -> 16-byte aligned)
    int k0c, k1c;  // spectrum digits owned after the forward transform
    int klow;      // k0c + 16*k1c : bins k = klow + 256*k2
    int src;       // lane holding bins 4096-k (partner)
    bool special;  // klow == 0: partner of slot k2 is own slot (16-k2)&15
};

__device__ __forceinline__ Lanes make_lanes()
{
    Lanes L;
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    L.t = t;
    L.base_a = t + 2 * (t >> 4);
    if (l < 16) { L.k0c = w; L.k1c = l; }
    else        { L.k0c = (w == 0) ? 8 : 16 - w; L.k1c = 31 - l; }
    L.n0b = l & 15;
    L.base_b = L.k0c * kPlane + L.n0b;
    L.base_c = L.k0c * kPlane + L.k1c * kRow;
    L.klow = L.k0c + 16 * L.k1c;
    if (w == 0) L.src = (l < 16) ? ((16 - l) & 15) : (47 - l);
    else        L.src = l ^ 16;
    L.special = (t == 0);
    return L;
}

// ---------------------------------------------------------------- complex helpers, packed FP32
// A complex value is a float2 in an aligned register pair; sm_100 executes FADD2 / FMUL2 / FFMA2 on
// such pairs with free operand swizzle (.LO_HI), per-lane sign (.NP/.PN), negation and scalar broadcast,
// so a complex add, "+- i*b", "a + s*b" and half of a complex multiply are ONE issued instruction each.
// (Same FP32 pipe throughput as two scalar instructions, half the issue slots -- the kernel is issue bound.)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cadd_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }   // a + i*b
__device__ __forceinline__ float2 csub_i(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }   // a - i*b
__device__ __forceinline__ float2 cfma_s(float2 b, float s, float2 a) { return __ffma2_rn(b, make_float2(s, s), a); }   // a + s*b
// a * (wr + i*wi)
__device__ __forceinline__ float2 cmul(float2 a, float wr, float wi)
{
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-wi, wi), __fmul2_rn(a, make_float2(wr, wr)));
}
template <bool INV>
__device__ __forceinline__ float2 cmul_tw(float2 a, float2 w)   // forward: a*w ; inverse: a*conj(w)
{
    return INV ? cmul(a, w.x, -w.y) : cmul(a, w.x, w.y);
}

// 4-point DFT, forward kernel e^{-2 pi i/4} (INV: conjugate)
template <bool INV>
__device__ __forceinline__ void radix4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    if (!INV) { a1 = csub_i(t1, t3); a3 = cadd_i(t1, t3); }
    else      { a1 = cadd_i(t1, t3); a3 = csub_i(t1, t3); }
}

// multiply by W16^e (forward) or its conjugate (INV); e in {1,3,9}
template <bool INV, int E>
__device__ __forceinline__ float2 mul_w16(float2 a)
{
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
    if (E == 1) return INV ? cmul(a, c1, s1) : cmul(a, c1, -s1);
    if (E == 3) return INV ? cmul(a, s1, c1) : cmul(a, s1, -c1);
    /* E == 9 */ return INV ? cmul(a, -c1, -s1) : cmul(a, -c1, s1);
}

// 4-point DFT over (b0, h*u1, W16^4*r2, h*u3): u1, u3 carry a pending factor h = sqrt(1/2) and r2 the
// pending rotation W16^4 = -i (forward) / +i (inverse); both are folded into the packed adds
template <bool INV>
__device__ __forceinline__ void radix4_h13(float2& a0, float2& u1, float2& r2, float2& u3)
{
    constexpr float h = 0.70710678118654752f;
    const float2 t0 = INV ? cadd_i(a0, r2) : csub_i(a0, r2);
    const float2 t1 = INV ? csub_i(a0, r2) : cadd_i(a0, r2);
    const float2 s2 = cadd(u1, u3), s3 = csub(u1, u3);
    a0 = cfma_s(s2, h, t0);
    r2 = cfma_s(s2, -h, t0);
    const float2 s3s = make_float2(s3.y, s3.x);
    if (!INV) { u1 = __ffma2_rn(s3s, make_float2(h, -h), t1); u3 = __ffma2_rn(s3s, make_float2(-h, h), t1); }   // t1 -+ i*h*s3
    else      { u1 = __ffma2_rn(s3s, make_float2(-h, h), t1); u3 = __ffma2_rn(s3s, make_float2(h, -h), t1); }
}
// 4-point DFT whose input b2 = h*u2 carries a pending factor h
template <bool INV>
__device__ __forceinline__ void radix4_h2(float2& a0, float2& a1, float2& u2, float2& a3)
{
    constexpr float h = 0.70710678118654752f;
    const float2 t0 = cfma_s(u2, h, a0), t1 = cfma_s(u2, -h, a0);
    const float2 t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    u2 = csub(t0, t2);
    if (!INV) { a1 = csub_i(t1, t3); a3 = cadd_i(t1, t3); }
    else      { a1 = cadd_i(t1, t3); a3 = csub_i(t1, t3); }
}

// 16-point DFT in registers, natural order in and out:  v[k] <- sum_m v[m] W16^{mk}
template <bool INV>
__device__ __forceinline__ void radix16(float2 (&v)[16])
{
    // step 1: over m1 for each m0 (elements m0, m0+4, m0+8, m0+12) -> B[m0][ka] at v[m0+4ka]
#pragma unroll
    for (int m0 = 0; m0 < 4; m0++) radix4<INV>(v[m0], v[m0 + 4], v[m0 + 8], v[m0 + 12]);
    // step 2: inner twiddles W16^{m0*ka}.  W16^2 = h(1 -+ i) and W16^6 = h(-1 -+ i) (upper sign forward)
    // keep their factor h pending, W16^4 = -+i stays pending entirely (folded into step 3)
    v[5]  = mul_w16<INV, 1>(v[5]);   v[13] = mul_w16<INV, 3>(v[13]);
    v[7]  = mul_w16<INV, 3>(v[7]);   v[15] = mul_w16<INV, 9>(v[15]);
    auto w2 = [](float2 a) { return INV ? cadd_i(a, a) : csub_i(a, a); };                           // a(1 -+ i)
    auto w6 = [](float2 a) {                                                                        // a(-1 -+ i)
        return INV ? __fadd2_rn(make_float2(-a.x, -a.y), make_float2(-a.y, a.x))
                   : __fadd2_rn(make_float2(-a.x, -a.y), make_float2(a.y, -a.x));
    };
    v[6] = w2(v[6]); v[9] = w2(v[9]); v[11] = w6(v[11]); v[14] = w6(v[14]);
    // step 3: over m0 for each ka (elements 4ka .. 4ka+3) -> X[ka + 4kb]
    float2 o[16];
    {
        float2 b0 = v[0], b1 = v[1], b2 = v[2], b3 = v[3];
        radix4<INV>(b0, b1, b2, b3);
        o[0] = b0; o[4] = b1; o[8] = b2; o[12] = b3;
    }
    {
        float2 b0 = v[4], b1 = v[5], b2 = v[6], b3 = v[7];
        radix4_h2<INV>(b0, b1, b2, b3);
        o[1] = b0; o[5] = b1; o[9] = b2; o[13] = b3;
    }
    {
        float2 b0 = v[8], b1 = v[9], b2 = v[10], b3 = v[11];
        radix4_h13<INV>(b0, b1, b2, b3);
        o[2] = b0; o[6] = b1; o[10] = b2; o[14] = b3;
    }
    {
        float2 b0 = v[12], b1 = v[13], b2 = v[14], b3 = v[15];
        radix4_h2<INV>(b0, b1, b2, b3);
        o[3] = b0; o[7] = b1; o[11] = b2; o[15] = b3;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = o[i];
}

// external twiddles v[k] *= w^k (forward) or conj(w)^k (inverse), w = this lane's root
template <bool INV, int MODE>
__device__ __forceinline__ void apply_twiddles(float2 (&v)[16], const float2* T, int rowlen, int lane)
{
    const float4* T4 = reinterpret_cast<const float4*>(T);
    auto mul = [](float2 x, float2 y) { return cmul(x, y.x, y.y); };
    if (MODE == 4) {
        const float4 a = T4[lane], b = T4[rowlen + lane];          // (w^1, w^2), (w^4, w^8)
        const float2 w1 = make_float2(a.x, a.y), w2 = make_float2(a.z, a.w), w4 = make_float2(b.x, b.y), w8 = make_float2(b.z, b.w);
        v[1] = cmul_tw<INV>(v[1], w1); v[2] = cmul_tw<INV>(v[2], w2); v[4] = cmul_tw<INV>(v[4], w4); v[8] = cmul_tw<INV>(v[8], w8);
        const float2 w3 = mul(w2, w1);
        v[3] = cmul_tw<INV>(v[3], w3);
        { const float2 w5 = mul(w4, w1); v[5] = cmul_tw<INV>(v[5], w5); v[13] = cmul_tw<INV>(v[13], mul(w8, w5)); }
        { const float2 w6 = mul(w4, w2); v[6] = cmul_tw<INV>(v[6], w6); v[14] = cmul_tw<INV>(v[14], mul(w8, w6)); }
        { const float2 w7 = mul(w4, w3); v[7] = cmul_tw<INV>(v[7], w7); v[15] = cmul_tw<INV>(v[15], mul(w8, w7)); }
        v[9] = cmul_tw<INV>(v[9], mul(w8, w1));
        v[10] = cmul_tw<INV>(v[10], mul(w8, w2));
        v[11] = cmul_tw<INV>(v[11], mul(w8, w3));
        v[12] = cmul_tw<INV>(v[12], mul(w8, w4));
    } else if (MODE == 8) {
        // rows q = 0..3 hold (w^(2q+1), w^(2q+2)): w^1..w^8
        float2 w[9];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 a = T4[q * rowlen + lane];
            w[2 * q + 1] = make_float2(a.x, a.y); w[2 * q + 2] = make_float2(a.z, a.w);
        }
#pragma unroll
        for (int k = 1; k <= 8; k++) v[k] = cmul_tw<INV>(v[k], w[k]);
#pragma unroll
        for (int k = 1; k <= 7; k++) v[8 + k] = cmul_tw<INV>(v[8 + k], mul(w[8], w[k]));
    } else {
        // T[(k>>1)][lane][k&1]: two twiddles per 128-bit shared-memory load
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 w = T4[q * rowlen + lane];
            if (q > 0) v[2 * q] = cmul_tw<INV>(v[2 * q], make_float2(w.x, w.y));
            v[2 * q + 1] = cmul_tw<INV>(v[2 * q + 1], make_float2(w.z, w.w));
        }
    }
}

// pattern (c): 16 consecutive cells per thread, 16-byte aligned -> 128-bit accesses
__device__ __forceinline__ void load_c(float2 (&v)[16], const float2* xb, int base_c)
{
    const float4* p = reinterpret_cast<const float4*>(xb + base_c);
#pragma unroll
    for (int q = 0; q < 8; q++) { const float4 w = p[q]; v[2 * q] = make_float2(w.x, w.y); v[2 * q + 1] = make_float2(w.z, w.w); }
}
__device__ __forceinline__ void store_c(const float2 (&v)[16], float2* xb, int base_c)
{
    // 64-bit stores, spelled in PTX so that they are not re-merged: a 128-bit store needs four consecutive
    // registers and costs four MOVs to assemble them out of two packed-FP32 register pairs
    const unsigned a = (unsigned)__cvta_generic_to_shared(xb + base_c);
#pragma unroll
    for (int c = 0; c < 16; c++)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a + 8u * c), "f"(v[c].x), "f"(v[c].y));
}

// ---------------------------------------------------------------- forward transform (DIF)
// in : v[m]  = z[t + 256*m]
// out: v[k2] = Z[klow + 256*k2]        (un-normalised, kernel e^{-2 pi i nk/N})
__device__ __forceinline__ void fft4096_fwd(float2 (&v)[16], const Smem& s, const Lanes& L)
{
    radix16<false>(v);                                               // n2 -> k0
    apply_twiddles<false, kT1Mode>(v, s.T1, 256, L.t);                        // W4096^(t*k0)
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) s.xb[k0 * kPlane + L.base_a] = v[k0];    // pattern (a)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = s.xb[L.base_b + r * kRow];           // pattern (b)
    radix16<false>(v);                                               // n1 -> k1
    apply_twiddles<false, kT2Mode>(v, s.T2, 16, L.n0b);                       // W256^(n0*k1)
#pragma unroll
    for (int r = 0; r < 16; r++) s.xb[L.base_b + r * kRow] = v[r];           // pattern (b), in place
    __syncwarp();                                                    // plane k0 is private to this half-warp
    load_c(v, s.xb, L.base_c);                                               // pattern (c)
    radix16<false>(v);                                               // n0 -> k2
}

// ---------------------------------------------------------------- inverse transform (DIT)
// in : v[k2] = Z[klow + 256*k2]
// out: v[m]  = sum_k Z[k] e^{+2 pi i nk/N},  n = t + 256*m   (caller folds in 1/N)
__device__ __forceinline__ void fft4096_inv(float2 (&v)[16], const Smem& s, const Lanes& L)
{
    radix16<true>(v);                                                // k2 -> n0
    apply_twiddles<true, kT2Mode>(v, s.T2, 16, L.k1c);                        // conj W256^(n0*k1), table is symmetric
    store_c(v, s.xb, L.base_c);                                              // pattern (c), in place
    __syncwarp();                                                    // plane k0 is private to this half-warp
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = s.xb[L.base_b + r * kRow];           // pattern (b)
    radix16<true>(v);                                                // k1 -> n1
#pragma unroll
    for (int r = 0; r < 16; r++) s.xb[L.base_b + r * kRow] = v[r];           // pattern (b), in place
    __syncthreads();
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) v[k0] = s.xb[k0 * kPlane + L.base_a];    // pattern (a)
    apply_twiddles<true, kT1Mode>(v, s.T1, 256, L.t);                         // conj W4096^(t*k0)
    radix16<true>(v);                                                // k0 -> n2
}

// ---------------------------------------------------------------- partner fetch
// P = Z[4096 - k] for the bins this thread holds (k = klow + 256*k2): generic lanes take slot 15-k2 of
// lane L.src with shfl2().  The shuffles are unconditional (every lane of every warp executes them); only
// the single special thread (klow == 0, whose partner is its own slot (16-k2)&15) recomputes its bins.
__device__ __forceinline__ float2 shfl2(float2 a, int src)
{
    return make_float2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
}
}  // namespace gomel
