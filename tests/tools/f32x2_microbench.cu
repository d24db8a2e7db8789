// f32x2_microbench.cu -- does packed FP32 (FADD2/FFMA2, sm_100+) double per-issue-slot throughput?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_microbench f32x2_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float2 seed, int iters)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, -0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }      // 2 scalar FFMA
            if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);                                              // 1 FFMA2
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                            // 2 scalar FADD
            if (MODE == 3) a[i] = __fadd2_rn(a[i], c);                                                 // 1 FADD2
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(float2* d, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, make_float2(1, 2), iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, make_float2(1, 2), iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2));
    const int iters = 20000;
    const double ops = 148.0 * 8 * 256 * 8 * 2 * (double)iters;      // fp32 lane-operations
    const char* names[4] = { "2x FFMA ", "FFMA2   ", "2x FADD ", "FADD2   " };
    float ms[4] = { run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters) };
    for (int i = 0; i < 4; i++) printf("%s %8.3f ms  %7.2f T lane-ops/s\n", names[i], ms[i], ops / ms[i] / 1e9);
    return 0;
}
