// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float2 &d, const float2 &a, const float2 &b) {
    unsigned long long dd, aa, bb;
    aa = *reinterpret_cast<const unsigned long long *>(&a);
    bb = *reinterpret_cast<const unsigned long long *>(&b);
    dd = *reinterpret_cast<unsigned long long *>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2 *>(&dd);
}
template <int MODE>
__global__ void k(float *out, float s, int iters) {
    float2 acc[8];
    float2 a = make_float2(s, s * 1.0001f), b = make_float2(1.0f - 1e-7f, 1.0f + 1e-7f);
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { acc[u].x = fmaf(a.x, b.x, acc[u].x); acc[u].y = fmaf(a.y, b.y, acc[u].y); }
            else fma2(acc[u], a, b);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 0.5f, iters); else k<1><<<148 * 8, 256>>>(d, 0.5f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * (double)iters * 16;
            if (rep) printf("mode %d (%s): %.3f ms, %.2f TFMA/s = %.1f FMA/clk/SM at 1.965 GHz\n", mode, mode ? "fma.rn.f32x2" : "scalar fmaf", ms, fma / ms / 1e9, fma / ms / 1e-3 / 148 / 1.965e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
