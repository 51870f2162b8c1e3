// Throughput of legacy mma.sync.m16n8k16 (f16 in, f32 accumulate) on sm_100a: how much Toeplitz-blur
// tensor work fits beside the CUDA-core work of K2.   nvcc -arch=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    unsigned a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        const int iters = 20000;
        k<<<148 * ctas, 256>>>(d, 100);
        cudaEventRecord(e0);
        k<<<148 * ctas, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 148.0 * ctas * 8 /*warps*/ * (double)iters * 8 * 4096.0;
        printf("ctas/SM %d: %.3f ms, %.1f TFLOP/s (mma.sync m16n8k16 f16/f32)\n", ctas, ms, flop / ms * 1e-9);
    }
    return 0;
}
