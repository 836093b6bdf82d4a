// Micro-benchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    const float m = 1.0f + s, c = s;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        }
    } else {
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
        const u64 mm = pk(m, m), cc = pk(c, c);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[2 * i] = __uint_as_float((unsigned)p[i]); a[2 * i + 1] = __uint_as_float((unsigned)(p[i] >> 32)); }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 1e-9f); else k<1><<<148 * 8, 256>>>(out, iters, 1e-9f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flops = 2.0 * 16 * iters * 148.0 * 8 * 256;
            printf("%s rep %d: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", rep, ms, flops / ms / 1e9);
        }
    }
    printf("err=%d\n", (int)cudaGetLastError());
    return 0;
}
