// On-chip roofline denominators for bench.py's per-stage `bound` (profiles/peaks_b200.json), measured on the box:
//   fp32      FFMA and packed FFMA2 (fma.rn.f32x2) flop rate, all SMs, 16 independent chains per thread
//   l1_data   shared-memory wavefront rate: conflict-free LDS.64 / LDS.128 (one wavefront moves 128 B) and STS.64
//   int_alu   LOP3 / IADD3 rate (the ALU pipe that bounds OSD)
//   issue     warp instructions per second with two independent pipes fed (FFMA + LOP3 interleaved)
//   mufu      MUFU.EX2 rate (the XU pipe behind tanhf / division sequences)
// Build + run (gpurun):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/peaks tools/micro/peaks.cu && /tmp/peaks > gpurun_out/peaks_b200.json
// Prints one JSON object.  Best of 5 repetitions each, CUDA events, 8 CTAs x 256 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// three-input XOR as one LOP3; volatile so that ptxas cannot fold the chain algebraically
__device__ __forceinline__ unsigned lop3(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

enum { FFMA, FFMA2, LOP3, MIX, MUFU, LDS64, LDS128, STS64 };

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s, int zero) {
    __shared__ float4 sm[1024];           // 16 KB
    float a[16];
    unsigned b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = threadIdx.x * 2654435761u + i; }
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const float m = 1.0f + s, c = s;
    if (MODE == FFMA) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        }
    } else if (MODE == FFMA2) {
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
        const u64 mm = pk(m, m), cc = pk(c, c);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[2 * i] = __uint_as_float((unsigned)p[i]); a[2 * i + 1] = __uint_as_float((unsigned)(p[i] >> 32)); }
    } else if (MODE == LOP3) {
        const unsigned x = __float_as_uint(s) | 0x9e3779b9u, y = x * 3u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) b[i] = lop3(b[i], x, y);       // one three-input LOP3 each, 16 independent chains
        }
    } else if (MODE == MIX) {
        const unsigned x = __float_as_uint(s) | 0x9e3779b9u, y = x * 3u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], m, c); b[i] = lop3(b[i], x, y); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i + 8] = fmaf(a[i + 8], m, c); b[i + 8] = lop3(b[i + 8], x, y); }
        }
    } else if (MODE == MUFU) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = exp2f(a[i]) * 1e-30f;       // MUFU.EX2 + FMUL
        }
    } else if (MODE == LDS64) {
        const float2* p = reinterpret_cast<const float2*>(sm);
        int idx = threadIdx.x + zero;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float2 v = p[(idx + 32 * i) & 2047]; a[i] += v.x + v.y; }
            idx += zero;
        }
    } else if (MODE == LDS128) {
        int idx = threadIdx.x + zero;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float4 v = sm[(idx + 32 * i) & 1023]; a[i] += v.x + v.w; }
            idx += zero;
        }
    } else if (MODE == STS64) {
        float2* p = reinterpret_cast<float2*>(sm);
        int idx = threadIdx.x + zero;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p[(idx + 32 * i) & 2047] = make_float2(a[i], (float)it);
            idx += zero;
        }
        __syncthreads();
        a[0] += p[threadIdx.x].x;
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i] + __uint_as_float(b[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> static double run(float* out, int nsm, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<nsm * 8, 256>>>(out, iters, 1e-9f, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    float* out; cudaMalloc(&out, (size_t)nsm * 8 * 256 * sizeof(float));
    const int iters = 8000;
    const double thr = (double)nsm * 8 * 256, per = 16.0 * iters;     // thread-level ops of the 16-wide body
    const double ms_ffma = run<FFMA>(out, nsm, iters), ms_ffma2 = run<FFMA2>(out, nsm, iters), ms_lop = run<LOP3>(out, nsm, iters);
    const double ms_mix = run<MIX>(out, nsm, iters), ms_mufu = run<MUFU>(out, nsm, iters);
    const double ms_l64 = run<LDS64>(out, nsm, iters), ms_l128 = run<LDS128>(out, nsm, iters), ms_s64 = run<STS64>(out, nsm, iters);
    const double warps = thr / 32.0;
    printf("{\"gpu\": \"%s\", \"n_sm\": %d, \"clock_khz_max\": %d,\n", p.name, nsm, p.clockRate);
    printf(" \"fp32_tflops_ffma\": %.2f, \"fp32_tflops_ffma2\": %.2f,\n", 2.0 * thr * per / ms_ffma / 1e9, 2.0 * thr * per * 2 / ms_ffma2 / 1e9);
    printf(" \"fma_warp_inst_per_s\": %.4g, \"int_alu_warp_inst_per_s\": %.4g, \"int_alu_thread_gops\": %.1f,\n", warps * per / (ms_ffma * 1e-3),
           warps * per / (ms_lop * 1e-3), thr * per / ms_lop / 1e6);
    printf(" \"issue_warp_inst_per_s\": %.4g, \"mufu_warp_inst_per_s\": %.4g,\n", warps * per * 2 / (ms_mix * 1e-3), warps * per / (ms_mufu * 1e-3));
    // one conflict-free LDS.64 warp request = 2 wavefronts of 128 B, LDS.128 = 4, STS.64 = 2
    printf(" \"l1_wavefronts_per_s_lds64\": %.4g, \"l1_wavefronts_per_s_lds128\": %.4g, \"l1_wavefronts_per_s_sts64\": %.4g,\n",
           warps * per * 2 / (ms_l64 * 1e-3), warps * per * 4 / (ms_l128 * 1e-3), warps * per * 2 / (ms_s64 * 1e-3));
    printf(" \"shared_gbs_lds128\": %.1f,\n", warps * per * 512.0 / (ms_l128 * 1e-3) / 1e9);
    printf(" \"how\": \"tools/micro/peaks.cu: 8 CTAs x 256 threads per SM, 16 independent chains per thread, best of 5, CUDA events\", \"err\": %d}\n",
           (int)cudaGetLastError());
    return 0;
}
