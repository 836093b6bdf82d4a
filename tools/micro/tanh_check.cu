// tanh_check.cu -- is csrc/ldpc.cuh's packed tanh_pair bit-identical to libdevice's tanhf?  All 2^32 inputs.
//   nvcc -arch=sm_100a -O3 -I pyft8_b200/csrc -o tanh_check tools/micro/tanh_check.cu && ./tanh_check
// Prints one JSON line: inputs compared, mismatches (NaN payloads compared as NaN == NaN), first mismatching input.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ldpc.cuh"
__global__ void k(unsigned long long* out) {
    const unsigned long long n = 1ull << 31;      // pairs (x, x | sign-flipped partner pattern)
    unsigned long long bad = 0, first = ~0ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t u0 = (uint32_t)i, u1 = (uint32_t)i | 0x80000000u;          // low half: sign 0, high half: sign 1 -> all 2^32 patterns
        float t0, t1;
        ft8::tanh_pair(ft8::pk(__uint_as_float(u0), __uint_as_float(u1)), t0, t1);
        const float w0 = tanhf(__uint_as_float(u0)), w1 = tanhf(__uint_as_float(u1));
        const bool ok0 = (__float_as_uint(t0) == __float_as_uint(w0)) || (t0 != t0 && w0 != w0);
        const bool ok1 = (__float_as_uint(t1) == __float_as_uint(w1)) || (t1 != t1 && w1 != w1);
        if (!ok0) { ++bad; if (u0 < first) first = u0; }
        if (!ok1) { ++bad; if (u1 < first) first = u1; }
    }
    atomicAdd(&out[0], bad);
    if (first != ~0ull) atomicMin(&out[1], first);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 16);
    unsigned long long h[2] = {0, ~0ull}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
    k<<<148 * 16, 256>>>(d);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("{\"inputs\": 4294967296, \"mismatches\": %llu, \"first_mismatch_bits\": \"%llx\", \"err\": %d}\n", h[0], h[1], (int)cudaGetLastError());
    return 0;
}
