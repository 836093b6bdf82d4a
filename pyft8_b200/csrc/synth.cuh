// synth.cuh -- workload generator on the device: GFSK FT8 signals + white Gaussian noise -> int16 cycles.
//
// Restates transmitter.symbols_to_complex_audio (transmitter.py:52-70) in closed form so that a sample can be
// computed independently of its neighbours: the reference accumulates dphi[m] = (2 pi/1920) * sum_i tone_i *
// pulse[m - 1920 i]; with CP = cumulative sum of the 3-symbol Gaussian pulse the phase at m is
// (2 pi/1920) * sum_i tone_i * CP[m - 1920 i], i.e. a prefix sum of completed symbols plus three table lookups.
// The reference's start/end phase patches (transmitter.py:62-63), carrier term, 1920-sample guard trim and
// 240-sample cosine ramps are reproduced.  Mixing: audio = round(noise + sum amp*Im(wf)), int16.
// This is measurement infrastructure (BASELINE configs 2/4/5 need 10^4..10^5 cycles), not part of the decode path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

namespace ft8 {

constexpr int SYNTH_MAX_SIG = 128, SYNTH_TILE = 1024, SYNTH_NT = 256;
constexpr int SPS = 1920, PULSE_LEN = 5760, WF_LEN = 79 * 1920;

// table layout: [0, 5760) pulse, [5760, 11520) cumulative pulse, both as float pairs (hi, lo) of a double
static std::vector<float> synth_pulse_table() {
    std::vector<float> t(4 * PULSE_LEN);
    const double c = M_PI * sqrt(2.0 / log(2.0));
    double acc = 0.0;
    for (int i = 0; i < PULSE_LEN; ++i) {
        const double tt = ((double)i - 1.5 * SPS) / SPS;
        const double p = 0.5 * (erf(c * 2.0 * (tt + 0.5)) - erf(c * 2.0 * (tt - 0.5)));
        acc += p;
        t[2 * i] = (float)p;
        t[2 * i + 1] = (float)(p - (double)(float)p);
        t[2 * PULSE_LEN + 2 * i] = (float)acc;
        t[2 * PULSE_LEN + 2 * i + 1] = (float)(acc - (double)(float)acc);
    }
    return t;
}

__device__ __forceinline__ double tab_d(const float* t, int i) { return (double)t[2 * i] + (double)t[2 * i + 1]; }

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// symbols: [B][n_sig][79]; f_hz, dt_s, amp: [B][n_sig]; audio: [B][180000]
__global__ void __launch_bounds__(SYNTH_NT)
k_synth(const uint8_t* __restrict__ symbols, const float* __restrict__ f_hz, const float* __restrict__ dt_s,
        const float* __restrict__ amp, int n_sig, float noise_sigma, uint64_t seed, const float* __restrict__ pulse,
        int16_t* __restrict__ audio) {
    __shared__ uint16_t pre[SYNTH_MAX_SIG][80];     // prefix sums of tones
    __shared__ uint8_t sym[SYNTH_MAX_SIG][80];
    __shared__ int start[SYNTH_MAX_SIG];
    __shared__ float s_f[SYNTH_MAX_SIG], s_a[SYNTH_MAX_SIG];
    const int cyc = blockIdx.y;
    const int t0 = blockIdx.x * SYNTH_TILE;
    for (int i = threadIdx.x; i < n_sig * 79; i += SYNTH_NT) sym[i / 79][i % 79] = symbols[((size_t)cyc * n_sig) * 79 + i];
    __syncthreads();
    for (int s = threadIdx.x; s < n_sig; s += SYNTH_NT) {
        int acc = 0;
        for (int i = 0; i < 79; ++i) { pre[s][i] = (uint16_t)acc; acc += sym[s][i]; }
        pre[s][79] = (uint16_t)acc;
        start[s] = (int)((0.5 + (double)dt_s[(size_t)cyc * n_sig + s]) * 12000.0);
        s_f[s] = f_hz[(size_t)cyc * n_sig + s];
        s_a[s] = amp[(size_t)cyc * n_sig + s];
    }
    __syncthreads();
    const float* cp = pulse + 2 * PULSE_LEN;
    const double total = tab_d(cp, PULSE_LEN - 1);
    for (int t = t0 + threadIdx.x; t < min(t0 + SYNTH_TILE, 180000); t += SYNTH_NT) {
        // noise: Box-Muller on two hashed uniforms
        const uint64_t r = mix64(seed ^ mix64(((uint64_t)cyc << 32) | (uint32_t)t));
        const float u1 = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777216.0f);
        const float u2 = (float)(uint32_t)((r >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        float v = noise_sigma * sqrtf(-2.0f * logf(u1)) * cs;
        for (int s = 0; s < n_sig; ++s) {
            const int n = t - start[s];                 // sample inside the trimmed waveform
            if (n < 0 || n >= WF_LEN) continue;
            const int m = n + SPS;                       // index in the untrimmed 81-symbol phase array
            const int i0 = m / SPS;                      // symbols i0-2..i0 overlap sample m (those < 79)
            double turns = 0.0;                          // phase / (2 pi)
            const int full = min(max(i0 - 2, 0), 79);    // symbols whose pulse has completely passed
            turns += (double)pre[s][full] * total;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int i = i0 - 2 + d;
                if (i >= 0 && i < 79) {
                    const int k = m - SPS * i;
                    if (k >= 0 && k < PULSE_LEN) turns += (double)sym[s][i] * tab_d(cp, k);
                }
            }
            if (m < 2 * SPS) turns += tab_d(pulse, SPS + m) * (double)sym[s][0];
            if (m >= 79 * SPS) turns += tab_d(pulse, m - 79 * SPS) * (double)sym[s][78];
            turns *= 1.0 / (double)SPS;
            turns += (double)s_f[s] * (double)m / 12000.0;
            const float fr = (float)(turns - floor(turns));
            float sv, cv;
            sincospif(2.0f * fr, &sv, &cv);
            float a = s_a[s];
            if (n < 240) a *= 0.5f * (1.0f - cospif((float)n / 239.0f));
            else if (n >= WF_LEN - 240) a *= 0.5f * (1.0f + cospif((float)(n - (WF_LEN - 240)) / 239.0f));
            v = fmaf(a, sv, v);
        }
        v = rintf(v);
        v = fminf(32767.0f, fmaxf(-32768.0f, v));
        audio[(size_t)cyc * 180000 + t] = (int16_t)v;
    }
}

}  // namespace ft8
