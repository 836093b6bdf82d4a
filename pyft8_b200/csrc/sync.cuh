// sync.cuh -- S2 Costas search, candidate top-K, and L0 (payload gather + max-log LLRs).
//
// Restates Receiver.search (receiver.py:338-367) and Candidate._dB_to_llr
// (receiver.py:208-222).
//   score(f0,h0) = sum_{k<7} sum_{j<14} grid[cycle_h0 + h0 + 148 + 4k, f0 + j] * csync[k, j]
// with csync = 1 on the two bins of Costas tone C[k] and -1/6 elsewhere: only the MIDDLE Costas
// block is scored.  Since csync is 1 / -1/6, a row's contribution is P2 + (S14 - P2) * (-1/6) with
// P2 the 2-bin sum at the Costas tone and S14 the 14-bin box sum; the kernel stages a tile of
// the 148 grid rows involved in shared memory once, builds the box sums there, and every thread then
// walks h0 in ascending order (strict '>' from 0: first maximum wins, at most one candidate per f0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ft8 {

constexpr int N_F0 = 928, F0_LO = 32, N_H0 = 124, H0_LO = -37;
constexpr int SY_TF = 58;                 // f0 bins per CTA (928 = 16 * 58)
constexpr int SY_W = SY_TF + 14;          // staged columns (13 extra + 1 pad)
constexpr int SY_ROWS = 148;              // grid rows h0+148+4k for h0 in [-37,87), k < 7 -> 111..258
constexpr int SY_NT = 256;
constexpr int LIVE_ROWS = 750;
constexpr int SY_SMEM_BYTES = SY_ROWS * (SY_W + SY_TF + 2) * 4 + 8 * SY_TF * 4;

__constant__ int c_costas[7] = {3, 1, 4, 0, 6, 5, 2};
__constant__ uint8_t c_payload_sym[58] = {7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27,
                                           28, 29, 30, 31, 32, 33, 34, 35, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54,
                                           55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67, 68, 69, 70, 71};

// grid value with the reference's ring semantics: row index taken mod 750; rows that are not stored hold 1.0
__device__ __forceinline__ float grid_at(const float* g, int grid_rows, int row, int col) {
    row %= LIVE_ROWS;
    if (row < 0) row += LIVE_ROWS;
    return row < grid_rows ? g[(size_t)row * GRID_COLS + col] : 1.0f;
}

// grid: [B][grid_rows][976].  best_score/best_h0: [B][928].
__global__ void __launch_bounds__(SY_NT)
k_sync_scores(const float* __restrict__ grid, int grid_rows, int cycle_h0, float* __restrict__ best_score,
              int16_t* __restrict__ best_h0) {
    extern __shared__ __align__(16) unsigned char sy_smem_raw[];
    float (*tile)[SY_W] = reinterpret_cast<float (*)[SY_W]>(sy_smem_raw);                                   // dB values
    float (*box)[SY_TF + 2] = reinterpret_cast<float (*)[SY_TF + 2]>(sy_smem_raw + SY_ROWS * SY_W * 4);     // 14-bin box sums
    float (*red_s)[SY_TF] = reinterpret_cast<float (*)[SY_TF]>(sy_smem_raw + SY_ROWS * (SY_W + SY_TF + 2) * 4);
    int (*red_h)[SY_TF] = reinterpret_cast<int (*)[SY_TF]>(sy_smem_raw + SY_ROWS * (SY_W + SY_TF + 2) * 4 + 4 * SY_TF * 4);
    const int cyc = blockIdx.y;
    const int f_base = F0_LO + blockIdx.x * SY_TF;
    const float* g = grid + (size_t)cyc * grid_rows * GRID_COLS;
    const int row0 = cycle_h0 + H0_LO + 148;    // 111 for the even cycle
    for (int i = threadIdx.x; i < SY_ROWS * SY_W; i += SY_NT) {
        const int r = i / SY_W, c = i - r * SY_W;
        const int col = f_base + c;
        tile[r][c] = (col < GRID_COLS) ? grid_at(g, grid_rows, row0 + r, col) : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SY_ROWS * SY_TF; i += SY_NT) {
        const int r = i / SY_TF, c = i - r * SY_TF;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 14; ++j) s += tile[r][c + j];
        box[r][c] = s;
    }
    __syncthreads();
    const int f = threadIdx.x & 63, hg = threadIdx.x >> 6;    // 4 groups of 31 h0 values
    float best = 0.0f;
    int best_h = 0;
    if (f < SY_TF) {
        const float c6 = -1.0f / 6.0f;
        for (int hh = hg * 31; hh < hg * 31 + 31; ++hh) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int r = hh + 4 * k;
                const int c = f + 2 * c_costas[k];
                const float p2 = tile[r][c] + tile[r][c + 1];
                s += fmaf(box[r][f] - p2, c6, p2);
            }
            if (s > best) { best = s; best_h = hh + H0_LO; }
        }
        red_s[hg][f] = best;
        red_h[hg][f] = best_h;
    }
    __syncthreads();
    if (hg == 0 && f < SY_TF) {
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (red_s[q][f] > best) { best = red_s[q][f]; best_h = red_h[q][f]; }
        const int fi = blockIdx.x * SY_TF + f;
        best_score[(size_t)cyc * N_F0 + fi] = best;
        best_h0[(size_t)cyc * N_F0 + fi] = (int16_t)best_h;
    }
}

// One CTA per cycle: stable descending rank of the f0 bins whose best score > score_min; first max_cands kept.
__global__ void __launch_bounds__(960)
k_topk(const float* __restrict__ best_score, const int16_t* __restrict__ best_h0, float score_min, int max_cands,
       int16_t* __restrict__ cand_f0, int16_t* __restrict__ cand_h0, float* __restrict__ cand_score,
       int32_t* __restrict__ n_cand) {
    __shared__ float sc[N_F0];
    __shared__ int n_valid;
    const int cyc = blockIdx.x, i = threadIdx.x;
    if (i == 0) n_valid = 0;
    float mine = 0.f;
    if (i < N_F0) {
        mine = best_score[(size_t)cyc * N_F0 + i];
        sc[i] = (mine > score_min) ? mine : -1.0f;     // scores are > 0 when valid
    }
    __syncthreads();
    if (i < N_F0 && mine > score_min) {
        int rank = 0;
        for (int j = 0; j < N_F0; ++j) {
            const float s = sc[j];
            rank += (s > mine || (s == mine && j < i)) ? 1 : 0;
        }
        atomicAdd(&n_valid, 1);
        if (rank < max_cands) {
            const size_t o = (size_t)cyc * max_cands + rank;
            cand_f0[o] = (int16_t)(F0_LO + i);
            cand_h0[o] = best_h0[(size_t)cyc * N_F0 + i];
            cand_score[o] = mine;
        }
    }
    __syncthreads();
    if (i == 0) n_cand[cyc] = min(n_valid, max_cands);
}

// Max-log LLRs of one 58x8 payload held by a warp: lanes 0..28 own symbols lane and lane+29.
// p[2][8] = this lane's two symbols (dB).  Writes llr[174] (scaled), returns sd and snr.  receiver.py:208-222
__device__ __forceinline__ void llr_from_payload_warp(const float (*p)[8], int lane, float* llr_out, float& sd_out, int& snr_out) {
    float l[2][3];
    float mx = -INFINITY, mn = INFINITY, sum = 0.f, sq = 0.f;
    if (lane < 29) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float* v = p[q];
            l[q][0] = fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])) - fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
            l[q][1] = fmaxf(fmaxf(v[2], v[3]), fmaxf(v[4], v[7])) - fmaxf(fmaxf(v[0], v[1]), fmaxf(v[5], v[6]));
            l[q][2] = fmaxf(fmaxf(v[1], v[2]), fmaxf(v[6], v[7])) - fmaxf(fmaxf(v[0], v[3]), fmaxf(v[4], v[5]));
#pragma unroll
            for (int t = 0; t < 8; ++t) { mx = fmaxf(mx, v[t]); mn = fminf(mn, v[t]); }
#pragma unroll
            for (int b = 0; b < 3; ++b) { sum += l[q][b]; sq = fmaf(l[q][b], l[q][b], sq); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = sum / 174.0f;
    const float var = sq / 174.0f - mean * mean;
    const float sd = sqrtf(var);
    sd_out = sd;
    const float d = mx - mn - 58.0f;
    int s = (int)d;                      // truncation toward zero, like int()
    snr_out = max(-24, min(24, s));
    if (lane < 29) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int b = 0; b < 3; ++b) llr_out[3 * (lane + 29 * q) + b] = __fdiv_rn(__fmul_rn(2.83f, l[q][b]), sd);
    }
}

}  // namespace ft8
