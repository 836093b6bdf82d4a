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

__constant__ int c_costas[7] = {3, 1, 4, 0, 6, 5, 2};
// payload symbol s' (0..57) sits at Costas-framed symbol s' + 7 (first half) or s' + 14 (second half): 7..35, 43..71
// grid value with the reference's ring semantics: row index taken mod 750; rows that are not stored hold 1.0
__device__ __forceinline__ float grid_at(const float* g, int grid_rows, int row, int col) {
    row %= LIVE_ROWS;
    if (row < 0) row += LIVE_ROWS;
    return row < grid_rows ? g[(size_t)row * GRID_COLS + col] : 1.0f;
}

// grid: [B][grid_rows][976].  best_score/best_h0: [B][928].
//
// Shared-memory plan per CTA (58 f0 bins x all 124 h0):
//   P2[r][c]  = g[r][c] + g[r][c+1]             148 x 70 (row stride 73: conflict-free for row-per-lane access)
//   S14[r][c] = sum_{i<7} P2[r][c+2i]           148 x 58 (row stride 61), built with a sliding window along c
// A hypothesis then costs 7 P2 reads (one per Costas symbol, at column f0 + 2*C[k]) plus a running sum
// U = sum_k S14[r+4k][f0] that slides along h0 in steps of 4:  score = sum P2 + (U - sum P2) * (-1/6).
// The 124 h0 hypotheses are split over SY_HS CTAs (62 each, 86 grid rows per CTA): 48 KB of shared memory instead of 81 KB,
// so four CTAs fit per SM and the tile loads of one hide behind the arithmetic of the others.  k_topk merges the halves.
constexpr int SY_PW = 73, SY_SW = 61;
constexpr int SY_HS = 2, SY_HPER = N_H0 / SY_HS, SY_TROWS = SY_HPER + 24;
constexpr int SY_SMEM_BYTES = SY_TROWS * (SY_PW + SY_SW) * 4 + 8 * SY_TF * 4;

__global__ void __launch_bounds__(SY_NT)
k_sync_scores(const float* __restrict__ grid, int grid_rows, int cycle_h0, float* __restrict__ best_score,
              int16_t* __restrict__ best_h0, int h0_lo, int h0_hi) {
    extern __shared__ __align__(16) unsigned char sy_smem_raw[];
    float* P2 = reinterpret_cast<float*>(sy_smem_raw);
    float* S14 = P2 + SY_TROWS * SY_PW;
    float (*red_s)[SY_TF] = reinterpret_cast<float (*)[SY_TF]>(S14 + SY_TROWS * SY_SW);
    int (*red_h)[SY_TF] = reinterpret_cast<int (*)[SY_TF]>(S14 + SY_TROWS * SY_SW + 4 * SY_TF);
    const int cyc = blockIdx.y;
    const int ftile = blockIdx.x % (N_F0 / SY_TF), half = blockIdx.x / (N_F0 / SY_TF);
    const int f_base = F0_LO + ftile * SY_TF;
    const float* g = grid + (size_t)cyc * grid_rows * GRID_COLS;
    const int row0 = cycle_h0 + H0_LO + 148 + half * SY_HPER;    // 111 for the even cycle, first half
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Centre the dB values on a per-cycle constant before summing.  Each row of the Costas kernel sums to zero (2 - 12/6),
    // so the score does not depend on it, but the fp32 partial sums (14-bin boxes, 7-row running sums) shrink from
    // O(5000) to O(100) and their rounding error with them (3e-4 -> 2e-5 on scores of O(100)), which keeps near-tied
    // candidates in the reference's order.  The constant is the mean of 32 fixed samples of grid row 185: identical in
    // every CTA of the cycle.  The float32 kernel value -1/6 is not exact; its residue c * 7 * (2 + 12 * fl(-1/6)) is added back.
    // (every warp derives it itself from the same 32 values: no barrier, and the load overlaps the first tile loads)
    float centre = grid_at(g, grid_rows, cycle_h0 + 185, 64 + 27 * lane);
    bool centre_pending = true;
    // ---- phase A: one warp per row; lanes hold columns lane, lane+32, lane+64 (71 needed), pair sums via shuffles.
    //      Rows are taken four at a time so that 12 independent global loads per lane are in flight.
    constexpr int NW = SY_NT / 32, RB = 4;
    for (int rb = warp * RB; rb < SY_TROWS; rb += NW * RB) {
        float v[RB][3];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r = rb + u;
            int row = (row0 + r) % LIVE_ROWS;
            if (row < 0) row += LIVE_ROWS;
            const bool stored = row < grid_rows && r < SY_TROWS;
            const float* gr = g + (size_t)row * GRID_COLS + f_base;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int c = lane + 32 * q;
                v[u][q] = (c < 71) ? (stored ? __ldg(gr + c) : 1.0f) : 0.0f;
            }
        }
        if (centre_pending) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) centre += __shfl_xor_sync(0xffffffffu, centre, o);
            centre *= (1.0f / 32.0f);
            centre_pending = false;
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int r = rb + u;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float nb = __shfl_down_sync(0xffffffffu, v[u][q], 1);
                const float wrap = __shfl_sync(0xffffffffu, (q < 2) ? v[u][q + 1] : 0.0f, 0);
                if (lane == 31) nb = wrap;
                const int c = lane + 32 * q;
                if (c < 70 && r < SY_TROWS) P2[r * SY_PW + c] = (v[u][q] - centre) + (nb - centre);
            }
        }
    }
    __syncthreads();
    // ---- phase B: S14 along each row with two interleaved sliding windows (even / odd columns)
    for (int task = threadIdx.x; task < 2 * SY_TROWS; task += SY_NT) {
        const int e = task / SY_TROWS, r = task - e * SY_TROWS;
        const float* p = P2 + r * SY_PW;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 7; ++i) s += p[e + 2 * i];
        float* o = S14 + r * SY_SW;
        o[e] = s;
        for (int c = e + 2; c < SY_TF; c += 2) {
            s += p[c + 12] - p[c - 2];
            o[c] = s;
        }
    }
    __syncthreads();
    // ---- phase C: thread = (f0, group of 16 h0); ascending h0, strict '>' from 0 (receiver.py:345-354)
    const int f = threadIdx.x & 63, hg = threadIdx.x >> 6;
    float best = 0.0f;
    int best_h = 0;
    if (f < SY_TF) {
        constexpr int C0 = 6, C1 = 2, C2 = 8, C3 = 0, C4 = 12, C5 = 10, C6 = 4;     // 2 * Costas tone
        const int h_lo = hg * 16;                // relative to this CTA's first hypothesis
        const float* pf = P2 + f;
        const float* sf = S14 + f;
        float U[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float u = 0.f;
            if (h_lo + j < SY_HPER) {
#pragma unroll
                for (int k = 0; k < 7; ++k) u += sf[(h_lo + j + 4 * k) * SY_SW];
            }
            U[j] = u;
        }
        const float c6 = -1.0f / 6.0f;
        const float residue = centre * -4.172325134277344e-07f;      // centre * 7 * (2 + 12 * fl32(-1/6))
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j;
                const int hh = h_lo + i;
                if (hh < SY_HPER) {
                    const float* p = pf + hh * SY_PW;
                    float sp = p[C0] + p[4 * SY_PW + C1];
                    sp += p[8 * SY_PW + C2];
                    sp += p[12 * SY_PW + C3];
                    sp += p[16 * SY_PW + C4];
                    sp += p[20 * SY_PW + C5];
                    sp += p[24 * SY_PW + C6];
                    const float sc = fmaf(c6, U[j] - sp, sp) + residue;      // sum(P2)*1 + (sum(S14) - sum(P2)) * fl(-1/6)
                    const int habs = half * SY_HPER + hh + H0_LO;             // search_time_range (receiver.py:319): [h0_lo, h0_hi)
                    if (sc > best && habs >= h0_lo && habs < h0_hi) { best = sc; best_h = habs; }
                    if (i + 4 < 16 && hh + 4 < SY_HPER) U[j] += sf[(hh + 28) * SY_SW] - sf[hh * SY_SW];
                }
            }
        }
        red_s[hg][f] = best;
        red_h[hg][f] = best_h;
    }
    __syncthreads();
    if (hg == 0 && f < SY_TF) {
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (red_s[q][f] > best) { best = red_s[q][f]; best_h = red_h[q][f]; }
        const int fi = ftile * SY_TF + f;        // halves are stored side by side: [cycle][half][928]
        best_score[((size_t)cyc * SY_HS + half) * N_F0 + fi] = best;
        best_h0[((size_t)cyc * SY_HS + half) * N_F0 + fi] = (int16_t)best_h;
    }
}

// One CTA per cycle: stable descending rank of the f0 bins whose best score > score_min; first max_cands kept.
// The bins above threshold are first compacted in f0 order (ballot scan), so the rank-counting loop runs over them only
// and "ties keep ascending f0" (Python's stable sort, receiver.py:366) becomes "ties keep compacted position".
__global__ void __launch_bounds__(960)
k_topk(const float* __restrict__ best_score, const int16_t* __restrict__ best_h0, float score_min, int max_cands,
       int16_t* __restrict__ cand_f0, int16_t* __restrict__ cand_h0, float* __restrict__ cand_score,
       int32_t* __restrict__ n_cand, int f0_lo, int f0_hi) {
    __shared__ float sc[N_F0];
    __shared__ int wcount[32];
    const int cyc = blockIdx.x, i = threadIdx.x, lane = i & 31, w = i >> 5;
    // merge the h0 halves in ascending-h0 order with the reference's strict '>' (the earlier half wins ties)
    float mine = 0.0f;
    int mine_h = 0;
    if (i < N_F0) {
#pragma unroll
        for (int hf = 0; hf < SY_HS; ++hf) {
            const float s = best_score[((size_t)cyc * SY_HS + hf) * N_F0 + i];
            if (s > mine) { mine = s; mine_h = best_h0[((size_t)cyc * SY_HS + hf) * N_F0 + i]; }
        }
    }
    const bool valid = (i < N_F0) && (mine > score_min) && (F0_LO + i >= f0_lo) && (F0_LO + i < f0_hi);   // search_freq_range
    const uint32_t m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) wcount[w] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
    for (int k = 0; k < 30; ++k) { const int c = wcount[k]; if (k < w) base += c; total += c; }
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (valid) sc[pos] = mine;
    __syncthreads();
    if (valid) {
        int rank = 0;
        for (int j = 0; j < total; ++j) {
            const float s = sc[j];
            rank += (s > mine || (s == mine && j < pos)) ? 1 : 0;
        }
        if (rank < max_cands) {
            const size_t o = (size_t)cyc * max_cands + rank;
            cand_f0[o] = (int16_t)(F0_LO + i);
            cand_h0[o] = (int16_t)mine_h;
            cand_score[o] = mine;
        }
    }
    if (i == 0) n_cand[cyc] = min(total, max_cands);
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
