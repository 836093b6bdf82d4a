// osd.cuh -- ordered-statistics decoding (order 0 + single/double flips), one warp per codeword.
//
// Restates osd_012 (decoders.py:223-272).  The reference runs Gauss-Jordan on a
// 91x174 uint8 matrix with row/column permutation lists.  Here the matrix is held COLUMN-major,
// bit-packed: a column is the 91-bit vector of its entries over the rows (3 x u32), so
//   * the initial matrix G0 = [I | A^T] needs no build step: column c < 91 is the unit vector
//     e_c and column 91+i is generator row i (constant table c_osd.col);
//   * columns are visited in reliability order (|llr| descending, ties by ascending index,
//     NaN last -- numpy's argsort(-|llr|) made stable); sorted position s lives in
//     lane s%32, register slot s/32, so all register indices are compile-time;
//   * a pivot step is: broadcast the column (3 shuffles), pick an unused row with a 1 (any such
//     row gives the same reduced matrix -- the reduced form for a fixed pivot-column set is
//     unique up to row order, and every result below is expressed through pivots, not row
//     numbers), then every lane conditionally XORs the pivot column into its 6 columns;
//   * T = G[:, :91] are the columns with original index < 91 wherever they sit;
//     trial word bit c = parity(u & col_c) with u[row of pivot k] = hard[column of pivot k];
//     flipping u at the row of pivot 90-i adds (row of T) = bit (row) of every col_c.
// CRC-14 is linear, so each of the 1+S vectors carries its 14-bit syndrome and a trial's CRC test is an XOR;
// the 91-bit word of a trial is only assembled when that XOR is zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "codec.cuh"

namespace ft8 {

struct OsdTables {
    uint32_t col[174][3];      // column c of G0 over the 91 rows, LSB-first
};
__constant__ OsdTables c_osd;

constexpr int OSD_MAX_FLIPS = 91;

struct OsdWarpScratch {
    uint8_t piv_row[96];             // pivot k -> row
    uint16_t syn[OSD_MAX_FLIPS + 1];      // CRC syndrome of [0] the order-0 word, [1+i] the row of T that flip i adds
};

// Returns trial index + 1 of the first accepted trial word (0 = none); bits = that word.
// llr: 174 floats in shared or global memory.
// CTA-shared copy of the generator columns, word-major [3][176]: after the sort every lane asks for a different column, and
// a lane-indexed read of __constant__ memory is replayed once per distinct address (up to 32 times per load).
constexpr int OSD_COL_PITCH = 176;
struct OsdCtaTables { uint32_t col[3 * OSD_COL_PITCH]; };
__device__ __forceinline__ void load_osd_tables(OsdCtaTables& t) {
    for (int i = threadIdx.x; i < 174; i += blockDim.x) {
        t.col[i] = c_osd.col[i][0]; t.col[OSD_COL_PITCH + i] = c_osd.col[i][1]; t.col[2 * OSD_COL_PITCH + i] = c_osd.col[i][2];
    }
}

__device__ __forceinline__ int osd_warp(OsdWarpScratch& s, const OsdCtaTables& g, const float* llr, int lane, const LaneSyn& ls, int S, int D, uint32_t* bits) {
    // ---- 1. reliability order: bitonic sort of 256 64-bit keys (8 per lane, slot e = 32*r + lane), descending.
    //      key = (|llr| bits + 1, or 0 for NaN) << 8 | (255 - index): larger |llr| first, ties by ascending index, NaN after
    //      every number, the 82 padding slots (key 0) last.  Slot e ends up holding the e-th column in reliability order.
    unsigned long long key[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = 32 * r + lane;
        unsigned long long k = 0ull;
        if (i < 174) {
            const float a = fabsf(llr[i]);
            const uint32_t u = (a != a) ? 0u : (__float_as_uint(a) + 1u);
            k = ((unsigned long long)u << 8) | (unsigned long long)(255 - i);
        }
        key[r] = k;
    }
#pragma unroll
    for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {                                   // partner slot lives in the same lane
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if ((r & jr) == 0) {
                        const int e = 32 * r;                // lane bits do not matter for k >= 64
                        const bool desc = (e & k) == 0;
                        const unsigned long long x = key[r], y = key[r | jr];
                        const bool sw = desc ? (x < y) : (x > y);
                        key[r] = sw ? y : x;
                        key[r | jr] = sw ? x : y;
                    }
                }
            } else {                                         // partner slot is in lane ^ j, same register
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int e = 32 * r + lane;
                    const bool desc = (e & k) == 0;
                    const unsigned long long x = key[r];
                    const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
                    const bool take_max = (lower == desc);
                    key[r] = (take_max == (x > y)) ? x : y;
                }
            }
        }
    }
    // ---- 2. load columns in sorted order; hard decisions per sorted position
    uint32_t c0[6], c1[6], c2[6];
    uint32_t hard_mask = 0;          // bit r: hard decision of this lane's slot r
    uint32_t orig[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const int sp = lane + 32 * r;
        if (sp < 174) {
            const int c = 255 - (int)(key[r] & 0xFFull);
            orig[r] = c;
            c0[r] = g.col[c]; c1[r] = g.col[OSD_COL_PITCH + c]; c2[r] = g.col[2 * OSD_COL_PITCH + c];
            if (llr[c] > 0.0f) hard_mask |= 1u << r;
        } else {
            orig[r] = 255; c0[r] = c1[r] = c2[r] = 0;
        }
    }
    // ---- 3. Gauss-Jordan over columns in sorted order
    uint32_t used0 = 0, used1 = 0, used2 = 0;     // rows already holding a pivot
    uint32_t u0 = 0, u1 = 0, u2 = 0;              // u[row] = hard decision of that row's pivot column
    int npiv = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        for (int l = 0; l < 32; ++l) {
            if (npiv >= 91 || r * 32 + l >= 174) break;
            const uint32_t v0 = __shfl_sync(0xffffffffu, c0[r], l);
            const uint32_t v1 = __shfl_sync(0xffffffffu, c1[r], l);
            const uint32_t v2 = __shfl_sync(0xffffffffu, c2[r], l);
            const uint32_t f0 = v0 & ~used0, f1 = v1 & ~used1, f2 = v2 & ~used2;
            if ((f0 | f1 | f2) == 0) continue;    // dependent column
            int p;
            if (f0) p = __ffs(f0) - 1; else if (f1) p = 32 + __ffs(f1) - 1; else p = 64 + __ffs(f2) - 1;
            const uint32_t pb = 1u << (p & 31);
            const int pw = p >> 5;
            uint32_t m0 = v0, m1 = v1, m2 = v2;   // pivot column with the pivot row cleared
            if (pw == 0) { m0 &= ~pb; used0 |= pb; } else if (pw == 1) { m1 &= ~pb; used1 |= pb; } else { m2 &= ~pb; used2 |= pb; }
            const uint32_t hb = __shfl_sync(0xffffffffu, hard_mask >> r, l) & 1u;
            if (hb) { if (pw == 0) u0 |= pb; else if (pw == 1) u1 |= pb; else u2 |= pb; }
            if (lane == 0) s.piv_row[npiv] = (uint8_t)p;
            ++npiv;
            // A column that is still a unit vector (an untouched systematic column) eliminates nothing: skip the update.
            if ((m0 | m1 | m2) == 0) continue;
            // pw is warp-uniform: three copies of the update, each testing a fixed word
            if (pw == 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c0[k] & pb) { c0[k] ^= m0; c1[k] ^= m1; c2[k] ^= m2; }
            } else if (pw == 1) {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c1[k] & pb) { c0[k] ^= m0; c1[k] ^= m1; c2[k] ^= m2; }
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c2[k] & pb) { c0[k] ^= m0; c1[k] ^= m1; c2[k] ^= m2; }
            }
        }
    }
    __syncwarp();
    // ---- 4. CRC syndromes of the 1+S vectors over the original columns 0..90 (vector 0 = order-0 word, vector 1+i = row
    //      of T that flip i adds).  Only the 14-bit syndromes are needed to test a trial; the 91-bit words themselves are
    //      built on demand (osd_word) for the rare trials whose syndrome is zero.
    //      Each slot first fetches the syndrome contribution of its original column from the lanes' register table.
    uint32_t slot_syn[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const uint32_t c = orig[k];
        const uint32_t a0 = __shfl_sync(0xffffffffu, ls.s0, c & 31);
        const uint32_t a1 = __shfl_sync(0xffffffffu, ls.s1, c & 31);
        const uint32_t a2 = __shfl_sync(0xffffffffu, ls.s2, c & 31);
        slot_syn[k] = (c < 32) ? a0 : ((c < 64) ? a1 : ((c < 91) ? a2 : 0u));
    }
    const int nvec = 1 + S;
    for (int vi = 0; vi < nvec; ++vi) {
        uint32_t syn = 0;
        if (vi == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if ((__popc(c0[k] & u0) + __popc(c1[k] & u1) + __popc(c2[k] & u2)) & 1) syn ^= slot_syn[k];
        } else {
            const int prow = s.piv_row[90 - (vi - 1)];
            const uint32_t pbit = 1u << (prow & 31);
            if (prow < 32) {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c0[k] & pbit) syn ^= slot_syn[k];
            } else if (prow < 64) {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c1[k] & pbit) syn ^= slot_syn[k];
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) if (c2[k] & pbit) syn ^= slot_syn[k];
            }
        }
        syn = __reduce_xor_sync(0xffffffffu, syn);
        if (lane == 0) s.syn[vi] = (uint16_t)syn;
    }
    __syncwarp();
    // 91-bit word of vector vi (-1: none), XOR-accumulated into w0..w2; every lane gets the result
    auto osd_word = [&](int vi, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
        if (vi < 0) return;
        uint32_t x0 = 0, x1 = 0, x2 = 0;
        const int prow = vi > 0 ? s.piv_row[90 - (vi - 1)] : 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t c = orig[k];
            if (c < 91) {
                uint32_t b;
                if (vi == 0) b = (__popc(c0[k] & u0) + __popc(c1[k] & u1) + __popc(c2[k] & u2)) & 1u;
                else b = (((prow < 32) ? c0[k] : ((prow < 64) ? c1[k] : c2[k])) >> (prow & 31)) & 1u;
                if (b) { const uint32_t bit = 1u << (c & 31); if (c < 32) x0 |= bit; else if (c < 64) x1 |= bit; else x2 |= bit; }
            }
        }
        w0 ^= __reduce_or_sync(0xffffffffu, x0);
        w1 ^= __reduce_or_sync(0xffffffffu, x1);
        w2 ^= __reduce_or_sync(0xffffffffu, x2);
    };
    // ---- 5. enumerate trials in the reference's order; first with payload != 0, CRC ok, valid payload
    //      trial 0: base; 1..S: single flips i = 0..S-1; then pairs (i, j), j < D, j < i, i-major.
    const int nsingle = S;
    int npair = 0;
    for (int i = 0; i < S; ++i) npair += min(i, D);
    const int ntrial = 1 + nsingle + npair;
    const uint32_t bs = s.syn[0];
    for (int base = 0; base < ntrial; base += 32) {
        const int tr = base + lane;
        bool pass = false;
        int fi = -1, fj = -1;
        if (tr < ntrial) {
            if (tr == 0) { }
            else if (tr <= nsingle) fi = tr - 1;
            else {
                int q = tr - 1 - nsingle;      // q-th pair
                int i = 0;
                while (true) { const int cnt = min(i, D); if (q < cnt) break; q -= cnt; ++i; }
                fi = i; fj = q;
            }
            uint32_t syn = bs;
            if (fi >= 0) syn ^= s.syn[1 + fi];
            if (fj >= 0) syn ^= s.syn[1 + fj];
            pass = (syn == 0);
        }
        uint32_t ballot = __ballot_sync(0xffffffffu, pass);
        while (ballot) {
            const int src = __ffs(ballot) - 1;
            ballot &= ballot - 1;
            const int sfi = __shfl_sync(0xffffffffu, fi, src), sfj = __shfl_sync(0xffffffffu, fj, src);
            uint32_t w[3] = {0u, 0u, 0u};
            osd_word(0, w[0], w[1], w[2]);
            osd_word(sfi >= 0 ? 1 + sfi : -1, w[0], w[1], w[2]);
            osd_word(sfj >= 0 ? 1 + sfj : -1, w[0], w[1], w[2]);
            if ((w[0] | w[1] | (w[2] & 0x1FFFu)) != 0 && payload_valid(w)) {
                bits[0] = w[0]; bits[1] = w[1]; bits[2] = w[2];
                return base + src + 1;
            }
        }
    }
    bits[0] = bits[1] = bits[2] = 0u;
    osd_word(0, bits[0], bits[1], bits[2]);           // nothing accepted: report the order-0 word
    return 0;
}

}  // namespace ft8
