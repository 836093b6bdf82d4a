// codec.cuh -- CRC-14 and the payload-validity predicate on the device.
//
// K1: crc_unpack91 (decoders.py:117-131) succeeds iff the 77-bit payload is non-zero, the CRC-14
// (poly 0x2757 over the payload zero-extended to 82 bits) equals bits 77..90, and unpack()
// (decoders.py:16-115) returns a message.  Text formatting stays in Python; what the LDPC / OSD
// schedules need from unpack() is only accept/reject, restated here as bit logic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ft8 {

// Device-resident constant tables (filled once per process by upload_tables(); identical for every handle).
struct CodecTables {
    uint16_t crc_syn[96];      // syndrome contribution of codeword bit j < 91 (CRC is linear over GF(2))
    uint32_t prefix2[41];      // 36x36 bitmap of accepted two-character callsign prefixes
};
__constant__ CodecTables c_codec;

// bit j of a 91-bit word packed LSB-first in 3 x u32
__device__ __forceinline__ uint32_t bit91(const uint32_t* w, int j) { return (w[j >> 5] >> (j & 31)) & 1u; }

// field of `len` <= 32 bits starting at codeword bit `first` (bit `first` is the most significant)
__device__ __forceinline__ uint32_t field91(const uint32_t* w, int first, int len) {
    uint32_t v = 0;
    for (int i = 0; i < len; ++i) v = (v << 1) | bit91(w, first + i);
    return v;
}

// Serial CRC check by one thread: true iff payload != 0 and syndrome == 0.
__device__ __forceinline__ bool crc_ok_serial(const uint32_t* w) {
    if ((w[0] | w[1] | (w[2] & 0x1FFFu)) == 0) return false;     // bits 0..76 all zero
    uint32_t syn = 0;
    for (int j = 0; j < 91; ++j)
        if (bit91(w, j)) syn ^= c_codec.crc_syn[j];
    return syn == 0;
}

// Per-lane copy of the syndrome table: lane l keeps the contributions of codeword bits l, 32+l, 64+l in registers
// (a lane-indexed read of __constant__ memory would serialise 32 ways).
struct LaneSyn { uint32_t s0, s1, s2; };
__device__ __forceinline__ LaneSyn load_lane_syn(int lane) {
    LaneSyn r;
    r.s0 = c_codec.crc_syn[lane];
    r.s1 = c_codec.crc_syn[32 + lane];
    r.s2 = (lane < 27) ? c_codec.crc_syn[64 + lane] : 0u;
    return r;
}

// CRC syndrome of a 91-bit word held identically by every lane (14 bits; 0 = CRC matches)
__device__ __forceinline__ uint32_t syndrome_warp(uint32_t w0, uint32_t w1, uint32_t w2, int lane, const LaneSyn& ls) {
    uint32_t syn = ((w0 >> lane) & 1u) ? ls.s0 : 0u;
    syn ^= ((w1 >> lane) & 1u) ? ls.s1 : 0u;
    syn ^= ((w2 >> lane) & 1u) ? ls.s2 : 0u;
    return __reduce_xor_sync(0xffffffffu, syn);
}

// Warp-cooperative CRC check: every lane passes the same three words.
__device__ __forceinline__ bool crc_ok_warp(uint32_t w0, uint32_t w1, uint32_t w2, int lane, const LaneSyn& ls) {
    if ((w0 | w1 | (w2 & 0x1FFFu)) == 0) return false;
    return syndrome_warp(w0, w1, w2, lane, ls) == 0;
}

// decoders.py:70-115 as accept/reject for one 29-bit call field.
__device__ __forceinline__ bool call29_ok(uint32_t c29, uint32_t i3) {
    const uint32_t n28 = c29 >> 1, p = c29 & 1u;
    if (n28 < 2063592u + 4194303u) return true;           // tokens, CQ nnn / CQ abcd, 22-bit hashes: always accepted
    // n28 == 6257895 wraps to 'ZZ9ZZZ' in the reference (negative index); same digits as the largest value
    uint32_t nn = (n28 == 6257895u) ? 262177559u : n28 - 6257896u;
    const uint32_t i5 = nn % 27u; nn /= 27u;
    const uint32_t i4 = nn % 27u; nn /= 27u;
    const uint32_t i3c = nn % 27u; nn /= 27u;
    const uint32_t i2 = nn % 10u; nn /= 10u;            // third char: digit index, or >= 10 ... see below
    // NB: the third alphabet has 27 entries (10 digits + 17 blanks) but the radix used by the packing is 10
    const uint32_t i1 = nn % 36u; nn /= 36u;
    const uint32_t i0 = nn;                              // 0..36 : ' ', 0-9, A-Z
    // suffix letters: blanks may only trail
    if ((i3c == 0 && (i4 | i5)) || (i4 == 0 && i5)) return false;
    (void)i2;                                            // always a digit with radix 10
    bool ok;
    uint32_t first_letter;                               // 0..25 when the first char of the stripped call is a letter, else 99
    if (i0 != 0) {
        // call = c0 c1 c2 [suffix]; c2 is a digit.  Shape 1: letter (not Q) + digit, except B,F,G,I,K,M,N,R,W + digit + digit.
        const bool c0_letter = i0 >= 11;
        const uint32_t l0 = i0 - 11;                     // A=0
        first_letter = c0_letter ? l0 : 99u;
        const bool c1_digit = i1 < 10;
        const uint32_t excl = (1u << 1) | (1u << 5) | (1u << 6) | (1u << 8) | (1u << 10) | (1u << 12) | (1u << 13) | (1u << 17) | (1u << 22);
        ok = c0_letter && l0 != 16 && c1_digit && !((excl >> l0) & 1u);
        if (!ok) {                                        // shape 2: two-character prefix + digit
            const uint32_t a = i0 - 1;                   // index in 0-9A-Z
            const uint32_t b = 36u * a + i1;
            ok = (c_codec.prefix2[b >> 5] >> (b & 31)) & 1u;
        }
    } else {
        // leading blank stripped: call = c1 c2 c3 ...; c2 digit, c3 must exist (length >= 3) and is a letter
        if (i3c == 0) return false;
        const bool c1_letter = i1 >= 10;
        const uint32_t l1 = i1 - 10;
        first_letter = c1_letter ? l1 : 99u;
        ok = c1_letter && l1 != 16;                      // shape 1 only (third char is a letter, so shape 2 cannot match)
    }
    if (!ok) return false;
    if (p && i3 == 1) {                                  // '/R' is kept only on A, K, N, W calls (decoders.py:90-91)
        if (!(first_letter == 0 || first_letter == 10 || first_letter == 13 || first_letter == 22)) return false;
    }
    return true;
}

// unpack(bits77) is not None  (decoders.py:16-68).  w = 91-bit word, only bits 0..76 are read.
__device__ __forceinline__ bool payload_valid(const uint32_t* w) {
    if ((w[0] | w[1] | (w[2] & 0x1FFFu)) == 0) return false;
    const uint32_t i3 = field91(w, 74, 3);
    if (i3 == 1 || i3 == 2) {
        const uint32_t g15 = field91(w, 59, 15);
        if (g15 == 0 || g15 == 32400u || g15 == 32401u) return false;
        return call29_ok(field91(w, 0, 29), i3) && call29_ok(field91(w, 29, 29), i3);
    }
    if (i3 == 4) {
        const uint32_t cq = bit91(w, 73), rrr = field91(w, 71, 2);
        return (cq != 0) != (rrr != 0);
    }
    return false;
}

// Out-of-line copy for the LDPC kernels: it runs only for the rare words whose CRC matches, and keeping its divisions out of
// the belief-propagation loop keeps that loop's instruction-cache footprint small (k_pass0 5.97 -> 5.79 ms; the OSD kernel
// is faster with the inlined form).
__device__ __noinline__ bool payload_valid_cold(const uint32_t* w) { return payload_valid(w); }

}  // namespace ft8
