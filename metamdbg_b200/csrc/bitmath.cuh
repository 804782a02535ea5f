// bitmath.cuh -- pure SIMD-in-register helpers of the sketch kernel (K1).  Host+device so that
// tests/cpp/device_math_test.cu can check them on a CPU against naive loops; the kernel itself is in sketch.cu.
#pragma once

#include "common.cuh"

namespace mdbg {

// 0x80 in every byte of x that is non-zero (exact for arbitrary bytes)
MDBG_HD uint32_t nonzero_bytes(uint32_t x) {
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
// four words of per-byte 0x80 flags -> 16-bit mask (bit 4q+b = byte b of word q)
MDBG_HD uint32_t flags_to_mask16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    // multiply gathers bits 7,15,23,31 into bits 28..31
    const uint32_t a = (f0 * 0x00204081u) >> 28;
    const uint32_t b = (f1 * 0x00204081u) >> 24;
    const uint32_t c = (f2 * 0x00204081u) >> 20;
    const uint32_t d = (f3 * 0x00204081u) >> 16;
    return a | (b & 0xF0u) | (c & 0xF00u) | (d & 0xF000u);
}

// gather the 16 even bits of x into the low 16 bits
MDBG_HD uint32_t even_bits16(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    return (x | (x >> 8)) & 0xFFFFu;
}

template <int L>
MDBG_HD uint32_t byte_of(const uint32_t (&W)[8], int t) {
    return (W[t >> 2] >> (8 * (t & 3))) & 0xFFu;
}

// 4 base codes (one per byte, each < 4) -> 8 bits, first base most significant / least significant
MDBG_HD uint32_t pack4_msb(uint32_t w) { return (w * 0x40100401u) >> 24; }
MDBG_HD uint32_t pack4_lsb(uint32_t w) { return (w * 0x01041040u) >> 24; }

// reverse complement of a 2L-bit l-mer value (A0 C1 T2 G3: complement = code ^ 2)
template <int L>
MDBG_HD uint32_t revcomp_lmer(uint32_t fwd) {
    constexpr uint32_t MASK = (L < 16) ? ((1u << (2 * L)) - 1u) : 0xFFFFFFFFu;
    uint32_t x = brev32(fwd ^ (0xAAAAAAAAu & MASK));
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
    return x >> (32 - 2 * L);
}

// Unrolled register path: 16 consecutive l-mers for this lane out of the
// 32 ring bytes W (no invalid code present).  Returns the 16-bit CANDIDATE
// mask (superset of the selected positions, see murmur_s1_u32); the forward
// l-mer of the last candidate is left in sel_fwd.
template <int L>
MDBG_HD uint32_t roll16_fast(const uint32_t (&W)[8], uint32_t thr_hi_plus1, uint32_t& sel_fwd) {
    constexpr uint32_t MASK = (L < 16) ? ((1u << (2 * L)) - 1u) : 0xFFFFFFFFu;
    constexpr uint32_t INIT_MASK = (1u << (2 * (L - 1))) - 1u;      // L-1 <= 15 bases
    // state after the first L-1 bases, built with two multiplies per 4 bases instead of L-1 roll steps
    const uint32_t pf = (pack4_msb(W[0]) << 24) | (pack4_msb(W[1]) << 16) | (pack4_msb(W[2]) << 8) | pack4_msb(W[3]);
    const uint32_t pr = pack4_lsb(W[0]) | (pack4_lsb(W[1]) << 8) | (pack4_lsb(W[2]) << 16) | (pack4_lsb(W[3]) << 24);
    uint32_t fwd = pf >> (2 * (16 - (L - 1)));
    uint32_t rc = ((pr & INIT_MASK) ^ (0xAAAAAAAAu & INIT_MASK)) << 2;
    uint32_t sel = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t c = byte_of<L>(W, L - 1 + j);
        fwd = ((fwd << 2) | c) & MASK;
        rc = (rc >> 2) | ((c ^ 2u) << (2 * L - 2));
        const uint32_t s1 = murmur_s1_u32(min(fwd, rc));
        if (s1 <= thr_hi_plus1) {
            sel |= 1u << j;
            sel_fwd = fwd;
        }
    }
    return sel;
}

// ---- variant 1 of the register path: the lane's 32 codes are packed ONCE into (s_hi:s_lo), first base most
// significant (code i at bits 63-2i, 62-2i), and into the complemented, reversed (r_hi:r_lo); every forward /
// reverse-complement l-mer is then one funnel shift + mask instead of a roll step, and the accept bits are
// collected by an add-with-carry chain.  lmer_from_packed() re-derives the forward l-mer of any position.
namespace k1v1 {

template <int L>
MDBG_HD uint32_t lmer_from_packed(uint32_t s_hi, uint32_t s_lo, uint32_t j) {
    constexpr uint32_t MASK = (L < 16) ? ((1u << (2 * L)) - 1u) : 0xFFFFFFFFu;
    const uint32_t sh = 64u - 2u * j - 2u * L;                      // 2 .. 34
    return (sh >= 32u ? (s_hi >> (sh - 32u)) : funnel_r(s_lo, s_hi, sh)) & MASK;
}

// core of variant 1 / 2: 16 consecutive l-mers out of (s_hi:s_lo) = the lane's 32 codes, first base most significant,
// and (r_lo, r_hi) = the same codes complemented, first base LEAST significant (r_lo = codes 0..15)
template <int L>
MDBG_HD uint32_t roll16_core(uint32_t s_hi, uint32_t s_lo, uint32_t r_lo, uint32_t r_hi, uint32_t thr_cand) {
    constexpr uint32_t MASK = (L < 16) ? ((1u << (2 * L)) - 1u) : 0xFFFFFFFFu;
    // rejected positions are shifted into `rej` (position 0 ends up in bit 15) by a carry chain: s1 + ~thr_cand
    // carries out of 32 bits exactly when s1 > thr_cand, and the add-with-carry doubles rej and takes the carry in
    uint32_t rej = 0, risk = 0;
    const uint32_t not_thr = ~thr_cand;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t fwd, rc;
        if constexpr (L == 15) {
            // a 32-bit window holds 16 codes = TWO consecutive 15-mers: one funnel shift per pair of positions and
            // orientation, then a shift (no mask needed) for one l-mer and a mask for the other
            const int e = j & ~1;
            const uint32_t wf = e == 0 ? s_hi : funnel_r(s_lo, s_hi, 32 - 2 * e);      // codes e .. e+15, first base on top
            const uint32_t wr = e == 0 ? r_lo : funnel_r(r_lo, r_hi, 2 * e);           // same codes, first base at the bottom
            fwd = (j & 1) ? (wf & MASK) : (wf >> 2);
            rc = (j & 1) ? (wr >> 2) : (wr & MASK);
        } else {
            const int sh = 64 - 2 * j - 2 * L;                      // S >> sh, low 2L bits
            fwd = (sh >= 32 ? (s_hi >> (sh - 32)) : funnel_r(s_lo, s_hi, sh)) & MASK;
            rc = (j == 0 ? r_lo : funnel_r(r_lo, r_hi, 2 * j)) & MASK;
        }
        const uint32_t s1 = murmur_s1_u32(min(fwd, rc), risk);
#ifdef __CUDA_ARCH__
        uint32_t scratch;
        asm("add.cc.u32 %1, %2, %3;\n\taddc.u32 %0, %0, %0;" : "+r"(rej), "=&r"(scratch) : "r"(s1), "r"(not_thr));
#else
        rej = 2u * rej + (s1 > thr_cand ? 1u : 0u);
#endif
    }
    if (risk >= S1_RISK) return 0xFFFFu;                            // 2^-26 per key: let the exact test decide all 16
    return brev32(~rej) >> 16;                                      // accepted positions, position j in bit j
}

// bit-reverse a word of 16 two-bit codes code-wise: code i moves from bits [2i, 2i+1] to [30-2i, 31-2i]
MDBG_HD uint32_t reverse_codes16(uint32_t x) {
    x = brev32(x);
    return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

template <int L>
MDBG_HD uint32_t roll16_fast(const uint32_t (&W)[8], uint32_t thr_cand, uint32_t& s_hi, uint32_t& s_lo) {
    // The ring bytes past the read's last base are whatever shared memory held (not necessarily codes, and not
    // necessarily flagged by the bit-2 test): keep 2 bits of every byte, or pack4's multiply lets such a byte carry
    // into the fields of the codes before it -- codes that valid positions of this lane still need.
    constexpr uint32_t M = 0x03030303u;
    s_hi = (pack4_msb(W[0] & M) << 24) | (pack4_msb(W[1] & M) << 16) | (pack4_msb(W[2] & M) << 8) | pack4_msb(W[3] & M);
    s_lo = (pack4_msb(W[4] & M) << 24) | (pack4_msb(W[5] & M) << 16) | (pack4_msb(W[6] & M) << 8) | pack4_msb(W[7] & M);
    const uint32_t r_lo = reverse_codes16(s_hi ^ 0xAAAAAAAAu), r_hi = reverse_codes16(s_lo ^ 0xAAAAAAAAu);
    return roll16_core<L>(s_hi, s_lo, r_lo, r_hi, thr_cand);
}

// variant 2 (packed ring): the ring already holds two-bit codes, 16 per word, first base in the low bits --
// `lo` = codes 0..15 of the lane, `hi` = codes 16..31.  That IS the reverse-complement orientation (up to the
// complement xor); the forward orientation is one code-wise reversal per word.
template <int L>
MDBG_HD uint32_t roll16_packed(uint32_t lo, uint32_t hi, uint32_t thr_cand, uint32_t& s_hi, uint32_t& s_lo) {
    s_hi = reverse_codes16(lo);
    s_lo = reverse_codes16(hi);
    return roll16_core<L>(s_hi, s_lo, lo ^ 0xAAAAAAAAu, hi ^ 0xAAAAAAAAu, thr_cand);
}

}  // namespace k1v1

}  // namespace mdbg
