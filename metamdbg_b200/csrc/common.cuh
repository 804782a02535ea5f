// common.cuh -- shared device helpers for libmdbg_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// Pure arithmetic helpers are host+device so that tests/cpp/device_math_test.cu can run them on a CPU (exhaustively
// for the 32-bit candidate test); the device code is unchanged by this (same intrinsics under __CUDA_ARCH__).
#define MDBG_HD __host__ __device__ __forceinline__

namespace mdbg {

MDBG_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {      // high word of (hi:lo) << s, s < 32
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return (uint32_t)((((uint64_t)hi << 32 | lo) << (s & 31)) >> 32);
#endif
}
MDBG_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {      // low word of (hi:lo) >> s, s < 32
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)(((uint64_t)hi << 32 | lo) >> (s & 31));
#endif
}
MDBG_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
MDBG_HD uint32_t brev32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// ------------------------------------------------------------------ MurmurHash3
// Arithmetic of MurmurHash3_x64_128 (reference: src/utils/MurmurHash3.cpp:246-405).

MDBG_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

MDBG_HD uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

constexpr uint64_t MURMUR_C1 = 0x87c37b91114253d5ULL;
constexpr uint64_t MURMUR_C2 = 0x4cf5ad432745937fULL;

// h1 of MurmurHash3_x64_128(&key, 8, seed=42): the 8-byte key takes the "tail"
// branch only (case 8..1), no body block (MurmurHash3.cpp:246-325, call site
// src/utils/kmer/Kmer.hpp:1421).
MDBG_HD uint64_t murmur_h1_u64(uint64_t key) {
    uint64_t k1 = key * MURMUR_C1;
    k1 = rotl64(k1, 31);
    k1 *= MURMUR_C2;
    uint64_t h1 = 42ULL ^ k1;
    uint64_t h2 = 42ULL;
    h1 ^= 8ULL; h2 ^= 8ULL;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    return h1 + h2;
}

// Cheap superset test for "murmur_h1_u64(key) <= T" with a 32-bit key.
//
//   h1 + h2 = fmix64(A) + fmix64(B),  A = (k1 ^ 34) + 34,  B = A + 34
//   (seed 42, len 8: h1 = 42 ^ k1 ^ 8, h2 = 42 ^ 8, then h1 += h2, h2 += h1).
//
// Only the HIGH words of the two fmix64 results are formed: with
// s0 = hi(fmix(A)) + hi(fmix(B)) (mod 2^32) the true high word of the sum is s0
// or s0 + 1 (carry out of the low words), so "hash <= T" implies s0 <= T_hi or
// s0 == 0xFFFFFFFF, i.e. (s0 + 1) <= T_hi + 1 in unsigned arithmetic.  The two
// +34 additions are done on the low word only; the 2^-26-rare case where one of
// them carries into the high word is detected (B_lo < 68) and reported as s1 = 0.
//
// murmur_s1_u32 returns s1 = s0 + 1 (or 0).  With T_hi = threshold >> 32:
//   s1 >  T_hi + 1            -> certainly not selected
//   1 <= s1 < T_hi            -> certainly selected (sum_hi <= s1 < T_hi)
//   otherwise (s1 in {0, T_hi, T_hi + 1}) -> undecided: callers run the exact murmur_h1_u64.
MDBG_HD uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }

// fmix64 up to its second multiply: returns the two words (plo, phi) of
// ((k ^ k>>33) * 0xff51afd7ed558ccd) ^ (.. >> 33); t = hi >> 1 and m = hi * 0xed558ccd are
// shared by A and B (same high word)
MDBG_HD void fmix64_front(uint32_t lo, uint32_t t, uint32_t m, uint32_t& plo, uint32_t& phi) {
    lo ^= t;                                                   // k ^= k >> 33
    const uint64_t p = (uint64_t)lo * 0xed558ccdu;             // k *= 0xff51afd7ed558ccd
    phi = (uint32_t)(p >> 32) + mad_lo(lo, 0xff51afd7u, m);
    plo = (uint32_t)p ^ (phi >> 1);                            // k ^= k >> 33
}

MDBG_HD uint32_t murmur_s1_u32(uint32_t key) {
    // k1 = key * c1 ; k1 = rotl64(k1, 31) ; k1 *= c2
    uint64_t p = (uint64_t)key * 0x114253d5u;
    uint32_t lo = (uint32_t)p;
    uint32_t hi = mad_lo(key, 0x87c37b91u, (uint32_t)(p >> 32));
    const uint32_t rlo = funnel_l(hi, lo, 31);          // low word of (k1 << 31) | (k1 >> 33)
    const uint32_t rhi = funnel_l(lo, hi, 31);
    p = (uint64_t)rlo * 0x2745937fu;
    hi = (uint32_t)(p >> 32) + mad_lo(rlo, 0x4cf5ad43u, rhi * 0x2745937fu);
    const uint32_t alo = ((uint32_t)p ^ 34u) + 34u;            // A (low word)
    const uint32_t blo = alo + 34u;                            // B (low word)
    const uint32_t t = hi >> 1;
    const uint32_t m = hi * 0xed558ccdu;
    uint32_t plo_a, phi_a, plo_b, phi_b;
    fmix64_front(alo, t, m, plo_a, phi_a);
    fmix64_front(blo, t, m, plo_b, phi_b);
    // high words of k * 0xc4ceb9fe1a85ec53 for A and B, summed (+1): the final k ^= k >> 33
    // only changes the low word
    uint32_t acc = mad_lo(phi_a, 0x1a85ec53u, 1u);
    acc = mad_lo(plo_a, 0xc4ceb9feu, acc);
    acc = mad_lo(phi_b, 0x1a85ec53u, acc);
    acc = mad_lo(plo_b, 0xc4ceb9feu, acc);
    const uint32_t s1 = acc + umulhi32(plo_a, 0x1a85ec53u) + umulhi32(plo_b, 0x1a85ec53u);
    return (blo < 68u) ? 0u : s1;                              // 0 = "undecided, run the exact hash"
}

// ---- variant 1 of the candidate filter (sketch kernel variant 1, see sketch.cu) -------------------------------------
// Same superset contract, fewer instructions.  The last fmix64 multiply is applied ONCE, to the sum of the two
// pre-images: h = (P_a ^ P_a >> 33) + (P_b ^ P_b >> 33) with P = k * 0xc4ceb9fe1a85ec53; the xor-shifts only touch
// the low words, so hi(h) = hi(P_a) + hi(P_b) + carry, and by linearity (k_a + k_b) * M has the same two high
// words plus ITS low-word carry.  s1(key) returns s' = s0 or s0 + 1, so that for a selected key (hash <= T)
//
//   s' <= T_hi + S1_SLACK     or     s' = 0xFFFFFFFF  (s0 = 0xFFFFFFFF with a carry: true high word 0).
//
// The second case, and the 2^-26-rare case where one of the two +34 additions (done on the low word only) carries
// into the high word, are reported through `risk`: s1 keeps risk = max(risk, ...) and the caller must treat ALL
// keys that went into one `risk` as candidates when risk >= S1_RISK.  Callers confirm every candidate with the
// exact murmur_h1_u64 (tests/cpp/device_math_test.cu walks all 2^32 keys: no selected key is ever rejected).
namespace k1v1 {

constexpr uint32_t S1_SLACK = 1;
constexpr uint32_t S1_RISK = 0xFFFFFFBCu;                      // (k1_lo ^ 34) + 68 wraps

MDBG_HD uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;                                                  // pinned: NVVM would re-associate a * b + c
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a * b + c;
#endif
}

MDBG_HD void fmix64_front(uint32_t lo, uint32_t t, uint32_t m, uint32_t& plo, uint32_t& phi) {
    lo ^= t;                                                   // k ^= k >> 33
    const uint64_t p = (uint64_t)lo * 0xed558ccdu;             // k *= 0xff51afd7ed558ccd
    phi = (uint32_t)(p >> 32) + mad_lo(lo, 0xff51afd7u, m);
    plo = (uint32_t)p ^ (phi >> 1);                            // k ^= k >> 33
}

MDBG_HD uint32_t murmur_s1_u32(uint32_t key, uint32_t& risk) {
    uint64_t p = (uint64_t)key * 0x114253d5u;
    uint32_t lo = (uint32_t)p;
    uint32_t hi = mad_lo(key, 0x87c37b91u, (uint32_t)(p >> 32));
    const uint32_t rlo = funnel_l(hi, lo, 31);
    const uint32_t rhi = funnel_l(lo, hi, 31);
    p = (uint64_t)rlo * 0x2745937fu;
    hi = mad_lo(rlo, 0x4cf5ad43u, mad_lo(rhi, 0x2745937fu, (uint32_t)(p >> 32)));
    const uint32_t x34 = (uint32_t)p ^ 34u;
    const uint32_t alo = x34 + 34u;                            // A (low word)
    const uint32_t blo = x34 + 68u;                            // B (low word)
    const uint32_t t = hi >> 1;
    const uint32_t m = hi * 0xed558ccdu;
    uint32_t plo_a, phi_a, plo_b, phi_b;
    fmix64_front(alo, t, m, plo_a, phi_a);
    fmix64_front(blo, t, m, plo_b, phi_b);
    uint32_t klo, khi;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=&r"(klo), "=r"(khi) : "r"(plo_a), "r"(plo_b), "r"(phi_a), "r"(phi_b));
#else
    klo = plo_a + plo_b;
    khi = phi_a + phi_b + (klo < plo_a ? 1u : 0u);
#endif
    uint32_t acc = mad_lo(klo, 0xc4ceb9feu, umulhi32(klo, 0x1a85ec53u));
    acc = mad_lo(khi, 0x1a85ec53u, acc);
    risk = max(risk, max(x34, acc));                           // acc >= S1_RISK also covers the wrap s' = 0xFFFFFFFF
    return acc;
}

}  // namespace k1v1

// MurmurHash3_x64_128_original(vec, 4*k bytes, seed 0) = KmerVec::hash128
// (src/Commons.hpp:941-969).  `get(i)` returns the i-th u32 of the normalized
// vector.  h1 -> high 64 bits of the u128, h2 -> low 64 bits.
template <typename Get>
MDBG_HD void murmur128_u32vec(Get get, int k, uint64_t& o1, uint64_t& o2) {
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = k >> 2;                      // 16-byte blocks = 4 u32
    for (int b = 0; b < nblocks; b++) {
        uint64_t k1 = (uint64_t)get(4 * b) | ((uint64_t)get(4 * b + 1) << 32);
        uint64_t k2 = (uint64_t)get(4 * b + 2) | ((uint64_t)get(4 * b + 3) << 32);
        k1 *= MURMUR_C1; k1 = rotl64(k1, 31); k1 *= MURMUR_C2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= MURMUR_C2; k2 = rotl64(k2, 33); k2 *= MURMUR_C1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const int rem = k & 3;                           // tail of 4, 8 or 12 bytes
    if (rem == 3) {
        uint64_t k2 = (uint64_t)get(4 * nblocks + 2);
        k2 *= MURMUR_C2; k2 = rotl64(k2, 33); k2 *= MURMUR_C1; h2 ^= k2;
    }
    if (rem >= 1) {
        uint64_t k1 = (uint64_t)get(4 * nblocks);
        if (rem >= 2) k1 |= (uint64_t)get(4 * nblocks + 1) << 32;
        k1 *= MURMUR_C1; k1 = rotl64(k1, 31); k1 *= MURMUR_C2; h1 ^= k1;
    }
    const uint64_t len = (uint64_t)k * 4;
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    o1 = h1; o2 = h2;
}

// ------------------------------------------------------------------ warp helpers

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// inclusive warp scan of a small integer
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (uint32_t)d) v += t;
    }
    return v;
}

// splitmix64 finaliser (synthetic-read generator only)
MDBG_HD uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace mdbg
