// common.cuh -- shared device helpers for libmdbg_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mdbg {

// ------------------------------------------------------------------ MurmurHash3
// Arithmetic of MurmurHash3_x64_128 (reference: src/utils/MurmurHash3.cpp:246-405).

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
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
__device__ __forceinline__ uint64_t murmur_h1_u64(uint64_t key) {
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

// MurmurHash3_x64_128_original(vec, 4*k bytes, seed 0) = KmerVec::hash128
// (src/Commons.hpp:941-969).  `get(i)` returns the i-th u32 of the normalized
// vector.  h1 -> high 64 bits of the u128, h2 -> low 64 bits.
template <typename Get>
__device__ __forceinline__ void murmur128_u32vec(Get get, int k, uint64_t& o1, uint64_t& o2) {
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
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace mdbg
