// table.cuh -- read-side helpers of the 32-byte-slot open-addressing tables, shared by kminmer.cu and unitig.cu.
#pragma once

#include "common.cuh"
#include "engine.cuh"

namespace mdbg {

// element i of the normalized k-min-mer a slot stands for (Slot::ref: store window, possibly reversed, or a foreign vector)
__device__ __forceinline__ uint32_t vec_elem(const uint32_t* mins, const uint32_t* foreign, uint64_t ref, int k, int i) {
    const uint64_t idx = ref & REF_INDEX_MASK;
    if (ref & REF_FOREIGN) return foreign[idx * (uint64_t)k + i];
    return (ref & REF_REV) ? mins[idx + k - 1 - i] : mins[idx + i];
}

// Slot of a key in a table of `mask + 1` slots.  The capacity is NOT restricted to powers of two (a table sized for load
// factor 0.6 is on average 1.4x smaller than the next power of two, and the random-access rate of the insert pass falls
// by 40 % between a 0.5 GB and a 1 GB table on a B200): the 64-bit hash is scaled to the capacity by a 64 x 64 -> high 64
// multiply, linear probing wraps at the end.
__device__ __forceinline__ uint64_t slot_of(uint64_t lo, uint64_t mask) {
#ifdef __CUDA_ARCH__
    return __umul64hi(lo, mask + 1);
#else
    return (uint64_t)(((unsigned __int128)lo * (unsigned __int128)(mask + 1)) >> 64);
#endif
}
__device__ __forceinline__ uint64_t next_slot(uint64_t idx, uint64_t mask) { return idx == mask ? 0 : idx + 1; }

// read-only probe (the table is not being modified while this runs)
__device__ __forceinline__ Slot* table_find(Slot* table, uint64_t mask, uint64_t lo, uint64_t hi) {
    uint64_t idx = slot_of(lo, mask);
    for (uint64_t probe = 0; probe <= mask && probe < 4096; probe++) {
        Slot* s = table + idx;
        const uint64_t clo = s->lo, chi = s->hi;
        if (clo == lo && chi == hi) return s;
        if ((clo | chi) == 0) return nullptr;
        idx = next_slot(idx, mask);
    }
    return nullptr;
}

}  // namespace mdbg
