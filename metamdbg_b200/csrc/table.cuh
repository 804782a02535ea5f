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

// read-only probe (the table is not being modified while this runs)
__device__ __forceinline__ Slot* table_find(Slot* table, uint64_t mask, uint64_t lo, uint64_t hi) {
    uint64_t idx = lo & mask;
    for (uint64_t probe = 0; probe <= mask && probe < 4096; probe++) {
        Slot* s = table + idx;
        const uint64_t clo = s->lo, chi = s->hi;
        if (clo == lo && chi == hi) return s;
        if ((clo | chi) == 0) return nullptr;
        idx = (idx + 1) & mask;
    }
    return nullptr;
}

}  // namespace mdbg
