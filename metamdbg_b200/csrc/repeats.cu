// repeats.cu -- K1b: the repetitive-minimizer histogram of the ONT path.
//
// Replaces ReadSelection::determineRepetitiveMinimizers + CountMinimizerFunctor
// (src/readSelection/ReadSelection.hpp:497-625): every minimizer of the first (up to) 1 M reads, sketched at the
// correction density with an empty blacklist, is counted; the max(1, int(1e-5f * #distinct)) most frequent values
// become repetitiveMinimizers.bin, the blacklist of every later sketch.  Upstream counts in an unordered_map under
// an omp critical; here the minimizer store is counted into an open-addressing table of 64-bit (value, count) words.
#include "common.cuh"
#include "engine.cuh"

namespace mdbg {

constexpr unsigned long long MINCOUNT_EMPTY = ~0ULL;

__device__ __forceinline__ uint32_t mincount_home(uint32_t v, uint64_t mask) {
    return (uint32_t)(mix64((uint64_t)v + 0x9E3779B97F4A7C15ULL) & mask);
}

__global__ void __launch_bounds__(256) mincount_insert_kernel(const uint32_t* mins, uint64_t n, unsigned long long* table,
                                                              uint64_t mask, unsigned long long* n_distinct, uint32_t* full_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = mins[i];
    uint64_t idx = mincount_home(v, mask);
    for (uint64_t probe = 0; probe <= mask; probe++) {
        unsigned long long cur = table[idx];
        if (cur == MINCOUNT_EMPTY) {
            const unsigned long long old = atomicCAS(&table[idx], MINCOUNT_EMPTY, (unsigned long long)v << 32);
            if (old == MINCOUNT_EMPTY) { atomicAdd(n_distinct, 1ULL); cur = (unsigned long long)v << 32; }
            else cur = old;
        }
        if ((uint32_t)(cur >> 32) == v) { atomicAdd(&table[idx], 1ULL); return; }     // count in the low word
        idx = (idx + 1) & mask;
    }
    atomicExch(full_flag, 1u);
}

// hist[min(count, n_bins - 1)]++ over the occupied slots
__global__ void __launch_bounds__(256) mincount_hist_kernel(const unsigned long long* table, uint64_t capacity,
                                                            unsigned long long* hist, uint32_t n_bins) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = table[i];
        if (w == MINCOUNT_EMPTY) continue;
        const uint32_t c = (uint32_t)w;
        atomicAdd(&hist[c < n_bins - 1 ? c : n_bins - 1], 1ULL);
    }
}

// (value, count) of every entry with count >= min_count, in one (unspecified) order
__global__ void __launch_bounds__(256) mincount_emit_kernel(const unsigned long long* table, uint64_t capacity, uint32_t min_count,
                                                            unsigned long long* out, unsigned long long* cursor, uint64_t out_cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = table[i];
        if (w == MINCOUNT_EMPTY || (uint32_t)w < min_count) continue;
        const unsigned long long pos = atomicAdd(cursor, 1ULL);
        if (pos < out_cap) out[pos] = w;
    }
}

void launch_mincount_insert(const uint32_t* mins, uint64_t n, unsigned long long* table, uint64_t mask,
                            unsigned long long* n_distinct, uint32_t* full_flag, cudaStream_t s) {
    if (n == 0) return;
    mincount_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mins, n, table, mask, n_distinct, full_flag);
}
void launch_mincount_hist(const unsigned long long* table, uint64_t capacity, unsigned long long* hist, uint32_t n_bins, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    mincount_hist_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, hist, n_bins);
}
void launch_mincount_emit(const unsigned long long* table, uint64_t capacity, uint32_t min_count, unsigned long long* out,
                          unsigned long long* cursor, uint64_t out_cap, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    mincount_emit_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, min_count, out, cursor, out_cap);
}

}  // namespace mdbg
