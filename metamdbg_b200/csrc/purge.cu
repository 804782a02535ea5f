// purge.cu -- K2: Commons::purgePalindrome (src/Commons.hpp:1617-1723) over the
// device-resident minimizer-space reads, plus the synthetic-read generator.
//
// The reference bans, one at a time, the first minimizer of the first
// palindromic window it finds scanning k = firstK..lastK-1 and then the start
// index (KmerVec::isPalindrome, src/Commons.hpp:918-921).  Because a palindrome
// of size k+2 always contains one of size k, only k = firstK and firstK+1 can
// ever be found, which turns the O(n^2 lastK) loop into O(n) per banned
// element; reads without such a window (almost all) are left untouched.
#include "common.cuh"
#include "engine.cuh"

namespace mdbg {

// True when the window of k minimizers starting at w is a palindrome
// (KmerVec::isPalindrome, Commons.hpp:918-921).
__device__ __forceinline__ bool window_is_palindrome(const uint32_t* w, uint32_t k) {
    for (uint32_t t = 0; t < k / 2; t++)
        if (w[t] != w[k - 1 - t]) return false;
    return true;
}

// purge_flag_kernel takes 32 consecutive reads per warp at a time: lane j loads the bounds of read r0 + j (one coalesced
// request), the warp walks the 32 reads with shuffles and votes once per batch -- a dependent pair of offset loads and
// a warp vote per read made the first version a chain of memory latencies (0.26 ms for 1 M reads, now 0.15).  The
// kernels that MOVE data per read (sketch.cu's compact_kernel, purge_compact_kernel) gain nothing from that form
// (measured): they run at DRAM speed.
//
// A palindromic window of size k+2 contains the palindromic window of size k
// with the same centre, and the reference scans k in ascending order
// (Commons.hpp:1631), so the first window it ever bans has size first_k or
// first_k+1.  A read is untouched iff it holds no palindrome of those two sizes.
__global__ void __launch_bounds__(256) purge_flag_kernel(const uint32_t* __restrict__ mins, const uint64_t* __restrict__ offs,
                                                         uint64_t n_reads, uint32_t first_k, uint32_t last_k,
                                                         uint8_t* __restrict__ flags, unsigned long long* n_flagged) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const bool two = first_k + 1 < last_k;               // sizes tested: first_k, and first_k + 1 when below last_k
    for (uint64_t r0 = warp * 32; r0 < n_reads; r0 += n_warps * 32) {
        const uint64_t rm = r0 + lane;
        uint64_t b_m = 0, e_m = 0;
        if (rm < n_reads) { b_m = offs[rm]; e_m = offs[rm + 1]; }
        const int cnt = n_reads - r0 < 32 ? (int)(n_reads - r0) : 32;
        uint32_t hits = 0;                               // bit j: this lane saw a palindrome in read r0 + j
        if (first_k == 4) {
#pragma unroll 2
            for (int j = 0; j < cnt; j++) {
                const uint64_t b = __shfl_sync(0xffffffffu, b_m, j), e = __shfl_sync(0xffffffffu, e_m, j);
                for (uint64_t g = b + lane; g + 4 <= e; g += 32) {
                    const bool five = two && g + 5 <= e;
                    const uint32_t w0 = mins[g], w1 = mins[g + 1], w2 = mins[g + 2], w3 = mins[g + 3];
                    const uint32_t w4 = five ? mins[g + 4] : 0u;
                    const bool hit = (w0 == w3 && w1 == w2) || (five && w0 == w4 && w1 == w3);
                    hits |= (hit ? 1u : 0u) << j;
                }
            }
        } else {
            for (int j = 0; j < cnt; j++) {
                const uint64_t b = __shfl_sync(0xffffffffu, b_m, j), e = __shfl_sync(0xffffffffu, e_m, j);
                for (uint64_t g = b + lane; g < e; g += 32)
                    for (uint32_t k = first_k; k < last_k && k < first_k + 2; k++)
                        if (g + k <= e && window_is_palindrome(mins + g, k)) hits |= 1u << j;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) hits |= __shfl_xor_sync(0xffffffffu, hits, d);
        if (rm < n_reads) flags[rm] = (uint8_t)((hits >> lane) & 1u);
        if (lane == 0 && hits) atomicAdd(n_flagged, (unsigned long long)__popc(hits));
    }
}

void launch_purge_flag(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, uint32_t first_k, uint32_t last_k,
                       uint8_t* flags, unsigned long long* n_flagged, cudaStream_t s) {
    if (n_reads == 0) return;
    uint64_t blocks = (n_reads + 255) / 256;                 // 32 reads per warp, 8 warps per block
    if (blocks > 148 * 16) blocks = 148 * 16;
    purge_flag_kernel<<<(unsigned)blocks, 256, 0, s>>>(mins, offs, n_reads, first_k, last_k, flags, n_flagged);
}

// Exact restatement of the banning loop for one flagged read, executed by one
// thread (flagged reads are rare).  keep[] is used as the banned bitmap.
__global__ void __launch_bounds__(128) purge_exact_kernel(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads,
                                                          const uint8_t* flags, uint32_t first_k, uint32_t last_k,
                                                          uint8_t* keep, uint32_t* new_cnt,
                                                          unsigned long long* n_changed) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t b = offs[r], e = offs[r + 1];
    const long n = (long)(e - b);
    if (!flags[r]) {
        new_cnt[r] = (uint32_t)n;
        return;
    }
    const uint32_t* m = mins + b;
    uint8_t* kp = keep + b;                              // 1 = kept (pre-set by the caller)
    for (;;) {                                           // Commons.hpp:1627
        bool has = false;
        // only sizes first_k and first_k+1 can be the first palindrome found (see purge_flag_kernel)
        for (long k = first_k; k < (long)last_k && k < (long)first_k + 2 && !has; k++) {
            const long i_max = n - k + 1;                // Commons.hpp:1635
            for (long i = 0; i < i_max && !has; i++) {
                if (!kp[i]) continue;
                // the next k kept minimizers from i: positions idx[0..k)
                // palindrome test needs pairs (t, k-1-t); gather lazily from both ends
                long cnt = 0, j = i;
                // find the k-th kept index from i (or fail)
                long last_idx = -1;
                for (; j < n; j++) {
                    if (!kp[j]) continue;
                    cnt++;
                    if (cnt == k) { last_idx = j; break; }
                }
                if (last_idx < 0) continue;
                long lo = i, hi = last_idx;
                bool pal = true;
                for (long t = 0; t < k / 2; t++) {
                    if (m[lo] != m[hi]) { pal = false; break; }
                    do { lo++; } while (!kp[lo]);
                    do { hi--; } while (!kp[hi]);
                }
                if (pal) { kp[i] = 0; has = true; }      // Commons.hpp:1669 bans the window's first minimizer
            }
        }
        if (!has) break;
    }
    uint32_t c = 0;
    for (long i = 0; i < n; i++) c += kp[i];
    new_cnt[r] = c;
    if (c != (uint32_t)n) atomicAdd(n_changed, 1ULL);
}

void launch_purge_exact(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, const uint8_t* flags,
                        uint32_t first_k, uint32_t last_k, uint8_t* keep, uint32_t* new_cnt,
                        unsigned long long* n_changed, cudaStream_t s) {
    if (n_reads == 0) return;
    purge_exact_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, s>>>(mins, offs, n_reads, flags, first_k, last_k,
                                                                        keep, new_cnt, n_changed);
}

// Utils::applyDensityThreshold (src/Commons.hpp:2507-2550): re-hash every stored minimizer value (u32 widened to
// u64, seed 42) and keep it iff hash < density * 2^64; keep[] / new_cnt[] feed the same scan + compaction as the purge.
__global__ void __launch_bounds__(256) density_filter_kernel(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads,
                                                             uint64_t threshold, uint32_t select_none, uint8_t* keep,
                                                             uint32_t* new_cnt, unsigned long long* n_changed) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        const uint64_t b = offs[r], e = offs[r + 1];
        uint32_t cnt = 0;
        for (uint64_t g0 = b; g0 < e; g0 += 32) {
            const uint64_t g = g0 + lane;
            bool k = false;
            if (g < e) {
                k = !select_none && murmur_h1_u64((uint64_t)mins[g]) <= threshold;
                keep[g] = k ? 1 : 0;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, k));
        }
        if (lane == 0) {
            new_cnt[r] = cnt;
            if ((uint64_t)cnt != e - b) atomicAdd(n_changed, 1ULL);
        }
    }
}

void launch_density_filter(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, uint64_t threshold,
                           uint32_t select_none, uint8_t* keep, uint32_t* new_cnt, unsigned long long* n_changed,
                           cudaStream_t s) {
    if (n_reads == 0) return;
    uint64_t blocks = (n_reads + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    density_filter_kernel<<<(unsigned)blocks, 256, 0, s>>>(mins, offs, n_reads, threshold, select_none, keep, new_cnt,
                                                           n_changed);
}

// out_rem (optional): rem[] of the NEW store (min(#minimizers from a position to the end of its read, 255); see
// fill_rem_kernel in kminmer.cu) as a by-product -- the table passes then need no pass of their own over the offsets.
__global__ void __launch_bounds__(256) purge_compact_kernel(const uint32_t* __restrict__ mins, const uint64_t* __restrict__ offs,
                                                            const uint64_t* __restrict__ new_offs,
                                                            const uint8_t* __restrict__ keep, uint64_t n_reads,
                                                            uint32_t* __restrict__ out_mins, uint8_t* __restrict__ out_rem) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        const uint64_t b = offs[r], e = offs[r + 1];
        uint64_t dst = new_offs[r];
        const uint64_t dst_end = new_offs[r + 1];
        for (uint64_t g0 = b; g0 < e; g0 += 32) {
            const uint64_t g = g0 + lane;
            const bool k = (g < e) && keep[g];
            const uint32_t mk = __ballot_sync(0xffffffffu, k);
            if (k) {
                const uint64_t o = dst + __popc(mk & ((1u << lane) - 1u));
                out_mins[o] = mins[g];
                if (out_rem) { const uint64_t left = dst_end - o; out_rem[o] = (uint8_t)(left > 255 ? 255 : left); }
            }
            dst += __popc(mk);
        }
    }
}

void launch_purge_compact(const uint32_t* mins, const uint64_t* offs, const uint64_t* new_offs, const uint8_t* keep,
                          uint64_t n_reads, uint32_t* out_mins, cudaStream_t s, uint8_t* out_rem) {
    if (n_reads == 0) return;
    uint64_t blocks = (n_reads + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    purge_compact_kernel<<<(unsigned)blocks, 256, 0, s>>>(mins, offs, new_offs, keep, n_reads, out_mins, out_rem);
}

// ------------------------------------------------------------------ synthetic reads (metamdbg_b200/synth.py)
__global__ void __launch_bounds__(256) synth_fill_kernel(uint8_t* bases, const uint64_t* offsets,
                                                         const uint64_t* vstart, const uint8_t* strand,
                                                         uint32_t n_reads, uint64_t read_index_base, uint64_t seed,
                                                         uint32_t err_q24) {
    const uint64_t GOLD = 0x9E3779B97F4A7C15ULL;
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint64_t b = offsets[r], n = offsets[r + 1] - b;
        const uint64_t vs = vstart[r];
        const bool rc = strand[r] != 0;
        const uint64_t rid = read_index_base + r;
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t q = rc ? vs + (n - 1) - i : vs + i;
            uint32_t x = (uint32_t)(mix64(q + seed) & 3);
            if (rc) x = 3 - x;
            const uint64_t e = mix64((rid * GOLD) ^ (i + seed * 0x632BE5ABULL));
            if ((e & 0xFFFFFFULL) < (uint64_t)err_q24) x = (x + 1 + (uint32_t)((e >> 24) % 3)) & 3;
            bases[b + i] = "ACGT"[x];
        }
    }
}

void launch_synth_fill(uint8_t* bases, const uint64_t* offsets, const uint64_t* vstart, const uint8_t* strand,
                       uint32_t n_reads, uint64_t read_index_base, uint64_t seed, uint32_t err_q24, cudaStream_t s) {
    if (n_reads == 0) return;
    unsigned blocks = n_reads < 148u * 16u ? n_reads : 148u * 16u;
    synth_fill_kernel<<<blocks, 256, 0, s>>>(bases, offsets, vstart, strand, n_reads, read_index_base, seed, err_q24);
}

}  // namespace mdbg
