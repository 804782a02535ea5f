// kminmer.cu -- K3/K4/K5: k-min-mer extraction, orientation normalisation,
// 128-bit Murmur hashing and the GPU-resident open-addressing count table.
//
// Replaces (reference, paths relative to the metaMDBG tree):
//   MDBG::getKminmers_complete     src/Commons.hpp:5282-5361 (active #else branch)
//   KmerVec::normalize / hash128   src/Commons.hpp:886-916, 941-969
//   KminmerCounter::partitionKminmer / dereplicatePartition / dumpKminmer
//                                  src/graph/CreateMdbg.hpp:3714-3883
// The reference sorts every occurrence on disk and run-length counts; here each
// occurrence is one 128-bit compare-and-swap probe into a 32-byte-slot table
// plus one 32-bit reduction, so the table is the only state.
#include "common.cuh"
#include "engine.cuh"
#include "table.cuh"

namespace mdbg {

// ------------------------------------------------------------------ 128-bit slot primitives

__device__ __forceinline__ void load_key(const Slot* s, uint64_t& lo, uint64_t& hi) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(s) : "memory");
}

// atom.cas.b128 (sm_90+): claims an empty slot with the full 128-bit key in one
// transaction, so a key is never visible half-written to another CAS.
__device__ __forceinline__ void cas_key(Slot* s, uint64_t new_lo, uint64_t new_hi, uint64_t& old_lo, uint64_t& old_hi) {
    asm volatile(
        "{\n\t"
        ".reg .b128 cmp, swp, old;\n\t"
        "mov.b128 cmp, {%3, %3};\n\t"
        "mov.b128 swp, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 old, [%2], cmp, swp;\n\t"
        "mov.b128 {%0, %1}, old;\n\t"
        "}"
        : "=l"(old_lo), "=l"(old_hi)
        : "l"(s), "l"(0ULL), "l"(new_lo), "l"(new_hi)
        : "memory");
}

// Insert-or-add.  Returns 0 when the probe limit is hit (table full), 1 when the key was there, 2 when this call
// claimed a fresh slot for it (callers count claims: that is the number of distinct keys, and the load factor).
__device__ __forceinline__ int table_add(Slot* table, uint64_t mask, uint64_t lo, uint64_t hi, uint32_t add,
                                         uint64_t ref) {
    uint64_t idx = slot_of(lo, mask);
    const uint64_t max_probe = (mask + 1 < 4096) ? mask + 1 : 4096;
    for (uint64_t probe = 0; probe < max_probe; probe++) {
        Slot* s = table + idx;
        uint64_t clo, chi;
        load_key(s, clo, chi);
        // A plain 16-byte load may in principle observe a half-written key, so a
        // hit is only trusted when both halves are non-zero; anything with a zero
        // half is re-examined by the atomic CAS.
        if (clo == lo && chi == hi && lo != 0 && hi != 0) {
            atomicAdd(&s->count, add);
            return 1;
        }
        if (clo == 0 || chi == 0) {
            uint64_t olo, ohi;
            cas_key(s, lo, hi, olo, ohi);
            if (olo == 0 && ohi == 0) {                  // claimed
                s->ref = ref;
                atomicAdd(&s->count, add);
                return 2;
            }
            if (olo == lo && ohi == hi) {
                atomicAdd(&s->count, add);
                return 1;
            }
        }
        idx = next_slot(idx, mask);
    }
    return 0;
}

// Block-wide bookkeeping of an insert pass.  block_begin: thread 0 reads the abandon flag ONCE for the block (a
// same-address load issued by every warp of a 200 000-block grid serialises in one L2 slice: 1.7 M requests at ~3
// cycles each were the whole 2.7 ms of the round-2 insert kernel, whatever the table size) and zeroes the block's
// counters; the caller does its independent loads and hashing, then block_ready() -- one barrier -- tells whether the
// pass is being abandoned (table too small: the host rebuilds it; blocks that start afterwards skip their probes).
// block_claims adds the block's number of claimed slots to *claims and raises *full_flag when a probe sequence ran out
// or the table passed its load limit.  Every thread of the block must call all three.
struct BlockPass { uint32_t claims, fail, abandon; };
__device__ __forceinline__ void block_begin(BlockPass& st, const uint32_t* full_flag) {
    if (threadIdx.x == 0) {
        st.claims = 0;
        st.fail = 0;
        st.abandon = *reinterpret_cast<const volatile uint32_t*>(full_flag);
    }
}
__device__ __forceinline__ bool block_ready(BlockPass& st) {
    __syncthreads();
    return st.abandon == 0;
}
__device__ __forceinline__ void block_claims(BlockPass& st, int status, bool active, unsigned long long* claims,
                                             unsigned long long claim_limit, uint32_t* full_flag) {
    const uint32_t mc = __ballot_sync(0xffffffffu, active && status == 2);
    const uint32_t mf = __ballot_sync(0xffffffffu, active && status == 0);
    if ((threadIdx.x & 31) == 0) {
        if (mc) atomicAdd(&st.claims, (uint32_t)__popc(mc));
        if (mf) atomicAdd(&st.fail, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool full = st.fail != 0;
        if (st.claims) full |= atomicAdd(claims, (unsigned long long)st.claims) + st.claims > claim_limit;
        if (full) atomicExch(full_flag, 1u);
    }
}

// Warp-wide bookkeeping of a pass, without any block barrier (PassAux, engine.cuh).  The abandon flag exists in 64
// copies, each in its own 128-byte line: lane 0 of every warp reads the copy of its warp index -- 1.7 M reads spread
// over 64 addresses instead of one (a single address is served by one L2 slice at about one request per 3 cycles).  The
// claim counter is sharded the same way; a shard that passes its share of the load limit raises every copy of the
// flag.  pass_fold_kernel folds the shards into *claims and the flag into *full_flag after the pass, which is what
// the host reads.  `shards` is a power of two <= 64 (1 for small tables, where a share would be a handful of slots).
__device__ __forceinline__ uint32_t warp_abandon(const PassAux* x, uint32_t wid) {
    uint32_t f = 0;
    if ((threadIdx.x & 31) == 0) f = *reinterpret_cast<const volatile uint32_t*>(&x->flag[wid & 63][0]);
    return __shfl_sync(0xffffffffu, f, 0);
}
__device__ __forceinline__ void warp_claims(PassAux* x, uint32_t shards, uint32_t wid, int status, bool active,
                                            unsigned long long claim_limit) {
    const uint32_t mc = __ballot_sync(0xffffffffu, active && status == 2);
    const uint32_t mf = __ballot_sync(0xffffffffu, active && status == 0);
    if ((mc | mf) == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t full = mf != 0;
    if (lane == 0 && mc) {
        const uint32_t c = (uint32_t)__popc(mc);
        full |= atomicAdd(&x->claims[wid & (shards - 1)][0], (unsigned long long)c) + c > claim_limit / shards;
    }
    if (__shfl_sync(0xffffffffu, full, 0)) {
        *reinterpret_cast<volatile uint32_t*>(&x->flag[lane][0]) = 1u;
        *reinterpret_cast<volatile uint32_t*>(&x->flag[lane + 32][0]) = 1u;
    }
}
__global__ void __launch_bounds__(64) pass_fold_kernel(PassAux* x, unsigned long long* claims, uint32_t* full_flag) {
    unsigned long long c = x->claims[threadIdx.x][0];
    uint32_t f = x->flag[threadIdx.x][0];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c += __shfl_down_sync(0xffffffffu, c, d);
        f |= __shfl_down_sync(0xffffffffu, f, d);
    }
    __shared__ unsigned long long sc[2];
    __shared__ uint32_t sf[2];
    if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = c; sf[threadIdx.x >> 5] = f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        *claims = sc[0] + sc[1];                           // the shards are cumulative over the passes into this table
        if (sf[0] | sf[1]) *full_flag = 1u;
    }
}

// normalized hash of the k-window starting at w
__device__ __forceinline__ void window_hash(const uint32_t* w, int k, uint64_t& h1, uint64_t& h2, bool& rev) {
    rev = true;
    for (int j = 0; j < k / 2; j++) {
        const uint32_t x = w[j], y = w[k - 1 - j];
        if (x != y) { rev = x > y; break; }
    }
    // ONE copy of the hash for both orientations (element i of the normalized vector = w[first + i * step]): a warp
    // whose lanes disagree on the orientation would otherwise run the ~90 instructions of Murmur twice
    const int first = rev ? k - 1 : 0, step = rev ? -1 : 1;
    murmur128_u32vec([&](int i) { return w[first + i * step]; }, k, h1, h2);
}

// The window starting at flat minimizer index g: does it exist (the read still holds k minimizers from g on --
// getKminmers_complete: i in [0, n-k]), its orientation (KmerVec::normalize, Commons.hpp:886-916: the first differing
// pair decides, a palindromic vector counts as reversed) and the hash128 of the normalized vector.  K_FIXED = 4: rem
// and the four minimizers are loaded side by side (one memory latency instead of three; the store's buffers end in
// >= 256 bytes of slack, so the three loads past a read's end stay inside the allocation) and every in-range lane hashes.
template <int K_FIXED>
__device__ __forceinline__ void window_key(const uint32_t* mins, const uint8_t* rem, uint64_t g, uint64_t g_hi, int k,
                                           bool& active, bool& rev, uint64_t& h1, uint64_t& h2) {
    active = false;
    if (g >= g_hi) return;
    if (K_FIXED == 4) {
        const uint32_t r = rem[g];
        const uint32_t* w = mins + g;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        active = r >= 4;
        rev = w0 != w3 ? w0 > w3 : (w1 != w2 ? w1 > w2 : true);
        const uint32_t x[4] = {rev ? w3 : w0, rev ? w2 : w1, rev ? w1 : w2, rev ? w0 : w3};
        murmur128_u32vec([&](int i) { return x[i]; }, 4, h1, h2);
    } else {
        active = (int)rem[g] >= k;
        if (active) window_hash(mins + g, k, h1, h2, rev);
    }
}

// ------------------------------------------------------------------ rem[] = minimizers left in the read
__global__ void __launch_bounds__(256) fill_rem_kernel(const uint64_t* __restrict__ offs, uint64_t read_lo, uint64_t read_hi,
                                                       uint8_t* __restrict__ rem) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    // 32 reads per warp and step: lane j loads the bounds of read r0 + j, then all lanes sweep the (contiguous) positions
    // of those reads; a position finds the end of its read among the 32 ends by a binary search over shuffles
    for (uint64_t r0 = read_lo + warp * 32; r0 < read_hi; r0 += n_warps * 32) {
        const uint64_t rm = r0 + lane;
        const uint64_t last = read_hi - r0 < 32 ? read_hi - r0 : 32;        // reads in this batch
        const uint64_t e_m = offs[(rm < read_hi ? rm : read_hi - 1) + 1];      // lanes past the batch repeat the last end
        const uint64_t b0 = offs[r0], e_last = __shfl_sync(0xffffffffu, e_m, (int)last - 1);
        for (uint64_t g0 = b0; g0 < e_last; g0 += 32) {                        // (uniform trip count: shuffles inside)
            const uint64_t g = g0 + lane;
            int lo = 0;                                                        // first read of the batch whose end is > g
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint64_t e_probe = __shfl_sync(0xffffffffu, e_m, lo + step - 1);
                if (e_probe <= g) lo += step;
            }
            const uint64_t e = __shfl_sync(0xffffffffu, e_m, lo > 31 ? 31 : lo);
            if (g < e_last) {
                const uint64_t left = e - g;
                rem[g] = (uint8_t)(left > 255 ? 255 : left);
            }
        }
    }
}

void launch_fill_rem(const uint64_t* offs, uint64_t read_lo, uint64_t read_hi, uint8_t* rem, cudaStream_t s) {
    if (read_hi <= read_lo) return;
    uint64_t blocks = (read_hi - read_lo + 255) / 256;       // 32 reads per warp, 8 warps per block
    if (blocks > 148 * 16) blocks = 148 * 16;
    fill_rem_kernel<<<(unsigned)blocks, 256, 0, s>>>(offs, read_lo, read_hi, rem);
}

// ------------------------------------------------------------------ insert: one thread per window
// Window starting at flat minimizer index g exists iff the read still holds k
// minimizers from g on (getKminmers_complete: i in [0, n-k]).
template <int K_FIXED>
__global__ void __launch_bounds__(256) insert_kernel(const InsertArgs a) {
    __shared__ BlockPass st;
    block_begin(st, a.full_flag);
    const uint64_t g = a.g_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = K_FIXED ? K_FIXED : (int)a.k;
    bool active = false, rev = true;
    uint64_t h1 = 0, h2 = 0;
    window_key<K_FIXED>(a.mins, a.rem, g, a.g_hi, k, active, rev, h1, h2);
    int status = 1;
    const bool go = block_ready(st);                                         // false: the pass is being abandoned
    if (go && active) status = table_add(a.table, a.mask, h2, h1, 1u, g | (rev ? REF_REV : 0ULL));
    block_claims(st, status, go && active, a.claims, a.claim_limit, a.full_flag);
}

// The same pass without a block barrier (see PassAux): a warp is done as soon as its own 32 probes are.
template <int K_FIXED>
__global__ void __launch_bounds__(128) insert_warp_kernel(const InsertArgs a) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t wid = (uint32_t)(t >> 5);
    const uint32_t abandon = warp_abandon(a.aux, wid);                        // (in flight behind the loads below)
    const uint64_t g = a.g_lo + t;
    const int k = K_FIXED ? K_FIXED : (int)a.k;
    bool active = false, rev = true;
    uint64_t h1 = 0, h2 = 0;
    window_key<K_FIXED>(a.mins, a.rem, g, a.g_hi, k, active, rev, h1, h2);
    int status = 1;
    const bool go = abandon == 0;
    if (go && active) status = table_add(a.table, a.mask, h2, h1, 1u, g | (rev ? REF_REV : 0ULL));
    warp_claims(a.aux, a.aux_shards, wid, status, go && active, a.claim_limit);
}

// PASS_VARIANT (MDBG_PASS_VARIANT, default 1): 0 = block form (one abandon-flag read and two barriers per 256 windows),
// 1 = warp form (no barrier; abandon flag replicated, claim counter sharded -- PassAux)
static int pass_variant() {
    static const int v = [] { const char* e = getenv("MDBG_PASS_VARIANT"); return e ? atoi(e) : 1; }();
    return v;
}

// kernels one launch_insert / launch_next_k call enqueues (the warp form adds pass_fold_kernel)
int pass_kernels_per_launch(bool have_aux) { return pass_variant() != 0 && have_aux ? 2 : 1; }

void launch_insert(const InsertArgs& a, cudaStream_t s) {
    if (a.g_hi <= a.g_lo) return;
    const uint64_t n = a.g_hi - a.g_lo;
    if (pass_variant() == 0 || !a.aux) {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (a.k == 4) insert_kernel<4><<<blocks, 256, 0, s>>>(a);
        else insert_kernel<0><<<blocks, 256, 0, s>>>(a);
    } else {
        const unsigned blocks = (unsigned)((n + 127) / 128);
        if (a.k == 4) insert_warp_kernel<4><<<blocks, 128, 0, s>>>(a);
        else insert_warp_kernel<0><<<blocks, 128, 0, s>>>(a);
        pass_fold_kernel<<<1, 64, 0, s>>>(a.aux, a.claims, a.full_flag);
    }
}

// Insert-if-absent with a VALUE (next-k tables: the abundance is a function of the key -- min over the two
// (k-1)-min-mers of the replicated previous-k table -- so every rank that met the key computed the same number
// and the first writer wins).
__device__ __forceinline__ int table_put(Slot* table, uint64_t mask, uint64_t lo, uint64_t hi, uint32_t value,
                                         uint64_t ref) {
    uint64_t idx = slot_of(lo, mask);
    const uint64_t max_probe = (mask + 1 < 4096) ? mask + 1 : 4096;
    for (uint64_t probe = 0; probe < max_probe; probe++) {
        Slot* s = table + idx;
        uint64_t clo, chi;
        load_key(s, clo, chi);
        if (clo == lo && chi == hi && lo != 0 && hi != 0) return 1;
        if (clo == 0 || chi == 0) {
            uint64_t olo, ohi;
            cas_key(s, lo, hi, olo, ohi);
            if (olo == 0 && ohi == 0) {                  // claimed
                s->ref = ref;
                s->count = value;
                return 2;
            }
            if (olo == lo && ohi == hi) return 1;
        }
        idx = next_slot(idx, mask);
    }
    return 0;
}

__global__ void __launch_bounds__(256) insert_vecs_kernel(const InsertVecArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int k = (int)a.k;
    const uint32_t* w = a.vecs + i * (uint64_t)k;
    uint64_t h1, h2;
    murmur128_u32vec([&](int j) { return w[j]; }, k, h1, h2);   // already normalized by the sender
    const uint64_t ref = REF_FOREIGN | (a.foreign_base + i);
    const int ok = a.assign ? table_put(a.table, a.mask, h2, h1, a.counts[i], ref)
                            : table_add(a.table, a.mask, h2, h1, a.counts[i], ref);
    if (!ok) atomicExch(a.full_flag, 1u);
}

void launch_insert_vecs(const InsertVecArgs& a, cudaStream_t s) {
    if (a.n == 0) return;
    insert_vecs_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ stats / emit
__global__ void __launch_bounds__(256) table_stats_kernel(const Slot* table, uint64_t capacity, uint32_t min_count,
                                                          TableStats* out) {
    unsigned long long ne = 0, nd = 0, ni = 0, cs = 0, nr = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 q0 = reinterpret_cast<const uint4*>(table + i)[0];
        const uint4 q1 = reinterpret_cast<const uint4*>(table + i)[1];
        const uint32_t count = q1.x, flags = q1.y;
        const uint64_t lo = (uint64_t)q0.x | ((uint64_t)q0.y << 32);
        const uint64_t hi = (uint64_t)q0.z | ((uint64_t)q0.w << 32);
        if (lo | hi) {
            nd++;
            ni += count;
            if (count >= min_count) { ne++; cs += (unsigned long long)count * lo; }
            else if (flags & SLOT_RESCUED) { ne++; nr++; cs += (unsigned long long)count * lo; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        ne += __shfl_down_sync(0xffffffffu, ne, d);
        nd += __shfl_down_sync(0xffffffffu, nd, d);
        ni += __shfl_down_sync(0xffffffffu, ni, d);
        cs += __shfl_down_sync(0xffffffffu, cs, d);
        nr += __shfl_down_sync(0xffffffffu, nr, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ne) atomicAdd(&out->n_entries, ne);
        if (nd) atomicAdd(&out->n_distinct, nd);
        if (ni) atomicAdd(&out->n_instances, ni);
        if (cs) atomicAdd(&out->checksum, cs);
        if (nr) atomicAdd(&out->n_rescued, nr);
    }
}

void launch_table_stats(const Slot* table, uint64_t capacity, uint32_t min_count, TableStats* d_stats, cudaStream_t s) {
    cudaMemsetAsync(d_stats, 0, sizeof(TableStats), s);
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    table_stats_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, min_count, d_stats);
}

// Block-aggregated compaction: one atomicAdd per 256-slot tile instead of one per warp/thread
// (tens of millions of same-address atomics serialise in L2).
__global__ void __launch_bounds__(256) table_emit_kernel(const EmitArgs a) {
    __shared__ uint32_t wcnt[8];
    __shared__ unsigned long long tile_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = (a.capacity + 255) / 256;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i = tile * 256 + threadIdx.x;
        bool take = false;
        uint64_t lo = 0, hi = 0, ref = 0;
        uint32_t count = 0;
        if (i < a.capacity) {
            const Slot sl = a.table[i];
            lo = sl.lo; hi = sl.hi; count = sl.count; ref = sl.ref;
            take = (lo | hi) != 0 && (count >= a.min_count || (sl.flags & SLOT_RESCUED));
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = wcnt[w]; wcnt[w] = tot; tot += c; }
            tile_base = tot ? atomicAdd(a.cursor, (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        if (take) {
            const uint64_t pos = tile_base + wcnt[warp] + __popc(m & ((1u << lane) - 1u));
            a.out_hashes[2 * pos] = lo;
            a.out_hashes[2 * pos + 1] = hi;
            a.out_abund[pos] = count;
            if (a.out_vecs)
                for (int j = 0; j < (int)a.k; j++)
                    a.out_vecs[pos * a.k + j] = vec_elem(a.mins, a.foreign_vecs, ref, (int)a.k, j);
        }
        __syncthreads();
    }
}

void launch_table_emit(const EmitArgs& a, cudaStream_t s) {
    uint64_t blocks = (a.capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    table_emit_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ lookups, rescue, next-k
// RescueKminmerFunctor (CreateMdbg.hpp:4579-4637), one warp per read.  The decision only needs
// "median * 0.1f > 1", which is unchanged when abundances are clamped at 22, so the median comes from a
// 22-bin histogram held one bin per lane (Utils::compute_median, Commons.hpp:2973-2988).
// REMOTE = false: one context; the abundance-1 windows of a rescued read are flagged in the table itself.
// REMOTE = true : multi-rank; abundances come from the replicated table of solid k-min-mers (a.table; absent => 1),
//                 and the non-solid windows of a rescued read are appended as normalized vectors to a.out_vecs --
//                 their slots live on the owner rank, which flags them (rescue_flag_kernel) after the exchange.
template <bool REMOTE>
__global__ void __launch_bounds__(256) rescue_kernel(const RescueArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int k = (int)a.k;
    for (uint64_t r = warp; r < a.n_reads; r += n_warps) {
        const uint64_t b = a.offs[r], e = a.offs[r + 1];
        if (e - b < (uint64_t)k) continue;
        const uint64_t nw = e - b - k + 1;
        uint32_t bin = 0;                                  // lane v holds #windows with clamped abundance v
        bool any_solid = false;
        for (uint64_t i0 = 0; i0 < nw; i0 += 32) {
            const uint64_t i = i0 + lane;
            uint32_t c = 0;
            if (i < nw) {
                uint64_t h1, h2; bool rev;
                window_hash(a.mins + b + i, k, h1, h2, rev);
                const Slot* s = table_find(a.table, a.mask, h2, h1);
                const uint32_t cnt = s ? s->count : 1u;
                const uint32_t ab = cnt >= 2 ? cnt : 1u;   // solid = listed with abundance != 1 (CreateMdbg.hpp:4553-4556)
                any_solid |= ab >= 2;
                c = ab > 22u ? 22u : ab;
            }
            for (uint32_t v = 1; v <= 22; v++) {
                const uint32_t m = __ballot_sync(0xffffffffu, c == v);
                if (lane == v) bin += __popc(m);
            }
        }
        if (!__any_sync(0xffffffffu, any_solid)) continue;  // allAbundanceOne
        const uint32_t incl = warp_inclusive_scan(bin);      // #windows with clamped abundance <= lane
        const uint64_t h = nw / 2;
        // value at sorted index j = smallest v with incl[v] > j
        const uint32_t hi_v = __ffs(__ballot_sync(0xffffffffu, incl > h)) - 1;
        uint32_t median = hi_v;
        if ((nw & 1) == 0) {
            const uint32_t lo_v = __ffs(__ballot_sync(0xffffffffu, incl > h - 1)) - 1;
            median = (lo_v + hi_v) / 2;
        }
        const double cutoff = (double)((float)median * 0.1f);    // CreateMdbg.hpp:4612
        if (cutoff > 1) continue;
        if constexpr (!REMOTE) {
            for (uint64_t i = lane; i < nw; i += 32) {
                uint64_t h1, h2; bool rev;
                window_hash(a.mins + b + i, k, h1, h2, rev);
                Slot* s = table_find(a.table, a.mask, h2, h1);
                if (s && s->count < 2) s->flags = SLOT_RESCUED;
            }
        } else {
            for (uint64_t i0 = 0; i0 < nw; i0 += 32) {
                const uint64_t i = i0 + lane;
                bool emit = false, rev = false;
                if (i < nw) {
                    uint64_t h1, h2;
                    window_hash(a.mins + b + i, k, h1, h2, rev);
                    emit = table_find(a.table, a.mask, h2, h1) == nullptr;      // not solid: exactly one occurrence
                }
                const uint32_t m = __ballot_sync(0xffffffffu, emit);
                unsigned long long base = 0;
                if (lane == 0 && m) base = atomicAdd(a.out_cursor, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (emit) {
                    const uint64_t pos = base + __popc(m & ((1u << lane) - 1u));
                    const uint32_t* w = a.mins + b + i;
                    for (int j = 0; j < k; j++) a.out_vecs[pos * k + j] = rev ? w[k - 1 - j] : w[j];
                }
            }
        }
        if (lane == 0) atomicAdd(a.n_reads_rescued, 1ULL);
    }
}

void launch_rescue(const RescueArgs& a, cudaStream_t s) {
    if (a.n_reads == 0) return;
    uint64_t blocks = (a.n_reads + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (a.out_vecs) rescue_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(a);
    else rescue_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(a);
}

// ---- multi-rank rescue: bucket the collected vectors by owner rank (pass 1 counts, pass 2 scatters; the order
// inside a bucket is irrelevant), and on the owner flag the slot of every received vector
__global__ void __launch_bounds__(256) bucket_vecs_kernel(const BucketVecArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int k = (int)a.k;
    const uint32_t* w = a.vecs + i * (uint64_t)k;
    uint64_t h1, h2;
    murmur128_u32vec([&](int j) { return w[j]; }, k, h1, h2);
    const uint32_t dst = owner_of(h1, a.n_ranks);
    const unsigned long long slot = atomicAdd(&a.bucket_count[dst], 1ULL);
    if (a.pass == 2) {
        const uint64_t pos = a.bucket_base[dst] + slot;
        for (int j = 0; j < k; j++) a.out_vecs[pos * k + j] = w[j];
    }
}

void launch_bucket_vecs(const BucketVecArgs& a, cudaStream_t s) {
    if (a.n == 0) return;
    bucket_vecs_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) rescue_flag_kernel(const uint32_t* vecs, uint64_t n, uint32_t k, Slot* table,
                                                          uint64_t mask, unsigned long long* n_missing) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* w = vecs + i * (uint64_t)k;
    uint64_t h1, h2;
    murmur128_u32vec([&](int j) { return w[j]; }, (int)k, h1, h2);
    Slot* sl = table_find(table, mask, h2, h1);
    if (!sl) { atomicAdd(n_missing, 1ULL); return; }            // every occurrence is in its owner's merged table
    if (sl->count < 2) sl->flags = SLOT_RESCUED;
}

void launch_rescue_flag(const uint32_t* vecs, uint64_t n, uint32_t k, Slot* table, uint64_t mask,
                        unsigned long long* n_missing, cudaStream_t s) {
    if (n == 0) return;
    rescue_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(vecs, n, k, table, mask, n_missing);
}

// insert-or-assign into the previous-k lookup table (value in `count`)
__device__ __forceinline__ bool prev_put(Slot* table, uint64_t mask, uint64_t lo, uint64_t hi, uint32_t value) {
    uint64_t idx = slot_of(lo, mask);
    for (uint64_t probe = 0; probe <= mask && probe < 4096; probe++) {
        Slot* s = table + idx;
        uint64_t clo, chi;
        load_key(s, clo, chi);
        // SLOT_RESCUED = "listed whatever its count": a pair loaded or patched by the host is never hidden by the
        // lookup-time abundance filter of a table that used to be a count table
        if (clo == lo && chi == hi && lo != 0 && hi != 0) { s->count = value; s->flags = SLOT_RESCUED; return true; }
        if (clo == 0 || chi == 0) {
            uint64_t olo, ohi;
            cas_key(s, lo, hi, olo, ohi);
            if ((olo == 0 && ohi == 0) || (olo == lo && ohi == hi)) { s->count = value; s->flags = SLOT_RESCUED; return true; }
        }
        idx = next_slot(idx, mask);
    }
    return false;
}

__global__ void __launch_bounds__(256) prev_from_table_kernel(const PrevFromTableArgs a) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.capacity;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const Slot sl = a.table[i];
        if ((sl.lo | sl.hi) == 0) continue;
        if (sl.count >= a.min_count || (sl.flags & SLOT_RESCUED))
            if (!prev_put(a.prev, a.prev_mask, sl.lo, sl.hi, sl.count)) atomicExch(a.full_flag, 1u);
    }
}

void launch_prev_from_table(const PrevFromTableArgs& a, cudaStream_t s) {
    uint64_t blocks = (a.capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    prev_from_table_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) prev_load_kernel(const PrevLoadArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    if (!prev_put(a.prev, a.prev_mask, a.hashes[2 * i], a.hashes[2 * i + 1], a.abund[i])) atomicExch(a.full_flag, 1u);
}

void launch_prev_load(const PrevLoadArgs& a, cudaStream_t s) {
    if (a.n == 0) return;
    prev_load_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
}

// k >= firstK+1: abundance of a k-min-mer = min over its two (k-1)-min-mers of the previous-k table,
// absent (or 0) => 1; kept (insert-if-absent, value = that abundance) only when > 1.
//   KminmerCounter::getRefinedAbundance  CreateMdbg.hpp:3933-4005  (k = firstK+1)
//   IndexKminmerFunctor                  CreateMdbg.hpp:988-1010, 1240-1265, 1268-1464  (k >= firstK+2)
// One thread per read POSITION: the (k-1)-min-mer starting there is hashed and looked up once; the k-min-mer starting
// at the same position takes its second value from the next lane (__shfl_down).  A warp covers 31 windows with 32
// positions, so every lookup but one per warp is used twice -- half the hashes and random reads of one thread per
// window.  prev_min_count filters the previous table at lookup time (an entry below it that is not rescued counts
// as absent), which lets the previous-k table BE the table of the previous pass, unfiltered and uncopied.
// WARP: bookkeeping per warp instead of per block (no barrier; PassAux) -- see insert_warp_kernel
template <bool WARP>
__global__ void __launch_bounds__(WARP ? 128 : 256) next_k_kernel(const NextKArgs a) {
    __shared__ BlockPass st;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t abandon = 0;
    if (WARP) abandon = warp_abandon(a.aux, (uint32_t)warp);
    else block_begin(st, a.full_flag);
    const uint64_t g = a.g_lo + warp * 31 + lane;
    const int k = (int)a.k;
    const uint32_t rem = g < a.g_hi ? (uint32_t)a.rem[g] : 0u;
    const uint32_t* w = a.mins + g;
    uint32_t v = 1;
    if ((int)rem >= k - 1) {
        uint64_t h1, h2; bool rev;
        window_hash(w, k - 1, h1, h2, rev);
        const Slot* s = table_find(const_cast<Slot*>(a.prev), a.prev_mask, h2, h1);
        if (s && (s->count >= a.prev_min_count || (s->flags & SLOT_RESCUED))) v = s->count;
        if (v == 0) v = 1;
    }
    const uint32_t v_next = __shfl_down_sync(0xffffffffu, v, 1);
    const bool active = lane < 31 && (int)rem >= k;              // (rem >= k implies the next position has rem >= k - 1)
    int status = 1;
    uint32_t out = 1;
    const bool go = WARP ? abandon == 0 : block_ready(st);       // false: the pass is being abandoned (it is redone as a whole)
    if (active) {
        const uint32_t ab = v < v_next ? v : v_next;
        if (ab > 1) {
            out = ab;
            if (go) {
                uint64_t h1, h2; bool rev;
                window_hash(w, k, h1, h2, rev);
                // insert-if-absent: the value is a function of the key, so concurrent inserters write the same number
                status = table_put(a.table, a.mask, h2, h1, ab, g | (rev ? REF_REV : 0ULL));
            }
        }
    }
    // value of the k-min-mer starting at g, as the NEXT pass would look it up (absent => 1): see next_k_stream_kernel
    if (a.val_out && lane < 31 && g < a.g_hi) a.val_out[g] = out;
    if (WARP) warp_claims(a.aux, a.aux_shards, (uint32_t)warp, status, go && active, a.claim_limit);
    else block_claims(st, status, go && active, a.claims, a.claim_limit, a.full_flag);
}

// The same pass without a single lookup.  The value a pass stores for a k-min-mer is a function of the key, and
// every window of every read was offered to it -- so what the NEXT pass would find in the table for the k-min-mer
// starting at position g is exactly what THIS pass computed at g (or "absent => 1" when that was <= 1).  When the
// previous-k table is nothing but the previous pass's table (no host patches, same store), pass k + 1 therefore
// reads val[g] and val[g + 1] -- two coalesced loads -- instead of hashing and probing two (k)-min-mers; the only
// random access left per window is the insert into the new table.  Identical tables, by induction on k.
template <bool WARP>
__global__ void __launch_bounds__(WARP ? 128 : 256) next_k_stream_kernel(const NextKArgs a) {
    __shared__ BlockPass st;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t abandon = 0;
    if (WARP) abandon = warp_abandon(a.aux, (uint32_t)(t >> 5));
    else block_begin(st, a.full_flag);
    const uint64_t g = a.g_lo + t;
    const int k = (int)a.k;
    const bool in_range = g < a.g_hi;
    const bool active = in_range && (int)a.rem[g] >= k;
    int status = 1;
    uint32_t out = 1;
    uint64_t h1 = 0, h2 = 0;
    bool rev = true, put = false;
    if (active) {
        const uint32_t v = a.val_in[g], v_next = a.val_in[g + 1];
        const uint32_t ab = v < v_next ? v : v_next;
        if (ab > 1) {
            out = ab;
            put = true;
            window_hash(a.mins + g, k, h1, h2, rev);
        }
    }
    if (in_range) a.val_out[g] = out;
    const bool go = WARP ? abandon == 0 : block_ready(st);       // false: the pass is being abandoned (it is redone as a whole)
    if (go && put) status = table_put(a.table, a.mask, h2, h1, out, g | (rev ? REF_REV : 0ULL));
    if (WARP) warp_claims(a.aux, a.aux_shards, (uint32_t)(t >> 5), status, go && active, a.claim_limit);
    else block_claims(st, status, go && active, a.claims, a.claim_limit, a.full_flag);
}

void launch_next_k(const NextKArgs& a, cudaStream_t s) {
    if (a.g_hi <= a.g_lo) return;
    const bool warp_form = pass_variant() != 0 && a.aux;
    const uint64_t n = a.g_hi - a.g_lo;
    if (a.val_in) {
        if (warp_form) next_k_stream_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
        else next_k_stream_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
    } else {
        const uint64_t n_warps = (n + 30) / 31;
        if (warp_form) next_k_kernel<true><<<(unsigned)((n_warps + 3) / 4), 128, 0, s>>>(a);
        else next_k_kernel<false><<<(unsigned)((n_warps + 7) / 8), 256, 0, s>>>(a);
    }
    if (warp_form) pass_fold_kernel<<<1, 64, 0, s>>>(a.aux, a.claims, a.full_flag);
}

// ------------------------------------------------------------------ edge keys of the node set (row F1)
// CreateMdbg::EdgeIndexer (src/graph/CreateMdbg.hpp:4010-4232): every node -- an emitted table entry, in the
// normalized orientation it is dumped in -- contributes the hash128 of its normalized (k-1)-prefix and
// (k-1)-suffix (partitionNode :4106-4120); the reference sorts and dereplicates them on disk, here they go into a
// second open-addressing table used as a set (value 1), one thread per table slot.
__global__ void __launch_bounds__(256) edge_insert_kernel(const EdgeArgs a) {
    const int k = (int)a.k, km = k - 1;
    // with a node list (slots of the emitted entries, unitig_nodes_kernel) the threads visit the nodes only; without
    // one they scan the whole table
    const uint64_t n_items = a.node_slot ? a.n_nodes : a.capacity;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const Slot sl = a.table[a.node_slot ? a.node_slot[i] : i];
        if ((sl.lo | sl.hi) == 0) continue;
        if (!(sl.count >= a.min_count || (sl.flags & SLOT_RESCUED))) continue;
        for (int side = 0; side < 2; side++) {            // 0: prefix = elements 0 .. k-2, 1: suffix = 1 .. k-1
            bool rev = true;                              // KmerVec::normalize on the k-1 elements
            for (int j = 0; j < km / 2; j++) {
                const uint32_t x = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + j);
                const uint32_t y = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - j);
                if (x != y) { rev = x > y; break; }
            }
            uint64_t h1, h2;
            if (rev) murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - t); }, km, h1, h2);
            else murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + t); }, km, h1, h2);
            if (!table_put(a.edges, a.edge_mask, h2, h1, 1u, 0ULL)) atomicExch(a.full_flag, 1u);
        }
    }
}

void launch_edge_insert(const EdgeArgs& a, cudaStream_t s) {
    uint64_t blocks = ((a.node_slot ? a.n_nodes : a.capacity) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks == 0) return;
    edge_insert_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

// Edge VALUES (CreateMdbg::indexEdge + successorExists, src/graph/CreateMdbg.cpp:1277-1500) in order-free form.
// Every node offers (extending minimizer, isReversed, isPrefix) to the key of its normalized suffix (first element,
// isPrefix = 0) and of its normalized prefix (last element, isPrefix = 1).  Offers to one key fall into two
// orientation classes, A = {isReversed == isPrefix} and B = {isReversed != isPrefix}; a palindromic key has only
// class A.  Upstream records the first offer of a class and marks hasMultipleSuccessors on any further one (which
// offer is "first" depends on its thread arrival order; its consumers use the recorded minimizer only while the
// mark is clear), so the order-free value of (key, class) is: empty, ONE offer with its fields, or "two or more".
//   vals[2 * slot + class]: bit 63 valid, bit 34 multi, bit 33 isPrefix, bit 32 isReversed, bits 0-31 minimizer
__global__ void __launch_bounds__(256) edge_values_kernel(const EdgeArgs a) {
    const int k = (int)a.k, km = k - 1;
    // with a node list (slots of the emitted entries, unitig_nodes_kernel) the threads visit the nodes only; without
    // one they scan the whole table
    const uint64_t n_items = a.node_slot ? a.n_nodes : a.capacity;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const Slot sl = a.table[a.node_slot ? a.node_slot[i] : i];
        if ((sl.lo | sl.hi) == 0) continue;
        if (!(sl.count >= a.min_count || (sl.flags & SLOT_RESCUED))) continue;
        for (int side = 0; side < 2; side++) {            // 0: prefix (isPrefix = 1), 1: suffix (isPrefix = 0)
            bool rev = true, pal = true;
            for (int j = 0; j < km / 2; j++) {
                const uint32_t x = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + j);
                const uint32_t y = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - j);
                if (x != y) { rev = x > y; pal = false; break; }
            }
            uint64_t h1, h2;
            if (rev) murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - t); }, km, h1, h2);
            else murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + t); }, km, h1, h2);
            const Slot* es = table_find(a.edges, a.edge_mask, h2, h1);
            if (!es) { atomicExch(a.full_flag, 1u); continue; }        // every key was inserted by edge_insert_kernel
            const uint32_t pre = side ? 0u : 1u, r = rev ? 1u : 0u;
            const uint32_t cls = pal ? 0u : (r == pre ? 0u : 1u);
            const uint32_t ext = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side ? 0 : k - 1);
            const unsigned long long one = EDGE_VALID | ((unsigned long long)pre << 33) | ((unsigned long long)r << 32) | ext;
            unsigned long long* w = a.edge_vals + 2 * (uint64_t)(es - a.edges) + cls;
            if (atomicCAS(w, 0ULL, one) != 0ULL) atomicExch(w, EDGE_VALID | EDGE_MULTI);
        }
    }
}

void launch_edge_values(const EdgeArgs& a, cudaStream_t s) {
    uint64_t blocks = ((a.node_slot ? a.n_nodes : a.capacity) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks == 0) return;
    edge_values_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

// keys and their two value words, in one (unspecified) order
__global__ void __launch_bounds__(256) edge_emit_kernel(const Slot* edges, const unsigned long long* vals, uint64_t capacity,
                                                        uint64_t* out_hashes, unsigned long long* out_vals,
                                                        unsigned long long* cursor) {
    __shared__ uint32_t wcnt[8];
    __shared__ unsigned long long tile_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = (capacity + 255) / 256;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {       // one cursor atomic per 256-slot tile
        const uint64_t i = tile * 256 + threadIdx.x;
        bool take = false;
        uint64_t lo = 0, hi = 0;
        if (i < capacity) { lo = edges[i].lo; hi = edges[i].hi; take = (lo | hi) != 0; }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = wcnt[w]; wcnt[w] = tot; tot += c; }
            tile_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        if (take) {
            const uint64_t pos = tile_base + wcnt[warp] + __popc(m & ((1u << lane) - 1u));
            out_hashes[2 * pos] = lo;
            out_hashes[2 * pos + 1] = hi;
            unsigned long long a0 = vals[2 * i], b0 = vals[2 * i + 1];
            if (a0 & EDGE_MULTI) a0 = EDGE_VALID | EDGE_MULTI;        // canonical: nothing but the mark
            if (b0 & EDGE_MULTI) b0 = EDGE_VALID | EDGE_MULTI;
            out_vals[2 * pos] = a0;
            out_vals[2 * pos + 1] = b0;
        }
        __syncthreads();
    }
}

void launch_edge_emit(const Slot* edges, const unsigned long long* vals, uint64_t capacity, uint64_t* out_hashes,
                      unsigned long long* out_vals, unsigned long long* cursor, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    edge_emit_kernel<<<(unsigned)blocks, 256, 0, s>>>(edges, vals, capacity, out_hashes, out_vals, cursor);
}

// multi-rank edge keys: bucket the local distinct keys by owner rank (pass 1 counts, pass 2 scatters), and on the
// owner insert what arrived into a fresh set table -- a key produced by nodes of several ranks is kept once
__global__ void __launch_bounds__(256) bucket_keys_kernel(const BucketKeyArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t W = a.rec_words;                       // 2: bare keys {lo, hi}; 3: offers {lo, hi, value word}
    const uint64_t* rec = a.keys + (uint64_t)W * i;
    const uint32_t dst = owner_of(rec[1], a.n_ranks);
    const unsigned long long slot = atomicAdd(&a.bucket_count[dst], 1ULL);
    if (a.pass == 2) {
        const uint64_t pos = a.bucket_base[dst] + slot;
        for (uint32_t w = 0; w < W; w++) a.out_keys[(uint64_t)W * pos + w] = rec[w];
    }
}

void launch_bucket_keys(const BucketKeyArgs& a, cudaStream_t s) {
    if (a.n == 0) return;
    bucket_keys_kernel<<<(unsigned)((a.n + 255) / 256), 256, 0, s>>>(a);
}

__global__ void __launch_bounds__(256) insert_keys_kernel(const uint64_t* keys, uint64_t n, Slot* table, uint64_t mask,
                                                          uint32_t* full_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!table_put(table, mask, keys[2 * i], keys[2 * i + 1], 1u, 0ULL)) atomicExch(full_flag, 1u);
}

// multi-rank edge values.  edge_offers_kernel: the two offers of every owned node as records {key lo, key hi, value
// word | class << 36}; after the owner exchange edge_apply_offers_kernel folds each record into the class word of
// its key exactly as edge_values_kernel does locally.
constexpr unsigned long long EDGE_CLASS_B = 1ULL << 36;

__global__ void __launch_bounds__(256) edge_offers_kernel(const EdgeArgs a, uint64_t* out_recs, unsigned long long* cursor) {
    const int k = (int)a.k, km = k - 1;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_rounds = (a.capacity + stride - 1) / stride;
    for (uint64_t round = 0; round < n_rounds; round++) {
        const uint64_t i = round * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool take = false;
        Slot sl{};
        if (i < a.capacity) {
            sl = a.table[i];
            take = (sl.lo | sl.hi) != 0 && (sl.count >= a.min_count || (sl.flags & SLOT_RESCUED));
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        unsigned long long base = 0;
        if (lane == 0 && m) base = atomicAdd(cursor, 2ULL * __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!take) continue;
        const uint64_t pos0 = base + 2ULL * __popc(m & ((1u << lane) - 1u));
        for (int side = 0; side < 2; side++) {
            bool rev = true, pal = true;
            for (int j = 0; j < km / 2; j++) {
                const uint32_t x = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + j);
                const uint32_t y = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - j);
                if (x != y) { rev = x > y; pal = false; break; }
            }
            uint64_t h1, h2;
            if (rev) murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + km - 1 - t); }, km, h1, h2);
            else murmur128_u32vec([&](int t) { return vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side + t); }, km, h1, h2);
            const uint32_t pre = side ? 0u : 1u, r = rev ? 1u : 0u;
            const uint32_t cls = pal ? 0u : (r == pre ? 0u : 1u);
            const uint32_t ext = vec_elem(a.mins, a.foreign_vecs, sl.ref, k, side ? 0 : k - 1);
            uint64_t* rec = out_recs + 3 * (pos0 + side);
            rec[0] = h2;
            rec[1] = h1;
            rec[2] = EDGE_VALID | ((unsigned long long)pre << 33) | ((unsigned long long)r << 32) | ext | (cls ? EDGE_CLASS_B : 0ULL);
        }
    }
}

void launch_edge_offers(const EdgeArgs& a, uint64_t* out_recs, unsigned long long* cursor, cudaStream_t s) {
    uint64_t blocks = (a.capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    edge_offers_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, out_recs, cursor);
}

__global__ void __launch_bounds__(256) edge_apply_offers_kernel(const uint64_t* recs, uint64_t n, const Slot* edges, uint64_t mask,
                                                                unsigned long long* vals, uint32_t* full_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* rec = recs + 3 * i;
    const Slot* es = table_find(const_cast<Slot*>(edges), mask, rec[0], rec[1]);
    if (!es) { atomicExch(full_flag, 1u); return; }                  // the owner holds every key its ranks produced
    const unsigned long long one = rec[2] & ~EDGE_CLASS_B;
    unsigned long long* w = vals + 2 * (uint64_t)(es - edges) + ((rec[2] & EDGE_CLASS_B) ? 1 : 0);
    if (atomicCAS(w, 0ULL, one) != 0ULL) atomicExch(w, EDGE_VALID | EDGE_MULTI);
}

void launch_edge_apply_offers(const uint64_t* recs, uint64_t n, const Slot* edges, uint64_t mask, unsigned long long* vals,
                              uint32_t* full_flag, cudaStream_t s) {
    if (n == 0) return;
    edge_apply_offers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(recs, n, edges, mask, vals, full_flag);
}

void launch_insert_keys(const uint64_t* keys, uint64_t n, Slot* table, uint64_t mask, uint32_t* full_flag, cudaStream_t s) {
    if (n == 0) return;
    insert_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys, n, table, mask, full_flag);
}

// ------------------------------------------------------------------ multi-GPU pack by owner rank
constexpr int PACK_MAX_RANKS = 64;

__global__ void __launch_bounds__(256) table_pack_kernel(const PackArgs a) {
    __shared__ uint32_t wcnt[8][PACK_MAX_RANKS];          // per-warp counts -> exclusive offsets inside the tile
    __shared__ unsigned long long bbase[PACK_MAX_RANKS];  // this tile's reservation in every bucket
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t R = a.n_ranks;
    const uint64_t n_tiles = (a.capacity + 255) / 256;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i = tile * 256 + threadIdx.x;
        bool take = false;
        uint32_t dst = 0, count = 0, my_prefix = 0;
        uint64_t ref = 0;
        if (i < a.capacity) {
            const Slot sl = a.table[i];
            take = (sl.lo | sl.hi) != 0;
            dst = owner_of(sl.hi, R);
            count = sl.count;
            ref = sl.ref;
        }
        for (uint32_t d = 0; d < R; d++) {
            const uint32_t m = __ballot_sync(0xffffffffu, take && dst == d);
            if (lane == 0) wcnt[warp][d] = __popc(m);
            if (take && dst == d) my_prefix = __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < R) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = tot; tot += c; }
            bbase[threadIdx.x] = tot ? atomicAdd(&a.bucket_count[threadIdx.x], (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        if (a.pass == 2 && take) {
            const uint64_t pos = a.bucket_base[dst] + bbase[dst] + wcnt[warp][dst] + my_prefix;
            a.out_counts[pos] = count;
            for (int j = 0; j < (int)a.k; j++)
                a.out_vecs[pos * a.k + j] = vec_elem(a.mins, a.foreign_vecs, ref, (int)a.k, j);
        }
        __syncthreads();
    }
}

void launch_table_pack(const PackArgs& a, cudaStream_t s) {
    uint64_t blocks = (a.capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    table_pack_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ multi-GPU merge, keys only
// For tables whose k-min-mer VECTORS are not wanted on the owner (the per-k tables of a multi-k loop: the next pass
// needs no table at all, see next_k_stream_kernel): records of 24 bytes {hash lo, hash hi, count | 0} instead of
// 4 k + 4 bytes, no gather of vectors on the sender, no re-hash on the owner.
// ONE pass: every destination has a fixed-capacity region of the send buffer (region_cap records; the host sizes it
// for the table's number of distinct keys, so it cannot overflow).  Per 256-slot tile the records are counted per
// destination in shared memory, the tile reserves its ranges with one global atomicAdd per destination, and the
// records are written behind them.
__global__ void __launch_bounds__(256) table_pack_hashes_kernel(const PackArgs a, uint64_t* out_recs, uint64_t region_cap) {
    __shared__ uint32_t cnt[PACK_MAX_RANKS];
    __shared__ unsigned long long base[PACK_MAX_RANKS];
    const uint32_t R = a.n_ranks;
    const uint64_t n_tiles = (a.capacity + 255) / 256;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < R) cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t i = tile * 256 + threadIdx.x;
        bool take = false;
        uint32_t dst = 0, count = 0, my = 0;
        uint64_t lo = 0, hi = 0;
        if (i < a.capacity) {
            const uint4 q0 = reinterpret_cast<const uint4*>(a.table + i)[0];
            lo = (uint64_t)q0.x | ((uint64_t)q0.y << 32);
            hi = (uint64_t)q0.z | ((uint64_t)q0.w << 32);
            take = (lo | hi) != 0;
            if (take) { dst = owner_of(hi, R); count = a.table[i].count; my = atomicAdd(&cnt[dst], 1u); }
        }
        __syncthreads();
        if (threadIdx.x < R) base[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&a.bucket_count[threadIdx.x], (unsigned long long)cnt[threadIdx.x]) : 0ULL;
        __syncthreads();
        if (take && base[dst] + my < region_cap) {
            const uint64_t pos = (uint64_t)dst * region_cap + base[dst] + my;
            out_recs[3 * pos] = lo;
            out_recs[3 * pos + 1] = hi;
            out_recs[3 * pos + 2] = count;
        }
        __syncthreads();
    }
}

void launch_table_pack_hashes(const PackArgs& a, uint64_t* out_recs, uint64_t region_cap, cudaStream_t s) {
    uint64_t blocks = (a.capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    table_pack_hashes_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, out_recs, region_cap);
}

__global__ void __launch_bounds__(256) insert_hash_recs_kernel(const uint64_t* recs, uint64_t n, Slot* table, uint64_t mask,
                                                               uint32_t assign, uint32_t* full_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t lo = recs[3 * i], hi = recs[3 * i + 1];
    const uint32_t count = (uint32_t)recs[3 * i + 2];
    const int ok = assign ? table_put(table, mask, lo, hi, count, REF_NONE) : table_add(table, mask, lo, hi, count, REF_NONE);
    if (!ok) atomicExch(full_flag, 1u);
}

void launch_insert_hash_recs(const uint64_t* recs, uint64_t n, Slot* table, uint64_t mask, uint32_t assign, uint32_t* full_flag,
                             cudaStream_t s) {
    if (n == 0) return;
    insert_hash_recs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(recs, n, table, mask, assign, full_flag);
}

// ------------------------------------------------------------------ postings: k-min-mer -> (read, window) lists
// Row (f)4: the inverted index the ONT correction stage builds over the low-density reads
// (ReadCorrection::IndexReadsFunctor, src/readSelection/ReadCorrection.hpp:3064-3130: for every window i of every
// read, _kminmer_to_readIndex[vec] gets (readIndex, positionIndex = i) when vec is a known k-min-mer).  Here the
// count table IS the key set and its abundances are the list lengths: one scan over the slots gives every list its
// place in one postings array (CSR), and one pass over the windows fills it.  Order inside a list is unspecified
// (upstream's is the arrival order of its OpenMP threads).
__global__ void __launch_bounds__(256) posting_counts_kernel(const Slot* table, uint64_t capacity, uint32_t min_count,
                                                             uint32_t* counts, uint32_t* flags) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (uint64_t)gridDim.x * blockDim.x) {
        const Slot sl = table[i];
        const bool take = (sl.lo | sl.hi) != 0 && (sl.count >= min_count || (sl.flags & SLOT_RESCUED));
        counts[i] = take ? sl.count : 0u;
        flags[i] = take ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) posting_keys_kernel(const Slot* table, uint64_t capacity, const uint32_t* flags,
                                                           const uint64_t* key_index, const uint64_t* post_off,
                                                           uint64_t* out_hashes, uint64_t* out_offsets) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!flags[i]) continue;
        const uint64_t j = key_index[i];
        out_hashes[2 * j] = table[i].lo;
        out_hashes[2 * j + 1] = table[i].hi;
        out_offsets[j] = post_off[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out_offsets[key_index[capacity]] = post_off[capacity];
}

__global__ void __launch_bounds__(256) read_of_kernel(const uint64_t* offs, uint64_t n_reads, uint32_t* read_of) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < n_reads; r += n_warps)
        for (uint64_t g = offs[r] + lane; g < offs[r + 1]; g += 32) read_of[g] = (uint32_t)r;
}

__global__ void __launch_bounds__(256) posting_fill_kernel(const PostingArgs a) {
    const uint64_t g = a.g_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.g_hi) return;
    const int k = (int)a.k;
    if ((int)a.rem[g] < k) return;
    uint64_t h1, h2; bool rev;
    window_hash(a.mins + g, k, h1, h2, rev);
    const Slot* s = table_find(const_cast<Slot*>(a.table), a.mask, h2, h1);
    if (!s) return;
    const uint64_t i = (uint64_t)(s - a.table);
    if (!a.flags[i]) return;
    const uint64_t p = a.post_off[i] + atomicAdd(&a.cursors[i], 1u);
    const uint32_t r = a.read_of[g];
    a.out_reads[p] = r;
    a.out_windows[p] = (uint32_t)(g - a.offs[r]);
}

void launch_posting_counts(const Slot* table, uint64_t capacity, uint32_t min_count, uint32_t* counts, uint32_t* flags, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    posting_counts_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, min_count, counts, flags);
}
void launch_posting_keys(const Slot* table, uint64_t capacity, const uint32_t* flags, const uint64_t* key_index,
                         const uint64_t* post_off, uint64_t* out_hashes, uint64_t* out_offsets, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    posting_keys_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, flags, key_index, post_off, out_hashes, out_offsets);
}
void launch_read_of(const uint64_t* offs, uint64_t n_reads, uint32_t* read_of, cudaStream_t s) {
    if (n_reads == 0) return;
    uint64_t blocks = (n_reads + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    read_of_kernel<<<(unsigned)blocks, 256, 0, s>>>(offs, n_reads, read_of);
}
void launch_posting_fill(const PostingArgs& a, cudaStream_t s) {
    if (a.g_hi <= a.g_lo) return;
    posting_fill_kernel<<<(unsigned)((a.g_hi - a.g_lo + 255) / 256), 256, 0, s>>>(a);
}

}  // namespace mdbg
