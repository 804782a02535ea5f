// unitig.cu -- unitig nodes of the node set on the device (third step).
//
// Replaces (reference, paths relative to the metaMDBG tree):
//   CreateMdbg::computeUnitigNodes                     src/graph/CreateMdbg.cpp:1521-1598
//   ComputeUnitigFunctor::computeUnitigNode2           src/graph/CreateMdbg.hpp:2513-2916
//   CreateMdbg::getNbSuccessors / getNbPredecessors    src/graph/CreateMdbg.cpp:1902-2135, 2227-2380
//   CreateMdbg::computeDeterministicUnitigs            src/graph/CreateMdbg.cpp:1001-1043 (normalize + hash; the sort is the host's)
//
// The reference walks from every not yet unitigged node forwards and backwards while "single successor whose single
// predecessor exists", under one critical section.  Here the walk is a list ranking over ORIENTED nodes (x = 2 * node
// + reversed): x -> y is a unitig link iff the key of x's (k-1)-suffix holds exactly one offer in the orientation class
// of that suffix (the single successor y = suffix + recorded minimizer) AND exactly one in the opposite class (y's
// single predecessor, which is then x); a palindromic key has one class.  Links are symmetric under reversal
// (x -> y  <=>  rev(y) -> rev(x)), every oriented node has at most one link in and one out, so the components are
// simple paths and cycles and every unitig appears twice (once per strand).  Pointer jumping gives every oriented node
// its path head and rank; what is left after ceil(log2(2n)) + 2 rounds lies on cycles, which are cut at the node with the
// smallest hash128 (the reference's rotation rule for circular unitigs) and ranked again.
#include "common.cuh"
#include "engine.cuh"
#include "table.cuh"

namespace mdbg {

constexpr uint32_t U_NONE = 0xFFFFFFFFu;
constexpr uint32_t U_HEAD = 0x80000000u;         // in the low word of a (pointer, rank) pair: the pointer is the path head

// ------------------------------------------------------------------ node ids: emitted slots -> 0 .. n-1
__global__ void __launch_bounds__(256) unitig_nodes_kernel(const Slot* table, uint64_t capacity, uint32_t min_count,
                                                           uint32_t* slot_node, uint32_t* node_slot,
                                                           unsigned long long* cursor) {
    __shared__ uint32_t wcnt[8];
    __shared__ unsigned long long tile_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_tiles = (capacity + 255) / 256;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i = tile * 256 + threadIdx.x;
        bool take = false;
        if (i < capacity) {
            const Slot sl = table[i];
            take = (sl.lo | sl.hi) != 0 && (sl.count >= min_count || (sl.flags & SLOT_RESCUED));
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = wcnt[w]; wcnt[w] = tot; tot += c; }
            tile_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        if (i < capacity) {
            uint32_t id = U_NONE;
            if (take) {
                id = (uint32_t)(tile_base + wcnt[warp] + __popc(m & ((1u << lane) - 1u)));
                node_slot[id] = (uint32_t)i;
            }
            if (slot_node) slot_node[i] = id;
        }
        __syncthreads();
    }
}

void launch_unitig_nodes(const Slot* table, uint64_t capacity, uint32_t min_count, uint32_t* slot_node, uint32_t* node_slot,
                         unsigned long long* cursor, cudaStream_t s) {
    uint64_t blocks = (capacity + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    unitig_nodes_kernel<<<(unsigned)blocks, 256, 0, s>>>(table, capacity, min_count, slot_node, node_slot, cursor);
}

// ------------------------------------------------------------------ links
// element j of oriented node x (bit 0 of x: the stored, normalized vector read backwards)
__device__ __forceinline__ uint32_t onode_elem(const UnitigArgs& a, uint64_t ref, uint32_t d, int j) {
    return vec_elem(a.mins, a.foreign_vecs, ref, (int)a.k, d ? (int)a.k - 1 - j : j);
}

__global__ void __launch_bounds__(256) unitig_link_kernel(const UnitigArgs a) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= 2 * a.n_nodes) return;
    const int k = (int)a.k, km = k - 1;
    const uint32_t d = x & 1;
    const uint64_t ref = a.table[a.node_slot[x >> 1]].ref;
    // the (k-1)-suffix S of x as written: S(j) = x(j + 1); KmerVec::normalize + isPalindrome on it
    bool rev = true, pal = true;
    for (int j = 0; j < km / 2; j++) {
        const uint32_t p = onode_elem(a, ref, d, 1 + j), q = onode_elem(a, ref, d, 1 + km - 1 - j);
        if (p != q) { rev = p > q; pal = false; break; }
    }
    uint64_t h1, h2;
    if (rev) murmur128_u32vec([&](int t) { return onode_elem(a, ref, d, 1 + km - 1 - t); }, km, h1, h2);
    else murmur128_u32vec([&](int t) { return onode_elem(a, ref, d, 1 + t); }, km, h1, h2);
    uint32_t next = U_NONE;
    const Slot* es = table_find(const_cast<Slot*>(a.edges), a.edge_mask, h2, h1);
    if (!es) {
        atomicExch(a.error_flag, 1u);                                // every node's keys are in the edge set
    } else {
        const unsigned long long* w = a.edge_vals + 2 * (uint64_t)(es - a.edges);
        // getNbSuccessors: offers whose oriented (k-1)-mer equals S lie in class A when S is the reversed key, in B otherwise
        const uint32_t cls = pal ? 0u : (rev ? 0u : 1u);
        const unsigned long long ws = w[cls], wp = w[cls ^ 1u];
        const bool single_s = (ws & EDGE_VALID) && !(ws & EDGE_MULTI);
        const bool single_p = pal || ((wp & EDGE_VALID) && !(wp & EDGE_MULTI));     // getNbPredecessors of the successor
        if (single_s && single_p) {
            const uint32_t ext = (uint32_t)ws;                       // successor y = S + recorded minimizer
            auto y = [&](int j) { return j < km ? onode_elem(a, ref, d, 1 + j) : ext; };
            bool yrev = true;
            for (int j = 0; j < k / 2; j++) {
                const uint32_t p = y(j), q = y(k - 1 - j);
                if (p != q) { yrev = p > q; break; }
            }
            uint64_t g1, g2;
            if (yrev) murmur128_u32vec([&](int t) { return y(k - 1 - t); }, k, g1, g2);
            else murmur128_u32vec([&](int t) { return y(t); }, k, g1, g2);
            const Slot* ns = table_find(const_cast<Slot*>(a.table), a.mask, g2, g1);
            const uint32_t id = ns ? a.slot_node[ns - a.table] : U_NONE;
            if (id == U_NONE) atomicExch(a.error_flag, 2u);          // an offer always comes from a node of the set
            else next = 2 * id + (yrev ? 1u : 0u);
        }
    }
    a.next[x] = next;
}

void launch_unitig_link(const UnitigArgs& a, cudaStream_t s) {
    if (!a.n_nodes) return;
    unitig_link_kernel<<<(unsigned)((2 * (uint64_t)a.n_nodes + 255) / 256), 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ list ranking by pointer jumping
// pair[x] = {low word: predecessor pointer, or U_HEAD | head once resolved; high word: distance to it}.  Pairs are read
// and written as single 64-bit words, so a jump may read a pair another thread already advanced in the same launch:
// any pair ever stored is a true (ancestor, distance) statement, which is all a jump needs.
__device__ __forceinline__ unsigned long long ld_pair(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__global__ void __launch_bounds__(256) unitig_rank_init_kernel(const uint32_t* next, uint32_t n2, unsigned long long* pair,
                                                               uint32_t* len) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n2) return;
    const uint32_t nr = next[x ^ 1u];                                // x's predecessor = reverse of the successor of rev(x)
    pair[x] = nr == U_NONE ? (unsigned long long)(U_HEAD | x) : ((1ULL << 32) | (unsigned long long)(nr ^ 1u));
    len[x] = 0;
}

void launch_unitig_rank_init(const uint32_t* next, uint32_t n2, unsigned long long* pair, uint32_t* len, cudaStream_t s) {
    if (!n2) return;
    unitig_rank_init_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(next, n2, pair, len);
}

// up to `steps` jumps per thread; *n_open counts the nodes still unresolved afterwards
__global__ void __launch_bounds__(256) unitig_jump_kernel(unsigned long long* pair, uint32_t n2, int steps,
                                                          unsigned long long* n_open) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    bool open = false;
    if (x < n2) {
        unsigned long long p = ld_pair(pair + x);
        bool changed = false;
        for (int it = 0; it < steps && !((uint32_t)p & U_HEAD); it++) {
            const unsigned long long q = ld_pair(pair + (uint32_t)p);
            p = (unsigned long long)(uint32_t)q | (((p >> 32) + (q >> 32)) << 32);
            changed = true;
        }
        if (changed) pair[x] = p;
        open = !((uint32_t)p & U_HEAD);
    }
    // one atomic per block: in the first rounds every warp has open nodes, and 120 000 same-address atomics per launch
    // would serialise in one L2 slice
    __shared__ uint32_t s_open;
    if (threadIdx.x == 0) s_open = 0;
    __syncthreads();
    const uint32_t m = __ballot_sync(0xffffffffu, open);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_open, (uint32_t)__popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_open) atomicAdd(n_open, (unsigned long long)s_open);
}

void launch_unitig_jump(unsigned long long* pair, uint32_t n2, int steps, unsigned long long* n_open, cudaStream_t s) {
    if (!n2) return;
    unitig_jump_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(pair, n2, steps, n_open);
}

// ------------------------------------------------------------------ cycles (what pointer jumping cannot resolve)
// compaction of the open nodes: cyc_list[c] = x, cyc_pos[x] = c
__global__ void __launch_bounds__(256) unitig_cycle_list_kernel(const unsigned long long* pair, uint32_t n2, uint32_t* cyc_list,
                                                                uint32_t* cyc_pos, unsigned long long* cursor) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool open = x < n2 && !((uint32_t)pair[x] & U_HEAD);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t m = __ballot_sync(0xffffffffu, open);
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (open) {
        const uint32_t c = (uint32_t)base + __popc(m & ((1u << lane) - 1u));
        cyc_list[c] = x;
        cyc_pos[x] = c;
    }
}

void launch_unitig_cycle_list(const unsigned long long* pair, uint32_t n2, uint32_t* cyc_list, uint32_t* cyc_pos,
                              unsigned long long* cursor, cudaStream_t s) {
    if (!n2) return;
    unitig_cycle_list_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(pair, n2, cyc_list, cyc_pos, cursor);
}

// best[c] = {h1, h2, x} of the cycle node itself, jump[c] = position of its successor
__global__ void __launch_bounds__(256) unitig_cycle_init_kernel(const UnitigArgs a, const uint32_t* cyc_list, const uint32_t* cyc_pos,
                                                                uint32_t m, uint64_t* best, uint32_t* jump) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint32_t x = cyc_list[c];
    const Slot* sl = a.table + a.node_slot[x >> 1];
    best[3 * (uint64_t)c] = sl->hi;                                  // u128 order of the reference: (h1 << 64) | h2
    best[3 * (uint64_t)c + 1] = sl->lo;
    best[3 * (uint64_t)c + 2] = x;
    jump[c] = cyc_pos[a.next[x]];
}

__device__ __forceinline__ bool best_less(const uint64_t* p, const uint64_t* q) {
    if (p[0] != q[0]) return p[0] < q[0];
    if (p[1] != q[1]) return p[1] < q[1];
    return p[2] < q[2];
}

// one synchronous doubling round (ping-pong buffers): the window of c doubles, its minimum follows
__global__ void __launch_bounds__(256) unitig_cycle_min_kernel(const uint64_t* best_in, const uint32_t* jump_in, uint32_t m,
                                                               uint64_t* best_out, uint32_t* jump_out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint32_t j = jump_in[c];
    const uint64_t* p = best_in + 3 * (uint64_t)c;
    const uint64_t* q = best_in + 3 * (uint64_t)j;
    const uint64_t* w = best_less(q, p) ? q : p;
    best_out[3 * (uint64_t)c] = w[0];
    best_out[3 * (uint64_t)c + 1] = w[1];
    best_out[3 * (uint64_t)c + 2] = w[2];
    jump_out[c] = jump_in[j];
}

// cut every cycle in front of its leader and restart the ranking of its nodes
__global__ void __launch_bounds__(256) unitig_cycle_cut_kernel(const uint32_t* next, const uint32_t* cyc_list, const uint64_t* best,
                                                               uint32_t m, unsigned long long* pair, uint8_t* is_cycle_head) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const uint32_t x = cyc_list[c];
    if ((uint32_t)best[3 * (uint64_t)c + 2] == x) {
        pair[x] = (unsigned long long)(U_HEAD | x);
        is_cycle_head[x] = 1;
    } else {
        pair[x] = (1ULL << 32) | (unsigned long long)(next[x ^ 1u] ^ 1u);
    }
}

void launch_unitig_cycle_init(const UnitigArgs& a, const uint32_t* cyc_list, const uint32_t* cyc_pos, uint32_t m, uint64_t* best,
                              uint32_t* jump, cudaStream_t s) {
    if (!m) return;
    unitig_cycle_init_kernel<<<(m + 255) / 256, 256, 0, s>>>(a, cyc_list, cyc_pos, m, best, jump);
}
void launch_unitig_cycle_min(const uint64_t* best_in, const uint32_t* jump_in, uint32_t m, uint64_t* best_out, uint32_t* jump_out,
                             cudaStream_t s) {
    if (!m) return;
    unitig_cycle_min_kernel<<<(m + 255) / 256, 256, 0, s>>>(best_in, jump_in, m, best_out, jump_out);
}
void launch_unitig_cycle_cut(const uint32_t* next, const uint32_t* cyc_list, const uint64_t* best, uint32_t m,
                             unsigned long long* pair, uint8_t* is_cycle_head, cudaStream_t s) {
    if (!m) return;
    unitig_cycle_cut_kernel<<<(m + 255) / 256, 256, 0, s>>>(next, cyc_list, best, m, pair, is_cycle_head);
}

// ------------------------------------------------------------------ which paths are written, and how long they are
__global__ void __launch_bounds__(256) unitig_len_kernel(const unsigned long long* pair, uint32_t n2, uint32_t* len) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n2) return;
    const unsigned long long p = pair[x];
    atomicMax(len + ((uint32_t)p & ~U_HEAD), (uint32_t)(p >> 32) + 1u);
}

// A linear path P (head h) and its mirror (head = reverse of P's tail = head of rev(h)) are the same unitig: the one
// with the smaller head is written.  Of a cycle and its mirror the one whose leader -- the node with the smallest
// hash128 -- stands in its normalized orientation is written, starting there (computeUnitigNode2's rotation rule).
__global__ void __launch_bounds__(256) unitig_select_kernel(const unsigned long long* pair, const uint8_t* is_cycle_head,
                                                            const uint32_t* len, uint32_t n2, uint32_t k, uint32_t* size,
                                                            uint32_t* flag) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n2) return;
    uint32_t sz = 0, fl = 0;
    if ((uint32_t)pair[x] == (U_HEAD | x)) {
        const bool emit = is_cycle_head[x] ? !(x & 1u) : x <= ((uint32_t)pair[x ^ 1u] & ~U_HEAD);
        if (emit) { sz = len[x] + k - 1; fl = 1; }
    }
    size[x] = sz;
    flag[x] = fl;
}

void launch_unitig_len(const unsigned long long* pair, uint32_t n2, uint32_t* len, cudaStream_t s) {
    if (!n2) return;
    unitig_len_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(pair, n2, len);
}
void launch_unitig_select(const unsigned long long* pair, const uint8_t* is_cycle_head, const uint32_t* len, uint32_t n2, uint32_t k,
                          uint32_t* size, uint32_t* flag, cudaStream_t s) {
    if (!n2) return;
    unitig_select_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(pair, is_cycle_head, len, n2, k, size, flag);
}

// every oriented node of a written path stores its share of the minimizer sequence: the head its k minimizers, the
// node of rank r its last minimizer at position k - 1 + r
__global__ void __launch_bounds__(256) unitig_scatter_kernel(const UnitigArgs a, const unsigned long long* pair, const uint32_t* flag,
                                                             const uint64_t* seq_off, const uint64_t* unitig_idx,
                                                             uint32_t* out_mins, uint64_t* out_off, uint8_t* out_circular,
                                                             const uint8_t* is_cycle_head, uint32_t* out_abund) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= 2 * a.n_nodes) return;
    const unsigned long long p = pair[x];
    const uint32_t h = (uint32_t)p & ~U_HEAD, r = (uint32_t)(p >> 32);
    if (!flag[h]) return;
    const uint64_t base = seq_off[h];
    const Slot* sl = a.table + a.node_slot[x >> 1];
    const uint64_t ref = sl->ref;
    const int k = (int)a.k;
    // dumpUnitigAbundances (CreateMdbg.cpp:3335-3390): the abundance of every k-min-mer of the unitig, in sequence order
    out_abund[base - unitig_idx[h] * (uint64_t)(k - 1) + r] = sl->count;
    if (r == 0) {
        for (int j = 0; j < k; j++) out_mins[base + j] = onode_elem(a, ref, x & 1u, j);
        out_off[unitig_idx[h]] = base;
        out_circular[unitig_idx[h]] = is_cycle_head[h];
    } else {
        out_mins[base + k - 1 + r] = onode_elem(a, ref, x & 1u, k - 1);
    }
}

void launch_unitig_scatter(const UnitigArgs& a, const unsigned long long* pair, const uint32_t* flag, const uint64_t* seq_off,
                           const uint64_t* unitig_idx, uint32_t* out_mins, uint64_t* out_off, uint8_t* out_circular,
                           const uint8_t* is_cycle_head, uint32_t* out_abund, cudaStream_t s) {
    if (!a.n_nodes) return;
    unitig_scatter_kernel<<<(unsigned)((2 * (uint64_t)a.n_nodes + 255) / 256), 256, 0, s>>>(a, pair, flag, seq_off, unitig_idx, out_mins,
                                                                                          out_off, out_circular, is_cycle_head, out_abund);
}

// computeDeterministicUnitigs: KmerVec::normalize on the whole minimizer sequence, hash128 of the normalized sequence.
// One thread per unitig (Murmur is sequential over the sequence); out_hashes[2u] = low word (h2), [2u + 1] = high (h1).
__global__ void __launch_bounds__(128) unitig_hash_kernel(const uint32_t* mins, const uint64_t* off, uint64_t n_unitigs,
                                                          uint64_t* out_hashes, uint8_t* out_rev) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_unitigs) return;
    const uint32_t* w = mins + off[u];
    const uint64_t L = off[u + 1] - off[u];
    bool rev = true;
    for (uint64_t j = 0; j < L / 2; j++) {
        const uint32_t p = w[j], q = w[L - 1 - j];
        if (p != q) { rev = p > q; break; }
    }
    uint64_t h1, h2;
    if (rev) murmur128_u32vec([&](int t) { return w[L - 1 - (uint64_t)t]; }, (int)L, h1, h2);
    else murmur128_u32vec([&](int t) { return w[t]; }, (int)L, h1, h2);
    out_hashes[2 * u] = h2;
    out_hashes[2 * u + 1] = h1;
    out_rev[u] = rev ? 1 : 0;
}

// sequences whose normalized form is the reversed one are reversed in place, one warp per unitig
__global__ void __launch_bounds__(256) unitig_reverse_kernel(uint32_t* mins, const uint64_t* off, uint64_t n_unitigs,
                                                             const uint8_t* rev, uint32_t* abund, uint32_t k) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t u = warp; u < n_unitigs; u += n_warps) {
        if (!rev[u]) continue;
        uint32_t* w = mins + off[u];
        const uint64_t L = off[u + 1] - off[u];
        for (uint64_t j = lane; j < L / 2; j += 32) {
            const uint32_t t = w[j];
            w[j] = w[L - 1 - j];
            w[L - 1 - j] = t;
        }
        uint32_t* ab = abund + (off[u] - u * (uint64_t)(k - 1));        // the k-min-mers of the reversed sequence, in its order
        const uint64_t nw = L - (k - 1);
        for (uint64_t j = lane; j < nw / 2; j += 32) {
            const uint32_t t = ab[j];
            ab[j] = ab[nw - 1 - j];
            ab[nw - 1 - j] = t;
        }
    }
}

void launch_unitig_hash(const uint32_t* mins, const uint64_t* off, uint64_t n_unitigs, uint64_t* out_hashes, uint8_t* out_rev,
                        cudaStream_t s) {
    if (!n_unitigs) return;
    unitig_hash_kernel<<<(unsigned)((n_unitigs + 127) / 128), 128, 0, s>>>(mins, off, n_unitigs, out_hashes, out_rev);
}
void launch_unitig_reverse(uint32_t* mins, const uint64_t* off, uint64_t n_unitigs, const uint8_t* rev, uint32_t* abund, uint32_t k,
                           cudaStream_t s) {
    if (!n_unitigs) return;
    uint64_t blocks = (n_unitigs + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    unitig_reverse_kernel<<<(unsigned)blocks, 256, 0, s>>>(mins, off, n_unitigs, rev, abund, k);
}

// ------------------------------------------------------------------ the reference's deterministic order, on the device
// computeDeterministicUnitigs sorts the unitigs by the u128 hash of their normalized sequence.  The hashes are uniform,
// so one counting pass over the top bits of the high word spreads n unitigs over >= n / 2 buckets of a few elements
// each; a thread per bucket finishes with an insertion sort on (high, low, index).  order[i] = unitig at position i.
__global__ void __launch_bounds__(256) unitig_bucket_count_kernel(const uint64_t* hashes, uint64_t n, uint32_t shift, uint32_t* cnt) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) atomicAdd(cnt + (hashes[2 * u + 1] >> shift), 1u);
}
__global__ void __launch_bounds__(256) unitig_bucket_fill_kernel(const uint64_t* hashes, uint64_t n, uint32_t shift,
                                                                 const uint64_t* bucket_off, uint32_t* cursor, uint32_t* order) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const uint64_t b = hashes[2 * u + 1] >> shift;
    order[bucket_off[b] + atomicAdd(cursor + b, 1u)] = (uint32_t)u;
}
__global__ void __launch_bounds__(256) unitig_bucket_sort_kernel(const uint64_t* hashes, uint64_t n_buckets, const uint64_t* bucket_off,
                                                                 uint32_t* order, uint32_t* pos_of) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    const uint64_t lo = bucket_off[b], hi = bucket_off[b + 1];
    for (uint64_t i = lo + 1; i < hi; i++) {
        const uint32_t x = order[i];
        const uint64_t xh = hashes[2 * (uint64_t)x + 1], xl = hashes[2 * (uint64_t)x];
        uint64_t j = i;
        while (j > lo) {
            const uint32_t y = order[j - 1];
            const uint64_t yh = hashes[2 * (uint64_t)y + 1], yl = hashes[2 * (uint64_t)y];
            const bool less = xh != yh ? xh < yh : (xl != yl ? xl < yl : x < y);
            if (!less) break;
            order[j] = y;
            j--;
        }
        order[j] = x;
    }
    for (uint64_t i = lo; i < hi; i++) pos_of[order[i]] = (uint32_t)i;
}

// The two figures the reference logs for the stage: "Checksum unitig nodes" = sum over records of minimizer * size *
// unitigIndex (CreateMdbg.cpp:3380, unitigIndex = 2 * position), "Checksum unitig abundance" = sum of abundance * number
// of abundances (:3384).  One thread per unitig, one atomic per warp.
__global__ void __launch_bounds__(256) unitig_checksum_kernel(const uint32_t* mins, const uint64_t* off, const uint32_t* abund,
                                                              const uint32_t* pos_of, uint64_t n, uint32_t k,
                                                              unsigned long long* sums) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cn = 0, ca = 0;
    if (u < n) {
        const uint64_t b = off[u], L = off[u + 1] - b, nw = L - (k - 1);
        unsigned long long sm = 0, sa = 0;
        for (uint64_t j = 0; j < L; j++) sm += mins[b + j];
        const uint32_t* ab = abund + (b - u * (uint64_t)(k - 1));
        for (uint64_t j = 0; j < nw; j++) sa += ab[j];
        cn = sm * L * (unsigned long long)(uint32_t)(2u * pos_of[u]);
        ca = sa * nw;
    }
    for (int d = 16; d; d >>= 1) {
        cn += __shfl_down_sync(0xffffffffu, cn, d);
        ca += __shfl_down_sync(0xffffffffu, ca, d);
    }
    if ((threadIdx.x & 31) == 0 && (cn | ca)) {
        atomicAdd(sums, cn);
        atomicAdd(sums + 1, ca);
    }
}

// ------------------------------------------------------------------ unitig graph edges
// CreateMdbg::indexUnitigEdges / computeUnitigEdges (src/graph/CreateMdbg.cpp:2915-3245; getSuccessors_unitig :2453-2530,
// getPredecessors_unitig :2631-2695).  The first and (when different) the last k-min-mer of every unitig, normalized, are
// entered under their normalized (k-1)-prefix and (k-1)-suffix keys -- the keys are in the edge set already, so a list
// per edge-set slot (count, scan, fill: a CSR) replaces the reference's second MPHF + vector-per-key map.  An entry is
// (seq << 3) | nodeReversed << 2 | keyReversed << 1 | isPrefix with seq = ((2 * position + isLastNode) * 2 + isSuffixKey):
// sorted by value, a list is in the order ONE reference thread would have pushed it (records in file order, first node
// before last node, prefix entry before suffix entry), so the successor lists come out in the reference's own order.
__device__ __forceinline__ void unitig_end_node(const uint32_t* seq, uint64_t L, uint32_t k, bool last, bool reversed_unitig,
                                                uint32_t* out) {
    // k-min-mer number 0 / L-k of the sequence, or of the reversed sequence
    const uint64_t at = (last != reversed_unitig) ? L - k : 0;
    for (uint32_t j = 0; j < k; j++) out[j] = reversed_unitig ? seq[at + k - 1 - j] : seq[at + j];
}

constexpr int UNITIG_MAX_K = 64;                                  // the end-node work areas live in registers / local memory

__global__ void __launch_bounds__(128) unitig_end_offers_kernel(const UnitigEdgeArgs a, int pass) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_unitigs) return;
    const uint32_t k = a.k, km = k - 1;
    const uint32_t* seq = a.mins + a.off[j];
    const uint64_t L = a.off[j + 1] - a.off[j];
    const uint64_t pos = a.pos_of[j];
    uint32_t node[UNITIG_MAX_K];
    const bool single = L == k;                                   // startNode == endNode: entered once
    bool same = single;
    if (!single) {                                                // (equal vectors at both ends also count once)
        same = true;
        for (uint32_t t = 0; t < k; t++) if (seq[t] != seq[L - k + t]) { same = false; break; }
    }
    for (int end = 0; end < (same ? 1 : 2); end++) {
        unitig_end_node(seq, L, k, end != 0, false, node);
        bool nrev = true;                                         // KmerVec::normalize of the node
        for (uint32_t t = 0; t < k / 2; t++) if (node[t] != node[k - 1 - t]) { nrev = node[t] > node[k - 1 - t]; break; }
        if (nrev) for (uint32_t t = 0; t < k / 2; t++) { const uint32_t x = node[t]; node[t] = node[k - 1 - t]; node[k - 1 - t] = x; }
        for (uint32_t side = 0; side < 2; side++) {               // prefix key, then suffix key (indexEdgeUnitig)
            const uint32_t* w = node + side;
            bool krev = true;
            for (uint32_t t = 0; t < km / 2; t++) if (w[t] != w[km - 1 - t]) { krev = w[t] > w[km - 1 - t]; break; }
            uint64_t h1, h2;
            if (krev) murmur128_u32vec([&](int t) { return w[km - 1 - t]; }, (int)km, h1, h2);
            else murmur128_u32vec([&](int t) { return w[t]; }, (int)km, h1, h2);
            const Slot* es = table_find(const_cast<Slot*>(a.edges), a.edge_mask, h2, h1);
            if (!es) { atomicExch(a.error_flag, 3u); continue; }
            const uint64_t slot = (uint64_t)(es - a.edges);
            if (pass == 1) {
                atomicAdd(a.slot_cnt + slot, 1u);
            } else {
                const uint64_t seqno = (2 * pos + (uint64_t)end) * 2 + side;
                a.entries[a.slot_off[slot] + atomicAdd(a.slot_cnt + slot, 1u)] =
                    (seqno << 3) | ((unsigned long long)(nrev ? 1 : 0) << 2) | ((unsigned long long)(krev ? 1 : 0) << 1) | (side ? 0ULL : 1ULL);
            }
        }
    }
}

// every list into ascending order (= the single-thread push order of the reference); lists are short
__global__ void __launch_bounds__(256) unitig_end_sort_kernel(const uint64_t* slot_off, uint64_t n_slots, unsigned long long* entries) {
    const uint64_t sl = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= n_slots) return;
    const uint64_t lo = slot_off[sl], hi = slot_off[sl + 1];
    for (uint64_t i = lo + 1; i < hi; i++) {
        const unsigned long long x = entries[i];
        uint64_t q = i;
        while (q > lo && entries[q - 1] > x) { entries[q] = entries[q - 1]; q--; }
        entries[q] = x;
    }
}

// successors of oriented unitig x = 2 * position + reversed (the predecessors of a unitig are the successors of its
// reverse): pass 1 counts, pass 2 writes targets and adds x * target to the checksum (_checksum_unitigEdges)
__global__ void __launch_bounds__(128) unitig_edges_query_kernel(const UnitigEdgeArgs a, int pass) {
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cs = 0;
    if (x < 2 * a.n_unitigs) {
        const uint32_t k = a.k, km = k - 1;
        const uint64_t j = a.order[x >> 1];
        const uint32_t* seq = a.mins + a.off[j];
        const uint64_t L = a.off[j + 1] - a.off[j];
        uint32_t node[UNITIG_MAX_K];
        unitig_end_node(seq, L, k, true, (x & 1) != 0, node);    // the last k-min-mer of x as oriented
        const uint32_t* S = node + 1;
        bool srev = true, pal = true;
        for (uint32_t t = 0; t < km / 2; t++) if (S[t] != S[km - 1 - t]) { srev = S[t] > S[km - 1 - t]; pal = false; break; }
        uint64_t h1, h2;
        if (srev) murmur128_u32vec([&](int t) { return S[km - 1 - t]; }, (int)km, h1, h2);
        else murmur128_u32vec([&](int t) { return S[t]; }, (int)km, h1, h2);
        const Slot* es = table_find(const_cast<Slot*>(a.edges), a.edge_mask, h2, h1);
        uint32_t n = 0;
        if (!es) {
            atomicExch(a.error_flag, 3u);
        } else {
            const uint64_t slot = (uint64_t)(es - a.edges);
            const uint64_t base = pass == 2 ? a.edge_off[x] : 0;
            for (uint64_t e = a.slot_off[slot]; e < a.slot_off[slot + 1]; e++) {
                const unsigned long long ent = a.entries[e];
                const uint32_t idx = (uint32_t)(2 * (ent >> 5)) + (uint32_t)((ent >> 2) & 1);   // seq >> 2 = position
                const bool krev = (ent >> 1) & 1, pre = ent & 1;
                uint32_t target;
                if (pre) {
                    if (!(pal || krev == srev)) continue;         // getSuccessors_unitig: vec_suffix == v
                    if ((uint32_t)x == (idx ^ 1u)) continue;
                    target = idx;
                } else {
                    if (!(pal || krev != srev)) continue;         // vec_suffix == reverse(v)
                    if ((uint32_t)x == idx) continue;
                    target = idx ^ 1u;
                }
                // (the reference multiplies two 32-bit UnitigType values: every product wraps at 2^32 before it is summed)
                if (pass == 2) { a.edge_targets[base + n] = target; cs += (unsigned long long)(uint32_t)((uint32_t)x * target); }
                n++;
            }
        }
        if (pass == 1) a.edge_cnt[x] = n;
    }
    if (pass == 2) {
        for (int d = 16; d; d >>= 1) cs += __shfl_down_sync(0xffffffffu, cs, d);
        if ((threadIdx.x & 31) == 0 && cs) atomicAdd(a.checksum, cs);
    }
}

void launch_unitig_end_offers(const UnitigEdgeArgs& a, int pass, cudaStream_t s) {
    if (!a.n_unitigs) return;
    unitig_end_offers_kernel<<<(unsigned)((a.n_unitigs + 127) / 128), 128, 0, s>>>(a, pass);
}
void launch_unitig_end_sort(const uint64_t* slot_off, uint64_t n_slots, unsigned long long* entries, cudaStream_t s) {
    if (!n_slots) return;
    unitig_end_sort_kernel<<<(unsigned)((n_slots + 255) / 256), 256, 0, s>>>(slot_off, n_slots, entries);
}
void launch_unitig_edges_query(const UnitigEdgeArgs& a, int pass, cudaStream_t s) {
    if (!a.n_unitigs) return;
    unitig_edges_query_kernel<<<(unsigned)((2 * a.n_unitigs + 127) / 128), 128, 0, s>>>(a, pass);
}

void launch_unitig_sort(const uint64_t* hashes, uint64_t n, uint32_t bucket_bits, uint32_t* cnt, uint64_t* bucket_off,
                        uint64_t* scan_scratch, uint32_t* order, uint32_t* pos_of, cudaStream_t s) {
    if (!n) return;
    const uint64_t nb = 1ull << bucket_bits;
    const uint32_t shift = 64 - bucket_bits;
    cudaMemsetAsync(cnt, 0, nb * 4, s);
    unitig_bucket_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hashes, n, shift, cnt);
    launch_scan_u32_to_u64(cnt, bucket_off, (uint32_t)nb, scan_scratch, s);
    cudaMemsetAsync(cnt, 0, nb * 4, s);
    unitig_bucket_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hashes, n, shift, bucket_off, cnt, order);
    unitig_bucket_sort_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, s>>>(hashes, nb, bucket_off, order, pos_of);
}
void launch_unitig_checksum(const uint32_t* mins, const uint64_t* off, const uint32_t* abund, const uint32_t* pos_of, uint64_t n,
                            uint32_t k, unsigned long long* sums, cudaStream_t s) {
    if (!n) return;
    unitig_checksum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mins, off, abund, pos_of, n, k, sums);
}

}  // namespace mdbg
