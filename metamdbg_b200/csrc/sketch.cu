// sketch.cu -- K1: homopolymer compression + canonical l-mer roll + MurmurHash3
// density-threshold selection, one warp per read.
//
// Replaces (reference, paths relative to the metaMDBG tree):
//   EncoderRLE::execute          src/Commons.hpp:4163-4203
//   KmerModel::iterate/first/next src/utils/kmer/Kmer.hpp:488-611
//   MinimizerParser::parse       src/utils/kmer/Kmer.hpp:1373-1456
//   MurmurHash3_x64_128          src/utils/MurmurHash3.cpp:246-325
//
// Data flow per warp (no block-level synchronisation anywhere):
//   HBM --16 B/lane coalesced loads (ASCII) or 4 B/lane (2-bit packed host batches)--> registers
//   --keep-mask (SIMD-in-register compare, warp scan)--> 2 KB shared-memory ring of HPC base codes
//   --2 x LDS.128 per lane--> 16 consecutive l-mers per lane rolled in registers
//   --high-word Murmur test--> candidates appended in position order to a 64-entry per-warp list
//   --32 at a time: exact hash, blacklist, ballot compaction--> coalesced (value, position, strand)
//   writes into the read's output slot.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "engine.cuh"
#include "bitmath.cuh"

namespace mdbg {

constexpr int WARPS_PER_CTA = 8;
constexpr int RING = 2048;        // bytes of HPC codes per warp (power of two)
constexpr int BLK = 512;          // l-mer positions per warp step (16 per lane)
constexpr int CHUNK = 512;        // raw bytes per warp load step (16 per lane)

__device__ __forceinline__ bool blacklisted(const uint32_t* bl, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (bl[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < n && bl[lo] == v;
}

// The fill phase writes each lane's bases linearly (the ring has 16 bytes of slack past RING).  When this chunk
// crosses the end of the ring exactly one lane straddles it; its overhang is copied back to the ring start by the
// first lanes of the warp (warp-uniform test, one byte per lane).
__device__ __forceinline__ void wrap_overhang(uint8_t* ring, uint32_t avail, uint32_t total, uint32_t base_idx,
                                              uint32_t cnt, uint32_t lane) {
    if ((avail & (RING - 1)) + total <= (uint32_t)RING) return;            // uniform
    const bool straddle = base_idx < (uint32_t)RING && base_idx + cnt > (uint32_t)RING;
    const uint32_t who = __ballot_sync(0xffffffffu, straddle);
    if (who == 0) return;                                                  // the boundary fell between two lanes
    const uint32_t over = __shfl_sync(0xffffffffu, base_idx + cnt - RING, __ffs(who) - 1);
    __syncwarp();
    if (lane < over) ring[lane] = ring[RING + lane];
}

// forward l-mer (2-bit codes of all l characters, valid or not) starting at HPC position p; inv != 0 when the
// window holds an invalid character
__device__ __forceinline__ uint32_t fwd_at(const uint8_t* ring, uint32_t p, uint32_t l, uint32_t& inv) {
    uint32_t fwd = 0;
    inv = 0;
    for (uint32_t t = 0; t < l; t++) {
        const uint32_t b = ring[(p + t) & (RING - 1)];
        inv |= b & 4u;
        fwd = (fwd << 2) | (b & 3u);
    }
    if (l < 16) fwd &= (1u << (2 * l)) - 1u;
    return fwd;
}

// One candidate-list entry per lane (lane < cnt): exact selection test, blacklist, then the survivors of the
// warp are written in list (= position) order with consecutive indices.  Returns the number written.
__device__ __forceinline__ uint32_t flush_candidates(const SketchArgs& a, const uint2* cand, uint32_t cnt, uint32_t lane,
                                                     uint64_t slot_lo, uint64_t slot_cap, uint32_t out_cnt) {
    const uint32_t l = a.l;
    bool ok = lane < cnt;
    uint32_t pos = 0, v32 = 0, dir = 0;
    if (ok) {
        const uint2 e = cand[lane];
        pos = e.x & 0x7FFFFFFFu;
        const uint32_t fwd = e.y;
        const uint32_t mask = (l < 16) ? ((1u << (2 * l)) - 1u) : 0xFFFFFFFFu;
        uint32_t x = brev32(fwd ^ (0xAAAAAAAAu & mask));               // reverse complement (code ^ 2, reversed)
        x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
        const uint32_t rc = x >> (32 - 2 * l);
        dir = (fwd < rc) ? 0u : 1u;                                     // KmerCanonical::updateChoice, Kmer.hpp:427
        const uint64_t key = (e.x >> 31) ? ~0ULL : (uint64_t)(dir ? rc : fwd);
        v32 = (uint32_t)key;                                            // Kmer.hpp:1441 truncation
        ok = murmur_h1_u64(key) <= a.threshold;                         // exact test, Kmer.hpp:1434
        if (ok && a.n_blacklist) ok = !blacklisted(a.blacklist, a.n_blacklist, v32);
    }
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
        const uint64_t idx = (uint64_t)out_cnt + __popc(m & ((1u << lane) - 1u));
        if (idx < slot_cap) {
            a.out_min[slot_lo + idx] = v32;
            a.out_pos[slot_lo + idx] = pos;
            a.out_dir[slot_lo + idx] = (uint8_t)dir;
        }
    }
    return __popc(m);
}

// Generic path: runtime l, invalid characters tracked as the reference does
// (indexBadChar, Kmer.hpp:552-556).
__device__ __forceinline__ uint32_t roll16_generic(const uint8_t* ring, uint32_t p0, uint32_t l, uint64_t threshold,
                                                   uint32_t valid_bits) {
    const uint32_t mask = (l < 16) ? ((1u << (2 * l)) - 1u) : 0xFFFFFFFFu;
    uint32_t fwd = 0, rc = 0, sel = 0;
    int bad = -1;
    for (uint32_t t = 0; t < l + 15; t++) {
        uint32_t b = ring[(p0 + t) & (RING - 1)];
        uint32_t c = b & 3;
        bad = (b & 4) ? (int)l - 1 : bad - 1;
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((c ^ 2u) << (2 * l - 2));
        if (t >= l - 1) {
            uint32_t j = t - (l - 1);
            if ((valid_bits >> j) & 1u) {
                uint64_t key = (bad < 0) ? (uint64_t)min(fwd, rc) : ~0ULL;
                if (murmur_h1_u64(key) <= threshold) sel |= 1u << j;
            }
        }
    }
    return sel;
}

// V selects the arithmetic of the unrolled register block (everything else is shared):
//   V = 0  roll step per position + murmur_s1_u32          (47.5 instructions per l-mer, 23.1 on the FMA pipe)
//   V = 1  k1v1:: funnel-shift l-mers from packed codes, one final fmix multiply on the summed pre-images,
//          carry-chain accept bits                         (40.6 instructions per l-mer, 18.1 on the FMA pipe)
// Both are supersets confirmed by the exact hash in flush_candidates, so their outputs are identical;
// mdbg_ctx_autotune_sketch() verifies that on the caller's own batch and keeps the faster one.
template <int L_FAST, int V>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, V ? 5 : 0) sketch_kernel(const SketchArgs a) {
    __shared__ __align__(16) uint8_t ring_all[WARPS_PER_CTA][RING + 16];   // +16: linear-write slack
    __shared__ uint2 cand_all[WARPS_PER_CTA][64];                           // pending candidates: (pos | invalid<<31, fwd)
    const uint32_t lane = threadIdx.x & 31;
    uint8_t* ring = ring_all[threadIdx.x >> 5];
    uint2* cand = cand_all[threadIdx.x >> 5];
    uint32_t n_list = 0;
    const uint32_t l = a.l;
#ifdef MDBG_POISON_SMEM
    // test builds only (tests/cpp/sketch_emu_test.cpp): shared memory is NOT zero when a CTA starts on a real SM --
    // it holds whatever the previous CTA left there.  Fill the ring with non-code bytes that pass the bit-2
    // "invalid character" test, so that any dependence on bytes past `avail` shows up as a wrong sketch.
    for (uint32_t i = lane; i < (uint32_t)RING + 16; i += 32) ring[i] = (uint8_t)(0x18u + 0x20u * (i % 7u) + (i & 3u));
    __syncwarp();
#endif
    // V = 0: T_hi + 1, V = 1: T_hi + S1_SLACK; a wrapped value (overflow) disables the fast path
    const uint32_t thr_cand = (uint32_t)(a.threshold >> 32) + (V ? k1v1::S1_SLACK : 1u);
    constexpr uint32_t THR_MIN = V ? k1v1::S1_SLACK : 1u;

    const uint32_t list_n = a.read_list ? *a.read_list_n : 0u;    // list mode: the reads variant 2 left behind
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) {
            r = atomicAdd(a.cursor, 1u);
            if (a.read_list) r = r < list_n ? a.read_list[r] : 0xFFFFFFFFu;
            else r += a.read_begin;
        }
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.read_end) break;

        const uint64_t start = a.offsets[r], end = a.offsets[r + 1];
        const uint32_t len = (uint32_t)(end - start);
        uint64_t slot_lo, slot_cap;
        if (a.exact_off) {
            slot_lo = a.exact_off[r];
            slot_cap = a.exact_off[r + 1] - slot_lo;
        } else {
            slot_lo = (start >> a.cap_shift) + (uint64_t)r * a.cap_const;
            slot_cap = ((end >> a.cap_shift) + (uint64_t)(r + 1) * a.cap_const) - slot_lo;
        }
        // where this read's bases live: the ASCII batch, or (host batches) its 2-bit packed copy
        const uint32_t* pk = nullptr;
        const uint8_t* base = a.bases + start;
        if (a.read_src) {
            const uint64_t src = a.read_src[r];
            if (src & SRC_ASCII) base = a.bases + (src & ~SRC_ASCII);
            else pk = a.packed + src;
        }
        const uint32_t skip = pk ? 0u : (uint32_t)((uintptr_t)base & 15);
        const uint8_t* abase = base - skip;             // 16-byte aligned
        const uint32_t x_end = skip + len;              // aligned-index space: x = skip + i
        const uint32_t n_chunks = (x_end + CHUNK - 1) / CHUNK;

        uint32_t avail = 0, done = 0, out_cnt = 0, chunk = 0;
        uint32_t carry = '#';                           // EncoderRLE's lastChar sentinel

        for (;;) {
            // ---- fill: append HPC base codes to the ring ----------------------
            while (chunk < n_chunks && avail - done < (uint32_t)BLK + l) {
                const uint32_t x0 = chunk * CHUNK + lane * 16;
                if (pk) {
                    // ---- 2-bit packed source: one u32 = my 16 bases (clean reads: only A, C, G, T) ----------
                    const uint32_t w = (x0 < len) ? pk[chunk * 32 + lane] : 0u;
                    const uint32_t nvalid = (len > x0) ? min(len - x0, 16u) : 0u;
                    const uint32_t vm = (1u << nvalid) - 1u;
                    uint32_t prev = __shfl_up_sync(0xffffffffu, w >> 30, 1);
                    if (lane == 0) prev = carry;
                    carry = __shfl_sync(0xffffffffu, w >> 30, 31);
                    uint32_t keep = vm;
                    if (a.hpc) {
                        const uint32_t x = w ^ ((w << 2) | (prev & 3u));
                        uint32_t k16 = even_bits16(x | (x >> 1));      // base differs from the previous base
                        if (x0 == 0) k16 |= 1u;                        // first base: previous char is the '#' sentinel
                        keep = k16 & vm;
                    }
                    const uint32_t cnt = __popc(keep);
                    const uint32_t incl = warp_inclusive_scan(cnt);
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    const uint32_t base_idx = (avail + incl - cnt) & (RING - 1);
                    uint8_t* dst = ring + base_idx;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if ((keep >> j) & 1u) {
                            *dst = (uint8_t)((w >> (2 * j)) & 3u);
                            dst++;
                        }
                    }
                    wrap_overhang(ring, avail, total, base_idx, cnt, lane);
                    avail += total;
                    chunk++;
                    continue;
                }
                // interior chunk: every byte of every lane belongs to the read (all but the first/last chunk)
                const bool interior = (chunk * CHUNK >= skip) && ((chunk + 1) * CHUNK <= x_end);
                uint4 w = make_uint4(0, 0, 0, 0);
                uint32_t vm = 0xFFFFu;
                if (interior) {
                    w = *reinterpret_cast<const uint4*>(abase + x0);
                } else {
                    if (x0 < x_end && x0 + 16 > skip) {
                        const uint8_t* p = abase + x0;
                        if (p + 16 <= a.bases_end) {
                            w = *reinterpret_cast<const uint4*>(p);
                        } else {                         // last 16 bytes of the buffer: byte loads
                            uint32_t t[4] = {0, 0, 0, 0};
                            for (int j = 0; j < 16; j++)
                                if (p + j < a.bases_end) t[j >> 2] |= (uint32_t)p[j] << (8 * (j & 3));
                            w = make_uint4(t[0], t[1], t[2], t[3]);
                        }
                    }
                    // validity of each of my 16 bytes: skip <= x0+j < x_end
                    const uint32_t lo = (skip > x0) ? min(skip - x0, 16u) : 0u;
                    const uint32_t hi = (x_end > x0) ? min(x_end - x0, 16u) : 0u;
                    vm = (hi > lo) ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                }
                uint32_t pb = __shfl_up_sync(0xffffffffu, w.w >> 24, 1);
                if (lane == 0) pb = carry;
                carry = __shfl_sync(0xffffffffu, w.w >> 24, 31);

                uint32_t keep = vm;
                if (a.hpc) {
                    // EncoderRLE (Commons.hpp:4172-4190): a base survives iff it differs from the previous
                    // char and is not the '#' sentinel.  x_q has a zero byte where byte == previous byte.
                    const uint32_t x0w = w.x ^ ((w.x << 8) | pb);
                    const uint32_t x1w = w.y ^ __funnelshift_l(w.x, w.y, 8);
                    const uint32_t x2w = w.z ^ __funnelshift_l(w.y, w.z, 8);
                    const uint32_t x3w = w.w ^ __funnelshift_l(w.z, w.w, 8);
                    uint32_t n0 = nonzero_bytes(x0w), n1 = nonzero_bytes(x1w), n2 = nonzero_bytes(x2w),
                             n3 = nonzero_bytes(x3w);
                    // '#' (0x23) has bit 6 clear; reads made of letters never take this branch
                    const uint32_t low = ~(w.x & w.y & w.z & w.w) & 0x40404040u;
                    if (__any_sync(0xffffffffu, (low != 0) && (vm != 0))) {
                        n0 &= nonzero_bytes(w.x ^ 0x23232323u);
                        n1 &= nonzero_bytes(w.y ^ 0x23232323u);
                        n2 &= nonzero_bytes(w.z ^ 0x23232323u);
                        n3 &= nonzero_bytes(w.w ^ 0x23232323u);
                    }
                    uint32_t k16 = flags_to_mask16(n0, n1, n2, n3);
                    if (!interior && skip >= x0 && skip < x0 + 16) {
                        // first base of the read: the previous char is the '#' sentinel, not the byte before it
                        const uint32_t sh = 8 * ((skip - x0) & 3);
                        const uint32_t wq = (skip - x0) < 4 ? w.x : (skip - x0) < 8 ? w.y : (skip - x0) < 12 ? w.z : w.w;
                        const bool is_hash = ((wq >> sh) & 0xFFu) == 0x23u;
                        k16 = is_hash ? (k16 & ~(1u << (skip - x0))) : (k16 | (1u << (skip - x0)));
                    }
                    keep = k16 & vm;
                }
                const uint32_t cnt = __popc(keep);
                const uint32_t incl = warp_inclusive_scan(cnt);
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t base_idx = (avail + incl - cnt) & (RING - 1);
                const uint32_t cw[4] = {(w.x >> 1) & 0x07070707u, (w.y >> 1) & 0x07070707u,
                                        (w.z >> 1) & 0x07070707u, (w.w >> 1) & 0x07070707u};
                // the ring has 16 bytes of slack past RING: write linearly, then wrap the overhang
                uint8_t* dst = ring + base_idx;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if ((keep >> j) & 1u) {
                        *dst = (uint8_t)(cw[j >> 2] >> (8 * (j & 3)));
                        dst++;
                    }
                }
                wrap_overhang(ring, avail, total, base_idx, cnt, lane);
                avail += total;
                chunk++;
            }
            if (chunk == n_chunks) {                    // all raw bases are in (reached exactly once per read)
                // EncoderRLE drops every run of '#' (its "no previous char" sentinel) EXCEPT a run that ends the
                // read: the final `rleSequence += lastChar` (Commons.hpp:4186) is unconditional.  Such a read is one
                // base longer in HPC space, which moves the last selectable position (n - 2) by one.  The base is
                // only ever part of the trimmed last l-mer; its code is (0x23 >> 1) & 7 = 1.  Packed reads hold
                // only A, C, G, T.
                if (a.hpc && !pk && len > 0 && base[len - 1] == '#') {
                    if (lane == 0) ring[avail & (RING - 1)] = 1;
                    avail += 1;
                }
                chunk = n_chunks + 1;
            }
            __syncwarp();

            // ---- roll + hash + select, 512 positions per step ------------------
            const bool final_ = (chunk > n_chunks);
            // Kmer.hpp:1395: positions 1 .. n-2 with n = L' - l + 1  =>  p <= L' - l - 1
            const int pmax_final = (int)avail - (int)l - 1;
            for (;;) {
                int pmax;
                if (avail >= done + (uint32_t)BLK + l) pmax = 0x7fffffff;   // (done may pass avail at the end)
                else if (final_ && (int)done <= pmax_final) pmax = pmax_final;
                else break;

                const uint32_t p0 = done + lane * 16;
                int nv = 16;
                if (pmax != 0x7fffffff) nv = max(0, min(16, pmax - (int)p0 + 1));
                uint32_t valid_bits = (1u << nv) - 1u;
                if (p0 == 0) valid_bits &= ~1u;          // position 0 is trimmed (Kmer.hpp:1362,1395)

                uint32_t W[8];
                {
                    const uint4 u0 = *reinterpret_cast<const uint4*>(ring + (p0 & (RING - 1)));
                    const uint4 u1 = *reinterpret_cast<const uint4*>(ring + ((p0 + 16) & (RING - 1)));
                    W[0] = u0.x; W[1] = u0.y; W[2] = u0.z; W[3] = u0.w;
                    W[4] = u1.x; W[5] = u1.y; W[6] = u1.z; W[7] = u1.w;
                }
                const uint32_t inv = (W[0] | W[1] | W[2] | W[3] | W[4] | W[5] | W[6] | W[7]) & 0x04040404u;
                const bool fast = (L_FAST != 0) && (l == (uint32_t)L_FAST) && thr_cand >= THR_MIN &&
                                  !__any_sync(0xffffffffu, inv != 0);

                uint32_t sel, sel_fwd = 0;           // V = 0: forward l-mer of the last candidate
                uint32_t s_hi = 0, s_lo = 0;         // V = 1: the lane's 32 codes, packed
                bool regs_ok = false;
                if (a.select_none) {
                    sel = 0;
                } else if (fast) {
                    if constexpr (V == 0) sel = roll16_fast<(L_FAST ? L_FAST : 15)>(W, thr_cand, sel_fwd);
                    else sel = k1v1::roll16_fast<(L_FAST ? L_FAST : 15)>(W, thr_cand, s_hi, s_lo);
                    regs_ok = (valid_bits == 0xFFFFu);
                    sel &= valid_bits;
                } else {
                    sel = roll16_generic(ring, p0, l, a.threshold, valid_bits);
                }
                const uint32_t hit_lanes = __ballot_sync(0xffffffffu, sel != 0);
                if (hit_lanes) {
                    // Candidates (~2 per 512 positions) are appended, in position order, to a small per-warp list;
                    // the list is resolved 32 entries at a time with every lane busy (flush_candidates).
                    const uint32_t n_c = __popc(sel);
                    // common case: at most one candidate per lane, all from the register path -> one store per hit lane
                    if (__all_sync(0xffffffffu, n_c <= 1 && (regs_ok || n_c == 0))) {
                        if (n_c) {
                            const uint32_t j = __ffs(sel) - 1;
                            if constexpr (V != 0) sel_fwd = k1v1::lmer_from_packed<(L_FAST ? L_FAST : 15)>(s_hi, s_lo, j);
                            cand[n_list + __popc(hit_lanes & ((1u << lane) - 1u))] = make_uint2(p0 + j, sel_fwd);
                        }
                        n_list += __popc(hit_lanes);
                        __syncwarp();
                        if (n_list >= 32) {
                            out_cnt += flush_candidates(a, cand, 32, lane, slot_lo, slot_cap, out_cnt);
                            __syncwarp();
                            uint2 keep_e = make_uint2(0, 0);
                            if (32 + lane < n_list) keep_e = cand[32 + lane];
                            __syncwarp();
                            if (32 + lane < n_list) cand[lane] = keep_e;
                            n_list -= 32;
                            __syncwarp();
                        }
                        done += BLK;
                        continue;
                    }
                    uint32_t before, total;
                    if (__ballot_sync(0xffffffffu, n_c > 1) == 0) {
                        before = __popc(hit_lanes & ((1u << lane) - 1u));
                        total = __popc(hit_lanes);
                    } else {
                        const uint32_t incl = warp_inclusive_scan(n_c);
                        total = __shfl_sync(0xffffffffu, incl, 31);
                        before = incl - n_c;
                    }
                    for (uint32_t base_rank = 0; base_rank < total; base_rank += 32) {   // one pass unless > 32 candidates
                        uint32_t rest = sel, rank = before;
                        while (rest) {
                            const uint32_t j = __ffs(rest) - 1;
                            rest &= rest - 1;
                            if (rank >= base_rank && rank < base_rank + 32) {
                                uint32_t fwd, inv = 0;
                                if (V == 0 && regs_ok && n_c == 1) fwd = sel_fwd;
                                else if (V != 0 && regs_ok) fwd = k1v1::lmer_from_packed<(L_FAST ? L_FAST : 15)>(s_hi, s_lo, j);
                                else fwd = fwd_at(ring, p0 + j, l, inv);
                                cand[n_list + (rank - base_rank)] = make_uint2((p0 + j) | (inv ? 0x80000000u : 0u), fwd);
                            }
                            rank++;
                        }
                        n_list += min(32u, total - base_rank);
                        __syncwarp();
                        if (n_list >= 32) {
                            out_cnt += flush_candidates(a, cand, 32, lane, slot_lo, slot_cap, out_cnt);
                            __syncwarp();
                            uint2 keep_e = make_uint2(0, 0);
                            if (32 + lane < n_list) keep_e = cand[32 + lane];
                            __syncwarp();
                            if (32 + lane < n_list) cand[lane] = keep_e;
                            n_list -= 32;
                            __syncwarp();
                        }
                    }
                }
                done += BLK;
            }
            __syncwarp();
            if (final_) break;
        }
        if (n_list) {
            out_cnt += flush_candidates(a, cand, n_list, lane, slot_lo, slot_cap, out_cnt);
            n_list = 0;
            __syncwarp();
        }
        if (lane == 0) {
            a.n_min[r] = out_cnt;
            if ((uint64_t)out_cnt > slot_cap) atomicAdd(a.n_overflow, 1ULL);
        }
    }
}


// ------------------------------------------------------------------ K1, packed form (variant 2)
// Same result as sketch_kernel, different data path -- the layout north_star names:
//   HBM: 2-bit packed bases (16 per u32)  --cp.async.bulk (TMA) 1 KB tiles, mbarrier, double buffered per warp-->
//   shared-memory stage  --LDS.64 per lane (32 bases)-->  in-register homopolymer compaction: four 10-bit lookups
//   per word in a 2 KB table ((previous code, 4 codes) -> compacted codes + count), warp prefix sum, and the lane's
//   compacted string is OR-ed at its bit offset into a 512-byte ring of two-bit codes (no byte stores, no per-base
//   predicates)  --2 x LDS.32 per lane-->  the variant-1 arithmetic on 16 consecutive l-mers, whose funnel-shift
//   form wants exactly this packed string (no pack multiplies)  -->  candidate list / exact confirm as before.
// No lane issues a global load or address arithmetic in the fill; one elected lane issues the bulk copies.
// Reads the host kept as ASCII (a byte outside "ACGT") are appended to a list for the byte-ring kernel.
constexpr int P_RING_WORDS = 128;                // 2048 codes per warp (power of two)
constexpr int P_TILE_WORDS = 256;                // one bulk copy: 1 KB = 4096 bases
constexpr int P_STEP_WORDS = 64;                 // one fill step: 2 words (32 bases) per lane = 1024 bases
constexpr int P_STEPS_PER_TILE = P_TILE_WORDS / P_STEP_WORDS;

struct __align__(16) PackedWarpSmem {
    uint32_t stage[2][P_TILE_WORDS];
    uint32_t ring[P_RING_WORDS];
    uint2 cand[64];
    unsigned long long bar[2];
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// arm the barrier with the byte count, then start the bulk copy global -> shared that completes it
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "MDBG_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MDBG_DONE;\n\t"
        "bra MDBG_WAIT;\n\t"
        "MDBG_DONE:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
#else   // CPU emulator (tests/cpp/warp_emu.hpp): the copy completes when it is issued, waiting is a warp rendezvous
static inline void mbar_init(unsigned long long*, uint32_t) {}
static inline void mbar_init_fence() {}
static inline void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long*) { memcpy(dst, src, bytes); }
static inline void mbar_wait(unsigned long long*, uint32_t) { __syncwarp(); }
#endif

// hpc_lut[prev | c0 << 2 | c1 << 4 | c2 << 6 | c3 << 8] = surviving codes (first survivor in the low bits) | 2 * count << 8
__device__ __forceinline__ uint32_t hpc_lut_entry(uint32_t idx) {
    uint32_t p = idx & 3u, out = 0, n = 0;
    for (int t = 0; t < 4; t++) {
        const uint32_t c = (idx >> (2 + 2 * t)) & 3u;
        if (c != p) { out |= c << (2 * n); n++; }
        p = c;
    }
    return out | (2u * n) << 8;
}

// four table entries (one word of 16 bases) -> compacted codes; n_bits = 2 * survivors
__device__ __forceinline__ uint32_t hpc_merge4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t& n_bits) {
    uint32_t acc = e3 & 0xFFu;
    acc = (acc << (e2 >> 8)) | (e2 & 0xFFu);
    acc = (acc << (e1 >> 8)) | (e1 & 0xFFu);
    acc = (acc << (e0 >> 8)) | (e0 & 0xFFu);
    n_bits = (e0 >> 8) + (e1 >> 8) + (e2 >> 8) + (e3 >> 8);
    return acc;
}

template <int L>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 5) sketch_packed_kernel(const SketchArgs a) {
    __shared__ PackedWarpSmem smem_all[WARPS_PER_CTA];
    __shared__ __align__(16) uint16_t hpc_lut[1024];
    const uint32_t lane = threadIdx.x & 31;
    PackedWarpSmem& sm = smem_all[threadIdx.x >> 5];
    uint32_t* ring = sm.ring;
    uint2* cand = sm.cand;
    const uint8_t* lut_bytes = reinterpret_cast<const uint8_t*>(hpc_lut);
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) hpc_lut[i] = (uint16_t)hpc_lut_entry(i);
    if (lane == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_init_fence();
    }
#ifdef MDBG_POISON_SMEM
    for (uint32_t i = lane; i < 2 * P_TILE_WORDS; i += 32) sm.stage[0][i] = 0x9E3779B9u * (i + 1);   // test builds: stale stage
    for (uint32_t i = lane; i < P_RING_WORDS; i += 32) ring[i] = 0xDEADBEEFu + i;
#endif
    __syncthreads();
    uint32_t n_list = 0, phase = 0;                     // bit s of phase: parity the next wait on stage s uses
    const uint32_t thr_cand = (uint32_t)(a.threshold >> 32) + k1v1::S1_SLACK;

    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = a.read_begin + atomicAdd(a.cursor, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.read_end) break;
        const uint64_t src = a.read_src[r];
        if (src & SRC_ASCII) {                          // not packable: the byte-ring kernel takes it (list mode)
            if (lane == 0) a.dirty_list[atomicAdd(a.dirty_count, 1u)] = r;
            continue;
        }
        const uint64_t start = a.offsets[r], end = a.offsets[r + 1];
        const uint32_t len = (uint32_t)(end - start);
        uint64_t slot_lo, slot_cap;
        if (a.exact_off) {
            slot_lo = a.exact_off[r];
            slot_cap = a.exact_off[r + 1] - slot_lo;
        } else {
            slot_lo = (start >> a.cap_shift) + (uint64_t)r * a.cap_const;
            slot_cap = ((end >> a.cap_shift) + (uint64_t)(r + 1) * a.cap_const) - slot_lo;
        }
        // stage coordinates: word x of the 16-byte aligned copy; base b = 16 x + j; the read is [b_lo, b_hi)
        const uint32_t skipw = (uint32_t)(src & 3u);
        const uint32_t* gsrc = a.packed + (src - skipw);
        const uint32_t b_lo = 16u * skipw, b_hi = b_lo + len;
        const uint32_t tot_words = len ? skipw + ((len + 15u) >> 4) : 0u;
        const uint32_t tot_bytes = ((tot_words + 3u) & ~3u) * 4u;            // bulk copies move multiples of 16 bytes
        const uint32_t n_tiles = (tot_words + P_TILE_WORDS - 1) / P_TILE_WORDS;
        const uint32_t n_steps = (tot_words + P_STEP_WORDS - 1) / P_STEP_WORDS;

        // the ring is written by OR: it must be all zero where codes are going to land
        *reinterpret_cast<uint4*>(ring + 4 * lane) = make_uint4(0, 0, 0, 0);
        __syncwarp();                                   // every lane is done with the previous read's stage and ring
        if (lane == 0) {
            if (n_tiles > 0) bulk_load(sm.stage[0], gsrc, min(tot_bytes, (uint32_t)P_TILE_WORDS * 4u), &sm.bar[0]);
            if (n_tiles > 1) bulk_load(sm.stage[1], gsrc + P_TILE_WORDS, min(tot_bytes - P_TILE_WORDS * 4u, (uint32_t)P_TILE_WORDS * 4u), &sm.bar[1]);
        }

        uint32_t avail = 0, done = 0, out_cnt = 0, step = 0;
        uint32_t carry = 0;                             // last code of the previous step (lane 31)

        for (;;) {
            // ---- fill: append compacted codes to the ring --------------------------
            while (step < n_steps && avail - done < (uint32_t)BLK + (uint32_t)L) {
                const uint32_t tile = step / P_STEPS_PER_TILE, sub = step % P_STEPS_PER_TILE, st = tile & 1u;
                if (sub == 0) {
                    mbar_wait(&sm.bar[st], (phase >> st) & 1u);
                    phase ^= 1u << st;
                }
                const uint2 w = *reinterpret_cast<const uint2*>(&sm.stage[st][sub * P_STEP_WORDS + 2 * lane]);
                const uint32_t b0 = step * (P_STEP_WORDS * 16u) + lane * 32u;           // my first base, stage coordinates
                const uint32_t step_lo = step * (P_STEP_WORDS * 16u);
                const bool head_ok = step_lo >= b_lo;                   // no foreign bases in front of mine (aligned read start)
                const bool interior = head_ok && step_lo + P_STEP_WORDS * 16u <= b_hi;
                uint32_t prev = __shfl_up_sync(0xffffffffu, w.y >> 30, 1);
                if (lane == 0) prev = carry;
                carry = __shfl_sync(0xffffffffu, w.y >> 30, 31);
                uint32_t lo, hi, cnt;
                if (head_ok) {
                    uint32_t wx = w.x, wy = w.y, nvalid = 32;
                    if (!interior) {
                        // last step of the read: the bases behind its end become copies of its last base -- homopolymer
                        // compression drops them, and without it the count cuts them off
                        nvalid = b_hi > b0 ? min(32u, b_hi - b0) : 0u;
                        if (nvalid > 0 && nvalid < 32) {
                            const uint64_t v = (uint64_t)wx | ((uint64_t)wy << 32);
                            const uint64_t m = (1ULL << (2 * nvalid)) - 1ULL;
                            const uint64_t fill = ((v >> (2 * nvalid - 2)) & 3ULL) * 0x5555555555555555ULL;
                            const uint64_t u = (v & m) | (fill & ~m);
                            wx = (uint32_t)u; wy = (uint32_t)(u >> 32);
                        }
                    }
                    if (a.hpc) {
                        if (b0 == b_lo) prev = (wx & 3u) ^ 1u;                         // first base of the read: always kept
                        const uint32_t X0 = (wx << 2) | prev;
                        const uint32_t X1 = __funnelshift_l(wx, wy, 2);
#define MDBG_LUT(addr) ((uint32_t) * reinterpret_cast<const uint16_t*>(lut_bytes + (addr)))
                        const uint32_t e0 = MDBG_LUT((X0 << 1) & 0x7FEu), e1 = MDBG_LUT((X0 >> 7) & 0x7FEu);
                        const uint32_t e2 = MDBG_LUT((X0 >> 15) & 0x7FEu), e3 = MDBG_LUT(__funnelshift_r(X0, X1, 23) & 0x7FEu);
                        const uint32_t e4 = MDBG_LUT((X1 << 1) & 0x7FEu), e5 = MDBG_LUT((X1 >> 7) & 0x7FEu);
                        const uint32_t e6 = MDBG_LUT((X1 >> 15) & 0x7FEu), e7 = MDBG_LUT((wy >> 21) & 0x7FEu);
#undef MDBG_LUT
                        uint32_t n0, n1;
                        const uint32_t a0 = hpc_merge4(e0, e1, e2, e3, n0), a1 = hpc_merge4(e4, e5, e6, e7, n1);
                        const uint64_t v = (uint64_t)a0 | ((uint64_t)a1 << n0);        // n0 <= 32
                        lo = (uint32_t)v; hi = (uint32_t)(v >> 32);
                        cnt = (n0 + n1) >> 1;
                    } else {
                        lo = wx; hi = wy; cnt = nvalid;
                        if (nvalid < 32) {
                            const uint64_t u = ((uint64_t)wx | ((uint64_t)wy << 32)) & ((1ULL << (2 * nvalid)) - 1ULL);
                            lo = (uint32_t)u; hi = (uint32_t)(u >> 32);
                        }
                    }
                    if (nvalid == 0) { lo = 0; hi = 0; cnt = 0; }
                } else {
                    // first step of a read that does not start on a 16-byte boundary of the packed buffer: base by base
                    uint64_t v = 0;
                    cnt = 0;
                    uint32_t pc = prev;
#pragma unroll 4
                    for (uint32_t j = 0; j < 32; j++) {
                        const uint32_t b = b0 + j;
                        const uint32_t c = ((j < 16 ? w.x >> (2 * j) : w.y >> (2 * j - 32)) & 3u);
                        if (b >= b_lo && b < b_hi && (!a.hpc || b == b_lo || c != pc)) {
                            v |= (uint64_t)c << (2 * cnt);
                            cnt++;
                        }
                        pc = c;
                    }
                    lo = (uint32_t)v; hi = (uint32_t)(v >> 32);
                }
                const uint32_t incl = warp_inclusive_scan(cnt);
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t bit = (2u * (avail + incl - cnt)) & (P_RING_WORDS * 32u - 1u);
                const uint32_t wi = bit >> 5, sh = bit & 31u;
                const uint32_t p0 = lo << sh, p1 = __funnelshift_l(lo, hi, sh), p2 = __funnelshift_l(hi, 0u, sh);
                // unconditional: OR-ing a zero word is harmless and cheaper than a branch around each atomic
                atomicOr(&ring[wi], p0);
                atomicOr(&ring[(wi + 1) & (P_RING_WORDS - 1)], p1);
                atomicOr(&ring[(wi + 2) & (P_RING_WORDS - 1)], p2);
                avail += total;
                step++;
                if (sub == P_STEPS_PER_TILE - 1 || step == n_steps) {      // the stage has been read completely: refill it
                    __syncwarp();
                    if (lane == 0 && tile + 2 < n_tiles) {
                        const uint32_t off = (tile + 2) * P_TILE_WORDS;
                        bulk_load(sm.stage[st], gsrc + off, min(tot_bytes - off * 4u, (uint32_t)P_TILE_WORDS * 4u), &sm.bar[st]);
                    }
                }
            }
            __syncwarp();                               // the ORs of every lane have landed before the ring is read

            // ---- roll + hash + select, 512 positions per step ------------------
            const bool final_ = (step == n_steps);
            const int pmax_final = (int)avail - L - 1;          // Kmer.hpp:1395: positions 1 .. n-2, n = L' - l + 1
            for (;;) {
                int pmax;
                if (avail >= done + (uint32_t)BLK + (uint32_t)L) pmax = 0x7fffffff;
                else if (final_ && (int)done <= pmax_final) pmax = pmax_final;
                else break;
                const uint32_t p0 = done + lane * 16;
                int nv = 16;
                if (pmax != 0x7fffffff) nv = max(0, min(16, pmax - (int)p0 + 1));
                uint32_t valid_bits = (1u << nv) - 1u;
                if (p0 == 0) valid_bits &= ~1u;                 // position 0 is trimmed (Kmer.hpp:1362,1395)
                const uint32_t wi = ((done >> 4) + lane) & (P_RING_WORDS - 1);
                const uint32_t lo = ring[wi], hi = ring[(wi + 1) & (P_RING_WORDS - 1)];
                uint32_t s_hi, s_lo;
                const uint32_t sel = k1v1::roll16_packed<L>(lo, hi, thr_cand, s_hi, s_lo) & valid_bits;
                const uint32_t hit_lanes = __ballot_sync(0xffffffffu, sel != 0);
                ring[wi] = 0;                                    // consumed: ready for the ORs of a later fill
                if (hit_lanes) {
                    const uint32_t n_c = __popc(sel);
                    if (__all_sync(0xffffffffu, n_c <= 1)) {     // common case: one store per hit lane
                        if (n_c) {
                            const uint32_t j = __ffs(sel) - 1;
                            cand[n_list + __popc(hit_lanes & ((1u << lane) - 1u))] =
                                make_uint2(p0 + j, k1v1::lmer_from_packed<L>(s_hi, s_lo, j));
                        }
                        n_list += __popc(hit_lanes);
                    } else {
                        const uint32_t incl = warp_inclusive_scan(n_c);
                        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                        const uint32_t before = incl - n_c;
                        for (uint32_t base_rank = 0; base_rank < total; base_rank += 32) {
                            uint32_t rest = sel, rank = before;
                            while (rest) {
                                const uint32_t j = __ffs(rest) - 1;
                                rest &= rest - 1;
                                if (rank >= base_rank && rank < base_rank + 32)
                                    cand[n_list + (rank - base_rank)] = make_uint2(p0 + j, k1v1::lmer_from_packed<L>(s_hi, s_lo, j));
                                rank++;
                            }
                            n_list += min(32u, total - base_rank);
                            __syncwarp();
                            if (n_list >= 32 && base_rank + 32 < total) {        // make room for the next 32
                                out_cnt += flush_candidates(a, cand, 32, lane, slot_lo, slot_cap, out_cnt);
                                __syncwarp();
                                uint2 keep_e = make_uint2(0, 0);
                                if (32 + lane < n_list) keep_e = cand[32 + lane];
                                __syncwarp();
                                if (32 + lane < n_list) cand[lane] = keep_e;
                                n_list -= 32;
                                __syncwarp();
                            }
                        }
                    }
                    __syncwarp();
                    if (n_list >= 32) {
                        out_cnt += flush_candidates(a, cand, 32, lane, slot_lo, slot_cap, out_cnt);
                        __syncwarp();
                        uint2 keep_e = make_uint2(0, 0);
                        if (32 + lane < n_list) keep_e = cand[32 + lane];
                        __syncwarp();
                        if (32 + lane < n_list) cand[lane] = keep_e;
                        n_list -= 32;
                        __syncwarp();
                    }
                }
                done += BLK;
            }
            __syncwarp();                               // the zeroing stores are ordered before the next fill's ORs
            if (final_) break;
        }
        if (n_list) {
            out_cnt += flush_candidates(a, cand, n_list, lane, slot_lo, slot_cap, out_cnt);
            n_list = 0;
            __syncwarp();
        }
        if (lane == 0) {
            a.n_min[r] = out_cnt;
            if ((uint64_t)out_cnt > slot_cap) atomicAdd(a.n_overflow, 1ULL);
        }
    }
}

bool sketch_packed_eligible(const SketchArgs& a) {
    const uint32_t t_hi = (uint32_t)(a.threshold >> 32);
    return a.l == 15 && !a.select_none && t_hi + k1v1::S1_SLACK >= k1v1::S1_SLACK;
}

template <int L_FAST, int V>
static void launch_sketch_as(const SketchArgs& a, int sm_count, cudaStream_t s) {
    // persistent grid: enough CTAs to fill every SM, reads are pulled dynamically
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (sketch_kernel<L_FAST, V>), WARPS_PER_CTA * 32, 0);
    if (per_sm < 1) per_sm = 1;
    uint64_t want = ((uint64_t)(a.read_end - a.read_begin) + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    uint64_t grid = (uint64_t)sm_count * per_sm;
    if (grid > want) grid = want;
    sketch_kernel<L_FAST, V><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(a);
}

static void launch_byte_ring(const SketchArgs& a, int sm_count, cudaStream_t s) {
    if (a.l != 15) launch_sketch_as<0, 0>(a, sm_count, s);          // generic l: no unrolled register block at all
    else if (a.variant == 0) launch_sketch_as<15, 0>(a, sm_count, s);
    else launch_sketch_as<15, 1>(a, sm_count, s);
}

int launch_sketch(const SketchArgs& a, int sm_count, cudaStream_t s) {
    if (a.read_end <= a.read_begin) return 0;
    if (a.variant == 2 && a.read_src && a.packed && a.dirty_list && sketch_packed_eligible(a)) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (sketch_packed_kernel<15>), WARPS_PER_CTA * 32, 0);
        if (per_sm < 1) per_sm = 1;
        const uint64_t want = ((uint64_t)(a.read_end - a.read_begin) + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        const uint64_t grid = std::min<uint64_t>((uint64_t)sm_count * per_sm, want);
        sketch_packed_kernel<15><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, s>>>(a);
        // the reads it could not take (ASCII spill of the host packer): byte-ring kernel over the list, a few CTAs
        SketchArgs b = a;
        b.variant = 1;
        b.read_list = a.dirty_list;
        b.read_list_n = a.dirty_count;
        b.cursor = a.dirty_cursor;
        b.read_begin = 0;                                   // list entries are absolute read indices
        b.read_end = a.n_reads;
        launch_byte_ring(b, sm_count, s);
        return 2;
    }
    launch_byte_ring(a, sm_count, s);
    return 1;
}

// ------------------------------------------------------------------ ASCII -> 2-bit device layout (variant 2's input)
// One warp per read, 16 bases -> one u32 per lane and step.  HBM-streaming: 1 B/bp in, 0.25 B/bp out.  A read holding
// any byte outside "ACGT" (N, IUPAC, lower case, '#') is left to the byte-ring kernel: read_src = SRC_ASCII | offset.
// 4 ASCII bytes -> their four 2-bit codes in byte 3 of the result (pack4_lsb before the shift), and in `diff` a
// non-zero byte for every input byte that is not one of A, C, G, T.  Exact: with y = x ^ 'A' the four letters are
// 00, 02, 06, 15 (hex), i.e. everything outside the two code bits must be 0x00 -- or 0x11 exactly for T (code 2).
__device__ __forceinline__ uint32_t pack4_checked(uint32_t x, uint32_t& diff) {
    const uint32_t s1 = x >> 1, s2 = x >> 2;
    const uint32_t t = s2 & ~s1 & 0x01010101u;                          // 1 where the code is 2 (T)
    diff |= ((x & 0xF9F9F9F9u) ^ (t * 0x11u)) ^ 0x41414141u;
    return (s1 & 0x03030303u) * 0x01041040u;                            // codes of bytes 0..3 in bits 24..31
}

__global__ void __launch_bounds__(256) pack_ascii_kernel(const PackArgsAscii a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = a.read_begin + warp; r < a.read_end; r += n_warps) {
        const uint64_t start = a.offsets[r], end = a.offsets[r + 1];
        const uint32_t len = (uint32_t)(end - start);
        const uint64_t w_off = pack_word_offset(start, r);
        const uint64_t byte0 = a.src_start ? a.src_start[r] : start;       // where the read's bytes are
        const uint8_t* base = a.bases + byte0;
        const uint32_t skip = (uint32_t)((uintptr_t)base & 15);
        const uint8_t* abase = base - skip;                       // 16-byte aligned
        // aligned 16-byte units of the text: unit m = bytes [16 m, 16 m + 16) of abase; the read is [skip, skip + len).
        // Every lane packs ONE aligned unit (coalesced 16-byte loads); output word j = bases [16 j, 16 j + 16) of the
        // read straddles units j and j + 1, so it is one funnel shift of the lane's word and its neighbour's: a warp
        // step loads 32 units and writes 31 words.
        const uint32_t x_end = skip + len, n_out = (len + 15) >> 4;
        uint32_t bad = 0;
        // aligned unit m of the text (zeros past the read / the buffer); loads of the NEXT step are issued before the
        // current one is packed, so two 16-byte loads per lane are in flight (the kernel is latency-bound otherwise)
        auto load_unit = [&](uint32_t m) -> uint4 {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (16 * (uint64_t)m < x_end) {
                const uint8_t* p = abase + 16 * (uint64_t)m;
                if (p + 16 <= a.bases_end) {
                    u = *reinterpret_cast<const uint4*>(p);
                } else {
                    uint32_t t[4] = {0, 0, 0, 0};
                    for (int j = 0; j < 16; j++)
                        if (p + j < a.bases_end) t[j >> 2] |= (uint32_t)p[j] << (8 * (j & 3));
                    u = make_uint4(t[0], t[1], t[2], t[3]);
                }
            }
            return u;
        };
        uint4 u_next = load_unit(lane);
        for (uint32_t m0 = 0; m0 < n_out; m0 += 31) {
            const uint32_t m = m0 + lane;
            const uint4 u = u_next;
            if (m0 + 31 < n_out) u_next = load_unit(m + 31);
            uint32_t word = 0, diff = 0;
            if (16 * m < x_end) {
                const uint32_t lo = m == 0 ? skip : 0u, hi = min(16u, x_end - 16 * m);
                if (lo == 0 && hi == 16) {
                    const uint32_t c0 = pack4_checked(u.x, diff), c1 = pack4_checked(u.y, diff);
                    const uint32_t c2 = pack4_checked(u.z, diff), c3 = pack4_checked(u.w, diff);
                    word = __byte_perm(__byte_perm(c0, c1, 0x0073), __byte_perm(c2, c3, 0x0073), 0x5410);
                } else {                                          // first / last unit: only bytes [lo, hi) are the read's
                    uint32_t d[4] = {0, 0, 0, 0};
                    const uint32_t c0 = pack4_checked(u.x, d[0]), c1 = pack4_checked(u.y, d[1]);
                    const uint32_t c2 = pack4_checked(u.z, d[2]), c3 = pack4_checked(u.w, d[3]);
                    word = __byte_perm(__byte_perm(c0, c1, 0x0073), __byte_perm(c2, c3, 0x0073), 0x5410);
                    for (uint32_t j = 0; j < 16; j++)
                        if (j >= lo && j < hi) diff |= (d[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                }
            }
            bad |= diff;
            const uint32_t next = __shfl_down_sync(0xffffffffu, word, 1);
            if (lane < 31 && m < n_out) {
                uint32_t out = skip ? __funnelshift_r(word, next, 2 * skip) : word;
                if (m == n_out - 1 && (len & 15u)) out &= (1u << (2 * (len & 15u))) - 1u;   // unused bits of the last word: zero
                a.packed[w_off + m] = out;
            }
        }
        const bool dirty = __any_sync(0xffffffffu, bad != 0);
        if (lane == 0) a.read_src[r] = dirty ? (SRC_ASCII | byte0) : w_off;
    }
}

void launch_pack_ascii(const PackArgsAscii& a, int sm_count, cudaStream_t s) {
    if (a.read_end <= a.read_begin) return;
    uint64_t blocks = ((uint64_t)(a.read_end - a.read_begin) + 7) / 8;
    if (blocks > (uint64_t)sm_count * 8) blocks = (uint64_t)sm_count * 8;
    pack_ascii_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------ shared-memory scrambler (verification aid)
// A CTA starts with whatever the previous CTA on that SM left in shared memory.  mdbg_ctx_autotune_sketch runs
// every kernel variant once more in many small launches with this kernel in between: it fills the shared memory
// of every SM with bytes that are neither base codes nor caught by the bit-2 "invalid character" test, so a
// variant whose result depends on ring bytes it never wrote is caught by the identity check instead of in
// production (that is how the first version of variant 1 failed: once per ~10^6 reads, only in multi-launch runs).
__global__ void __launch_bounds__(256) smem_scramble_kernel(uint32_t seed, uint32_t* sink) {
    __shared__ uint32_t buf[12 * 1024];                           // 48 KB static; 4 CTAs cover 192 KB of an SM
    for (uint32_t i = threadIdx.x; i < 12 * 1024; i += blockDim.x) {
        const uint32_t b = 0x18u + 0x20u * ((i + seed) % 7u);     // bit 2 clear, value >= 0x18 in every byte
        buf[i] = (b | (b << 8) | (b << 16) | (b << 24)) ^ ((i * 0x01010101u) & 0x03030303u);
    }
    __syncthreads();
    if (seed == 0xFFFFFFFFu) *sink = buf[threadIdx.x];             // never true: keeps the stores alive
}

void launch_smem_scramble(int sm_count, uint32_t seed, uint32_t* sink, cudaStream_t s) {
    const unsigned grid = (unsigned)sm_count * 4u;
    smem_scramble_kernel<<<grid, 256, 0, s>>>(seed & 0x7FFFFFFFu, sink);
}

// ------------------------------------------------------------------ byte-wise comparison of two device arrays
__global__ void __launch_bounds__(256) count_diff_kernel(const uint8_t* a, const uint8_t* b, size_t n, unsigned long long* n_diff) {
    unsigned long long mine = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        mine += a[i] != b[i];
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_diff, mine);
}

void launch_count_diff(const void* a, const void* b, size_t n_bytes, unsigned long long* n_diff, cudaStream_t s) {
    if (n_bytes == 0) return;
    const uint64_t grid = std::min<uint64_t>((n_bytes + 255) / 256, 148 * 8);
    count_diff_kernel<<<(unsigned)grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(a), reinterpret_cast<const uint8_t*>(b), n_bytes, n_diff);
}

// ------------------------------------------------------------------ scan
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_TILE = SCAN_T * SCAN_E;

size_t scan_scratch_elems(uint32_t n) { return (size_t)(n + SCAN_TILE - 1) / SCAN_TILE + 1; }

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total) {
    __shared__ uint64_t wsum[SCAN_T / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += t;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    uint64_t woff = 0, tot = 0;
    for (int i = 0; i < SCAN_T / 32; i++) {
        if (i < (int)wid) woff += wsum[i];
        tot += wsum[i];
    }
    __syncthreads();
    *total = tot;
    return woff + x - v;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_sums(const uint32_t* counts, uint32_t n, uint64_t* tile_sums) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_E;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_E; i++)
        if (base + i < n) s += counts[base + i];
    uint64_t tot;
    block_exclusive_scan_u64(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_prefix(uint64_t* tile_sums, uint32_t n_tiles) {
    uint64_t running = 0;
    for (uint32_t base = 0; base < n_tiles; base += SCAN_T) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = (i < n_tiles) ? tile_sums[i] : 0;
        uint64_t tot;
        const uint64_t ex = block_exclusive_scan_u64(v, &tot);
        if (i < n_tiles) tile_sums[i] = running + ex;
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = running;
}

__global__ void __launch_bounds__(SCAN_T) scan_write(const uint32_t* counts, uint32_t n, const uint64_t* tile_sums,
                                                      uint32_t n_tiles, uint64_t* offsets) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_E;
    uint32_t c[SCAN_E];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_E; i++) {
        c[i] = (base + i < n) ? counts[base + i] : 0;
        s += c[i];
    }
    uint64_t tot;
    uint64_t ex = block_exclusive_scan_u64(s, &tot) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_E; i++) {
        if (base + i < n) offsets[base + i] = ex;
        ex += c[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = tile_sums[n_tiles];
}

void launch_scan_u32_to_u64(const uint32_t* counts, uint64_t* offsets, uint32_t n, uint64_t* scratch, cudaStream_t s) {
    if (n == 0) {
        cudaMemsetAsync(offsets, 0, sizeof(uint64_t), s);
        return;
    }
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<<<n_tiles, SCAN_T, 0, s>>>(counts, n, scratch);
    scan_tile_prefix<<<1, SCAN_T, 0, s>>>(scratch, n_tiles);
    scan_write<<<n_tiles, SCAN_T, 0, s>>>(counts, n, scratch, n_tiles, offsets);
}

// ------------------------------------------------------------------ compact padded slots -> tight CSR
// (One warp per read.  The kernel moves ~0.5 GB out of 128-byte-line fragments of the padded slot arrays and writes 0.5 GB:
// it runs at DRAM speed -- two forms that batch the per-read metadata over 32 reads were measured on a B200 and were
// slower, 0.49 / 0.59 ms against 0.40.)
__global__ void __launch_bounds__(256) compact_kernel(const CompactArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i_read = warp; i_read < a.n_reads; i_read += n_warps) {
        const uint64_t r = a.read_begin + i_read;
        const uint64_t start = a.base_offsets[r], end = a.base_offsets[r + 1];
        const uint64_t slot_lo = (start >> a.cap_shift) + r * a.cap_const;
        const uint64_t slot_cap = ((end >> a.cap_shift) + (r + 1) * a.cap_const) - slot_lo;
        uint64_t n = a.n_min[r];
        if (n > slot_cap) n = slot_cap;                  // overflowed read: the caller re-runs in exact mode
        const uint64_t dst = a.tight_base + a.tight_off[i_read];
        for (uint64_t i = lane; i < n; i += 32) {
            a.out_min[dst + i] = a.in_min[slot_lo + i];
            a.out_pos[dst + i] = a.in_pos[slot_lo + i];
            a.out_dir[dst + i] = a.in_dir[slot_lo + i];
            if (a.out_qual) a.out_qual[dst + i] = a.in_qual[slot_lo + i];
        }
    }
}

void launch_compact(const CompactArgs& a, cudaStream_t s) {
    if (a.n_reads == 0) return;
    uint64_t blocks = ((uint64_t)a.n_reads + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    compact_kernel<<<(unsigned)blocks, 256, 0, s>>>(a);
}

__global__ void append_offsets_kernel(const uint64_t* batch_off, uint64_t* store_off, uint32_t n_reads,
                                      uint64_t dst_read_base, uint64_t dst_min_base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_reads) store_off[dst_read_base + 1 + i] = dst_min_base + batch_off[i + 1];
    if (i == 0 && dst_read_base == 0) store_off[0] = 0;
}

void launch_append_offsets(const uint64_t* batch_off, uint64_t* store_off, uint32_t n_reads, uint64_t dst_read_base,
                           uint64_t dst_min_base, cudaStream_t s) {
    if (n_reads == 0) return;
    append_offsets_kernel<<<(n_reads + 255) / 256, 256, 0, s>>>(batch_off, store_off, n_reads, dst_read_base,
                                                                dst_min_base);
}

}  // namespace mdbg
