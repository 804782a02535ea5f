// aux.cu -- row A3b: the per-read side outputs of ReadSelectionFunctor::operator()
// (src/readSelection/ReadSelection.hpp:870-920, 1047-1138, 1171-1228, 1302-1320):
//   * exact sum of the per-base error rates (the `long double errorSum` loop, :870-875) as a 128-bit
//     fixed-point integer -- the host side of the C ABI turns it into meanReadQuality with the same
//     long double / float / log10f operations the reference executes;
//   * computeSequenceComplexity(seq, 64, 32) (:1171-1228) in IEEE double, same operation order,
//     and the "score > 5 => clear the read's minimizers" filter (:894-903);
//   * per-minimizer minimum base quality over rlePositions[pos] .. rlePositions[pos+l] (:1135, :1302-1320).
// One warp per read; runs between the sketch kernel and the scan, on the padded slots.
#include "common.cuh"
#include "engine.cuh"

namespace mdbg {

constexpr int AUX_WARPS = 8;

__device__ __forceinline__ uint32_t nonzero_flags(uint32_t x) {
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t flags16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    const uint32_t a = (f0 * 0x00204081u) >> 28, b = (f1 * 0x00204081u) >> 24;
    const uint32_t c = (f2 * 0x00204081u) >> 20, d = (f3 * 0x00204081u) >> 16;
    return a | (b & 0xF0u) | (c & 0xF00u) | (d & 0xF000u);
}

__global__ void __launch_bounds__(AUX_WARPS * 32) read_aux_kernel(const AuxArgs a) {
    __shared__ uint8_t cnt_all[AUX_WARPS][32][64];        // 3-mer counters, one 64-byte row per lane
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t* cnt = cnt_all[wib][lane];
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;

    for (uint64_t r = warp; r < a.n_reads; r += n_warps) {
        const uint64_t start = a.offsets[r], end = a.offsets[r + 1];
        const uint32_t len = (uint32_t)(end - start);
        const uint8_t* seq = a.bases + start;

        // ---- (1) exact error sum ------------------------------------------------------------------
        if (a.quals) {
            const uint8_t* q = a.quals + start;
            uint64_t lo = 0, hi = 0;
            uint32_t lmin = 255;
            for (uint32_t i = lane; i < len; i += 32) {
                const uint32_t c = q[i];
                const uint64_t v = a.err_fixed[c];          // float error rate * 2^ERR_SHIFT, exact
                const uint64_t nlo = lo + v;
                hi += (nlo < lo) ? 1u : 0u;
                lo = nlo;
                lmin = min(lmin, (uint32_t)a.err_tz[c]);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const uint64_t olo = __shfl_xor_sync(0xffffffffu, lo, d), ohi = __shfl_xor_sync(0xffffffffu, hi, d);
                const uint64_t nlo = lo + olo;
                hi += ohi + ((nlo < lo) ? 1u : 0u);
                lo = nlo;
                lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, d));
            }
            if (lane == 0) {
                a.err_sum_lo[r] = lo;
                a.err_sum_hi[r] = hi;
                a.err_lmin[r] = (uint8_t)lmin;
            }
        }

        // ---- (2) sequence complexity --------------------------------------------------------------------
        // kmers[i] = direct 3-mer at i (KmerModelDirect, Kmer.hpp:701-870), i in [0, len-3]; windows start
        // at ii = 0, 32, 64, ... and use kmers[ii .. ii+64) when all 64 exist (:1188-1203).
        const uint32_t nk = len >= 3 ? len - 2 : 0;
        const uint32_t n_win = nk >= 64 ? (nk - 64) / 32 + 1 : 0;
        double acc = 0.0;
        for (uint32_t w0 = 0; w0 < n_win; w0 += 32) {
            const uint32_t w = w0 + lane;
            double score = 0.0;
            if (w < n_win) {
                uint4* z = reinterpret_cast<uint4*>(cnt);
                z[0] = z[1] = z[2] = z[3] = make_uint4(0, 0, 0, 0);
                const uint8_t* p = seq + (size_t)w * 32;
                uint32_t c0 = (p[0] >> 1) & 3, c1 = (p[1] >> 1) & 3;
                for (int i = 0; i < 64; i++) {
                    const uint32_t c2 = (p[i + 2] >> 1) & 3;
                    cnt[(c0 << 4) | (c1 << 2) | c2]++;
                    c0 = c1; c1 = c2;
                }
                uint32_t s2 = 0;                               // sum c*(c-1): even, so /2 is exact
                for (int v = 0; v < 64; v++) { const uint32_t c = cnt[v]; s2 += c * (c - 1); }
                // sa[i] = c*(c-1)/2.0 summed in double (all exact integers), then score /= (l-1), l = w-2 = 62
                score = __ddiv_rn((double)(s2 >> 1), 61.0);
            }
            const uint32_t lim = min(32u, n_win - w0);
            for (uint32_t t = 0; t < lim; t++) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, score, t));   // window order
        }
        const double complexity = __ddiv_rn(acc, (double)n_win);   // 0/0 = NaN for short reads: "NaN > 5" is false
        const bool low = complexity > 5.0;
        if (lane == 0) {
            a.complexity[r] = complexity;
            a.low_complexity[r] = low ? 1 : 0;
        }

        // slot of this read in the padded sketch output
        uint64_t slot_lo, slot_cap;
        if (a.exact_off) {
            slot_lo = a.exact_off[r];
            slot_cap = a.exact_off[r + 1] - slot_lo;
        } else {
            slot_lo = (start >> a.cap_shift) + r * a.cap_const;
            slot_cap = ((end >> a.cap_shift) + (r + 1) * a.cap_const) - slot_lo;
        }
        uint32_t nm = a.n_min[r];
        if (low && a.filter_low_complexity) {                  // ReadSelection.hpp:894-903
            if (lane == 0) a.n_min[r] = 0;
            nm = 0;
        }
        if (!a.out_qual) continue;
        if (nm > slot_cap) nm = (uint32_t)slot_cap;            // overflowed read: redone after the exact re-sketch
        if (nm == 0) continue;
        if (!a.quals) {                                        // ReadSelection.hpp:1049-1053: no qualities => 1
            for (uint32_t j = lane; j < nm; j += 32) a.out_qual[slot_lo + j] = 1;
            continue;
        }

        // ---- (3) raw coordinates of HPC positions pos and pos+l, then the min quality in between ------
        const uint32_t* pos = a.pad_pos + slot_lo;
        uint32_t* rawA = a.pad_raw_a + slot_lo;
        uint32_t* rawB = a.pad_raw_b + slot_lo;
        if (!a.hpc) {
            for (uint32_t j = lane; j < nm; j += 32) { rawA[j] = pos[j]; rawB[j] = pos[j] + a.l; }
        } else {
            // re-scan the read: keep mask per 16-byte lane piece, exactly as the sketch kernel's fill phase
            const uint32_t skip = (uint32_t)((uintptr_t)seq & 15);
            const uint8_t* abase = seq - skip;
            const uint32_t x_end = skip + len;
            const uint32_t n_chunks = (x_end + 511) / 512;
            uint32_t hpc_base = 0, carry = '#';
            uint32_t ja = 0, jb = 0;                           // next unresolved target in each sorted list
            for (uint32_t chunk = 0; chunk < n_chunks && (ja < nm || jb < nm); chunk++) {
                const uint32_t x0 = chunk * 512 + lane * 16;
                uint4 w = make_uint4(0, 0, 0, 0);
                if (x0 < x_end && x0 + 16 > skip) {
                    const uint8_t* p = abase + x0;
                    if (p + 16 <= a.bases_end) w = *reinterpret_cast<const uint4*>(p);
                    else {
                        uint32_t t[4] = {0, 0, 0, 0};
                        for (int j = 0; j < 16; j++)
                            if (p + j < a.bases_end) t[j >> 2] |= (uint32_t)p[j] << (8 * (j & 3));
                        w = make_uint4(t[0], t[1], t[2], t[3]);
                    }
                }
                const uint32_t lo = (skip > x0) ? min(skip - x0, 16u) : 0u;
                const uint32_t hi = (x_end > x0) ? min(x_end - x0, 16u) : 0u;
                const uint32_t vm = (hi > lo) ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                uint32_t pb = __shfl_up_sync(0xffffffffu, w.w >> 24, 1);
                if (lane == 0) pb = carry;
                carry = __shfl_sync(0xffffffffu, w.w >> 24, 31);
                uint32_t n0 = nonzero_flags(w.x ^ ((w.x << 8) | pb));
                uint32_t n1 = nonzero_flags(w.y ^ __funnelshift_l(w.x, w.y, 8));
                uint32_t n2 = nonzero_flags(w.z ^ __funnelshift_l(w.y, w.z, 8));
                uint32_t n3 = nonzero_flags(w.w ^ __funnelshift_l(w.z, w.w, 8));
                n0 &= nonzero_flags(w.x ^ 0x23232323u); n1 &= nonzero_flags(w.y ^ 0x23232323u);
                n2 &= nonzero_flags(w.z ^ 0x23232323u); n3 &= nonzero_flags(w.w ^ 0x23232323u);
                uint32_t k16 = flags16(n0, n1, n2, n3);
                if (skip >= x0 && skip < x0 + 16) {
                    const uint32_t sh = 8 * ((skip - x0) & 3);
                    const uint32_t wq = (skip - x0) < 4 ? w.x : (skip - x0) < 8 ? w.y : (skip - x0) < 12 ? w.z : w.w;
                    const bool is_hash = ((wq >> sh) & 0xFFu) == 0x23u;
                    k16 = is_hash ? (k16 & ~(1u << (skip - x0))) : (k16 | (1u << (skip - x0)));
                }
                const uint32_t keep = k16 & vm;
                const uint32_t c = __popc(keep);
                const uint32_t incl = warp_inclusive_scan(c);
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t hpc_end = hpc_base + total;
                // resolve every target that falls in [hpc_base, hpc_end): two passes (A list, B list)
                for (int which = 0; which < 2; which++) {
                    uint32_t& jn = which ? jb : ja;
                    for (;;) {
                        // each lane looks at target jn + lane
                        const uint32_t j = jn + lane;
                        uint32_t tgt = 0xFFFFFFFFu;
                        if (j < nm) tgt = pos[j] + (which ? a.l : 0u);
                        const bool mine = tgt < hpc_end;           // sorted: a prefix of lanes qualifies
                        const uint32_t mm = __ballot_sync(0xffffffffu, mine);
                        if (mm == 0) break;
                        // owner lane of my target: first lane with incl > tgt - hpc_base
                        const uint32_t rel = mine ? tgt - hpc_base : 0u;
                        uint32_t owner = 0;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) {
                            const uint32_t probe = owner + s - 1;
                            const uint32_t v = __shfl_sync(0xffffffffu, incl, probe & 31);
                            if (v <= rel) owner += s;
                        }
                        owner &= 31;
                        const uint32_t o_keep = __shfl_sync(0xffffffffu, keep, owner);
                        const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, owner);
                        const uint32_t o_cnt = __popc(o_keep);
                        if (mine) {
                            const uint32_t rank = rel - (o_incl - o_cnt);          // 0-based among the owner's kept bases
                            const uint32_t bit = __fns(o_keep, 0, rank + 1);
                            const uint32_t raw = chunk * 512 + owner * 16 + bit - skip;
                            (which ? rawB : rawA)[j] = raw;
                        }
                        const uint32_t n_res = __popc(mm);
                        jn += n_res;
                        if (n_res < 32) break;
                    }
                }
                hpc_base = hpc_end;
            }
        }
        __syncwarp();
        // getMinQuality (ReadSelection.hpp:1302-1320): u8 min of (qual[i] - 33) over [rawA, rawB)
        for (uint32_t j = lane; j < nm; j += 32) {
            const uint32_t ra = rawA[j], rb = rawB[j];
            uint8_t mq = 255;
            for (uint32_t i = ra; i < rb; i++) {
                const uint8_t q = (uint8_t)(a.quals[start + i] - 33);
                mq = q < mq ? q : mq;
            }
            a.out_qual[slot_lo + j] = mq;
        }
    }
}

void launch_read_aux(const AuxArgs& a, cudaStream_t s) {
    if (a.n_reads == 0) return;
    uint64_t blocks = ((uint64_t)a.n_reads + AUX_WARPS - 1) / AUX_WARPS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    read_aux_kernel<<<(unsigned)blocks, AUX_WARPS * 32, 0, s>>>(a);
}

}  // namespace mdbg
