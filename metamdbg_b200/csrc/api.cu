// api.cu -- the C ABI of libmdbg_b200 (include/mdbg_b200.h): context, device
// memory, stream plumbing and the host-side sequencing of the kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/mdbg_b200.h"
#include "engine.cuh"
#include "pack_host.hpp"
#include <functional>

using namespace mdbg;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int MAX_SUB = 64;       // H2D/compute pipeline depth of one host batch

struct SmallDev {                 // device-side scalars, one allocation
    uint32_t cursor;
    uint32_t sub_cursor[MAX_SUB];
    uint32_t dirty_n[1 + MAX_SUB];       // variant 2: reads left to the byte-ring kernel, whole batch [0] / per piece [1 + i]
    uint32_t dirty_cur[1 + MAX_SUB];     // ... and the list-mode cursors
    uint32_t full_flag;
    unsigned long long n_overflow;
    unsigned long long n_flagged;
    unsigned long long n_changed;
    unsigned long long emit_cursor;
    unsigned long long t_claims;         // claimed slots (= distinct keys) of the current count table
    TableStats stats;
};

}  // namespace

struct mdbg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int sketch_variant = 2;            // sketch kernel variant (sketch.cu): 2 = packed kernel (default), 0 / 1 = byte-ring
                                       // kernel with the two arithmetic forms of the l = 15 block; same results
    int host_packing = -1;             // 2-bit pack ASCII host batches before H2D: -1 auto, 0 off, 1 on, 2 hybrid
    HostPool* pool = nullptr;
    uint64_t last_direct_pieces = 0, last_pieces = 0;   // hybrid transfer statistics of the last host batch
    double last_pack_gbs = 0;          // host packer throughput of the last packed batch (raw GB/s)
    uint64_t last_pipelined = 0, last_grows = 0;   // piece pipeline: pieces finished piece-wise, mid-batch buffer growths
    int last_overflow_fallback = 0, last_packed = 0;
    int auto_pack_pause = 0;           // auto mode: batches still to send as ASCII after the packer proved too slow
    DevBuf d_pack, d_src, d_dirty;
    DevBuf pg_counts, pg_flags, pg_keyidx, pg_off, pg_readof, pg_hash, pg_koff, pg_reads, pg_wins;   // postings index
    PinBuf hpg_hash, hpg_koff, hpg_reads, hpg_wins;
    DevBuf r_table, r_hist, r_out;                                // repetitive-minimizer histogram (K1b)
    std::vector<uint32_t> r_sel_min, r_sel_cnt;
    DevBuf f_raw, f_cnt, f_off, f_nl, f_start, f_len, f_qstart;   // FASTQ / FASTA text ingest (mdbg_sketch_fastx)
    PinBuf h_pack, h_src, h_asc;
    cudaEvent_t sub_ev[MAX_SUB] = {};
    cudaEvent_t copy_gate = nullptr;
    // piece pipeline of a host batch: per-piece scan / compaction / D2H of the CSR behind the sketch of that piece
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t tot_ev[MAX_SUB] = {}, cmp_ev[MAX_SUB] = {};
    uint64_t* h_piece = nullptr;       // pinned, [MAX_SUB][2]: minimizers of the piece, overflow counter snapshot
    DevBuf loc_off;                    // piece-local exclusive offsets: u64 [n_reads + n_pieces]
    bool piece_pipeline = true;        // MDBG_PIECE_PIPELINE=0 restores the whole-batch tail
    std::string error;
    uint64_t launches = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;   // bytes enqueued across PCIe by the host-buffer entry points
    bool timing = false;
    cudaEvent_t ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool ev_valid[2] = {false, false};

    // parameters
    uint32_t l = 15;
    float density = 0.005f;
    uint32_t hpc = 1;
    uint64_t threshold = 0;
    uint32_t select_none = 0;
    DevBuf d_blacklist;
    uint32_t n_blacklist = 0;
    uint32_t cap_shift = 5, cap_const = 32;

    // small device scalars + pinned mirror
    SmallDev* d_small = nullptr;
    SmallDev* h_small = nullptr;   // pinned
    uint64_t* h_scalar = nullptr;  // pinned, 8 u64

    // last sketch batch
    DevBuf d_bases, d_offsets;                  // host-API staging
    DevBuf pad_min, pad_pos, pad_dir, n_min, scan_scratch;
    DevBuf b_off, b_min, b_pos, b_dir;          // tight CSR of the batch
    uint32_t b_reads = 0;
    uint64_t b_total = 0;
    PinBuf h_off, h_min, h_pos, h_dir;
    bool host_csr_valid = false;                // the piece pipeline already copied this batch's CSR into h_*

    // read side outputs (row A3b)
    bool filter_low_complexity = false;
    DevBuf d_quals, pad_qual, pad_raw_a, pad_raw_b, b_qual, x_sum_lo, x_sum_hi, x_lmin, x_cplx, x_low, d_err_fixed, d_err_tz;
    PinBuf hx_sum_lo, hx_sum_hi, hx_lmin, hx_cplx, hx_low, hx_meanq, h_qual;
    bool aux_valid = false, aux_has_quals = false;

    // minimizer-space read store
    DevBuf s_min, s_off, s_rem;
    uint64_t s_reads = 0, s_mins = 0;
    DevBuf p_flags, p_keep, p_cnt, p_newoff, p_newmin;

    // count table
    DevBuf table;
    uint64_t t_capacity = 0;
    uint32_t t_k = 0;
    bool t_active = false;
    bool t_value_mode = false;         // the table was filled by the next-k pass: `count` is a value, not an occurrence count
    bool t_merged = false;             // multi-rank: mdbg_count_merge has moved every key to its owner
    // The table is sized for the EXPECTED number of distinct k-min-mers (load factor <= 0.6), not for the worst case
    // "every window distinct": a pass that outgrows it is abandoned early (kernels watch the claim counter) and the
    // table is rebuilt 4x larger from the store ranges inserted so far -- both kinds of pass are reproducible from
    // the store.  distinct_ratio remembers distinct keys per stored minimizer of the last first-pass table, so
    // steady-state batches are sized right the first time.
    struct Range { uint64_t read_lo, read_hi; };
    std::vector<Range> t_ranges;       // store ranges inserted into the current table (rebuild recipe)
    bool t_rebuildable = false;        // false once foreign vectors were merged in
    bool t_autogrow = true;            // MDBG_TABLE_AUTOGROW=0: report MDBG_ERR_TABLE_FULL instead of rebuilding (tests)
    uint64_t t_claim_limit = 0;
    double distinct_ratio = 0.0;
    uint64_t n_dev_allocs = 0;         // ensure(): device buffers (re)allocated so far
    uint64_t n_table_trades = 0;       // table_reset: times the two table buffers traded places instead of a reallocation
    bool phase_prof = false;
    double phase_ms[16] = {};
    uint64_t store_gen = 0, rem_gen = ~0ull;   // s_rem (minimizers left in the read) is valid for this store generation
    uint32_t prev_min_count = 0;       // lookup-time filter of the previous-k table (see mdbg_prev_from_current)
    // Per-position values of the last whole-store pass (kminmer.cu, next_k_stream_kernel): when the previous-k table is
    // exactly that pass's table, the next pass needs no lookups -- and, with several ranks, no replicated table.
    DevBuf val_a, val_b, prev_src;
    bool val_valid = false;            // val_a holds the values of pass val_k over store generation val_gen
    uint32_t val_k = 0;
    uint64_t val_gen = 0;
    bool t_whole = false;              // the current table = ONE pass over the whole store (nothing else inserted)
    bool t_no_vecs = false;            // keys-only merge: the slots carry no k-min-mer vectors (no kminmers output, no edges)
    uint32_t prev_k = 0;               // k of the table that became the previous-k table
    bool prev_pure = false;            // ... and it is that pass's table as it was (no host pairs loaded or patched in)
    bool prev_replicated = true;       // several ranks: prev_table holds every rank's entries (false: only the owned ones)
    DevBuf foreign_vecs;
    uint64_t foreign_n = 0;
    DevBuf prev_table, prev_stage_h, prev_stage_a, rescue_table, pass_aux;
    DevBuf edge_table, edge_vals, o_edge_vals;
    PinBuf ho_edge_vals;
    uint64_t edge_cap = 0;             // slots of the device-resident edge set left by the last edges / unitigs call
    // unitigs (mdbg_unitigs_build)
    DevBuf u_slot_node, u_node_slot, u_next, u_pair, u_len, u_size, u_flag, u_cychead, u_seqoff, u_idx, u_cyclist, u_cycpos, u_best,
        u_jump, u_mins, u_off, u_hash, u_rev, u_circ, u_abund, u_bcnt, u_boff, u_order, u_pos, u_scnt, u_soff, u_ents, u_ecnt, u_eoff, u_etgt;
    PinBuf hu_mins, hu_off, hu_hash, hu_circ, hu_order, hu_abund, hu_eoff, hu_etgt;
    uint64_t prev_capacity = 0;
    DevBuf o_hash, o_abund, o_vecs;
    PinBuf ho_hash, ho_abund, ho_vecs;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, n_ranks = 1;
    DevBuf m_send_vecs, m_send_counts, m_recv_vecs, m_recv_counts, m_bucket;
};

namespace {

// Optional per-phase wall-clock profile of the table / collective paths (mdbg_ctx_phase_profile): when enabled, every
// phase boundary synchronises the stream, so the figures are exclusive phase times (and the run is slower).
enum Phase { PH_MERGE_PACK_COUNT, PH_MERGE_PLAN, PH_MERGE_PACK_SCATTER, PH_MERGE_EXCHANGE, PH_MERGE_INSERT,
             PH_PREV_EMIT, PH_PREV_PLAN, PH_PREV_EXCHANGE, PH_PREV_INSERT, PH_PASS, PH_STATS, PH_EMIT, PH_TABLE_RESET, PH_N };
const char* const kPhaseNames[PH_N] = {"merge.pack_count", "merge.plan_allgather", "merge.pack_scatter", "merge.exchange",
                                       "merge.insert", "prev.stats_emit", "prev.plan_allgather", "prev.exchange",
                                       "prev.insert", "pass.insert_or_next_k", "table.stats", "table.emit", "table.reset"};
struct PhaseClock {
    mdbg_ctx* c;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseClock(mdbg_ctx* ctx);
    void lap(Phase p);
};

mdbg_status fail(mdbg_ctx* c, mdbg_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->error = buf; else g_create_error = buf;
    return st;
}

PhaseClock::PhaseClock(mdbg_ctx* ctx) : c(ctx) {
    if (c->phase_prof) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); }
}
void PhaseClock::lap(Phase p) {
    if (!c->phase_prof) return;
    cudaStreamSynchronize(c->stream);
    const auto t1 = std::chrono::steady_clock::now();
    c->phase_ms[p] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? MDBG_ERR_OOM : MDBG_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CKS(expr)                          \
    do {                                   \
        mdbg_status s_ = (expr);           \
        if (s_ != MDBG_OK) return s_;      \
    } while (0)

mdbg_status ensure(mdbg_ctx* ctx, DevBuf& b, size_t bytes, bool keep = false) {
    if (bytes <= b.cap && b.p) return MDBG_OK;
    size_t want = bytes + bytes / 8 + 256;
    void* np = nullptr;
    CK(cudaMalloc(&np, want));
    ctx->n_dev_allocs++;
    static const bool trace = getenv("MDBG_TRACE_ALLOC") != nullptr;        // which buffer grew (offset inside the context)
    if (trace)
        fprintf(stderr, "[mdbg alloc] ctx %p buffer +%zu: %zu -> %zu bytes\n", (void*)ctx,
                (size_t)(reinterpret_cast<char*>(&b) - reinterpret_cast<char*>(ctx)), b.cap, want);
    if (keep && b.p && b.cap) {
        CK(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (b.p) cudaFree(b.p);
    b.p = np;
    b.cap = want;
    return MDBG_OK;
}

mdbg_status ensure_pin(mdbg_ctx* ctx, PinBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return MDBG_OK;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CK(cudaMallocHost(&b.p, want));
    b.cap = want;
    return MDBG_OK;
}

void release(DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }
void release(PinBuf& b) { if (b.p) cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }

// Largest T with (double)T < bound (reference compares u64 < double, Kmer.hpp:1434).
uint64_t threshold_from_density(float density, uint32_t* none) {
    const uint64_t max_hash = (uint64_t)-1;
    const double bound = (double)density * (double)max_hash;       // Kmer.hpp:1354-1356
    *none = 0;
    if (!((double)(uint64_t)0 < bound)) { *none = 1; return 0; }
    uint64_t lo = 0, hi = (uint64_t)-1;
    if ((double)hi < bound) return hi;
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if ((double)mid < bound) lo = mid; else hi = mid;
    }
    return lo;
}

uint64_t pow2ceil(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

mdbg_status check_launch(mdbg_ctx* ctx, const char* what, int n_kernels) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, MDBG_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    ctx->launches += (uint64_t)n_kernels;
    return MDBG_OK;
}

// ---- sketch of a device-resident batch into the tight CSR b_* -------------------
// error-rate table of ReadSelection::execute (ReadSelection.hpp:101-104): pow(10.0f, -q/10.0f) as float, handed
// to the device as exact integers (value * 2^ERR_SHIFT)
mdbg_status ensure_err_table(mdbg_ctx* ctx) {
    if (ctx->d_err_fixed.p) return MDBG_OK;
    uint64_t fixed[256];
    uint8_t tz[256];
    for (int c = 0; c < 256; c++) {
        float e = 0.0f;
        if (c >= 33 && c <= 127) { float q = (float)(uint8_t)(c - 33); e = powf(10.0f, -q / 10.0f); }   // Commons.hpp:2338-2341
        const long double scaled = ldexpl((long double)e, ERR_SHIFT);
        fixed[c] = (uint64_t)scaled;
        tz[c] = fixed[c] ? (uint8_t)__builtin_ctzll(fixed[c]) : 255;
    }
    CKS(ensure(ctx, ctx->d_err_fixed, sizeof fixed));
    CKS(ensure(ctx, ctx->d_err_tz, sizeof tz));
    CK(cudaMemcpy(ctx->d_err_fixed.p, fixed, sizeof fixed, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_err_tz.p, tz, sizeof tz, cudaMemcpyHostToDevice));
    return MDBG_OK;
}

mdbg_status run_aux(mdbg_ctx* ctx, const uint8_t* d_bases, const uint8_t* d_quals, const uint64_t* d_offsets,
                    uint32_t n_reads, uint64_t n_bases, const uint64_t* exact_off, const uint32_t* pos, uint8_t* qual_out) {
    AuxArgs x{};
    x.bases = d_bases; x.bases_end = d_bases + n_bases; x.quals = d_quals;
    x.offsets = d_offsets; x.n_reads = n_reads; x.l = ctx->l; x.hpc = ctx->hpc;
    x.err_fixed = ctx->d_err_fixed.as<uint64_t>(); x.err_tz = ctx->d_err_tz.as<uint8_t>();
    x.exact_off = exact_off; x.cap_shift = ctx->cap_shift; x.cap_const = ctx->cap_const;
    x.n_min = ctx->n_min.as<uint32_t>(); x.pad_pos = pos;
    x.pad_raw_a = ctx->pad_raw_a.as<uint32_t>(); x.pad_raw_b = ctx->pad_raw_b.as<uint32_t>();
    x.out_qual = qual_out;
    x.err_sum_lo = ctx->x_sum_lo.as<uint64_t>(); x.err_sum_hi = ctx->x_sum_hi.as<uint64_t>();
    x.err_lmin = ctx->x_lmin.as<uint8_t>(); x.complexity = ctx->x_cplx.as<double>();
    x.low_complexity = ctx->x_low.as<uint8_t>();
    x.filter_low_complexity = ctx->filter_low_complexity ? 1 : 0;
    launch_read_aux(x, ctx->stream);
    return check_launch(ctx, "read_aux_kernel", 1);
}

struct SubRange { uint32_t r0, r1; };

// Per-piece tail of a host batch.  A host batch reaches the device in pieces (sketch_host_batch); instead of
// waiting for the last piece before the scan / compaction / device-to-host copy of the minimizer CSR start, every
// piece gets its own scan right behind its sketch launch, and LAG pieces later -- when its total has long landed
// in pinned memory, so the host does not block -- its compaction into the tight CSR at the running offset and, if
// the caller wants the CSR on the host, its D2H copy on a third stream.  What is left after the last piece is
// one piece's worth of work.  A read that overflows its padded slot (low-complexity sequence) cancels the
// pipeline: the whole batch then takes the classic tail (global scan, exact re-sketch), results unchanged.
struct PiecePipeline {
    static constexpr size_t LAG = 2;
    mdbg_ctx* ctx;
    const uint64_t* d_offsets;           // device: read byte offsets of the batch
    const std::vector<SubRange>* subs;
    uint32_t n_reads;
    bool fetch;                          // also copy the CSR to the pinned host arrays h_off / h_min / h_pos / h_dir
    uint64_t run_total = 0;
    size_t launched_n = 0, finished_n = 0;
    bool overflow = false;
    bool in_flight = false;              // begin() is over: growth from here on happens with copies in flight

    mdbg_status grow_dev(DevBuf& b, size_t bytes) {
        if (bytes <= b.cap && b.p) return MDBG_OK;
        CK(cudaStreamSynchronize(ctx->d2h_stream));               // a copy may still read the old allocation
        if (in_flight) ctx->last_grows++;
        return ensure(ctx, b, in_flight ? 2 * bytes : bytes, true);   // the estimate was low: double, do not creep
    }
    mdbg_status grow_pin(PinBuf& b, size_t bytes, size_t keep_bytes) {
        if (bytes <= b.cap && b.p) return MDBG_OK;
        CK(cudaStreamSynchronize(ctx->d2h_stream));
        if (in_flight) ctx->last_grows++;
        PinBuf bigger;
        CKS(ensure_pin(ctx, bigger, in_flight ? 2 * bytes : bytes));
        if (b.p && keep_bytes) memcpy(bigger.p, b.p, keep_bytes);
        release(b);
        b = bigger;
        return MDBG_OK;
    }
    // sizes everything for the expected number of minimizers, so that growing mid-batch is the exception
    mdbg_status begin(uint64_t n_bases) {
        const uint64_t est = (uint64_t)((double)n_bases * (double)ctx->density * 1.25) + 1024;
        CKS(ensure(ctx, ctx->loc_off, ((size_t)n_reads + subs->size() + 1) * sizeof(uint64_t)));
        CKS(grow_dev(ctx->b_min, (est + 1) * 4));
        CKS(grow_dev(ctx->b_pos, (est + 1) * 4));
        CKS(grow_dev(ctx->b_dir, est + 1));
        if (fetch) {
            CKS(grow_pin(ctx->h_off, ((size_t)n_reads + 1) * 8, 0));
            CKS(grow_pin(ctx->h_min, (est + 1) * 4, 0));
            CKS(grow_pin(ctx->h_pos, (est + 1) * 4, 0));
            CKS(grow_pin(ctx->h_dir, est + 1, 0));
            ctx->h_off.as<uint64_t>()[0] = 0;
        }
        in_flight = true;
        ctx->last_pipelined = ctx->last_grows = 0;
        ctx->last_overflow_fallback = 0;
        return MDBG_OK;
    }
    // call right after the sketch launch of piece i (pieces are launched in order)
    mdbg_status launched(size_t i) {
        cudaStream_t s = ctx->stream;
        const SubRange& sr = (*subs)[i];
        const uint32_t np = sr.r1 - sr.r0;
        uint64_t* loc = ctx->loc_off.as<uint64_t>() + sr.r0 + i;                  // np + 1 entries
        launch_scan_u32_to_u64(ctx->n_min.as<uint32_t>() + sr.r0, loc, np, ctx->scan_scratch.as<uint64_t>(), s);
        CKS(check_launch(ctx, "scan", np ? 3 : 0));
        CK(cudaMemcpyAsync(&ctx->h_piece[2 * i], loc + np, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_piece[2 * i + 1], &ctx->d_small->n_overflow, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->tot_ev[i], s));
        launched_n = i + 1;
        while (finished_n + LAG < launched_n) CKS(finish(finished_n));
        return MDBG_OK;
    }
    mdbg_status finish(size_t j) {
        cudaStream_t s = ctx->stream, ds = ctx->d2h_stream;
        finished_n = j + 1;
        CK(cudaEventSynchronize(ctx->tot_ev[j]));
        const uint64_t total = ctx->h_piece[2 * j];
        if (ctx->h_piece[2 * j + 1] != 0) overflow = true;
        if (overflow) { ctx->last_overflow_fallback = 1; return MDBG_OK; }        // the classic tail redoes the batch
        ctx->last_pipelined++;
        const SubRange& sr = (*subs)[j];
        const uint32_t np = sr.r1 - sr.r0;
        const uint64_t* loc = ctx->loc_off.as<uint64_t>() + sr.r0 + j;
        CKS(grow_dev(ctx->b_min, (run_total + total + 1) * 4));
        CKS(grow_dev(ctx->b_pos, (run_total + total + 1) * 4));
        CKS(grow_dev(ctx->b_dir, run_total + total + 1));
        CompactArgs c{};
        c.base_offsets = d_offsets;
        c.cap_shift = ctx->cap_shift;
        c.cap_const = ctx->cap_const;
        c.n_min = ctx->n_min.as<uint32_t>();
        c.tight_off = loc;
        c.tight_base = run_total;
        c.read_begin = sr.r0;
        c.n_reads = np;
        c.in_min = ctx->pad_min.as<uint32_t>();
        c.in_pos = ctx->pad_pos.as<uint32_t>();
        c.in_dir = ctx->pad_dir.as<uint8_t>();
        c.out_min = ctx->b_min.as<uint32_t>();
        c.out_pos = ctx->b_pos.as<uint32_t>();
        c.out_dir = ctx->b_dir.as<uint8_t>();
        launch_compact(c, s);
        CKS(check_launch(ctx, "compact_kernel", np ? 1 : 0));
        // b_off[r0 + 1 + t] = run_total + loc[t + 1]  (and b_off[0] = 0 with the first piece)
        launch_append_offsets(loc, ctx->b_off.as<uint64_t>(), np, sr.r0, run_total, s);
        CKS(check_launch(ctx, "append_offsets_kernel", np ? 1 : 0));
        if (fetch) {
            CKS(grow_pin(ctx->h_min, (run_total + total + 1) * 4, run_total * 4));
            CKS(grow_pin(ctx->h_pos, (run_total + total + 1) * 4, run_total * 4));
            CKS(grow_pin(ctx->h_dir, run_total + total + 1, run_total));
            CK(cudaEventRecord(ctx->cmp_ev[j], s));
            CK(cudaStreamWaitEvent(ds, ctx->cmp_ev[j], 0));
            if (np)
                CK(cudaMemcpyAsync(ctx->h_off.as<uint64_t>() + sr.r0 + 1, ctx->b_off.as<uint64_t>() + sr.r0 + 1, (size_t)np * 8,
                                   cudaMemcpyDeviceToHost, ds));
            if (total) {
                CK(cudaMemcpyAsync(ctx->h_min.as<uint32_t>() + run_total, ctx->b_min.as<uint32_t>() + run_total, total * 4,
                                   cudaMemcpyDeviceToHost, ds));
                CK(cudaMemcpyAsync(ctx->h_pos.as<uint32_t>() + run_total, ctx->b_pos.as<uint32_t>() + run_total, total * 4,
                                   cudaMemcpyDeviceToHost, ds));
                CK(cudaMemcpyAsync(ctx->h_dir.as<uint8_t>() + run_total, ctx->b_dir.as<uint8_t>() + run_total, total,
                                   cudaMemcpyDeviceToHost, ds));
            }
            ctx->d2h_bytes += (uint64_t)np * 8 + total * 9;
        }
        run_total += total;
        return MDBG_OK;
    }
    // after the last piece: the remaining tails; the caller then checks `overflow`
    mdbg_status drain() {
        while (finished_n < launched_n) CKS(finish(finished_n));
        return MDBG_OK;
    }
};
// A feeder enqueues the sketch launches itself (host batches arriving in pieces); it may edit the argument block
// (input pointers) and must leave read_begin/read_end/cursor covering the whole batch when it returns.
using Feeder = std::function<mdbg_status(SketchArgs&)>;


// Variant 2 on ASCII bytes that are already in HBM (device batches, host batches that travelled unpacked): buffers
// for the 2-bit copy, and the streaming pack pass over a read range (launched right in front of that range's sketch).
mdbg_status device_pack_prepare(mdbg_ctx* ctx, SketchArgs& a, uint32_t n_reads, uint64_t n_bases) {
    CKS(ensure(ctx, ctx->d_pack, pack_words_capacity(n_bases, n_reads) * 4 + 64));
    CKS(ensure(ctx, ctx->d_src, (size_t)n_reads * 8));
    a.packed = ctx->d_pack.as<uint32_t>();
    a.read_src = ctx->d_src.as<uint64_t>();
    return MDBG_OK;
}
mdbg_status device_pack_range(mdbg_ctx* ctx, const SketchArgs& a, uint32_t r0, uint32_t r1, cudaStream_t s) {
    PackArgsAscii pa{};
    pa.bases = a.bases; pa.bases_end = a.bases_end; pa.offsets = a.offsets;
    pa.read_begin = r0; pa.read_end = r1;
    pa.packed = ctx->d_pack.as<uint32_t>(); pa.read_src = ctx->d_src.as<uint64_t>();
    launch_pack_ascii(pa, ctx->sm_count, s);
    return check_launch(ctx, "pack_ascii_kernel", r1 > r0 ? 1 : 0);
}
// per-piece launches of a host batch: each piece has its own cursor, dirty list segment and counters
void piece_args(mdbg_ctx* ctx, SketchArgs& a, size_t i, uint32_t r0, uint32_t r1) {
    a.read_begin = r0;
    a.read_end = r1;
    a.cursor = &ctx->d_small->sub_cursor[i];
    a.dirty_list = ctx->d_dirty.as<uint32_t>() + r0;
    a.dirty_count = &ctx->d_small->dirty_n[1 + i];
    a.dirty_cursor = &ctx->d_small->dirty_cur[1 + i];
}

mdbg_status sketch_internal(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                            uint64_t n_bases, int append, bool want_aux = false, const uint8_t* d_quals = nullptr,
                            const Feeder* feeder = nullptr, PiecePipeline* pipe = nullptr) {
    cudaStream_t s = ctx->stream;
    ctx->b_reads = n_reads;
    ctx->b_total = 0;
    ctx->aux_valid = false;
    ctx->host_csr_valid = false;
    CKS(ensure(ctx, ctx->b_off, ((size_t)n_reads + 1) * sizeof(uint64_t)));
    if (n_reads == 0) {
        CK(cudaMemsetAsync(ctx->b_off.p, 0, sizeof(uint64_t), s));
        return MDBG_OK;
    }
    const uint64_t pad_cap = (n_bases >> ctx->cap_shift) + (uint64_t)n_reads * ctx->cap_const + 1;
    CKS(ensure(ctx, ctx->pad_min, pad_cap * 4));
    CKS(ensure(ctx, ctx->pad_pos, pad_cap * 4));
    CKS(ensure(ctx, ctx->pad_dir, pad_cap));
    CKS(ensure(ctx, ctx->n_min, (size_t)n_reads * 4));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems(n_reads) * sizeof(uint64_t)));

    SketchArgs a{};
    a.bases = d_bases;
    a.bases_end = d_bases + n_bases;
    a.offsets = d_offsets;
    a.n_reads = n_reads;
    a.read_begin = 0;
    a.read_end = n_reads;
    a.l = ctx->l;
    a.hpc = ctx->hpc;
    a.threshold = ctx->threshold;
    a.select_none = ctx->select_none;
    a.blacklist = ctx->d_blacklist.as<uint32_t>();
    a.n_blacklist = ctx->n_blacklist;
    a.exact_off = nullptr;
    a.cap_shift = ctx->cap_shift;
    a.cap_const = ctx->cap_const;
    a.out_min = ctx->pad_min.as<uint32_t>();
    a.out_pos = ctx->pad_pos.as<uint32_t>();
    a.out_dir = ctx->pad_dir.as<uint8_t>();
    a.n_min = ctx->n_min.as<uint32_t>();
    a.cursor = &ctx->d_small->cursor;
    a.n_overflow = &ctx->d_small->n_overflow;
    a.variant = (uint32_t)ctx->sketch_variant;
    CKS(ensure(ctx, ctx->d_dirty, (size_t)n_reads * 4));
    a.dirty_list = ctx->d_dirty.as<uint32_t>();
    a.dirty_count = &ctx->d_small->dirty_n[0];
    a.dirty_cursor = &ctx->d_small->dirty_cur[0];

    // cursor, sub_cursor[], dirty_n[], dirty_cur[] are adjacent
    CK(cudaMemsetAsync(&ctx->d_small->cursor, 0, sizeof(uint32_t) * (1 + MAX_SUB + 2 * (1 + MAX_SUB)), s));
    CK(cudaMemsetAsync(&ctx->d_small->n_overflow, 0, sizeof(unsigned long long), s));
    if (!feeder && d_bases && a.variant == 2 && sketch_packed_eligible(a)) {
        // ASCII batch resident in HBM: one streaming pass turns it into the 2-bit layout the packed kernel reads
        CKS(device_pack_prepare(ctx, a, n_reads, n_bases));
        CKS(device_pack_range(ctx, a, 0, n_reads, s));
    }
    if (feeder) {
        if (pipe) CKS(pipe->begin(n_bases));
        CKS((*feeder)(a));
        if (pipe) CKS(pipe->drain());
        a.read_begin = 0;
        a.read_end = n_reads;
        a.cursor = &ctx->d_small->cursor;
        a.dirty_list = ctx->d_dirty.as<uint32_t>();
        a.dirty_count = &ctx->d_small->dirty_n[0];
        a.dirty_cursor = &ctx->d_small->dirty_cur[0];
    } else {
        if (ctx->timing) CK(cudaEventRecord(ctx->ev[0][0], s));
        const int n_k = launch_sketch(a, ctx->sm_count, s);
        if (ctx->timing) { CK(cudaEventRecord(ctx->ev[0][1], s)); ctx->ev_valid[0] = true; }
        CKS(check_launch(ctx, "sketch_kernel", n_k + 0));
    }
    if (want_aux) {
        CKS(ensure_err_table(ctx));
        CKS(ensure(ctx, ctx->pad_qual, pad_cap));
        CKS(ensure(ctx, ctx->pad_raw_a, pad_cap * 4));
        CKS(ensure(ctx, ctx->pad_raw_b, pad_cap * 4));
        CKS(ensure(ctx, ctx->x_sum_lo, (size_t)n_reads * 8));
        CKS(ensure(ctx, ctx->x_sum_hi, (size_t)n_reads * 8));
        CKS(ensure(ctx, ctx->x_lmin, n_reads));
        CKS(ensure(ctx, ctx->x_cplx, (size_t)n_reads * 8));
        CKS(ensure(ctx, ctx->x_low, n_reads));
        CKS(run_aux(ctx, d_bases, d_quals, d_offsets, n_reads, n_bases, nullptr, ctx->pad_pos.as<uint32_t>(),
                    ctx->pad_qual.as<uint8_t>()));
        ctx->aux_valid = true;
        ctx->aux_has_quals = d_quals != nullptr;
    }
    const bool piecewise_done = pipe && !pipe->overflow;          // every piece already sits in the tight CSR
    uint64_t total = 0, n_over = 0;
    if (piecewise_done) {
        total = pipe->run_total;
        ctx->b_total = total;
    } else {
        if (pipe) {                                                // cancelled by a slot overflow: classic tail, no D2H yet
            CK(cudaStreamSynchronize(ctx->d2h_stream));
            pipe->fetch = false;
        }
        launch_scan_u32_to_u64(ctx->n_min.as<uint32_t>(), ctx->b_off.as<uint64_t>(), n_reads,
                               ctx->scan_scratch.as<uint64_t>(), s);
        CKS(check_launch(ctx, "scan", 3));
        CK(cudaMemcpyAsync(&ctx->h_scalar[0], ctx->b_off.as<uint64_t>() + n_reads, sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_scalar[1], &ctx->d_small->n_overflow, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        total = ctx->h_scalar[0];
        n_over = ctx->h_scalar[1];
        ctx->b_total = total;
        CKS(ensure(ctx, ctx->b_min, (total + 1) * 4));
        CKS(ensure(ctx, ctx->b_pos, (total + 1) * 4));
        CKS(ensure(ctx, ctx->b_dir, total + 1));
        if (want_aux) CKS(ensure(ctx, ctx->b_qual, total + 1));
    }
    if (piecewise_done) {
        // nothing left to do before the store append
    } else if (n_over == 0) {
        CompactArgs c{};
        c.base_offsets = d_offsets;
        c.cap_shift = ctx->cap_shift;
        c.cap_const = ctx->cap_const;
        c.n_min = ctx->n_min.as<uint32_t>();
        c.tight_off = ctx->b_off.as<uint64_t>();
        c.in_min = ctx->pad_min.as<uint32_t>();
        c.in_pos = ctx->pad_pos.as<uint32_t>();
        c.in_dir = ctx->pad_dir.as<uint8_t>();
        c.out_min = ctx->b_min.as<uint32_t>();
        c.out_pos = ctx->b_pos.as<uint32_t>();
        c.out_dir = ctx->b_dir.as<uint8_t>();
        c.n_reads = n_reads;
        c.in_qual = want_aux ? ctx->pad_qual.as<uint8_t>() : nullptr;
        c.out_qual = want_aux ? ctx->b_qual.as<uint8_t>() : nullptr;
        launch_compact(c, s);
        CKS(check_launch(ctx, "compact_kernel", 1));
    } else {
        // some read produced more minimizers than its padded slot holds (low-complexity
        // sequence): the exact counts are known now, so sketch again straight into the
        // tight CSR.
        a.exact_off = ctx->b_off.as<uint64_t>();
        a.out_min = ctx->b_min.as<uint32_t>();
        a.out_pos = ctx->b_pos.as<uint32_t>();
        a.out_dir = ctx->b_dir.as<uint8_t>();
        a.dirty_list = ctx->d_dirty.as<uint32_t>();
        a.dirty_count = &ctx->d_small->dirty_n[0];
        a.dirty_cursor = &ctx->d_small->dirty_cur[0];
        CK(cudaMemsetAsync(&ctx->d_small->cursor, 0, sizeof(uint32_t) * (1 + MAX_SUB + 2 * (1 + MAX_SUB)), s));
        CK(cudaMemsetAsync(&ctx->d_small->n_overflow, 0, sizeof(unsigned long long), s));
        const int n_k = launch_sketch(a, ctx->sm_count, s);
        CKS(check_launch(ctx, "sketch_kernel(exact)", n_k + 0));
        if (want_aux) {
            // qualities again on the exact slots (the filter was already applied to n_min; a filtered read has an
            // empty exact slot, and the re-sketch writes nothing into it)
            CKS(ensure(ctx, ctx->pad_raw_a, (total + 1) * 4));
            CKS(ensure(ctx, ctx->pad_raw_b, (total + 1) * 4));
            CKS(run_aux(ctx, d_bases, d_quals, d_offsets, n_reads, n_bases, ctx->b_off.as<uint64_t>(),
                        ctx->b_pos.as<uint32_t>(), ctx->b_qual.as<uint8_t>()));
        }
    }
    if (append) {
        CKS(ensure(ctx, ctx->s_min, (ctx->s_mins + total + 1) * 4, true));
        CKS(ensure(ctx, ctx->s_off, (ctx->s_reads + n_reads + 2) * sizeof(uint64_t), true));
        if (ctx->s_reads == 0) CK(cudaMemsetAsync(ctx->s_off.p, 0, sizeof(uint64_t), s));
        CK(cudaMemcpyAsync(ctx->s_min.as<uint32_t>() + ctx->s_mins, ctx->b_min.p, total * 4,
                           cudaMemcpyDeviceToDevice, s));
        launch_append_offsets(ctx->b_off.as<uint64_t>(), ctx->s_off.as<uint64_t>(), n_reads, ctx->s_reads,
                              ctx->s_mins, s);
        CKS(check_launch(ctx, "append_offsets_kernel", 1));
        ctx->s_reads += n_reads;
        ctx->s_mins += total;
        ctx->store_gen++;
    }
    return MDBG_OK;
}

// ---- NCCL through dlopen ---------------------------------------------------------
struct Id128 { char b[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128 /* ncclUniqueId by value */, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
#define SYM(field, name)                                                     \
    *(void**)(&g_nccl.field) = dlsym(h, name);                               \
    if (!g_nccl.field) { err = std::string("missing symbol ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllGather, "ncclAllGather")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return true;
}

constexpr int NCCL_UINT8 = 1;    // ncclUint8 / ncclChar aliases: ncclInt8=0, ncclUint8=1
constexpr int NCCL_UINT64 = 5;   // ncclUint64

#define NK(call)                                                                                     \
    do {                                                                                             \
        int r_ = (call);                                                                             \
        if (r_ != 0) return fail(ctx, MDBG_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// Owner-partitioned all-to-all, host side.  After a bucket-count pass has left this rank's per-destination record
// counts in m_bucket[0, R), plan() all-gathers the R x R count matrix and derives what this rank sends to and
// receives from everybody; run() moves records of rec_bytes each in one grouped ncclSend / ncclRecv.
struct OwnerExchange {
    std::vector<uint64_t> send_cnt, send_base, recv_cnt, recv_base;
    uint64_t send_total = 0, recv_total = 0;

    mdbg_status plan(mdbg_ctx* ctx) {
        cudaStream_t s = ctx->stream;
        const uint32_t R = (uint32_t)ctx->n_ranks;
        uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
        uint64_t* d_all = d_cnt + 2 * R;
        NK(g_nccl.AllGather(d_cnt, d_all, R, NCCL_UINT64, ctx->nccl_comm, s));
        std::vector<uint64_t> all((size_t)R * R);
        CK(cudaMemcpyAsync(all.data(), d_all, (size_t)R * R * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        send_cnt.assign(R, 0); send_base.assign(R, 0); recv_cnt.assign(R, 0); recv_base.assign(R, 0);
        send_total = recv_total = 0;
        for (uint32_t d = 0; d < R; d++) {
            send_cnt[d] = all[(size_t)ctx->rank * R + d];
            send_base[d] = send_total;
            send_total += send_cnt[d];
            recv_cnt[d] = all[(size_t)d * R + ctx->rank];
            recv_base[d] = recv_total;
            recv_total += recv_cnt[d];
        }
        return MDBG_OK;
    }
    // bucket bases for the scatter pass (m_bucket[R, 2R)) and fresh cursors (m_bucket[0, R))
    mdbg_status upload_bases(mdbg_ctx* ctx) {
        const uint32_t R = (uint32_t)ctx->n_ranks;
        uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
        CK(cudaMemcpyAsync(d_cnt + R, send_base.data(), R * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(d_cnt, 0, R * 8, ctx->stream));
        return MDBG_OK;
    }
    mdbg_status run(mdbg_ctx* ctx, const void* send, void* recv, size_t rec_bytes) {
        const uint32_t R = (uint32_t)ctx->n_ranks;
        NK(g_nccl.GroupStart());
        for (uint32_t d = 0; d < R; d++) {
            if (send_cnt[d])
                NK(g_nccl.Send((const char*)send + send_base[d] * rec_bytes, send_cnt[d] * rec_bytes, NCCL_UINT8, (int)d,
                               ctx->nccl_comm, ctx->stream));
            if (recv_cnt[d])
                NK(g_nccl.Recv((char*)recv + recv_base[d] * rec_bytes, recv_cnt[d] * rec_bytes, NCCL_UINT8, (int)d,
                               ctx->nccl_comm, ctx->stream));
        }
        NK(g_nccl.GroupEnd());
        return MDBG_OK;
    }
};

}  // namespace

// =====================================================================================
extern "C" {

mdbg_status mdbg_ctx_create(int device, const mdbg_params* p, mdbg_ctx** out) {
    mdbg_ctx* ctx = nullptr;   // for the CK macro: errors go to g_create_error
    if (!p || !out) return fail(nullptr, MDBG_ERR_ARG, "mdbg_ctx_create: null argument");
    if (p->minimizer_size < 2 || p->minimizer_size > 16)
        return fail(nullptr, MDBG_ERR_ARG, "minimizer_size %u unsupported (2..16; metaMDBG caps l at 16)",
                    p->minimizer_size);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, MDBG_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(nullptr, MDBG_ERR_ARG, "device %d out of range (%d)", device, n_dev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, MDBG_ERR_CUDA, "device %d is sm_%d%d; libmdbg_b200 is built for sm_100a only", device,
                    prop.major, prop.minor);

    mdbg_ctx* c = new mdbg_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l = p->minimizer_size;
    c->density = p->density;
    c->hpc = p->use_hpc ? 1 : 0;
    c->threshold = threshold_from_density(p->density, &c->select_none);
    // padded output slots: 2^-shift >= 4 * density minimizers per base
    {
        int shift = 0;
        double d4 = 4.0 * (double)p->density;
        while (shift < 8 && 1.0 / (double)(1u << (shift + 1)) >= d4) shift++;
        c->cap_shift = (uint32_t)shift;
        c->cap_const = 32;
    }
    cudaError_t e2 = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e2 != cudaSuccess) { delete c; return fail(nullptr, MDBG_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e2)); }
    c->stream = c->own_stream;
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_gate, cudaEventDisableTiming) != cudaSuccess) {
        mdbg_ctx_destroy(c);
        return fail(nullptr, MDBG_ERR_CUDA, "copy stream creation failed");
    }
    for (int i = 0; i < MAX_SUB; i++)
        if (cudaEventCreateWithFlags(&c->sub_ev[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->tot_ev[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->cmp_ev[i], cudaEventDisableTiming) != cudaSuccess) {
            mdbg_ctx_destroy(c);
            return fail(nullptr, MDBG_ERR_CUDA, "event creation failed");
        }
    if (cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_piece, 2 * MAX_SUB * sizeof(uint64_t)) != cudaSuccess) {
        mdbg_ctx_destroy(c);
        return fail(nullptr, MDBG_ERR_CUDA, "piece pipeline allocation failed");
    }
    if (const char* v = getenv("MDBG_PIECE_PIPELINE")) c->piece_pipeline = atoi(v) != 0;
    if (cudaMalloc((void**)&c->d_small, sizeof(SmallDev)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_small, sizeof(SmallDev)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_scalar, 8 * sizeof(uint64_t)) != cudaSuccess) {
        mdbg_ctx_destroy(c);
        return fail(nullptr, MDBG_ERR_OOM, "context allocation failed");
    }
    cudaMemset(c->d_small, 0, sizeof(SmallDev));
    if (const char* v = getenv("MDBG_TABLE_AUTOGROW")) c->t_autogrow = atoi(v) != 0;
    if (const char* v = getenv("MDBG_SKETCH_VARIANT")) {
        const int want = atoi(v);
        if (want >= 0 && want < SKETCH_VARIANTS) c->sketch_variant = want;
    }
    if (p->blacklist && p->n_blacklist) {
        std::vector<uint32_t> bl(p->blacklist, p->blacklist + p->n_blacklist);
        std::sort(bl.begin(), bl.end());
        bl.erase(std::unique(bl.begin(), bl.end()), bl.end());
        mdbg_status st = ensure(c, c->d_blacklist, bl.size() * 4);
        if (st != MDBG_OK) { g_create_error = c->error; mdbg_ctx_destroy(c); return st; }
        cudaMemcpy(c->d_blacklist.p, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice);
        c->n_blacklist = (uint32_t)bl.size();
    }
    *out = c;
    return MDBG_OK;
}

void mdbg_ctx_destroy(mdbg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* devs[] = {&c->d_blacklist, &c->d_bases, &c->d_offsets, &c->pad_min, &c->pad_pos, &c->pad_dir, &c->n_min,
                      &c->scan_scratch, &c->b_off, &c->b_min, &c->b_pos, &c->b_dir, &c->s_min, &c->s_off, &c->s_rem,
                      &c->p_flags, &c->p_keep, &c->p_cnt, &c->p_newoff, &c->p_newmin, &c->table, &c->foreign_vecs,
                      &c->o_hash, &c->o_abund, &c->o_vecs, &c->d_pack, &c->d_src, &c->d_dirty, &c->pg_counts, &c->pg_flags, &c->pg_keyidx, &c->pg_off, &c->pg_readof, &c->pg_hash, &c->pg_koff, &c->pg_reads, &c->pg_wins, &c->r_table, &c->r_hist, &c->r_out, &c->f_raw, &c->f_cnt, &c->f_off, &c->f_nl, &c->f_start, &c->f_len, &c->f_qstart, &c->d_quals, &c->pad_qual, &c->pad_raw_a, &c->pad_raw_b, &c->b_qual, &c->x_sum_lo, &c->x_sum_hi, &c->x_lmin, &c->x_cplx, &c->x_low, &c->d_err_fixed, &c->d_err_tz, &c->prev_table, &c->prev_stage_h, &c->prev_stage_a, &c->rescue_table, &c->pass_aux, &c->val_a, &c->val_b, &c->prev_src, &c->m_send_vecs, &c->m_send_counts, &c->m_recv_vecs,
                      &c->m_recv_counts, &c->m_bucket, &c->loc_off, &c->edge_table, &c->edge_vals, &c->o_edge_vals,
                      &c->u_slot_node, &c->u_node_slot, &c->u_next, &c->u_pair, &c->u_len, &c->u_size, &c->u_flag, &c->u_cychead,
                      &c->u_seqoff, &c->u_idx, &c->u_cyclist, &c->u_cycpos, &c->u_best, &c->u_jump, &c->u_mins, &c->u_off, &c->u_hash,
                      &c->u_rev, &c->u_circ, &c->u_abund, &c->u_bcnt, &c->u_boff, &c->u_order, &c->u_pos,
                      &c->u_scnt, &c->u_soff, &c->u_ents, &c->u_ecnt, &c->u_eoff, &c->u_etgt};
    for (DevBuf* b : devs) release(*b);
    PinBuf* pins[] = {&c->h_off, &c->h_min, &c->h_pos, &c->h_dir, &c->ho_hash, &c->ho_abund, &c->ho_vecs, &c->hx_sum_lo, &c->hx_sum_hi, &c->hx_lmin, &c->hx_cplx, &c->hx_low, &c->hx_meanq, &c->h_qual, &c->h_pack, &c->h_src, &c->h_asc, &c->ho_edge_vals, &c->hpg_hash, &c->hpg_koff, &c->hpg_reads, &c->hpg_wins,
                      &c->hu_mins, &c->hu_off, &c->hu_hash, &c->hu_circ, &c->hu_order, &c->hu_abund, &c->hu_eoff, &c->hu_etgt};
    for (PinBuf* b : pins) release(*b);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->d_small) cudaFree(c->d_small);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->h_scalar) cudaFreeHost(c->h_scalar);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            if (c->ev[i][j]) cudaEventDestroy(c->ev[i][j]);
    if (c->pool) host_pool_destroy(c->pool);
    for (int i = 0; i < MAX_SUB; i++) {
        if (c->sub_ev[i]) cudaEventDestroy(c->sub_ev[i]);
        if (c->tot_ev[i]) cudaEventDestroy(c->tot_ev[i]);
        if (c->cmp_ev[i]) cudaEventDestroy(c->cmp_ev[i]);
    }
    if (c->h_piece) cudaFreeHost(c->h_piece);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->copy_gate) cudaEventDestroy(c->copy_gate);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* mdbg_last_error(mdbg_ctx* c) { return c ? c->error.c_str() : g_create_error.c_str(); }

mdbg_status mdbg_ctx_set_stream(mdbg_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MDBG_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_synchronize(mdbg_ctx* ctx) {
    if (!ctx) return MDBG_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return MDBG_OK;
}

uint64_t mdbg_ctx_kernel_launches(mdbg_ctx* ctx) { return ctx ? ctx->launches : 0; }

mdbg_status mdbg_ctx_bytes_moved(mdbg_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!ctx) return MDBG_ERR_ARG;
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_allocations(mdbg_ctx* ctx, uint64_t* n_device_allocations, uint64_t* n_table_buffer_trades) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_device_allocations) *n_device_allocations = ctx->n_dev_allocs;
    if (n_table_buffer_trades) *n_table_buffer_trades = ctx->n_table_trades;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_phase_profile(mdbg_ctx* ctx, int on) {
    if (!ctx) return MDBG_ERR_ARG;
    ctx->phase_prof = on != 0;
    for (double& v : ctx->phase_ms) v = 0;
    return MDBG_OK;
}

int mdbg_ctx_phase_times(mdbg_ctx* ctx, double* ms_out, const char** names_out, int max_n) {
    if (!ctx) return 0;
    int n = 0;
    for (; n < PH_N && n < max_n; n++) {
        if (ms_out) ms_out[n] = ctx->phase_ms[n];
        if (names_out) names_out[n] = kPhaseNames[n];
    }
    return n;
}

mdbg_status mdbg_ctx_enable_timing(mdbg_ctx* ctx, int on) {
    if (!ctx) return MDBG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0][0])
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) CK(cudaEventCreate(&ctx->ev[i][j]));
    ctx->timing = on != 0;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_kernel_time_ms(mdbg_ctx* ctx, int which, float* ms) {
    if (!ctx || !ms || which < 0 || which > 1) return MDBG_ERR_ARG;
    if (!ctx->ev_valid[which]) return fail(ctx, MDBG_ERR_STATE, "no timed launch of kernel %d yet", which);
    CK(cudaEventSynchronize(ctx->ev[which][1]));
    CK(cudaEventElapsedTime(ms, ctx->ev[which][0], ctx->ev[which][1]));
    return MDBG_OK;
}

static mdbg_status sketch_host_batch(mdbg_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                                     uint32_t n_reads, uint64_t n_bases, int append, bool want_aux, bool fetch = false);

// Host CSR offsets: start at 0, never decrease, and no read reaches 2^31 bases (positions are u32 and the candidate
// list keeps a flag in bit 31).  A bad array would otherwise turn into out-of-range device addresses.
static mdbg_status check_host_offsets(mdbg_ctx* ctx, const uint64_t* offsets, uint32_t n_reads) {
    if (n_reads == 0) return MDBG_OK;
    if (offsets[0] != 0) return fail(ctx, MDBG_ERR_ARG, "offsets[0] must be 0");
    uint64_t bad = 0;
    for (uint32_t r = 0; r < n_reads; r++) bad |= (offsets[r + 1] - offsets[r]) >> 31;      // wraps if decreasing
    if (bad) return fail(ctx, MDBG_ERR_ARG, "offsets must be non-decreasing with reads shorter than 2^31 bases");
    return MDBG_OK;
}

// ---- sketch -----------------------------------------------------------------------
mdbg_status mdbg_sketch_batch_device(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets,
                                     uint32_t n_reads, uint64_t n_bases, int append_to_store, mdbg_sketch_dev* out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!d_bases || !d_offsets)) return fail(ctx, MDBG_ERR_ARG, "null device buffer");
    if ((uintptr_t)d_bases & 15) return fail(ctx, MDBG_ERR_ARG, "d_bases must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    CKS(sketch_internal(ctx, d_bases, d_offsets, n_reads, n_bases, append_to_store));
    if (out) {
        out->n_reads = n_reads;
        out->n_minimizers = ctx->b_total;
        out->d_min_offsets = ctx->b_off.as<uint64_t>();
        out->d_minimizers = ctx->b_min.as<uint32_t>();
        out->d_positions = ctx->b_pos.as<uint32_t>();
        out->d_directions = ctx->b_dir.as<uint8_t>();
    }
    return MDBG_OK;
}

mdbg_status mdbg_ctx_set_sketch_variant(mdbg_ctx* ctx, int variant) {
    if (!ctx) return MDBG_ERR_ARG;
    if (variant < 0 || variant >= SKETCH_VARIANTS) return fail(ctx, MDBG_ERR_ARG, "sketch variant %d unknown (0..%d)", variant, SKETCH_VARIANTS - 1);
    ctx->sketch_variant = variant;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_get_sketch_variant(mdbg_ctx* ctx, int* variant) {
    if (!ctx || !variant) return MDBG_ERR_ARG;
    *variant = ctx->sketch_variant;
    return MDBG_OK;
}

// Runs every arithmetic variant of the sketch kernel on the caller's own device batch, compares the complete
// results (offsets, minimizers, positions, strands) with variant 0 byte for byte on the device, and keeps the
// fastest variant whose output is identical.  Variant 0 is the reference point and the fallback.
mdbg_status mdbg_ctx_autotune_sketch(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                                     uint64_t n_bases, mdbg_autotune_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    if (n_reads == 0 || !d_bases || !d_offsets) return fail(ctx, MDBG_ERR_ARG, "autotune needs a non-empty device batch");
    if ((uintptr_t)d_bases & 15) return fail(ctx, MDBG_ERR_ARG, "d_bases must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int before = ctx->sketch_variant;
    memset(out, 0, sizeof *out);
    out->n_variants = SKETCH_VARIANTS;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    DevBuf r_off, r_min, r_pos, r_dir;                            // variant 0's result
    uint64_t r_total = 0;
    mdbg_status st = MDBG_OK;
    auto run = [&](int v, float* ms) -> mdbg_status {
        ctx->sketch_variant = v;
        CKS(sketch_internal(ctx, d_bases, d_offsets, n_reads, n_bases, 0));      // warm-up (allocations, caches)
        float best = 0;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0, s));
            CKS(sketch_internal(ctx, d_bases, d_offsets, n_reads, n_bases, 0));
            CK(cudaEventRecord(e1, s));
            CK(cudaEventSynchronize(e1));
            float t = 0;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (rep == 0 || t < best) best = t;
        }
        *ms = best;
        return MDBG_OK;
    };
    // second opinion: the same batch in launches of at most 2048 reads with the shared memory of every SM
    // scrambled in between -- every read then meets a ring it did not write (fresh-CTA conditions, 500 times per
    // million reads instead of once), which a whole-batch launch hardly ever exercises
    const Feeder stress_feeder = [&](SketchArgs& a) -> mdbg_status {
        const uint32_t step = 2048;
        uint32_t it = 0;
        for (uint32_t r0 = 0; r0 < n_reads; r0 += step, it++) {
            launch_smem_scramble(ctx->sm_count, it, &ctx->d_small->full_flag, s);
            CK(cudaMemsetAsync(&ctx->d_small->cursor, 0, sizeof(uint32_t), s));
            CK(cudaMemsetAsync(&ctx->d_small->dirty_n[0], 0, sizeof(uint32_t), s));
            CK(cudaMemsetAsync(&ctx->d_small->dirty_cur[0], 0, sizeof(uint32_t), s));
            a.read_begin = r0;
            a.read_end = std::min<uint64_t>(n_reads, (uint64_t)r0 + step);
            a.cursor = &ctx->d_small->cursor;
            if (it == 0 && a.variant == 2 && sketch_packed_eligible(a)) {
                CKS(device_pack_prepare(ctx, a, n_reads, n_bases));
                CKS(device_pack_range(ctx, a, 0, n_reads, s));
            }
            const int n_k = launch_sketch(a, ctx->sm_count, s);
            CKS(check_launch(ctx, "sketch_kernel(stress)", n_k + 1));
        }
        return MDBG_OK;
    };
    auto same_as_reference = [&](bool* same) -> mdbg_status {
        unsigned long long* n_diff = &ctx->d_small->n_changed;
        CK(cudaMemsetAsync(n_diff, 0, sizeof(unsigned long long), s));
        *same = ctx->b_total == r_total;
        if (!*same) return MDBG_OK;
        launch_count_diff(r_off.p, ctx->b_off.p, ((size_t)n_reads + 1) * 8, n_diff, s);
        launch_count_diff(r_min.p, ctx->b_min.p, r_total * 4, n_diff, s);
        launch_count_diff(r_pos.p, ctx->b_pos.p, r_total * 4, n_diff, s);
        launch_count_diff(r_dir.p, ctx->b_dir.p, r_total, n_diff, s);
        CKS(check_launch(ctx, "count_diff_kernel", 4));
        CK(cudaMemcpyAsync(&ctx->h_scalar[2], n_diff, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        *same = ctx->h_scalar[2] == 0;
        return MDBG_OK;
    };
    auto body = [&]() -> mdbg_status {
        CKS(run(0, &out->ms[0]));
        out->identical[0] = 1;
        r_total = ctx->b_total;
        CKS(ensure(ctx, r_off, ((size_t)n_reads + 1) * 8));
        CKS(ensure(ctx, r_min, (r_total + 1) * 4));
        CKS(ensure(ctx, r_pos, (r_total + 1) * 4));
        CKS(ensure(ctx, r_dir, r_total + 1));
        CK(cudaMemcpyAsync(r_off.p, ctx->b_off.p, ((size_t)n_reads + 1) * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(r_min.p, ctx->b_min.p, r_total * 4, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(r_pos.p, ctx->b_pos.p, r_total * 4, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(r_dir.p, ctx->b_dir.p, r_total, cudaMemcpyDeviceToDevice, s));
        int best_v = 0;
        {   // the reference itself must survive the stress run, or the batch is not a usable yardstick
            bool same = false;
            CKS(sketch_internal(ctx, d_bases, d_offsets, n_reads, n_bases, 0, false, nullptr, &stress_feeder));
            CKS(same_as_reference(&same));
            if (!same) return fail(ctx, MDBG_ERR_STATE, "autotune: variant 0 is not reproducible under the fresh-CTA stress run");
        }
        for (int v = 1; v < SKETCH_VARIANTS; v++) {
            CKS(run(v, &out->ms[v]));
            bool same = false;
            CKS(same_as_reference(&same));
            if (same) {
                CKS(sketch_internal(ctx, d_bases, d_offsets, n_reads, n_bases, 0, false, nullptr, &stress_feeder));
                CKS(same_as_reference(&same));
            }
            out->identical[v] = same ? 1 : 0;
            if (same && out->ms[v] < 0.99f * out->ms[best_v]) best_v = v;
        }
        out->chosen = best_v;
        return MDBG_OK;
    };
    st = body();
    ctx->sketch_variant = (st == MDBG_OK) ? out->chosen : before;
    out->n_reads = n_reads;
    out->n_minimizers = r_total;
    cudaStreamSynchronize(s);
    release(r_off); release(r_min); release(r_pos); release(r_dir);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return st;
}

mdbg_status mdbg_sketch_batch_device_packed(mdbg_ctx* ctx, const uint32_t* d_packed, const uint64_t* d_word_offsets,
                                            const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases,
                                            int append_to_store, mdbg_sketch_dev* out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!d_packed || !d_word_offsets || !d_offsets)) return fail(ctx, MDBG_ERR_ARG, "null device buffer");
    if ((uintptr_t)d_packed & 15) return fail(ctx, MDBG_ERR_ARG, "d_packed must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = d_word_offsets;            // bit 63 clear everywhere: every read is packed
        a.packed = d_packed;
        a.bases = nullptr;
        a.bases_end = nullptr;
        if (ctx->timing) CK(cudaEventRecord(ctx->ev[0][0], ctx->stream));
        const int n_k = launch_sketch(a, ctx->sm_count, ctx->stream);
        if (ctx->timing) { CK(cudaEventRecord(ctx->ev[0][1], ctx->stream)); ctx->ev_valid[0] = true; }
        return check_launch(ctx, "sketch_kernel", n_k + 0);
    };
    CKS(sketch_internal(ctx, nullptr, d_offsets, n_reads, n_bases, append_to_store, false, nullptr, &feeder));
    if (out) {
        out->n_reads = n_reads;
        out->n_minimizers = ctx->b_total;
        out->d_min_offsets = ctx->b_off.as<uint64_t>();
        out->d_minimizers = ctx->b_min.as<uint32_t>();
        out->d_positions = ctx->b_pos.as<uint32_t>();
        out->d_directions = ctx->b_dir.as<uint8_t>();
    }
    return MDBG_OK;
}

uint64_t mdbg_pack_device_words(uint64_t n_bases, uint64_t n_reads) { return pack_words_capacity(n_bases, n_reads) + 16; }

mdbg_status mdbg_pack_device(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                             uint64_t n_bases, uint32_t* d_packed_out, uint64_t* d_read_src_out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads == 0) return MDBG_OK;
    if (!d_bases || !d_offsets || !d_packed_out || !d_read_src_out) return fail(ctx, MDBG_ERR_ARG, "null device buffer");
    if (((uintptr_t)d_bases & 15) || ((uintptr_t)d_packed_out & 15)) return fail(ctx, MDBG_ERR_ARG, "d_bases and d_packed_out must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    PackArgsAscii pa{};
    pa.bases = d_bases; pa.bases_end = d_bases + n_bases; pa.offsets = d_offsets;
    pa.read_begin = 0; pa.read_end = n_reads;
    pa.packed = d_packed_out; pa.read_src = d_read_src_out;
    launch_pack_ascii(pa, ctx->sm_count, ctx->stream);
    return check_launch(ctx, "pack_ascii_kernel", 1);
}

mdbg_status mdbg_sketch_batch_device_packed2(mdbg_ctx* ctx, const uint32_t* d_packed, const uint64_t* d_read_src,
                                             const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                                             uint64_t n_bases, int append_to_store, mdbg_sketch_dev* out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!d_packed || !d_read_src || !d_offsets)) return fail(ctx, MDBG_ERR_ARG, "null device buffer");
    if (((uintptr_t)d_packed & 15) || ((uintptr_t)d_bases & 15)) return fail(ctx, MDBG_ERR_ARG, "d_packed and d_bases must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = d_read_src;
        a.packed = d_packed;
        a.bases = d_bases;
        a.bases_end = d_bases ? d_bases + n_bases : nullptr;
        if (ctx->timing) CK(cudaEventRecord(ctx->ev[0][0], ctx->stream));
        const int n_k = launch_sketch(a, ctx->sm_count, ctx->stream);
        if (ctx->timing) { CK(cudaEventRecord(ctx->ev[0][1], ctx->stream)); ctx->ev_valid[0] = true; }
        return check_launch(ctx, "sketch_kernel", n_k);
    };
    CKS(sketch_internal(ctx, nullptr, d_offsets, n_reads, n_bases, append_to_store, false, nullptr, &feeder));
    if (out) {
        out->n_reads = n_reads;
        out->n_minimizers = ctx->b_total;
        out->d_min_offsets = ctx->b_off.as<uint64_t>();
        out->d_minimizers = ctx->b_min.as<uint32_t>();
        out->d_positions = ctx->b_pos.as<uint32_t>();
        out->d_directions = ctx->b_dir.as<uint8_t>();
    }
    return MDBG_OK;
}

mdbg_status mdbg_sketch_fetch(mdbg_ctx* ctx, mdbg_sketch_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->b_reads;
    const uint64_t t = ctx->b_total;
    if (ctx->host_csr_valid) {                  // copied piece by piece behind the sketch: wait for the last copies
        CK(cudaStreamSynchronize(ctx->d2h_stream));
        out->n_reads = n;
        out->n_minimizers = t;
        out->min_offsets = ctx->h_off.as<uint64_t>();
        out->minimizers = ctx->h_min.as<uint32_t>();
        out->positions = ctx->h_pos.as<uint32_t>();
        out->directions = ctx->h_dir.as<uint8_t>();
        return MDBG_OK;
    }
    CKS(ensure_pin(ctx, ctx->h_off, ((size_t)n + 1) * 8));
    CKS(ensure_pin(ctx, ctx->h_min, (t + 1) * 4));
    CKS(ensure_pin(ctx, ctx->h_pos, (t + 1) * 4));
    CKS(ensure_pin(ctx, ctx->h_dir, t + 1));
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->h_off.p, ctx->b_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (t) {
        CK(cudaMemcpyAsync(ctx->h_min.p, ctx->b_min.p, t * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_pos.p, ctx->b_pos.p, t * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_dir.p, ctx->b_dir.p, t, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += ((uint64_t)n + 1) * 8 + t * 9;
    out->n_reads = n;
    out->n_minimizers = t;
    out->min_offsets = ctx->h_off.as<uint64_t>();
    out->minimizers = ctx->h_min.as<uint32_t>();
    out->positions = ctx->h_pos.as<uint32_t>();
    out->directions = ctx->h_dir.as<uint8_t>();
    return MDBG_OK;
}

mdbg_status mdbg_sketch_batch(mdbg_ctx* ctx, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads,
                              int append_to_store, mdbg_sketch_out* out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!bases || !offsets)) return fail(ctx, MDBG_ERR_ARG, "null host buffer");
    CK(cudaSetDevice(ctx->device));
    CKS(check_host_offsets(ctx, offsets, n_reads));
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    CKS(ensure(ctx, ctx->d_bases, n_bases + 64));
    CKS(ensure(ctx, ctx->d_offsets, ((size_t)n_reads + 1) * 8));
    CKS(sketch_host_batch(ctx, bases, nullptr, offsets, n_reads, n_bases, append_to_store, false, out != nullptr));
    if (out) return mdbg_sketch_fetch(ctx, out);
    return MDBG_OK;
}

// Host batch -> device in up to MAX_SUB pieces on the copy stream, so that the sketch of piece i overlaps the
// transfer of piece i+1 (pieces end on read boundaries).  Without qualities the pieces travel 2-bit packed
// (pack_host.cpp, AVX2 + worker threads): 4x fewer PCIe bytes; reads holding a byte outside "ACGT" keep their
// ASCII form so that the result is unchanged.
static void split_pieces(const uint64_t* offsets, uint32_t n_reads, uint64_t n_bases, std::vector<SubRange>& subs) {
    subs.clear();
    // pieces of at least 128 MB (MDBG_PIECE_BYTES overrides the floor: tests use it to cut small batches into many
    // pieces), at most MAX_SUB of them
    uint64_t floor_bytes = uint64_t(128) << 20;
    if (const char* e = getenv("MDBG_PIECE_BYTES")) { const long long v = atoll(e); if (v > 0) floor_bytes = (uint64_t)v; }
    const uint64_t piece = std::max<uint64_t>(floor_bytes, (n_bases + (MAX_SUB - 8) - 1) / (MAX_SUB - 8));
    // Ramp: nothing overlaps the packing of the first piece, and nothing overlaps the copy / sketch / scan / compaction
    // of the last one, so a batch of several pieces starts with a quarter and a half piece and ends by halving what
    // is left (down to an eighth): at most 6 extra pieces.
    const bool ramp = n_bases >= 4 * piece;
    uint32_t r0 = 0;
    while (r0 < n_reads) {
        uint64_t size = piece;
        if (ramp) {
            const uint64_t rem = n_bases - offsets[r0];
            if (subs.size() == 0) size = piece / 4;
            else if (subs.size() == 1) size = piece / 2;
            else if (rem < 2 * piece) size = rem / 2 >= piece / 8 ? rem / 2 : rem;
        }
        const uint64_t target = offsets[r0] + std::max<uint64_t>(size, 1);
        uint32_t r1 = (uint32_t)(std::upper_bound(offsets + r0 + 1, offsets + n_reads + 1, target) - offsets);
        if (r1 <= r0 + 1) r1 = r0 + 1; else r1 -= 1;
        if (r1 > n_reads || subs.size() + 1 == (size_t)MAX_SUB) r1 = n_reads;
        subs.push_back(SubRange{r0, r1});
        r0 = r1;
    }
}

static mdbg_status sketch_host_batch(mdbg_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                                     uint32_t n_reads, uint64_t n_bases, int append, bool want_aux, bool fetch) {
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    ctx->host_csr_valid = false;
    CKS(ensure(ctx, ctx->d_offsets, ((size_t)n_reads + 1) * 8));
    if (n_reads == 0)
        return sketch_internal(ctx, nullptr, ctx->d_offsets.as<uint64_t>(), 0, 0, append, want_aux, nullptr);
    std::vector<SubRange> subs;
    split_pieces(offsets, n_reads, n_bases, subs);
    PiecePipeline pipeline{ctx, ctx->d_offsets.as<uint64_t>(), &subs, n_reads, fetch};
    bool want_pipe = ctx->piece_pipeline;
    if (const char* e = getenv("MDBG_PIECE_PIPELINE")) want_pipe = atoi(e) != 0;            // A/B runs, tests
    PiecePipeline* pipe = (want_pipe && !quals && !want_aux) ? &pipeline : nullptr;
    const auto done = [&](mdbg_status st) {
        if (st == MDBG_OK && pipe && !pipe->overflow && pipe->fetch) ctx->host_csr_valid = true;
        return st;
    };
    // the copy stream must not overwrite buffers the compute stream is still reading
    CK(cudaEventRecord(ctx->copy_gate, s));
    CK(cudaStreamWaitEvent(cs, ctx->copy_gate, 0));
    CK(cudaMemcpyAsync(ctx->d_offsets.p, offsets, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, cs));
    ctx->h2d_bytes += ((uint64_t)n_reads + 1) * 8;

    // auto: packing pays when the host can pack faster than PCIe moves ASCII (~55 GB/s), i.e. with >= 12 usable CPUs
    bool want_pack = ctx->host_packing >= 1;
    if (ctx->host_packing < 0) {
        const char* env = getenv("MDBG_HOST_THREADS");
        want_pack = ((env && atoi(env) > 0) ? atoi(env) : host_default_threads()) >= 12;
        // ... and only while the packer keeps ahead of what PCIe would move as ASCII (busy or NUMA-remote hosts)
        if (want_pack && ctx->auto_pack_pause > 0) { ctx->auto_pack_pause--; want_pack = false; }
    }
    uint64_t pack_floor = uint64_t(1) << 20;            // below 1 MB the transfer is not worth a pass over the bases
    if (const char* e = getenv("MDBG_PACK_MIN_BYTES")) { const long long v = atoll(e); if (v >= 0) pack_floor = (uint64_t)v; }   // tests
    const bool packed = want_pack && !quals && !want_aux && n_bases >= pack_floor;
    ctx->last_packed = packed ? 1 : 0;
    ctx->last_pieces = subs.size();
    if (!pipe) { ctx->last_pipelined = ctx->last_grows = 0; ctx->last_overflow_fallback = 0; }
    if (!packed) {
        CKS(ensure(ctx, ctx->d_bases, n_bases + 64));
        if (quals) CKS(ensure(ctx, ctx->d_quals, n_bases + 64));
        const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
            // the ASCII bytes are packed on the device, piece by piece, in front of the packed kernel (variant 2)
            const bool dev_pack = a.variant == 2 && sketch_packed_eligible(a);
            if (dev_pack) CKS(device_pack_prepare(ctx, a, n_reads, n_bases));
            for (size_t i = 0; i < subs.size(); i++) {
                const uint64_t lo = offsets[subs[i].r0], hi = offsets[subs[i].r1];
                if (hi > lo) {
                    CK(cudaMemcpyAsync(ctx->d_bases.as<uint8_t>() + lo, bases + lo, hi - lo, cudaMemcpyHostToDevice, cs));
                    if (quals)
                        CK(cudaMemcpyAsync(ctx->d_quals.as<uint8_t>() + lo, quals + lo, hi - lo, cudaMemcpyHostToDevice, cs));
                    ctx->h2d_bytes += (hi - lo) * (quals ? 2 : 1);
                }
                CK(cudaEventRecord(ctx->sub_ev[i], cs));
                CK(cudaStreamWaitEvent(s, ctx->sub_ev[i], 0));
                piece_args(ctx, a, i, subs[i].r0, subs[i].r1);
                if (dev_pack) CKS(device_pack_range(ctx, a, subs[i].r0, subs[i].r1, s));
                const int n_k = launch_sketch(a, ctx->sm_count, s);
                CKS(check_launch(ctx, "sketch_kernel", n_k + 0));
                if (pipe) CKS(pipe->launched(i));
            }
            return MDBG_OK;
        };
        return done(sketch_internal(ctx, ctx->d_bases.as<uint8_t>(), ctx->d_offsets.as<uint64_t>(), n_reads, n_bases, append,
                                    want_aux, quals ? ctx->d_quals.as<uint8_t>() : nullptr, &feeder, pipe));
    }

    // ---- packed transfer ----------------------------------------------------------------------------------------
    if (!ctx->pool) ctx->pool = host_pool_create(0);
    std::vector<uint64_t> pk_off((size_t)n_reads + 1);
    pk_off[0] = 0;
    // every read starts on a 16-byte boundary of the packed buffer: the packed kernel fetches it by bulk copy
    for (uint32_t r = 0; r < n_reads; r++) pk_off[r + 1] = pk_off[r] + (((offsets[r + 1] - offsets[r] + 63) >> 6) << 2);
    const uint64_t n_words = pk_off[n_reads];
    const uint64_t asc_cap = n_bases + 16ull * n_reads + 64;                 // worst case: every read kept as ASCII
    CKS(ensure_pin(ctx, ctx->h_pack, (n_words + 1) * 4));
    CKS(ensure_pin(ctx, ctx->h_src, (size_t)n_reads * 8));
    CKS(ensure(ctx, ctx->d_pack, (n_words + 1) * 4 + 64));
    CKS(ensure(ctx, ctx->d_src, (size_t)n_reads * 8));
    // Hybrid transfer: when the caller's buffer is pinned, a piece can also travel as plain ASCII by DMA alone
    // (no CPU work).  Each piece picks its mode by looking at the copy stream: if the previous pieces have already
    // landed, PCIe is idle and the piece is sent as it is; otherwise the host threads use the waiting time to pack
    // it.  Both resources (copy engine, CPU packer) stay busy and the split adapts to the machine.
    bool can_direct = false;
    if (ctx->host_packing == 2) {      // opt-in: on the round-1 test box the DMA reads slowed the packer more than they saved
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, bases) == cudaSuccess && attr.type == cudaMemoryTypeHost) can_direct = true;
        else cudaGetLastError();                                         // pageable memory: not an error
    }
    // device ASCII area: [0, n_bases) mirrors the batch (direct pieces), the spill of packed pieces follows it
    const uint64_t spill_base = can_direct ? ((n_bases + 15) & ~uint64_t(15)) : 0;
    if (can_direct) CKS(ensure(ctx, ctx->d_bases, spill_base + (uint64_t(16) << 20)));
    std::atomic<uint64_t> asc_cursor{0};
    uint64_t asc_sent = 0;
    uint64_t n_direct_pieces = 0;
    double pack_seconds = 0;
    uint64_t pack_bytes = 0;
    // worst-case ASCII spill of a piece must fit the pinned staging before the workers start on it
    const auto reserve_spill = [&](uint32_t r0, uint32_t r1, uint64_t asc_used) -> mdbg_status {
        const uint64_t piece_bytes = offsets[r1] - offsets[r0] + 16ull * (r1 - r0);
        if (ctx->h_asc.cap < asc_used + piece_bytes) {
            PinBuf bigger;
            CKS(ensure_pin(ctx, bigger, asc_used + piece_bytes));
            if (ctx->h_asc.p && asc_used) {
                CK(cudaStreamSynchronize(cs));                           // earlier spill copies still read the old buffer
                memcpy(bigger.p, ctx->h_asc.p, asc_used);
            }
            release(ctx->h_asc);
            ctx->h_asc = bigger;
        }
        return MDBG_OK;
    };
    const auto start_pack = [&](uint32_t r0, uint32_t r1) {
        return host_pack_start(ctx->pool, bases, offsets, r0, r1, pk_off.data(), ctx->h_pack.as<uint32_t>(),
                               ctx->h_src.as<uint64_t>(), ctx->h_asc.as<uint8_t>(), &asc_cursor);
    };
    // copies of a packed piece [r0, r1) whose spill area ends at asc_end, then its sketch launch
    const auto enqueue_packed_piece = [&](SketchArgs& a, size_t i, uint32_t r0, uint32_t r1, uint64_t asc_end) -> mdbg_status {
        if (spill_base) {                                                // spilled reads live behind the mirror area
            uint64_t* src = ctx->h_src.as<uint64_t>();
            for (uint32_t r = r0; r < r1; r++)
                if (src[r] & SRC_ASCII) src[r] = SRC_ASCII | (spill_base + (src[r] & ~SRC_ASCII));
        }
        const uint64_t w0 = pk_off[r0], w1 = pk_off[r1];
        if (w1 > w0)
            CK(cudaMemcpyAsync(ctx->d_pack.as<uint32_t>() + w0, ctx->h_pack.as<uint32_t>() + w0, (w1 - w0) * 4,
                               cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(ctx->d_src.as<uint64_t>() + r0, ctx->h_src.as<uint64_t>() + r0, (size_t)(r1 - r0) * 8,
                           cudaMemcpyHostToDevice, cs));
        ctx->h2d_bytes += (w1 - w0) * 4 + (uint64_t)(r1 - r0) * 8;
        if (asc_end > asc_sent) {
            if (ctx->d_bases.cap < spill_base + asc_end + 64) {          // grow, keeping what is already there
                CK(cudaStreamSynchronize(cs));
                CK(cudaStreamSynchronize(s));
                CKS(ensure(ctx, ctx->d_bases,
                           std::min<uint64_t>(spill_base + asc_cap, spill_base + 2 * asc_end + (uint64_t(64) << 20)), true));
            }
            CK(cudaMemcpyAsync(ctx->d_bases.as<uint8_t>() + spill_base + asc_sent, ctx->h_asc.as<uint8_t>() + asc_sent,
                               asc_end - asc_sent, cudaMemcpyHostToDevice, cs));
            ctx->h2d_bytes += asc_end - asc_sent;
            asc_sent = asc_end;
        }
        CK(cudaEventRecord(ctx->sub_ev[i], cs));
        CK(cudaStreamWaitEvent(s, ctx->sub_ev[i], 0));
        a.bases = ctx->d_bases.as<uint8_t>();
        a.bases_end = ctx->d_bases.p ? ctx->d_bases.as<uint8_t>() + ctx->d_bases.cap - 32 : nullptr;
        piece_args(ctx, a, i, r0, r1);
        const int n_k = launch_sketch(a, ctx->sm_count, s);
        CKS(check_launch(ctx, "sketch_kernel", n_k + 0));
        if (pipe) CKS(pipe->launched(i));
        return MDBG_OK;
    };
    // Plain packed transfer: piece i+1 is handed to the workers BEFORE piece i's copies, launches and piece-pipeline
    // bookkeeping are enqueued (~20 driver calls, ~0.1 ms: 8 % of a piece's packing time on 16 threads), so the
    // packer -- the bottleneck of the host path -- never waits for the enqueuing thread.
    PackJob* job_in_flight = nullptr;
    struct JobGuard {                                                    // an early error return must not leave workers
        HostPool*& pool; PackJob*& job;                                  // running on the caller's buffers
        ~JobGuard() { if (job) { host_pack_wait(pool, job); job = nullptr; } }
    } job_guard{ctx->pool, job_in_flight};
    const Feeder overlapped_feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = ctx->d_src.as<uint64_t>();
        a.packed = ctx->d_pack.as<uint32_t>();
        CKS(reserve_spill(subs[0].r0, subs[0].r1, 0));
        auto t_pack = std::chrono::steady_clock::now();
        job_in_flight = start_pack(subs[0].r0, subs[0].r1);
        for (size_t i = 0; i < subs.size(); i++) {
            const uint32_t r0 = subs[i].r0, r1 = subs[i].r1;
            host_pack_wait(ctx->pool, job_in_flight);
            job_in_flight = nullptr;
            pack_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pack).count();
            pack_bytes += offsets[r1] - offsets[r0];
            const uint64_t asc_end = asc_cursor.load();                  // spill of pieces 0 .. i, complete
            if (i + 1 < subs.size()) {
                CKS(reserve_spill(subs[i + 1].r0, subs[i + 1].r1, asc_end));
                t_pack = std::chrono::steady_clock::now();
                job_in_flight = start_pack(subs[i + 1].r0, subs[i + 1].r1);
            }
            CKS(enqueue_packed_piece(a, i, r0, r1, asc_end));
        }
        return MDBG_OK;
    };
    const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = ctx->d_src.as<uint64_t>();
        a.packed = ctx->d_pack.as<uint32_t>();
        for (size_t i = 0; i < subs.size(); i++) {
            const uint32_t r0 = subs[i].r0, r1 = subs[i].r1;
            bool direct = false;
            if (can_direct) direct = (i == 0) || cudaEventQuery(ctx->sub_ev[i - 1]) == cudaSuccess;
            if (direct) {
                uint64_t* src = ctx->h_src.as<uint64_t>();
                for (uint32_t r = r0; r < r1; r++) src[r] = SRC_ASCII | offsets[r];
                const uint64_t lo = offsets[r0], hi = offsets[r1];
                if (hi > lo)
                    CK(cudaMemcpyAsync(ctx->d_bases.as<uint8_t>() + lo, bases + lo, hi - lo, cudaMemcpyHostToDevice, cs));
                CK(cudaMemcpyAsync(ctx->d_src.as<uint64_t>() + r0, src + r0, (size_t)(r1 - r0) * 8,
                                   cudaMemcpyHostToDevice, cs));
                ctx->h2d_bytes += (hi - lo) + (uint64_t)(r1 - r0) * 8;
                n_direct_pieces++;
                CK(cudaEventRecord(ctx->sub_ev[i], cs));
                CK(cudaStreamWaitEvent(s, ctx->sub_ev[i], 0));
                a.bases = ctx->d_bases.as<uint8_t>();
                a.bases_end = ctx->d_bases.p ? ctx->d_bases.as<uint8_t>() + ctx->d_bases.cap - 32 : nullptr;
                piece_args(ctx, a, i, r0, r1);
                const int n_k = launch_sketch(a, ctx->sm_count, s);
                CKS(check_launch(ctx, "sketch_kernel", n_k + 0));
                if (pipe) CKS(pipe->launched(i));
            } else {
                CKS(reserve_spill(r0, r1, asc_cursor.load()));
                const auto t_pack = std::chrono::steady_clock::now();
                host_pack_reads(ctx->pool, bases, offsets, r0, r1, pk_off.data(), ctx->h_pack.as<uint32_t>(),
                                ctx->h_src.as<uint64_t>(), ctx->h_asc.as<uint8_t>(), &asc_cursor);
                pack_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pack).count();
                pack_bytes += offsets[r1] - offsets[r0];
                CKS(enqueue_packed_piece(a, i, r0, r1, asc_cursor.load()));
            }
        }
        return MDBG_OK;
    };
    ctx->last_direct_pieces = 0;
    const mdbg_status st = done(sketch_internal(ctx, ctx->d_bases.as<uint8_t>(), ctx->d_offsets.as<uint64_t>(), n_reads, n_bases,
                                                append, false, nullptr, can_direct ? &feeder : &overlapped_feeder, pipe));
    ctx->last_direct_pieces = n_direct_pieces;
    ctx->last_pieces = subs.size();
    if (pack_seconds > 0) {
        ctx->last_pack_gbs = (double)pack_bytes / pack_seconds / 1e9;
        // slower than the ~48 GB/s the pipelined ASCII path sustains over PCIe: stop packing for a while
        if (ctx->host_packing < 0 && pack_bytes >= (uint64_t(256) << 20) && ctx->last_pack_gbs < 45.0) ctx->auto_pack_pause = 16;
    }
    return st;
}

// Raw FASTQ / FASTA text -> sketch, record split on the device (ingest.cu; row (f)3).
mdbg_status mdbg_sketch_fastx(mdbg_ctx* ctx, const uint8_t* text, uint64_t n_bytes, int is_final, int append_to_store,
                              mdbg_sketch_out* out, mdbg_fastx_info* info) {
    if (!ctx || !info) return MDBG_ERR_ARG;
    memset(info, 0, sizeof *info);
    if (n_bytes && !text) return fail(ctx, MDBG_ERR_ARG, "null text buffer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    ctx->host_csr_valid = false;
    uint32_t lines = 0;
    if (n_bytes) {
        lines = text[0] == '@' ? 4u : text[0] == '>' ? 2u : 0u;
        if (!lines) return fail(ctx, MDBG_ERR_ARG, "the text does not start with '@' (FASTQ) or '>' (FASTA)");
    }
    // a last line without '\n' at the end of the input still ends a record
    const bool virtual_nl = is_final && n_bytes && text[n_bytes - 1] != '\n';
    const uint64_t n_dev = n_bytes + (virtual_nl ? 1 : 0);
    CKS(ensure(ctx, ctx->f_raw, n_dev + 64));
    if (n_bytes) CK(cudaMemcpyAsync(ctx->f_raw.p, text, n_bytes, cudaMemcpyHostToDevice, s));
    if (virtual_nl) CK(cudaMemsetAsync(ctx->f_raw.as<uint8_t>() + n_bytes, '\n', 1, s));
    ctx->h2d_bytes += n_bytes;
    const uint8_t* d_text = ctx->f_raw.as<uint8_t>();
    uint64_t n_nl = 0;
    if (n_dev) {
        const uint64_t tiles = newline_tiles(n_dev);
        if (tiles > 0xFFFFFFF0ull) return fail(ctx, MDBG_ERR_ARG, "text block too large");
        CKS(ensure(ctx, ctx->f_cnt, tiles * 4));
        CKS(ensure(ctx, ctx->f_off, (tiles + 1) * 8));
        CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)tiles) * sizeof(uint64_t)));
        launch_newline_count(d_text, n_dev, ctx->f_cnt.as<uint32_t>(), s);
        launch_scan_u32_to_u64(ctx->f_cnt.as<uint32_t>(), ctx->f_off.as<uint64_t>(), (uint32_t)tiles, ctx->scan_scratch.as<uint64_t>(), s);
        CKS(check_launch(ctx, "newline_count_kernel + scan", 4));
        CK(cudaMemcpyAsync(&ctx->h_scalar[0], ctx->f_off.as<uint64_t>() + tiles, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        n_nl = ctx->h_scalar[0];
        CKS(ensure(ctx, ctx->f_nl, (n_nl + 1) * 8));
        launch_newline_write(d_text, n_dev, ctx->f_off.as<uint64_t>(), ctx->f_nl.as<uint64_t>(), s);
        CKS(check_launch(ctx, "newline_write_kernel", 1));
    }
    const uint64_t n_rec64 = lines ? n_nl / lines : 0;
    if (n_rec64 > 0xFFFFFFF0ull) return fail(ctx, MDBG_ERR_ARG, "too many records in one block");
    const uint32_t n_rec = (uint32_t)n_rec64;
    CKS(ensure(ctx, ctx->d_offsets, ((size_t)n_rec + 1) * 8));
    uint64_t n_bases = 0;
    if (n_rec) {
        CKS(ensure(ctx, ctx->f_start, (size_t)n_rec * 8));
        CKS(ensure(ctx, ctx->f_len, (size_t)n_rec * 4));
        CKS(ensure(ctx, ctx->f_qstart, (size_t)n_rec * 8));
        CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems(n_rec) * sizeof(uint64_t)));
        CK(cudaMemsetAsync(&ctx->d_small->n_flagged, 0, sizeof(unsigned long long), s));
        FastxArgs fa{};
        fa.text = d_text; fa.nl = ctx->f_nl.as<uint64_t>(); fa.n_records = n_rec; fa.lines = lines;
        fa.seq_start = ctx->f_start.as<uint64_t>(); fa.seq_len = ctx->f_len.as<uint32_t>();
        fa.qual_start = ctx->f_qstart.as<uint64_t>(); fa.n_bad = &ctx->d_small->n_flagged;
        launch_fastx_records(fa, s);
        launch_scan_u32_to_u64(ctx->f_len.as<uint32_t>(), ctx->d_offsets.as<uint64_t>(), n_rec, ctx->scan_scratch.as<uint64_t>(), s);
        CKS(check_launch(ctx, "fastx_records_kernel + scan", 4));
        CK(cudaMemcpyAsync(&ctx->h_scalar[0], &ctx->d_small->n_flagged, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_scalar[1], ctx->f_nl.as<uint64_t>() + ((uint64_t)n_rec * lines - 1), 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_scalar[2], ctx->d_offsets.as<uint64_t>() + n_rec, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (ctx->h_scalar[0])
            return fail(ctx, MDBG_ERR_ARG, "%llu of %u records are not well-formed 4-line FASTQ / 2-line FASTA (multi-line "
                        "sequences need the host parser)", (unsigned long long)ctx->h_scalar[0], n_rec);
        info->consumed_bytes = std::min<uint64_t>(ctx->h_scalar[1] + 1, n_bytes);
        n_bases = ctx->h_scalar[2];
    } else {
        CK(cudaMemsetAsync(ctx->d_offsets.p, 0, 8, s));
    }
    info->n_records = n_rec;
    info->n_bases = n_bases;
    info->format = lines == 4 ? 1 : lines == 2 ? 2 : 0;
    // sequence bytes -> 2-bit layout straight from the text; reads with a byte outside "ACGT" are sketched from the
    // text itself (SRC_ASCII | position) by the byte-ring kernel
    CKS(ensure(ctx, ctx->d_pack, pack_words_capacity(n_bases, n_rec) * 4 + 64));
    CKS(ensure(ctx, ctx->d_src, (size_t)n_rec * 8 + 8));
    if (n_rec) {
        PackArgsAscii pa{};
        pa.bases = d_text; pa.bases_end = d_text + n_dev; pa.offsets = ctx->d_offsets.as<uint64_t>();
        pa.read_begin = 0; pa.read_end = n_rec;
        pa.packed = ctx->d_pack.as<uint32_t>(); pa.read_src = ctx->d_src.as<uint64_t>();
        pa.src_start = ctx->f_start.as<uint64_t>();
        launch_pack_ascii(pa, ctx->sm_count, s);
        CKS(check_launch(ctx, "pack_ascii_kernel", 1));
    }
    const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = ctx->d_src.as<uint64_t>();
        a.packed = ctx->d_pack.as<uint32_t>();
        a.bases = d_text;
        a.bases_end = d_text + n_dev;
        const int n_k = launch_sketch(a, ctx->sm_count, s);
        return check_launch(ctx, "sketch_kernel", n_k);
    };
    CKS(sketch_internal(ctx, nullptr, ctx->d_offsets.as<uint64_t>(), n_rec, n_bases, append_to_store, false, nullptr, &feeder));
    if (out) return mdbg_sketch_fetch(ctx, out);
    return MDBG_OK;
}

int mdbg_host_pack_read(const uint8_t* bases, uint64_t len, uint32_t* words_out) {
    return host_pack_one(bases, len, words_out) ? 1 : 0;
}

// Host batch handed over 2-bit packed: pieces of the word array cross PCIe on the copy stream while the packed kernel
// works on the piece before; scan / compaction / D2H of the CSR piece by piece behind it (PiecePipeline).
mdbg_status mdbg_sketch_batch_packed(mdbg_ctx* ctx, const uint32_t* packed, uint64_t n_words, const uint64_t* read_src,
                                     const uint8_t* ascii, uint64_t n_ascii_bytes, const uint64_t* offsets,
                                     uint32_t n_reads, int append_to_store, mdbg_sketch_out* out) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!read_src || !offsets || (n_words && !packed))) return fail(ctx, MDBG_ERR_ARG, "null host buffer");
    CK(cudaSetDevice(ctx->device));
    CKS(check_host_offsets(ctx, offsets, n_reads));
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    // every read must lie inside its buffer, packed reads in non-decreasing word order
    uint64_t last_w = 0;
    for (uint32_t r = 0; r < n_reads; r++) {
        const uint64_t len = offsets[r + 1] - offsets[r], src = read_src[r];
        if (src & SRC_ASCII) {
            if ((src & ~SRC_ASCII) + len > n_ascii_bytes || !ascii) return fail(ctx, MDBG_ERR_ARG, "read %u lies outside the ASCII spill buffer", r);
        } else {
            if (src < last_w || src + ((len + 15) >> 4) > n_words) return fail(ctx, MDBG_ERR_ARG, "read %u: packed reads must follow each other without overlap inside the packed buffer", r);
            last_w = src + ((len + 15) >> 4);
        }
    }
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    ctx->host_csr_valid = false;
    CKS(ensure(ctx, ctx->d_offsets, ((size_t)n_reads + 1) * 8));
    if (n_reads == 0) {
        CKS(sketch_internal(ctx, nullptr, ctx->d_offsets.as<uint64_t>(), 0, 0, append_to_store));
        return out ? mdbg_sketch_fetch(ctx, out) : MDBG_OK;
    }
    std::vector<SubRange> subs;
    split_pieces(offsets, n_reads, n_bases, subs);
    const bool fetch = out != nullptr;
    PiecePipeline pipeline{ctx, ctx->d_offsets.as<uint64_t>(), &subs, n_reads, fetch};
    bool want_pipe = ctx->piece_pipeline;
    if (const char* e = getenv("MDBG_PIECE_PIPELINE")) want_pipe = atoi(e) != 0;
    PiecePipeline* pipe = want_pipe ? &pipeline : nullptr;
    CKS(ensure(ctx, ctx->d_pack, (n_words + 4) * 4 + 64));
    CKS(ensure(ctx, ctx->d_src, (size_t)n_reads * 8));
    CKS(ensure(ctx, ctx->d_bases, n_ascii_bytes + 64));
    CK(cudaEventRecord(ctx->copy_gate, s));                          // the copy stream must not overwrite buffers still being read
    CK(cudaStreamWaitEvent(cs, ctx->copy_gate, 0));
    CK(cudaMemcpyAsync(ctx->d_offsets.p, offsets, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(ctx->d_src.p, read_src, (size_t)n_reads * 8, cudaMemcpyHostToDevice, cs));
    if (n_ascii_bytes) CK(cudaMemcpyAsync(ctx->d_bases.p, ascii, n_ascii_bytes, cudaMemcpyHostToDevice, cs));
    ctx->h2d_bytes += ((uint64_t)n_reads + 1) * 8 + (uint64_t)n_reads * 8 + n_ascii_bytes;
    // first word of each piece = word offset of its first packed read
    std::vector<uint64_t> piece_w(subs.size() + 1, n_words);
    {
        size_t i = subs.size();
        uint64_t next = n_words;
        while (i-- > 0) {
            for (uint32_t r = subs[i].r1; r-- > subs[i].r0;)
                if (!(read_src[r] & SRC_ASCII)) next = read_src[r];
            piece_w[i] = next;
        }
        piece_w[0] = std::min<uint64_t>(piece_w[0], n_words);
    }
    ctx->last_packed = 0; ctx->last_pieces = subs.size(); ctx->last_direct_pieces = 0;
    if (!pipe) { ctx->last_pipelined = ctx->last_grows = 0; ctx->last_overflow_fallback = 0; }
    const Feeder feeder = [&](SketchArgs& a) -> mdbg_status {
        a.read_src = ctx->d_src.as<uint64_t>();
        a.packed = ctx->d_pack.as<uint32_t>();
        a.bases = ctx->d_bases.as<uint8_t>();
        a.bases_end = ctx->d_bases.as<uint8_t>() + n_ascii_bytes;
        for (size_t i = 0; i < subs.size(); i++) {
            const uint64_t w0 = (i == 0) ? 0 : piece_w[i], w1 = piece_w[i + 1];
            if (w1 > w0) {
                CK(cudaMemcpyAsync(ctx->d_pack.as<uint32_t>() + w0, packed + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, cs));
                ctx->h2d_bytes += (w1 - w0) * 4;
            }
            CK(cudaEventRecord(ctx->sub_ev[i], cs));
            CK(cudaStreamWaitEvent(s, ctx->sub_ev[i], 0));
            piece_args(ctx, a, i, subs[i].r0, subs[i].r1);
            const int n_k = launch_sketch(a, ctx->sm_count, s);
            CKS(check_launch(ctx, "sketch_kernel", n_k));
            if (pipe) CKS(pipe->launched(i));
        }
        return MDBG_OK;
    };
    CKS(sketch_internal(ctx, ctx->d_bases.as<uint8_t>(), ctx->d_offsets.as<uint64_t>(), n_reads, n_bases, append_to_store, false,
                        nullptr, &feeder, pipe));
    if (pipe && !pipe->overflow && pipe->fetch) ctx->host_csr_valid = true;
    if (out) return mdbg_sketch_fetch(ctx, out);
    return MDBG_OK;
}

mdbg_status mdbg_ctx_last_batch_info(mdbg_ctx* ctx, mdbg_batch_info* info) {
    if (!ctx || !info) return MDBG_ERR_ARG;
    info->n_pieces = ctx->last_pieces;
    info->n_pieces_pipelined = ctx->last_pipelined;
    info->n_buffer_growths = ctx->last_grows;
    info->n_direct_pieces = ctx->last_direct_pieces;
    info->overflow_fallback = ctx->last_overflow_fallback;
    info->packed = ctx->last_packed;
    info->pack_gb_per_s = ctx->last_packed ? ctx->last_pack_gbs : 0.0;
    info->pack_isa = host_pack_isa();
    info->host_threads = ctx->pool ? host_pool_size(ctx->pool) : 0;
    return MDBG_OK;
}

mdbg_status mdbg_ctx_set_host_packing(mdbg_ctx* ctx, int on) {
    if (!ctx) return MDBG_ERR_ARG;
    ctx->host_packing = on < 0 ? -1 : (on > 2 ? 1 : on);
    return MDBG_OK;
}

mdbg_status mdbg_ctx_set_read_filters(mdbg_ctx* ctx, int filter_low_complexity) {
    if (!ctx) return MDBG_ERR_ARG;
    ctx->filter_low_complexity = filter_low_complexity != 0;
    return MDBG_OK;
}

mdbg_status mdbg_sketch_batch_q(mdbg_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                                uint32_t n_reads, int append_to_store, mdbg_sketch_out* out, mdbg_aux_out* aux) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads && (!bases || !offsets)) return fail(ctx, MDBG_ERR_ARG, "null host buffer");
    CK(cudaSetDevice(ctx->device));
    CKS(check_host_offsets(ctx, offsets, n_reads));
    const uint64_t n_bases = n_reads ? offsets[n_reads] : 0;
    // '#' is EncoderRLE's internal sentinel (Commons.hpp:4172): a base string holding it makes the reference record
    // shifted rlePositions, which the quality windows of the side outputs do not reproduce.  No FASTA/FASTQ parser
    // produces it, so such a batch is refused here instead of returning windows that differ (the plain sketch entry
    // points accept it and stay bit-exact).
    if (ctx->hpc && n_bases && memchr(bases, '#', n_bases))
        return fail(ctx, MDBG_ERR_ARG, "a read contains '#': not supported by the side-output entry point with HPC on");
    CKS(ensure(ctx, ctx->d_bases, n_bases + 64));
    CKS(ensure(ctx, ctx->d_offsets, ((size_t)n_reads + 1) * 8));
    cudaStream_t s = ctx->stream;
    CKS(sketch_host_batch(ctx, bases, quals, offsets, n_reads, n_bases, append_to_store, true));
    if (out) CKS(mdbg_sketch_fetch(ctx, out));
    if (aux) {
        const uint64_t t = ctx->b_total;
        CKS(ensure_pin(ctx, ctx->hx_sum_lo, (size_t)n_reads * 8 + 8));
        CKS(ensure_pin(ctx, ctx->hx_sum_hi, (size_t)n_reads * 8 + 8));
        CKS(ensure_pin(ctx, ctx->hx_lmin, n_reads + 1));
        CKS(ensure_pin(ctx, ctx->hx_cplx, (size_t)n_reads * 8 + 8));
        CKS(ensure_pin(ctx, ctx->hx_low, n_reads + 1));
        CKS(ensure_pin(ctx, ctx->hx_meanq, (size_t)n_reads * 4 + 4));
        CKS(ensure_pin(ctx, ctx->h_qual, t + 1));
        if (n_reads) {
            CK(cudaMemcpyAsync(ctx->hx_sum_lo.p, ctx->x_sum_lo.p, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(ctx->hx_sum_hi.p, ctx->x_sum_hi.p, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(ctx->hx_cplx.p, ctx->x_cplx.p, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(ctx->hx_low.p, ctx->x_low.p, n_reads, cudaMemcpyDeviceToHost, s));
        }
        if (t) CK(cudaMemcpyAsync(ctx->h_qual.p, ctx->b_qual.p, t, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        // ReadSelection.hpp:870-879 on the exact error sum, with the reference's own operand types
        float* mq = ctx->hx_meanq.as<float>();
        for (uint32_t r = 0; r < n_reads; r++) {
            const uint64_t len = offsets[r + 1] - offsets[r];
            if (!quals) {                                  // empty _qual: errorSum / 0 = NaN, carried through log10f
                volatile long double z = 0, n0 = 0;
                float e = z / n0;
                mq[r] = -10.0f * log10f(e);
                continue;
            }
            long double errorSum = ldexpl((long double)ctx->hx_sum_hi.as<uint64_t>()[r], 64 - ERR_SHIFT) +
                                   ldexpl((long double)ctx->hx_sum_lo.as<uint64_t>()[r], -ERR_SHIFT);
            float meanReadError = errorSum / len;
            mq[r] = -10.0f * log10f(meanReadError);
        }
        aux->n_reads = n_reads;
        aux->mean_quality = mq;
        aux->complexity = ctx->hx_cplx.as<double>();
        aux->low_complexity = ctx->hx_low.as<uint8_t>();
        aux->qualities = ctx->h_qual.as<uint8_t>();
    }
    return MDBG_OK;
}

// ---- store ------------------------------------------------------------------------
// The slots of a count table reference positions of the store (Slot::ref): anything that rewrites the store ends
// the current table -- finalize / merge / edges first, then clear or purge (mdbg_store_append keeps positions).
static void store_rewritten(mdbg_ctx* ctx) {
    ctx->t_active = false;
    ctx->t_ranges.clear();
    ctx->store_gen++;
}

mdbg_status mdbg_store_clear(mdbg_ctx* ctx) {
    if (!ctx) return MDBG_ERR_ARG;
    store_rewritten(ctx);
    ctx->s_reads = 0;
    ctx->s_mins = 0;
    ctx->store_gen++;
    return MDBG_OK;
}

mdbg_status mdbg_store_append(mdbg_ctx* ctx, const uint32_t* minimizers, const uint64_t* min_offsets,
                              uint32_t n_reads) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads == 0) return MDBG_OK;
    if (!min_offsets) return fail(ctx, MDBG_ERR_ARG, "null min_offsets");
    for (uint32_t r = 0; r < n_reads; r++)
        if (min_offsets[r + 1] < min_offsets[r]) return fail(ctx, MDBG_ERR_ARG, "min_offsets must be non-decreasing");
    if (min_offsets[n_reads] > min_offsets[0] && !minimizers) return fail(ctx, MDBG_ERR_ARG, "null minimizers");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t total = min_offsets[n_reads] - min_offsets[0];
    CKS(ensure(ctx, ctx->s_min, (ctx->s_mins + total + 1) * 4, true));
    CKS(ensure(ctx, ctx->s_off, (ctx->s_reads + n_reads + 2) * 8, true));
    CKS(ensure(ctx, ctx->b_off, ((size_t)n_reads + 1) * 8));
    if (ctx->s_reads == 0) CK(cudaMemsetAsync(ctx->s_off.p, 0, 8, s));
    if (total)
        CK(cudaMemcpyAsync(ctx->s_min.as<uint32_t>() + ctx->s_mins, minimizers + min_offsets[0], total * 4,
                           cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->b_off.p, min_offsets, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
    launch_append_offsets(ctx->b_off.as<uint64_t>(), ctx->s_off.as<uint64_t>(), n_reads, ctx->s_reads,
                          ctx->s_mins - min_offsets[0], s);
    CKS(check_launch(ctx, "append_offsets_kernel", 1));
    CK(cudaStreamSynchronize(s));
    ctx->b_reads = 0;
    ctx->b_total = 0;
    ctx->s_reads += n_reads;
    ctx->s_mins += total;
    ctx->store_gen++;
    return MDBG_OK;
}

mdbg_status mdbg_store_size(mdbg_ctx* ctx, uint64_t* n_reads, uint64_t* n_minimizers) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads) *n_reads = ctx->s_reads;
    if (n_minimizers) *n_minimizers = ctx->s_mins;
    return MDBG_OK;
}

mdbg_status mdbg_store_fetch(mdbg_ctx* ctx, uint64_t* min_offsets, uint32_t* minimizers) {
    if (!ctx) return MDBG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    if (ctx->s_reads == 0) {
        if (min_offsets) min_offsets[0] = 0;
        return MDBG_OK;
    }
    if (min_offsets)
        CK(cudaMemcpyAsync(min_offsets, ctx->s_off.p, (ctx->s_reads + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (minimizers && ctx->s_mins)
        CK(cudaMemcpyAsync(minimizers, ctx->s_min.p, ctx->s_mins * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MDBG_OK;
}

mdbg_status mdbg_purge_palindromes(mdbg_ctx* ctx, uint32_t first_k, uint32_t last_k, uint64_t* n_reads_changed) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads_changed) *n_reads_changed = 0;
    if (ctx->s_reads == 0 || last_k <= first_k) return MDBG_OK;
    if (first_k < 2) return fail(ctx, MDBG_ERR_ARG, "first_k must be >= 2");
    if (ctx->s_reads > 0xFFFFFFFFull) return fail(ctx, MDBG_ERR_ARG, "store too large for purge");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CKS(ensure(ctx, ctx->p_flags, ctx->s_reads));
    CK(cudaMemsetAsync(&ctx->d_small->n_flagged, 0, 2 * sizeof(unsigned long long), s));   // n_flagged, n_changed
    launch_purge_flag(ctx->s_min.as<uint32_t>(), ctx->s_off.as<uint64_t>(), ctx->s_reads, first_k, last_k,
                      ctx->p_flags.as<uint8_t>(), &ctx->d_small->n_flagged, s);
    CKS(check_launch(ctx, "purge_flag_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], &ctx->d_small->n_flagged, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_scalar[0] == 0) return MDBG_OK;

    CKS(ensure(ctx, ctx->p_keep, ctx->s_mins + 1));
    CKS(ensure(ctx, ctx->p_cnt, ctx->s_reads * 4));
    CK(cudaMemsetAsync(ctx->p_keep.p, 1, ctx->s_mins + 1, s));
    launch_purge_exact(ctx->s_min.as<uint32_t>(), ctx->s_off.as<uint64_t>(), ctx->s_reads, ctx->p_flags.as<uint8_t>(),
                       first_k, last_k, ctx->p_keep.as<uint8_t>(), ctx->p_cnt.as<uint32_t>(),
                       &ctx->d_small->n_changed, s);
    CKS(check_launch(ctx, "purge_exact_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_scalar[1], &ctx->d_small->n_changed, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t changed = ctx->h_scalar[1];
    if (n_reads_changed) *n_reads_changed = changed;
    if (changed == 0) return MDBG_OK;

    CKS(ensure(ctx, ctx->p_newoff, (ctx->s_reads + 2) * 8));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)ctx->s_reads) * 8));
    launch_scan_u32_to_u64(ctx->p_cnt.as<uint32_t>(), ctx->p_newoff.as<uint64_t>(), (uint32_t)ctx->s_reads,
                           ctx->scan_scratch.as<uint64_t>(), s);
    CKS(check_launch(ctx, "scan", 3));
    CK(cudaMemcpyAsync(&ctx->h_scalar[2], ctx->p_newoff.as<uint64_t>() + ctx->s_reads, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t new_total = ctx->h_scalar[2];
    CKS(ensure(ctx, ctx->p_newmin, (new_total + 1) * 4));
    CKS(ensure(ctx, ctx->s_rem, ctx->s_mins + 1));                 // (sized for the old store: the new one is smaller)
    launch_purge_compact(ctx->s_min.as<uint32_t>(), ctx->s_off.as<uint64_t>(), ctx->p_newoff.as<uint64_t>(),
                         ctx->p_keep.as<uint8_t>(), ctx->s_reads, ctx->p_newmin.as<uint32_t>(), s, ctx->s_rem.as<uint8_t>());
    CKS(check_launch(ctx, "purge_compact_kernel", 1));
    std::swap(ctx->s_min, ctx->p_newmin);                          // (everything that follows is ordered on the same stream)
    std::swap(ctx->s_off, ctx->p_newoff);
    ctx->s_mins = new_total;
    store_rewritten(ctx);
    ctx->rem_gen = ctx->store_gen;                                 // rem[] of the new store came out of the compaction
    return MDBG_OK;
}

mdbg_status mdbg_store_apply_density(mdbg_ctx* ctx, float density, uint64_t* n_reads_changed) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads_changed) *n_reads_changed = 0;
    if (ctx->s_reads == 0) return MDBG_OK;
    if (ctx->s_reads > 0xFFFFFFFFull) return fail(ctx, MDBG_ERR_ARG, "store too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint32_t none = 0;
    const uint64_t thr = threshold_from_density(density, &none);
    CKS(ensure(ctx, ctx->p_keep, ctx->s_mins + 1));
    CKS(ensure(ctx, ctx->p_cnt, ctx->s_reads * 4));
    CK(cudaMemsetAsync(&ctx->d_small->n_changed, 0, sizeof(unsigned long long), s));
    launch_density_filter(ctx->s_min.as<uint32_t>(), ctx->s_off.as<uint64_t>(), ctx->s_reads, thr, none,
                          ctx->p_keep.as<uint8_t>(), ctx->p_cnt.as<uint32_t>(), &ctx->d_small->n_changed, s);
    CKS(check_launch(ctx, "density_filter_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_scalar[1], &ctx->d_small->n_changed, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t changed = ctx->h_scalar[1];
    if (n_reads_changed) *n_reads_changed = changed;
    if (changed == 0) return MDBG_OK;
    CKS(ensure(ctx, ctx->p_newoff, (ctx->s_reads + 2) * 8));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)ctx->s_reads) * 8));
    launch_scan_u32_to_u64(ctx->p_cnt.as<uint32_t>(), ctx->p_newoff.as<uint64_t>(), (uint32_t)ctx->s_reads,
                           ctx->scan_scratch.as<uint64_t>(), s);
    CKS(check_launch(ctx, "scan", 3));
    CK(cudaMemcpyAsync(&ctx->h_scalar[2], ctx->p_newoff.as<uint64_t>() + ctx->s_reads, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t new_total = ctx->h_scalar[2];
    CKS(ensure(ctx, ctx->p_newmin, (new_total + 1) * 4));
    CKS(ensure(ctx, ctx->s_rem, ctx->s_mins + 1));                 // (sized for the old store: the new one is smaller)
    launch_purge_compact(ctx->s_min.as<uint32_t>(), ctx->s_off.as<uint64_t>(), ctx->p_newoff.as<uint64_t>(),
                         ctx->p_keep.as<uint8_t>(), ctx->s_reads, ctx->p_newmin.as<uint32_t>(), s, ctx->s_rem.as<uint8_t>());
    CKS(check_launch(ctx, "purge_compact_kernel", 1));
    std::swap(ctx->s_min, ctx->p_newmin);                          // (everything that follows is ordered on the same stream)
    std::swap(ctx->s_off, ctx->p_newoff);
    ctx->s_mins = new_total;
    store_rewritten(ctx);
    ctx->rem_gen = ctx->store_gen;                                 // rem[] of the new store came out of the compaction
    return MDBG_OK;
}

// ---- repetitive minimizers (K1b) -------------------------------------------------------
static mdbg_status check_full(mdbg_ctx* ctx, const char* what);

mdbg_status mdbg_ctx_set_blacklist(mdbg_ctx* ctx, const uint32_t* values, uint64_t n) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n && !values) return fail(ctx, MDBG_ERR_ARG, "null blacklist");
    if (n > 0xFFFFFFFFull) return fail(ctx, MDBG_ERR_ARG, "blacklist too large");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> bl(values, values + n);
    std::sort(bl.begin(), bl.end());
    bl.erase(std::unique(bl.begin(), bl.end()), bl.end());
    if (!bl.empty()) {
        CKS(ensure(ctx, ctx->d_blacklist, bl.size() * 4));
        CK(cudaMemcpy(ctx->d_blacklist.p, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice));
    }
    ctx->n_blacklist = (uint32_t)bl.size();
    return MDBG_OK;
}

mdbg_status mdbg_store_repetitive_minimizers(mdbg_ctx* ctx, float fraction, mdbg_repeats_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    memset(out, 0, sizeof *out);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t n = ctx->s_mins;
    if (n == 0) return MDBG_OK;
    const uint64_t cap = pow2ceil(std::max<uint64_t>(1024, 2 * n));
    constexpr uint32_t BINS = 65536;
    CKS(ensure(ctx, ctx->r_table, cap * 8));
    CKS(ensure(ctx, ctx->r_hist, (size_t)BINS * 8));
    CK(cudaMemsetAsync(ctx->r_table.p, 0xFF, cap * 8, s));
    CK(cudaMemsetAsync(ctx->r_hist.p, 0, (size_t)BINS * 8, s));
    CK(cudaMemsetAsync(&ctx->d_small->n_changed, 0, 8, s));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, 4, s));
    launch_mincount_insert(ctx->s_min.as<uint32_t>(), n, ctx->r_table.as<unsigned long long>(), cap - 1,
                           &ctx->d_small->n_changed, &ctx->d_small->full_flag, s);
    launch_mincount_hist(ctx->r_table.as<unsigned long long>(), cap, ctx->r_hist.as<unsigned long long>(), BINS, s);
    CKS(check_launch(ctx, "mincount_insert_kernel + mincount_hist_kernel", 2));
    std::vector<unsigned long long> hist(BINS);
    CK(cudaMemcpyAsync(hist.data(), ctx->r_hist.p, (size_t)BINS * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], &ctx->d_small->n_changed, 8, cudaMemcpyDeviceToHost, s));
    CKS(check_full(ctx, "mdbg_store_repetitive_minimizers"));
    const uint64_t n_distinct = ctx->h_scalar[0];
    // ReadSelection.hpp:524-525: int nb = fraction(float) * vec.size(); nb = max(nb, 1)
    int nb = (int)(fraction * (float)n_distinct);
    if (nb < 1) nb = 1;
    uint64_t want = std::min<uint64_t>((uint64_t)nb, n_distinct);
    // the smallest count still selected: largest T with #(count >= T) >= want
    uint64_t ge = 0;
    uint32_t T = 1;
    for (uint32_t c = BINS - 1; c >= 1; c--) {
        ge += hist[c];
        if (ge >= want) { T = c; break; }
    }
    CKS(ensure(ctx, ctx->r_out, (ge + 1) * 8));
    CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
    launch_mincount_emit(ctx->r_table.as<unsigned long long>(), cap, T, ctx->r_out.as<unsigned long long>(),
                         &ctx->d_small->emit_cursor, ge, s);
    CKS(check_launch(ctx, "mincount_emit_kernel", 1));
    std::vector<unsigned long long> cand(ge);
    if (ge) CK(cudaMemcpyAsync(cand.data(), ctx->r_out.p, ge * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += ge * 8 + (uint64_t)BINS * 8;
    // most frequent first; upstream's order among equal counts is whatever std::sort makes of an unordered_map's
    // iteration order, here it is the smaller value first
    std::sort(cand.begin(), cand.end(), [](unsigned long long a, unsigned long long b) {
        const uint32_t ca = (uint32_t)a, cb = (uint32_t)b;
        return ca != cb ? ca > cb : (a >> 32) < (b >> 32);
    });
    ctx->r_sel_min.resize(want);
    ctx->r_sel_cnt.resize(want);
    uint64_t ties_all = 0, ties_sel = 0;
    const uint32_t c_last = want ? (uint32_t)cand[want - 1] : 0;
    for (uint64_t i = 0; i < ge; i++) ties_all += (uint32_t)cand[i] == c_last;
    for (uint64_t i = 0; i < want; i++) {
        ctx->r_sel_min[i] = (uint32_t)(cand[i] >> 32);
        ctx->r_sel_cnt[i] = (uint32_t)cand[i];
        ties_sel += (uint32_t)cand[i] == c_last;
    }
    out->n_distinct = n_distinct;
    out->n_selected = want;
    out->minimizers = ctx->r_sel_min.data();
    out->counts = ctx->r_sel_cnt.data();
    out->min_count_selected = c_last;
    out->n_with_min_count = ties_all;
    out->n_with_min_count_selected = ties_sel;
    return MDBG_OK;
}

// ---- count table --------------------------------------------------------------------
// capacity for `expect` distinct keys at load factor <= 0.6, and the claim count at which a pass gives up (0.8)
// (any multiple of 1024 slots: the probes scale the hash to the capacity, table.cuh; MDBG_TABLE_POW2=1 restores the
// power-of-two sizes of the earlier rounds for comparisons)
static uint64_t table_capacity_for(uint64_t expect) {
    static const bool pow2 = [] { const char* e = getenv("MDBG_TABLE_POW2"); return e && atoi(e) != 0; }();
    if (expect < 512) expect = 512;
    const uint64_t want = expect + (expect * 2) / 3 + 1;
    return pow2 ? pow2ceil(want) : (want + 1023) & ~uint64_t(1023);
}

// The count table and the previous-k table swap buffers with every k of a loop (mdbg_prev_from_current), so each of
// the two buffers has to hold the largest table of the loop sooner or later: a buffer of the pair is never allocated
// smaller than its sibling, and a loop stops allocating after its first pass instead of after two sweeps (a
// cudaMalloc / cudaFree pair of this size costs 5 - 10 ms on the B200 boxes, more than two next-k passes).
static mdbg_status ensure_table_buf(mdbg_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return MDBG_OK;
    // (with several ranks a third buffer, prev_src, rotates with the two: mdbg_count_add_store_next_k's lazy replication)
    size_t sib_cap = 0;
    for (const DevBuf* o : {&ctx->table, &ctx->prev_table, &ctx->prev_src})
        if (o != &b && o->cap > sib_cap) sib_cap = o->cap;
    const size_t with_slack = bytes + bytes / 8;
    // ensure() adds its own slack of bytes / 8 + 256: ask for what makes the result at least the largest sibling's size
    const size_t sib_payload = sib_cap > 256 ? (sib_cap - 256) - (sib_cap - 256) / 9 + 16 : 0;
    return ensure(ctx, b, with_slack > sib_payload ? bytes : sib_payload);
}

// claim-counter shards of the warp-form passes: one per 16 k slots, at most 64 (a shard's share of the load limit must
// be large against the 32 claims a warp can add at once and against the imbalance between shards)
static uint32_t pass_shards(uint64_t cap) {
    uint32_t n = 1;
    while (n < 64 && (uint64_t)n * 2 * 16384 <= cap) n *= 2;
    return n;
}

// the claim accounting of a fresh table: the counter the host reads and the shards / flag copies of the warp-form passes
// (a merged table starts from zero as well: its limit is its capacity, so only a probe sequence that runs out can still
// raise the flag for passes added after the merge)
static mdbg_status reset_claims(mdbg_ctx* ctx) {
    CK(cudaMemsetAsync(&ctx->d_small->t_claims, 0, sizeof(unsigned long long), ctx->stream));
    CKS(ensure(ctx, ctx->pass_aux, sizeof(PassAux)));
    CK(cudaMemsetAsync(ctx->pass_aux.p, 0, sizeof(PassAux), ctx->stream));
    return MDBG_OK;
}

static mdbg_status table_reset(mdbg_ctx* ctx, uint64_t cap) {
    PhaseClock clk(ctx);
    // mdbg_prev_from_current swaps the two table buffers, so a loop over k leaves the large first-pass buffer on either
    // side.  When the current buffer is too small and the previous-k table sits in one that is large enough, its live
    // slots move into the small one (a device copy of a few hundred MB) and the buffers trade places: no cudaMalloc /
    // cudaFree in the loop (on the B200 boxes a reallocation of that size cost 5 - 70 ms, measured).
    if (cap * sizeof(Slot) > ctx->table.cap && ctx->table.p && ctx->prev_table.p && ctx->prev_table.cap >= cap * sizeof(Slot) &&
        ctx->prev_capacity * sizeof(Slot) <= ctx->table.cap) {
        if (ctx->prev_capacity)
            CK(cudaMemcpyAsync(ctx->table.p, ctx->prev_table.p, ctx->prev_capacity * sizeof(Slot), cudaMemcpyDeviceToDevice, ctx->stream));
        std::swap(ctx->table, ctx->prev_table);
        ctx->n_table_trades++;
    }
    CKS(ensure_table_buf(ctx, ctx->table, cap * sizeof(Slot)));
    CK(cudaMemsetAsync(ctx->table.p, 0, cap * sizeof(Slot), ctx->stream));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), ctx->stream));
    CKS(reset_claims(ctx));
    ctx->t_capacity = cap;
    ctx->t_claim_limit = cap - cap / 5;
    clk.lap(PH_TABLE_RESET);
    return MDBG_OK;
}

mdbg_status mdbg_count_begin(mdbg_ctx* ctx, uint32_t k, uint64_t expected_distinct) {
    if (!ctx) return MDBG_ERR_ARG;
    if (k < 2 || k > 255) return fail(ctx, MDBG_ERR_ARG, "k=%u unsupported (2..255)", k);
    CK(cudaSetDevice(ctx->device));
    // expected_distinct = 0: distinct k-min-mers per stored minimizer as the last first-pass table saw them (+ 15 %),
    // a quarter of the windows before anything is known; a table that turns out too small is rebuilt (count_pass)
    uint64_t expect = expected_distinct;
    if (!expect) {
        const double ratio = ctx->distinct_ratio > 0 ? ctx->distinct_ratio * 1.15 : 0.25;
        expect = (uint64_t)((double)ctx->s_mins * ratio) + 1024;
        if (expect > ctx->s_mins + 1024) expect = ctx->s_mins + 1024;
    }
    CKS(table_reset(ctx, table_capacity_for(expect)));
    ctx->t_k = k;
    ctx->t_active = true;
    ctx->t_value_mode = false;
    ctx->t_merged = false;
    ctx->t_ranges.clear();
    ctx->t_rebuildable = true;
    ctx->t_whole = false;
    ctx->t_no_vecs = false;
    ctx->foreign_n = 0;
    return MDBG_OK;
}

// minimizers-left-in-the-read array of the store, rebuilt only when the store changed
static mdbg_status ensure_rem(mdbg_ctx* ctx) {
    if (ctx->rem_gen == ctx->store_gen && ctx->s_rem.p) return MDBG_OK;
    CKS(ensure(ctx, ctx->s_rem, ctx->s_mins + 1));
    launch_fill_rem(ctx->s_off.as<uint64_t>(), 0, ctx->s_reads, ctx->s_rem.as<uint8_t>(), ctx->stream);
    CKS(check_launch(ctx, "fill_rem_kernel", ctx->s_reads ? 1 : 0));
    ctx->rem_gen = ctx->store_gen;
    return MDBG_OK;
}

// flat minimizer range of a read range (no device round trip for the usual "whole store")
static mdbg_status flat_range(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi, uint64_t* g_lo, uint64_t* g_hi) {
    if (read_lo == 0 && read_hi == ctx->s_reads) { *g_lo = 0; *g_hi = ctx->s_mins; return MDBG_OK; }
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], ctx->s_off.as<uint64_t>() + read_lo, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[1], ctx->s_off.as<uint64_t>() + read_hi, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *g_lo = ctx->h_scalar[0]; *g_hi = ctx->h_scalar[1];
    return MDBG_OK;
}

// one pass over a store range into the current table: occurrence counts (first pass) or next-k values
static mdbg_status launch_pass(mdbg_ctx* ctx, uint64_t g_lo, uint64_t g_hi, bool next_k, bool timed,
                               const uint32_t* val_in = nullptr, uint32_t* val_out = nullptr) {
    cudaStream_t s = ctx->stream;
    if (timed && ctx->timing) CK(cudaEventRecord(ctx->ev[1][0], s));
    if (!next_k) {
        InsertArgs a{};
        a.mins = ctx->s_min.as<uint32_t>();
        a.rem = ctx->s_rem.as<uint8_t>();
        a.g_lo = g_lo; a.g_hi = g_hi;
        a.k = ctx->t_k;
        a.table = ctx->table.as<Slot>();
        a.mask = ctx->t_capacity - 1;
        a.full_flag = &ctx->d_small->full_flag;
        a.claims = &ctx->d_small->t_claims;
        a.claim_limit = ctx->t_claim_limit;
        a.aux = ctx->pass_aux.as<PassAux>();
        a.aux_shards = pass_shards(ctx->t_capacity);
        launch_insert(a, s);
    } else {
        NextKArgs a{};
        a.mins = ctx->s_min.as<uint32_t>();
        a.rem = ctx->s_rem.as<uint8_t>();
        a.g_lo = g_lo; a.g_hi = g_hi;
        a.k = ctx->t_k;
        a.prev = ctx->prev_table.as<Slot>();
        a.prev_mask = ctx->prev_capacity - 1;
        a.prev_min_count = ctx->prev_min_count;
        a.table = ctx->table.as<Slot>();
        a.mask = ctx->t_capacity - 1;
        a.full_flag = &ctx->d_small->full_flag;
        a.claims = &ctx->d_small->t_claims;
        a.claim_limit = ctx->t_claim_limit;
        a.val_in = val_in;
        a.val_out = val_out;
        a.aux = ctx->pass_aux.as<PassAux>();
        a.aux_shards = pass_shards(ctx->t_capacity);
        launch_next_k(a, s);
    }
    if (timed && ctx->timing) { CK(cudaEventRecord(ctx->ev[1][1], s)); ctx->ev_valid[1] = true; }
    return check_launch(ctx, next_k ? "next-k pass kernel (+ pass_fold_kernel)" : "insert pass kernel (+ pass_fold_kernel)",
                        g_hi > g_lo ? pass_kernels_per_launch(ctx->pass_aux.p != nullptr) : 0);
}

// Runs the pass for [read_lo, read_hi); when the table proves too small (probe limit or load limit hit -- the
// kernels abandon the pass early) it is rebuilt 4x larger from every range inserted so far and the pass is redone.
static mdbg_status replicate_into_prev(mdbg_ctx* ctx, const Slot* src, uint64_t src_cap, uint32_t thr);

static mdbg_status count_pass(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi, bool next_k) {
    cudaStream_t s = ctx->stream;
    CKS(ensure_rem(ctx));
    uint64_t g_lo, g_hi;
    CKS(flat_range(ctx, read_lo, read_hi, &g_lo, &g_hi));
    // a pass over the WHOLE store into a fresh table leaves per-position values for the next k (val_b), and can itself
    // run without lookups when the previous-k table is nothing but the previous such pass's table
    const bool whole = read_lo == 0 && read_hi == ctx->s_reads && ctx->t_ranges.empty() && ctx->foreign_n == 0;
    const uint32_t* val_in = nullptr;
    uint32_t* val_out = nullptr;
    if (next_k) {
        const bool stream = whole && ctx->val_valid && ctx->val_gen == ctx->store_gen && ctx->val_k + 1 == ctx->t_k &&
                            ctx->prev_pure && ctx->prev_k + 1 == ctx->t_k && !getenv("MDBG_NO_STREAM_NEXT_K");
        if (stream) val_in = ctx->val_a.as<uint32_t>();
        else if (ctx->n_ranks > 1 && !ctx->prev_replicated) {
            // the pass needs lookups in the COMPLETE previous-k table: collective exchange of the owned entries
            std::swap(ctx->prev_table, ctx->prev_src);
            CKS(replicate_into_prev(ctx, ctx->prev_src.as<Slot>(), ctx->prev_capacity, ctx->prev_min_count));
        }
        if (whole) {
            CKS(ensure(ctx, ctx->val_b, (ctx->s_mins + 2) * 4));
            val_out = ctx->val_b.as<uint32_t>();
        }
    }
    PhaseClock clk(ctx);
    CKS(launch_pass(ctx, g_lo, g_hi, next_k, true, val_in, val_out));
    for (int attempt = 0;; attempt++) {
        CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_small->t_claims, &ctx->d_small->t_claims, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (!ctx->h_small->full_flag) break;
        const uint64_t worst = table_capacity_for(ctx->s_mins + 1024);
        if (!ctx->t_rebuildable || !ctx->t_autogrow || ctx->t_capacity >= worst || attempt >= 16)
            return fail(ctx, MDBG_ERR_TABLE_FULL, "count table (capacity %llu slots) is full; raise expected_distinct",
                        (unsigned long long)ctx->t_capacity);
        CKS(table_reset(ctx, std::min<uint64_t>(ctx->t_capacity * 4, worst)));
        for (const mdbg_ctx::Range& r : ctx->t_ranges) {
            uint64_t a_lo, a_hi;
            CKS(flat_range(ctx, r.read_lo, r.read_hi, &a_lo, &a_hi));
            CKS(launch_pass(ctx, a_lo, a_hi, next_k, false));
        }
        CKS(launch_pass(ctx, g_lo, g_hi, next_k, true, val_in, val_out));
    }
    clk.lap(PH_PASS);
    ctx->t_whole = whole;
    ctx->t_ranges.push_back(mdbg_ctx::Range{read_lo, read_hi});
    if (val_out) {
        std::swap(ctx->val_a, ctx->val_b);
        ctx->val_valid = true;
        ctx->val_k = ctx->t_k;
        ctx->val_gen = ctx->store_gen;
    } else {
        ctx->val_valid = false;                            // (a count pass included: the next k looks its global counts up)
    }
    if (!next_k && ctx->s_mins) ctx->distinct_ratio = (double)ctx->h_small->t_claims / (double)ctx->s_mins;
    return MDBG_OK;
}

mdbg_status mdbg_count_add_store(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_add_store before mdbg_count_begin");
    if (read_hi > ctx->s_reads) read_hi = ctx->s_reads;
    if (read_lo >= read_hi) return MDBG_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->t_merged = false;
    return count_pass(ctx, read_lo, read_hi, false);
}

mdbg_status mdbg_count_add(mdbg_ctx* ctx, const uint32_t* minimizers, const uint64_t* min_offsets, uint32_t n_reads) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_add before mdbg_count_begin");
    const uint64_t lo = ctx->s_reads;
    CKS(mdbg_store_append(ctx, minimizers, min_offsets, n_reads));
    return mdbg_count_add_store(ctx, lo, ctx->s_reads);
}

static mdbg_status table_stats(mdbg_ctx* ctx, uint32_t thr, TableStats* st, const Slot* table = nullptr, uint64_t capacity = 0) {
    cudaStream_t s = ctx->stream;
    PhaseClock clk(ctx);
    if (!table) { table = ctx->table.as<Slot>(); capacity = ctx->t_capacity; }
    launch_table_stats(table, capacity, thr, &ctx->d_small->stats, s);
    CKS(check_launch(ctx, "table_stats_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_small->stats, &ctx->d_small->stats, sizeof(TableStats), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *st = ctx->h_small->stats;
    clk.lap(PH_STATS);
    return MDBG_OK;
}

// dumpKminmer (CreateMdbg.hpp:3862-3869): drop abundance <= 1, and on the first
// pass abundance < min_abundance.
// Later passes (value tables of the next-k pass) keep every k-min-mer with refined abundance > 1 whatever
// --min-abundance says (`_isFirstPass && abundance < _minAbundance`, CreateMdbg.hpp:3868-3869).
static uint32_t count_threshold(const mdbg_ctx* ctx, uint32_t min_abundance) {
    if (ctx->t_value_mode) return 2;
    return min_abundance > 2 ? min_abundance : 2;
}

mdbg_status mdbg_count_stats(mdbg_ctx* ctx, uint32_t min_abundance, uint64_t* n_entries, uint64_t* n_distinct,
                             uint64_t* n_instances, uint64_t* checksum) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "no active count table");
    CK(cudaSetDevice(ctx->device));
    TableStats st;
    CKS(table_stats(ctx, count_threshold(ctx, min_abundance), &st));
    if (n_entries) *n_entries = st.n_entries;
    if (n_distinct) *n_distinct = st.n_distinct;
    if (n_instances) *n_instances = st.n_instances;
    if (checksum) *checksum = st.checksum;
    return MDBG_OK;
}

// statistics + compaction of the qualifying entries into o_hash / o_abund / o_vecs (device); *st_out = the statistics
static mdbg_status emit_table(mdbg_ctx* ctx, uint32_t thr, TableStats* st_out) {
    cudaStream_t s = ctx->stream;
    const uint32_t k = ctx->t_k;
    TableStats st;
    CKS(table_stats(ctx, thr, &st));
    const uint64_t n = st.n_entries;
    CKS(ensure(ctx, ctx->o_hash, (n + 1) * 16));
    CKS(ensure(ctx, ctx->o_abund, (n + 1) * 4));
    CKS(ensure(ctx, ctx->o_vecs, (n + 1) * 4 * k));
    CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
    EmitArgs e{};
    e.table = ctx->table.as<Slot>();
    e.capacity = ctx->t_capacity;
    e.min_count = thr;
    e.k = k;
    e.mins = ctx->s_min.as<uint32_t>();
    e.foreign_vecs = ctx->foreign_vecs.as<uint32_t>();
    e.out_hashes = ctx->o_hash.as<uint64_t>();
    e.out_abund = ctx->o_abund.as<uint32_t>();
    e.out_vecs = ctx->t_no_vecs ? nullptr : ctx->o_vecs.as<uint32_t>();
    e.cursor = &ctx->d_small->emit_cursor;
    PhaseClock clk(ctx);
    launch_table_emit(e, s);
    CKS(check_launch(ctx, "table_emit_kernel", 1));
    clk.lap(PH_EMIT);
    *st_out = st;
    return MDBG_OK;
}

mdbg_status mdbg_count_finalize_device(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_table_dev* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_finalize_device before mdbg_count_begin");
    CK(cudaSetDevice(ctx->device));
    TableStats st;
    CKS(emit_table(ctx, count_threshold(ctx, min_abundance), &st));
    out->k = ctx->t_k;
    out->n_entries = st.n_entries;
    out->d_hashes = ctx->o_hash.as<uint64_t>();
    out->d_abundances = ctx->o_abund.as<uint32_t>();
    out->d_kminmers = ctx->t_no_vecs ? nullptr : ctx->o_vecs.as<uint32_t>();
    out->n_instances = st.n_instances;
    out->n_distinct = st.n_distinct;
    out->checksum = st.checksum;
    out->n_rescued = st.n_rescued;
    return MDBG_OK;
}

mdbg_status mdbg_count_finalize(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_table_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_finalize before mdbg_count_begin");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t k = ctx->t_k;
    TableStats st;
    CKS(emit_table(ctx, count_threshold(ctx, min_abundance), &st));
    const uint64_t n = st.n_entries;
    CKS(ensure_pin(ctx, ctx->ho_hash, (n + 1) * 16));
    CKS(ensure_pin(ctx, ctx->ho_abund, (n + 1) * 4));
    CKS(ensure_pin(ctx, ctx->ho_vecs, (n + 1) * 4 * k));
    if (n) {
        CK(cudaMemcpyAsync(ctx->ho_hash.p, ctx->o_hash.p, n * 16, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->ho_abund.p, ctx->o_abund.p, n * 4, cudaMemcpyDeviceToHost, s));
        if (!ctx->t_no_vecs) CK(cudaMemcpyAsync(ctx->ho_vecs.p, ctx->o_vecs.p, n * 4 * k, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += n * (16 + 4 + (ctx->t_no_vecs ? 0ull : 4ull * k));
    out->k = k;
    out->n_entries = n;
    out->hashes = ctx->ho_hash.as<uint64_t>();
    out->abundances = ctx->ho_abund.as<uint32_t>();
    out->kminmers = ctx->t_no_vecs ? nullptr : ctx->ho_vecs.as<uint32_t>();
    out->n_instances = st.n_instances;
    out->n_distinct = st.n_distinct;
    out->checksum = st.checksum;
    out->n_rescued = st.n_rescued;
    return MDBG_OK;
}

static mdbg_status replicate_into_prev(mdbg_ctx* ctx, const Slot* src, uint64_t src_cap, uint32_t thr);

// Multi-rank rescue (collective, after mdbg_count_merge).  The decision per read needs the abundance of every
// window, the flag belongs to the window's owner:
//   1. the solid k-min-mers (abundance >= 2) of all ranks are replicated into the lookup table (the same exchange
//      the next-k pass uses);
//   2. every rank runs the rescue decision on ITS reads against that table and collects the normalized vectors of
//      the non-solid windows of rescued reads (each occurs exactly once in the whole read set);
//   3. the vectors are bucketed by owner rank and exchanged in one grouped all-to-all;
//   4. the owner flags the slots (count 1) of the vectors it received.
static mdbg_status count_rescue_all_ranks(mdbg_ctx* ctx, uint64_t* n_reads_rescued) {
    if (!ctx->nccl_comm) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_rescue before mdbg_comm_init");
    if (!ctx->t_merged || ctx->t_value_mode || ctx->t_no_vecs)
        return fail(ctx, MDBG_ERR_STATE, "multi-rank mdbg_count_rescue needs the merged count table: call mdbg_count_merge first");
    cudaStream_t s = ctx->stream;
    const uint32_t R = (uint32_t)ctx->n_ranks, k = ctx->t_k;
    // step 1: the lookup table is built in ctx->prev_table; a previous-k table the caller loaded is set aside and
    // put back at the end (mdbg_prev_load + refined-abundance patches survive a rescue)
    struct PrevStash {
        mdbg_ctx* c; DevBuf table; uint64_t cap; uint32_t thr, k; bool pure, repl;
        explicit PrevStash(mdbg_ctx* ctx) : c(ctx), table(ctx->prev_table), cap(ctx->prev_capacity), thr(ctx->prev_min_count),
                                            k(ctx->prev_k), pure(ctx->prev_pure), repl(ctx->prev_replicated) {
            c->prev_table = c->rescue_table; c->prev_capacity = 0;
        }
        ~PrevStash() {
            c->rescue_table = c->prev_table; c->prev_table = table; c->prev_capacity = cap; c->prev_min_count = thr;
            c->prev_k = k; c->prev_pure = pure; c->prev_replicated = repl;
        }
    } stash(ctx);
    CKS(replicate_into_prev(ctx, ctx->table.as<Slot>(), ctx->t_capacity, 2));
    // step 2
    const uint64_t max_windows = ctx->s_mins + 1;
    CKS(ensure(ctx, ctx->m_recv_vecs, max_windows * 4 * k));        // flat list of collected vectors
    CK(cudaMemsetAsync(&ctx->d_small->n_changed, 0, sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, sizeof(unsigned long long), s));
    if (ctx->s_reads) {
        RescueArgs a{};
        a.mins = ctx->s_min.as<uint32_t>();
        a.offs = ctx->s_off.as<uint64_t>();
        a.n_reads = ctx->s_reads;
        a.k = k;
        a.table = ctx->prev_table.as<Slot>();
        a.mask = ctx->prev_capacity - 1;
        a.n_reads_rescued = &ctx->d_small->n_changed;
        a.out_vecs = ctx->m_recv_vecs.as<uint32_t>();
        a.out_cursor = &ctx->d_small->emit_cursor;
        launch_rescue(a, s);
        CKS(check_launch(ctx, "rescue_kernel(collect)", 1));
    }
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], &ctx->d_small->n_changed, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[1], &ctx->d_small->emit_cursor, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_reads_rescued) *n_reads_rescued = ctx->h_scalar[0];
    const uint64_t n_list = ctx->h_scalar[1];
    // step 3: bucket by owner, exchange
    CKS(ensure(ctx, ctx->m_bucket, (size_t)(2 * R + R * R) * 8));
    uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
    uint64_t* d_base = d_cnt + R;
    CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
    BucketVecArgs b{};
    b.vecs = ctx->m_recv_vecs.as<uint32_t>();
    b.n = n_list;
    b.k = k;
    b.n_ranks = R;
    b.bucket_count = reinterpret_cast<unsigned long long*>(d_cnt);
    b.pass = 1;
    launch_bucket_vecs(b, s);
    CKS(check_launch(ctx, "bucket_vecs_kernel(count)", n_list ? 1 : 0));
    OwnerExchange xr;
    CKS(xr.plan(ctx));
    const uint64_t recv_total = xr.recv_total;
    CKS(ensure(ctx, ctx->m_send_vecs, (xr.send_total + 1) * 4 * k));
    CKS(ensure(ctx, ctx->o_vecs, (recv_total + 1) * 4 * k));       // receive side
    CKS(xr.upload_bases(ctx));
    b.bucket_base = d_base;
    b.out_vecs = ctx->m_send_vecs.as<uint32_t>();
    b.pass = 2;
    launch_bucket_vecs(b, s);
    CKS(check_launch(ctx, "bucket_vecs_kernel(scatter)", n_list ? 1 : 0));
    CKS(xr.run(ctx, ctx->m_send_vecs.p, ctx->o_vecs.p, (size_t)4 * k));
    // step 4
    CK(cudaMemsetAsync(&ctx->d_small->n_flagged, 0, sizeof(unsigned long long), s));
    launch_rescue_flag(ctx->o_vecs.as<uint32_t>(), recv_total, k, ctx->table.as<Slot>(), ctx->t_capacity - 1,
                       &ctx->d_small->n_flagged, s);
    CKS(check_launch(ctx, "rescue_flag_kernel", recv_total ? 1 : 0));
    CK(cudaMemcpyAsync(&ctx->h_scalar[2], &ctx->d_small->n_flagged, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_scalar[2])
        return fail(ctx, MDBG_ERR_STATE, "multi-rank rescue: %llu rescued k-min-mers have no slot on their owner "
                    "(were reads added after mdbg_count_merge?)", (unsigned long long)ctx->h_scalar[2]);
    return MDBG_OK;
}

mdbg_status mdbg_count_rescue(mdbg_ctx* ctx, uint64_t* n_reads_rescued) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n_reads_rescued) *n_reads_rescued = 0;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_rescue before mdbg_count_begin");
    if (ctx->n_ranks > 1) {
        CK(cudaSetDevice(ctx->device));
        return count_rescue_all_ranks(ctx, n_reads_rescued);
    }
    if (ctx->s_reads == 0) return MDBG_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(&ctx->d_small->n_changed, 0, sizeof(unsigned long long), s));
    RescueArgs a{};
    a.mins = ctx->s_min.as<uint32_t>();
    a.offs = ctx->s_off.as<uint64_t>();
    a.n_reads = ctx->s_reads;
    a.k = ctx->t_k;
    a.table = ctx->table.as<Slot>();
    a.mask = ctx->t_capacity - 1;
    a.n_reads_rescued = &ctx->d_small->n_changed;
    launch_rescue(a, s);
    CKS(check_launch(ctx, "rescue_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], &ctx->d_small->n_changed, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n_reads_rescued) *n_reads_rescued = ctx->h_scalar[0];
    return MDBG_OK;
}

static mdbg_status prev_alloc(mdbg_ctx* ctx, uint64_t expect) {
    if (expect < 512) expect = 512;
    const uint64_t cap = pow2ceil(expect * 2);
    CKS(ensure_table_buf(ctx, ctx->prev_table, cap * sizeof(Slot)));
    CK(cudaMemsetAsync(ctx->prev_table.p, 0, cap * sizeof(Slot), ctx->stream));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), ctx->stream));
    ctx->prev_capacity = cap;
    ctx->prev_min_count = 0;           // a table built from listed pairs: every entry counts
    return MDBG_OK;
}

static mdbg_status check_full(mdbg_ctx* ctx, const char* what) {
    CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_small->full_flag) return fail(ctx, MDBG_ERR_TABLE_FULL, "%s: table full", what);
    return MDBG_OK;
}

// Multi-rank: after mdbg_count_merge every rank holds the keys it owns with their global abundances.  A pass that
// looks up arbitrary (k-1)-min-mers needs all of them, so the qualifying (hash, abundance) pairs of all ranks are
// replicated: local emit -> all-gather of the counts -> one grouped exchange of the pairs over NVLink -> insert into
// the previous-k table of every rank ("replicate if it fits": 20 B per solid k-min-mer).  `src` is the
// rank's owned table (the current table, or the one mdbg_prev_from_current set aside); the result is ctx->prev_table.
static mdbg_status replicate_into_prev(mdbg_ctx* ctx, const Slot* src, uint64_t src_cap, uint32_t thr) {
    if (!ctx->nccl_comm) return fail(ctx, MDBG_ERR_STATE, "previous-k replication before mdbg_comm_init");
    cudaStream_t s = ctx->stream;
    const uint32_t R = (uint32_t)ctx->n_ranks;
    PhaseClock clk(ctx);
    TableStats st;
    CKS(table_stats(ctx, thr, &st, src, src_cap));
    const uint64_t n_local = st.n_entries;
    CKS(ensure(ctx, ctx->o_hash, (n_local + 1) * 16));
    CKS(ensure(ctx, ctx->o_abund, (n_local + 1) * 4));
    CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
    EmitArgs e{};
    e.table = src;
    e.capacity = src_cap;
    e.min_count = thr;
    e.k = ctx->t_k;
    e.mins = ctx->s_min.as<uint32_t>();
    e.foreign_vecs = ctx->foreign_vecs.as<uint32_t>();
    e.out_hashes = ctx->o_hash.as<uint64_t>();
    e.out_abund = ctx->o_abund.as<uint32_t>();
    e.out_vecs = nullptr;
    e.cursor = &ctx->d_small->emit_cursor;
    launch_table_emit(e, s);
    CKS(check_launch(ctx, "table_emit_kernel", 1));
    clk.lap(PH_PREV_EMIT);
    // every rank's pair count
    CKS(ensure(ctx, ctx->m_bucket, (size_t)(2 * R + R * R) * 8));
    uint64_t* d_mine = ctx->m_bucket.as<uint64_t>();
    uint64_t* d_all = d_mine + R;
    ctx->h_scalar[3] = n_local;
    CK(cudaMemcpyAsync(d_mine, &ctx->h_scalar[3], 8, cudaMemcpyHostToDevice, s));
    NK(g_nccl.AllGather(d_mine, d_all, 1, NCCL_UINT64, ctx->nccl_comm, s));
    std::vector<uint64_t> cnt(R), base(R);
    CK(cudaMemcpyAsync(cnt.data(), d_all, (size_t)R * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (uint32_t d = 0; d < R; d++) { base[d] = total; total += cnt[d]; }
    CKS(ensure(ctx, ctx->prev_stage_h, (total + 1) * 16));
    CKS(ensure(ctx, ctx->prev_stage_a, (total + 1) * 4));
    uint64_t* g_hash = ctx->prev_stage_h.as<uint64_t>();
    uint32_t* g_abund = ctx->prev_stage_a.as<uint32_t>();
    clk.lap(PH_PREV_PLAN);
    NK(g_nccl.GroupStart());
    for (uint32_t d = 0; d < R; d++) {
        if ((int)d == ctx->rank) continue;
        if (n_local) {
            NK(g_nccl.Send(ctx->o_hash.p, n_local * 16, NCCL_UINT8, (int)d, ctx->nccl_comm, s));
            NK(g_nccl.Send(ctx->o_abund.p, n_local * 4, NCCL_UINT8, (int)d, ctx->nccl_comm, s));
        }
        if (cnt[d]) {
            NK(g_nccl.Recv(g_hash + 2 * base[d], cnt[d] * 16, NCCL_UINT8, (int)d, ctx->nccl_comm, s));
            NK(g_nccl.Recv(g_abund + base[d], cnt[d] * 4, NCCL_UINT8, (int)d, ctx->nccl_comm, s));
        }
    }
    NK(g_nccl.GroupEnd());
    if (n_local) {
        CK(cudaMemcpyAsync(g_hash + 2 * base[ctx->rank], ctx->o_hash.p, n_local * 16, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(g_abund + base[ctx->rank], ctx->o_abund.p, n_local * 4, cudaMemcpyDeviceToDevice, s));
    }
    clk.lap(PH_PREV_EXCHANGE);
    CKS(prev_alloc(ctx, total));
    if (total) {
        PrevLoadArgs a{};
        a.hashes = g_hash;
        a.abund = g_abund;
        a.n = total;
        a.prev = ctx->prev_table.as<Slot>();
        a.prev_mask = ctx->prev_capacity - 1;
        a.full_flag = &ctx->d_small->full_flag;
        launch_prev_load(a, s);
        CKS(check_launch(ctx, "prev_load_kernel", 1));
    }
    const mdbg_status st_full = check_full(ctx, "previous-k replication (all ranks)");
    clk.lap(PH_PREV_INSERT);
    ctx->prev_replicated = true;
    return st_full;
}

mdbg_status mdbg_prev_from_current(mdbg_ctx* ctx, uint32_t min_abundance) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_prev_from_current without a count table");
    // Several ranks: an occurrence-count table must be merged first (abundances are sums over all ranks).  A VALUE table
    // of a whole-store pass may stay rank-local: its values are functions of the key, and every (k-1)-min-mer the
    // rank's next pass can ask for is a window of the rank's own reads, i.e. already in this table.
    const bool local_value_table = ctx->t_value_mode && ctx->t_whole && !ctx->t_merged;
    if (ctx->n_ranks > 1 && !ctx->t_merged && !local_value_table)
        return fail(ctx, MDBG_ERR_STATE, "multi-rank mdbg_prev_from_current needs the merged table: call mdbg_count_merge first");
    CK(cudaSetDevice(ctx->device));
    const uint32_t thr = count_threshold(ctx, min_abundance);
    // The current table BECOMES the previous-k table -- no copy, no scan.  Entries the reference would not have dumped
    // (abundance below the threshold and not rescued) are skipped at lookup time (NextKArgs::prev_min_count).  The
    // old previous-k buffer is recycled as the next current table.  With several ranks the table holds the OWNED
    // entries only; they are exchanged (replicate_into_prev) when -- and only when -- the next pass needs lookups.
    std::swap(ctx->prev_table, ctx->table);
    ctx->prev_capacity = ctx->t_capacity;
    ctx->prev_min_count = thr;
    ctx->prev_k = ctx->t_k;
    ctx->prev_pure = ctx->t_whole;
    ctx->prev_replicated = ctx->n_ranks == 1 || local_value_table;
    ctx->t_capacity = 0;
    ctx->t_active = false;
    ctx->t_ranges.clear();
    return MDBG_OK;
}

mdbg_status mdbg_prev_load(mdbg_ctx* ctx, const uint64_t* hashes, const uint32_t* abundances, uint64_t n, int clear) {
    if (!ctx) return MDBG_ERR_ARG;
    if (n && (!hashes || !abundances)) return fail(ctx, MDBG_ERR_ARG, "null table arrays");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    ctx->prev_pure = false;                                // host pairs: the next pass must look them up
    if (!clear && ctx->n_ranks > 1 && !ctx->prev_replicated && ctx->prev_capacity) {
        std::swap(ctx->prev_table, ctx->prev_src);         // patches go on top of the COMPLETE table
        CKS(replicate_into_prev(ctx, ctx->prev_src.as<Slot>(), ctx->prev_capacity, ctx->prev_min_count));
    }
    if (clear || ctx->prev_capacity == 0) CKS(prev_alloc(ctx, n));
    ctx->prev_replicated = true;
    if (n == 0) return MDBG_OK;
    CKS(ensure(ctx, ctx->prev_stage_h, n * 16));
    CKS(ensure(ctx, ctx->prev_stage_a, n * 4));
    CK(cudaMemcpyAsync(ctx->prev_stage_h.p, hashes, n * 16, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->prev_stage_a.p, abundances, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), s));
    PrevLoadArgs a{};
    a.hashes = ctx->prev_stage_h.as<uint64_t>();
    a.abund = ctx->prev_stage_a.as<uint32_t>();
    a.n = n;
    a.prev = ctx->prev_table.as<Slot>();
    a.prev_mask = ctx->prev_capacity - 1;
    a.full_flag = &ctx->d_small->full_flag;
    launch_prev_load(a, s);
    CKS(check_launch(ctx, "prev_load_kernel", 1));
    return check_full(ctx, "mdbg_prev_load (patching needs free slots: load the base table with clear=1 first)");
}

mdbg_status mdbg_count_add_store_next_k(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_add_store_next_k before mdbg_count_begin");
    if (ctx->prev_capacity == 0) return fail(ctx, MDBG_ERR_STATE, "no previous-k table (mdbg_prev_load / mdbg_prev_from_current)");
    if (ctx->t_k < 3) return fail(ctx, MDBG_ERR_ARG, "k must be >= 3 for a next-k pass");
    // the table now holds abundance VALUES (also on a rank that has no read to add: it still owns keys in the merge)
    ctx->t_value_mode = true;
    ctx->t_merged = false;
    if (read_hi > ctx->s_reads) read_hi = ctx->s_reads;
    if (read_lo > read_hi) read_lo = read_hi;
    // no early return for an empty range: with several ranks the pass may start with a collective (previous-k
    // replication), and a rank without reads must take the same decisions as the others
    CK(cudaSetDevice(ctx->device));
    return count_pass(ctx, read_lo, read_hi, true);
}

// CreateMdbg::EdgeIndexer (CreateMdbg.hpp:4010-4232; first step of indexEdges, CreateMdbg.cpp:1177-1187): the
// dereplicated hash128 of the normalized (k-1)-prefix and (k-1)-suffix of every node of the current table.
static mdbg_status edges_index_impl(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_edges_out* out, bool to_host,
                                    bool want_slot_node = false) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_edges_index before mdbg_count_begin");
    if (ctx->t_k < 2) return fail(ctx, MDBG_ERR_ARG, "k must be >= 2");
    if (ctx->t_no_vecs) return fail(ctx, MDBG_ERR_STATE, "mdbg_edges_index needs the k-min-mer vectors: merge with mdbg_count_merge, not mdbg_count_merge_hashes");
    if (ctx->n_ranks > 1 && (!ctx->nccl_comm || !ctx->t_merged))
        return fail(ctx, MDBG_ERR_STATE, "multi-rank mdbg_edges_index needs the merged table: call mdbg_count_merge first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t thr = count_threshold(ctx, min_abundance);
    TableStats st;
    CKS(table_stats(ctx, thr, &st));
    uint64_t expect = 2 * st.n_entries;                          // at most two keys per node
    if (expect < 512) expect = 512;
    const uint64_t cap = pow2ceil(expect * 2);
    CKS(ensure(ctx, ctx->edge_table, cap * sizeof(Slot)));
    CK(cudaMemsetAsync(ctx->edge_table.p, 0, cap * sizeof(Slot), s));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), s));
    EdgeArgs e{};
    e.table = ctx->table.as<Slot>();
    e.capacity = ctx->t_capacity;
    e.min_count = thr;
    e.k = ctx->t_k;
    e.mins = ctx->s_min.as<uint32_t>();
    e.foreign_vecs = ctx->foreign_vecs.as<uint32_t>();
    e.edges = ctx->edge_table.as<Slot>();
    e.edge_mask = cap - 1;
    e.full_flag = &ctx->d_small->full_flag;
    // the slots of the nodes, once: the insert / values kernels visit 2 M nodes instead of scanning a 32 M-slot table twice
    if (ctx->t_capacity <= 0xFFFFFFF0ull && st.n_entries) {
        CKS(ensure(ctx, ctx->u_node_slot, (st.n_entries + 1) * 4));
        if (want_slot_node) CKS(ensure(ctx, ctx->u_slot_node, (ctx->t_capacity + 1) * 4));      // the inverse map, for the unitig links
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
        launch_unitig_nodes(ctx->table.as<Slot>(), ctx->t_capacity, thr, want_slot_node ? ctx->u_slot_node.as<uint32_t>() : nullptr,
                            ctx->u_node_slot.as<uint32_t>(), &ctx->d_small->emit_cursor, s);
        e.node_slot = ctx->u_node_slot.as<uint32_t>();
        e.n_nodes = st.n_entries;
    }
    launch_edge_insert(e, s);
    CKS(check_launch(ctx, "edge_insert_kernel", e.node_slot ? 2 : 1));
    CKS(check_full(ctx, "mdbg_edges_index"));
    uint64_t set_cap = cap;
    if (ctx->n_ranks > 1) {
        // Several ranks: the nodes a rank owns give it a local key set; a key can come from nodes of different
        // ranks, so the distinct local keys travel to owner(key) in one grouped all-to-all and the owner
        // dereplicates what it receives.  Afterwards every rank holds (and returns) the keys it owns.
        const uint32_t R = (uint32_t)ctx->n_ranks;
        launch_table_stats(ctx->edge_table.as<Slot>(), cap, 1u, &ctx->d_small->stats, s);
        CKS(check_launch(ctx, "table_stats_kernel", 1));
        CK(cudaMemcpyAsync(&ctx->h_small->stats, &ctx->d_small->stats, sizeof(TableStats), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        const uint64_t n_local = ctx->h_small->stats.n_entries;
        CKS(ensure(ctx, ctx->o_hash, (n_local + 1) * 16));
        CKS(ensure(ctx, ctx->o_abund, (n_local + 1) * 4));
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
        EmitArgs el{};
        el.table = ctx->edge_table.as<Slot>();
        el.capacity = cap;
        el.min_count = 1;
        el.k = ctx->t_k - 1;
        el.out_hashes = ctx->o_hash.as<uint64_t>();
        el.out_abund = ctx->o_abund.as<uint32_t>();
        el.out_vecs = nullptr;
        el.cursor = &ctx->d_small->emit_cursor;
        launch_table_emit(el, s);
        CKS(check_launch(ctx, "table_emit_kernel", 1));
        CKS(ensure(ctx, ctx->m_bucket, (size_t)(2 * R + R * R) * 8));
        uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
        uint64_t* d_base = d_cnt + R;
        CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
        BucketKeyArgs b{};
        b.keys = ctx->o_hash.as<uint64_t>();
        b.n = n_local;
        b.n_ranks = R;
        b.rec_words = 2;
        b.bucket_count = reinterpret_cast<unsigned long long*>(d_cnt);
        b.pass = 1;
        launch_bucket_keys(b, s);
        CKS(check_launch(ctx, "bucket_keys_kernel(count)", n_local ? 1 : 0));
        OwnerExchange xk;
        CKS(xk.plan(ctx));
        const uint64_t recv_total = xk.recv_total;
        CKS(ensure(ctx, ctx->prev_stage_h, (xk.send_total + 1) * 16));     // send side (free at this point of the flow)
        CKS(ensure(ctx, ctx->m_recv_vecs, (recv_total + 1) * 16));         // receive side
        CKS(xk.upload_bases(ctx));
        b.bucket_base = d_base;
        b.out_keys = ctx->prev_stage_h.as<uint64_t>();
        b.pass = 2;
        launch_bucket_keys(b, s);
        CKS(check_launch(ctx, "bucket_keys_kernel(scatter)", n_local ? 1 : 0));
        CKS(xk.run(ctx, ctx->prev_stage_h.p, ctx->m_recv_vecs.p, 16));
        set_cap = pow2ceil((recv_total < 512 ? 512 : recv_total) * 2);
        CKS(ensure(ctx, ctx->edge_table, set_cap * sizeof(Slot)));
        CK(cudaMemsetAsync(ctx->edge_table.p, 0, set_cap * sizeof(Slot), s));
        CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), s));
        launch_insert_keys(ctx->m_recv_vecs.as<uint64_t>(), recv_total, ctx->edge_table.as<Slot>(), set_cap - 1,
                           &ctx->d_small->full_flag, s);
        CKS(check_launch(ctx, "insert_keys_kernel", recv_total ? 1 : 0));
        CKS(check_full(ctx, "mdbg_edges_index (owner side)"));
    }
    // the set's statistics and contents: every entry has value 1
    launch_table_stats(ctx->edge_table.as<Slot>(), set_cap, 1u, &ctx->d_small->stats, s);
    CKS(check_launch(ctx, "table_stats_kernel", 1));
    CK(cudaMemcpyAsync(&ctx->h_small->stats, &ctx->d_small->stats, sizeof(TableStats), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t n = ctx->h_small->stats.n_entries;
    CKS(ensure(ctx, ctx->o_hash, (n + 1) * 16));
    CKS(ensure(ctx, ctx->o_abund, (n + 1) * 4));
    CKS(ensure_pin(ctx, ctx->ho_hash, (n + 1) * 16));
    CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
    // the keys come out together with their order-free edge values (indexEdge / successorExists)
    CKS(ensure(ctx, ctx->edge_vals, set_cap * 16));
    CKS(ensure(ctx, ctx->o_edge_vals, (n + 1) * 16));
    CKS(ensure_pin(ctx, ctx->ho_edge_vals, (n + 1) * 16));
    CK(cudaMemsetAsync(ctx->edge_vals.p, 0, set_cap * 16, s));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), s));
    if (ctx->n_ranks == 1) {
        e.edge_vals = ctx->edge_vals.as<unsigned long long>();
        launch_edge_values(e, s);
        CKS(check_launch(ctx, "edge_values_kernel", 1));
        CKS(check_full(ctx, "mdbg_edges_index (values: a key is missing from the set)"));
    } else {
        // several ranks: the two offers of every owned node travel to the owner of their key as 3-word records and
        // are folded into that key's class words there
        const uint32_t R = (uint32_t)ctx->n_ranks;
        const uint64_t off_cap = 2 * st.n_entries;
        CKS(ensure(ctx, ctx->m_send_vecs, (off_cap + 1) * 24));
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
        launch_edge_offers(e, ctx->m_send_vecs.as<uint64_t>(), &ctx->d_small->emit_cursor, s);
        CKS(check_launch(ctx, "edge_offers_kernel", 1));
        CK(cudaMemcpyAsync(&ctx->h_scalar[4], &ctx->d_small->emit_cursor, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (ctx->h_scalar[4] != off_cap)
            return fail(ctx, MDBG_ERR_STATE, "mdbg_edges_index: %llu offers for %llu nodes", (unsigned long long)ctx->h_scalar[4],
                        (unsigned long long)st.n_entries);
        uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
        uint64_t* d_base = d_cnt + R;
        CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
        BucketKeyArgs b{};
        b.keys = ctx->m_send_vecs.as<uint64_t>();
        b.n = off_cap;
        b.n_ranks = R;
        b.rec_words = 3;
        b.bucket_count = reinterpret_cast<unsigned long long*>(d_cnt);
        b.pass = 1;
        launch_bucket_keys(b, s);
        CKS(check_launch(ctx, "bucket_keys_kernel(count)", off_cap ? 1 : 0));
        OwnerExchange xo;
        CKS(xo.plan(ctx));
        const uint64_t recv_total = xo.recv_total;
        CKS(ensure(ctx, ctx->prev_stage_h, (xo.send_total + 1) * 24));
        CKS(ensure(ctx, ctx->m_recv_vecs, (recv_total + 1) * 24));
        CKS(xo.upload_bases(ctx));
        b.bucket_base = d_base;
        b.out_keys = ctx->prev_stage_h.as<uint64_t>();
        b.pass = 2;
        launch_bucket_keys(b, s);
        CKS(check_launch(ctx, "bucket_keys_kernel(scatter)", off_cap ? 1 : 0));
        CKS(xo.run(ctx, ctx->prev_stage_h.p, ctx->m_recv_vecs.p, 24));
        launch_edge_apply_offers(ctx->m_recv_vecs.as<uint64_t>(), recv_total, ctx->edge_table.as<Slot>(), set_cap - 1,
                                 ctx->edge_vals.as<unsigned long long>(), &ctx->d_small->full_flag, s);
        CKS(check_launch(ctx, "edge_apply_offers_kernel", recv_total ? 1 : 0));
        CKS(check_full(ctx, "mdbg_edges_index (values: an offer arrived for a key its owner does not hold)"));
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
    }
    ctx->edge_cap = set_cap;
    if (!to_host) {                                              // the caller works on the device-resident set (unitigs)
        out->k = ctx->t_k;
        out->n_edges = n;
        out->n_nodes = st.n_entries;
        return MDBG_OK;
    }
    launch_edge_emit(ctx->edge_table.as<Slot>(), ctx->edge_vals.as<unsigned long long>(), set_cap, ctx->o_hash.as<uint64_t>(),
                     ctx->o_edge_vals.as<unsigned long long>(), &ctx->d_small->emit_cursor, s);
    CKS(check_launch(ctx, "edge_emit_kernel", 1));
    if (n) CK(cudaMemcpyAsync(ctx->ho_edge_vals.p, ctx->o_edge_vals.p, n * 16, cudaMemcpyDeviceToHost, s));
    out->values = ctx->ho_edge_vals.as<uint64_t>();
    ctx->d2h_bytes += n * 16;
    if (n) CK(cudaMemcpyAsync(ctx->ho_hash.p, ctx->o_hash.p, n * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += n * 16;
    out->k = ctx->t_k;
    out->n_edges = n;
    out->hashes = ctx->ho_hash.as<uint64_t>();
    out->checksum = ctx->h_small->stats.checksum;                // sum of the low words (value 1 each)
    out->n_nodes = st.n_entries;
    return MDBG_OK;
}

mdbg_status mdbg_edges_index(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_edges_out* out) {
    return edges_index_impl(ctx, min_abundance, out, true);
}

// CreateMdbg::computeUnitigNodes + computeDeterministicUnitigs (CreateMdbg.cpp:1521-1598, 1001-1043) on the device:
// node ids -> edge set + class values (as mdbg_edges_index) -> unitig links between oriented nodes -> list ranking by
// pointer jumping (cycles cut at their smallest hash128) -> minimizer sequences -> normalize + hash128 per unitig.
// The host only sorts the unitig indices by hash (the reference's deterministic order).
mdbg_status mdbg_unitigs_build(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_unitigs_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_unitigs_build before mdbg_count_begin");
    if (ctx->n_ranks > 1)
        return fail(ctx, MDBG_ERR_STATE, "mdbg_unitigs_build is a single-context call (the walk needs the whole node set on one device)");
    if (ctx->t_capacity > 0xFFFFFFF0ull) return fail(ctx, MDBG_ERR_ARG, "table too large for 32-bit node ids");
    mdbg_edges_out eo;
    memset(&eo, 0, sizeof eo);
    CKS(edges_index_impl(ctx, min_abundance, &eo, false, true)); // device-resident edge set + class values, node ids
    cudaStream_t s = ctx->stream;
    const uint32_t thr = count_threshold(ctx, min_abundance);
    const uint64_t cap = ctx->t_capacity;
    const uint64_t n = eo.n_nodes;
    if (2 * n + 2 > 0x7FFFFFF0ull) return fail(ctx, MDBG_ERR_ARG, "too many nodes for 32-bit oriented node ids");
    const uint32_t n2 = (uint32_t)(2 * n);
    out->k = ctx->t_k;
    out->n_nodes = n;
    CKS(ensure(ctx, ctx->u_next, ((uint64_t)n2 + 1) * 4));
    CKS(ensure(ctx, ctx->u_pair, ((uint64_t)n2 + 1) * 8));
    CKS(ensure(ctx, ctx->u_len, ((uint64_t)n2 + 1) * 4));
    CKS(ensure(ctx, ctx->u_size, ((uint64_t)n2 + 1) * 4));
    CKS(ensure(ctx, ctx->u_flag, ((uint64_t)n2 + 1) * 4));
    CKS(ensure(ctx, ctx->u_cychead, (uint64_t)n2 + 1));
    CKS(ensure(ctx, ctx->u_seqoff, ((uint64_t)n2 + 2) * 8));
    CKS(ensure(ctx, ctx->u_idx, ((uint64_t)n2 + 2) * 8));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems(n2 + 1) * sizeof(uint64_t)));
    CK(cudaMemsetAsync(&ctx->d_small->full_flag, 0, sizeof(uint32_t), s));
    CK(cudaMemsetAsync(ctx->u_cychead.p, 0, (uint64_t)n2 + 1, s));
    UnitigArgs ua{};
    ua.table = ctx->table.as<Slot>(); ua.mask = cap - 1;
    ua.node_slot = ctx->u_node_slot.as<uint32_t>(); ua.slot_node = ctx->u_slot_node.as<uint32_t>();
    ua.n_nodes = (uint32_t)n; ua.k = ctx->t_k;
    ua.mins = ctx->s_min.as<uint32_t>(); ua.foreign_vecs = ctx->foreign_vecs.as<uint32_t>();
    ua.edges = ctx->edge_table.as<Slot>(); ua.edge_mask = ctx->edge_cap - 1;
    ua.edge_vals = ctx->edge_vals.as<unsigned long long>();
    ua.next = ctx->u_next.as<uint32_t>(); ua.error_flag = &ctx->d_small->full_flag;
    launch_unitig_link(ua, s);
    unsigned long long* pair = ctx->u_pair.as<unsigned long long>();
    launch_unitig_rank_init(ua.next, n2, pair, ctx->u_len.as<uint32_t>(), s);
    CKS(check_launch(ctx, "unitig_link_kernel + unitig_rank_init_kernel", n ? 2 : 0));
    CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_small->full_flag)
        return fail(ctx, MDBG_ERR_STATE, "mdbg_unitigs_build: %s", ctx->h_small->full_flag == 1 ? "a node's edge key is missing from the edge set"
                                                                                                  : "an edge offer names a k-min-mer that is not a node");
    // pointer jumping: the pointer distance of an open node at least doubles per launch, so paths are resolved after
    // ceil(log2(2n)) launches; what is still open then lies on cycles
    int max_rounds = 2;
    while ((1ull << (max_rounds - 2)) < (uint64_t)n2 + 1) max_rounds++;
    auto rank_all = [&](uint64_t* n_open_out) -> mdbg_status {
        uint64_t n_open = n2;
        for (int r = 0; r < max_rounds && n_open; r++) {
            CK(cudaMemsetAsync(&ctx->d_small->n_flagged, 0, 8, s));
            launch_unitig_jump(pair, n2, 2, &ctx->d_small->n_flagged, s);
            CKS(check_launch(ctx, "unitig_jump_kernel", n2 ? 1 : 0));
            CK(cudaMemcpyAsync(&ctx->h_scalar[1], &ctx->d_small->n_flagged, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            n_open = ctx->h_scalar[1];
        }
        *n_open_out = n_open;
        return MDBG_OK;
    };
    uint64_t n_open = 0;
    CKS(rank_all(&n_open));
    out->n_cycle_nodes = n_open;
    if (n_open) {
        const uint32_t m = (uint32_t)n_open;
        CKS(ensure(ctx, ctx->u_cyclist, ((uint64_t)m + 1) * 4));
        CKS(ensure(ctx, ctx->u_cycpos, ((uint64_t)n2 + 1) * 4));
        CKS(ensure(ctx, ctx->u_best, 2 * ((uint64_t)m + 1) * 24));
        CKS(ensure(ctx, ctx->u_jump, 2 * ((uint64_t)m + 1) * 4));
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
        launch_unitig_cycle_list(pair, n2, ctx->u_cyclist.as<uint32_t>(), ctx->u_cycpos.as<uint32_t>(), &ctx->d_small->emit_cursor, s);
        uint64_t* best[2] = {ctx->u_best.as<uint64_t>(), ctx->u_best.as<uint64_t>() + 3 * ((uint64_t)m + 1)};
        uint32_t* jump[2] = {ctx->u_jump.as<uint32_t>(), ctx->u_jump.as<uint32_t>() + ((uint64_t)m + 1)};
        launch_unitig_cycle_init(ua, ctx->u_cyclist.as<uint32_t>(), ctx->u_cycpos.as<uint32_t>(), m, best[0], jump[0], s);
        int cur = 0, rounds = 1;
        while ((1ull << rounds) < (uint64_t)m + 1) rounds++;
        for (int r = 0; r < rounds; r++, cur ^= 1)
            launch_unitig_cycle_min(best[cur], jump[cur], m, best[cur ^ 1], jump[cur ^ 1], s);
        launch_unitig_cycle_cut(ua.next, ctx->u_cyclist.as<uint32_t>(), best[cur], m, pair, ctx->u_cychead.as<uint8_t>(), s);
        CKS(check_launch(ctx, "unitig_cycle_list/init/min/cut kernels", 3 + rounds));
        CKS(rank_all(&n_open));
        if (n_open) return fail(ctx, MDBG_ERR_STATE, "mdbg_unitigs_build: %llu oriented nodes are still unranked after the cycle cut",
                                (unsigned long long)n_open);
    }
    launch_unitig_len(pair, n2, ctx->u_len.as<uint32_t>(), s);
    launch_unitig_select(pair, ctx->u_cychead.as<uint8_t>(), ctx->u_len.as<uint32_t>(), n2, ctx->t_k, ctx->u_size.as<uint32_t>(),
                         ctx->u_flag.as<uint32_t>(), s);
    launch_scan_u32_to_u64(ctx->u_size.as<uint32_t>(), ctx->u_seqoff.as<uint64_t>(), n2, ctx->scan_scratch.as<uint64_t>(), s);
    launch_scan_u32_to_u64(ctx->u_flag.as<uint32_t>(), ctx->u_idx.as<uint64_t>(), n2, ctx->scan_scratch.as<uint64_t>(), s);
    CKS(check_launch(ctx, "unitig_len_kernel + unitig_select_kernel + scans", n2 ? 8 : 0));
    CK(cudaMemcpyAsync(&ctx->h_scalar[2], ctx->u_seqoff.as<uint64_t>() + n2, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[3], ctx->u_idx.as<uint64_t>() + n2, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t total = n2 ? ctx->h_scalar[2] : 0, nu = n2 ? ctx->h_scalar[3] : 0;
    CKS(ensure(ctx, ctx->u_mins, (total + 1) * 4));
    CKS(ensure(ctx, ctx->u_off, (nu + 2) * 8));
    CKS(ensure(ctx, ctx->u_hash, (nu + 1) * 16));
    CKS(ensure(ctx, ctx->u_rev, nu + 1));
    CKS(ensure(ctx, ctx->u_circ, nu + 1));
    const uint64_t n_windows = total - nu * (uint64_t)(ctx->t_k - 1);       // k-min-mers over all unitigs
    CKS(ensure(ctx, ctx->u_abund, (n_windows + 1) * 4));
    launch_unitig_scatter(ua, pair, ctx->u_flag.as<uint32_t>(), ctx->u_seqoff.as<uint64_t>(), ctx->u_idx.as<uint64_t>(),
                          ctx->u_mins.as<uint32_t>(), ctx->u_off.as<uint64_t>(), ctx->u_circ.as<uint8_t>(), ctx->u_cychead.as<uint8_t>(),
                          ctx->u_abund.as<uint32_t>(), s);
    CK(cudaMemcpyAsync(ctx->u_off.as<uint64_t>() + nu, &ctx->h_scalar[2], 8, cudaMemcpyHostToDevice, s));
    launch_unitig_hash(ctx->u_mins.as<uint32_t>(), ctx->u_off.as<uint64_t>(), nu, ctx->u_hash.as<uint64_t>(), ctx->u_rev.as<uint8_t>(), s);
    launch_unitig_reverse(ctx->u_mins.as<uint32_t>(), ctx->u_off.as<uint64_t>(), nu, ctx->u_rev.as<uint8_t>(), ctx->u_abund.as<uint32_t>(),
                          ctx->t_k, s);
    CKS(check_launch(ctx, "unitig_scatter_kernel + unitig_hash_kernel + unitig_reverse_kernel", nu ? 3 : 0));
    // computeDeterministicUnitigs' order (ascending u128 hash; unitigIndex = 2 * position) and the two checksums the
    // reference logs for the stage, both on the device
    uint32_t bucket_bits = 10;
    while ((1ull << bucket_bits) < nu / 2 && bucket_bits < 28) bucket_bits++;
    const uint64_t n_buckets = 1ull << bucket_bits;
    CKS(ensure(ctx, ctx->u_bcnt, n_buckets * 4));
    CKS(ensure(ctx, ctx->u_boff, (n_buckets + 1) * 8));
    CKS(ensure(ctx, ctx->u_order, (nu + 1) * 4));
    CKS(ensure(ctx, ctx->u_pos, (nu + 1) * 4));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)n_buckets) * sizeof(uint64_t)));
    launch_unitig_sort(ctx->u_hash.as<uint64_t>(), nu, bucket_bits, ctx->u_bcnt.as<uint32_t>(), ctx->u_boff.as<uint64_t>(),
                       ctx->scan_scratch.as<uint64_t>(), ctx->u_order.as<uint32_t>(), ctx->u_pos.as<uint32_t>(), s);
    CK(cudaMemsetAsync(&ctx->d_small->n_flagged, 0, 2 * sizeof(unsigned long long), s));       // n_flagged, n_changed: the two sums
    launch_unitig_checksum(ctx->u_mins.as<uint32_t>(), ctx->u_off.as<uint64_t>(), ctx->u_abund.as<uint32_t>(), ctx->u_pos.as<uint32_t>(),
                           nu, ctx->t_k, &ctx->d_small->n_flagged, s);
    CKS(check_launch(ctx, "unitig_bucket_count/fill/sort kernels + scan + unitig_checksum_kernel", nu ? 7 : 0));
    // unitig graph edges (indexUnitigEdges + computeUnitigEdges): end-node lists per edge-set slot, then the successor /
    // predecessor lists of every unitig, in the order a single reference thread would write them
    uint64_t n_uedges = 0;
    const bool with_edges = nu > 0 && ctx->t_k <= 64;
    if (with_edges) {
        const uint64_t ecap = ctx->edge_cap;
        CKS(ensure(ctx, ctx->u_scnt, ecap * 4));
        CKS(ensure(ctx, ctx->u_soff, (ecap + 1) * 8));
        CKS(ensure(ctx, ctx->u_ents, (4 * nu + 1) * 8));
        CKS(ensure(ctx, ctx->u_ecnt, (2 * nu + 1) * 4));
        CKS(ensure(ctx, ctx->u_eoff, (2 * nu + 2) * 8));
        CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)std::max<uint64_t>(ecap, 2 * nu)) * sizeof(uint64_t)));
        UnitigEdgeArgs ea{};
        ea.mins = ctx->u_mins.as<uint32_t>(); ea.off = ctx->u_off.as<uint64_t>(); ea.n_unitigs = nu; ea.k = ctx->t_k;
        ea.pos_of = ctx->u_pos.as<uint32_t>(); ea.order = ctx->u_order.as<uint32_t>();
        ea.edges = ctx->edge_table.as<Slot>(); ea.edge_mask = ecap - 1;
        ea.slot_cnt = ctx->u_scnt.as<uint32_t>(); ea.slot_off = ctx->u_soff.as<uint64_t>();
        ea.entries = ctx->u_ents.as<unsigned long long>();
        ea.edge_cnt = ctx->u_ecnt.as<uint32_t>(); ea.edge_off = ctx->u_eoff.as<uint64_t>();
        ea.checksum = &ctx->d_small->emit_cursor; ea.error_flag = &ctx->d_small->full_flag;
        CK(cudaMemsetAsync(ctx->u_scnt.p, 0, ecap * 4, s));
        launch_unitig_end_offers(ea, 1, s);
        launch_scan_u32_to_u64(ctx->u_scnt.as<uint32_t>(), ctx->u_soff.as<uint64_t>(), (uint32_t)ecap, ctx->scan_scratch.as<uint64_t>(), s);
        CK(cudaMemsetAsync(ctx->u_scnt.p, 0, ecap * 4, s));
        launch_unitig_end_offers(ea, 2, s);
        launch_unitig_end_sort(ctx->u_soff.as<uint64_t>(), ecap, ctx->u_ents.as<unsigned long long>(), s);
        launch_unitig_edges_query(ea, 1, s);
        launch_scan_u32_to_u64(ctx->u_ecnt.as<uint32_t>(), ctx->u_eoff.as<uint64_t>(), (uint32_t)(2 * nu), ctx->scan_scratch.as<uint64_t>(), s);
        CKS(check_launch(ctx, "unitig_end_offers/sort kernels + unitig_edges_query_kernel(count) + scans", 10));
        CK(cudaMemcpyAsync(&ctx->h_scalar[6], ctx->u_eoff.as<uint64_t>() + 2 * nu, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (ctx->h_small->full_flag) return fail(ctx, MDBG_ERR_STATE, "mdbg_unitigs_build: a unitig end's key is missing from the edge set");
        n_uedges = ctx->h_scalar[6];
        CKS(ensure(ctx, ctx->u_etgt, (n_uedges + 1) * 4));
        ea.edge_targets = ctx->u_etgt.as<uint32_t>();
        CK(cudaMemsetAsync(&ctx->d_small->emit_cursor, 0, 8, s));
        launch_unitig_edges_query(ea, 2, s);
        CKS(check_launch(ctx, "unitig_edges_query_kernel(fill)", 1));
        CKS(ensure_pin(ctx, ctx->hu_eoff, (2 * nu + 2) * 8));
        CKS(ensure_pin(ctx, ctx->hu_etgt, (n_uedges + 1) * 4));
        CK(cudaMemcpyAsync(ctx->hu_eoff.p, ctx->u_eoff.p, (2 * nu + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (n_uedges) CK(cudaMemcpyAsync(ctx->hu_etgt.p, ctx->u_etgt.p, n_uedges * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&ctx->h_scalar[7], &ctx->d_small->emit_cursor, 8, cudaMemcpyDeviceToHost, s));
        ctx->d2h_bytes += (2 * nu + 1) * 8 + n_uedges * 4 + 8;
    }
    CKS(ensure_pin(ctx, ctx->hu_mins, (total + 1) * 4));
    CKS(ensure_pin(ctx, ctx->hu_off, (nu + 2) * 8));
    CKS(ensure_pin(ctx, ctx->hu_hash, (nu + 1) * 16));
    CKS(ensure_pin(ctx, ctx->hu_circ, nu + 1));
    CKS(ensure_pin(ctx, ctx->hu_order, (nu + 1) * 4));
    CKS(ensure_pin(ctx, ctx->hu_abund, (n_windows + 1) * 4));
    CK(cudaMemcpyAsync(ctx->hu_off.p, ctx->u_off.p, (nu + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (nu) {
        CK(cudaMemcpyAsync(ctx->hu_hash.p, ctx->u_hash.p, nu * 16, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->hu_circ.p, ctx->u_circ.p, nu, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->hu_order.p, ctx->u_order.p, nu * 4, cudaMemcpyDeviceToHost, s));
    }
    if (total) CK(cudaMemcpyAsync(ctx->hu_mins.p, ctx->u_mins.p, total * 4, cudaMemcpyDeviceToHost, s));
    if (n_windows) CK(cudaMemcpyAsync(ctx->hu_abund.p, ctx->u_abund.p, n_windows * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[4], &ctx->d_small->n_flagged, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (!nu) ctx->hu_off.as<uint64_t>()[0] = 0;
    ctx->d2h_bytes += total * 4 + n_windows * 4 + nu * 29 + 24;
    uint32_t* order = ctx->hu_order.as<uint32_t>();
    const uint64_t cs_nodes = nu ? ctx->h_scalar[4] : 0, cs_ab = nu ? ctx->h_scalar[5] : 0;
    uint64_t n_circ = 0;
    for (uint64_t i = 0; i < nu; i++) n_circ += ctx->hu_circ.as<uint8_t>()[i];
    out->n_unitigs = nu;
    out->n_minimizers = total;
    out->n_circular = n_circ;
    out->offsets = ctx->hu_off.as<uint64_t>();
    out->minimizers = ctx->hu_mins.as<uint32_t>();
    out->hashes = ctx->hu_hash.as<uint64_t>();
    out->circular = ctx->hu_circ.as<uint8_t>();
    out->order = order;
    out->node_abundances = ctx->hu_abund.as<uint32_t>();
    if (with_edges) {
        out->n_unitig_edges = n_uedges;
        out->checksum_edges = ctx->h_scalar[7];
        out->edge_offsets = ctx->hu_eoff.as<uint64_t>();
        out->edge_targets = ctx->hu_etgt.as<uint32_t>();
    }
    out->checksum_nodes = cs_nodes;
    out->checksum_abundances = cs_ab;
    out->d_offsets = ctx->u_off.as<uint64_t>();
    out->d_minimizers = ctx->u_mins.as<uint32_t>();
    return MDBG_OK;
}

// Inverted index over the current table: for every emitted k-min-mer the (read, window) pairs of its occurrences in
// the stored reads (ReadCorrection::IndexReadsFunctor, src/readSelection/ReadCorrection.hpp:3064-3130).
mdbg_status mdbg_count_postings(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_postings_out* out) {
    if (!ctx || !out) return MDBG_ERR_ARG;
    memset(out, 0, sizeof *out);
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_postings before mdbg_count_begin");
    if (ctx->n_ranks > 1) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_postings is a single-context call (postings refer to the rank's own reads)");
    if (ctx->t_value_mode || !ctx->t_whole)
        return fail(ctx, MDBG_ERR_STATE, "mdbg_count_postings needs the occurrence-count table of the whole store (one mdbg_count_add_store over all reads)");
    if (ctx->t_capacity > 0xFFFFFFF0ull || ctx->s_reads > 0xFFFFFFF0ull) return fail(ctx, MDBG_ERR_ARG, "table too large for the postings scan");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t thr = count_threshold(ctx, min_abundance);
    const uint64_t cap = ctx->t_capacity;
    CKS(ensure_rem(ctx));
    CKS(ensure(ctx, ctx->pg_counts, cap * 4));
    CKS(ensure(ctx, ctx->pg_flags, cap * 4));
    CKS(ensure(ctx, ctx->pg_keyidx, (cap + 1) * 8));
    CKS(ensure(ctx, ctx->pg_off, (cap + 1) * 8));
    CKS(ensure(ctx, ctx->pg_readof, (ctx->s_mins + 1) * 4));
    CKS(ensure(ctx, ctx->scan_scratch, scan_scratch_elems((uint32_t)cap) * sizeof(uint64_t)));
    launch_posting_counts(ctx->table.as<Slot>(), cap, thr, ctx->pg_counts.as<uint32_t>(), ctx->pg_flags.as<uint32_t>(), s);
    launch_scan_u32_to_u64(ctx->pg_flags.as<uint32_t>(), ctx->pg_keyidx.as<uint64_t>(), (uint32_t)cap, ctx->scan_scratch.as<uint64_t>(), s);
    launch_scan_u32_to_u64(ctx->pg_counts.as<uint32_t>(), ctx->pg_off.as<uint64_t>(), (uint32_t)cap, ctx->scan_scratch.as<uint64_t>(), s);
    launch_read_of(ctx->s_off.as<uint64_t>(), ctx->s_reads, ctx->pg_readof.as<uint32_t>(), s);
    CKS(check_launch(ctx, "posting_counts_kernel + scans + read_of_kernel", 8));
    CK(cudaMemcpyAsync(&ctx->h_scalar[0], ctx->pg_keyidx.as<uint64_t>() + cap, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ctx->h_scalar[1], ctx->pg_off.as<uint64_t>() + cap, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint64_t n_keys = ctx->h_scalar[0], n_post = ctx->h_scalar[1];
    CKS(ensure(ctx, ctx->pg_hash, (n_keys + 1) * 16));
    CKS(ensure(ctx, ctx->pg_koff, (n_keys + 2) * 8));
    CKS(ensure(ctx, ctx->pg_reads, (n_post + 1) * 4));
    CKS(ensure(ctx, ctx->pg_wins, (n_post + 1) * 4));
    launch_posting_keys(ctx->table.as<Slot>(), cap, ctx->pg_flags.as<uint32_t>(), ctx->pg_keyidx.as<uint64_t>(),
                        ctx->pg_off.as<uint64_t>(), ctx->pg_hash.as<uint64_t>(), ctx->pg_koff.as<uint64_t>(), s);
    CK(cudaMemsetAsync(ctx->pg_counts.p, 0, cap * 4, s));               // reused as the per-list fill cursors
    PostingArgs pa{};
    pa.mins = ctx->s_min.as<uint32_t>(); pa.rem = ctx->s_rem.as<uint8_t>(); pa.offs = ctx->s_off.as<uint64_t>();
    pa.read_of = ctx->pg_readof.as<uint32_t>();
    pa.g_lo = 0; pa.g_hi = ctx->s_mins; pa.k = ctx->t_k;
    pa.table = ctx->table.as<Slot>(); pa.mask = cap - 1;
    pa.flags = ctx->pg_flags.as<uint32_t>(); pa.post_off = ctx->pg_off.as<uint64_t>(); pa.cursors = ctx->pg_counts.as<uint32_t>();
    pa.out_reads = ctx->pg_reads.as<uint32_t>(); pa.out_windows = ctx->pg_wins.as<uint32_t>();
    launch_posting_fill(pa, s);
    CKS(check_launch(ctx, "posting_keys_kernel + posting_fill_kernel", 2));
    CKS(ensure_pin(ctx, ctx->hpg_hash, (n_keys + 1) * 16));
    CKS(ensure_pin(ctx, ctx->hpg_koff, (n_keys + 2) * 8));
    CKS(ensure_pin(ctx, ctx->hpg_reads, (n_post + 1) * 4));
    CKS(ensure_pin(ctx, ctx->hpg_wins, (n_post + 1) * 4));
    if (n_keys) CK(cudaMemcpyAsync(ctx->hpg_hash.p, ctx->pg_hash.p, n_keys * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->hpg_koff.p, ctx->pg_koff.p, (n_keys + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (n_post) {
        CK(cudaMemcpyAsync(ctx->hpg_reads.p, ctx->pg_reads.p, n_post * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->hpg_wins.p, ctx->pg_wins.p, n_post * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    ctx->d2h_bytes += n_keys * 24 + n_post * 8;
    out->k = ctx->t_k;
    out->n_keys = n_keys;
    out->n_postings = n_post;
    out->hashes = ctx->hpg_hash.as<uint64_t>();
    out->offsets = ctx->hpg_koff.as<uint64_t>();
    out->reads = ctx->hpg_reads.as<uint32_t>();
    out->windows = ctx->hpg_wins.as<uint32_t>();
    out->d_hashes = ctx->pg_hash.as<uint64_t>();
    out->d_offsets = ctx->pg_koff.as<uint64_t>();
    out->d_reads = ctx->pg_reads.as<uint32_t>();
    out->d_windows = ctx->pg_wins.as<uint32_t>();
    return MDBG_OK;
}

// ---- multi-GPU --------------------------------------------------------------------------
mdbg_status mdbg_nccl_unique_id(uint8_t id_out[128]) {
    mdbg_ctx* ctx = nullptr;
    std::string err;
    if (!load_nccl(err)) return fail(nullptr, MDBG_ERR_NCCL, "%s", err.c_str());
    NK(g_nccl.GetUniqueId(id_out));
    return MDBG_OK;
}

mdbg_status mdbg_comm_init(mdbg_ctx* ctx, int rank, int n_ranks, const uint8_t id[128]) {
    if (!ctx || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return MDBG_ERR_ARG;
    if (n_ranks > 64) return fail(ctx, MDBG_ERR_ARG, "at most 64 ranks are supported");
    std::string err;
    if (!load_nccl(err)) return fail(ctx, MDBG_ERR_NCCL, "%s", err.c_str());
    CK(cudaSetDevice(ctx->device));
    Id128 uid;
    memcpy(uid.b, id, 128);
    NK(g_nccl.CommInitRank(&ctx->nccl_comm, n_ranks, uid, rank));
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    return MDBG_OK;
}

mdbg_status mdbg_count_merge(mdbg_ctx* ctx) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_merge before mdbg_count_begin");
    if (ctx->n_ranks == 1) return MDBG_OK;
    if (!ctx->nccl_comm) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_merge before mdbg_comm_init");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t R = (uint32_t)ctx->n_ranks, k = ctx->t_k;
    // layout of m_bucket (u64): [0,R) send counts | [R,2R) send bases | [2R, 2R+R*R) all ranks' send counts
    CKS(ensure(ctx, ctx->m_bucket, (size_t)(2 * R + R * R) * 8));
    uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
    uint64_t* d_base = d_cnt + R;
    uint64_t* d_all = d_cnt + 2 * R;
    CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
    PhaseClock clk(ctx);
    PackArgs p{};
    p.table = ctx->table.as<Slot>();
    p.capacity = ctx->t_capacity;
    p.k = k;
    p.n_ranks = R;
    p.mins = ctx->s_min.as<uint32_t>();
    p.foreign_vecs = ctx->foreign_vecs.as<uint32_t>();
    p.bucket_count = reinterpret_cast<unsigned long long*>(d_cnt);
    p.pass = 1;
    launch_table_pack(p, s);
    CKS(check_launch(ctx, "table_pack_kernel(count)", 1));
    clk.lap(PH_MERGE_PACK_COUNT);
    NK(g_nccl.AllGather(d_cnt, d_all, R, NCCL_UINT64, ctx->nccl_comm, s));
    std::vector<uint64_t> all((size_t)R * R);
    CK(cudaMemcpyAsync(all.data(), d_all, (size_t)R * R * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<uint64_t> send_cnt(R), send_base(R), recv_cnt(R), recv_base(R);
    uint64_t send_total = 0, recv_total = 0;
    for (uint32_t d = 0; d < R; d++) {
        send_cnt[d] = all[(size_t)ctx->rank * R + d];
        send_base[d] = send_total;
        send_total += send_cnt[d];
        recv_cnt[d] = all[(size_t)d * R + ctx->rank];
        recv_base[d] = recv_total;
        recv_total += recv_cnt[d];
    }
    CKS(ensure(ctx, ctx->m_send_vecs, (send_total + 1) * 4 * k));
    CKS(ensure(ctx, ctx->m_send_counts, (send_total + 1) * 4));
    CKS(ensure(ctx, ctx->m_recv_vecs, (recv_total + 1) * 4 * k));
    CKS(ensure(ctx, ctx->m_recv_counts, (recv_total + 1) * 4));
    clk.lap(PH_MERGE_PLAN);
    CK(cudaMemcpyAsync(d_base, send_base.data(), R * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
    p.bucket_base = d_base;
    p.out_vecs = ctx->m_send_vecs.as<uint32_t>();
    p.out_counts = ctx->m_send_counts.as<uint32_t>();
    p.pass = 2;
    launch_table_pack(p, s);
    CKS(check_launch(ctx, "table_pack_kernel(scatter)", 1));
    clk.lap(PH_MERGE_PACK_SCATTER);
    // one grouped all-to-all over NVLink: vectors and counts
    NK(g_nccl.GroupStart());
    for (uint32_t d = 0; d < R; d++) {
        if (send_cnt[d]) {
            NK(g_nccl.Send(ctx->m_send_vecs.as<uint32_t>() + send_base[d] * k, send_cnt[d] * 4 * k, NCCL_UINT8, (int)d,
                           ctx->nccl_comm, s));
            NK(g_nccl.Send(ctx->m_send_counts.as<uint32_t>() + send_base[d], send_cnt[d] * 4, NCCL_UINT8, (int)d,
                           ctx->nccl_comm, s));
        }
        if (recv_cnt[d]) {
            NK(g_nccl.Recv(ctx->m_recv_vecs.as<uint32_t>() + recv_base[d] * k, recv_cnt[d] * 4 * k, NCCL_UINT8, (int)d,
                           ctx->nccl_comm, s));
            NK(g_nccl.Recv(ctx->m_recv_counts.as<uint32_t>() + recv_base[d], recv_cnt[d] * 4, NCCL_UINT8, (int)d,
                           ctx->nccl_comm, s));
        }
    }
    NK(g_nccl.GroupEnd());
    clk.lap(PH_MERGE_EXCHANGE);
    // rebuild the table with only the keys this rank owns (recv_total bounds its distinct keys)
    const uint64_t cap = table_capacity_for(recv_total);
    std::swap(ctx->foreign_vecs, ctx->m_recv_vecs);       // received vectors become the table's vector store
    ctx->foreign_n = recv_total;
    CKS(ensure_table_buf(ctx, ctx->table, cap * sizeof(Slot)));
    CK(cudaMemsetAsync(ctx->table.p, 0, cap * sizeof(Slot), s));
    CKS(reset_claims(ctx));
    ctx->t_capacity = cap;
    InsertVecArgs iv{};
    iv.vecs = ctx->foreign_vecs.as<uint32_t>();
    iv.counts = ctx->m_recv_counts.as<uint32_t>();
    iv.n = recv_total;
    iv.foreign_base = 0;
    iv.k = k;
    iv.table = ctx->table.as<Slot>();
    iv.mask = cap - 1;
    iv.full_flag = &ctx->d_small->full_flag;
    iv.assign = ctx->t_value_mode ? 1u : 0u;      // next-k abundances are values (equal on every rank), not counts to add
    launch_insert_vecs(iv, s);
    CKS(check_launch(ctx, "insert_vecs_kernel", recv_total ? 1 : 0));
    CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_small->full_flag) return fail(ctx, MDBG_ERR_TABLE_FULL, "merged table full");
    clk.lap(PH_MERGE_INSERT);
    ctx->t_merged = true;
    ctx->t_rebuildable = false;            // the table now holds foreign vectors: it cannot be rebuilt from the store
    ctx->t_ranges.clear();
    ctx->t_claim_limit = cap;
    return MDBG_OK;
}

// Keys-only owner merge: (hash128, abundance) records of 24 bytes instead of (k-min-mer vector, abundance).  Afterwards
// rank r holds exactly the keys it owns with their global abundances, as after mdbg_count_merge, but the slots carry no
// vectors: finalize returns kminmers = NULL and mdbg_edges_index / mdbg_count_rescue are refused.  Made for the per-k
// tables of a multi-k loop, where the next pass needs no table from other ranks at all (next_k_stream_kernel) and the
// merge only has to put every key on one rank: 4.4 x fewer bytes over NVLink at k = 21, no vector gather, no re-hash.
mdbg_status mdbg_count_merge_hashes(mdbg_ctx* ctx) {
    if (!ctx) return MDBG_ERR_ARG;
    if (!ctx->t_active) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_merge_hashes before mdbg_count_begin");
    if (ctx->n_ranks == 1) return MDBG_OK;
    if (!ctx->nccl_comm) return fail(ctx, MDBG_ERR_STATE, "mdbg_count_merge_hashes before mdbg_comm_init");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t R = (uint32_t)ctx->n_ranks;
    CKS(ensure(ctx, ctx->m_bucket, (size_t)(2 * R + R * R) * 8));
    uint64_t* d_cnt = ctx->m_bucket.as<uint64_t>();
    CK(cudaMemsetAsync(d_cnt, 0, R * 8, s));
    PhaseClock clk(ctx);
    // ONE pack pass: a region of the send buffer per destination, each large enough for every key of the table (the
    // number of distinct keys is known from the pass's claim counter; a table with foreign entries is scanned once)
    uint64_t region_cap = ctx->h_small->t_claims;
    if (!ctx->t_rebuildable || region_cap == 0 || region_cap > ctx->t_capacity) {
        TableStats st;
        CKS(table_stats(ctx, 2, &st));
        region_cap = st.n_distinct;
    }
    region_cap += 1;
    {   // the tables of a multi-k loop grow a little with every k: allocate with headroom instead of re-allocating per k
        const size_t need = (size_t)R * region_cap * 24;
        if (need > ctx->m_send_vecs.cap) CKS(ensure(ctx, ctx->m_send_vecs, need + need / 2));
    }
    PackArgs p{};
    p.table = ctx->table.as<Slot>();
    p.capacity = ctx->t_capacity;
    p.k = ctx->t_k;
    p.n_ranks = R;
    p.bucket_count = reinterpret_cast<unsigned long long*>(d_cnt);
    launch_table_pack_hashes(p, ctx->m_send_vecs.as<uint64_t>(), region_cap, s);
    CKS(check_launch(ctx, "table_pack_hashes_kernel", 1));
    clk.lap(PH_MERGE_PACK_SCATTER);
    OwnerExchange x;
    CKS(x.plan(ctx));                                               // all-gather of the R x R counts + one D2H
    for (uint32_t d = 0; d < R; d++) {
        if (x.send_cnt[d] > region_cap) return fail(ctx, MDBG_ERR_STATE, "keys-only merge: send region overflow");
        x.send_base[d] = (uint64_t)d * region_cap;                  // regions, not a packed prefix
    }
    if ((x.recv_total + 1) * 24 > ctx->m_recv_vecs.cap) CKS(ensure(ctx, ctx->m_recv_vecs, (x.recv_total + 1) * 36));
    clk.lap(PH_MERGE_PLAN);
    CKS(x.run(ctx, ctx->m_send_vecs.p, ctx->m_recv_vecs.p, 24));
    clk.lap(PH_MERGE_EXCHANGE);
    const uint64_t cap = table_capacity_for(x.recv_total);
    CKS(ensure_table_buf(ctx, ctx->table, cap * sizeof(Slot)));
    CK(cudaMemsetAsync(ctx->table.p, 0, cap * sizeof(Slot), s));
    CKS(reset_claims(ctx));
    ctx->t_capacity = cap;
    launch_insert_hash_recs(ctx->m_recv_vecs.as<uint64_t>(), x.recv_total, ctx->table.as<Slot>(), cap - 1,
                            ctx->t_value_mode ? 1u : 0u, &ctx->d_small->full_flag, s);
    CKS(check_launch(ctx, "insert_hash_recs_kernel", x.recv_total ? 1 : 0));
    CK(cudaMemcpyAsync(&ctx->h_small->full_flag, &ctx->d_small->full_flag, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_small->full_flag) return fail(ctx, MDBG_ERR_TABLE_FULL, "merged table full");
    clk.lap(PH_MERGE_INSERT);
    ctx->t_merged = true;
    ctx->t_no_vecs = true;
    ctx->t_rebuildable = false;
    ctx->t_ranges.clear();
    ctx->t_claim_limit = cap;
    ctx->foreign_n = 0;
    return MDBG_OK;
}

// The multi-k loop as ONE library call (the host-side control of metaMDBG's `graph` passes, src/graph/CreateMdbg.cpp:386-468
// driven per k by src/pipeline/AssemblyPipeline.hpp:606-671, without the contig stage in between): count at first_k
// [+ merge, + rescue], then for every further k the previous-k table from the current one, the next-k pass over the
// store and [the merge]; statistics of every k are returned, tables stay on the device (the last one is current).
mdbg_status mdbg_multi_k_run(mdbg_ctx* ctx, uint32_t first_k, uint32_t last_k, uint32_t min_abundance, int rescue,
                             int merge_mode, mdbg_k_stats* stats_out) {
    if (!ctx || !stats_out) return MDBG_ERR_ARG;
    if (first_k < 2 || last_k < first_k || last_k > 255) return fail(ctx, MDBG_ERR_ARG, "need 2 <= first_k <= last_k <= 255");
    if (merge_mode < 0 || merge_mode > 2) return fail(ctx, MDBG_ERR_ARG, "merge_mode: 0 none, 1 with vectors, 2 keys only for k > first_k");
    const int R = ctx->n_ranks;
    uint64_t n_keys = 0;
    for (uint32_t k = first_k; k <= last_k; k++) {
        mdbg_k_stats& st = stats_out[k - first_k];
        memset(&st, 0, sizeof st);
        st.k = k;
        if (k == first_k) {
            CKS(mdbg_count_begin(ctx, k, 0));
            CKS(mdbg_count_add_store(ctx, 0, UINT64_MAX));
            if (merge_mode && R > 1) CKS(mdbg_count_merge(ctx));
            if (rescue) CKS(mdbg_count_rescue(ctx, &st.n_reads_rescued));
        } else {
            CKS(mdbg_prev_from_current(ctx, min_abundance));
            const uint64_t expect = (uint64_t)(1.3 * (double)std::max<uint64_t>(256, n_keys));
            CKS(mdbg_count_begin(ctx, k, expect));
            CKS(mdbg_count_add_store_next_k(ctx, 0, UINT64_MAX));
            if (merge_mode == 1 && R > 1) CKS(mdbg_count_merge(ctx));
            if (merge_mode == 2 && R > 1) CKS(mdbg_count_merge_hashes(ctx));
        }
        CKS(mdbg_count_stats(ctx, min_abundance, &st.n_entries, &st.n_distinct, &st.n_instances, &st.checksum));
        n_keys = st.n_entries * (uint64_t)((merge_mode && R > 1) ? R : 1);       // keys a rank-local table of the next k will hold
    }
    return MDBG_OK;
}

// ---- synthetic reads -------------------------------------------------------------------
mdbg_status mdbg_synth_fill_reads(mdbg_ctx* ctx, uint8_t* d_bases, const uint64_t* d_offsets, const uint64_t* d_vstart,
                                  const uint8_t* d_strand, uint32_t n_reads, uint64_t read_index_base, uint64_t seed,
                                  uint32_t err_q24) {
    if (!ctx) return MDBG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    launch_synth_fill(d_bases, d_offsets, d_vstart, d_strand, n_reads, read_index_base, seed, err_q24, ctx->stream);
    CKS(check_launch(ctx, "synth_fill_kernel", n_reads ? 1 : 0));
    return MDBG_OK;
}

}  // extern "C"
