// engine.cuh -- kernel launchers shared between the .cu files of libmdbg_b200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mdbg {

// ------------------------------------------------------------------ sketch
struct SketchArgs {
    const uint8_t* bases;        // concatenated ASCII reads (16-byte aligned allocation)
    const uint8_t* bases_end;    // bases + n_bases: bytes at/after this address are never dereferenced
    const uint64_t* offsets;     // [n_reads+1]
    // optional per-read source table of a host batch that travelled 2-bit packed (bit 63 clear: index of the
    // read's first u32 in `packed`, 16 bases per word, base j at bits [2j,2j+1], code (c>>1)&3; bit 63 set: byte
    // offset in `bases` of a read holding a character outside "ACGT", kept as ASCII).  nullptr = plain ASCII batch.
    const uint64_t* read_src;
    const uint32_t* packed;
    uint32_t n_reads;
    uint32_t read_begin, read_end;   // this launch handles reads [read_begin, read_end) (H2D/compute pipelining)
    uint32_t l;                  // minimizer size (2..16)
    uint32_t hpc;                // homopolymer compression on/off
    uint64_t threshold;          // select iff murmur_h1 <= threshold ...
    uint32_t select_none;        // ... unless no hash value qualifies at all
    const uint32_t* blacklist;   // sorted repetitive minimizers (device) or nullptr
    uint32_t n_blacklist;
    // output slots: either exact (tight CSR offsets) or padded closed form
    const uint64_t* exact_off;   // [n_reads+1] or nullptr
    uint32_t cap_shift;          // padded: slot_lo(r) = (offsets[r] >> cap_shift) + r * cap_const
    uint32_t cap_const;
    uint32_t* out_min;
    uint32_t* out_pos;
    uint8_t* out_dir;
    uint32_t* n_min;             // [n_reads] true minimizer count per read
    uint32_t* cursor;            // dynamic read scheduler (zeroed before launch)
    unsigned long long* n_overflow;  // reads whose slot was too small (zeroed before launch)
    uint32_t variant;            // 0 / 1: arithmetic of the unrolled l = 15 block of the byte-ring kernel; 2: the packed
                                 // kernel (2-bit input by bulk copy, bit-packed ring) where it applies (sketch.cu)
    // variant 2 only: reads the packed kernel cannot take (kept as ASCII because they hold a byte outside "ACGT") are
    // appended here and sketched by the byte-ring kernel in list mode right behind it.  dirty_list has room for
    // read_end - read_begin indices; the two counters are zeroed before the launch.
    uint32_t* dirty_list;
    uint32_t* dirty_count;
    uint32_t* dirty_cursor;
    // list mode of the byte-ring kernel (set by launch_sketch): reads read_list[0 .. *read_list_n) instead of a range
    const uint32_t* read_list;
    const uint32_t* read_list_n;
};
constexpr int SKETCH_VARIANTS = 3;

// true when variant 2's packed kernel can run this configuration (l = 15, a usable 32-bit candidate threshold,
// something to select); everything else stays on the byte-ring kernel
bool sketch_packed_eligible(const SketchArgs& a);

// ASCII reads -> the 2-bit device layout of SketchArgs::packed (16 bases per u32, base j at bits [2j, 2j+1], code
// (c >> 1) & 3); read r starts at word pack_word_offset(offsets[r], r) (16-byte aligned, closed form, no prefix sum).
// read_src[r] = that word offset, or SRC_ASCII | offsets[r] for a read holding a byte outside "ACGT".
__host__ __device__ inline uint64_t pack_word_offset(uint64_t base_offset, uint64_t r) { return ((base_offset >> 6) + r) << 2; }
inline uint64_t pack_words_capacity(uint64_t n_bases, uint64_t n_reads) { return ((n_bases >> 6) + n_reads + 2) << 2; }
struct PackArgsAscii {
    const uint8_t* bases; const uint8_t* bases_end; const uint64_t* offsets;
    uint32_t read_begin, read_end;
    uint32_t* packed; uint64_t* read_src;
    // optional: read r's bytes start at bases[src_start[r]] instead of bases[offsets[r]] (reads located inside raw
    // FASTQ text, ingest.cu); offsets still give the lengths and the packed layout
    const uint64_t* src_start;
};
void launch_pack_ascii(const PackArgsAscii& a, int sm_count, cudaStream_t s);

// ------------------------------------------------------------------ FASTQ / FASTA text on the device (ingest.cu)
uint64_t newline_tiles(uint64_t n_bytes);                           // entries of the per-tile count array
void launch_newline_count(const uint8_t* text, uint64_t n, uint32_t* counts, cudaStream_t s);
void launch_newline_write(const uint8_t* text, uint64_t n, const uint64_t* tile_off, uint64_t* nl_pos, cudaStream_t s);
struct FastxArgs {
    const uint8_t* text; const uint64_t* nl; uint64_t n_records; uint32_t lines;   // lines per record: 4 (FASTQ) or 2 (FASTA)
    uint64_t* seq_start; uint32_t* seq_len; uint64_t* qual_start;                 // qual_start may be nullptr
    unsigned long long* n_bad;                                                    // records that are not well formed
};
void launch_fastx_records(const FastxArgs& a, cudaStream_t s);

int launch_sketch(const SketchArgs& a, int sm_count, cudaStream_t s);   // returns the number of kernels launched

// fills the shared memory of every SM with non-code bytes (verification aid of the autotuner, see sketch.cu)
void launch_smem_scramble(int sm_count, uint32_t seed, uint32_t* sink, cudaStream_t s);

// *n_diff += number of bytes at which a[0..n_bytes) and b[0..n_bytes) differ (autotune's identity check)
void launch_count_diff(const void* a, const void* b, size_t n_bytes, unsigned long long* n_diff, cudaStream_t s);

// counts (u32[n]) -> exclusive offsets (u64[n+1]); scratch must hold ceil(n/2048)+1 u64
void launch_scan_u32_to_u64(const uint32_t* counts, uint64_t* offsets, uint32_t n, uint64_t* scratch, cudaStream_t s);
size_t scan_scratch_elems(uint32_t n);

struct CompactArgs {
    const uint64_t* base_offsets;  // read byte offsets (for the padded slot formula)
    uint32_t cap_shift, cap_const;
    const uint32_t* n_min;
    const uint64_t* tight_off;     // [n_reads+1]
    const uint32_t* in_min; const uint32_t* in_pos; const uint8_t* in_dir;
    uint32_t* out_min; uint32_t* out_pos; uint8_t* out_dir;
    uint32_t n_reads;
    const uint8_t* in_qual; uint8_t* out_qual;   // optional per-minimizer qualities (nullptr = none)
    // piece-wise compaction of a host batch (api.cu, PiecePipeline): reads [read_begin, read_begin + n_reads) only,
    // tight_off = the piece's own exclusive offsets (indexed from 0), destination shifted by tight_base
    uint32_t read_begin;
    uint64_t tight_base;
};
void launch_compact(const CompactArgs& a, cudaStream_t s);

// store_off[dst_read_base + 1 + i] = dst_min_base + batch_off[i + 1]
void launch_append_offsets(const uint64_t* batch_off, uint64_t* store_off, uint32_t n_reads, uint64_t dst_read_base,
                           uint64_t dst_min_base, cudaStream_t s);

// ------------------------------------------------------------------ read side outputs (row A3b)
constexpr int ERR_SHIFT = 60;   // error rates are handed over as exact integers: float value * 2^60
struct AuxArgs {
    const uint8_t* bases; const uint8_t* bases_end; const uint8_t* quals;   // quals may be nullptr
    const uint64_t* offsets; uint32_t n_reads;
    uint32_t l, hpc;
    const uint64_t* err_fixed;   // [256]
    const uint8_t* err_tz;       // [256] index of the lowest set bit of err_fixed (255 for zero)
    const uint64_t* exact_off;   // tight slots (after an overflow re-sketch) or nullptr = padded formula
    uint32_t cap_shift, cap_const;
    uint32_t* n_min;             // in/out: zeroed for filtered reads
    const uint32_t* pad_pos;
    uint32_t* pad_raw_a; uint32_t* pad_raw_b;   // scratch, padded layout
    uint8_t* out_qual;           // padded layout, may be nullptr
    uint64_t* err_sum_lo; uint64_t* err_sum_hi; uint8_t* err_lmin;
    double* complexity; uint8_t* low_complexity;
    uint32_t filter_low_complexity;
};
void launch_read_aux(const AuxArgs& a, cudaStream_t s);

// ------------------------------------------------------------------ purge palindromes
// flags[r] = 1 when read r may contain a palindromic window (cheap necessary test)
void launch_purge_flag(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, uint32_t first_k, uint32_t last_k,
                       uint8_t* flags, unsigned long long* n_flagged, cudaStream_t s);
// exact Commons::purgePalindrome on flagged reads: keep[g] = 0 for banned minimizers, new_cnt[r] = survivors
void launch_purge_exact(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, const uint8_t* flags,
                        uint32_t first_k, uint32_t last_k, uint8_t* keep, uint32_t* new_cnt,
                        unsigned long long* n_changed, cudaStream_t s);
void launch_density_filter(const uint32_t* mins, const uint64_t* offs, uint64_t n_reads, uint64_t threshold,
                           uint32_t select_none, uint8_t* keep, uint32_t* new_cnt, unsigned long long* n_changed,
                           cudaStream_t s);
// out_rem (optional): rem[] of the compacted store (launch_fill_rem's output) as a by-product
void launch_purge_compact(const uint32_t* mins, const uint64_t* offs, const uint64_t* new_offs, const uint8_t* keep,
                          uint64_t n_reads, uint32_t* out_mins, cudaStream_t s, uint8_t* out_rem = nullptr);

// ------------------------------------------------------------------ k-min-mer table
struct __align__(32) Slot {
    uint64_t lo;      // Murmur h2 = low 64 bits of KmerVec::hash128 (0,0 = empty)
    uint64_t hi;      // Murmur h1 = high 64 bits
    uint32_t count;   // abundance
    uint32_t flags;   // bit 0: rescued (abundance-1 k-min-mer kept by rescueKminmers)
    uint64_t ref;     // bit63: vector is the reversed window; bit62: index into foreign vecs; low bits: index
};

constexpr uint64_t SRC_ASCII = 1ULL << 63;
constexpr uint64_t REF_REV = 1ULL << 63;
constexpr uint64_t REF_FOREIGN = 1ULL << 62;
constexpr uint64_t REF_INDEX_MASK = (1ULL << 62) - 1;
constexpr uint64_t REF_NONE = ~0ULL;             // the slot's k-min-mer vector is not on this rank (keys-only merge)

void launch_fill_rem(const uint64_t* offs, uint64_t read_lo, uint64_t read_hi, uint8_t* rem, cudaStream_t s);

// Warp-form passes (kminmer.cu): 64 copies of the abandon flag and 64 shards of the claim counter, one 128-byte line each
struct PassAux {
    uint32_t flag[64][32];
    unsigned long long claims[64][16];
};

struct InsertArgs {
    const uint32_t* mins;   // store minimizers
    const uint8_t* rem;     // min(#minimizers from g to end of its read, 255)
    uint64_t g_lo, g_hi;    // flat minimizer range
    uint32_t k;
    Slot* table;
    uint64_t mask;          // capacity - 1 (any capacity: slot_of() in table.cuh scales the hash to it)
    uint32_t* full_flag;    // raised when a probe sequence ran out or the table passed claim_limit distinct keys
    unsigned long long* claims;        // running number of claimed slots (= distinct keys) of the table
    unsigned long long claim_limit;
    PassAux* aux;           // warp-form bookkeeping (nullptr: block form), zeroed with the table
    uint32_t aux_shards;    // claim-counter shards in use: a power of two <= 64
};
void launch_insert(const InsertArgs& a, cudaStream_t s);
int pass_kernels_per_launch(bool have_aux);   // 1, or 2 when the warp form adds pass_fold_kernel

// insert pre-normalized foreign vectors with counts (multi-GPU merge, receive side)
struct InsertVecArgs {
    const uint32_t* vecs;   // [n*k]
    const uint32_t* counts; // [n]
    uint64_t n;
    uint64_t foreign_base;  // index of vecs[0] in the context's foreign vector array
    uint32_t k;
    Slot* table;
    uint64_t mask;
    uint32_t* full_flag;
    uint32_t assign;        // 0: counts are added (occurrence counts); 1: insert-if-absent with the value (next-k tables)
};
void launch_insert_vecs(const InsertVecArgs& a, cudaStream_t s);

constexpr uint32_t SLOT_RESCUED = 1u;

struct TableStats {
    unsigned long long n_entries;    // count >= threshold, or rescued
    unsigned long long n_distinct;   // occupied
    unsigned long long n_instances;  // sum of counts
    unsigned long long checksum;     // sum count*lo over qualifying entries
    unsigned long long n_rescued;    // rescued entries among n_entries
};
void launch_table_stats(const Slot* table, uint64_t capacity, uint32_t min_count, TableStats* d_stats, cudaStream_t s);

struct EmitArgs {
    const Slot* table;
    uint64_t capacity;
    uint32_t min_count;
    uint32_t k;
    const uint32_t* mins;          // store minimizers
    const uint32_t* foreign_vecs;  // merged-in vectors (may be nullptr)
    uint64_t* out_hashes;          // [2n]: lo, hi
    uint32_t* out_abund;           // [n]
    uint32_t* out_vecs;            // [n*k], or nullptr when only (hash, abundance) pairs are wanted
    unsigned long long* cursor;    // zeroed before launch
};
void launch_table_emit(const EmitArgs& a, cudaStream_t s);

// ---- rescue (RescueKminmerFunctor, CreateMdbg.hpp:4579-4637): flag the abundance-1 k-min-mers of reads whose
// median solid abundance is <= 10
struct RescueArgs {
    const uint32_t* mins;
    const uint64_t* offs;
    uint64_t n_reads;
    uint32_t k;
    Slot* table;                 // one context: the count table; multi-rank: the replicated table of solid k-min-mers
    uint64_t mask;
    unsigned long long* n_reads_rescued;
    // multi-rank only (nullptr otherwise): normalized vectors of the rescued, non-solid windows, appended through
    // *out_cursor (zeroed before launch; capacity = number of windows of the local reads)
    uint32_t* out_vecs;
    unsigned long long* out_cursor;
};
void launch_rescue(const RescueArgs& a, cudaStream_t s);

struct BucketVecArgs {
    const uint32_t* vecs; uint64_t n; uint32_t k; uint32_t n_ranks;
    unsigned long long* bucket_count;   // [n_ranks], zeroed before each pass
    const uint64_t* bucket_base;        // [n_ranks], pass 2
    uint32_t* out_vecs;                 // pass 2
    int pass;
};
void launch_bucket_vecs(const BucketVecArgs& a, cudaStream_t s);
// flags = SLOT_RESCUED on the slot of every vector (count < 2); *n_missing counts vectors without a slot
void launch_rescue_flag(const uint32_t* vecs, uint64_t n, uint32_t k, Slot* table, uint64_t mask,
                        unsigned long long* n_missing, cudaStream_t s);

// ---- previous-k lookup table + next-k pass (getRefinedAbundance / IndexKminmerFunctor)
struct PrevFromTableArgs {
    const Slot* table; uint64_t capacity; uint32_t min_count;
    Slot* prev; uint64_t prev_mask; uint32_t* full_flag;
};
void launch_prev_from_table(const PrevFromTableArgs& a, cudaStream_t s);
struct PrevLoadArgs {
    const uint64_t* hashes;   // [2n]: lo, hi
    const uint32_t* abund;    // [n]
    uint64_t n;
    Slot* prev; uint64_t prev_mask; uint32_t* full_flag;
};
void launch_prev_load(const PrevLoadArgs& a, cudaStream_t s);
struct NextKArgs {
    const uint32_t* mins; const uint8_t* rem; uint64_t g_lo, g_hi; uint32_t k;
    const Slot* prev; uint64_t prev_mask;
    uint32_t prev_min_count;           // previous-table entries below it (and not rescued) count as absent
    Slot* table; uint64_t mask; uint32_t* full_flag;
    unsigned long long* claims; unsigned long long claim_limit;
    // per-position values (kminmer.cu, next_k_stream_kernel): val_out[g] = what the next pass would look up for the
    // k-min-mer starting at g (nullptr = not wanted); val_in != nullptr selects the lookup-free form of the pass,
    // reading the previous pass's values instead of the previous-k table
    uint32_t* val_out; const uint32_t* val_in;
    PassAux* aux; uint32_t aux_shards;  // warp-form bookkeeping, as in InsertArgs
};
void launch_next_k(const NextKArgs& a, cudaStream_t s);

// ---- edge keys (CreateMdbg::EdgeIndexer): prefix / suffix hashes of every emitted node into a set table
struct EdgeArgs {
    const Slot* table; uint64_t capacity; uint32_t min_count; uint32_t k;
    const uint32_t* mins; const uint32_t* foreign_vecs;
    Slot* edges; uint64_t edge_mask; uint32_t* full_flag;
    unsigned long long* edge_vals;      // launch_edge_values only: [2 * edge capacity], zeroed
    const uint32_t* node_slot;          // optional: table slots of the emitted entries (launch_unitig_nodes); the insert /
    uint64_t n_nodes;                   // values kernels then visit these instead of scanning the whole table
};
void launch_edge_insert(const EdgeArgs& a, cudaStream_t s);
constexpr unsigned long long EDGE_VALID = 1ULL << 63, EDGE_MULTI = 1ULL << 34;
void launch_edge_values(const EdgeArgs& a, cudaStream_t s);
void launch_edge_emit(const Slot* edges, const unsigned long long* vals, uint64_t capacity, uint64_t* out_hashes,
                      unsigned long long* out_vals, unsigned long long* cursor, cudaStream_t s);
struct BucketKeyArgs {
    const uint64_t* keys; uint64_t n; uint32_t n_ranks;      // records: [rec_words * n], {lo, hi[, value word]}
    uint32_t rec_words;                                      // 2 or 3
    unsigned long long* bucket_count;                        // [n_ranks], zeroed before each pass
    const uint64_t* bucket_base;                             // [n_ranks], pass 2
    uint64_t* out_keys;                                      // pass 2
    int pass;
};
void launch_bucket_keys(const BucketKeyArgs& a, cudaStream_t s);
void launch_insert_keys(const uint64_t* keys, uint64_t n, Slot* table, uint64_t mask, uint32_t* full_flag, cudaStream_t s);
// multi-rank edge values: offers of the owned nodes as 3-word records, folded into the class words on the key's owner
void launch_edge_offers(const EdgeArgs& a, uint64_t* out_recs, unsigned long long* cursor, cudaStream_t s);
void launch_edge_apply_offers(const uint64_t* recs, uint64_t n, const Slot* edges, uint64_t mask, unsigned long long* vals,
                              uint32_t* full_flag, cudaStream_t s);

// ---- unitigs (unitig.cu): oriented nodes x = 2 * node id + reversed; links, list ranking, sequences
struct UnitigArgs {
    const Slot* table; uint64_t mask;
    const uint32_t* node_slot; const uint32_t* slot_node;      // node id <-> table slot
    uint32_t n_nodes; uint32_t k;
    const uint32_t* mins; const uint32_t* foreign_vecs;
    const Slot* edges; uint64_t edge_mask; const unsigned long long* edge_vals;
    uint32_t* next;                                              // [2 * n_nodes] linked successor or 0xFFFFFFFF
    uint32_t* error_flag;
};
void launch_unitig_nodes(const Slot* table, uint64_t capacity, uint32_t min_count, uint32_t* slot_node, uint32_t* node_slot,
                         unsigned long long* cursor, cudaStream_t s);
void launch_unitig_link(const UnitigArgs& a, cudaStream_t s);
void launch_unitig_rank_init(const uint32_t* next, uint32_t n2, unsigned long long* pair, uint32_t* len, cudaStream_t s);
void launch_unitig_jump(unsigned long long* pair, uint32_t n2, int steps, unsigned long long* n_open, cudaStream_t s);
void launch_unitig_cycle_list(const unsigned long long* pair, uint32_t n2, uint32_t* cyc_list, uint32_t* cyc_pos,
                              unsigned long long* cursor, cudaStream_t s);
void launch_unitig_cycle_init(const UnitigArgs& a, const uint32_t* cyc_list, const uint32_t* cyc_pos, uint32_t m, uint64_t* best,
                              uint32_t* jump, cudaStream_t s);
void launch_unitig_cycle_min(const uint64_t* best_in, const uint32_t* jump_in, uint32_t m, uint64_t* best_out, uint32_t* jump_out,
                             cudaStream_t s);
void launch_unitig_cycle_cut(const uint32_t* next, const uint32_t* cyc_list, const uint64_t* best, uint32_t m,
                             unsigned long long* pair, uint8_t* is_cycle_head, cudaStream_t s);
void launch_unitig_len(const unsigned long long* pair, uint32_t n2, uint32_t* len, cudaStream_t s);
void launch_unitig_select(const unsigned long long* pair, const uint8_t* is_cycle_head, const uint32_t* len, uint32_t n2, uint32_t k,
                          uint32_t* size, uint32_t* flag, cudaStream_t s);
void launch_unitig_scatter(const UnitigArgs& a, const unsigned long long* pair, const uint32_t* flag, const uint64_t* seq_off,
                           const uint64_t* unitig_idx, uint32_t* out_mins, uint64_t* out_off, uint8_t* out_circular,
                           const uint8_t* is_cycle_head, uint32_t* out_abund, cudaStream_t s);
void launch_unitig_hash(const uint32_t* mins, const uint64_t* off, uint64_t n_unitigs, uint64_t* out_hashes, uint8_t* out_rev,
                        cudaStream_t s);
void launch_unitig_reverse(uint32_t* mins, const uint64_t* off, uint64_t n_unitigs, const uint8_t* rev, uint32_t* abund, uint32_t k,
                           cudaStream_t s);

// unitig graph edges: lists of unitig end nodes per edge-set slot, successor / predecessor lists per oriented unitig
struct UnitigEdgeArgs {
    const uint32_t* mins; const uint64_t* off; uint64_t n_unitigs; uint32_t k;      // unitig sequences (unsorted index j)
    const uint32_t* pos_of; const uint32_t* order;                                  // j -> position, position -> j
    const Slot* edges; uint64_t edge_mask;
    uint32_t* slot_cnt; const uint64_t* slot_off; unsigned long long* entries;      // per edge-set slot
    uint32_t* edge_cnt; const uint64_t* edge_off; uint32_t* edge_targets;           // per oriented unitig x = 2 * position + reversed
    unsigned long long* checksum; uint32_t* error_flag;
};
void launch_unitig_end_offers(const UnitigEdgeArgs& a, int pass, cudaStream_t s);   // pass 1: count, pass 2: fill (slot_cnt zeroed before each)
void launch_unitig_end_sort(const uint64_t* slot_off, uint64_t n_slots, unsigned long long* entries, cudaStream_t s);
void launch_unitig_edges_query(const UnitigEdgeArgs& a, int pass, cudaStream_t s);  // pass 1: edge_cnt, pass 2: edge_targets + checksum

// order[i] = unitig at position i of the ascending (high, low) hash order, pos_of = its inverse; cnt: 2^bucket_bits u32,
// bucket_off: 2^bucket_bits + 1 u64
void launch_unitig_sort(const uint64_t* hashes, uint64_t n, uint32_t bucket_bits, uint32_t* cnt, uint64_t* bucket_off,
                        uint64_t* scan_scratch, uint32_t* order, uint32_t* pos_of, cudaStream_t s);
void launch_unitig_checksum(const uint32_t* mins, const uint64_t* off, const uint32_t* abund, const uint32_t* pos_of, uint64_t n,
                            uint32_t k, unsigned long long* sums, cudaStream_t s);

// ---- postings (kminmer.cu): k-min-mer -> (read, window) lists over the count table
struct PostingArgs {
    const uint32_t* mins; const uint8_t* rem; const uint64_t* offs; const uint32_t* read_of;
    uint64_t g_lo, g_hi; uint32_t k;
    const Slot* table; uint64_t mask;
    const uint32_t* flags; const uint64_t* post_off; uint32_t* cursors;
    uint32_t* out_reads; uint32_t* out_windows;
};
void launch_posting_counts(const Slot* table, uint64_t capacity, uint32_t min_count, uint32_t* counts, uint32_t* flags, cudaStream_t s);
void launch_posting_keys(const Slot* table, uint64_t capacity, const uint32_t* flags, const uint64_t* key_index,
                         const uint64_t* post_off, uint64_t* out_hashes, uint64_t* out_offsets, cudaStream_t s);
void launch_read_of(const uint64_t* offs, uint64_t n_reads, uint32_t* read_of, cudaStream_t s);
void launch_posting_fill(const PostingArgs& a, cudaStream_t s);

// multi-GPU pack: bucket every occupied slot by owner rank
struct PackArgs {
    const Slot* table;
    uint64_t capacity;
    uint32_t k;
    uint32_t n_ranks;
    const uint32_t* mins;
    const uint32_t* foreign_vecs;
    unsigned long long* bucket_count;   // [n_ranks] (pass 1 output / pass 2 cursors)
    const uint64_t* bucket_base;        // [n_ranks] pass 2: start record of each bucket
    uint32_t* out_vecs;                 // pass 2: [total*k]
    uint32_t* out_counts;               // pass 2: [total]
    int pass;                           // 1 = count, 2 = scatter
};
void launch_table_pack(const PackArgs& a, cudaStream_t s);
// keys-only form, ONE pass: 24-byte records {hash lo, hash hi, count} appended to a fixed-capacity region per destination
void launch_table_pack_hashes(const PackArgs& a, uint64_t* out_recs, uint64_t region_cap, cudaStream_t s);
void launch_insert_hash_recs(const uint64_t* recs, uint64_t n, Slot* table, uint64_t mask, uint32_t assign, uint32_t* full_flag,
                             cudaStream_t s);

__host__ __device__ inline uint32_t owner_of(uint64_t hi, uint32_t n_ranks) {
    return (uint32_t)(((hi >> 32) * (uint64_t)n_ranks) >> 32);
}

// ------------------------------------------------------------------ repetitive minimizers (repeats.cu, K1b)
// table: capacity 64-bit words, all ones = empty, else (minimizer << 32) | count
void launch_mincount_insert(const uint32_t* mins, uint64_t n, unsigned long long* table, uint64_t mask,
                            unsigned long long* n_distinct, uint32_t* full_flag, cudaStream_t s);
void launch_mincount_hist(const unsigned long long* table, uint64_t capacity, unsigned long long* hist, uint32_t n_bins, cudaStream_t s);
void launch_mincount_emit(const unsigned long long* table, uint64_t capacity, uint32_t min_count, unsigned long long* out,
                          unsigned long long* cursor, uint64_t out_cap, cudaStream_t s);

// ------------------------------------------------------------------ synthetic reads
void launch_synth_fill(uint8_t* bases, const uint64_t* offsets, const uint64_t* vstart, const uint8_t* strand,
                       uint32_t n_reads, uint64_t read_index_base, uint64_t seed, uint32_t err_q24, cudaStream_t s);

}  // namespace mdbg
