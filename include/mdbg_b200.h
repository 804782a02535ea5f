/*
 * mdbg_b200.h -- C ABI of the Blackwell-native minimizer-sketch + k-min-mer
 * count engine (libmdbg_b200.so).
 *
 * This is the drop-in boundary for metaMDBG's hot path.  metaMDBG has no
 * plugin/FFI interface of its own (stages are functor callbacks inside one
 * binary), so every entry point below names the reference
 * code it replaces (paths relative to the metaMDBG source tree):
 *
 *   mdbg_sketch_batch         <- ReadSelectionFunctor::operator() lines that
 *                                chain EncoderRLE::execute and
 *                                MinimizerParser::parse
 *                                (src/readSelection/ReadSelection.hpp:682-690,
 *                                 src/Commons.hpp:4163-4203,
 *                                 src/utils/kmer/Kmer.hpp:1373-1456)
 *   mdbg_purge_palindromes    <- ReadSelection::purgePalindromes /
 *                                Commons::purgePalindrome
 *                                (src/readSelection/ReadSelection.hpp:1374-1431,
 *                                 src/Commons.hpp:1617-1723)
 *   mdbg_store_append         <- the read_data_*.txt minimizer-space records
 *                                KminmerParserParallel iterates
 *                                (src/Commons.hpp:7367-7495)
 *   mdbg_count_begin/add/finalize
 *                             <- CreateMdbg::KminmerCounter::execute
 *                                (src/graph/CreateMdbg.hpp:3591-3883) with
 *                                MDBG::getKminmers_complete (src/Commons.hpp:5282-5361)
 *                                and KmerVec::normalize/hash128 (src/Commons.hpp:886-969)
 *   mdbg_count_merge          <- the `hash128 % P` partitioning of
 *                                KminmerCounter::partitionKminmer
 *                                (src/graph/CreateMdbg.hpp:3714-3724), as an
 *                                owner-partitioned all-to-all between GPUs
 *
 * Conventions: plain C, no exceptions; every function returns an mdbg_status
 * (0 = OK) and records a message readable with mdbg_last_error(); all buffers
 * named "host" are caller-owned unless documented as library-owned; one
 * context per device; calls on one context must be serialised by the caller
 * (the host batches reads inside the reference's existing critical section);
 * different contexts are independent.  There is NO CPU fallback: creating a
 * context without a usable CUDA device fails with MDBG_ERR_CUDA.
 */
#ifndef MDBG_B200_H
#define MDBG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdbg_ctx mdbg_ctx;

typedef enum {
    MDBG_OK = 0,
    MDBG_ERR_CUDA = 1,        /* a CUDA call failed (message holds the CUDA error) */
    MDBG_ERR_ARG = 2,         /* invalid argument / unsupported parameter */
    MDBG_ERR_STATE = 3,       /* call sequence error (e.g. count_add before count_begin) */
    MDBG_ERR_TABLE_FULL = 4,  /* count table capacity exceeded: re-run with a larger expected_distinct */
    MDBG_ERR_NCCL = 5,        /* NCCL unavailable or a collective failed */
    MDBG_ERR_OOM = 6
} mdbg_status;

/* MinimizerParser(minimizerSize, density, repetitive set) -- Kmer.hpp:1351-1366;
 * use_hpc = Params::_useHomopolymerCompression (AssemblyPipeline.hpp:301,318). */
typedef struct {
    uint32_t minimizer_size;      /* l, 2..16 (AssemblyPipeline.hpp:201-202 caps it at 16) */
    float    density;             /* --density-assembly / --density-correction, as float */
    uint32_t use_hpc;             /* 1 = homopolymer-compress before sketching (HiFi) */
    const uint32_t* blacklist;    /* host: repetitiveMinimizers.bin values, may be NULL */
    uint64_t n_blacklist;
} mdbg_params;

/* Result of one sketch batch.  Host pointers are library-owned pinned memory,
 * valid until the next sketch call on the same context. */
typedef struct {
    uint32_t n_reads;
    uint64_t n_minimizers;
    const uint64_t* min_offsets;  /* [n_reads+1] CSR offsets into the three arrays below */
    const uint32_t* minimizers;   /* canonical l-mer value truncated to u32 (Kmer.hpp:1441) */
    const uint32_t* positions;    /* l-mer start in (HPC) sequence coordinates (Kmer.hpp:1442) */
    const uint8_t*  directions;   /* 0 = forward strand is canonical, 1 = reverse complement */
} mdbg_sketch_out;

/* Device-resident view of the same CSR (pointers into context-owned HBM, valid
 * until the next sketch call). */
typedef struct {
    uint32_t n_reads;
    uint64_t n_minimizers;
    const uint64_t* d_min_offsets;
    const uint32_t* d_minimizers;
    const uint32_t* d_positions;
    const uint8_t*  d_directions;
} mdbg_sketch_dev;

/* Finalised count table (host, library-owned pinned memory, valid until the
 * next finalize/destroy).  hashes holds the 16 on-disk bytes of each u128
 * exactly as MDBG::writeKminmerAbundance emits them (Commons.hpp:4463-4471):
 * hashes[2*i] = low 64 bits (Murmur h2), hashes[2*i+1] = high 64 bits (h1). */
typedef struct {
    uint32_t k;
    uint64_t n_entries;           /* entries with abundance >= max(2, min_abundance) */
    const uint64_t* hashes;       /* [2*n_entries] */
    const uint32_t* abundances;   /* [n_entries] */
    const uint32_t* kminmers;     /* [n_entries*k] normalized vectors (kminmerData_min.txt rows) */
    uint64_t n_instances;         /* k-min-mer occurrences inserted */
    uint64_t n_distinct;          /* distinct k-min-mers in the table (any abundance) */
    uint64_t checksum;            /* sum abundance * low64(hash) mod 2^64 (CreateMdbg.cpp:3321) */
    uint64_t n_rescued;           /* abundance-1 entries kept by mdbg_count_rescue (included in n_entries) */
} mdbg_table_out;

/* ---- context ------------------------------------------------------------ */
mdbg_status mdbg_ctx_create(int device, const mdbg_params* params, mdbg_ctx** out);
void        mdbg_ctx_destroy(mdbg_ctx* ctx);
const char* mdbg_last_error(mdbg_ctx* ctx);              /* ctx may be NULL: last create error */
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed
 * as void*); NULL restores the context's own stream. */
mdbg_status mdbg_ctx_set_stream(mdbg_ctx* ctx, void* cuda_stream);
mdbg_status mdbg_ctx_synchronize(mdbg_ctx* ctx);
/* Number of this library's kernels launched on the context so far. */
uint64_t    mdbg_ctx_kernel_launches(mdbg_ctx* ctx);
/* Bytes the host-buffer entry points (sketch batches, fetches, finalize) have sent over PCIe so far. */
mdbg_status mdbg_ctx_bytes_moved(mdbg_ctx* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Device buffers the context has (re)allocated so far (a steady-state loop must not add to it: cudaMalloc / cudaFree
 * synchronise the device and cost milliseconds at these sizes), and how often the count table and the previous-k table
 * traded buffers instead (see mdbg_prev_from_current). */
mdbg_status mdbg_ctx_allocations(mdbg_ctx* ctx, uint64_t* n_device_allocations, uint64_t* n_table_buffer_trades);
/* Per-kernel device timing (CUDA events recorded on the context's stream around
 * the launch).  which: 0 = sketch kernel (K1), 1 = k-min-mer insert kernel (K3).
 * Returns the duration of the most recent launch of that kernel in ms. */
mdbg_status mdbg_ctx_enable_timing(mdbg_ctx* ctx, int on);
mdbg_status mdbg_ctx_kernel_time_ms(mdbg_ctx* ctx, int which, float* ms);

/* Per-phase wall-clock profile of the table and collective paths (merge: pack / plan / exchange / insert; previous-k
 * replication; insert or next-k pass; statistics; emit; table reset).  While it is on, every phase boundary
 * synchronises the stream: the figures are exclusive phase times in ms, accumulated since the profile was switched
 * on, and the calls are slower -- a diagnostic, not for timed runs.  mdbg_ctx_phase_times returns the number of
 * phases written (names are static strings). */
mdbg_status mdbg_ctx_phase_profile(mdbg_ctx* ctx, int on);
int         mdbg_ctx_phase_times(mdbg_ctx* ctx, double* ms_out, const char** names_out, int max_n);

/* ---- sketch (rows A1-A3 of SURVEY.md section 8a) -------------------------- */
/* Host reads: read r = bases[offsets[r] .. offsets[r+1]) (ASCII, as Read::_seq).
 * The batch is copied to the device, sketched, the CSR is copied back into
 * `out`, and the minimizer-space reads are appended to the context's device
 * store when append_to_store != 0.  offsets[0] must be 0, offsets must not
 * decrease and a read must be shorter than 2^31 bases (MDBG_ERR_ARG otherwise;
 * positions are u32 as upstream).  The device-buffer variants trust their
 * offsets. */
mdbg_status mdbg_sketch_batch(mdbg_ctx* ctx, const uint8_t* bases, const uint64_t* offsets,
                              uint32_t n_reads, int append_to_store, mdbg_sketch_out* out);
/* A host batch is cut into pieces of >= 128 MB on read boundaries; packing / copying piece i+1, sketching piece i and
 * the scan, compaction and device-to-host copy of piece i-2's share of the CSR overlap, so `out` is complete shortly
 * after the last piece has been sketched.  `out` arrays live in pinned memory owned by the context and stay valid
 * until the next call on it.
 * Host batches without qualities cross PCIe 2-bit packed (worker threads + AVX-512 / AVX2 inside the library, unpacked again
 * by the sketch kernel; reads holding a byte outside "ACGT" stay ASCII, so results are identical).  on = 0 sends the
 * ASCII bytes as they are, on = 1 always packs, on = -1 (default) packs when the process has at least 12 usable CPUs
 * (cgroup quota and ranks-per-node aware), i.e. when packing outruns the PCIe transfer it saves.  on = 2
 * (experimental, pinned caller buffers only) additionally sends a piece as plain ASCII whenever the copy engine
 * is idle, so that DMA and the packer work side by side. */
mdbg_status mdbg_ctx_set_host_packing(mdbg_ctx* ctx, int on);
/* How the last host batch (mdbg_sketch_batch / _q) travelled: pieces it was cut into, pieces whose scan /
 * compaction / CSR copy ran piece-wise behind their sketch (all of them unless a read overflowed its padded slot,
 * which sends the whole batch through the exact re-sketch: overflow_fallback = 1), buffer growths with copies in
 * flight, and the host packer's throughput on the raw ASCII bytes. */
typedef struct mdbg_batch_info {
    uint64_t n_pieces, n_pieces_pipelined, n_buffer_growths, n_direct_pieces;
    int32_t overflow_fallback, packed;
    double pack_gb_per_s;
    const char* pack_isa;                            /* "avx512" | "avx2" | "scalar" */
    int32_t host_threads;
} mdbg_batch_info;
mdbg_status mdbg_ctx_last_batch_info(mdbg_ctx* ctx, mdbg_batch_info* info);
/* Host batch that is ALREADY 2-bit packed (a reader that packs while it parses -- on all of its parser threads --
 * hands over a quarter of the bytes and nothing is left for the library's packer to do):
 *   mdbg_host_pack_read    packs one read on the calling thread (thread safe, stateless; AVX-512 / AVX2 / scalar):
 *                          ceil(len / 16) words, base j of a word at bits [2j, 2j+1], code (c >> 1) & 3.  Returns 1, or
 *                          0 when the read holds a byte outside "ACGT" (words undefined): such a read must travel as
 *                          ASCII in the spill buffer.
 *   mdbg_sketch_batch_packed  packed     host words of the batch, n_words of them
 *                          read_src[r]   first word of read r (use multiples of 4: 16-byte aligned reads; must not
 *                                        decrease with r), or MDBG_SRC_ASCII | byte offset of the read in `ascii`
 *                          ascii         spill buffer of the reads that could not be packed (may be NULL)
 *                          offsets       [n_reads+1] base offsets (read lengths), offsets[0] = 0
 * Same pipelining, outputs and store behaviour as mdbg_sketch_batch. */
int         mdbg_host_pack_read(const uint8_t* bases, uint64_t len, uint32_t* words_out);
mdbg_status mdbg_sketch_batch_packed(mdbg_ctx* ctx, const uint32_t* packed, uint64_t n_words, const uint64_t* read_src,
                                     const uint8_t* ascii, uint64_t n_ascii_bytes, const uint64_t* offsets,
                                     uint32_t n_reads, int append_to_store, mdbg_sketch_out* out);
/* Raw, uncompressed FASTQ / FASTA TEXT -> sketch: the record split runs on the device (newline index, sequence /
 * quality line of every record, 2-bit packing straight from the text), replacing the parsing half of
 * ReadParserParallel::parse (src/Commons.hpp:5846-5911, kseq on one thread inside an omp critical).  Supported:
 * 4-line FASTQ and 2-line FASTA ('\n' or '\r\n'); any other shape returns MDBG_ERR_ARG and the caller uses its host
 * parser.  The block may end inside a record: only complete records are sketched and info->consumed_bytes says where
 * the next block must start (is_final != 0: a last line without a newline still ends its record).  Read r of `out`
 * is record r of the block. */
typedef struct {
    uint64_t n_records;        /* complete records sketched */
    uint64_t consumed_bytes;   /* text[0 .. consumed_bytes) was used */
    uint64_t n_bases;
    int32_t  format;           /* 1 = FASTQ, 2 = FASTA */
} mdbg_fastx_info;
mdbg_status mdbg_sketch_fastx(mdbg_ctx* ctx, const uint8_t* text, uint64_t n_bytes, int is_final, int append_to_store,
                              mdbg_sketch_out* out, mdbg_fastx_info* info);
/* Same with the reads already in HBM.  d_bases must be 16-byte aligned;
 * nothing is copied to the host.  `out` may be NULL. */
mdbg_status mdbg_sketch_batch_device(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets,
                                     uint32_t n_reads, uint64_t n_bases, int append_to_store,
                                     mdbg_sketch_dev* out);
/* The sketch kernel exists in MDBG_SKETCH_VARIANTS variants with identical results (sketch.cu):
 *   2 (default)  the packed kernel: 2-bit bases fetched by bulk copy (TMA) into shared memory, homopolymer
 *                compression by table lookups in registers, bit-packed ring, funnel-shift l-mers.  ASCII input is
 *                turned into the 2-bit layout by one streaming pass on the device first; reads holding a byte
 *                outside "ACGT", minimizer sizes other than 15 and degenerate densities use the byte-ring kernel.
 *   0 / 1        the byte-ring kernel on ASCII bytes (0 = one roll step + candidate filter per position, 1 =
 *                funnel-shift l-mers from codes packed in registers, one final hash multiply on the summed
 *                pre-images, carry-chain accept bits).
 * The environment variable MDBG_SKETCH_VARIANT presets the variant at context creation.  mdbg_ctx_autotune_sketch
 * runs every variant on the caller's own device-resident batch (nothing is appended to the store), compares the
 * complete results with variant 0 byte for byte ON THE DEVICE, and keeps the fastest variant that is identical.
 * Every variant is run twice: once as one launch (timed), once in launches of 2048 reads with the shared memory of
 * all SMs overwritten in between, so that a variant depending on stale shared memory fails the comparison. */
#define MDBG_SKETCH_VARIANTS 3
typedef struct mdbg_autotune_out {
    int32_t n_variants;
    int32_t chosen;                                  /* variant now active in the context */
    int32_t identical[4];                            /* [v]: 1 when variant v reproduced variant 0 exactly */
    float ms[4];                                     /* [v]: best-of-2 time of sketch + scan + compaction */
    uint32_t n_reads;
    uint64_t n_minimizers;
} mdbg_autotune_out;
mdbg_status mdbg_ctx_set_sketch_variant(mdbg_ctx* ctx, int variant);
mdbg_status mdbg_ctx_get_sketch_variant(mdbg_ctx* ctx, int* variant);
mdbg_status mdbg_ctx_autotune_sketch(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets,
                                     uint32_t n_reads, uint64_t n_bases, mdbg_autotune_out* out);
/* Reads resident in HBM in 2-bit packed form: 16 bases per u32, base j of a word at bits [2j, 2j+1], code
 * (c >> 1) & 3 (A=0 C=1 T=2 G=3); read r starts at word d_word_offsets[r] and has d_offsets[r+1] - d_offsets[r]
 * bases.  Only for reads made of the letters A, C, G, T (anything else must use the ASCII entry points). */
mdbg_status mdbg_sketch_batch_device_packed(mdbg_ctx* ctx, const uint32_t* d_packed, const uint64_t* d_word_offsets,
                                            const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases,
                                            int append_to_store, mdbg_sketch_dev* out);
/* The packed layout is the device-resident input format of the sketch (0.25 B per base).  d_packed must be
 * 16-byte aligned and readable up to the next multiple of 16 bytes behind the last word of the last read (any
 * cudaMalloc'ed buffer is): the kernel fetches each read with 16-byte granular bulk copies.  Reads that start on a
 * 16-byte boundary (word offset % 4 == 0, what mdbg_pack_device produces) avoid a slower first step.
 *
 * mdbg_pack_device: ASCII reads in HBM -> that layout, one streaming pass (1 B/bp read, 0.25 B/bp written).
 *   d_packed_out      caller-owned, mdbg_pack_device_words(n_bases, n_reads) u32
 *   d_read_src_out    [n_reads]: word offset of the read, or MDBG_SRC_ASCII | byte offset in d_bases for a read
 *                     holding a byte outside "ACGT" (not representable; it stays ASCII)
 * mdbg_sketch_batch_device_packed2 takes exactly these two arrays plus the ASCII buffer the flagged reads live in
 * (d_bases may be NULL when no entry is flagged). */
#define MDBG_SRC_ASCII (1ULL << 63)
uint64_t    mdbg_pack_device_words(uint64_t n_bases, uint64_t n_reads);
mdbg_status mdbg_pack_device(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                             uint64_t n_bases, uint32_t* d_packed_out, uint64_t* d_read_src_out);
mdbg_status mdbg_sketch_batch_device_packed2(mdbg_ctx* ctx, const uint32_t* d_packed, const uint64_t* d_read_src,
                                             const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                                             uint64_t n_bases, int append_to_store, mdbg_sketch_dev* out);
/* Copy the last batch's CSR to pinned host memory (after *_device). */
mdbg_status mdbg_sketch_fetch(mdbg_ctx* ctx, mdbg_sketch_out* out);

/* ---- read side outputs (row A3b: ReadSelection.hpp:870-920, 1047-1138, 1171-1228, 1302-1320) ------------
 * What ReadSelectionFunctor hands to writeRead besides the sketch.  Host pointers are library-owned. */
typedef struct {
    uint32_t n_reads;
    const float*   mean_quality;    /* [n_reads] meanReadQuality; NaN when the batch has no qualities */
    const double*  complexity;      /* [n_reads] computeSequenceComplexity(seq, 64, 32) */
    const uint8_t* low_complexity;  /* [n_reads] 1 = complexity > 5 (minimizers cleared when the filter is on) */
    const uint8_t* qualities;       /* [n_minimizers] per-minimizer min base quality, aligned with mdbg_sketch_out
                                       (all 1 when the batch has no qualities, ReadSelection.hpp:1049-1053) */
} mdbg_aux_out;

/* filter_low_complexity != 0: reads with complexity > 5 lose their minimizers (ReadSelection.hpp:894-903), as the
 * reference always does.  Only applied by mdbg_sketch_batch_q.  --min-read-quality > 0 is not supported. */
mdbg_status mdbg_ctx_set_read_filters(mdbg_ctx* ctx, int filter_low_complexity);
/* mdbg_sketch_batch plus the side outputs.  quals = Read::_qual bytes laid out like bases (NULL for FASTA).
 * With HPC on, a base string containing '#' (EncoderRLE's internal sentinel, never produced by a FASTA/FASTQ
 * parser) is refused with MDBG_ERR_ARG: the reference records shifted rlePositions after a '#', which the quality
 * windows here would not reproduce.  The plain sketch entry points accept such reads and stay bit-exact. */
mdbg_status mdbg_sketch_batch_q(mdbg_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                                uint32_t n_reads, int append_to_store, mdbg_sketch_out* out, mdbg_aux_out* aux);

/* ---- minimizer-space read store (device-resident read_data_*.txt) -------- */
mdbg_status mdbg_store_clear(mdbg_ctx* ctx);
/* Append host minimizer-space reads (what KminmerParserParallel would read). */
mdbg_status mdbg_store_append(mdbg_ctx* ctx, const uint32_t* minimizers, const uint64_t* min_offsets,
                              uint32_t n_reads);
mdbg_status mdbg_store_size(mdbg_ctx* ctx, uint64_t* n_reads, uint64_t* n_minimizers);
/* Copy the store back (min_offsets [n_reads+1], minimizers [n_minimizers]). */
mdbg_status mdbg_store_fetch(mdbg_ctx* ctx, uint64_t* min_offsets, uint32_t* minimizers);
/* Commons::purgePalindrome on every stored read, in place (row A4). */
mdbg_status mdbg_purge_palindromes(mdbg_ctx* ctx, uint32_t first_k, uint32_t last_k,
                                   uint64_t* n_reads_changed);

/* Utils::applyDensityThreshold (src/Commons.hpp:2507-2550) on every stored read, in place: keeps the minimizers
 * whose Murmur hash (of the stored u32 value) is below density * 2^64 -- how reads sketched at the correction
 * density (0.025) are parsed at the assembly density (0.005), src/Commons.hpp:7457. */
mdbg_status mdbg_store_apply_density(mdbg_ctx* ctx, float density, uint64_t* n_reads_changed);

/* The whole multi-k loop in one call: count at first_k [merge, rescue], then for k = first_k + 1 .. last_k the
 * previous-k table from the current one, the next-k pass over the store and [the merge].  merge_mode: 0 = none (one
 * context), 1 = mdbg_count_merge after every k, 2 = mdbg_count_merge at first_k and mdbg_count_merge_hashes after every
 * later k.  COLLECTIVE with several ranks.  stats_out has last_k - first_k + 1 entries; the table of last_k stays
 * current (finalize / edges as usual).  Equivalent to the call sequence shown in INTEGRATION.md. */
typedef struct {
    uint32_t k;
    uint64_t n_entries, n_distinct, n_instances, checksum, n_reads_rescued;
} mdbg_k_stats;
mdbg_status mdbg_multi_k_run(mdbg_ctx* ctx, uint32_t first_k, uint32_t last_k, uint32_t min_abundance, int rescue,
                             int merge_mode, mdbg_k_stats* stats_out);

/* ---- inverted index: k-min-mer -> (read, window) postings (row (f)4 of SURVEY.md section 8) -----------------------
 * What ReadCorrection::indexReads builds over the low-density reads for the ONT all-vs-all chaining
 * (IndexReadsFunctor, src/readSelection/ReadCorrection.hpp:3064-3130): for every emitted k-min-mer of the current
 * occurrence-count table, the list of (read index in the store, window index = positionIndex) of its occurrences.
 * CSR over the keys: list j is reads / windows[offsets[j] .. offsets[j+1]); its length is the k-min-mer's abundance;
 * the order inside a list is unspecified (upstream: arrival order of the OpenMP threads).  Host arrays are pinned and
 * library-owned; the d_* pointers are the same arrays in HBM.  Single context; the table must be the count of the
 * whole store (mdbg_count_begin + one mdbg_count_add_store). */
typedef struct {
    uint32_t k;
    uint64_t n_keys, n_postings;
    const uint64_t* hashes;     /* [2*n_keys]: low 64 bits, high 64 bits */
    const uint64_t* offsets;    /* [n_keys+1] */
    const uint32_t* reads;      /* [n_postings] */
    const uint32_t* windows;    /* [n_postings] */
    const uint64_t* d_hashes; const uint64_t* d_offsets; const uint32_t* d_reads; const uint32_t* d_windows;
} mdbg_postings_out;
mdbg_status mdbg_count_postings(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_postings_out* out);

/* ---- repetitive minimizers (ONT path; ReadSelection::determineRepetitiveMinimizers + CountMinimizerFunctor,
 * src/readSelection/ReadSelection.hpp:497-625) ------------------------------------------------------------
 * Counts every minimizer of the stored reads (upstream: the first 1 M reads sketched at the correction density with
 * an empty blacklist) and returns the max(1, int(fraction * #distinct)) most frequent values with their counts,
 * most frequent first -- the contents of repetitiveMinimizers.bin (fraction = 0.00001f upstream).  Upstream's choice
 * among values of EQUAL count at the cut-off depends on the iteration order of an unordered_map; here the smaller
 * value wins, and n_with_min_count / n_with_min_count_selected say whether the cut-off fell inside a tie.
 * mdbg_ctx_set_blacklist installs such a list as the blacklist of all later sketches (copied; NULL / 0 clears it). */
typedef struct {
    uint64_t n_distinct;                 /* distinct minimizer values in the store */
    uint64_t n_selected;
    const uint32_t* minimizers;          /* [n_selected] host, library-owned, valid until the next call */
    const uint32_t* counts;              /* [n_selected] */
    uint32_t min_count_selected;         /* count of the last selected value */
    uint64_t n_with_min_count;           /* values having exactly that count ... */
    uint64_t n_with_min_count_selected;  /* ... and how many of them were selected */
} mdbg_repeats_out;
mdbg_status mdbg_store_repetitive_minimizers(mdbg_ctx* ctx, float fraction, mdbg_repeats_out* out);
mdbg_status mdbg_ctx_set_blacklist(mdbg_ctx* ctx, const uint32_t* values, uint64_t n);

/* ---- k-min-mer count table (rows A5-A7) ----------------------------------- */
/* expected_distinct = 0 sizes the table from the store (upper bound: one slot
 * pair per k-min-mer occurrence). */
mdbg_status mdbg_count_begin(mdbg_ctx* ctx, uint32_t k, uint64_t expected_distinct);
/* Insert every k-min-mer of stored reads [read_lo, read_hi); (0, UINT64_MAX) = all. */
mdbg_status mdbg_count_add_store(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi);
/* Insert the k-min-mers of host minimizer-space reads (appends them to the store first). */
mdbg_status mdbg_count_add(mdbg_ctx* ctx, const uint32_t* minimizers, const uint64_t* min_offsets,
                           uint32_t n_reads);
/* Keep abundance >= max(2, min_abundance) (dumpKminmer, CreateMdbg.hpp:3862-3869)
 * and copy the table to the host.  out->hashes etc. are in unspecified order
 * (as upstream, whose writer runs under an omp critical). */
mdbg_status mdbg_count_finalize(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_table_out* out);
/* The same output ((u128, u32)[n] + u32[n * k]) left in HBM: pointers into context-owned device memory,
 * valid until the next finalize / edges call.  Nothing but the five statistics crosses PCIe. */
typedef struct {
    uint32_t k;
    uint64_t n_entries;
    const uint64_t* d_hashes;       /* [2*n_entries]: low 64 bits, high 64 bits */
    const uint32_t* d_abundances;   /* [n_entries] */
    const uint32_t* d_kminmers;     /* [n_entries*k] */
    uint64_t n_instances, n_distinct, checksum, n_rescued;
} mdbg_table_dev;
mdbg_status mdbg_count_finalize_device(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_table_dev* out);
/* Table statistics without the host copy (device-side reduction only). */
mdbg_status mdbg_count_stats(mdbg_ctx* ctx, uint32_t min_abundance, uint64_t* n_entries,
                             uint64_t* n_distinct, uint64_t* n_instances, uint64_t* checksum);

/* rescueKminmers / RescueKminmerFunctor (CreateMdbg.hpp:4517-4640; default mode, --min-abundance <= 1):
 * second pass over the stored reads; reads whose median solid abundance m satisfies m * 0.1f <= 1 (and that
 * hold at least one solid k-min-mer) keep their abundance-1 k-min-mers, which are then emitted by
 * mdbg_count_finalize / counted by mdbg_count_stats next to the abundance >= 2 entries.
 * With more than one rank (mdbg_comm_init) the call is COLLECTIVE and must follow mdbg_count_merge: the solid
 * k-min-mers of all ranks are replicated for the per-read decision, every rank decides for its own reads, and
 * the abundance-1 k-min-mers of rescued reads are sent to their owner ranks, which flag them. */
mdbg_status mdbg_count_rescue(mdbg_ctx* ctx, uint64_t* n_reads_rescued);

/* ---- multi-k: previous-k abundance table and the k >= firstK+1 passes (rows A8/A9) ------------
 * The reference derives the abundance of a k-min-mer from the table of the previous k
 * (_kminmerAbundances, loaded by loadRefinedAbundances, CreateMdbg.cpp:3401-3709): minimum over its two
 * (k-1)-min-mers, absent or 0 => 1, kept when > 1 (getRefinedAbundance CreateMdbg.hpp:3933-4005 for
 * k = firstK+1; IndexKminmerFunctor CreateMdbg.hpp:988-1010,1240-1265,1268-1464 for k >= firstK+2). */
/* previous-k table := the current table's emitted entries (abundance >= max(2,min_abundance) or rescued).
 * With more than one rank the call is COLLECTIVE and must follow mdbg_count_merge: every rank's owned entries
 * are exchanged so that each rank holds the complete previous-k table (20 B per entry); the next-k pass then
 * runs on the rank's own reads and mdbg_count_merge moves the resulting (k-min-mer -> abundance) entries to
 * their owners without adding them up (the abundance is a function of the key). */
mdbg_status mdbg_prev_from_current(mdbg_ctx* ctx, uint32_t min_abundance);
/* insert-or-assign host (hash, abundance) pairs, hashes laid out as in mdbg_table_out / on disk;
 * clear != 0 starts from an empty table (e.g. kminmerData_abundance_prev.txt), clear == 0 patches the
 * existing one (e.g. unitigGraph.nodes.refined_abundances) */
mdbg_status mdbg_prev_load(mdbg_ctx* ctx, const uint64_t* hashes, const uint32_t* abundances, uint64_t n, int clear);
/* after mdbg_count_begin(ctx, k, ...): insert-if-absent every k-min-mer of stored reads [read_lo, read_hi)
 * whose derived abundance is > 1, with that abundance (unitig sequences are appended to the store by the
 * host like reads). mdbg_count_finalize(ctx, 0, ..) then returns kminmerData_abundance.txt of this k. */
mdbg_status mdbg_count_add_store_next_k(mdbg_ctx* ctx, uint64_t read_lo, uint64_t read_hi);

/* ---- edge keys of the node set (first step beyond the count table) ----------------
 * CreateMdbg::EdgeIndexer (src/graph/CreateMdbg.hpp:4010-4232, called first by CreateMdbg::indexEdges,
 * src/graph/CreateMdbg.cpp:1177-1187): every node of the current table (entries mdbg_count_finalize would emit, in
 * their normalized orientation) contributes the hash128 of its normalized (k-1)-prefix and (k-1)-suffix; the
 * dereplicated keys are what the reference writes to edges.bin and builds its edge MPHF over.  hashes[2*i] = low
 * 64 bits, hashes[2*i+1] = high 64 bits (the u128 as it lies in edges.bin); order unspecified; n_edges =
 * EdgeIndexer::_nbEdges; checksum = EdgeIndexer::_checksum (sum of the keys truncated to 64 bits).  The arrays
 * stay valid until the next call on the context.  With more than one rank the call is COLLECTIVE and must follow
 * mdbg_count_merge: every rank derives the keys of the nodes it owns, the distinct local keys travel to their owner
 * rank (same owner function as the count table) and are dereplicated there; each rank returns the keys it owns, and
 * the union over the ranks is the reference's key set (n_edges and checksum add up). */
typedef struct {
    uint32_t k;
    uint64_t n_nodes;             /* table entries that contributed */
    uint64_t n_edges;
    const uint64_t* hashes;       /* [2*n_edges] */
    uint64_t checksum;
    /* Edge values (with several ranks: the offers of a rank's nodes travel to the owner of their key, which folds them):
     * CreateMdbg::indexEdge / successorExists
     * (src/graph/CreateMdbg.cpp:1277-1500) in order-free form.  values[2*i + c], c = orientation class of the offers a
     * key received (0: isReversed == isPrefix, and every offer to a palindromic key; 1: isReversed != isPrefix):
     *   0                               no node extends the key in this class
     *   bit 63 set, bit 34 clear        exactly one node: bits 0-31 = the extending minimizer (KminmerEdge33::_minimizer),
     *                                   bit 32 = _isReversed, bit 33 = _isPrefix
     *   bit 63 and bit 34 set           two or more nodes (_hasMultipleSuccessors; upstream keeps whichever arrived first) */
    const uint64_t* values;       /* [2*n_edges] */
} mdbg_edges_out;
mdbg_status mdbg_edges_index(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_edges_out* out);

/* ---- unitig nodes of the node set (third step) -------------------------------------
 * CreateMdbg::computeUnitigNodes (src/graph/CreateMdbg.cpp:1521-1598; walker ComputeUnitigFunctor::computeUnitigNode2,
 * src/graph/CreateMdbg.hpp:2513-2916, over getNbSuccessors / getNbPredecessors, CreateMdbg.cpp:1902-2380) followed by
 * computeDeterministicUnitigs (CreateMdbg.cpp:1001-1043): the maximal non-branching paths of the node set of the current
 * table (entries mdbg_count_finalize would emit), each as its minimizer sequence -- the first node's k minimizers, then
 * the last minimizer of every further node; circular unitigs start at their k-min-mer with the smallest hash128, in
 * that k-min-mer's normalized orientation -- normalized as a whole (KmerVec::normalize) and hashed (hash128 of the
 * normalized sequence).  Unitig u = minimizers[offsets[u] .. offsets[u+1]); order[i] = the unitig the reference
 * writes as record i of unitigGraph.nodes.bin (ascending u128 hash), i.e. with unitigIndex 2 * i.  hashes[2u] = low,
 * hashes[2u+1] = high 64 bits.  Builds the edge set on the way (as mdbg_edges_index, nothing of it is copied to the
 * host).  Single-context call (n_ranks == 1): with several ranks, gather the node set on one context first.  The
 * arrays stay valid until the next unitigs call on the context. */
typedef struct {
    uint32_t k;
    uint64_t n_nodes;
    uint64_t n_unitigs;
    uint64_t n_minimizers;        /* offsets[n_unitigs] */
    uint64_t n_circular;          /* unitigs that close on themselves */
    uint64_t n_cycle_nodes;       /* oriented nodes on cycles (diagnostic: 2 per node of a circular unitig) */
    const uint64_t* offsets;      /* [n_unitigs + 1] */
    const uint32_t* minimizers;   /* [n_minimizers] */
    const uint64_t* hashes;       /* [2 * n_unitigs] */
    const uint8_t* circular;      /* [n_unitigs] */
    const uint32_t* order;        /* [n_unitigs] */
    /* dumpUnitigAbundances (CreateMdbg.cpp:3335-3390): the abundance of every k-min-mer of unitig u, in sequence order,
     * at node_abundances[offsets[u] - u * (k - 1) ..] (offsets[u+1] - offsets[u] - (k - 1) of them) =
     * the record of unitigGraph.nodes.abundances.bin */
    const uint32_t* node_abundances;
    /* Unitig graph edges -- indexUnitigEdges + computeUnitigEdges (CreateMdbg.cpp:2915-3245; getSuccessors_unitig :2453-2530,
     * getPredecessors_unitig :2631-2695, dumpUnitigEdge :2853-2912) -- as a CSR over ORIENTED unitigs in the reference's
     * numbering (unitigIndex 2 i = record i = unitig order[i], 2 i + 1 = its reverse): edge_targets[edge_offsets[2 i] ..
     * edge_offsets[2 i + 1]) = the successors of record i, edge_targets[edge_offsets[2 i + 1] .. edge_offsets[2 i + 2]) its
     * predecessors, each list in the order one reference thread writes it (= the record of
     * unitigGraph.edges.successors.bin).  n_unitig_edges = _nbUnitigEdges, checksum_edges = _checksum_unitigEdges.
     * Computed for k <= 64; NULL / 0 beyond. */
    uint64_t n_unitig_edges;
    uint64_t checksum_edges;
    const uint64_t* edge_offsets; /* [2 * n_unitigs + 1] */
    const uint32_t* edge_targets; /* [n_unitig_edges] */
    uint64_t checksum_nodes;      /* "Checksum unitig nodes": sum over records of minimizer * size * unitigIndex (CreateMdbg.cpp:3380) */
    uint64_t checksum_abundances; /* "Checksum unitig abundance": sum of abundance * number of abundances (CreateMdbg.cpp:3384) */
    const uint64_t* d_offsets;    /* the same CSR in device memory */
    const uint32_t* d_minimizers;
} mdbg_unitigs_out;
mdbg_status mdbg_unitigs_build(mdbg_ctx* ctx, uint32_t min_abundance, mdbg_unitigs_out* out);

/* ---- multi-GPU (one process per GPU) --------------------------------------- */
/* NCCL is loaded at run time (dlopen libnccl.so.2).  Rank 0 creates an id,
 * the host distributes its 128 bytes to every rank by its own means. */
mdbg_status mdbg_nccl_unique_id(uint8_t id_out[128]);
mdbg_status mdbg_comm_init(mdbg_ctx* ctx, int rank, int n_ranks, const uint8_t id[128]);
/* Owner-partitioned all-to-all of the local table's (k-min-mer, count) pairs
 * followed by reduce-by-key: afterwards this context's table holds exactly the
 * keys it owns with their global abundances.  Collective over all ranks.  Tables filled by
 * mdbg_count_add_store_next_k hold abundance VALUES: equal keys from several ranks are kept once, not summed. */
mdbg_status mdbg_count_merge(mdbg_ctx* ctx);
/* The same merge with (hash128, abundance) records only -- 24 bytes per entry instead of 4 k + 4, no vector gather on
 * the sender, no re-hash on the owner.  The merged table carries no k-min-mer vectors: mdbg_count_finalize returns
 * kminmers = NULL (hashes and abundances as usual), mdbg_edges_index and mdbg_count_rescue are refused.  For the per-k
 * tables of a multi-k loop (the next pass needs no table of another rank, so the merge only has to place every key
 * on exactly one rank). */
mdbg_status mdbg_count_merge_hashes(mdbg_ctx* ctx);

/* ---- synthetic input (benchmark/test support, device side) ------------------ */
/* Fill d_bases with the reads described by (vstart, strand, offsets, lengths) using
 * the counter-based generator of metamdbg_b200/synth.py. */
mdbg_status mdbg_synth_fill_reads(mdbg_ctx* ctx, uint8_t* d_bases, const uint64_t* d_offsets,
                                  const uint64_t* d_vstart, const uint8_t* d_strand,
                                  uint32_t n_reads, uint64_t read_index_base, uint64_t seed, uint32_t err_q24);

#ifdef __cplusplus
}
#endif
#endif /* MDBG_B200_H */
