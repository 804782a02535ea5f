/*
 * mdbg_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of metaMDBG v1.4's minimizer-sketch + k-min-mer
 * count path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (metamdbg_b200/csrc, libmdbg_b200.so) never links, loads or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  The restatement is pinned against the real
 * reference sources compiled in oracle/_ref (see oracle/ref_shim.cpp and
 * tests/test_oracle_vs_ref.py) and against the golden vectors in
 * tests/golden/.
 */
#ifndef MDBG_ORACLE_H
#define MDBG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/utils/MurmurHash3.cpp:246-325 -- returns h1 only. */
uint64_t orc_murmur3_x64_128_h1(const void* key, int len, uint32_t seed);
/* src/utils/MurmurHash3.cpp:328-405 -- out[0]=h1, out[1]=h2. */
void orc_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t out[2]);

/* Selection bound of MinimizerParser (src/utils/kmer/Kmer.hpp:1354-1356):
 * (double)(float)density * (double)(uint64_t)-1. */
double orc_minimizer_bound(float density);
/* Largest u64 T with (double)T < bound, i.e. "select iff hash <= T";
 * returns 0 and sets *none=1 when no hash value is selected. */
uint64_t orc_minimizer_threshold(float density, int* none);

/* EncoderRLE::execute, src/Commons.hpp:4163-4203.  out must hold len+1 bytes
 * (an empty read yields "#", as upstream), rle_pos len+2 entries (may be NULL).  Returns the compressed length L'.
 * hpc==0 is the identity copy of the else-branch. */
size_t orc_hpc(const char* seq, size_t len, int hpc, char* out, uint64_t* rle_pos);

/* KmerModel::iterate (src/utils/kmer/Kmer.hpp:570-611) on an (already
 * compressed) sequence: values[i] = canonical l-mer or UINT64_MAX when the
 * window holds a character with bit 3 set; dirs[i] = 0 fwd / 1 revcomp.
 * Returns the number of l-mers (0 if len < l). */
size_t orc_lmers(const char* seq, size_t len, int l, uint64_t* values, uint8_t* dirs);

/* EncoderRLE + MinimizerParser::parse (src/utils/kmer/Kmer.hpp:1373-1456),
 * as chained by ReadSelectionFunctor (src/readSelection/ReadSelection.hpp:682-690).
 * blacklist = sorted u32 array of repetitive minimizers (may be NULL/0).
 * Writes at most cap entries; returns the true number of minimizers. */
size_t orc_sketch_read(const char* seq, size_t len, int l, float density, int hpc,
                       const uint32_t* blacklist, size_t n_blacklist,
                       uint32_t* minimizers, uint32_t* positions, uint8_t* directions,
                       size_t cap);

/* Batch form over a concatenated buffer: read r = bases[offsets[r]..offsets[r+1]).
 * min_offsets has n_reads+1 entries.  Returns total minimizers (writes at
 * most cap). */
size_t orc_sketch_batch(const char* bases, const uint64_t* offsets, size_t n_reads,
                        int l, float density, int hpc,
                        const uint32_t* blacklist, size_t n_blacklist,
                        uint64_t* min_offsets, uint32_t* minimizers, uint32_t* positions,
                        uint8_t* directions, size_t cap);

/* Side outputs of ReadSelectionFunctor::operator() (src/readSelection/ReadSelection.hpp):
 *   mean_quality  = -10.0f * log10f((float)(errorSum / n)) with errorSum the long double sum of
 *                   pow(10.0f, -(q-33)/10.0f) over the quality bytes (:870-879; NaN when qual is empty)
 *   complexity    = computeSequenceComplexity(seq, 64, 32) (:1171-1228), the read is dropped when > 5
 *   qualities[j]  = getMinQuality over raw [rlePos[pos_j], rlePos[pos_j + l]) (:1135, :1302-1320), or 1
 *                   when qual is empty (:1049-1053)
 * positions = HPC positions of the read's minimizers (as returned by orc_sketch_read). */
void orc_read_aux(const char* seq, const char* qual, size_t len, size_t qual_len, int l, int hpc,
                  const uint32_t* positions, size_t n_minimizers,
                  float* mean_quality, double* complexity, uint8_t* qualities);

/* Utils::applyDensityThreshold, src/Commons.hpp:2507-2550: keep[i] = 1 iff Murmur(u64(m[i]), seed 42) < density*2^64.
 * Returns the number kept (written to out in order). */
size_t orc_apply_density(const uint32_t* m, size_t n, float density, uint32_t* out);

/* Commons::purgePalindrome, src/Commons.hpp:1617-1723.  out holds n entries;
 * keep (may be NULL) receives 0/1 per input position.  Returns n'. */
size_t orc_purge_palindrome(const uint32_t* m, size_t n, size_t first_k, size_t last_k,
                            uint32_t* out, uint8_t* keep);

/* MDBG::getKminmers_complete active branch (src/Commons.hpp:5282-5361) +
 * KmerVec::normalize (src/Commons.hpp:886-916): for i in [0, n-k] the
 * normalized window, k u32 each, and its isReversed flag.  Returns n-k+1
 * (0 if n < k). */
size_t orc_kminmers(const uint32_t* m, size_t n, int k, uint32_t* vecs, uint8_t* reversed);

/* KmerVec::hash128, src/Commons.hpp:941-969: out[0]=h1 (high 64 bits of the
 * u128), out[1]=h2 (low 64 bits). */
void orc_hash128(const uint32_t* vec, int k, uint64_t out[2]);

/* KminmerCounter first pass (src/graph/CreateMdbg.hpp:3652-3883): exact
 * multiplicity of every distinct normalized k-min-mer of the reads in the
 * CSR (mins, offs); keeps abundance >= max(2, min_abundance).
 * Outputs (malloc'ed, caller frees with orc_free): vecs (n*k u32, sorted
 * lexicographically), hashes (n*2 u64: h1,h2), abundances (n u32).
 * *n_instances / *n_distinct receive the pre-filter totals.  Returns n. */
size_t orc_count(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                 uint32_t min_abundance, uint32_t** vecs, uint64_t** hashes,
                 uint32_t** abundances, uint64_t* n_instances, uint64_t* n_distinct);

/* rescueKminmers / RescueKminmerFunctor (src/graph/CreateMdbg.hpp:4517-4640), default mode.
 * solid_hashes (n_solid x {h1,h2}) / solid_ab = the abundance >= 2 table.  For every read whose
 * median abundance m (solid abundance, else 1; Utils::compute_median, src/Commons.hpp:2973-2988)
 * satisfies m * 0.1f <= 1 and that has at least one solid k-min-mer, the non-solid k-min-mers are
 * emitted with abundance 1, in read order.  Outputs malloc'ed: vecs (n*k), hashes (n*2: h1,h2).
 * Returns n; *n_reads_rescued = number of qualifying reads. */
size_t orc_rescue(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                  const uint64_t* solid_hashes, const uint32_t* solid_ab, size_t n_solid,
                  uint32_t** vecs, uint64_t** hashes, uint64_t* n_reads_rescued);

/* k >= firstK+1 pass: KminmerCounter::getRefinedAbundance (src/graph/CreateMdbg.hpp:3933-4005)
 * and IndexKminmerFunctor (src/graph/CreateMdbg.hpp:988-1010, 1240-1265, 1268-1464) give the same
 * table: every distinct k-min-mer of the reads whose min over its two (k-1)-min-mers of the previous
 * table (absent or 0 => 1) is > 1, with that value.  prev_hashes: n_prev x {h1,h2}.  Outputs
 * malloc'ed and sorted by (h1,h2): vecs (n*k), hashes (n*2), abundances (n).  Returns n. */
size_t orc_next_k(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                  const uint64_t* prev_hashes, const uint32_t* prev_ab, size_t n_prev,
                  uint32_t** vecs, uint64_t** hashes, uint32_t** abundances);

/* CreateMdbg::EdgeIndexer (CreateMdbg.hpp:4010-4232): the distinct hash128 of the normalized (k-1)-prefix and
 * (k-1)-suffix of every node.  vecs: n normalized k-min-mers (kminmerData_min.txt rows).  *hashes: malloc'ed
 * n_edges x {h1,h2} sorted by (h1,h2); *checksum = sum of the low words.  Returns n_edges (_nbEdges). */
size_t orc_edge_index(const uint32_t* vecs, size_t n, int k, uint64_t** hashes, uint64_t* checksum);

/* CreateMdbg::indexEdge / successorExists (CreateMdbg.cpp:1277-1500) in order-free form: per distinct edge key
 * (sorted by (h1,h2)) two orientation classes x {count (0, 1, 2 = two or more), minimizer, isReversed, isPrefix};
 * the last three are those of the single offer when count == 1, else 0.  *values: n_edges x 8 u32. */
size_t orc_edge_values(const uint32_t* vecs, size_t n, int k, uint64_t** hashes, uint32_t** values);

/* CreateMdbg::computeUnitigNodes + computeDeterministicUnitigs (CreateMdbg.cpp:1521-1598, 1001-1043; walker
 * ComputeUnitigFunctor::computeUnitigNode2, CreateMdbg.hpp:2513-2916): the content of unitigGraph.nodes.bin -- every
 * unitig's normalized minimizer sequence, sorted by the hash128 of that sequence (unitigIndex = 2 * position).
 * *offs: n_unitigs + 1, *mins: concatenated sequences, *hashes (may be NULL): n_unitigs x {h1,h2}.  Returns n_unitigs. */
size_t orc_unitigs(const uint32_t* vecs, size_t n, int k, uint64_t** offs, uint32_t** mins, uint64_t** hashes);

/* CreateMdbg::indexUnitigEdges + computeUnitigEdges (CreateMdbg.cpp:2915-3245, getSuccessors_unitig :2453-2530,
 * getPredecessors_unitig :2631-2695, dumpUnitigEdge :2853-2912) on the records of unitigGraph.nodes.bin (mins / offs, record i
 * = unitigIndex 2 i).  CSR over oriented unitigs: list 2 i = successors of record i, list 2 i + 1 = its predecessors, each
 * in the order one thread produces.  Returns the number of edges (_nbUnitigEdges); *checksum = _checksum_unitigEdges. */
size_t orc_unitig_edges(const uint32_t* mins, const uint64_t* offs, size_t n_unitigs, int k, uint64_t** edge_offs,
                        uint32_t** edge_targets, uint64_t* checksum);

/* Order-free fingerprint used by the reference's debug log
 * (src/graph/CreateMdbg.cpp:3321): sum abundance * (u64)hash128 mod 2^64,
 * where (u64)hash128 = low 64 bits = h2. */
uint64_t orc_table_checksum(const uint64_t* hashes, const uint32_t* abundances, size_t n);

void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
