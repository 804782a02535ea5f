/*
 * mdbg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mdbg_oracle.h).
 *
 * Plain-C restatement of the reference algorithm, written from its behaviour;
 * each function cites the reference lines it follows.  Parity status: PINNED
 * against the reference's own sources compiled by oracle/Makefile into
 * oracle/_ref/libmdbg_ref.so (tests/test_oracle_vs_ref.py, run where
 * /root/reference exists) and against tests/golden/ (npz files) minted from that
 * library (tests/golden/make_golden.py).
 */
#include "mdbg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ murmur */

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

/* MurmurHash3.cpp:328-405 (the _original variant); :246-325 is the same
 * arithmetic returning only h1. */
void orc_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t out[2]) {
    const uint8_t* data = (const uint8_t*)key;
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;

    for (int i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * (size_t)i, 8);
        memcpy(&k2, data + 16 * (size_t)i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }

    const uint8_t* tail = data + (size_t)nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    const int rem = len & 15;
    for (int i = rem - 1; i >= 8; i--) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem > 8 ? 7 : rem - 1); i >= 0; i--) k1 ^= (uint64_t)tail[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }

    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

uint64_t orc_murmur3_x64_128_h1(const void* key, int len, uint32_t seed) {
    uint64_t o[2];
    orc_murmur3_x64_128(key, len, seed, o);
    return o[0];
}

/* ------------------------------------------------------------------ bound */

/* Kmer.hpp:1354-1356: u_int64_t maxHashValue = -1; bound = density * maxHashValue
 * with density a double parameter fed from a float (ReadSelection.hpp:574). */
double orc_minimizer_bound(float density) {
    uint64_t max_hash = (uint64_t)-1;
    return (double)density * (double)max_hash;
}

/* Kmer.hpp:1434 compares `u64 < double`, i.e. (double)h < bound with
 * round-to-nearest conversion; (double)h is monotone in h, so the selected
 * set is {h <= T}.  Binary search for T. */
uint64_t orc_minimizer_threshold(float density, int* none) {
    const double bound = orc_minimizer_bound(density);
    if (none) *none = 0;
    if (!((double)(uint64_t)0 < bound)) { if (none) *none = 1; return 0; }
    uint64_t lo = 0, hi = (uint64_t)-1;            /* invariant: (double)lo < bound */
    if ((double)hi < bound) return hi;
    while (hi - lo > 1) {                            /* (double)hi >= bound */
        uint64_t mid = lo + (hi - lo) / 2;
        if ((double)mid < bound) lo = mid; else hi = mid;
    }
    return lo;
}

/* -------------------------------------------------------------------- HPC */

size_t orc_hpc(const char* seq, size_t len, int hpc, char* out, uint64_t* rle_pos) {
    if (!hpc) {                                      /* Commons.hpp:4192-4199 */
        /* the reference builds string(sequence): stops at the first NUL */
        size_t n = 0;
        while (n < len && seq[n] != '\0') n++;
        memcpy(out, seq, n);
        if (rle_pos) for (size_t i = 0; i < n; i++) rle_pos[i] = i;
        return n;
    }
    /* Commons.hpp:4172-4190.  lastChar starts as '#': a leading '#' run is
     * swallowed exactly as upstream does. */
    size_t n = 0;
    char last = '#';
    uint64_t last_pos = 0;
    for (size_t i = 0; i < len; i++) {
        char c = seq[i];
        if (c == last) continue;
        if (last != '#') {
            out[n] = last;
            if (rle_pos) rle_pos[n] = last_pos;
            n++;
            last_pos = i;
        }
        last = c;
    }
    out[n] = last;
    if (rle_pos) { rle_pos[n] = last_pos; rle_pos[n + 1] = len; }
    n++;
    return n;
}

/* ------------------------------------------------------------------ l-mers */

/* Kmer.hpp:31 comp_NT, :462 ConvertASCII, :488-507 polynom, :509-518 revcomp,
 * :570-589 iterate, :594-611 first/next, :427 updateChoice. */
size_t orc_lmers(const char* seq, size_t len, int l, uint64_t* values, uint8_t* dirs) {
    if (len < (size_t)l) return 0;
    const size_t n = len - (size_t)l + 1;
    const uint64_t mask = (l >= 32) ? ~0ULL : ((1ULL << (2 * l)) - 1);
    static const uint64_t comp[4] = {2, 3, 0, 1};
    uint64_t fwd = 0, rc = 0;
    int bad = -1;                                    /* index of last bad char in window */
    for (int i = 0; i < l; i++) {
        unsigned char ch = (unsigned char)seq[i];
        uint64_t c = (ch >> 1) & 3;
        fwd = (fwd << 2) + c;
        if ((ch >> 3) & 1) bad = i;
    }
    /* revcomp of the first window: complement each base, reverse order */
    for (int i = 0; i < l; i++) {
        uint64_t c = (fwd >> (2 * i)) & 3;           /* base l-1-i */
        rc = (rc << 2) | comp[c];
    }
    size_t idx = 0;
    {
        int dir = (fwd < rc) ? 0 : 1;
        values[idx] = (bad < 0) ? (dir ? rc : fwd) : ~0ULL;
        dirs[idx] = (uint8_t)dir;
        idx++;
    }
    for (size_t p = (size_t)l; p < len; p++) {
        unsigned char ch = (unsigned char)seq[p];
        uint64_t c = (ch >> 1) & 3;
        if ((ch >> 3) & 1) bad = l - 1; else bad--;
        fwd = ((fwd << 2) + c) & mask;
        rc = ((rc >> 2) + (comp[c] << (2 * (l - 1)))) & mask;
        int dir = (fwd < rc) ? 0 : 1;
        values[idx] = (bad < 0) ? (dir ? rc : fwd) : ~0ULL;
        dirs[idx] = (uint8_t)dir;
        idx++;
    }
    return n;
}

/* ------------------------------------------------------------------ sketch */

static int bl_contains(const uint32_t* bl, size_t n, uint32_t v) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (bl[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo < n && bl[lo] == v;
}

size_t orc_sketch_read(const char* seq, size_t len, int l, float density, int hpc,
                       const uint32_t* blacklist, size_t n_blacklist,
                       uint32_t* minimizers, uint32_t* positions, uint8_t* directions,
                       size_t cap) {
    char* hs = (char*)malloc(len + 1);
    size_t hl = orc_hpc(seq, len, hpc, hs, NULL);
    size_t n_out = 0;
    if (hl >= (size_t)l) {
        size_t n = hl - (size_t)l + 1;
        uint64_t* vals = (uint64_t*)malloc(n * sizeof(uint64_t));
        uint8_t* dirs = (uint8_t*)malloc(n);
        orc_lmers(hs, hl, l, vals, dirs);
        const double bound = orc_minimizer_bound(density);
        /* Kmer.hpp:1395: for(pos=_trimBps; pos<kmers.size()-_trimBps; pos++), _trimBps=1
         * (unsigned arithmetic: n>=1 here so n-1 cannot wrap). */
        for (size_t pos = 1; pos + 1 < n; pos++) {
            uint64_t v = vals[pos];
            uint64_t h = orc_murmur3_x64_128_h1(&v, 8, 42);       /* Kmer.hpp:1421 */
            if ((double)h < bound) {                                 /* Kmer.hpp:1434 */
                /* Kmer.hpp:1437: unordered_set<u32>::find(u64) truncates the key */
                if (n_blacklist > 0 && bl_contains(blacklist, n_blacklist, (uint32_t)v)) continue;
                if (n_out < cap) {
                    minimizers[n_out] = (uint32_t)v;                 /* Kmer.hpp:1441, u32 truncation */
                    positions[n_out] = (uint32_t)pos;
                    directions[n_out] = dirs[pos];
                }
                n_out++;
            }
        }
        free(vals);
        free(dirs);
    }
    free(hs);
    return n_out;
}

size_t orc_sketch_batch(const char* bases, const uint64_t* offsets, size_t n_reads,
                        int l, float density, int hpc,
                        const uint32_t* blacklist, size_t n_blacklist,
                        uint64_t* min_offsets, uint32_t* minimizers, uint32_t* positions,
                        uint8_t* directions, size_t cap) {
    size_t total = 0;
    for (size_t r = 0; r < n_reads; r++) {
        min_offsets[r] = total;
        size_t room = total < cap ? cap - total : 0;
        size_t w = total < cap ? total : cap;
        total += orc_sketch_read(bases + offsets[r], (size_t)(offsets[r + 1] - offsets[r]), l, density,
                                 hpc, blacklist, n_blacklist, minimizers + w, positions + w,
                                 directions + w, room);
    }
    min_offsets[n_reads] = total;
    return total;
}

/* ---------------------------------------------------------------- side outputs */

void orc_read_aux(const char* seq, const char* qual, size_t len, size_t qual_len, int l, int hpc,
                  const uint32_t* positions, size_t n_minimizers,
                  float* mean_quality, double* complexity, uint8_t* qualities) {
    /* ReadSelection.hpp:870-879 */
    long double error_sum = 0;
    for (size_t i = 0; i < qual_len; i++) {
        unsigned char c = (unsigned char)qual[i];
        float e = 0.0f;                              /* table is zero outside 33..127 (ReadSelection.hpp:101-104) */
        if (c >= 33 && c <= 127) { float q = (float)(unsigned char)(c - 33); e = powf(10.0f, -q / 10.0f); }
        error_sum += e;
    }
    float mean_err = (float)(error_sum / (long double)qual_len);
    *mean_quality = -10.0f * log10f(mean_err);

    /* computeSequenceComplexity(seq, 64, 32), ReadSelection.hpp:1171-1228 */
    {
        const size_t w = 64, step = 32;
        const double lw = (double)w - 2;
        size_t nk = len >= 3 ? len - 2 : 0;
        double nb_windows = 0, sum = 0;
        for (size_t ii = 0; ii < nk; ii += step) {
            double counts[64];
            for (int i = 0; i < 64; i++) counts[i] = 0;
            size_t n = 0;
            for (size_t i = ii; i < nk; i++) {
                unsigned c0 = ((unsigned char)seq[i] >> 1) & 3, c1 = ((unsigned char)seq[i + 1] >> 1) & 3,
                         c2 = ((unsigned char)seq[i + 2] >> 1) & 3;
                counts[(c0 << 4) | (c1 << 2) | c2] += 1;
                if (++n == w) break;
            }
            if (n < w) continue;
            double score = 0;
            for (int i = 0; i < 64; i++) score += counts[i] * (counts[i] - 1) / 2.0;
            score /= (lw - 1);
            nb_windows += 1;
            sum += score;
        }
        *complexity = sum / nb_windows;
    }

    /* per-minimizer min quality */
    if (n_minimizers == 0) return;
    if (qual_len == 0) {
        for (size_t j = 0; j < n_minimizers; j++) qualities[j] = 1;
        return;
    }
    char* hs = (char*)malloc(len + 1);
    uint64_t* rle = (uint64_t*)malloc((len + 2) * sizeof(uint64_t));
    size_t hl = orc_hpc(seq, len, hpc, hs, rle);
    if (!hpc) rle[hl] = len;
    for (size_t j = 0; j < n_minimizers; j++) {
        size_t a = (size_t)rle[positions[j]], b = (size_t)rle[positions[j] + (size_t)l];
        uint8_t mq = 255;
        for (size_t i = a; i < b; i++) {
            uint8_t q = (uint8_t)((unsigned char)qual[i] - 33);
            if (q < mq) mq = q;
        }
        qualities[j] = mq;
    }
    free(hs); free(rle);
}

size_t orc_apply_density(const uint32_t* m, size_t n, float density, uint32_t* out) {
    uint64_t max_hash = (uint64_t)-1;
    double bound = density * max_hash;              /* Commons.hpp:2515-2516: float * u64 -> float, widened */
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t v = m[i];
        uint64_t h = orc_murmur3_x64_128_h1(&v, 8, 42);
        if (h < bound) out[o++] = m[i];
    }
    return o;
}

/* -------------------------------------------------------- purge palindromes */

/* KmerVec::isPalindrome, Commons.hpp:918-921: first size/2 entries equal the
 * reversed last size/2. */
static int is_palindrome(const uint32_t* v, size_t k) {
    for (size_t i = 0; i < k / 2; i++)
        if (v[i] != v[k - 1 - i]) return 0;
    return 1;
}

size_t orc_purge_palindrome(const uint32_t* m, size_t n, size_t first_k, size_t last_k,
                            uint32_t* out, uint8_t* keep) {
    uint8_t* banned = (uint8_t*)calloc(n ? n : 1, 1);
    uint32_t* win = (uint32_t*)malloc((last_k ? last_k : 1) * sizeof(uint32_t));
    for (;;) {                                       /* Commons.hpp:1627 */
        int has = 0;
        for (size_t k = first_k; k < last_k && !has; k++) {
            long i_max = (long)n - (long)k + 1;      /* Commons.hpp:1635 */
            for (long i = 0; i < i_max && !has; i++) {
                if (banned[i]) continue;
                size_t cnt = 0;
                for (size_t j = (size_t)i; j < n; j++) {   /* next k non-banned from i */
                    if (banned[j]) continue;
                    win[cnt++] = m[j];
                    if (cnt == k) break;
                }
                if (cnt == k && is_palindrome(win, k)) {
                    banned[i] = 1;                   /* Commons.hpp:1669: ban the first */
                    has = 1;
                }
            }
        }
        if (!has) break;
    }
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        if (keep) keep[i] = !banned[i];
        if (!banned[i]) out[o++] = m[i];
    }
    free(banned);
    free(win);
    return o;
}

/* ---------------------------------------------------------------- k-min-mers */

size_t orc_kminmers(const uint32_t* m, size_t n, int k, uint32_t* vecs, uint8_t* reversed) {
    if (n < (size_t)k) return 0;
    size_t nw = n - (size_t)k + 1;
    for (size_t i = 0; i < nw; i++) {
        const uint32_t* w = m + i;
        /* KmerVec::normalize Commons.hpp:886-916: first differing position
         * decides; all-equal (palindromic vector) => reversed. */
        int rev = 1;
        for (int j = 0; j < k; j++) {
            uint32_t a = w[j], b = w[k - 1 - j];
            if (a == b) continue;
            rev = (a < b) ? 0 : 1;
            break;
        }
        for (int j = 0; j < k; j++) vecs[i * (size_t)k + j] = rev ? w[k - 1 - j] : w[j];
        if (reversed) reversed[i] = (uint8_t)rev;
    }
    return nw;
}

void orc_hash128(const uint32_t* vec, int k, uint64_t out[2]) {
    orc_murmur3_x64_128(vec, k * 4, 0, out);         /* Commons.hpp:956-961 */
}

/* -------------------------------------------------------------------- count */

static int g_cmp_k;
static int cmp_vec(const void* a, const void* b) {   /* KmerVec operator<, Commons.hpp:754-773 */
    const uint32_t* x = (const uint32_t*)a;
    const uint32_t* y = (const uint32_t*)b;
    for (int i = 0; i < g_cmp_k; i++) {
        if (x[i] == y[i]) continue;
        return x[i] < y[i] ? -1 : 1;
    }
    return 0;
}

size_t orc_count(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                 uint32_t min_abundance, uint32_t** vecs_out, uint64_t** hashes_out,
                 uint32_t** abundances_out, uint64_t* n_instances, uint64_t* n_distinct) {
    /* partitionKminmers (CreateMdbg.hpp:3652-3724) only groups equal vectors;
     * the partition count is unobservable, so one partition is used. */
    size_t total = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t n = (size_t)(offs[r + 1] - offs[r]);
        if (n >= (size_t)k) total += n - (size_t)k + 1;
    }
    uint32_t* all = (uint32_t*)malloc((total ? total : 1) * (size_t)k * sizeof(uint32_t));
    size_t w = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t n = (size_t)(offs[r + 1] - offs[r]);
        w += orc_kminmers(mins + offs[r], n, k, all + w * (size_t)k, NULL);
    }
    g_cmp_k = k;
    qsort(all, total, (size_t)k * sizeof(uint32_t), cmp_vec);   /* CreateMdbg.hpp:3797 */

    uint32_t* vecs = (uint32_t*)malloc((total ? total : 1) * (size_t)k * sizeof(uint32_t));
    uint64_t* hashes = (uint64_t*)malloc((total ? total : 1) * 2 * sizeof(uint64_t));
    uint32_t* abs_ = (uint32_t*)malloc((total ? total : 1) * sizeof(uint32_t));
    size_t n_out = 0, distinct = 0;
    size_t i = 0;
    while (i < total) {                              /* run-length count, CreateMdbg.hpp:3808-3838 */
        size_t j = i + 1;
        while (j < total && cmp_vec(all + i * (size_t)k, all + j * (size_t)k) == 0) j++;
        uint32_t ab = (uint32_t)(j - i);
        distinct++;
        /* dumpKminmer CreateMdbg.hpp:3862-3869 (first pass); min_abundance == UINT32_MAX is a
         * test-only "keep every distinct k-min-mer" mode used to check the multi-GPU merge */
        if (min_abundance == 0xFFFFFFFFu || (ab > 1 && ab >= min_abundance)) {
            memcpy(vecs + n_out * (size_t)k, all + i * (size_t)k, (size_t)k * sizeof(uint32_t));
            orc_hash128(all + i * (size_t)k, k, hashes + 2 * n_out);
            abs_[n_out] = ab;
            n_out++;
        }
        i = j;
    }
    free(all);
    if (n_instances) *n_instances = total;
    if (n_distinct) *n_distinct = distinct;
    *vecs_out = vecs; *hashes_out = hashes; *abundances_out = abs_;
    return n_out;
}

/* ------------------------------------------------------- hash -> abundance lookup (sorted array) */

typedef struct { uint64_t h1, h2; uint32_t ab; uint32_t idx; } HEntry;

static int cmp_hentry(const void* a, const void* b) {
    const HEntry* x = (const HEntry*)a; const HEntry* y = (const HEntry*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

static HEntry* build_lookup(const uint64_t* hashes, const uint32_t* ab, size_t n) {
    HEntry* t = (HEntry*)malloc((n ? n : 1) * sizeof(HEntry));
    for (size_t i = 0; i < n; i++) { t[i].h1 = hashes[2 * i]; t[i].h2 = hashes[2 * i + 1]; t[i].ab = ab[i]; t[i].idx = (uint32_t)i; }
    qsort(t, n, sizeof(HEntry), cmp_hentry);
    return t;
}

/* returns 1 and *ab when present (the LAST loaded duplicate wins, like map[key] = value) */
static int lookup(const HEntry* t, size_t n, uint64_t h1, uint64_t h2, uint32_t* ab) {
    size_t lo = 0, hi = n;
    while (lo < hi) {                                /* upper bound of (h1,h2) */
        size_t mid = (lo + hi) / 2;
        if (t[mid].h1 < h1 || (t[mid].h1 == h1 && t[mid].h2 <= h2)) lo = mid + 1; else hi = mid;
    }
    if (lo == 0 || t[lo - 1].h1 != h1 || t[lo - 1].h2 != h2) return 0;
    *ab = t[lo - 1].ab;
    return 1;
}

static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

size_t orc_rescue(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                  const uint64_t* solid_hashes, const uint32_t* solid_ab, size_t n_solid,
                  uint32_t** vecs_out, uint64_t** hashes_out, uint64_t* n_reads_rescued) {
    HEntry* t = build_lookup(solid_hashes, solid_ab, n_solid);
    size_t cap = 1024, n_out = 0, n_rr = 0;
    uint32_t* vecs = (uint32_t*)malloc(cap * (size_t)k * 4);
    uint64_t* hashes = (uint64_t*)malloc(cap * 16);
    for (size_t r = 0; r < n_reads; r++) {
        size_t n = (size_t)(offs[r + 1] - offs[r]);
        if (n < (size_t)k) continue;
        size_t nw = n - (size_t)k + 1;
        uint32_t* w = (uint32_t*)malloc(nw * (size_t)k * 4);
        uint32_t* ab = (uint32_t*)malloc(nw * 4);
        uint32_t* sorted = (uint32_t*)malloc(nw * 4);
        uint8_t* solid = (uint8_t*)malloc(nw);
        orc_kminmers(mins + offs[r], n, k, w, NULL);
        int all_one = 1;
        for (size_t i = 0; i < nw; i++) {            /* CreateMdbg.hpp:4593-4607 */
            uint64_t h[2]; uint32_t a;
            orc_hash128(w + i * (size_t)k, k, h);
            if (lookup(t, n_solid, h[0], h[1], &a)) { ab[i] = a; solid[i] = 1; all_one = 0; }
            else { ab[i] = 1; solid[i] = 0; }
            sorted[i] = ab[i];
        }
        qsort(sorted, nw, 4, cmp_u32);               /* Utils::compute_median, Commons.hpp:2973-2988 */
        uint32_t median = (nw % 2 == 0) ? (uint32_t)(sorted[nw / 2 - 1] + sorted[nw / 2]) / 2 : sorted[nw / 2];
        double cutoff = median * 0.1f;               /* CreateMdbg.hpp:4612: u32 * float */
        if (!(cutoff > 1) && !all_one) {
            n_rr++;
            for (size_t i = 0; i < nw; i++) {
                if (solid[i]) continue;
                if (n_out == cap) {
                    cap *= 2;
                    vecs = (uint32_t*)realloc(vecs, cap * (size_t)k * 4);
                    hashes = (uint64_t*)realloc(hashes, cap * 16);
                }
                memcpy(vecs + n_out * (size_t)k, w + i * (size_t)k, (size_t)k * 4);
                orc_hash128(w + i * (size_t)k, k, hashes + 2 * n_out);
                n_out++;
            }
        }
        free(w); free(ab); free(sorted); free(solid);
    }
    free(t);
    *vecs_out = vecs; *hashes_out = hashes;
    if (n_reads_rescued) *n_reads_rescued = n_rr;
    return n_out;
}

typedef struct { uint64_t h1, h2; uint32_t ab; const uint32_t* vec; } NEntry;
static int cmp_nentry(const void* a, const void* b) {
    const NEntry* x = (const NEntry*)a; const NEntry* y = (const NEntry*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : 1;
    return 0;
}

size_t orc_next_k(const uint32_t* mins, const uint64_t* offs, size_t n_reads, int k,
                  const uint64_t* prev_hashes, const uint32_t* prev_ab, size_t n_prev,
                  uint32_t** vecs_out, uint64_t** hashes_out, uint32_t** abundances_out) {
    HEntry* t = build_lookup(prev_hashes, prev_ab, n_prev);
    size_t total = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t n = (size_t)(offs[r + 1] - offs[r]);
        if (n >= (size_t)k) total += n - (size_t)k + 1;
    }
    uint32_t* all = (uint32_t*)malloc((total ? total : 1) * (size_t)k * 4);
    NEntry* ent = (NEntry*)malloc((total ? total : 1) * sizeof(NEntry));
    size_t ne = 0, wpos = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t n = (size_t)(offs[r + 1] - offs[r]);
        if (n < (size_t)k) continue;
        size_t nw = n - (size_t)k + 1, np = n - (size_t)k + 2;     /* (k-1)-windows */
        uint32_t* sub = (uint32_t*)malloc(np * (size_t)(k - 1) * 4);
        uint32_t* prev = (uint32_t*)malloc(np * 4);
        orc_kminmers(mins + offs[r], n, k - 1, sub, NULL);          /* getPrevAbundances, CreateMdbg.hpp:1240-1265 */
        for (size_t i = 0; i < np; i++) {
            uint64_t h[2]; uint32_t a;
            orc_hash128(sub + i * (size_t)(k - 1), k - 1, h);
            prev[i] = lookup(t, n_prev, h[0], h[1], &a) ? a : 1;
        }
        orc_kminmers(mins + offs[r], n, k, all + wpos * (size_t)k, NULL);
        for (size_t i = 0; i < nw; i++) {
            uint32_t a = prev[i] < prev[i + 1] ? prev[i] : prev[i + 1];   /* getAbundance, CreateMdbg.hpp:988-1010 */
            if (a <= 1) continue;                                           /* CreateMdbg.hpp:1429-1431 */
            NEntry* e = &ent[ne++];
            e->vec = all + (wpos + i) * (size_t)k;
            uint64_t h[2];
            orc_hash128(e->vec, k, h);
            e->h1 = h[0]; e->h2 = h[1]; e->ab = a;
        }
        wpos += nw;
        free(sub); free(prev);
    }
    qsort(ent, ne, sizeof(NEntry), cmp_nentry);
    uint32_t* vecs = (uint32_t*)malloc((ne ? ne : 1) * (size_t)k * 4);
    uint64_t* hashes = (uint64_t*)malloc((ne ? ne : 1) * 16);
    uint32_t* abs_ = (uint32_t*)malloc((ne ? ne : 1) * 4);
    size_t n_out = 0;
    for (size_t i = 0; i < ne; i++) {                /* lazy_emplace_l: first insertion wins (all equal anyway) */
        if (i > 0 && ent[i].h1 == ent[i - 1].h1 && ent[i].h2 == ent[i - 1].h2) continue;
        memcpy(vecs + n_out * (size_t)k, ent[i].vec, (size_t)k * 4);
        hashes[2 * n_out] = ent[i].h1; hashes[2 * n_out + 1] = ent[i].h2;
        abs_[n_out] = ent[i].ab;
        n_out++;
    }
    free(all); free(ent); free(t);
    *vecs_out = vecs; *hashes_out = hashes; *abundances_out = abs_;
    return n_out;
}

/* ------------------------------------------------------- edge keys of the node set (row F1)
 * CreateMdbg::EdgeIndexer (src/graph/CreateMdbg.hpp:4010-4232): every node (normalized k-min-mer, as dumped to
 * kminmerData_min.txt) contributes the hash128 of its normalized prefix (first k-1 minimizers, KmerVec::prefix
 * Commons.hpp:851-856) and of its normalized suffix (last k-1, :858-862) -- partitionNode :4106-4120; the hashes
 * are sorted and dereplicated per partition (:4140-4200).  The partition order is not observable in the key set,
 * so the restatement returns the distinct keys sorted by (h1,h2); _nbEdges = their number, _checksum = sum of the
 * keys truncated to 64 bits (u_int64_t += u_int128_t), i.e. the sum of the low words (h2). */
typedef struct { uint64_t h1, h2; } EKey;
static int cmp_ekey(const void* a, const void* b) {
    const EKey* x = (const EKey*)a; const EKey* y = (const EKey*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : (x->h2 > y->h2 ? 1 : 0);
    return 0;
}

size_t orc_edge_index(const uint32_t* vecs, size_t n, int k, uint64_t** hashes, uint64_t* checksum) {
    const int km = k - 1;
    EKey* keys = (EKey*)malloc((2 * n + 1) * sizeof(EKey));
    uint32_t* tmp = (uint32_t*)malloc((size_t)(km > 0 ? km : 1) * sizeof(uint32_t));
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        for (int side = 1; side >= 0; side--) {          /* suffix first, then prefix (order is irrelevant) */
            const uint32_t* w = vecs + i * (size_t)k + side;
            int rev = 1;                                  /* KmerVec::normalize, Commons.hpp:886-916 */
            for (int j = 0; j < km; j++) {
                uint32_t a = w[j], b = w[km - 1 - j];
                if (a == b) continue;
                rev = (a < b) ? 0 : 1;
                break;
            }
            for (int j = 0; j < km; j++) tmp[j] = rev ? w[km - 1 - j] : w[j];
            uint64_t h[2];
            orc_hash128(tmp, km, h);
            keys[m].h1 = h[0]; keys[m].h2 = h[1];
            m++;
        }
    }
    qsort(keys, m, sizeof(EKey), cmp_ekey);
    uint64_t* out = (uint64_t*)malloc((2 * m + 2) * sizeof(uint64_t));
    size_t n_out = 0;
    uint64_t cs = 0;
    for (size_t i = 0; i < m; i++) {
        if (i && keys[i].h1 == keys[i - 1].h1 && keys[i].h2 == keys[i - 1].h2) continue;
        out[2 * n_out] = keys[i].h1; out[2 * n_out + 1] = keys[i].h2;
        cs += keys[i].h2;
        n_out++;
    }
    free(keys);
    free(tmp);
    *hashes = out;
    if (checksum) *checksum = cs;
    return n_out;
}

/* ------------------------------------------------------- edge values of the node set (row F1, second step)
 * CreateMdbg::indexEdge + successorExists (src/graph/CreateMdbg.cpp:1277-1500), order-free form.  Every node offers
 *   to the key of its normalized SUFFIX : (minimizer = first element, isReversed = suffix was reversed, isPrefix = 0)
 *   to the key of its normalized PREFIX : (minimizer = last element,  isReversed = prefix was reversed, isPrefix = 1)
 * successorExists folds an offer into an existing slot when the key's vector is a palindrome, or when
 * (isReversed, isPrefix) equals the slot's or is its exact complement -- i.e. there are two orientation classes,
 * A = {(0,0),(1,1)} and B = {(0,1),(1,0)} (a palindromic key has one class) -- and then only marks
 * hasMultipleSuccessors; which offer of a class was recorded first depends on the thread arrival order upstream,
 * and the consumers (getSuccessors_unitig / getPredecessors_unitig, :2039-2130, :2287-2380) use the recorded
 * minimizer only while the class is not marked.  Canonical value per (key, class): count = 0, 1 or 2 (= two or
 * more), and for count == 1 the single offer's minimizer / isReversed / isPrefix. */
typedef struct { uint64_t h1, h2; uint32_t cls, min, rev, pre; } EOffer;
static int cmp_eoffer(const void* a, const void* b) {
    const EOffer* x = (const EOffer*)a; const EOffer* y = (const EOffer*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : 1;
    if (x->cls != y->cls) return x->cls < y->cls ? -1 : 1;
    return 0;
}

size_t orc_edge_values(const uint32_t* vecs, size_t n, int k, uint64_t** hashes, uint32_t** values) {
    const int km = k - 1;
    EOffer* of = (EOffer*)malloc((2 * n + 1) * sizeof(EOffer));
    uint32_t* tmp = (uint32_t*)malloc((size_t)(km > 0 ? km : 1) * sizeof(uint32_t));
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        const uint32_t* v = vecs + i * (size_t)k;
        for (int side = 1; side >= 0; side--) {          /* 1: suffix (isPrefix = 0), 0: prefix (isPrefix = 1) */
            const uint32_t* w = v + side;
            int rev = 1, pal = 1;
            for (int j = 0; j < km; j++) {
                uint32_t a = w[j], b = w[km - 1 - j];
                if (a == b) continue;
                rev = (a < b) ? 0 : 1;
                break;
            }
            for (int j = 0; j < km / 2; j++) if (w[j] != w[km - 1 - j]) { pal = 0; break; }   /* KmerVec::isPalindrome */
            for (int j = 0; j < km; j++) tmp[j] = rev ? w[km - 1 - j] : w[j];
            uint64_t h[2];
            orc_hash128(tmp, km, h);
            const uint32_t pre = side ? 0u : 1u;
            of[m].h1 = h[0]; of[m].h2 = h[1];
            of[m].cls = pal ? 0u : (((uint32_t)rev == pre) ? 0u : 1u);
            of[m].min = side ? v[0] : v[k - 1];
            of[m].rev = (uint32_t)rev; of[m].pre = pre;
            m++;
        }
    }
    qsort(of, m, sizeof(EOffer), cmp_eoffer);
    uint64_t* hk = (uint64_t*)malloc((2 * m + 2) * sizeof(uint64_t));
    uint32_t* val = (uint32_t*)calloc(8 * m + 8, sizeof(uint32_t));      /* per key: 2 classes x {count, min, rev, pre} */
    size_t ne = 0;
    for (size_t i = 0; i < m;) {
        size_t j = i;
        while (j < m && of[j].h1 == of[i].h1 && of[j].h2 == of[i].h2) j++;
        hk[2 * ne] = of[i].h1; hk[2 * ne + 1] = of[i].h2;
        for (size_t t = i; t < j; t++) {
            uint32_t* c = val + 8 * ne + 4 * of[t].cls;
            if (c[0] == 0) { c[0] = 1; c[1] = of[t].min; c[2] = of[t].rev; c[3] = of[t].pre; }
            else { c[0] = 2; c[1] = c[2] = c[3] = 0; }
        }
        ne++;
        i = j;
    }
    free(of);
    free(tmp);
    *hashes = hk;
    *values = val;
    return ne;
}

/* ------------------------------------------------------- unitig nodes of the node set (row F1, third step)
 * CreateMdbg::computeUnitigNodes / ComputeUnitigFunctor::computeUnitigNode2 (src/graph/CreateMdbg.cpp:1521-1598,
 * CreateMdbg.hpp:2513-2916) followed by computeDeterministicUnitigs (CreateMdbg.cpp:1001-1043), run sequentially over
 * the nodes in file order.  getNbSuccessors (CreateMdbg.cpp:1902-2135): the successors of an oriented node are the
 * offers, in the orientation class that matches its (k-1)-suffix as written, to the key of that suffix -- exactly one
 * (unmarked) offer <=> single successor = suffix + recorded minimizer; getNbPredecessors (:2227-2380) is the mirror
 * image (predecessor of v = reverse of the successor of reverse(v)).  A unitig grows from its source node forwards
 * while "single successor whose single predecessor exists", closes as a circle when the successor is the source
 * again, else grows backwards; circular unitigs are rotated to start at their k-min-mer with the smallest normalized
 * hash128, in that k-min-mer's normalized orientation.  computeDeterministicUnitigs normalizes every unitig's
 * minimizer sequence (KmerVec::normalize), sorts by the hash128 of the sequence and numbers them 0, 2, 4, ...
 * Returns the unitigs in that final order as a CSR (offs[n_unitigs + 1], mins). */
typedef struct {
    const uint32_t* vecs; size_t n; int k;
    uint32_t* order;                 /* node ids sorted by vector */
    const uint64_t* ekeys; const uint32_t* evals; size_t ne;
} UCtx;
static int g_ucmp_k;
static const uint32_t* g_ucmp_vecs;
static int cmp_node_idx(const void* a, const void* b) {
    const uint32_t* x = g_ucmp_vecs + (size_t)(*(const uint32_t*)a) * g_ucmp_k;
    const uint32_t* y = g_ucmp_vecs + (size_t)(*(const uint32_t*)b) * g_ucmp_k;
    for (int j = 0; j < g_ucmp_k; j++) if (x[j] != y[j]) return x[j] < y[j] ? -1 : 1;
    return 0;
}
/* KmerVec::normalize (Commons.hpp:886-916) of w[0..len): out = normalized, returns isReversed */
static int u_normalize(const uint32_t* w, int len, uint32_t* out) {
    int rev = 1;
    for (int j = 0; j < len; j++) {
        uint32_t a = w[j], b = w[len - 1 - j];
        if (a == b) continue;
        rev = (a < b) ? 0 : 1;
        break;
    }
    for (int j = 0; j < len; j++) out[j] = rev ? w[len - 1 - j] : w[j];
    return rev;
}
static long u_node_id(const UCtx* c, const uint32_t* vec) {           /* id of the node whose normalized form is norm(vec) */
    uint32_t tmp[256];
    u_normalize(vec, c->k, tmp);
    size_t lo = 0, hi = c->n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        const uint32_t* y = c->vecs + (size_t)c->order[mid] * c->k;
        int cmp = 0;
        for (int j = 0; j < c->k; j++) if (tmp[j] != y[j]) { cmp = tmp[j] < y[j] ? -1 : 1; break; }
        if (cmp == 0) return (long)c->order[mid];
        if (cmp < 0) hi = mid; else lo = mid + 1;
    }
    return -1;
}
static int u_succ(const UCtx* c, const uint32_t* vec, uint32_t* out) {
    const int km = c->k - 1;
    uint32_t tmp[256];
    const uint32_t* S = vec + 1;
    const int rev = u_normalize(S, km, tmp);
    int pal = 1;
    for (int j = 0; j < km / 2; j++) if (S[j] != S[km - 1 - j]) { pal = 0; break; }
    uint64_t h[2];
    orc_hash128(tmp, km, h);
    size_t lo = 0, hi = c->ne;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        uint64_t a = c->ekeys[2 * mid], b = c->ekeys[2 * mid + 1];
        if (a == h[0] && b == h[1]) { lo = mid; break; }
        if (a < h[0] || (a == h[0] && b < h[1])) lo = mid + 1; else hi = mid;
    }
    if (lo >= c->ne || c->ekeys[2 * lo] != h[0] || c->ekeys[2 * lo + 1] != h[1]) return 0;
    const uint32_t* cl = c->evals + 8 * lo + 4 * (pal ? 0 : (rev ? 0 : 1));
    if (cl[0] != 1) return 0;
    for (int j = 0; j < km; j++) out[j] = S[j];
    out[km] = cl[1];
    return 1;
}
static int u_pred(const UCtx* c, const uint32_t* vec, uint32_t* out) {
    uint32_t r[256], t[256];
    for (int j = 0; j < c->k; j++) r[j] = vec[c->k - 1 - j];
    if (!u_succ(c, r, t)) return 0;
    for (int j = 0; j < c->k; j++) out[j] = t[c->k - 1 - j];
    return 1;
}
typedef struct { uint64_t h1, h2; uint32_t* m; uint32_t len; } UOut;
static int cmp_uout(const void* a, const void* b) {
    const UOut* x = (const UOut*)a; const UOut* y = (const UOut*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : 1;
    return 0;
}

size_t orc_unitigs(const uint32_t* vecs, size_t n, int k, uint64_t** offs_out, uint32_t** mins_out, uint64_t** hashes_out) {
    UCtx c; c.vecs = vecs; c.n = n; c.k = k;
    c.order = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) c.order[i] = (uint32_t)i;
    g_ucmp_k = k; g_ucmp_vecs = vecs;
    qsort(c.order, n, sizeof(uint32_t), cmp_node_idx);
    uint64_t* ek; uint32_t* ev;
    c.ne = orc_edge_values(vecs, n, k, &ek, &ev);
    c.ekeys = ek; c.evals = ev;
    uint8_t* marked = (uint8_t*)calloc(n + 1, 1);
    UOut* outs = (UOut*)malloc((n + 1) * sizeof(UOut));
    size_t n_out = 0;
    uint32_t a[256], b[256], t[256];
    for (size_t src = 0; src < n; src++) {
        const uint32_t* source = vecs + src * (size_t)k;
        if (marked[src]) continue;                                   /* isNodeUnitigged(source) || (sourceRC) */
        size_t cap = (size_t)k + 16, len = (size_t)k;
        uint32_t* u = (uint32_t*)malloc(cap * sizeof(uint32_t));
        memcpy(u, source, (size_t)k * 4);
        uint32_t start[256], end[256];
        memcpy(start, source, (size_t)k * 4); memcpy(end, source, (size_t)k * 4);
        int circular = 0;
        for (;;) {                                                    /* forwards */
            if (!u_succ(&c, end, a)) break;
            if (!u_pred(&c, a, b)) break;
            if (memcmp(a, source, (size_t)k * 4) == 0) { circular = 1; memcpy(end, a, (size_t)k * 4); memcpy(start, a, (size_t)k * 4); break; }
            memcpy(end, a, (size_t)k * 4);
            if (len + 1 > cap) { cap *= 2; u = (uint32_t*)realloc(u, cap * sizeof(uint32_t)); }
            u[len++] = end[k - 1];
        }
        if (!circular) {                                              /* backwards */
            size_t bcap = 16, blen = 0;
            uint32_t* back = (uint32_t*)malloc(bcap * sizeof(uint32_t));
            for (;;) {
                if (!u_pred(&c, start, a)) break;
                if (!u_succ(&c, a, b)) break;
                memcpy(start, a, (size_t)k * 4);
                if (blen + 1 > bcap) { bcap *= 2; back = (uint32_t*)realloc(back, bcap * sizeof(uint32_t)); }
                back[blen++] = start[0];
            }
            uint32_t* full = (uint32_t*)malloc((len + blen + 1) * sizeof(uint32_t));
            for (size_t i = 0; i < blen; i++) full[i] = back[blen - 1 - i];
            memcpy(full + blen, u, len * 4);
            free(u); free(back);
            u = full; len += blen;
        } else {
            /* rotate to the window with the smallest normalized hash, in its normalized orientation */
            const size_t nw = len - (size_t)k + 1;
            uint64_t bh1 = ~0ULL, bh2 = ~0ULL; size_t bi = (size_t)-1; int brev = 0;
            for (size_t i = 0; i < nw; i++) {
                int r = u_normalize(u + i, k, t);
                uint64_t h[2];
                orc_hash128(t, k, h);
                if (h[0] < bh1 || (h[0] == bh1 && h[1] < bh2)) { bh1 = h[0]; bh2 = h[1]; bi = i; brev = r; }
            }
            if (brev) {
                for (size_t i = 0; i < len / 2; i++) { uint32_t x = u[i]; u[i] = u[len - 1 - i]; u[len - 1 - i] = x; }
                for (size_t i = 0; i < nw; i++) {
                    uint64_t h[2];
                    orc_hash128(u + i, k, h);                         /* as written, not normalized */
                    if (h[0] == bh1 && h[1] == bh2) { bi = i; break; }
                }
            }
            uint32_t* rot = (uint32_t*)malloc((len + 1) * sizeof(uint32_t));
            size_t rl = 0;
            for (int j = 0; j < k; j++) rot[rl++] = u[bi + (size_t)j];
            for (size_t i = bi + 1; i < nw; i++) rot[rl++] = u[i + (size_t)k - 1];
            for (size_t i = 0; i < bi; i++) rot[rl++] = u[i + (size_t)k - 1];
            free(u);
            u = rot; len = rl;
            u_normalize(u, k, start); memcpy(end, start, (size_t)k * 4);
        }
        /* isValid: none of start / end / their reverses is unitigged yet */
        long is = u_node_id(&c, start), ie = u_node_id(&c, end);
        if (is < 0 || ie < 0 || marked[is] || marked[ie]) { free(u); continue; }
        for (size_t i = 0; i + (size_t)k <= len; i++) {
            long id = u_node_id(&c, u + i);
            if (id >= 0) marked[id] = 1;
        }
        outs[n_out].m = u; outs[n_out].len = (uint32_t)len;
        n_out++;
    }
    /* computeDeterministicUnitigs */
    size_t total = 0;
    for (size_t i = 0; i < n_out; i++) {
        uint32_t* nm = (uint32_t*)malloc(((size_t)outs[i].len + 1) * sizeof(uint32_t));
        u_normalize(outs[i].m, (int)outs[i].len, nm);
        free(outs[i].m);
        outs[i].m = nm;
        uint64_t h[2];
        orc_hash128(nm, (int)outs[i].len, h);
        outs[i].h1 = h[0]; outs[i].h2 = h[1];
        total += outs[i].len;
    }
    qsort(outs, n_out, sizeof(UOut), cmp_uout);
    uint64_t* offs = (uint64_t*)malloc((n_out + 2) * sizeof(uint64_t));
    uint32_t* mins = (uint32_t*)malloc((total + 1) * sizeof(uint32_t));
    uint64_t* hs = (uint64_t*)malloc((2 * n_out + 2) * sizeof(uint64_t));
    size_t pos = 0;
    for (size_t i = 0; i < n_out; i++) {
        offs[i] = pos;
        memcpy(mins + pos, outs[i].m, (size_t)outs[i].len * 4);
        pos += outs[i].len;
        hs[2 * i] = outs[i].h1; hs[2 * i + 1] = outs[i].h2;
        free(outs[i].m);
    }
    offs[n_out] = pos;
    free(outs); free(marked); free(c.order); free(ek); free(ev);
    *offs_out = offs; *mins_out = mins;
    if (hashes_out) *hashes_out = hs; else free(hs);
    return n_out;
}

/* ------------------------------------------------------- unitig graph edges (row F1, fourth step)
 * CreateMdbg::indexUnitigEdges / indexUnitigEdge / indexEdgeUnitig (src/graph/CreateMdbg.cpp:2915-3086) and
 * computeUnitigEdges / computeUnitigEdge / getSuccessors_unitig / getPredecessors_unitig / dumpUnitigEdge
 * (:3088-3245, 2453-2530, 2631-2695, 2853-2912), run sequentially over unitigGraph.nodes.bin in file order (record i
 * has unitigIndex 2 i; 2 i + 1 is its reverse).  Index: the first and (when different) the last k-min-mer of every unitig,
 * normalized, enter the lists of their normalized (k-1)-prefix and (k-1)-suffix keys as {unitigIndex of the orientation
 * in which the normalized k-min-mer appears, isReversed of the key, isPrefix}.  Query: the successors of a unitig are
 * the entries of its last k-min-mer's suffix key whose oriented (k-1)-mer continues that suffix (isPrefix: the entry's
 * unitig, unless that is the querying unitig's own reverse; else: the reverse of the entry's unitig, unless that is...
 * the same exclusion seen from the other side); the predecessors are the successors of the reversed unitig.  Lists are
 * returned in the order a single thread produces (file order of the offering unitigs, first node before last node,
 * prefix entry before suffix entry).  CSR over ORIENTED unitigs: list 2 i = successors of record i, list 2 i + 1 = its
 * predecessors.  *checksum = _checksum_unitigEdges (sum of from * to, predecessors with the reversed from index). */
typedef struct { uint64_t h1, h2, seq; uint32_t idx, rev, pre; } UEnt;
static int cmp_uent(const void* a, const void* b) {
    const UEnt* x = (const UEnt*)a; const UEnt* y = (const UEnt*)b;
    if (x->h1 != y->h1) return x->h1 < y->h1 ? -1 : 1;
    if (x->h2 != y->h2) return x->h2 < y->h2 ? -1 : 1;
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
    return 0;
}
static void ue_offer(UEnt* ents, size_t* m, const uint32_t* node, int k, uint32_t u, uint64_t seq) {
    uint32_t nn[256], key[256];
    const int nrev = u_normalize(node, k, nn);
    const uint32_t idx = nrev ? u + 1 : u;
    for (int side = 0; side < 2; side++) {               /* prefix entry first, then suffix entry (indexEdgeUnitig) */
        const int rev = u_normalize(nn + side, k - 1, key);
        uint64_t h[2];
        orc_hash128(key, k - 1, h);
        UEnt* e = ents + (*m)++;
        e->h1 = h[0]; e->h2 = h[1]; e->seq = seq * 2 + (uint64_t)side; e->idx = idx; e->rev = (uint32_t)rev; e->pre = side ? 0u : 1u;
    }
}
/* successors of oriented unitig x whose last k-min-mer (as oriented) is `end` */
static size_t ue_succ(const UEnt* ents, size_t m, const uint32_t* end, int k, uint32_t x, uint32_t* out) {
    const int km = k - 1;
    uint32_t key[256];
    const uint32_t* S = end + 1;
    const int srev = u_normalize(S, km, key);
    int pal = 1;
    for (int j = 0; j < km / 2; j++) if (S[j] != S[km - 1 - j]) { pal = 0; break; }
    uint64_t h[2];
    orc_hash128(key, km, h);
    size_t lo = 0, hi = m;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (ents[mid].h1 < h[0] || (ents[mid].h1 == h[0] && ents[mid].h2 < h[1])) lo = mid + 1; else hi = mid;
    }
    size_t n = 0;
    for (size_t i = lo; i < m && ents[i].h1 == h[0] && ents[i].h2 == h[1]; i++) {
        const UEnt* e = ents + i;
        if (e->pre) {
            if (!(pal || (int)e->rev == srev)) continue;                  /* vec_suffix == v */
            if (x == (e->idx ^ 1u)) continue;
            out[n++] = e->idx;
        } else {
            if (!(pal || (int)e->rev != srev)) continue;                  /* vec_suffix == reverse(v) */
            if (x == e->idx) continue;
            out[n++] = e->idx ^ 1u;
        }
    }
    return n;
}

size_t orc_unitig_edges(const uint32_t* mins, const uint64_t* offs, size_t n_unitigs, int k, uint64_t** eoff_out,
                        uint32_t** etgt_out, uint64_t* checksum) {
    UEnt* ents = (UEnt*)malloc((4 * n_unitigs + 1) * sizeof(UEnt));
    size_t m = 0;
    for (size_t i = 0; i < n_unitigs; i++) {
        const uint32_t* u = mins + offs[i];
        const size_t L = (size_t)(offs[i + 1] - offs[i]);
        const uint32_t* start = u; const uint32_t* end = u + L - (size_t)k;
        ue_offer(ents, &m, start, k, (uint32_t)(2 * i), 2 * (uint64_t)i);
        if (memcmp(start, end, (size_t)k * 4) != 0) ue_offer(ents, &m, end, k, (uint32_t)(2 * i), 2 * (uint64_t)i + 1);
    }
    qsort(ents, m, sizeof(UEnt), cmp_uent);
    uint64_t* eoff = (uint64_t*)malloc((2 * n_unitigs + 2) * sizeof(uint64_t));
    size_t cap = 4 * n_unitigs + 16, tot = 0;
    uint32_t* etgt = (uint32_t*)malloc(cap * sizeof(uint32_t));
    uint32_t* tmp = (uint32_t*)malloc((m + 1) * sizeof(uint32_t));
    uint32_t rc[256];
    uint64_t cs = 0;
    for (size_t i = 0; i < n_unitigs; i++) {
        const uint32_t* u = mins + offs[i];
        const size_t L = (size_t)(offs[i + 1] - offs[i]);
        for (int o = 0; o < 2; o++) {
            const uint32_t x = (uint32_t)(2 * i) + (uint32_t)o;
            const uint32_t* end;
            if (!o) end = u + L - (size_t)k;
            else { for (int j = 0; j < k; j++) rc[j] = u[k - 1 - j]; end = rc; }     /* last k-min-mer of the reversed unitig */
            const size_t n = ue_succ(ents, m, end, k, x, tmp);
            eoff[2 * i + (size_t)o] = tot;
            if (tot + n > cap) { cap = 2 * (tot + n); etgt = (uint32_t*)realloc(etgt, cap * sizeof(uint32_t)); }
            /* dumpUnitigEdge: `_checksum_unitigEdges += unitigIndexFrom * unitigIndexTo` multiplies two 32-bit UnitigType
             * values, so every product wraps at 2^32 before it is added to the 64-bit sum */
            for (size_t j = 0; j < n; j++) { etgt[tot++] = tmp[j]; cs += (uint64_t)(uint32_t)(x * tmp[j]); }
        }
    }
    eoff[2 * n_unitigs] = tot;
    free(ents); free(tmp);
    *eoff_out = eoff; *etgt_out = etgt;
    if (checksum) *checksum = cs;
    return tot;
}

uint64_t orc_table_checksum(const uint64_t* hashes, const uint32_t* abundances, size_t n) {
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) s += (uint64_t)abundances[i] * hashes[2 * i + 1];
    return s;
}

void orc_free(void* p) { free(p); }
