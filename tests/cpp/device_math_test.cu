// device_math_test.cu -- runs the kernel's pure arithmetic (metamdbg_b200/csrc/common.cuh, bitmath.cuh) on the
// CPU and checks it against the oracle (oracle/mdbg_oracle.c) and naive loops.  Built with nvcc, host code only:
// the functions are __host__ __device__, the device side uses the same source with PTX intrinsics.
//
//   device_math_test [--exhaustive]     exit 0 = all good
//
// --exhaustive walks all 2^32 keys of the candidate test (about a minute on a few cores); the default samples.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

#include "../../metamdbg_b200/csrc/bitmath.cuh"
extern "C" {
#include "../../oracle/mdbg_oracle.h"
}

using namespace mdbg;
#ifndef MDBG_TEST_VARIANT
#define MDBG_TEST_VARIANT 0                  // which arithmetic variant of the sketch kernel's register block is tested
#endif
#if MDBG_TEST_VARIANT == 1
using namespace mdbg::k1v1;                  // murmur_s1_u32(key, risk), roll16_fast(W, thr, s_hi, s_lo), lmer_from_packed
#endif

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state += 0x9E3779B97F4A7C15ull; return mix64(rng_state); }

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

// ---- murmur_h1_u64 / murmur128_u32vec vs the oracle's byte-wise restatement -------------------
static void test_murmur() {
    for (int i = 0; i < 2000000; i++) {
        uint64_t key = (i & 1) ? rnd() : (rnd() & 0xFFFFFFFFull);
        if (i < 4) key = (i == 0) ? 0 : (i == 1) ? ~0ull : (i == 2) ? 0xFFFFFFFFull : 1;
        CHECK(murmur_h1_u64(key) == orc_murmur3_x64_128_h1(&key, 8, 42), "h1 key=%llx", (unsigned long long)key);
    }
    for (int k = 1; k <= 40; k++)
        for (int rep = 0; rep < 2000; rep++) {
            uint32_t v[40];
            for (int i = 0; i < k; i++) v[i] = (uint32_t)rnd();
            uint64_t o1, o2, ref[2];
            murmur128_u32vec([&](int i) { return v[i]; }, k, o1, o2);
            orc_hash128(v, k, ref);                      // ref[0] = h1 (high 64 bits), ref[1] = h2 (low 64 bits)
            CHECK(o1 == ref[0] && o2 == ref[1], "hash128 k=%d", k);
        }
}

// ---- candidate test: no selected key may be rejected -------------------------------------------
struct CandStats { uint64_t keys = 0, selected = 0, candidates = 0, undecided0 = 0, missed = 0; };

static void cand_range(uint64_t lo, uint64_t hi, uint64_t step, uint64_t T, CandStats* st) {
#if MDBG_TEST_VARIANT == 0
    const uint32_t thp1 = (uint32_t)(T >> 32) + 1u;
#else
    const uint32_t thp1 = (uint32_t)(T >> 32) + S1_SLACK;
#endif
    CandStats s;
    for (uint64_t k = lo; k < hi; k += step) {
        const uint32_t key = (uint32_t)k;
#if MDBG_TEST_VARIANT == 0
        const uint32_t s1 = murmur_s1_u32(key);
        const bool cand = s1 <= thp1;
#else
        uint32_t risk = 0;
        const uint32_t s1 = murmur_s1_u32(key, risk);
        const bool cand = s1 <= thp1 || risk >= S1_RISK;
#endif
        const bool sel = murmur_h1_u64((uint64_t)key) <= T;
#if MDBG_TEST_VARIANT == 0
        s.keys++; s.selected += sel; s.candidates += cand; s.undecided0 += (s1 == 0);
#else
        s.keys++; s.selected += sel; s.candidates += cand; s.undecided0 += (risk >= S1_RISK);
#endif
        if (sel && !cand) s.missed++;
#if MDBG_TEST_VARIANT == 0
        // documented three-way classification (common.cuh): 1 <= s1 < T_hi  =>  certainly selected
        if (s1 >= 1 && s1 < (uint32_t)(T >> 32) && !sel) s.missed++;
#else
#endif
    }
    *st = s;
}

static void test_candidates(bool exhaustive) {
    const float densities[] = {0.005f, 0.0025f, 0.05f, 0.5f, 0.9f, 1e-6f};
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for (size_t d = 0; d < sizeof densities / sizeof *densities; d++) {
        int none = 0;
        const uint64_t T = orc_minimizer_threshold(densities[d], &none);
        if (none) continue;
        const bool full = exhaustive && d < 2;            // the two production densities (assembly, ONT correction)
        const uint64_t step = full ? 1 : 257;              // odd stride: a 2^32/257 sample touching all residues
        std::vector<CandStats> st(nt);
        std::vector<std::thread> th;
        const uint64_t span = (1ull << 32) / nt;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back(cand_range, t * span, (t + 1 == nt) ? (1ull << 32) : (t + 1) * span, step, T, &st[t]);
        for (auto& x : th) x.join();
        CandStats s;
        for (auto& x : st) { s.keys += x.keys; s.selected += x.selected; s.candidates += x.candidates;
                             s.undecided0 += x.undecided0; s.missed += x.missed; }
#if MDBG_TEST_VARIANT == 0
        printf("density %-8g T=%016llx keys %llu%s selected %llu candidates %llu (s1==0: %llu) misclassified %llu\n",
#else
        printf("density %-8g T=%016llx keys %llu%s selected %llu candidates %llu (carry case: %llu) misclassified %llu\n",
#endif
               densities[d], (unsigned long long)T, (unsigned long long)s.keys, full ? " (all)" : "",
               (unsigned long long)s.selected, (unsigned long long)s.candidates, (unsigned long long)s.undecided0,
               (unsigned long long)s.missed);
        CHECK(s.missed == 0, "candidate test rejected a selected key at density %g", densities[d]);
        CHECK(s.candidates >= s.selected, "candidates < selected");
    }
}

// ---- bit tricks vs naive loops -----------------------------------------------------------------------
static void test_bits() {
    for (int i = 0; i < 4000000; i++) {
        uint32_t x = (uint32_t)rnd();
        if (i & 1) x &= (uint32_t)rnd() & (uint32_t)rnd();            // sparse words: many zero bytes
        if ((i & 7) == 3) x &= 0x00FF00FFu << (8 * (i & 8 ? 1 : 0));
        uint32_t nz = 0, ev = 0;
        for (int b = 0; b < 4; b++) if ((x >> (8 * b)) & 0xFF) nz |= 0x80u << (8 * b);
        for (int b = 0; b < 16; b++) ev |= ((x >> (2 * b)) & 1u) << b;
        CHECK(nonzero_bytes(x) == nz, "nonzero_bytes %08x", x);
        CHECK(even_bits16(x) == ev, "even_bits16 %08x", x);
        uint32_t r = 0;
        for (int b = 0; b < 32; b++) r |= ((x >> b) & 1u) << (31 - b);
        CHECK(brev32(x) == r, "brev32 %08x", x);
        const uint32_t y = (uint32_t)rnd(), s = (uint32_t)rnd() & 31;
        CHECK(funnel_l(x, y, s) == (uint32_t)((((uint64_t)y << 32 | x) << s) >> 32), "funnel_l");
        CHECK(umulhi32(x, y) == (uint32_t)(((unsigned __int128)x * y) >> 32), "umulhi32");
    }
    for (uint32_t m = 0; m < 65536; m++) {                              // all 16-bit keep masks
        uint32_t f[4] = {0, 0, 0, 0};
        for (int b = 0; b < 16; b++) if ((m >> b) & 1u) f[b >> 2] |= 0x80u << (8 * (b & 3));
        CHECK(flags_to_mask16(f[0], f[1], f[2], f[3]) == m, "flags_to_mask16 %04x", m);
    }
    for (uint32_t v = 0; v < 256; v++) {                                // all 4-base words
        const uint32_t c[4] = {v & 3, (v >> 2) & 3, (v >> 4) & 3, (v >> 6) & 3};
        const uint32_t w = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        CHECK(pack4_msb(w) == ((c[0] << 6) | (c[1] << 4) | (c[2] << 2) | c[3]), "pack4_msb %02x", v);
        CHECK(pack4_lsb(w) == (c[0] | (c[1] << 2) | (c[2] << 4) | (c[3] << 6)), "pack4_lsb %02x", v);
    }
}

// ---- 16 rolled l-mers of one lane vs the oracle's l-mer iterator ---------------------------------
template <int L>
static void test_roll(uint64_t T) {
    static const char ALPHA[4] = {'A', 'C', 'T', 'G'};                  // code = (c >> 1) & 3
#if MDBG_TEST_VARIANT == 0
    const uint32_t thp1 = (uint32_t)(T >> 32) + 1u;
#else
    const uint32_t thp1 = (uint32_t)(T >> 32) + S1_SLACK;
#endif
    uint64_t n_sel = 0, n_cand = 0;
    for (int rep = 0; rep < 200000; rep++) {
        uint32_t W[8];
        char seq[32];
        for (int q = 0; q < 8; q++) {
            W[q] = 0;
            for (int b = 0; b < 4; b++) {
                uint32_t c = (uint32_t)rnd() & 3;
                if (rep % 5 == 0 && (rnd() & 3)) c = (rep / 5) & 3;       // low-complexity stretches too
                W[q] |= c << (8 * b);
                seq[4 * q + b] = ALPHA[c];
            }
        }
        uint64_t vals[32]; uint8_t dirs[32];
        const size_t n = orc_lmers(seq, L + 15, L, vals, dirs);
        CHECK(n == 16, "orc_lmers returned %zu", n);
#if MDBG_TEST_VARIANT == 0
        uint32_t sel_fwd = 0xDEADBEEF;
        const uint32_t cand = roll16_fast<L>(W, thp1, sel_fwd);
        uint32_t exact = 0, last = 0;
        for (int j = 0; j < 16; j++) {
#else
        uint32_t s_hi = 0, s_lo = 0;
        const uint32_t cand = roll16_fast<L>(W, thp1, s_hi, s_lo);
        uint32_t exact = 0;
        for (int j = 0; j < 16; j++)
#endif
            if (orc_murmur3_x64_128_h1(&vals[j], 8, 42) <= T) exact |= 1u << j;
#if MDBG_TEST_VARIANT == 0
            if ((cand >> j) & 1u) last = j;
        }
#else
#endif
        CHECK((exact & ~cand) == 0, "roll16_fast<%d> lost a selected position (%04x vs %04x)", L, cand, exact);
#if MDBG_TEST_VARIANT == 0
        if (cand) {                                                     // sel_fwd = forward l-mer of the last candidate
#else
        for (int j = 0; j < 16; j++) {                                  // every position's l-mer from the packed codes
#endif
            uint32_t fwd = 0;
#if MDBG_TEST_VARIANT == 0
            for (int t = 0; t < L; t++) fwd = (fwd << 2) | ((W[(last + t) >> 2] >> (8 * ((last + t) & 3))) & 3u);
            if (L < 16) fwd &= (1u << (2 * L)) - 1u;
            CHECK(sel_fwd == fwd, "sel_fwd");
#else
            for (int t = 0; t < L; t++) fwd = (fwd << 2) | ((W[(j + t) >> 2] >> (8 * ((j + t) & 3))) & 3u);
            if (L < 16) fwd &= (1u << ((2 * L) & 31)) - 1u;
            CHECK(lmer_from_packed<L>(s_hi, s_lo, j) == fwd, "lmer_from_packed j=%d", j);
#endif
            const uint32_t rc = revcomp_lmer<L>(fwd);
#if MDBG_TEST_VARIANT == 0
            CHECK((uint64_t)(fwd < rc ? fwd : rc) == vals[last] && dirs[last] == (fwd < rc ? 0 : 1), "canonical/dir");
#else
            CHECK((uint64_t)(fwd < rc ? fwd : rc) == vals[j] && dirs[j] == (fwd < rc ? 0 : 1), "canonical/dir");
#endif
        }
        n_sel += __builtin_popcount(exact); n_cand += __builtin_popcount(cand);
    }
    printf("roll16_fast<%d>: %llu selected, %llu candidates over 3.2M positions\n", L, (unsigned long long)n_sel,
           (unsigned long long)n_cand);
}

int main(int argc, char** argv) {
    const bool exhaustive = argc > 1 && !strcmp(argv[1], "--exhaustive");
    test_murmur();
    test_bits();
    int none = 0;
    test_roll<15>(orc_minimizer_threshold(0.005f, &none));
    test_roll<15>(orc_minimizer_threshold(0.3f, &none));
    test_roll<13>(orc_minimizer_threshold(0.05f, &none));
    test_roll<16>(orc_minimizer_threshold(0.05f, &none));
    test_candidates(exhaustive);
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
