// aux_emu_test.cpp -- runs the REAL read_aux_kernel source (metamdbg_b200/csrc/aux.cu) in the warp emulator and compares mean read quality, DUST-like complexity, the low-complexity
// filter and the per-minimizer minimum qualities with the oracle.  The error table and the final float arithmetic
// restate what api.cu does around the kernel (ensure_err_table, mdbg_sketch_batch_q).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "warp_emu.hpp"
#include AUX_SOURCE

extern "C" {
#include "../../oracle/mdbg_oracle.h"
}

using namespace mdbg;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 20) { printf("FAIL line %d: ", __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

static bool same_float(float a, float b) { return (std::isnan(a) && std::isnan(b)) || memcmp(&a, &b, 4) == 0; }
static bool same_double(double a, double b) { return (std::isnan(a) && std::isnan(b)) || memcmp(&a, &b, 8) == 0; }

static void run_case(const std::vector<std::string>& seqs, const std::vector<std::string>& quals, bool have_q, uint32_t l,
                     float density, int hpc, uint64_t* n_checked) {
    const uint32_t n = (uint32_t)seqs.size();
    std::vector<uint64_t> offs(n + 1, 0);
    for (uint32_t r = 0; r < n; r++) offs[r + 1] = offs[r] + seqs[r].size();
    const uint64_t nb = offs[n];
    std::vector<uint8_t> bstore(nb + 96), qstore(nb + 96);
    uint8_t* bases = bstore.data() + ((16 - ((uintptr_t)bstore.data() & 15)) & 15);
    uint8_t* q = qstore.data() + 16;
    for (uint32_t r = 0; r < n; r++) {
        memcpy(bases + offs[r], seqs[r].data(), seqs[r].size());
        if (have_q) memcpy(q + offs[r], quals[r].data(), quals[r].size());
    }
    // minimizer positions from the oracle (the kernel's input is the sketch kernel's output), tight CSR
    std::vector<uint64_t> moff(n + 1, 0);
    std::vector<uint32_t> pos, nmin(n);
    for (uint32_t r = 0; r < n; r++) {
        const size_t len = seqs[r].size();
        std::vector<uint32_t> m(len + 1), p(len + 1);
        std::vector<uint8_t> d(len + 1);
        const size_t k = orc_sketch_read(seqs[r].data(), len, (int)l, density, hpc, nullptr, 0, m.data(), p.data(), d.data(), len + 1);
        pos.insert(pos.end(), p.begin(), p.begin() + k);
        nmin[r] = (uint32_t)k;
        moff[r + 1] = moff[r] + k;
    }
    const uint64_t total = moff[n];
    // error table (api.cu ensure_err_table; Commons.hpp:2338-2341)
    uint64_t fixed[256]; uint8_t tz[256];
    for (int c = 0; c < 256; c++) {
        float e = 0.0f;
        if (c >= 33 && c <= 127) { float qq = (float)(uint8_t)(c - 33); e = powf(10.0f, -qq / 10.0f); }
        fixed[c] = (uint64_t)ldexpl((long double)e, ERR_SHIFT);
        tz[c] = fixed[c] ? (uint8_t)__builtin_ctzll(fixed[c]) : 255;
    }
    std::vector<uint32_t> rawA(total + 1), rawB(total + 1);
    std::vector<uint8_t> out_qual(total + 1, 0xEE), lmin(n), low(n);
    std::vector<uint64_t> slo(n), shi(n);
    std::vector<double> cplx(n);
    std::vector<uint32_t> nmin_io = nmin;
    AuxArgs a{};
    a.bases = bases; a.bases_end = bases + nb; a.quals = have_q ? q : nullptr;
    a.offsets = offs.data(); a.n_reads = n; a.l = l; a.hpc = (uint32_t)hpc;
    a.err_fixed = fixed; a.err_tz = tz;
    a.exact_off = moff.data(); a.cap_shift = 5; a.cap_const = 32;
    a.n_min = nmin_io.data(); a.pad_pos = pos.data(); a.pad_raw_a = rawA.data(); a.pad_raw_b = rawB.data();
    a.out_qual = out_qual.data();
    a.err_sum_lo = slo.data(); a.err_sum_hi = shi.data(); a.err_lmin = lmin.data();
    a.complexity = cplx.data(); a.low_complexity = low.data();
    a.filter_low_complexity = 1;
    launch_read_aux(a, nullptr);                                     // the product's launcher (8 warps per block)

    for (uint32_t r = 0; r < n; r++) {
        const size_t len = seqs[r].size();
        float mq_ref = 0; double cx_ref = 0;
        std::vector<uint8_t> q_ref(nmin[r] + 1);
        orc_read_aux(seqs[r].data(), have_q ? quals[r].data() : "", len, have_q ? len : 0, (int)l, hpc, pos.data() + moff[r],
                     nmin[r], &mq_ref, &cx_ref, q_ref.data());
        CHECK(same_double(cplx[r], cx_ref), "complexity read %u: %.17g vs %.17g", r, cplx[r], cx_ref);
        const bool low_ref = cx_ref > 5.0;
        CHECK((low[r] != 0) == low_ref, "low-complexity flag read %u", r);
        CHECK(nmin_io[r] == (low_ref ? 0u : nmin[r]), "filter read %u", r);
        if (have_q) {                                                    // mdbg_sketch_batch_q's finishing code
            long double errorSum = ldexpl((long double)shi[r], 64 - ERR_SHIFT) + ldexpl((long double)slo[r], -ERR_SHIFT);
            float meanReadError = errorSum / (uint64_t)len;
            const float mq = -10.0f * log10f(meanReadError);
            CHECK(same_float(mq, mq_ref), "mean quality read %u: %a vs %a", r, mq, mq_ref);
        }
        if (!low_ref)
            for (uint32_t j = 0; j < nmin[r]; j++)
                CHECK(out_qual[moff[r] + j] == q_ref[j], "min quality read %u minimizer %u: %u vs %u", r, j,
                      out_qual[moff[r] + j], q_ref[j]);
        *n_checked += nmin[r];
    }
}

int main() {
    std::mt19937_64 rng(77);
    auto rnd_seq = [&](size_t n, int kind) {
        std::string s(n, 'A');
        for (size_t i = 0; i < n; i++) s[i] = "ACGT"[rng() & 3];
        if (kind == 1) for (size_t i = 1; i < n; i++) if (rng() % 3 == 0) s[i] = s[i - 1];               // homopolymers
        if (kind == 2) { const size_t u = 1 + rng() % 4; for (size_t i = u; i < n; i++) s[i] = s[i - u]; } // low complexity
        if (kind == 3) for (size_t i = 0; i < n; i += 1 + rng() % 40) s[i] = (rng() & 1) ? s[i] : "ACGT"[rng() & 3];
        return s;
    };
    auto rnd_qual = [&](size_t n) {
        std::string q(n, '!');
        uint32_t level = 20 + rng() % 20;
        for (size_t i = 0; i < n; i++) {
            if (rng() % 50 == 0) level = rng() % 60;
            q[i] = (char)(33 + std::min<uint32_t>(93, level + rng() % 5));
        }
        return q;
    };
    std::vector<std::string> seqs, quals;
    for (size_t n : {0u, 1u, 2u, 3u, 14u, 15u, 16u, 64u, 65u, 66u, 67u, 97u, 98u, 99u, 130u, 511u, 512u, 513u, 1100u})
        seqs.push_back(rnd_seq(n, 0));
    for (int i = 0; i < 45; i++) seqs.push_back(rnd_seq(200 + rng() % 5000, i % 4));
    seqs.push_back(rnd_seq(30000, 1));
    seqs.push_back(std::string(2000, 'C'));
    for (auto& s : seqs) quals.push_back(rnd_qual(s.size()));
    uint64_t n_checked = 0;
    for (int hpc = 0; hpc < 2; hpc++) {
        run_case(seqs, quals, true, 15, 0.005f, hpc, &n_checked);
        run_case(seqs, quals, true, 15, 0.05f, hpc, &n_checked);
        run_case(seqs, quals, true, 11, 0.2f, hpc, &n_checked);
        run_case(seqs, quals, false, 15, 0.05f, hpc, &n_checked);       // FASTA: no qualities
    }
    printf("%llu per-minimizer qualities compared\n", (unsigned long long)n_checked);
    CHECK(n_checked > 50000, "coverage too small");
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
