// table_emu_test.cpp -- runs the REAL minimizer-space kernels (metamdbg_b200/csrc/purge.cu and kminmer.cu) in the
// warp emulator, sequenced as api.cu sequences them, and compares with the oracle:
//   purgePalindrome (flag / exact / compact), the density re-threshold, the k-min-mer count table (insert, stats,
//   emit) for k = 4 (specialised kernel) and generic k, rescue, the next-k pass with a previous-k table built on
//   the device or loaded from arrays, and the multi-GPU owner merge (pack by owner, insert-add of foreign vectors)
//   for 3 "ranks" living in one process.
// The two inline-PTX slot primitives (16-byte load, 128-bit CAS) are replaced by their plain C meaning by
// tests/_emu.py; everything else is the product source.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

#include "warp_emu.hpp"
#include PURGE_SOURCE
#include KMINMER_SOURCE

extern "C" {
#include "../../oracle/mdbg_oracle.h"
}

using namespace mdbg;

static int fails = 0;
static uint64_t cov_entries = 0, cov_rescued = 0, cov_nextk = 0, cov_merged = 0, cov_purged = 0, cov_density = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 20) { printf("FAIL line %d: ", __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

struct Store { std::vector<uint32_t> mins; std::vector<uint64_t> offs{0}; uint64_t n() const { return offs.size() - 1; } };
typedef std::pair<uint64_t, uint64_t> Key;                        // (h1, h2)
struct Entry { uint32_t ab; std::vector<uint32_t> vec; };
typedef std::map<Key, Entry> Table;

static uint64_t pow2ceil(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

static Store random_store(std::mt19937_64& rng, int n_reads, uint32_t alpha, int max_len) {
    Store s;
    for (int r = 0; r < n_reads; r++) {
        const int len = (int)(rng() % (uint64_t)(max_len + 1));
        for (int i = 0; i < len; i++) s.mins.push_back((uint32_t)(rng() % alpha) * 2654435761u + 17u * (alpha > 1000));
        s.offs.push_back(s.mins.size());
    }
    const uint64_t n0 = s.n();                                     // duplicates (forward and reversed) -> abundance
    for (uint64_t r = 0; r < n0; r += 2) {
        std::vector<uint32_t> v(s.mins.begin() + s.offs[r], s.mins.begin() + s.offs[r + 1]);
        if (rng() & 1) std::reverse(v.begin(), v.end());
        s.mins.insert(s.mins.end(), v.begin(), v.end());
        s.offs.push_back(s.mins.size());
    }
    s.mins.resize(s.mins.size() + 300, 0xDEADBEEFu);               // slack like the device allocation
    return s;
}

static Table oracle_count(const Store& s, int k, uint32_t min_ab, uint64_t* ninst = nullptr, uint64_t* ndist = nullptr) {
    uint32_t *v = nullptr, *a = nullptr; uint64_t* h = nullptr; uint64_t ni = 0, nd = 0;
    const size_t n = orc_count(s.mins.data(), s.offs.data(), s.n(), k, min_ab, &v, &h, &a, &ni, &nd);
    Table t;
    for (size_t i = 0; i < n; i++) t[Key(h[2 * i], h[2 * i + 1])] = Entry{a[i], std::vector<uint32_t>(v + i * k, v + (i + 1) * k)};
    orc_free(v); orc_free(h); orc_free(a);
    if (ninst) *ninst = ni;
    if (ndist) *ndist = nd;
    return t;
}

struct DevTable {
    std::vector<Slot> slots; uint64_t cap = 0; uint32_t k = 0;
    std::vector<uint32_t> foreign;
    void begin(uint64_t expect, uint32_t kk) { cap = pow2ceil(std::max<uint64_t>(expect, 512) * 2); slots.assign(cap, Slot{}); k = kk; foreign.clear(); }
};

static void insert_store(DevTable& t, const Store& s, uint64_t read_lo, uint64_t read_hi) {
    std::vector<uint8_t> rem(s.mins.size() + 1, 0);
    launch_fill_rem(s.offs.data(), read_lo, read_hi, rem.data(), nullptr);
    uint32_t full = 0;
    InsertArgs a{};
    a.mins = s.mins.data(); a.rem = rem.data(); a.g_lo = s.offs[read_lo]; a.g_hi = s.offs[read_hi]; a.k = t.k;
    static unsigned long long claims; claims = 0;
    a.table = t.slots.data(); a.mask = t.cap - 1; a.full_flag = &full; a.claims = &claims; a.claim_limit = t.cap;
    launch_insert(a, nullptr);
    CHECK(full == 0, "table full");
}

static Table emit(const DevTable& t, const Store& s, uint32_t min_count, TableStats* st_out = nullptr) {
    TableStats st{};
    launch_table_stats(t.slots.data(), t.cap, min_count, &st, nullptr);
    std::vector<uint64_t> h(2 * st.n_entries + 2); std::vector<uint32_t> ab(st.n_entries + 1), vecs((st.n_entries + 1) * t.k);
    unsigned long long cursor = 0;
    EmitArgs e{};
    e.table = t.slots.data(); e.capacity = t.cap; e.min_count = min_count; e.k = t.k; e.mins = s.mins.data();
    e.foreign_vecs = t.foreign.empty() ? nullptr : t.foreign.data();
    e.out_hashes = h.data(); e.out_abund = ab.data(); e.out_vecs = vecs.data(); e.cursor = &cursor;
    launch_table_emit(e, nullptr);
    CHECK(cursor == st.n_entries, "emit wrote %llu entries, stats say %llu", cursor, st.n_entries);
    Table out;
    uint64_t cs = 0;
    for (uint64_t i = 0; i < cursor; i++) {                       // device order: lo = h2, hi = h1
        out[Key(h[2 * i + 1], h[2 * i])] = Entry{ab[i], std::vector<uint32_t>(vecs.begin() + i * t.k, vecs.begin() + (i + 1) * t.k)};
        cs += (uint64_t)ab[i] * h[2 * i];
    }
    CHECK(out.size() == cursor, "duplicate keys emitted");
    CHECK(cs == st.checksum, "checksum");
    if (st_out) *st_out = st;
    return out;
}

static void expect_equal(const Table& got, const Table& want, const char* tag) {
    cov_entries += want.size();
    CHECK(got.size() == want.size(), "%s: %zu entries, oracle %zu", tag, got.size(), want.size());
    for (auto& kv : want) {
        auto it = got.find(kv.first);
        CHECK(it != got.end(), "%s: key missing", tag);
        if (it == got.end()) continue;
        CHECK(it->second.ab == kv.second.ab, "%s: abundance %u vs %u", tag, it->second.ab, kv.second.ab);
        CHECK(it->second.vec == kv.second.vec, "%s: vector differs", tag);
    }
}

// ---- purge + density, sequenced like mdbg_purge_palindromes / mdbg_store_apply_density ------------------------
static Store compact_with(const Store& s, const std::vector<uint8_t>& keep, const std::vector<uint32_t>& cnt) {
    Store out;
    out.offs.assign(s.n() + 1, 0);
    for (uint64_t r = 0; r < s.n(); r++) out.offs[r + 1] = out.offs[r] + cnt[r];
    out.mins.assign(out.offs[s.n()] + 300, 0xDEADBEEFu);
    // rem[] of the compacted store comes out of the compaction as a by-product: it must equal fill_rem's on the new store
    std::vector<uint8_t> rem_by(out.mins.size() + 1, 0xEE), rem_fill(out.mins.size() + 1, 0xEE);
    launch_purge_compact(s.mins.data(), s.offs.data(), out.offs.data(), keep.data(), s.n(), out.mins.data(), nullptr, rem_by.data());
    launch_fill_rem(out.offs.data(), 0, s.n(), rem_fill.data(), nullptr);
    for (uint64_t g = 0; g < out.offs[s.n()]; g++) {
        uint64_t r = 0;
        while (out.offs[r + 1] <= g) r++;
        const uint64_t left = out.offs[r + 1] - g;
        CHECK(rem_fill[g] == (uint8_t)(left > 255 ? 255 : left), "fill_rem[%llu]", (unsigned long long)g);
        CHECK(rem_by[g] == rem_fill[g], "rem by-product[%llu]", (unsigned long long)g);
    }
    for (uint64_t g = out.offs[s.n()]; g < rem_by.size(); g++) CHECK(rem_by[g] == 0xEE && rem_fill[g] == 0xEE, "rem written past the store end at %llu", (unsigned long long)g);
    return out;
}

static void test_purge_and_density(std::mt19937_64& rng) {
    for (uint32_t alpha : {2u, 3u, 6u, 50000u}) {
        const Store s = random_store(rng, 150, alpha, 60);
        const uint32_t first_k = 4, last_k = 5 + (uint32_t)(rng() % 30);
        std::vector<uint8_t> flags(s.n() + 1, 0), keep(s.mins.size() + 1, 1);
        std::vector<uint32_t> cnt(s.n() + 1, 0);
        unsigned long long n_flagged = 0, n_changed = 0;
        launch_purge_flag(s.mins.data(), s.offs.data(), s.n(), first_k, last_k, flags.data(), &n_flagged, nullptr);
        launch_purge_exact(s.mins.data(), s.offs.data(), s.n(), flags.data(), first_k, last_k, keep.data(), cnt.data(), &n_changed, nullptr);
        const Store p = compact_with(s, keep, cnt);
        uint64_t changed_ref = 0;
        for (uint64_t r = 0; r < s.n(); r++) {
            const uint64_t lo = s.offs[r], n = s.offs[r + 1] - lo;
            std::vector<uint32_t> out(n + 1);
            const size_t m = orc_purge_palindrome(s.mins.data() + lo, n, first_k, last_k, out.data(), nullptr);
            changed_ref += (m != n);
            cov_purged += n - m;
            CHECK(p.offs[r + 1] - p.offs[r] == m, "purge alpha %u read %llu: %llu vs %zu", alpha, (unsigned long long)r,
                  (unsigned long long)(p.offs[r + 1] - p.offs[r]), m);
            if (p.offs[r + 1] - p.offs[r] == m) CHECK(!memcmp(p.mins.data() + p.offs[r], out.data(), m * 4), "purge content read %llu", (unsigned long long)r);
        }
        CHECK(n_changed == changed_ref, "purge n_changed %llu vs %llu", n_changed, (unsigned long long)changed_ref);
        // density re-threshold (Utils::applyDensityThreshold) on full-range values
        Store d = random_store(rng, 80, 4000000000u, 70);
        for (float dens : {0.0025f, 0.3f}) {
            int none = 0;
            const uint64_t thr = orc_minimizer_threshold(dens, &none);
            std::vector<uint8_t> k2(d.mins.size() + 1, 1); std::vector<uint32_t> c2(d.n() + 1, 0);
            unsigned long long ch = 0;
            launch_density_filter(d.mins.data(), d.offs.data(), d.n(), thr, (uint32_t)none, k2.data(), c2.data(), &ch, nullptr);
            const Store f = compact_with(d, k2, c2);
            for (uint64_t r = 0; r < d.n(); r++) {
                const uint64_t lo = d.offs[r], n = d.offs[r + 1] - lo;
                std::vector<uint32_t> out(n + 1);
                const size_t m = orc_apply_density(d.mins.data() + lo, n, dens, out.data());
                cov_density += n - m;
                CHECK(f.offs[r + 1] - f.offs[r] == m && !memcmp(f.mins.data() + f.offs[r], out.data(), m * 4), "density read %llu", (unsigned long long)r);
            }
        }
    }
}

// ---- count / rescue / next-k / merge ------------------------------------------------------------------------------
static void test_tables(std::mt19937_64& rng) {
    for (uint32_t alpha : {2u, 5u, 40u, 3000000u})
        for (int k : {4, 2, 5, 7, 21}) {
            const Store s = random_store(rng, 120, alpha, 70);
            // count, in two batches like mdbg_count_add_store called twice
            DevTable t;
            t.begin(s.offs[s.n()], (uint32_t)k);
            insert_store(t, s, 0, s.n() / 3);
            insert_store(t, s, s.n() / 3, s.n());
            uint64_t ninst = 0, ndist = 0;
            const Table want2 = oracle_count(s, k, 2, &ninst, &ndist);
            TableStats st{};
            expect_equal(emit(t, s, 2, &st), want2, "count>=2");
            CHECK(st.n_instances == ninst && st.n_distinct == ndist, "instances/distinct k=%d", k);
            expect_equal(emit(t, s, 3), oracle_count(s, k, 3), "count>=3");

            // rescue (default mode): flags abundance-1 entries of low-coverage reads
            {
                std::vector<uint64_t> sh; std::vector<uint32_t> sa;
                for (auto& kv : want2) { sh.push_back(kv.first.first); sh.push_back(kv.first.second); sa.push_back(kv.second.ab); }
                uint32_t* rv = nullptr; uint64_t* rh = nullptr; uint64_t nrr = 0;
                const size_t nr = orc_rescue(s.mins.data(), s.offs.data(), s.n(), k, sh.data(), sa.data(), sa.size(), &rv, &rh, &nrr);
                Table want = want2;
                for (size_t i = 0; i < nr; i++) want[Key(rh[2 * i], rh[2 * i + 1])] = Entry{1, std::vector<uint32_t>(rv + i * k, rv + (i + 1) * k)};
                orc_free(rv); orc_free(rh);
                DevTable t2 = t;
                unsigned long long n_rescued_reads = 0;
                RescueArgs ra{};
                ra.mins = s.mins.data(); ra.offs = s.offs.data(); ra.n_reads = s.n(); ra.k = (uint32_t)k;
                ra.table = t2.slots.data(); ra.mask = t2.cap - 1; ra.n_reads_rescued = &n_rescued_reads;
                launch_rescue(ra, nullptr);
                TableStats rs{};
                expect_equal(emit(t2, s, 2, &rs), want, "rescue");
                CHECK(n_rescued_reads == nrr, "rescued reads %llu vs %llu", n_rescued_reads, (unsigned long long)nrr);
                CHECK(rs.n_rescued == want.size() - want2.size(), "n_rescued");
                cov_rescued += rs.n_rescued;
            }

            // next k: previous table from the device table, and loaded from arrays (+ a patch on top)
            if (k >= 2 && k < 21) {
                std::vector<uint64_t> ph; std::vector<uint32_t> pa;
                for (auto& kv : want2) { ph.push_back(kv.first.first); ph.push_back(kv.first.second); pa.push_back(kv.second.ab); }
                uint32_t *nv = nullptr, *na = nullptr; uint64_t* nh = nullptr;
                const size_t nn = orc_next_k(s.mins.data(), s.offs.data(), s.n(), k + 1, ph.data(), pa.data(), pa.size(), &nv, &nh, &na);
                Table want;
                for (size_t i = 0; i < nn; i++) want[Key(nh[2 * i], nh[2 * i + 1])] = Entry{na[i], std::vector<uint32_t>(nv + i * (k + 1), nv + (i + 1) * (k + 1))};
                orc_free(nv); orc_free(nh); orc_free(na);
                for (int mode = 0; mode < 3; mode++) {            // 2: the count table itself as previous-k table (lookup-time filter)
                    const uint64_t pcap = pow2ceil(std::max<uint64_t>(want2.size(), 512) * 2);
                    std::vector<Slot> prev(pcap, Slot{});
                    uint32_t full = 0;
                    if (mode == 0) {
                        PrevFromTableArgs pf{};
                        pf.table = t.slots.data(); pf.capacity = t.cap; pf.min_count = 2; pf.prev = prev.data(); pf.prev_mask = pcap - 1; pf.full_flag = &full;
                        launch_prev_from_table(pf, nullptr);
                    } else if (mode == 1) {                       // device layout of hashes: lo = h2, hi = h1
                        std::vector<uint64_t> lohi; for (size_t i = 0; i < pa.size(); i++) { lohi.push_back(ph[2 * i + 1]); lohi.push_back(ph[2 * i]); }
                        PrevLoadArgs pl{};
                        pl.hashes = lohi.data(); pl.abund = pa.data(); pl.n = pa.size(); pl.prev = prev.data(); pl.prev_mask = pcap - 1; pl.full_flag = &full;
                        launch_prev_load(pl, nullptr);
                    }
                    CHECK(full == 0, "prev table full");
                    DevTable t3;
                    t3.begin(s.offs[s.n()], (uint32_t)k + 1);
                    std::vector<uint8_t> rem(s.mins.size() + 1, 0);
                    launch_fill_rem(s.offs.data(), 0, s.n(), rem.data(), nullptr);
                    NextKArgs nk{};
                    nk.mins = s.mins.data(); nk.rem = rem.data(); nk.g_lo = 0; nk.g_hi = s.offs[s.n()]; nk.k = (uint32_t)k + 1;
                    nk.prev = prev.data(); nk.prev_mask = pcap - 1; nk.table = t3.slots.data(); nk.mask = t3.cap - 1; nk.full_flag = &full;
                    unsigned long long nk_claims = 0; nk.claims = &nk_claims; nk.claim_limit = t3.cap; nk.prev_min_count = 0;
                    if (mode == 2) { nk.prev = t.slots.data(); nk.prev_mask = t.cap - 1; nk.prev_min_count = 2; }
                    launch_next_k(nk, nullptr);
                    CHECK(full == 0, "next-k table full");
                    expect_equal(emit(t3, s, 2), want, mode == 2 ? "next-k (count table as prev)" : mode ? "next-k (loaded prev)" : "next-k (device prev)");
                    cov_nextk += want.size();
                }
            }

            // multi-GPU merge, 3 ranks in one process: shard reads, count locally, pack by owner, insert-add
            {
                const uint32_t R = 3;
                std::vector<DevTable> loc(R);
                std::vector<std::vector<uint32_t>> send_vecs(R), send_cnt(R);
                std::vector<std::vector<uint64_t>> base(R, std::vector<uint64_t>(R, 0)), cntm(R, std::vector<uint64_t>(R, 0));
                for (uint32_t r = 0; r < R; r++) {
                    loc[r].begin(s.offs[s.n()], (uint32_t)k);
                    insert_store(loc[r], s, s.n() * r / R, s.n() * (r + 1) / R);
                    std::vector<unsigned long long> bc(R, 0);
                    PackArgs p{};
                    p.table = loc[r].slots.data(); p.capacity = loc[r].cap; p.k = (uint32_t)k; p.n_ranks = R; p.mins = s.mins.data();
                    p.bucket_count = bc.data(); p.pass = 1;
                    launch_table_pack(p, nullptr);
                    uint64_t tot = 0;
                    for (uint32_t d = 0; d < R; d++) { cntm[r][d] = bc[d]; base[r][d] = tot; tot += bc[d]; }
                    send_vecs[r].assign((tot + 1) * k, 0); send_cnt[r].assign(tot + 1, 0);
                    std::fill(bc.begin(), bc.end(), 0);
                    p.bucket_base = base[r].data(); p.out_vecs = send_vecs[r].data(); p.out_counts = send_cnt[r].data(); p.pass = 2;
                    launch_table_pack(p, nullptr);
                    for (uint32_t d = 0; d < R; d++) CHECK(bc[d] == cntm[r][d], "pack pass 2 count");
                }
                Table merged;
                uint64_t merged_instances = 0;
                for (uint32_t d = 0; d < R; d++) {                 // receive side of rank d
                    std::vector<uint32_t> rv, rc;
                    for (uint32_t r = 0; r < R; r++) {
                        rv.insert(rv.end(), send_vecs[r].begin() + base[r][d] * k, send_vecs[r].begin() + (base[r][d] + cntm[r][d]) * k);
                        rc.insert(rc.end(), send_cnt[r].begin() + base[r][d], send_cnt[r].begin() + base[r][d] + cntm[r][d]);
                    }
                    DevTable own;
                    own.begin(rc.size(), (uint32_t)k);
                    own.foreign = rv; own.foreign.resize(rv.size() + 64, 0);
                    uint32_t full = 0;
                    InsertVecArgs iv{};
                    iv.vecs = own.foreign.data(); iv.counts = rc.data(); iv.n = rc.size(); iv.foreign_base = 0; iv.k = (uint32_t)k;
                    iv.table = own.slots.data(); iv.mask = own.cap - 1; iv.full_flag = &full;
                    launch_insert_vecs(iv, nullptr);
                    CHECK(full == 0, "merged table full");
                    TableStats ms{};
                    const Table part = emit(own, s, 2, &ms);
                    merged_instances += ms.n_instances;
                    for (auto& kv : part) {
                        CHECK(owner_of(kv.first.first, R) == d, "key on the wrong owner");
                        CHECK(!merged.count(kv.first), "key on two owners");
                        merged[kv.first] = kv.second;
                    }
                }
                expect_equal(merged, want2, "3-rank merge");
                cov_merged += merged.size();
                CHECK(merged_instances == ninst, "occurrence conservation across ranks");
            }
        }
}

int main() {
    std::mt19937_64 rng(4242);
    test_purge_and_density(rng);
    test_tables(rng);
    printf("compared: %llu table entries, %llu rescued, %llu next-k, %llu merged; %llu minimizers purged, %llu dropped by density\n",
           (unsigned long long)cov_entries, (unsigned long long)cov_rescued, (unsigned long long)cov_nextk,
           (unsigned long long)cov_merged, (unsigned long long)cov_purged, (unsigned long long)cov_density);
    CHECK(cov_entries > 20000 && cov_rescued > 100 && cov_nextk > 1000 && cov_merged > 5000 && cov_purged > 100 && cov_density > 1000,
          "coverage too small");
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
