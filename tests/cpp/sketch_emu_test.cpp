// sketch_emu_test.cpp -- runs the REAL sketch / scan / compact kernel sources (metamdbg_b200/csrc/sketch.cu) on the CPU inside the warp emulator and compares every read's minimizers with the
// oracle.  This is a CPU regression test of the GPU kernel's logic (HPC fill, ring, roll, candidate list, flush,
// slot overflow + exact re-run, packed 2-bit input); it says nothing about speed.
//
// The kernels are started through the product's own launch_* functions (grid sizing included); `<<<...>>>` is
// rewritten to emu::launch by tests/_emu.py.
//
//   sketch_emu_test            exit 0 = identical to the oracle on every case
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "warp_emu.hpp"
#include SKETCH_SOURCE                       // the kernel source, `<<<...>>>` launches stripped

extern "C" {
#include "../../oracle/mdbg_oracle.h"
}

using namespace mdbg;

static int fails = 0;
static uint32_t g_variant = 0;             // arithmetic variant of the unrolled l = 15 block under test
static bool g_per_read = false;            // one launch per read: every read is the first one a (poisoned) ring sees
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 20) { printf("FAIL line %d: ", __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

struct Batch {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets{0};
    void add(const std::string& s) { bases.insert(bases.end(), s.begin(), s.end()); offsets.push_back(bases.size()); }
    uint32_t n() const { return (uint32_t)offsets.size() - 1; }
};

static uint64_t threshold_of(float density, uint32_t* none) {
    int nn = 0;
    const uint64_t t = orc_minimizer_threshold(density, &nn);
    *none = (uint32_t)nn;
    return t;
}

struct Result { std::vector<uint64_t> off; std::vector<uint32_t> min, pos; std::vector<uint8_t> dir; };

// what api.cu's sketch_internal does around the kernel, on host memory: padded slots, overflow -> exact re-run
static Result run_kernel(const Batch& b, uint32_t l, float density, int hpc, const std::vector<uint32_t>& blacklist,
                         bool packed, uint64_t* n_overflow_seen) {
    const uint32_t n = b.n();
    uint32_t none = 0;
    SketchArgs a{};
    std::vector<uint8_t> bases = b.bases;
    bases.resize(bases.size() + 64, 0);                       // the library allocates n_bases + 64
    // the device buffer is 16-byte aligned; keep that property
    std::vector<uint8_t> store(bases.size() + 16);
    uint8_t* aligned = store.data() + ((16 - ((uintptr_t)store.data() & 15)) & 15);
    memcpy(aligned, bases.data(), bases.size());
    a.bases = aligned;
    a.bases_end = aligned + b.bases.size();
    a.offsets = b.offsets.data();
    a.n_reads = n;
    a.read_begin = 0;
    a.read_end = n;
    a.variant = g_variant;
    a.l = l;
    a.hpc = hpc;
    a.threshold = threshold_of(density, &none);
    a.select_none = none;
    a.blacklist = blacklist.empty() ? nullptr : blacklist.data();
    a.n_blacklist = (uint32_t)blacklist.size();
    int shift = 0;
    while (shift < 8 && 1.0 / (double)(1u << (shift + 1)) >= 4.0 * (double)density) shift++;
    a.cap_shift = (uint32_t)shift;
    a.cap_const = 32;
    const uint64_t pad_cap = (b.bases.size() >> a.cap_shift) + (uint64_t)n * a.cap_const + 1;
    std::vector<uint32_t> pad_min(pad_cap), pad_pos(pad_cap), n_min(n + 1);
    std::vector<uint8_t> pad_dir(pad_cap);
    a.out_min = pad_min.data(); a.out_pos = pad_pos.data(); a.out_dir = pad_dir.data();
    a.n_min = n_min.data();
    uint32_t cursor = 0;
    unsigned long long n_overflow = 0;
    a.cursor = &cursor;
    a.n_overflow = &n_overflow;
    // packed input: every clean read as 2-bit words, dirty reads stay ASCII (what host_pack_reads produces)
    std::vector<uint32_t> words;
    std::vector<uint64_t> src(n);
    if (packed) {
        for (uint32_t r = 0; r < n; r++) {
            const uint64_t lo = b.offsets[r], hi = b.offsets[r + 1];
            bool clean = true;
            for (uint64_t i = lo; i < hi; i++) clean &= (b.bases[i] == 'A' || b.bases[i] == 'C' || b.bases[i] == 'G' || b.bases[i] == 'T');
            if (!clean) { src[r] = SRC_ASCII | lo; continue; }
            src[r] = words.size();
            for (uint64_t i = lo; i < hi; i += 16) {
                uint32_t w = 0;
                for (uint64_t j = i; j < hi && j < i + 16; j++) w |= (uint32_t)((b.bases[j] >> 1) & 3) << (2 * (j - i));
                words.push_back(w);
            }
        }
        words.resize(words.size() + 64, 0);
        a.read_src = src.data();
        a.packed = words.data();
    }
    // variant 2 (packed kernel + byte-ring kernel over the dirty list): an ASCII batch goes through the device-side
    // packer first (16-byte aligned read starts); the host-style packing above gives it arbitrary word alignment
    std::vector<uint32_t> dirty(n + 1);
    uint32_t dirty_n = 0, dirty_cur = 0;
    a.dirty_list = dirty.data(); a.dirty_count = &dirty_n; a.dirty_cursor = &dirty_cur;
    if (g_variant == 2 && !packed) {
        words.assign(pack_words_capacity(b.bases.size(), n) + 64, 0xA5A5A5A5u);
        PackArgsAscii pa{};
        pa.bases = a.bases; pa.bases_end = a.bases_end; pa.offsets = a.offsets;
        pa.read_begin = 0; pa.read_end = n; pa.packed = words.data(); pa.read_src = src.data();
        launch_pack_ascii(pa, 2, nullptr);
        for (uint32_t r = 0; r < n; r++) {
            bool clean = true;
            for (uint64_t i = b.offsets[r]; i < b.offsets[r + 1]; i++) clean &= (b.bases[i] == 'A' || b.bases[i] == 'C' || b.bases[i] == 'G' || b.bases[i] == 'T');
            CHECK(src[r] == (clean ? pack_word_offset(b.offsets[r], r) : (SRC_ASCII | b.offsets[r])), "pack kernel: read_src of read %u", r);
        }
        a.read_src = src.data();
        a.packed = words.data();
    }
    auto launch = [&]() {
        if (g_per_read) {
            for (uint32_t r = 0; r < n; r++) {
                cursor = 0; dirty_n = 0; dirty_cur = 0;
                a.read_begin = r; a.read_end = r + 1;
                launch_sketch(a, 2, nullptr);
            }
            a.read_begin = 0; a.read_end = n;
            return;
        }
        cursor = 0; dirty_n = 0; dirty_cur = 0;
        launch_sketch(a, /*sm_count=*/2, nullptr);            // 2 "SMs" x 2 CTAs x 8 warps, reads pulled dynamically
    };
    launch();
    Result res;
    res.off.assign(n + 1, 0);
    std::vector<uint64_t> scratch(scan_scratch_elems(n) + 1);
    launch_scan_u32_to_u64(n_min.data(), res.off.data(), n, scratch.data(), nullptr);      // the real 3-kernel scan
    for (uint32_t r = 0, acc = 0; r < n; r++) { CHECK(res.off[r] == acc, "scan at %u", r); acc += n_min[r]; }
    const uint64_t total = res.off[n];
    res.min.resize(total + 1); res.pos.resize(total + 1); res.dir.resize(total + 1);
    *n_overflow_seen = n_overflow;
    if (n_overflow == 0) {
        CompactArgs c{};                                      // the real compact_kernel
        c.base_offsets = b.offsets.data(); c.cap_shift = a.cap_shift; c.cap_const = a.cap_const;
        c.n_min = n_min.data(); c.tight_off = res.off.data();
        c.in_min = pad_min.data(); c.in_pos = pad_pos.data(); c.in_dir = pad_dir.data();
        c.out_min = res.min.data(); c.out_pos = res.pos.data(); c.out_dir = res.dir.data();
        c.n_reads = n;
        launch_compact(c, nullptr);
    } else {                                                  // exact re-run straight into the tight CSR
        a.exact_off = res.off.data();
        a.out_min = res.min.data(); a.out_pos = res.pos.data(); a.out_dir = res.dir.data();
        n_overflow = 0;
        launch();
        CHECK(n_overflow == 0, "overflow in the exact re-run");
    }
    res.min.resize(total); res.pos.resize(total); res.dir.resize(total);
    return res;
}

static void compare(const Batch& b, uint32_t l, float density, int hpc, const std::vector<uint32_t>& bl, bool packed,
                    const char* tag, uint64_t* totals) {
    uint64_t n_over = 0;
    const Result got = run_kernel(b, l, density, hpc, bl, packed, &n_over);
    std::vector<uint32_t> bls = bl;
    std::sort(bls.begin(), bls.end());
    for (uint32_t r = 0; r < b.n(); r++) {
        const uint64_t lo = b.offsets[r], len = b.offsets[r + 1] - lo;
        std::vector<uint32_t> m(len + 1), p(len + 1);
        std::vector<uint8_t> d(len + 1);
        const size_t nm = orc_sketch_read((const char*)b.bases.data() + lo, len, (int)l, density, hpc,
                                          bls.empty() ? nullptr : bls.data(), bls.size(), m.data(), p.data(), d.data(), len + 1);
        const uint64_t g0 = got.off[r], gn = got.off[r + 1] - g0;
        CHECK(gn == nm, "%s read %u: %llu minimizers, oracle %zu", tag, r, (unsigned long long)gn, nm);
        if (gn != nm) continue;
        for (size_t i = 0; i < nm; i++)
            CHECK(got.min[g0 + i] == m[i] && got.pos[g0 + i] == p[i] && got.dir[g0 + i] == d[i],
                  "%s read %u minimizer %zu", tag, r, i);
        totals[0] += nm;
    }
    totals[1] += n_over;
}

int main() {
    std::mt19937_64 rng(2024);
    auto rnd_read = [&](size_t n, int kind) {
        std::string s(n, 'A');
        static const char* alpha = "ACGT";
        for (size_t i = 0; i < n; i++) s[i] = alpha[rng() & 3];
        if (kind == 1) for (size_t i = 0; i < n; i++) if (rng() % 3 == 0 && i) s[i] = s[i - 1];           // homopolymers
        if (kind == 2) for (size_t i = 0; i < n; i += 1 + rng() % 97) s[i] = "NnacgtRY#"[rng() % 9];      // dirty
        if (kind == 3) { const size_t u = 1 + rng() % 6; for (size_t i = u; i < n; i++) s[i] = s[i - u]; } // tandem repeat
        return s;
    };
    Batch b;
    for (size_t n : {0u, 1u, 14u, 15u, 16u, 17u, 31u, 32u, 33u, 511u, 512u, 513u, 527u, 528u, 1039u, 2047u, 2048u, 2049u})
        b.add(rnd_read(n, 0));
    for (int i = 0; i < 40; i++) b.add(rnd_read(rng() % 6000, i % 4));
    b.add(rnd_read(40000, 0));                                   // many ring wraps
    b.add(rnd_read(9000, 1));
    b.add(std::string(3000, 'A'));
    b.add("####ACGTACGGTCA#ACGTTTGACCATGACCAGTAGGACCATTAGGGACCCATAGAC");
    // EncoderRLE keeps a '#' run that ENDS the read (its final push is unconditional): one more HPC base, so the last
    // selectable position moves by one.  Dense selections below make that position count.
    for (int i = 0; i < 24; i++) b.add(rnd_read(40 + rng() % 1500, i % 3 == 1 ? 2 : 0) + (i % 2 ? "#" : "###"));
    b.add("#"); b.add("###"); b.add("ACGTACGTTGCATGCA#"); b.add(rnd_read(527, 0) + "#"); b.add(rnd_read(511 + 15, 0) + "##");
    uint64_t totals[2] = {0, 0};
    std::vector<uint32_t> none;
    for (g_variant = 0; g_variant < (uint32_t)SKETCH_VARIANTS; g_variant++)
    for (int hpc = 0; hpc < 2; hpc++) {
        compare(b, 15, 0.005f, hpc, none, false, "l15 d0.005", totals);
        compare(b, 15, 0.05f, hpc, none, false, "l15 d0.05", totals);
        compare(b, 15, 0.05f, hpc, none, true, "l15 d0.05 packed", totals);
        compare(b, 15, 0.6f, hpc, none, false, "l15 d0.6 (slot overflow)", totals);
        compare(b, 11, 0.02f, hpc, none, false, "l11 generic", totals);
        compare(b, 16, 0.02f, hpc, none, true, "l16 generic packed", totals);
        compare(b, 15, 1.0f, hpc, none, false, "l15 d1.0 (exact path)", totals);
        compare(b, 15, 0.0f, hpc, none, false, "l15 d0 (select none)", totals);
    }
    for (g_variant = 0; g_variant < (uint32_t)SKETCH_VARIANTS; g_variant++) {   // blacklist = every third minimizer of a plain run
        uint64_t n_over = 0;
        const Result r0 = run_kernel(b, 15, 0.05f, 1, none, false, &n_over);
        std::vector<uint32_t> bl;
        for (size_t i = 0; i < r0.min.size(); i += 3) bl.push_back(r0.min[i]);
        std::sort(bl.begin(), bl.end());
        bl.erase(std::unique(bl.begin(), bl.end()), bl.end());
        compare(b, 15, 0.05f, 1, bl, false, "blacklist", totals);
    }
#ifdef MDBG_POISON_SMEM
    // Shared memory is not zero when a CTA starts: with the ring poisoned at kernel start and one launch per read,
    // every read meets bytes past `avail` that are neither codes nor flagged invalid.  Reads whose HPC length puts
    // `avail` exactly on a block boundary + l (the only non-final case) are included on purpose.
    {
        Batch pb;
        for (size_t n : {527u, 528u, 529u, 1039u, 1040u, 1041u, 1551u, 1552u, 2063u, 2064u, 3000u}) pb.add(rnd_read(n, 0));
        for (int i = 0; i < 300; i++) pb.add(rnd_read(500 + rng() % 2200, 0));
        g_per_read = true;
        for (g_variant = 0; g_variant < (uint32_t)SKETCH_VARIANTS; g_variant++)
            for (int hpc = 0; hpc < 2; hpc++) {
                compare(pb, 15, 0.02f, hpc, none, false, "poisoned ring", totals);
                compare(pb, 15, 0.02f, hpc, none, true, "poisoned ring packed", totals);
            }
        g_per_read = false;
    }
#endif
    printf("%llu minimizers compared, %llu slot overflows exercised\n", (unsigned long long)totals[0],
           (unsigned long long)totals[1]);
    CHECK(totals[0] > 40000 && totals[1] > 0, "coverage too small");
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
