// host_glue_test.cpp -- CPU checks of the pure parts of metamdbg_b200/host/mdbg_host.hpp (no GPU, no library
// call): the read_stats scalars against the reference's own Utils/Commons functions (oracle/_ref/libmdbg_ref.so),
// and the file helper's error behaviour.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "mdbg_host.hpp"

extern "C" {
uint64_t ref_compute_n50(const uint32_t* lengths, size_t n);
uint64_t ref_compute_mean_length(const uint32_t* lengths, size_t n);
int ref_compute_last_k(float density, size_t n50, size_t first_k, size_t max_k);
}

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 20) { printf("FAIL line %d: ", __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

int main() {
    std::mt19937_64 rng(12345);
    for (int trial = 0; trial < 20000; trial++) {
        const size_t n = 1 + rng() % 40;
        std::vector<uint32_t> len(n);
        const int kind = trial % 4;
        for (auto& x : len)
            x = kind == 0 ? (uint32_t)(rng() % 30000) : kind == 1 ? (uint32_t)(rng() % 4) : kind == 2 ? 15000u
                                                                                           : (uint32_t)(rng() >> 33);
        CHECK(mdbg_host::ReadDataWriter::computeN50(len) == ref_compute_n50(len.data(), n), "n50 trial %d", trial);
        CHECK(mdbg_host::ReadDataWriter::computeMeanLength(len) == ref_compute_mean_length(len.data(), n), "mean %d", trial);
    }
    CHECK(mdbg_host::ReadDataWriter::computeN50({}) == 0, "n50 of nothing");
    const float dens[] = {0.005f, 0.0025f, 0.01f, 0.3f};
    for (float d : dens)
        for (size_t n50 = 0; n50 < 200000; n50 += 37)
            CHECK((int)mdbg_host::computeLastK(d, n50, 4) == ref_compute_last_k(d, n50, 4, 0), "lastK d=%g n50=%zu", d, n50);

    // File: unopenable path and a device that refuses data both surface as exceptions
    bool threw = false;
    try { mdbg_host::File f("/nonexistent-dir/x.bin"); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw, "open of an impossible path must throw");
    threw = false;
    try {
        mdbg_host::File f("/dev/full");
        std::vector<char> big(1 << 20, 'x');
        f.put(big.data(), 1, big.size());
        f.close();
    } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw, "a failed write must throw");
    printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
