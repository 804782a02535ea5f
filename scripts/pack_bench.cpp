// host packer throughput vs thread count (diagnostic): g++ -O2 -std=c++17 -pthread scripts/pack_bench.cpp metamdbg_b200/csrc/pack_host.o
#include "../metamdbg_b200/csrc/pack_host.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace mdbg;
int main(int argc, char** argv) {
    const size_t n_reads = 65536, len = 15000;
    std::vector<uint8_t> bases(n_reads * len);
    uint64_t x = 88172645463325252ull;
    for (auto& b : bases) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = "ACGT"[x & 3]; }
    std::vector<uint64_t> offs(n_reads + 1), pk(n_reads + 1);
    for (size_t r = 0; r <= n_reads; r++) { offs[r] = r * len; pk[r] = r * ((len + 15) / 16); }
    std::vector<uint32_t> pack(pk[n_reads] + 16);
    std::vector<uint64_t> src(n_reads);
    std::vector<uint8_t> asc(1 << 20);
    for (int t : {1, 4, 8, 16, 32, 64, 128}) {
        HostPool* p = host_pool_create(t);
        std::atomic<uint64_t> cur{0};
        host_pack_reads(p, bases.data(), offs.data(), 0, n_reads, pk.data(), pack.data(), src.data(), asc.data(), &cur);
        auto t0 = std::chrono::high_resolution_clock::now();
        const int reps = 5;
        for (int it = 0; it < reps; it++)
            for (int piece = 0; piece < 8; piece++)
                host_pack_reads(p, bases.data(), offs.data(), piece * 8192, (piece + 1) * 8192, pk.data(), pack.data(),
                                src.data(), asc.data(), &cur);
        double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        printf("threads %3d: %.1f GB/s\n", t, reps * (double)bases.size() / dt / 1e9);
        host_pool_destroy(p);
    }
}
