// e2e_probe.cu -- stand-alone probe (no Python) of the host-buffer path: device-generated reads are copied to pinned
// host memory once, then mdbg_sketch_batch (host ASCII in, minimizer CSR out on the host, store appended) is timed
// with the host wall clock.  One JSON line.  Environment toggles: MDBG_PACK_ISA=avx2, MDBG_PIECE_PIPELINE=0.
//   nvcc -O2 -std=c++17 -o scripts/e2e_probe scripts/e2e_probe.cu -Lmetamdbg_b200 -lmdbg_b200 -Xlinker -rpath='$ORIGIN/../metamdbg_b200'
//   scripts/e2e_probe [n_reads=262144] [read_len=15000] [iterations=4]
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/mdbg_b200.h"

static uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
#define DIE(...) do { printf("{\"error\": \""); printf(__VA_ARGS__); printf("\"}\n"); return 1; } while (0)

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 262144u;
    const uint64_t len = argc > 2 ? (uint64_t)atol(argv[2]) : 15000u;
    const int iters = argc > 3 ? atoi(argv[3]) : 4;
    mdbg_params p{};
    p.minimizer_size = 15; p.density = 0.005f; p.use_hpc = 1;
    mdbg_ctx* ctx = nullptr;
    if (mdbg_ctx_create(0, &p, &ctx) != MDBG_OK) DIE("%s", mdbg_last_error(nullptr));
    std::vector<uint64_t> off(n + 1), vs(n);
    std::vector<uint8_t> st(n);
    for (uint32_t r = 0; r <= n; r++) off[r] = r * len;
    for (uint32_t r = 0; r < n; r++) { vs[r] = mix64(r + 17) % 300000000ull; st[r] = (uint8_t)(mix64(r + 99) & 1); }
    uint8_t *d_bases, *d_st, *h_bases; uint64_t *d_off, *d_vs;
    if (cudaMalloc(&d_bases, n * len + 64) || cudaMalloc(&d_off, (n + 1) * 8) || cudaMalloc(&d_vs, n * 8) || cudaMalloc(&d_st, n) ||
        cudaMallocHost(&h_bases, n * len + 64)) DIE("allocation");
    cudaMemcpy(d_off, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_vs, vs.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_st, st.data(), n, cudaMemcpyHostToDevice);
    if (mdbg_synth_fill_reads(ctx, d_bases, d_off, d_vs, d_st, n, 0, 20260924ull, (uint32_t)(0.001 * (1 << 24))) != MDBG_OK ||
        mdbg_ctx_synchronize(ctx) != MDBG_OK) DIE("%s", mdbg_last_error(ctx));
    cudaMemcpy(h_bases, d_bases, n * len, cudaMemcpyDeviceToHost);
    mdbg_autotune_out tune{};
    if (mdbg_ctx_autotune_sketch(ctx, d_bases, d_off, n, n * len, &tune) != MDBG_OK) DIE("%s", mdbg_last_error(ctx));
    cudaFree(d_bases);
    double best = 1e30, sum = 0;
    uint64_t n_min = 0;
    mdbg_batch_info info{};
    for (int it = 0; it < iters + 1; it++) {                          // iteration 0 = warm-up (allocations)
        mdbg_store_clear(ctx);
        mdbg_sketch_out out{};
        const auto t0 = std::chrono::steady_clock::now();
        if (mdbg_sketch_batch(ctx, h_bases, off.data(), n, 1, &out) != MDBG_OK) DIE("%s", mdbg_last_error(ctx));
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        n_min = out.n_minimizers;
        if (out.min_offsets[n] != n_min) DIE("offsets[n] != n_minimizers");
        if (it) { sum += dt; if (dt < best) best = dt; }
    }
    mdbg_ctx_last_batch_info(ctx, &info);
    uint64_t h2d = 0, d2h = 0;
    mdbg_ctx_bytes_moved(ctx, &h2d, &d2h);
    printf("{\"n_reads\": %u, \"read_len\": %llu, \"n_minimizers\": %llu, \"matches_device_batch\": %s, \"sketch_variant\": %d, "
           "\"ms_best\": %.2f, \"ms_mean\": %.2f, \"gbp_per_s_best\": %.1f, \"gbp_per_s_mean\": %.1f, \"pieces\": %llu, "
           "\"pieces_pipelined\": %llu, \"growths\": %llu, \"packed\": %d, \"pack_gb_per_s\": %.1f, \"pack_isa\": \"%s\", "
           "\"host_threads\": %d, \"h2d_bytes_per_call\": %llu, \"d2h_bytes_per_call\": %llu}\n",
           n, (unsigned long long)len, (unsigned long long)n_min, n_min == tune.n_minimizers ? "true" : "false", tune.chosen,
           best * 1e3, sum / iters * 1e3, n * len / best / 1e9, n * len / (sum / iters) / 1e9,
           (unsigned long long)info.n_pieces, (unsigned long long)info.n_pieces_pipelined, (unsigned long long)info.n_buffer_growths,
           info.packed, info.pack_gb_per_s, info.pack_isa, info.host_threads, (unsigned long long)(h2d / (iters + 1)),
           (unsigned long long)(d2h / (iters + 1)));
    mdbg_ctx_destroy(ctx);
    return 0;
}
