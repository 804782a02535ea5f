// k1_probe.cu -- stand-alone probe (no Python): device-generated HiFi-like reads, mdbg_ctx_autotune_sketch on them,
// one JSON line with the per-variant times and the device-side identity verdicts.
//   nvcc -O2 -std=c++17 -o scripts/k1_probe scripts/k1_probe.cu -Lmetamdbg_b200 -lmdbg_b200 -Xlinker -rpath='$ORIGIN/../metamdbg_b200'
//   scripts/k1_probe [n_reads=200000] [read_len=15000]
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

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 200000u;
    const uint64_t len = argc > 2 ? (uint64_t)atol(argv[2]) : 15000u;
    mdbg_params p{};
    p.minimizer_size = 15; p.density = 0.005f; p.use_hpc = 1;
    mdbg_ctx* ctx = nullptr;
    if (mdbg_ctx_create(0, &p, &ctx) != MDBG_OK) { printf("{\"error\": \"%s\"}\n", mdbg_last_error(nullptr)); return 1; }
    std::vector<uint64_t> off(n + 1), vs(n);
    std::vector<uint8_t> st(n);
    for (uint32_t r = 0; r <= n; r++) off[r] = r * len;
    for (uint32_t r = 0; r < n; r++) { vs[r] = mix64(r + 17) % 300000000ull; st[r] = (uint8_t)(mix64(r + 99) & 1); }
    uint8_t *d_bases, *d_st; uint64_t *d_off, *d_vs;
    if (cudaMalloc(&d_bases, n * len + 64) || cudaMalloc(&d_off, (n + 1) * 8) || cudaMalloc(&d_vs, n * 8) || cudaMalloc(&d_st, n)) {
        printf("{\"error\": \"cudaMalloc\"}\n"); return 1;
    }
    cudaMemcpy(d_off, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_vs, vs.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_st, st.data(), n, cudaMemcpyHostToDevice);
    if (mdbg_synth_fill_reads(ctx, d_bases, d_off, d_vs, d_st, n, 0, 20260924ull, (uint32_t)(0.001 * (1 << 24))) != MDBG_OK ||
        mdbg_ctx_synchronize(ctx) != MDBG_OK) { printf("{\"error\": \"%s\"}\n", mdbg_last_error(ctx)); return 1; }
    mdbg_autotune_out o{};
    if (mdbg_ctx_autotune_sketch(ctx, d_bases, d_off, n, n * len, &o) != MDBG_OK) { printf("{\"error\": \"%s\"}\n", mdbg_last_error(ctx)); return 1; }
    printf("{\"n_reads\": %u, \"read_len\": %llu, \"n_minimizers\": %llu, \"chosen\": %d, \"identical\": [%d, %d, %d], \"ms\": [%.3f, %.3f, %.3f], "
           "\"gbp_per_s\": [%.1f, %.1f, %.1f]", n, (unsigned long long)len, (unsigned long long)o.n_minimizers, o.chosen, o.identical[0],
           o.identical[1], o.identical[2], o.ms[0], o.ms[1], o.ms[2], n * len / (o.ms[0] * 1e6), n * len / (o.ms[1] * 1e6), n * len / (o.ms[2] * 1e6));
    // packed-resident input: the pack pass alone, then the packed kernel alone (mdbg_ctx_kernel_time_ms(0))
    {
        uint32_t* d_words; uint64_t* d_src;
        const uint64_t n_words = mdbg_pack_device_words(n * len, n);
        if (cudaMalloc(&d_words, n_words * 4) || cudaMalloc(&d_src, (uint64_t)n * 8)) { printf(", \"error\": \"cudaMalloc packed\"}\n"); return 1; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float pack_ms = 1e9f, k_ms = 1e9f, all_ms = 1e9f;
        mdbg_ctx_set_sketch_variant(ctx, 2);
        mdbg_ctx_enable_timing(ctx, 1);
        for (int rep = 0; rep < 3; rep++) {
            mdbg_ctx_synchronize(ctx);
            cudaEventRecord(e0, 0);                       // legacy default stream synchronises with the context's blocking... not: use sync
            cudaDeviceSynchronize();
            auto t0 = std::chrono::steady_clock::now();
            if (mdbg_pack_device(ctx, d_bases, d_off, n, n * len, d_words, d_src) != MDBG_OK) { printf(", \"error\": \"%s\"}\n", mdbg_last_error(ctx)); return 1; }
            mdbg_ctx_synchronize(ctx);
            const float t = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (t < pack_ms) pack_ms = t;
            t0 = std::chrono::steady_clock::now();
            mdbg_sketch_dev sd{};
            if (mdbg_sketch_batch_device_packed2(ctx, d_words, d_src, d_bases, d_off, n, n * len, 0, &sd) != MDBG_OK) { printf(", \"error\": \"%s\"}\n", mdbg_last_error(ctx)); return 1; }
            mdbg_ctx_synchronize(ctx);
            const float ta = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ta < all_ms) all_ms = ta;
            float km = 0; mdbg_ctx_kernel_time_ms(ctx, 0, &km);
            if (km < k_ms) k_ms = km;
            if (sd.n_minimizers != o.n_minimizers) { printf(", \"error\": \"packed2 gave %llu minimizers\"}\n", (unsigned long long)sd.n_minimizers); return 1; }
        }
        printf(", \"pack_device_ms_wall\": %.3f, \"packed_kernels_ms\": %.3f, \"packed2_call_ms_wall\": %.3f, \"packed_gbp_per_s\": %.1f",
               pack_ms, k_ms, all_ms, n * len / (k_ms * 1e6));
    }
    printf("}\n");
    mdbg_ctx_destroy(ctx);
    return 0;
}
