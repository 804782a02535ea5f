// pack_host.cpp -- see pack_host.hpp.  Compiled by g++ (not nvcc) so that the AVX2 path can use intrinsics.
#include "pack_host.hpp"

#include <immintrin.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mdbg {

// CPUs this process may actually use: hardware threads, capped by the cgroup CPU quota (containers), divided
// by the ranks sharing the node (torchrun exports LOCAL_WORLD_SIZE); at most 64.
int host_default_threads() {
    int n = (int)std::thread::hardware_concurrency();
    if (n <= 0) n = 4;
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char quota[32];
        long period = 0;
        if (fscanf(f, "%31s %ld", quota, &period) == 2 && strcmp(quota, "max") != 0 && period > 0) {
            const long q = atol(quota);
            if (q > 0) {
                const int cap = (int)((q + period - 1) / period);
                if (cap < n) n = cap;
            }
        }
        fclose(f);
    }
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    if (lws && atoi(lws) > 1) n /= atoi(lws);
    if (n > 64) n = 64;
    if (n < 1) n = 1;
    return n;
}

class HostPool {
public:
    explicit HostPool(int n) {
        if (n <= 0) {
            // one worker per hardware thread this process may fairly use: MDBG_HOST_THREADS overrides, else
            // hardware threads / ranks on this node (torchrun exports LOCAL_WORLD_SIZE), at most 64
            const char* env = getenv("MDBG_HOST_THREADS");
            if (env && atoi(env) > 0) n = atoi(env);
            else n = host_default_threads();
        }
        for (int i = 0; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // run f(thread_index) on every worker and wait
    void run(const std::function<void(int)>& f) {
        std::unique_lock<std::mutex> g(m_);
        job_ = &f;
        pending_ = (int)workers_.size();
        gen_++;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            if (job) (*job)(idx);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

HostPool* host_pool_create(int n_threads) { return new HostPool(n_threads); }
int host_pool_size(const HostPool* p) { return p->size(); }
void host_pool_destroy(HostPool* p) { delete p; }

namespace {

// scalar: returns false when a byte outside "ACGT" is met
inline bool pack_scalar(const uint8_t* s, uint64_t len, uint32_t* dst) {
    static const uint8_t expect[4] = {'A', 'C', 'T', 'G'};          // code (c >> 1) & 3 -> the only byte allowed
    for (uint64_t i = 0; i < len; i += 16) {
        uint32_t word = 0;
        const uint64_t n = len - i < 16 ? len - i : 16;
        for (uint64_t j = 0; j < n; j++) {
            const uint8_t c = s[i + j];
            const uint32_t code = (c >> 1) & 3;
            if (expect[code] != c) return false;
            word |= code << (2 * j);
        }
        dst[i >> 4] = word;
    }
    return true;
}

__attribute__((target("avx2"))) bool pack_avx2(const uint8_t* s, uint64_t len, uint32_t* dst) {
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                          'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w2 = _mm256_set1_epi16(0x0401);         // code[2i] + 4*code[2i+1]
    const __m256i w4 = _mm256_set1_epi32(0x00100001);     // lo + 16*hi
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                             0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t i = 0;
    __m256i bad = _mm256_setzero_si256();
    const __m256i pick = _mm256_setr_epi32(0, 4, 0, 4, 0, 4, 0, 4);
    for (; i + 64 <= len; i += 64) {
        const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        const __m256i c0 = _mm256_and_si256(_mm256_srli_epi16(v0, 1), three);
        const __m256i c1 = _mm256_and_si256(_mm256_srli_epi16(v1, 1), three);
        bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_xor_si256(_mm256_shuffle_epi8(lut, c0), v0),
                                                   _mm256_xor_si256(_mm256_shuffle_epi8(lut, c1), v1)));
        const __m256i q0 = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c0, w2), w4), gather);
        const __m256i q1 = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(c1, w2), w4), gather);
        const __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(q0, pick));   // words 0,1
        const __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(q1, pick));   // words 2,3
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (i >> 4)), _mm_unpacklo_epi64(lo, hi));
    }
    for (; i + 32 <= len; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i code = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i exp = _mm256_shuffle_epi8(lut, code);
        bad = _mm256_or_si256(bad, _mm256_xor_si256(exp, v));
        const __m256i p2 = _mm256_maddubs_epi16(code, w2);            // 16-bit lanes: 2 bases
        const __m256i p4 = _mm256_madd_epi16(p2, w4);                 // 32-bit lanes: 4 bases in the low byte
        const __m256i p16 = _mm256_shuffle_epi8(p4, gather);          // dword 0 of each 128-bit lane: 16 bases
        dst[i >> 4] = (uint32_t)_mm256_extract_epi32(p16, 0);
        dst[(i >> 4) + 1] = (uint32_t)_mm256_extract_epi32(p16, 4);
    }
    if (!_mm256_testz_si256(bad, bad)) return false;
    if (i < len) return pack_scalar(s + i, len - i, dst + (i >> 4));
    return true;
}

const bool g_have_avx2 = __builtin_cpu_supports("avx2");

inline bool pack_read(const uint8_t* s, uint64_t len, uint32_t* dst) {
    return g_have_avx2 ? pack_avx2(s, len, dst) : pack_scalar(s, len, dst);
}

}  // namespace

void host_pack_reads(HostPool* pool, const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1,
                     const uint64_t* pk_off, uint32_t* pack_out, uint64_t* src_out, uint8_t* asc_out,
                     std::atomic<uint64_t>* asc_cursor) {
    std::atomic<uint32_t> next{r0};
    const uint32_t grain = 32;
    const std::function<void(int)> body = [&](int) {
        for (;;) {
            const uint32_t b = next.fetch_add(grain);
            if (b >= r1) break;
            const uint32_t e = b + grain < r1 ? b + grain : r1;
            for (uint32_t r = b; r < e; r++) {
                const uint8_t* s = bases + offsets[r];
                const uint64_t len = offsets[r + 1] - offsets[r];
                if (pack_read(s, len, pack_out + pk_off[r])) {
                    src_out[r] = pk_off[r];
                } else {
                    const uint64_t slot = asc_cursor->fetch_add((len + 15) & ~uint64_t(15));
                    memcpy(asc_out + slot, s, len);
                    src_out[r] = (uint64_t(1) << 63) | slot;
                }
            }
        }
    };
    pool->run(body);
}

}  // namespace mdbg
