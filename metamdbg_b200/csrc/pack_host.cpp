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
            gen_.fetch_add(1);
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // run f(thread_index) on every worker and wait.  A host batch is packed piece by piece with only a copy and a
    // kernel launch enqueued in between, so both sides first spin for a moment (~100 us) before they sleep on the
    // condition variable: a futex round trip per worker per piece otherwise costs ~0.2 ms of a ~1.5 ms piece.
    void run(const std::function<void(int)>& f) {
        start(f);
        wait();
    }
    // start() hands f to every worker and returns; wait() blocks until all of them are done with it.  One job at a
    // time; f must stay alive until wait() returns.
    void start(const std::function<void(int)>& f) {
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &f;
            pending_.store((int)workers_.size(), std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
    }
    void wait() {
        for (int spin = 0; spin < kSpin && pending_.load(std::memory_order_acquire) != 0; spin++) _mm_pause();
        if (pending_.load(std::memory_order_acquire) != 0) {
            std::unique_lock<std::mutex> g(m_);
            done_.wait(g, [this] { return pending_.load(std::memory_order_acquire) == 0; });
        }
        job_ = nullptr;
    }

private:
    static constexpr int kSpin = 4000;
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            for (int spin = 0; spin < kSpin && gen_.load(std::memory_order_acquire) == seen; spin++) _mm_pause();
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                seen = gen_.load(std::memory_order_acquire);
                if (stop_) return;
                job = job_;
            }
            if (job) (*job)(idx);
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> g(m_);                    // pairs with the waiter's predicate check
                done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
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
        _mm_prefetch(reinterpret_cast<const char*>(s + i + 2048), _MM_HINT_T0);
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

// AVX-512 (F/BW/VL/VBMI -- every Xeon since Ice Lake, Zen 4+): 256 bases -> one 64-byte line per iteration.
// The input is streamed from DRAM exactly once, so the loop (a) software-prefetches 4 KB ahead -- one core's
// hardware prefetcher alone sustains only ~6 GB/s here, the explicit prefetch ~10 GB/s -- and (b) writes the
// packed words with non-temporal stores once the destination is 64-byte aligned (no read-for-ownership of lines
// the DMA engine is about to read anyway).  Head and tail use masked loads/stores: nothing outside
// [s, s+len) is read and nothing outside the read's own ceil(len/16) words is written.
#define MDBG_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))

struct Pack512 {
    __m512i three, lut, w2, w4, ix;
    MDBG_AVX512 Pack512() {
        three = _mm512_set1_epi8(3);
        lut = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
        w2 = _mm512_set1_epi16(0x0401);                  // code[2i] + 4*code[2i+1]
        w4 = _mm512_set1_epi32(0x00100001);              // lo + 16*hi
        alignas(64) uint8_t idx[64];
        // byte 4j of the first operand for j < 16, byte 4(j-16) of the second for 16 <= j < 32
        for (int j = 0; j < 64; j++) idx[j] = (uint8_t)(((j & 15) * 4) | (((j >> 4) & 1) ? 64 : 0));
        ix = _mm512_load_si512(idx);
    }
    // 64 bases -> 16 dwords whose low bytes hold 4 bases each; bad |= lanes that are not one of "ACGT"
    MDBG_AVX512 inline __m512i quads(__m512i v, __mmask64 live, __mmask64& bad) const {
        const __m512i c = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);
        bad |= _mm512_mask_cmpneq_epi8_mask(live, _mm512_shuffle_epi8(lut, c), v);
        return _mm512_madd_epi16(_mm512_maddubs_epi16(c, w2), w4);
    }
    // 64 bases -> 4 packed words
    MDBG_AVX512 inline __m128i words4(__m512i q) const { return _mm512_cvtepi32_epi8(q); }
};

MDBG_AVX512 bool pack_avx512(const uint8_t* s, uint64_t len, uint32_t* dst) {
    static const Pack512 K;
    const __mmask64 all = ~__mmask64(0);
    __mmask64 bad = 0;
    uint64_t i = 0;
    // head: bring dst to a 64-byte boundary (dst is only word-aligned: reads are packed back to back)
    uint64_t head_words = ((64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63) >> 2;
    if (head_words * 16 > len) head_words = (len + 15) >> 4;            // short read: everything is "head"
    while (head_words) {
        const uint64_t nw = head_words < 4 ? head_words : 4;
        const uint64_t left = len - i, nb = left < nw * 16 ? left : nw * 16;
        const __mmask64 live = nb >= 64 ? all : ((__mmask64(1) << nb) - 1);
        const __m512i v = _mm512_maskz_loadu_epi8(live, s + i);
        _mm_mask_storeu_epi32(dst + (i >> 4), (__mmask8)((1u << nw) - 1), K.words4(K.quads(v, live, bad)));
        i += nw * 16;
        head_words -= nw;
    }
    if (i >= len) return bad == 0;
    for (; i + 256 <= len; i += 256) {
        _mm_prefetch(reinterpret_cast<const char*>(s + i + 4096), _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char*>(s + i + 4096 + 64), _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char*>(s + i + 4096 + 128), _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char*>(s + i + 4096 + 192), _MM_HINT_T0);
        const __m512i q0 = K.quads(_mm512_loadu_si512(s + i), all, bad);
        const __m512i q1 = K.quads(_mm512_loadu_si512(s + i + 64), all, bad);
        const __m512i q2 = K.quads(_mm512_loadu_si512(s + i + 128), all, bad);
        const __m512i q3 = K.quads(_mm512_loadu_si512(s + i + 192), all, bad);
        const __m512i a = _mm512_permutex2var_epi8(q0, K.ix, q1);      // low 32 bytes: words of q0, q1
        const __m512i b = _mm512_permutex2var_epi8(q2, K.ix, q3);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + (i >> 4)), _mm512_inserti64x4(a, _mm512_castsi512_si256(b), 1));
    }
    for (; i < len; i += 64) {
        const uint64_t left = len - i;
        const __mmask64 live = left >= 64 ? all : ((__mmask64(1) << left) - 1);
        const uint64_t nw = left >= 64 ? 4 : (left + 15) >> 4;
        const __m512i v = _mm512_maskz_loadu_epi8(live, s + i);
        _mm_mask_storeu_epi32(dst + (i >> 4), (__mmask8)((1u << nw) - 1), K.words4(K.quads(v, live, bad)));
    }
    _mm_sfence();                                                      // the streamed lines are read by DMA next
    return bad == 0;
}

// 0 scalar, 1 AVX2, 2 AVX-512; MDBG_PACK_ISA=scalar|avx2|avx512 lowers it (tests)
int detect_isa() {
    int isa = 0;
    if (__builtin_cpu_supports("avx2")) isa = 1;
    if (isa == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
        __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vbmi"))
        isa = 2;
    if (const char* e = getenv("MDBG_PACK_ISA")) {
        const int want = !strcmp(e, "scalar") ? 0 : !strcmp(e, "avx2") ? 1 : 2;
        if (want < isa) isa = want;
    }
    return isa;
}
const int g_isa = detect_isa();

inline bool pack_read(const uint8_t* s, uint64_t len, uint32_t* dst) {
    if (g_isa == 2) return pack_avx512(s, len, dst);
    return g_isa == 1 ? pack_avx2(s, len, dst) : pack_scalar(s, len, dst);
}

}  // namespace

const char* host_pack_isa() { return g_isa == 2 ? "avx512" : g_isa == 1 ? "avx2" : "scalar"; }

bool host_pack_one(const uint8_t* bases, uint64_t len, uint32_t* words_out) { return pack_read(bases, len, words_out); }

// one piece being packed: lives on the heap so that the workers can go on while the caller enqueues the previous
// piece's copies and launches
struct PackJob {
    std::atomic<uint32_t> next;
    std::function<void(int)> body;
};

PackJob* host_pack_start(HostPool* pool, const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1,
                         const uint64_t* pk_off, uint32_t* pack_out, uint64_t* src_out, uint8_t* asc_out,
                         std::atomic<uint64_t>* asc_cursor) {
    PackJob* job = new PackJob();
    job->next.store(r0);
    job->body = [=](int) {
        const uint32_t grain = 32;
        for (;;) {
            const uint32_t b = job->next.fetch_add(grain);
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
    pool->start(job->body);
    return job;
}

void host_pack_wait(HostPool* pool, PackJob* job) {
    if (!job) return;
    pool->wait();
    delete job;
}

void host_pack_reads(HostPool* pool, const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1,
                     const uint64_t* pk_off, uint32_t* pack_out, uint64_t* src_out, uint8_t* asc_out,
                     std::atomic<uint64_t>* asc_cursor) {
    host_pack_wait(pool, host_pack_start(pool, bases, offsets, r0, r1, pk_off, pack_out, src_out, asc_out, asc_cursor));
}

}  // namespace mdbg
