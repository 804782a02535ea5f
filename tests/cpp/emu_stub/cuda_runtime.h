// Stand-in for <cuda_runtime.h> when the product's .cu files are compiled by g++ for the CPU emulator (TESTS ONLY).
// Kernels get warp_emu.hpp; the host side of the C ABI (api.cu) gets a synchronous single-"device" runtime: device
// memory is host memory, every asynchronous operation completes before the call returns (a valid serialisation:
// the library only ever waits on work it has already enqueued), streams and events are tokens.
//
// With -DEMU_DEFERRED the runtime picks the OPPOSITE serialisation (warp_emu.hpp, emu::Stream): asynchronous work
// stays queued on its stream until something really waits for it (stream / event / device synchronisation, an event
// another stream waits on, cudaFree).  Copies follow CUDA's rules for host memory: an asynchronous copy FROM pageable
// memory is staged when it is issued, one INTO pageable memory completes before the call returns, pinned memory
// (cudaMallocHost) is read and written when the copy executes.  A host read of a pinned result, or a reuse of a
// pinned / device buffer, that is not ordered behind the work by a real synchronisation then sees poisoned or stale
// bytes and the parity tests fail -- the class of error eager execution hides.
#pragma once
#include "../warp_emu.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };
struct EmuEvent { std::chrono::steady_clock::time_point t; emu::Stream* s = nullptr; uint64_t seq = 0; };
namespace emu {
inline std::mutex& pinned_mu() { static std::mutex m; return m; }
inline std::map<uintptr_t, size_t>& pinned() { static std::map<uintptr_t, size_t> m; return m; }
inline bool is_pinned(const void* p) {
    std::lock_guard<std::mutex> g(pinned_mu());
    auto it = pinned().upper_bound((uintptr_t)p);
    if (it == pinned().begin()) return false;
    --it;
    return (uintptr_t)p < it->first + it->second;
}
}  // namespace emu
typedef EmuEvent* cudaEvent_t;

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "emulated runtime error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p); p->major = 10; p->minor = 0; p->multiProcessorCount = 2; strcpy(p->name, "warp emulator"); return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize() { emu::drain_all(); return cudaSuccess; }
// 256-byte aligned like the driver's allocations (the kernels rely on 16-byte aligned buffers)
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, n);                                      // device memory is NOT zero on allocation
    *p = (T*)q; return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) {
    const cudaError_t e = cudaMalloc(p, n);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(emu::pinned_mu()); emu::pinned()[(uintptr_t)*p] = n ? n : 1; }
    return e;
}
// cudaFree / cudaFreeHost synchronise the device before they release memory
static inline cudaError_t cudaFree(void* p) { emu::drain_all(); free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) {
    emu::drain_all();
    { std::lock_guard<std::mutex> g(emu::pinned_mu()); emu::pinned().erase((uintptr_t)p); }
    free(p);
    return cudaSuccess;
}
// synchronous calls on the legacy default stream do not wait for non-blocking streams: they run at once
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t st) {
    emu::Stream* s = (emu::Stream*)st;
    if (s) emu::drain(s, s->submitted);
    return cudaSuccess;
}
// hook for tests/cpp/fake_nccl.cpp (found with dlsym): a collective first drains the rank's stream, as the CUDA-aware
// build does with cudaStreamSynchronize
extern "C" __attribute__((used, visibility("default"))) inline void emu_stream_synchronize(void* st) { cudaStreamSynchronize((cudaStream_t)st); }

static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
#ifdef EMU_DEFERRED
    if (st && n) {
        const bool src_host = kind == cudaMemcpyHostToDevice || kind == cudaMemcpyHostToHost;
        const bool dst_host = kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyHostToHost;
        if (dst_host && !emu::is_pinned(d)) {                 // pageable destination: done when the call returns
            cudaStreamSynchronize(st);
            memmove(d, s, n);
            return cudaSuccess;
        }
        if (src_host && !emu::is_pinned(s)) {                 // pageable source: staged now, copied later
            std::shared_ptr<std::vector<char>> stage = std::make_shared<std::vector<char>>((const char*)s, (const char*)s + n);
            emu::enqueue(st, [d, stage, n]() { memcpy(d, stage->data(), n); });
            return cudaSuccess;
        }
        emu::enqueue(st, [d, s, n]() { memmove(d, s, n); });
        return cudaSuccess;
    }
#endif
    (void)kind; (void)st;
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) {
    emu::Stream* st = new emu::Stream();
    emu::my_streams().push_back(st);
    *s = (cudaStream_t)st;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t st) {
    emu::Stream* s = (emu::Stream*)st;
    emu::drain(s, s->submitted);
    auto& v = emu::my_streams();
    v.erase(std::remove(v.begin(), v.end(), s), v.end());
    delete s;
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st) {
    e->t = std::chrono::steady_clock::now();
    e->s = (emu::Stream*)st;
    e->seq = e->s ? e->s->submitted : 0;                      // everything enqueued on the stream so far
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { if (e->s) emu::drain(e->s, e->seq); return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t e) { return (!e->s || e->s->executed >= e->seq) ? cudaSuccess : cudaErrorNotReady; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned) {
    emu::Stream* es = e->s;
    const uint64_t seq = e->seq;                              // the event's state at the time of this call
    if (es && es != (emu::Stream*)st) emu::enqueue(st, [es, seq]() { emu::drain(es, seq); });
    return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
    memset(a, 0, sizeof *a); a->type = cudaMemoryTypeUnregistered; return cudaSuccess;
}
