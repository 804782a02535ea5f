// warp_emu.hpp -- a tiny SIMT emulator for CPU tests of warp-synchronous CUDA kernels (TEST INFRASTRUCTURE).
//
// One warp = 32 fibers (ucontext) inside one OS thread.  A lane runs until it reaches a warp collective
// (__shfl*_sync, __ballot_sync, __any/__all_sync, __syncwarp) and yields; when every live lane has arrived at the
// SAME collective the values are exchanged and all lanes continue.  A lane that finishes while others wait in a
// full-mask collective, or lanes arriving at different collectives, abort the test: on hardware that would be a
// hang or undefined behaviour.  Only full masks are supported (all this repository's kernels use 0xffffffff).
//
// The kernel source is compiled unchanged by g++: this header supplies the CUDA spellings it uses (qualifiers,
// threadIdx/blockIdx, the integer intrinsics, atomics on plain memory -- there is only one OS thread).
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// the vector types the kernels use, with CUDA's alignment (no CUDA header is involved in an emulated build)
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

namespace emu {

constexpr int W = 32;
enum Op : int { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_BALLOT, OP_ANY, OP_ALL, OP_SYNCWARP };

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[W];
    std::vector<char> stack[W];
    bool done[W], waiting[W];
    int cur = 0;
    Op pending_op[W];
    uint64_t pending[W], result[W];
    std::function<void()> body;
    uint64_t n_collectives = 0;
    Dim3 block_idx, block_dim, grid_dim;
};

inline Warp*& current() { static Warp* w = nullptr; return w; }

[[noreturn]] inline void die(const char* msg) {
    fprintf(stderr, "warp_emu: %s\n", msg);
    abort();
}

inline void lane_entry() {
    Warp* w = current();
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
    die("finished lane resumed");
}

// Runs `body` once per lane of one warp (threadIdx.x = 0..31) to completion.
inline void run_warp(const std::function<void()>& body, Dim3 block_idx = Dim3(), Dim3 grid_dim = Dim3()) {
    Warp w;
    w.body = body;
    w.block_idx = block_idx;
    w.grid_dim = grid_dim;
    w.block_dim.x = W;
    Warp* outer = current();
    current() = &w;
    for (int i = 0; i < W; i++) {
        w.done[i] = w.waiting[i] = false;
        w.pending_op[i] = OP_NONE;
        w.stack[i].resize(512 * 1024);
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stack[i].data();
        w.ctx[i].uc_stack.ss_size = w.stack[i].size();
        w.ctx[i].uc_link = nullptr;
        makecontext(&w.ctx[i], (void (*)())lane_entry, 0);
    }
    for (;;) {
        bool progressed = false;
        for (int i = 0; i < W; i++) {
            if (w.done[i] || w.waiting[i]) continue;
            w.cur = i;
            swapcontext(&w.sched, &w.ctx[i]);              // runs lane i up to its next collective (or its end)
            progressed = true;
        }
        int n_done = 0, n_wait = 0;
        for (int i = 0; i < W; i++) { n_done += w.done[i]; n_wait += w.waiting[i]; }
        if (n_done == W) break;
        if (n_wait && n_done) die("some lanes exited while others wait in a full-mask collective");
        if (n_wait == W) {                                  // everyone arrived: same collective?
            for (int i = 1; i < W; i++)
                if (w.pending_op[i] != w.pending_op[0]) die("lanes arrived at different collectives (divergent call)");
            memcpy(w.result, w.pending, sizeof w.result);
            for (int i = 0; i < W; i++) w.waiting[i] = false;
            w.n_collectives++;
            continue;
        }
        if (!progressed) die("scheduler stuck");
    }
    current() = outer;
}

inline int lane() { return current()->cur; }

// arrive at a collective with my value; returns after all lanes arrived (results in current()->result[])
inline void collective(Op op, uint64_t v, unsigned mask) {
    if (mask != 0xffffffffu) die("only full-mask collectives are emulated");
    Warp* w = current();
    const int me = w->cur;
    w->pending[me] = v;
    w->pending_op[me] = op;
    w->waiting[me] = true;
    swapcontext(&w->ctx[me], &w->sched);
    w->cur = me;                                            // (the scheduler set it already; kept for clarity)
}

struct ThreadIdx { struct X { operator unsigned() const { return (unsigned)lane(); } } x; };
struct BlockIdxT { struct X { operator unsigned() const { return current()->block_idx.x; } } x; };
struct BlockDimT { struct X { operator unsigned() const { return current()->block_dim.x; } } x; };
struct GridDimT { struct X { operator unsigned() const { return current()->grid_dim.x; } } x; };

}  // namespace emu

// ---- CUDA spellings ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __restrict__

static emu::ThreadIdx threadIdx;
static emu::BlockIdxT blockIdx;
static emu::BlockDimT blockDim;
static emu::GridDimT gridDim;
typedef void* cudaStream_t;

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    emu::collective(emu::OP_SHFL, raw, mask);
    T out; memcpy(&out, &emu::current()->result[src & 31], sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    emu::collective(emu::OP_SHFL_UP, raw, mask);
    const int me = emu::lane();
    const int src = me - (int)delta;
    T out; memcpy(&out, &emu::current()->result[src >= 0 ? src : me], sizeof(T));   // out of range: own value
    return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::collective(emu::OP_BALLOT, pred ? 1 : 0, mask);
    unsigned b = 0;
    for (int i = 0; i < emu::W; i++) b |= (unsigned)(emu::current()->result[i] & 1) << i;
    return b;
}
static inline int __any_sync(unsigned mask, int pred) {
    emu::collective(emu::OP_ANY, pred ? 1 : 0, mask);
    for (int i = 0; i < emu::W; i++) if (emu::current()->result[i]) return 1;
    return 0;
}
static inline int __all_sync(unsigned mask, int pred) {
    emu::collective(emu::OP_ALL, pred ? 1 : 0, mask);
    for (int i = 0; i < emu::W; i++) if (!emu::current()->result[i]) return 0;
    return 1;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::collective(emu::OP_SYNCWARP, 0, mask); }
static inline void __syncthreads() { emu::die("__syncthreads is not emulated (one warp only)"); }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }

static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
