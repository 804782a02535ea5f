// warp_emu.hpp -- a tiny SIMT emulator for CPU tests of warp-synchronous CUDA kernels (TEST INFRASTRUCTURE).
//
// One thread block = n fibers (ucontext) inside one OS thread, grouped in warps of 32.  A thread runs until it
// reaches a collective and yields: warp collectives (__shfl*_sync, __ballot_sync, __any/__all_sync, __syncwarp)
// complete when every lane of that warp has arrived at the SAME collective, __syncthreads when every thread of the
// block has.  A lane that finishes while others wait in a full-mask collective, lanes arriving at different
// collectives, or a barrier some threads never reach abort the test: on hardware that would be a hang or undefined
// behaviour.  Only full masks are supported (all this repository's kernels use 0xffffffff).  Blocks of a grid run
// one after the other.
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
#include <deque>
#include <functional>
#include <vector>

// the vector types the kernels use, with CUDA's alignment (no CUDA header is involved in an emulated build)
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

// ---- fiber switch ------------------------------------------------------------------------------------------------
// glibc's swapcontext makes a sigprocmask system call per switch; a block of 256 threads switches millions of
// times, so x86-64 gets a 12-instruction switch of its own (callee-saved registers + stack pointer).
#if defined(__x86_64__) && !defined(EMU_USE_UCONTEXT)
extern "C" void emu_switch_x86(void** from_sp, void** to_sp);
asm(".text\n"
    ".weak emu_switch_x86\n"
    ".type emu_switch_x86,@function\n"
    "emu_switch_x86:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq (%rsi), %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n"
    "    ret\n"
    ".size emu_switch_x86, .-emu_switch_x86\n");
namespace emu {
struct Fiber { void* sp = nullptr; };
inline void fiber_init(Fiber& f, char* stack, size_t bytes, void (*entry)()) {
    uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
    void** p = (void**)top;
    *--p = nullptr;                 // return address of `entry` (it never returns)
    *--p = (void*)entry;            // popped by the switch's `ret`
    for (int i = 0; i < 6; i++) *--p = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = p;
}
inline void fiber_switch(Fiber& from, Fiber& to) { emu_switch_x86(&from.sp, &to.sp); }
}  // namespace emu
#else
namespace emu {
struct Fiber { ucontext_t uc; };
inline void fiber_init(Fiber& f, char* stack, size_t bytes, void (*entry)()) {
    getcontext(&f.uc);
    f.uc.uc_stack.ss_sp = stack;
    f.uc.uc_stack.ss_size = bytes;
    f.uc.uc_link = nullptr;
    makecontext(&f.uc, entry, 0);
}
inline void fiber_switch(Fiber& from, Fiber& to) { swapcontext(&from.uc, &to.uc); }
}  // namespace emu
#endif

namespace emu {

constexpr int W = 32;
constexpr int MAX_THREADS = 1024;
enum Op : int { OP_NONE = 0, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_ANY, OP_ALL, OP_SYNCWARP,
                OP_SYNCTHREADS };

struct Dim3 { unsigned x = 1, y = 1, z = 1; };

struct Block {
    Fiber sched;
    std::vector<Fiber> ctx;
    std::vector<char> done, waiting;
    std::vector<Op> pending_op;
    std::vector<uint64_t> pending, result;
    int n_threads = 0, cur = 0;
    std::function<void()> body;
    uint64_t n_collectives = 0;
    Dim3 block_idx, block_dim, grid_dim;
};

inline Block*& current() { static thread_local Block* b = nullptr; return b; }   // one emulated GPU per OS thread
inline bool outer_running() { return current() != nullptr; }
constexpr size_t STACK_BYTES = 256 * 1024;
inline char* fiber_stack(int i) {                           // stacks are reused by every block of every launch
    static thread_local std::vector<std::vector<char>> pool(MAX_THREADS);
    if (pool[i].empty()) pool[i].resize(STACK_BYTES);
    return pool[i].data();
}

[[noreturn]] inline void die(const char* msg) {
    fprintf(stderr, "warp_emu: %s\n", msg);
    abort();
}

inline void thread_entry() {
    Block* b = current();
    b->body();
    b->done[b->cur] = 1;
    fiber_switch(b->ctx[b->cur], b->sched);
    die("finished thread resumed");
}

// Runs `body` once per thread of one block (threadIdx.x = 0 .. n_threads-1, warps of 32) to completion.
inline void run_block(int n_threads, const std::function<void()>& body, Dim3 block_idx = Dim3(), Dim3 grid_dim = Dim3()) {
    if (n_threads <= 0 || n_threads > MAX_THREADS || (n_threads % W) != 0) die("block size must be a multiple of 32");
    Block b;
    b.body = body;
    b.n_threads = n_threads;
    b.block_idx = block_idx;
    b.grid_dim = grid_dim;
    b.block_dim.x = (unsigned)n_threads;
    if (outer_running()) die("nested run_block");
    b.ctx.resize(n_threads);
    b.done.assign(n_threads, 0); b.waiting.assign(n_threads, 0);
    b.pending_op.assign(n_threads, OP_NONE);
    b.pending.assign(n_threads, 0); b.result.assign(n_threads, 0);
    Block* outer = current();
    current() = &b;
    for (int i = 0; i < n_threads; i++) {
        fiber_init(b.ctx[i], fiber_stack(i), STACK_BYTES, thread_entry);
    }
    const int n_warps = n_threads / W;
    for (;;) {
        bool progressed = false;
        for (int i = 0; i < n_threads; i++) {
            if (b.done[i] || b.waiting[i]) continue;
            b.cur = i;
            fiber_switch(b.sched, b.ctx[i]);                // runs thread i up to its next collective (or its end)
            progressed = true;
        }
        int n_done = 0;
        for (int i = 0; i < n_threads; i++) n_done += b.done[i];
        if (n_done == n_threads) break;
        bool released = false;
        // warp collectives: a warp continues when all of its lanes wait at the same warp-level collective
        for (int w = 0; w < n_warps; w++) {
            int wd = 0, ww = 0;
            for (int l = 0; l < W; l++) {
                const int t = w * W + l;
                wd += b.done[t];
                ww += b.waiting[t] && b.pending_op[t] != OP_SYNCTHREADS;
            }
            if (ww == 0) continue;
            if (wd) die("some lanes exited while others wait in a full-mask warp collective");
            if (ww != W) {
                bool any_block = false;
                for (int l = 0; l < W; l++) any_block |= b.waiting[w * W + l] && b.pending_op[w * W + l] == OP_SYNCTHREADS;
                if (any_block) die("lanes of one warp wait at a warp collective and at __syncthreads");
                continue;                                    // (cannot happen after a full sweep, kept for safety)
            }
            for (int l = 1; l < W; l++)
                if (b.pending_op[w * W + l] != b.pending_op[w * W]) die("lanes arrived at different collectives (divergent call)");
            for (int l = 0; l < W; l++) { b.result[w * W + l] = b.pending[w * W + l]; b.waiting[w * W + l] = 0; }
            b.n_collectives++;
            released = true;
        }
        // block barrier: continues when every live thread waits at __syncthreads
        if (!released) {
            int at_barrier = 0;
            for (int i = 0; i < n_threads; i++) at_barrier += b.waiting[i] && b.pending_op[i] == OP_SYNCTHREADS;
            if (at_barrier) {
                if (at_barrier + n_done != n_threads) die("__syncthreads not reached by every live thread");
                if (n_done) die("threads exited before a __syncthreads the others wait at");
                for (int i = 0; i < n_threads; i++) b.waiting[i] = 0;
                released = true;
            }
        }
        if (!progressed && !released) die("scheduler stuck");
    }
    current() = outer;
}

inline void run_warp(const std::function<void()>& body, Dim3 block_idx = Dim3(), Dim3 grid_dim = Dim3()) {
    run_block(W, body, block_idx, grid_dim);
}

// kernel<<<grid, block>>>: blocks run one after the other
inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    Dim3 g; g.x = grid;
    for (unsigned bx = 0; bx < grid; bx++) { Dim3 bi; bi.x = bx; run_block((int)block, body, bi, g); }
}

// ---- streams ------------------------------------------------------------------------------------------------------
// Default build: every enqueued operation runs at once (a valid serialisation of stream semantics).
// -DEMU_DEFERRED: the ADVERSARIAL serialisation -- work enqueued on a stream sits in that stream's queue until the
// host (or another stream, through an event) actually waits for it, so a missing synchronisation shows up as stale
// or poisoned data instead of being hidden by eager execution.  Streams belong to the thread that created them.
struct Stream {
    std::deque<std::function<void()>> q;
    uint64_t submitted = 0, executed = 0;
};
inline std::vector<Stream*>& my_streams() { static thread_local std::vector<Stream*> v; return v; }
inline void drain(Stream* s, uint64_t upto) {
    while (s && s->executed < upto && !s->q.empty()) {
        std::function<void()> f = std::move(s->q.front());
        s->q.pop_front();
        s->executed++;
        f();
    }
}
inline void drain_all() { for (Stream* s : my_streams()) drain(s, s->submitted); }
inline void enqueue(void* stream, std::function<void()> f) {
#ifdef EMU_DEFERRED
    if (stream) {
        Stream* s = (Stream*)stream;
        s->q.push_back(std::move(f));
        s->submitted++;
        return;
    }
#endif
    (void)stream;
    f();
}
// kernel<<<grid, block, 0, stream>>>(args): the body must hold its arguments by value (tests/_emu.py writes [=])
inline void launch(unsigned grid, unsigned block, void* stream, std::function<void()> body) {
    enqueue(stream, [grid, block, body]() { launch(grid, block, body); });
}

inline int tid() { return current()->cur; }
inline int lane() { return current()->cur & (W - 1); }
inline int warp_base() { return current()->cur & ~(W - 1); }

// arrive at a collective with my value; returns after all participants arrived (warp results in result[])
inline void collective(Op op, uint64_t v, unsigned mask) {
    if (mask != 0xffffffffu) die("only full-mask collectives are emulated");
    Block* b = current();
    const int me = b->cur;
    b->pending[me] = v;
    b->pending_op[me] = op;
    b->waiting[me] = 1;
    fiber_switch(b->ctx[me], b->sched);
    b->cur = me;
}
inline uint64_t peer(int lane_idx) { return current()->result[warp_base() + (lane_idx & 31)]; }

struct ThreadIdx { struct X { operator unsigned() const { return (unsigned)tid(); } } x; };
struct BlockIdxT { struct X { operator unsigned() const { return current()->block_idx.x; } } x; };
struct BlockDimT { struct X { operator unsigned() const { return current()->block_dim.x; } } x; };
struct GridDimT { struct X { operator unsigned() const { return current()->grid_dim.x; } } x; };

}  // namespace emu

// ---- CUDA spellings ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static thread_local      /* per OS thread: several emulated ranks may run the same kernel */
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __restrict__

static emu::ThreadIdx threadIdx;
static emu::BlockIdxT blockIdx;
static emu::BlockDimT blockDim;
static emu::GridDimT gridDim;
typedef void* cudaStream_t;
// the two runtime calls the launch_* functions make
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kernel, threads, smem) (*(out) = 2)
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t s) { emu::enqueue(s, [=]() { memset(p, v, n); }); return 0; }

template <typename T> static inline uint64_t emu_raw(T v) { static_assert(sizeof(T) <= 8, "payload"); uint64_t r = 0; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline T emu_val(uint64_t r) { T v; memcpy(&v, &r, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    emu::collective(emu::OP_SHFL, emu_raw(v), mask);
    return emu_val<T>(emu::peer(src));
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    emu::collective(emu::OP_SHFL_UP, emu_raw(v), mask);
    const int me = emu::lane(), src = me - (int)delta;
    return emu_val<T>(emu::peer(src >= 0 ? src : me));        // out of range: own value
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
    emu::collective(emu::OP_SHFL_DOWN, emu_raw(v), mask);
    const int me = emu::lane(), src = me + (int)delta;
    return emu_val<T>(emu::peer(src < emu::W ? src : me));
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    emu::collective(emu::OP_SHFL_XOR, emu_raw(v), mask);
    return emu_val<T>(emu::peer(emu::lane() ^ lane_mask));
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {    // lanes holding the same value as mine
    emu::collective(emu::OP_SHFL, emu_raw(v), mask);
    unsigned m = 0;
    for (int i = 0; i < emu::W; i++) m |= (unsigned)(emu::peer(i) == emu_raw(v)) << i;
    return m;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::collective(emu::OP_BALLOT, pred ? 1 : 0, mask);
    unsigned b = 0;
    for (int i = 0; i < emu::W; i++) b |= (unsigned)(emu::peer(i) & 1) << i;
    return b;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::collective(emu::OP_SYNCWARP, 0, mask); }
static inline void __syncthreads() { emu::collective(emu::OP_SYNCTHREADS, 0, 0xffffffffu); }

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
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {     // selector nibbles 0..7 only (no sign replication)
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {     // n-th set bit at/after base (offset > 0 only)
    if (offset <= 0) emu::die("__fns: only positive offsets are emulated");
    for (unsigned b = base; b < 32; b++)
        if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xFFFFFFFFu;
}
static inline double __ddiv_rn(double a, double b) { return a / b; }          // build with -ffp-contract=off
static inline double __dadd_rn(double a, double b) { return a + b; }
template <typename T> static inline T atomicExch(T* p, T v) { T old = *p; *p = v; return old; }
template <typename T> static inline T atomicCAS(T* p, T cmp, T v) { T old = *p; if (old == cmp) *p = v; return old; }
template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> static inline T atomicOr(T* p, U v) { T old = *p; *p = (T)(old | (T)v); return old; }
template <typename T, typename U> static inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U> static inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }

static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
