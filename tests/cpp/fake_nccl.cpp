// fake_nccl.cpp -- an in-process stand-in for libnccl.so.2 (TESTS ONLY).  Ranks are OS threads of one process, each
// driving its own emulated context (tests/test_capi_emulated_cpu.py); "device" buffers are host memory, so the
// collectives are memcpy between threads behind barriers.  Only what libmdbg_b200 dlsym()s is provided:
// ncclGetUniqueId, ncclCommInitRank, ncclCommDestroy, ncclAllGather, ncclGroupStart/End + ncclSend/Recv,
// ncclGetErrorString.  Built as libnccl.so.2 into a temporary directory that is put first on LD_LIBRARY_PATH.
//
// With -DFAKE_NCCL_CUDA (scripts/gpu_selftest.sh) the buffers are real device memory: N ranks = N threads sharing
// ONE GPU (real NCCL refuses two ranks on one device), each collective first drains the rank's stream and then
// copies with cudaMemcpy.  That runs the product's multi-rank kernels and host sequencing on real hardware when
// only a single GPU is available; it is not a performance path.
#ifdef FAKE_NCCL_CUDA
#include <cuda_runtime.h>
#endif
#include <dlfcn.h>

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

#ifdef FAKE_NCCL_CUDA
inline void copy_bytes(void* d, const void* s, size_t n) { cudaMemcpy(d, s, n, cudaMemcpyDefault); }
inline void drain(void* stream) { cudaStreamSynchronize((cudaStream_t)stream); }
#else
inline void copy_bytes(void* d, const void* s, size_t n) { memcpy(d, s, n); }
// the emulated library built with -DEMU_DEFERRED queues stream work; it exports a hook to flush a stream
inline void drain(void* stream) {
    static void (*hook)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "emu_stream_synchronize");
    if (hook && stream) hook(stream);
}
#endif

struct Barrier {
    std::mutex m; std::condition_variable cv; int n = 0, waiting = 0; uint64_t gen = 0;
    void wait() {
        std::unique_lock<std::mutex> g(m);
        const uint64_t my = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(g, [&] { return gen != my; });
    }
};
struct Msg { const void* p; size_t bytes; };
struct Group {
    int n = 0, joined = 0;
    Barrier bar;
    std::vector<const void*> gather_src;
    std::vector<std::vector<std::vector<Msg>>> box;       // [src][dst] -> messages in posting order
};
struct Comm { Group* g; int rank; };
struct PendingOp { bool send; const void* sp; void* rp; size_t bytes; int peer; };

std::mutex g_mu;
std::map<std::string, Group*> g_groups;
uint64_t g_next_id = 1;
thread_local std::vector<PendingOp> t_ops;
thread_local int t_depth = 0;
thread_local Comm* t_comm = nullptr;
thread_local void* t_stream = nullptr;

size_t dtype_size(int t) { static const size_t s[] = {1, 1, 4, 4, 8, 8, 2, 4, 8, 2}; return (t >= 0 && t < 10) ? s[t] : 1; }

int flush_group_ops() {
    if (!t_comm) { t_ops.clear(); return 0; }                // no collective seen on this thread yet
    Group* g = t_comm->g;                                    // (a rank with nothing to send still takes the barriers)
    const int me = t_comm->rank;
    for (const PendingOp& op : t_ops)
        if (op.send) g->box[me][op.peer].push_back(Msg{op.sp, op.bytes});
    drain(t_stream);                                         // what this rank sends has been produced
    g->bar.wait();                                           // everybody has posted
    std::vector<size_t> next(g->n, 0);
    int rc = 0;
    for (const PendingOp& op : t_ops)
        if (!op.send) {
            auto& q = g->box[op.peer][me];
            if (next[op.peer] >= q.size() || q[next[op.peer]].bytes != op.bytes) { rc = 3; continue; }   // mismatched send/recv
            copy_bytes(op.rp, q[next[op.peer]].p, op.bytes);
            next[op.peer]++;
        }
    g->bar.wait();                                           // everybody has copied
    for (int d = 0; d < g->n; d++) g->box[me][d].clear();
    g->bar.wait();
    t_ops.clear();
    return rc;
}

}  // namespace

extern "C" {

struct ncclUniqueId { char internal[128]; };

int ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> l(g_mu);
    memset(id, 0, sizeof *id);
    const uint64_t v = g_next_id++;
    memcpy(id->internal, &v, sizeof v);
    memcpy(id->internal + 8, "fake-nccl", 9);
    return 0;
}

int ncclCommInitRank(void** comm, int nranks, ncclUniqueId id, int rank) {
    Group* g;
    {
        std::lock_guard<std::mutex> l(g_mu);
        const std::string key(id.internal, sizeof id.internal);
        auto it = g_groups.find(key);
        if (it == g_groups.end()) {
            g = new Group();
            g->n = nranks; g->bar.n = nranks;
            g->gather_src.assign(nranks, nullptr);
            g->box.assign(nranks, std::vector<std::vector<Msg>>(nranks));
            g_groups[key] = g;
        } else g = it->second;
        if (g->n != nranks || rank < 0 || rank >= nranks) return 4;
        g->joined++;
    }
    *comm = new Comm{g, rank};
    g->bar.wait();                                           // like NCCL: returns when every rank has joined
    return 0;
}

int ncclCommDestroy(void* comm) { delete (Comm*)comm; return 0; }

int ncclAllGather(const void* send, void* recv, size_t count, int dtype, void* comm, void* stream) {
    Comm* c = (Comm*)comm;
    t_comm = c;
    drain(stream);
    const size_t bytes = count * dtype_size(dtype);
    c->g->gather_src[c->rank] = send;
    c->g->bar.wait();
    for (int r = 0; r < c->g->n; r++) copy_bytes((char*)recv + (size_t)r * bytes, c->g->gather_src[r], bytes);
    c->g->bar.wait();
    return 0;
}

int ncclGroupStart() { t_depth++; return 0; }
int ncclGroupEnd() { if (--t_depth == 0) return flush_group_ops(); return 0; }

int ncclSend(const void* p, size_t count, int dtype, int peer, void* comm, void* stream) {
    t_comm = (Comm*)comm;
    t_stream = stream;
    t_ops.push_back(PendingOp{true, p, nullptr, count * dtype_size(dtype), peer});
    return t_depth ? 0 : flush_group_ops();
}
int ncclRecv(void* p, size_t count, int dtype, int peer, void* comm, void* stream) {
    t_comm = (Comm*)comm;
    t_stream = stream;
    t_ops.push_back(PendingOp{false, nullptr, p, count * dtype_size(dtype), peer});
    return t_depth ? 0 : flush_group_ops();
}

const char* ncclGetErrorString(int e) { return e == 0 ? "no error" : e == 3 ? "fake nccl: send/recv mismatch" : "fake nccl: invalid argument"; }

}  // extern "C"
