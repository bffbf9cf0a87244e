// Shared declarations of the qtt_b200 library: handle, workspace cache, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <nvtx3/nvToolsExt.h>
#include "../../include/qtt_b200.h"

// NVTX ranges around the sweeps / solver phases (header-only NVTX3: a no-op unless a tool such as nsys / ncu --nvtx is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

typedef long long ll;

struct WsBlock { char* ptr; size_t bytes; bool busy; };

struct qtt_handle_s {
    int device = 0;
    std::string err;
    std::vector<WsBlock> blocks;      // cached device allocations (never shrinks below the limit)
    size_t ws_limit = (size_t)64 << 30;
    size_t ws_total = 0;
    int* h_ranks = nullptr;           // pinned host scratch for rank tables
    size_t h_ranks_cap = 0;
    double gemm_flops = 0.0;
    double leaf_flops = 0.0;          // executed flops of the panel QR leaves (left-looking sub-panels)
    ll launches = 0;
    int sm_count = 148;
    int max_cluster = 8;              // largest thread-block cluster the tall-panel leaf may use (16 when the GPU can schedule it)
    // secondary stream: the exact head compression of operator terms (small, latency-bound kernels on the first cores) runs beside the
    // tail of the right-to-left sweep, which does not need its result before core kc (QTT_HEAD_OVERLAP=0 turns the overlap off)
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool head_overlap = true;
    double* coop_scratch[2] = {nullptr, nullptr};   // reduction scratch of the cooperative QR leaf: [0] main stream, [1] secondary stream
    size_t coop_cap[2] = {0, 0};
    // optional per-launch timing of the GEMM kernel (CUDA events on the launching stream)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;      // pairs: start, stop
    std::vector<double> prof_fl;           // flops of the launch timed by pair i
    size_t prof_used = 0;                  // pairs in flight
    double prof_ms = 0.0, prof_flops = 0.0;
    ll prof_count = 0;
    // optional timeline mode (qtt_set_profiling(h, 2)): one event after every launch, interval attributed to the
    // launch that just finished; classes: 0 gemm, 1 larfb, 2 qr leaf, 3 gram/larft, 4 jacobi, 5 other
    bool tl_on = false;
    std::vector<cudaEvent_t> tl_ev; std::vector<int> tl_cls; size_t tl_used = 0;
    double tl_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // CUDA-graph cache of whole sweeps (qtt_b200.cu: sweep_group).  A sweep of small cores is a chain of ~100-250 kernels of a few
    // microseconds each, bound by the host's launch rate; inside a solver the same sweep (same buffers, same ranks) recurs every
    // iteration, so its launch sequence is captured once and replayed with one cudaGraphLaunch.
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        std::vector<int> ws_idx; std::vector<char*> ws_ptr;   // workspace blocks the captured kernels use (must be free and unmoved at replay)
        double gemm_flops = 0.0, leaf_flops = 0.0; ll launches = 0;
        unsigned long long ws_gen = 0, stamp = 0;
    };
    std::map<std::vector<ll>, GraphEntry> graphs;
    cudaStream_t st_cap = nullptr;    // capture stream (the caller's stream may be the legacy default stream, which cannot capture)
    bool graphs_on = false;           // QTT_GRAPHS=1 turns the cache on (off by default: measured, no gain - see sweep_group)
    bool ws_tracing = false;          // WsScope records the blocks it hands out (during a capture)
    std::vector<int> ws_trace;
    unsigned long long ws_gen = 0;    // bumped whenever a cached workspace block is freed: invalidates every captured graph
    unsigned long long graph_stamp = 0;
    ll graph_replays = 0, graph_captures = 0;
};

static inline void tl_flush(qtt_handle_s* h) {
    if (h->tl_used < 2) { h->tl_used = 0; return; }
    cudaEventSynchronize(h->tl_ev[h->tl_used - 1]);
    for (size_t i = 1; i < h->tl_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->tl_ev[i - 1], h->tl_ev[i]) == cudaSuccess) h->tl_ms[h->tl_cls[i] & 7] += ms;
    }
    h->tl_used = 0;
}
static inline void tl_mark(qtt_handle_s* h, cudaStream_t st, int cls) {
    if (!h->tl_on) return;
    if (h->tl_used >= h->tl_ev.size()) {
        // keep continuity: flush, then re-record a starting point
        tl_flush(h);
        cudaEventRecord(h->tl_ev[0], st); h->tl_cls[0] = 5; h->tl_used = 1;
    }
    cudaEventRecord(h->tl_ev[h->tl_used], st); h->tl_cls[h->tl_used] = cls; h->tl_used++;
}

static inline void prof_flush(qtt_handle_s* h) {
    if (h->prof_used == 0) return;
    for (size_t i = 0; i < h->prof_used; ++i) {
        float ms = 0.f;
        cudaEventSynchronize(h->prof_ev[2 * i + 1]);      // pairs may sit on two streams (head compression overlaps the sweep)
        if (cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]) == cudaSuccess) {
            h->prof_ms += ms; h->prof_flops += h->prof_fl[i]; h->prof_count += 1;
        }
    }
    h->prof_used = 0;
}

#define QTT_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                     cudaGetErrorString(e__));                                                  \
            h->err = buf__;                                                                     \
            return QTT_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define QTT_CHECK(expr)                                                                         \
    do {                                                                                        \
        int rc__ = (expr);                                                                      \
        if (rc__ != QTT_OK) return rc__;                                                        \
    } while (0)

#define QTT_REQUIRE(cond, code, msg)                                                            \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            char buf__[512];                                                                    \
            snprintf(buf__, sizeof buf__, "%s:%d %s", __FILE__, __LINE__, msg);                 \
            h->err = buf__;                                                                     \
            return code;                                                                        \
        }                                                                                       \
    } while (0)

// Workspace: a per-call scope takes cached blocks (best fit) or allocates new ones; everything is
// released (marked free, not cudaFree'd) when the scope ends.  All work of a handle runs on one
// stream, so reuse by the next call is stream-ordered.
struct WsScope {
    qtt_handle_s* h;
    std::vector<int> taken;
    explicit WsScope(qtt_handle_s* hh) : h(hh) {}
    ~WsScope() { for (int i : taken) h->blocks[i].busy = false; }
    int alloc(void** out, size_t bytes) {
        if (bytes == 0) bytes = 256;
        bytes = (bytes + 255) & ~(size_t)255;
        int best = -1;
        for (int i = 0; i < (int)h->blocks.size(); ++i) {
            WsBlock& b = h->blocks[i];
            if (!b.busy && b.bytes >= bytes && (best < 0 || b.bytes < h->blocks[best].bytes)) best = i;
        }
        if (best >= 0 && h->blocks[best].bytes <= 2 * bytes + (1 << 20)) {
            h->blocks[best].busy = true; taken.push_back(best); *out = h->blocks[best].ptr;
            if (h->ws_tracing) h->ws_trace.push_back(best);
            return QTT_OK;
        }
        // allocate; if over the limit, drop idle blocks first
        if (h->ws_total + bytes > h->ws_limit) {
            cudaDeviceSynchronize();
            for (auto& b : h->blocks) if (!b.busy && b.ptr) { cudaFree(b.ptr); h->ws_total -= b.bytes; b.ptr = nullptr; b.bytes = 0; h->ws_gen++; }
        }
        char* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceSynchronize();
            for (auto& b : h->blocks) if (!b.busy && b.ptr) { cudaFree(b.ptr); h->ws_total -= b.bytes; b.ptr = nullptr; b.bytes = 0; h->ws_gen++; }
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess) {
            char buf[256]; snprintf(buf, sizeof buf, "workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            h->err = buf; cudaGetLastError(); return QTT_ERR_CUDA;
        }
        h->ws_total += bytes;
        int slot = -1;
        for (int i = 0; i < (int)h->blocks.size(); ++i) if (!h->blocks[i].ptr) { slot = i; break; }
        WsBlock nb{p, bytes, true};
        if (slot < 0) { h->blocks.push_back(nb); slot = (int)h->blocks.size() - 1; } else h->blocks[slot] = nb;
        taken.push_back(slot); *out = p;
        if (h->ws_tracing) h->ws_trace.push_back(slot);
        return QTT_OK;
    }
    template <typename T> int alloc_t(T** out, size_t count) { return alloc((void**)out, count * sizeof(T)); }
    // take exactly these blocks again (replay of a captured graph): all must still exist at the same address and be free
    bool retake(const std::vector<int>& idx, const std::vector<char*>& ptr) {
        for (size_t i = 0; i < idx.size(); ++i) {
            if (idx[i] < 0 || idx[i] >= (int)h->blocks.size()) return false;
            const WsBlock& b = h->blocks[idx[i]];
            if (b.busy || b.ptr != ptr[i]) return false;
        }
        for (int i : idx) { h->blocks[i].busy = true; taken.push_back(i); }
        return true;
    }
};

static inline ll ceil_div(ll a, ll b) { return (a + b - 1) / b; }

// development switches are read from the environment once per process, not per launch
#define QTT_ENV_FLAG(name) ([]() -> bool { static const bool v = (getenv(name) != nullptr); return v; }())
