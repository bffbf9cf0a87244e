// qtt_b200: host orchestration and C ABI (include/qtt_b200.h) of the batched FP64 tensor-train kernels.
//
// The rounding sweep (qtt_round_sum_batched) works on a SUM OF TERMS, each either a plain TT or an
// operator applied to a TT, and never materialises the summed / applied tensor:
//   right -> left   for every core k: the column block of every term of the working matrix
//                   M_k (n_k q_{k+1} x p_k, column-major) is produced by pushing the triangular factor of
//                   core k+1 through the term's core (one GEMM for a plain term, two for MPO x MPS),
//                   then M_k is factorised by blocked Householder QR; reflectors (V, T, tau) are kept.
//   left -> right   one-sided Jacobi SVD of the carry core with the reference's rank rule, then the
//                   carry S V^T is pushed through the stored reflectors of the next core.
// Reference behaviour restated: tensormethods.py:90-133 (sweeps), :176-194 (core product), :135-149 (sum).
#include <cstdlib>
#include <memory>
#include "common.cuh"
#include "gemm.cuh"
#include "gemm_push.cuh"
#include "qr.cuh"
#include "svd.cuh"
#include "larfb.cuh"
#include "qr2.cuh"
#include "qr3.cuh"
#include "tsqr.cuh"
#include "tsqr2.cuh"
#include "tt_ops.cuh"

#define QR_NB 64          // outer panel width of the blocked QR (= LARFB_NB)
// push_t_kernel / head_z_generic_kernel stage one vector core (rl x m x rr doubles) in shared memory: vector ranks up to
// 79 at mode size 4 (2D), 111 at mode size 2 (1D).  Larger cores are rejected with QTT_ERR_UNSUPPORTED (documented envelope).
#define XCORE_SMEM_MAX ((size_t)200 * 1024)
#define GRAM_SPLITS 8      // row splits of the panel Gram matrix

// Tall-skinny panels (tsqr.cuh): a full-width outer panel with at least g_ts_min_rows rows is factorised in row blocks
// and carries one T per row block.  QTT_TS_MIN_ROWS=0 disables the path; tests lower the threshold to reach it at small sizes.
static int g_leaf_small_rows = 768;  // panels with at most this many rows use the 256-thread leaf (QTT_LEAF_SMALL_ROWS)
static int g_leaf_loop = 1;
static int g_leaf_vres = 1;       // short panels keep their reflectors in shared memory inside the looped leaf (QTT_LEAF_VRES=0: off)
static int g_leaf_nt = 1024;      // threads of the left-looking leaf on tall panels (QTT_LEAF_NT=512: development switch)
static int g_force_cluster = 0;   // QTT_FORCE_CLUSTER_LEAF=1 (read at qtt_create): panels of >= 64 rows take the cluster leaf (tests reach it at small sizes)
#define TS_MIN_ROWS_DEFAULT 0    // full-width outer panels with at least this many rows take the tall-skinny (flat-tree) path
static int g_ts_min_rows = TS_MIN_ROWS_DEFAULT;
static ll g_graph_max_bytes = (ll)64 << 20;   // QTT_GRAPH_MAX_MB (read at qtt_create): sweeps with more workspace than this are not captured
static int g_svd_smem = 1;        // QTT_SVD_SMEM=0 (read at qtt_create): Jacobi vectors stay in global memory (A/B)
static int g_r_in_place = 1;      // QTT_R_IN_PLACE=0 (read at qtt_create): always extract R into its own buffer (A/B)
static int g_ts_leaf_gen = 2;     // 2: shared-memory row-block leaf (tsqr2.cuh); 1: register-resident first generation (QTT_TS_LEAF=1, A/B)
static inline int ts_nblk(bool ts, int rows, int jb) {
    if (!ts || g_ts_min_rows <= 0 || jb != QR_NB || rows < g_ts_min_rows || rows <= TS_TOP + TS_BS) return 0;
    return (int)((rows - TS_TOP + TS_BS - 1) / TS_BS);
}
// Outer panels of a factorisation with kmax reflectors: the NARROW panel (kmax mod 64 columns) comes FIRST, every later panel
// is 64 wide.  With the remainder at the end (644 = 10 x 64 + 4) every trailing update carried a 4-column block per instance,
// i.e. one CTA that streams the whole 64-column reflector panel for 4 columns of work (~10 % of the update's CTAs at the C5
// shapes); with the remainder in front the trailing widths are multiples of 64 and the 4-column panel streams 4 reflectors.
static int g_narrow_first = 0;    // QTT_NARROW_FIRST=1 (read at qtt_create): remainder panel first.  Measured: larfb 33.9 vs 31.7 ms per sweep - the 4-reflector
                                  // update of 640 columns still streams C at the cost of a full panel (it would need its own rank-4 kernel), so: off
static inline int qr_npan(int kmax) { return (int)((kmax + QR_NB - 1) / QR_NB); }
static inline int qr_first(int kmax) { const int r = kmax % QR_NB; return (g_narrow_first && r) ? r : std::min(QR_NB, kmax); }
static inline int qr_pstart(int kmax, int pi) { return pi == 0 ? 0 : qr_first(kmax) + (pi - 1) * QR_NB; }
static inline int qr_pwidth(int kmax, int pi) { return pi == 0 ? qr_first(kmax) : std::min(QR_NB, kmax - qr_pstart(kmax, pi)); }
// T slots (64 x 64 each) used by the outer panels before panel pi of an a x b factorisation
static inline ll tb_units_before(bool ts, int a, int kmax, int pi) {
    ll u = 0;
    for (int p = 0; p < pi && p < qr_npan(kmax); ++p) u += std::max(1, ts_nblk(ts, a - qr_pstart(kmax, p), qr_pwidth(kmax, p)));
    return u;
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
extern "C" int qtt_version(void) { return 100; }
static void graph_cache_clear(qtt_handle_s* h);

extern "C" int qtt_create(qtt_handle_t* handle, int device) {
    if (!handle) return QTT_ERR_INVALID;
    qtt_handle_s* h = new qtt_handle_s();
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return QTT_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    cudaFuncSetAttribute(qr_leaf_kernel<QR_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf_kernel<QR_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf2_kernel<QR_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf2_kernel<QR_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf3_kernel<QR_THREADS_BIG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf3_kernel<QR_THREADS_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    cudaFuncSetAttribute(qr_leaf3_kernel<QR_THREADS_BIG, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);   // clusters of 16 CTAs (panels > 20k rows)
    {   // can this GPU schedule a 16-CTA cluster of the leaf at its largest footprint?
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16, 1); cfg.blockDim = dim3(QR_THREADS_BIG); cfg.dynamicSmemBytes = qr_leaf3_smem(qr_leaf3_max_slice(QR_THREADS_BIG), QR_THREADS_BIG);
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, qr_leaf3_kernel<QR_THREADS_BIG, true>, &cfg) == cudaSuccess && ncl > 0) h->max_cluster = 16;
        cudaGetLastError();
        if (getenv("QTT_MAX_CLUSTER")) h->max_cluster = atoi(getenv("QTT_MAX_CLUSTER"));
    }
    cudaFuncSetAttribute(larfb_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_smem<8, 2>());
    cudaFuncSetAttribute(larfb_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_smem<16, 2>());
    cudaFuncSetAttribute(larfb_kernel<64, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_smem<64, 1>());
    cudaFuncSetAttribute(larfb_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_smem<32, 1>());
    cudaFuncSetAttribute(larfb2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb2_smem());
    if (getenv("QTT_LARFB2")) g_larfb2 = atoi(getenv("QTT_LARFB2"));
    cudaFuncSetAttribute(inner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(head_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(push_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XCORE_SMEM_MAX);
    cudaFuncSetAttribute(head_z_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XCORE_SMEM_MAX);
    cudaFuncSetAttribute(larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaFuncSetAttribute(tsqr_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsqr_leaf_smem());
    cudaFuncSetAttribute(larfb_ts_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_ts_smem<8, 2>());
    cudaFuncSetAttribute(larfb_ts_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)larfb_ts_smem<32, 1>());
    cudaFuncSetAttribute(gemm_push_kernel<56, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_push_smem<56, false>());
    cudaFuncSetAttribute(gemm_push_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_push_smem<64, false>());
    cudaFuncSetAttribute(gemm_push_kernel<56, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_push_smem<56, true>());
    cudaFuncSetAttribute(gemm_push_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_push_smem<64, true>());
    g_gemm_push = getenv("QTT_GEMM_PUSH") ? atoi(getenv("QTT_GEMM_PUSH")) : 1;
    h->graphs_on = getenv("QTT_GRAPHS") ? atoi(getenv("QTT_GRAPHS")) != 0 : false;
    if (getenv("QTT_GRAPH_MAX_MB")) g_graph_max_bytes = (ll)atoll(getenv("QTT_GRAPH_MAX_MB")) << 20;
    if (cudaStreamCreateWithFlags(&h->st_cap, cudaStreamNonBlocking) != cudaSuccess) { h->st_cap = nullptr; cudaGetLastError(); }
    cudaFuncSetAttribute(jacobi_svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    g_svd_smem = getenv("QTT_SVD_SMEM") ? atoi(getenv("QTT_SVD_SMEM")) : 1;
    g_r_in_place = getenv("QTT_R_IN_PLACE") ? atoi(getenv("QTT_R_IN_PLACE")) : 1;
    g_narrow_first = getenv("QTT_NARROW_FIRST") ? atoi(getenv("QTT_NARROW_FIRST")) : 0;
    cudaFuncSetAttribute(tsqr2_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsqr2_leaf_smem());
    g_ts_min_rows = getenv("QTT_TS_MIN_ROWS") ? atoi(getenv("QTT_TS_MIN_ROWS")) : TS_MIN_ROWS_DEFAULT;
    g_ts_leaf_gen = getenv("QTT_TS_LEAF") ? atoi(getenv("QTT_TS_LEAF")) : 2;
    if (getenv("QTT_HEAD_OVERLAP")) h->head_overlap = atoi(getenv("QTT_HEAD_OVERLAP")) != 0;
    if (cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) { h->head_overlap = false; cudaGetLastError(); }
    g_force_cluster = getenv("QTT_FORCE_CLUSTER_LEAF") ? atoi(getenv("QTT_FORCE_CLUSTER_LEAF")) : 0;
    if (getenv("QTT_LEAF_NT")) g_leaf_nt = atoi(getenv("QTT_LEAF_NT"));
    if (getenv("QTT_LEAF_LOOP")) g_leaf_loop = atoi(getenv("QTT_LEAF_LOOP"));
    if (getenv("QTT_LEAF_VRES")) g_leaf_vres = atoi(getenv("QTT_LEAF_VRES"));
    if (getenv("QTT_LEAF_SMALL_ROWS")) g_leaf_small_rows = atoi(getenv("QTT_LEAF_SMALL_ROWS"));
    cudaFuncSetAttribute(qr_leaf2_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    {   // A/B switches of kernel variants (development; defaults are the measured-faster paths)
        int v = getenv("QTT_LARFB_CP16") ? atoi(getenv("QTT_LARFB_CP16")) : 1;
        cudaMemcpyToSymbol(g_larfb_cp16, &v, sizeof v);
    }
    cudaGetLastError();
    *handle = h;
    return QTT_OK;
}

extern "C" int qtt_destroy(qtt_handle_t h) {
    if (!h) return QTT_ERR_INVALID;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& b : h->blocks) if (b.ptr) cudaFree(b.ptr);
    for (auto& e : h->prof_ev) cudaEventDestroy(e);
    if (h->h_ranks) cudaFreeHost(h->h_ranks);
    for (int i = 0; i < 2; ++i) if (h->coop_scratch[i]) cudaFree(h->coop_scratch[i]);
    graph_cache_clear(h);
    if (h->st_cap) cudaStreamDestroy(h->st_cap);
    if (h->st2) cudaStreamDestroy(h->st2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return QTT_OK;
}

extern "C" const char* qtt_last_error(qtt_handle_t h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int qtt_set_workspace_limit(qtt_handle_t h, size_t bytes) { if (!h) return QTT_ERR_INVALID; h->ws_limit = bytes; return QTT_OK; }
extern "C" int qtt_get_counters(qtt_handle_t h, double* f, ll* l) { if (!h) return QTT_ERR_INVALID; if (f) *f = h->gemm_flops; if (l) *l = h->launches; return QTT_OK; }
extern "C" int qtt_reset_counters(qtt_handle_t h) { if (!h) return QTT_ERR_INVALID; h->gemm_flops = 0; h->leaf_flops = 0; h->launches = 0; return QTT_OK; }
extern "C" int qtt_get_flop_counters(qtt_handle_t h, double* dmma, double* leaf) {
    if (!h) return QTT_ERR_INVALID;
    if (dmma) *dmma = h->gemm_flops;
    if (leaf) *leaf = h->leaf_flops;
    return QTT_OK;
}

extern "C" int qtt_get_graph_counters(qtt_handle_t h, ll* captures, ll* replays) {
    if (!h) return QTT_ERR_INVALID;
    if (captures) *captures = h->graph_captures;
    if (replays) *replays = h->graph_replays;
    return QTT_OK;
}

extern "C" int qtt_set_profiling(qtt_handle_t h, int on) {
    if (!h) return QTT_ERR_INVALID;
    if (on && h->prof_ev.empty()) {
        h->prof_ev.resize(2 * 8192); h->prof_fl.resize(8192);
        for (auto& e : h->prof_ev) if (cudaEventCreate(&e) != cudaSuccess) { h->err = "cudaEventCreate failed"; return QTT_ERR_CUDA; }
    }
    if (!on) prof_flush(h);
    h->prof_on = (on == 1);
    if (on == 2 && h->tl_ev.empty()) {
        h->tl_ev.resize(16384); h->tl_cls.resize(16384);
        for (auto& e : h->tl_ev) if (cudaEventCreate(&e) != cudaSuccess) { h->err = "cudaEventCreate failed"; return QTT_ERR_CUDA; }
    }
    if (on != 2) tl_flush(h);
    h->tl_on = (on == 2);
    return QTT_OK;
}
extern "C" int qtt_get_class_profile(qtt_handle_t h, double* ms8, int reset) {
    if (!h || !ms8) return QTT_ERR_INVALID;
    tl_flush(h);
    for (int i = 0; i < 8; ++i) { ms8[i] = h->tl_ms[i]; if (reset) h->tl_ms[i] = 0; }
    return QTT_OK;
}
extern "C" int qtt_get_profile(qtt_handle_t h, double* gemm_ms, double* gemm_flops, long long* gemm_launches, int reset) {
    if (!h) return QTT_ERR_INVALID;
    prof_flush(h);
    if (gemm_ms) *gemm_ms = h->prof_ms;
    if (gemm_flops) *gemm_flops = h->prof_flops;
    if (gemm_launches) *gemm_launches = h->prof_count;
    if (reset) { h->prof_ms = 0; h->prof_flops = 0; h->prof_count = 0; }
    return QTT_OK;
}

static cudaStream_t g_tl_stream = nullptr;     // stream of the call in flight (timeline profiling only)
static int check_launch(qtt_handle_s* h, const char* what) {
    if (h->tl_on) {
        int cls = 5;
        if (strstr(what, "qr_leaf")) cls = 2; else if (strstr(what, "gram") || strstr(what, "larft")) cls = 3; else if (strstr(what, "jacobi")) cls = 4;
        tl_mark(h, g_tl_stream, cls);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string(what) + ": " + cudaGetErrorString(e); return QTT_ERR_CUDA; }
    h->launches += 1;
    return QTT_OK;
}

static inline unsigned grid1d(ll total, int threads, int cap = 4096) {
    ll b = ceil_div(std::max<ll>(total, 1), threads);
    return (unsigned)std::min<ll>(b, cap);
}

static int read_ranks_host(qtt_handle_s* h, cudaStream_t st, const qtt_batch* x, std::vector<int>& out) {
    const size_t cnt = (size_t)x->N * (x->L + 1);
    if (h->h_ranks_cap < cnt) {
        if (h->h_ranks) cudaFreeHost(h->h_ranks);
        QTT_CUDA(cudaMallocHost(&h->h_ranks, cnt * sizeof(int)));
        h->h_ranks_cap = cnt;
    }
    QTT_CUDA(cudaMemcpyAsync(h->h_ranks, x->ranks, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    QTT_CUDA(cudaStreamSynchronize(st));
    out.assign(h->h_ranks, h->h_ranks + cnt);
    return QTT_OK;
}

// ------------------------------------------------------------------------------------------------
// materialised ops
// ------------------------------------------------------------------------------------------------
extern "C" int qtt_matmul_batched(qtt_handle_t h, const qtt_batch* a, const qtt_mpo* as, const qtt_batch* x,
                                  const int* m, qtt_batch* y, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE((a != nullptr) != (as != nullptr), QTT_ERR_INVALID, "exactly one of a / a_shared");
    QTT_REQUIRE(x && y && m, QTT_ERR_INVALID, "null argument");
    const int L = x->L, N = x->N;
    QTT_REQUIRE(L >= 1 && L <= QTT_MAX_L && N >= 1 && N <= 65535, QTT_ERR_INVALID, "L or N out of range");
    QTT_REQUIRE(y->L == L && y->N == N && (as ? as->L == L : (a->L == L && a->N == N)), QTT_ERR_INVALID, "shape mismatch");
    WsScope ws(h);
    int* d_ar = nullptr;
    if (as) {
        QTT_CHECK(ws.alloc_t(&d_ar, L + 1));
        QTT_CUDA(cudaMemcpyAsync(d_ar, as->ranks, (L + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    matmul_ranks_kernel<<<grid1d((ll)N * (L + 1), 256), 256, 0, st>>>(a ? a->ranks : nullptr, d_ar, x->ranks, y->ranks, L, N);
    QTT_CHECK(check_launch(h, "matmul_ranks"));
    for (int k = 0; k < L; ++k) {
        ContractParams p;
        memset(&p, 0, sizeof p);
        int an = as ? as->n_out[k] * as->n_in[k] : a->n[k];
        if (m[k] == 0) {        // Hadamard core product: same mode index on both sides, no contraction
            QTT_REQUIRE(!as && an == x->n[k] && y->n[k] == an, QTT_ERR_INVALID, "hadamard product needs equal mode sizes (per-instance operands)");
            p.na = an; p.m = 1; p.nx = 1; p.hadamard = 1;
        } else {
        QTT_REQUIRE(an % m[k] == 0 && x->n[k] % m[k] == 0, QTT_ERR_INVALID, "contracted mode does not divide");
        p.na = an / m[k]; p.m = m[k]; p.nx = x->n[k] / m[k];
        }
        QTT_REQUIRE(y->n[k] == p.na * p.nx, QTT_ERR_INVALID, "output mode size mismatch");
        if (as) { QTT_REQUIRE(as->n_in[k] == m[k], QTT_ERR_INVALID, "operator n_in != m");
                  p.A = as->cores[k]; p.a_shared = 1; p.pa = as->ranks[k]; p.qa = as->ranks[k + 1]; }
        else { p.A = a->cores[k]; p.a_slot = a->slot[k]; p.a_ranks = a->ranks; }
        p.X = x->cores[k]; p.x_slot = x->slot[k]; p.x_ranks = x->ranks;
        p.Y = y->cores[k]; p.y_slot = y->slot[k]; p.y_idx = 1;
        p.rstride = L + 1; p.k = k; p.coef = nullptr; p.coef_scale = 1.0;
        dim3 grid(grid1d(y->slot[k], 256, 1024), N);
        contract_kernel<<<grid, 256, 0, st>>>(p);
        QTT_CHECK(check_launch(h, "contract"));
    }
    return QTT_OK;
}

extern "C" int qtt_axpby_batched(qtt_handle_t h, const double* alpha, const qtt_batch* a, const double* beta,
                                 const qtt_batch* b, qtt_batch* y, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(a && b && y, QTT_ERR_INVALID, "null argument");
    const int L = a->L, N = a->N;
    QTT_REQUIRE(b->L == L && y->L == L && b->N == N && y->N == N && N <= 65535, QTT_ERR_INVALID, "shape mismatch");
    axpby_ranks_kernel<<<grid1d((ll)N * (L + 1), 256), 256, 0, st>>>(a->ranks, b->ranks, y->ranks, L, N);
    QTT_CHECK(check_launch(h, "axpby_ranks"));
    for (int k = 0; k < L; ++k) {
        QTT_REQUIRE(a->n[k] == b->n[k] && a->n[k] == y->n[k], QTT_ERR_INVALID, "mode size mismatch");
        AxpbyParams p;
        p.A = a->cores[k]; p.a_slot = a->slot[k]; p.a_ranks = a->ranks;
        p.B = b->cores[k]; p.b_slot = b->slot[k]; p.b_ranks = b->ranks;
        p.Y = y->cores[k]; p.y_slot = y->slot[k];
        p.rstride = L + 1; p.k = k; p.n = a->n[k]; p.L = L; p.alpha = alpha; p.beta = beta;
        dim3 grid(grid1d(y->slot[k], 256, 1024), N);
        axpby_kernel<<<grid, 256, 0, st>>>(p);
        QTT_CHECK(check_launch(h, "axpby"));
    }
    return QTT_OK;
}

// Function encoders, batched (reference utils.py:188-225) - see tt_ops.cuh.
extern "C" int qtt_encode_linear_batched(qtt_handle_t h, int nc, const double* coef, const double* const* shared_cores,
                                         const int* ranks, qtt_batch* y, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(coef && shared_cores && ranks && y && nc >= 1, QTT_ERR_INVALID, "null argument");
    const int L = y->L, N = y->N;
    QTT_REQUIRE(L >= 1 && L <= QTT_MAX_L && N >= 1 && N <= 65535 && ranks[0] == 1 && ranks[L] == 1, QTT_ERR_INVALID, "L, N or ranks out of range");
    std::vector<int> hr((size_t)N * (L + 1));
    for (int i = 0; i < N; ++i) for (int k = 0; k <= L; ++k) hr[(size_t)i * (L + 1) + k] = ranks[k];
    QTT_CUDA(cudaMemcpyAsync(y->ranks, hr.data(), hr.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < L; ++k) {
        EncodeLinearParams p;
        p.coef = coef; p.nc = nc; p.src = shared_cores[k]; p.count = (ll)ranks[k] * y->n[k] * ranks[k + 1];
        p.dst = y->cores[k]; p.dst_slot = y->slot[k]; p.k = k;
        QTT_REQUIRE(y->slot[k] >= p.count, QTT_ERR_SLOT, "output slot too small");
        encode_linear_kernel<<<dim3(grid1d(p.count, 128, 64), N), 128, 0, st>>>(p);
        QTT_CHECK(check_launch(h, "encode_linear"));
    }
    QTT_CUDA(cudaStreamSynchronize(st));       // hr is a host temporary
    return QTT_OK;
}

extern "C" int qtt_encode_trig_batched(qtt_handle_t h, const double* abnu, qtt_batch* y, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(abnu && y, QTT_ERR_INVALID, "null argument");
    const int Lc = y->L, N = y->N, L = Lc - 1;              // L level cores + the trailing corner core
    QTT_REQUIRE(Lc >= 2 && Lc <= QTT_MAX_L && N >= 1, QTT_ERR_INVALID, "need at least one level");
    std::vector<int> hr((size_t)N * (Lc + 1), 2);
    for (int i = 0; i < N; ++i) { hr[(size_t)i * (Lc + 1)] = 1; hr[(size_t)i * (Lc + 1) + Lc] = 1; }
    QTT_CUDA(cudaMemcpyAsync(y->ranks, hr.data(), hr.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < Lc; ++k) {
        QTT_REQUIRE(y->n[k] == 2, QTT_ERR_INVALID, "trigonometric trains have mode size 2");
        QTT_REQUIRE(y->slot[k] >= ((k == 0 || k == L) ? 4 : 8), QTT_ERR_SLOT, "output slot too small");
        EncodeTrigParams p; p.abnu = abnu; p.dst = y->cores[k]; p.dst_slot = y->slot[k]; p.k = k; p.L = L;
        encode_trig_kernel<<<grid1d(N, 128), 128, 0, st>>>(p, N);
        QTT_CHECK(check_launch(h, "encode_trig"));
    }
    QTT_CUDA(cudaStreamSynchronize(st));
    return QTT_OK;
}

static int inner_plain(qtt_handle_s* h, cudaStream_t st, const qtt_batch* a, const qtt_batch* b, double* out) {
    const int L = a->L, N = a->N;
    // shared-memory budget from the actual ranks: E <= max bond product, F <= ra * n * rb
    std::vector<int> ra, rb;
    QTT_CHECK(read_ranks_host(h, st, a, ra));
    QTT_CHECK(read_ranks_host(h, st, b, rb));
    ll ecap = 1, fcap = 1;
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < L; ++k) {
            const int* pa = &ra[(size_t)i * (L + 1)]; const int* pb = &rb[(size_t)i * (L + 1)];
            ecap = std::max<ll>(ecap, std::max<ll>((ll)pa[k] * pb[k], (ll)pa[k + 1] * pb[k + 1]));
            fcap = std::max<ll>(fcap, (ll)pa[k] * a->n[k] * pb[k + 1]);
        }
    const size_t smem = (size_t)(2 * ecap + fcap) * sizeof(double);
    QTT_REQUIRE(smem <= 200 * 1024, QTT_ERR_UNSUPPORTED, "inner product: ranks too large for the shared-memory kernel");
    InnerParams p;
    memset(&p, 0, sizeof p);
    for (int k = 0; k < L; ++k) {
        QTT_REQUIRE(a->n[k] == b->n[k], QTT_ERR_INVALID, "mode size mismatch");
        p.a_cores[k] = a->cores[k]; p.a_slot[k] = a->slot[k];
        p.b_cores[k] = b->cores[k]; p.b_slot[k] = b->slot[k]; p.n[k] = a->n[k];
    }
    p.a_ranks = a->ranks; p.b_ranks = b->ranks; p.L = L; p.rstride = L + 1; p.out = out; p.e_cap = (int)ecap;
    inner_kernel<<<N, 256, smem, st>>>(p);
    return check_launch(h, "inner");
}

extern "C" int qtt_inner_batched(qtt_handle_t h, const qtt_batch* a, const qtt_mpo* op, const qtt_batch* b,
                                 double* out, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(a && b && out, QTT_ERR_INVALID, "null argument");
    const int L = a->L, N = a->N;
    QTT_REQUIRE(b->L == L && b->N == N && L <= QTT_MAX_L && N <= 65535, QTT_ERR_INVALID, "shape mismatch");
    if (!op) return inner_plain(h, st, a, b, out);
    // bilinear form <a, op @ b>: materialise op @ b in the workspace (tensormethods.py:183-186), then contract
    QTT_REQUIRE(op->L == L, QTT_ERR_INVALID, "operator length mismatch");
    std::vector<int> rb;
    QTT_CHECK(read_ranks_host(h, st, b, rb));
    std::vector<int> bmax(L + 1, 1);
    for (int i = 0; i < N; ++i) for (int k = 0; k <= L; ++k) bmax[k] = std::max(bmax[k], rb[(size_t)i * (L + 1) + k]);
    WsScope ws(h);
    std::vector<double*> cores(L); std::vector<ll> slot(L); std::vector<int> nout(L), m(L);
    for (int k = 0; k < L; ++k) {
        QTT_REQUIRE(op->n_in[k] == b->n[k] && op->n_out[k] == a->n[k], QTT_ERR_INVALID, "operator mode sizes");
        nout[k] = op->n_out[k]; m[k] = op->n_in[k];
        slot[k] = (ll)op->ranks[k] * bmax[k] * nout[k] * op->ranks[k + 1] * bmax[k + 1];
        QTT_CHECK(ws.alloc_t(&cores[k], (size_t)slot[k] * N));
    }
    int* ranks = nullptr;
    QTT_CHECK(ws.alloc_t(&ranks, (size_t)N * (L + 1)));
    qtt_batch y; y.L = L; y.N = N; y.n = nout.data(); y.ranks = ranks; y.cores = cores.data(); y.slot = slot.data();
    QTT_CHECK(qtt_matmul_batched(h, nullptr, op, b, m.data(), &y, stream));
    return inner_plain(h, st, a, &y, out);
}

// ------------------------------------------------------------------------------------------------
// the sweeps
// ------------------------------------------------------------------------------------------------
struct Term {
    const qtt_mpo* op; const qtt_batch* x; const double* coef; double cs;
    std::vector<int> xr;     // ranks of x for this group, [L+1]
    std::vector<int> pr;     // bond dims of the term, [L+1]
    std::vector<int> off;    // column offset of the term inside the summed bond, [L+1]
    // Head compression of an operator term (exact): the left bonds of Op @ x cannot exceed prod(n_0..n_{k-1}),
    // so cores 0..kc-1 are replaced by small dense cores with orthonormal columns (left-to-right QR, no
    // truncation) and core kc by the dense Z = R_{kc-1} . (Op_kc (x) x_kc).  kc == 0: no head.
    int kc = 0;
    std::vector<int> prod;   // uncompressed bonds of the term, [L+1]
    std::vector<int> c;      // compressed bonds, [kc+1]
    std::vector<double*> head; std::vector<ll> head_stride;   // head cores k < kc, (c[k], n_k, c[k+1]) C-order, group-local
    double* Z = nullptr; ll z_stride = 0;                     // dense core kc, column-major (c[kc] n_kc) x prod[kc+1]
};
enum SweepMode { MODE_ROUND = 0, MODE_ORTH = 1, MODE_NORM = 2 };
struct SweepArgs {
    int L; const int* n; std::vector<Term> terms; qtt_batch* y; const int* caps; double eps;
    double* norm2; int* status; SweepMode mode;
};
struct SweepPlan {
    std::vector<int> p, q, a, sb;       // summed bond dims, orthogonalised dims, rows of M_k, truncation bounds
    ll v_elems = 0, m_elems = 0, r_elems = 0, t_elems = 0, x_elems = 0, g_elems = 0;
    int wk_cols = 0, nvec_max = 0, tb_panels = 0, ts_blocks = 0, tau_len = 0;
    std::vector<ll> v_off, tb_off, ts_off, tau_off;
    ll per_inst_bytes = 0;
};

static int leaf_width(int rows) {
    // widest sub-panel (<= QR_W) whose shared-memory footprint fits one CTA
    const size_t cap = 220 * 1024;
    for (int w = QR_W; w >= 1; --w)
        if (qr_leaf_smem(rows, w, QR_THREADS_BIG) <= cap) return w;
    return 0;
}

// Tallest panel the leaves can factorise: one CTA (shared-memory leaf) or a cluster of up to max_cluster CTAs, rows split.
static bool leaf_supports(const qtt_handle_s* h, int rows) {
    return leaf_width(rows) > 0 || (ll)rows <= (ll)h->max_cluster * (qr_leaf3_max_slice(QR_THREADS_BIG) - 16);
}

static int make_plan(qtt_handle_s* h, const SweepArgs& A, SweepPlan& P) {
    const int L = A.L;
    P.p.assign(L + 1, 1); P.q.assign(L + 1, 1); P.a.assign(L, 0); P.sb.assign(L + 1, 1);
    for (int k = 1; k < L; ++k) { int s = 0; for (auto& t : A.terms) s += t.pr[k]; P.p[k] = s; }
    for (int k = L - 1; k >= 1; --k) { P.a[k] = A.n[k] * P.q[k + 1]; P.q[k] = std::min(P.a[k], P.p[k]); }
    P.a[0] = A.n[0] * P.q[1];
    for (int k = 0; k < L - 1; ++k) {
        int cap = A.caps ? A.caps[k] : 2147483647;
        P.sb[k + 1] = std::min(std::min(P.sb[k] * A.n[k], P.q[k + 1]), std::max(cap, 0));
    }
    P.v_off.assign(L, 0); P.tb_off.assign(L, 0); P.ts_off.assign(L, 0); P.tau_off.assign(L, 0);
    ll voff = 0, tboff = 0, tsoff = 0, tauoff = 0;
    const bool keep = (A.mode != MODE_NORM);
    for (int k = 1; k < L; ++k) {
        const ll ve = (ll)P.a[k] * P.q[k];
        if (keep) { P.v_off[k] = voff; voff += ve; } else { P.v_off[k] = 0; voff = std::max(voff, ve); }
        const int panels = (int)tb_units_before(true, P.a[k], P.q[k], qr_npan(P.q[k]));   // T slots: one per panel, or per row block (tall-skinny)
        const int blocks = P.q[k];   // upper bound on the number of sub-panels (width >= 1)
        if (keep) { P.tb_off[k] = tboff; tboff += (ll)panels * QR_NB * QR_NB; P.ts_off[k] = tsoff; tsoff += (ll)blocks * 64; P.tau_off[k] = tauoff; tauoff += P.q[k]; }
        else { tboff = std::max<ll>(tboff, (ll)panels * QR_NB * QR_NB); tsoff = std::max<ll>(tsoff, (ll)blocks * 64); tauoff = std::max<ll>(tauoff, P.q[k]); }
        P.m_elems = std::max(P.m_elems, (ll)P.a[k] * P.p[k]);
        P.r_elems = std::max(P.r_elems, (ll)P.q[k] * P.p[k]);
        P.wk_cols = std::max(P.wk_cols, std::max(P.p[k], P.q[k]));
        if (!leaf_supports(h, P.a[k])) { h->err = "QR panel taller than the cluster leaf supports (16 x ~2500 rows)"; return QTT_ERR_UNSUPPORTED; }
    }
    P.m_elems = std::max(P.m_elems, (ll)P.a[0]);
    for (auto& t : A.terms)
        if (t.op)
            for (int k = (t.kc > 0 ? t.kc + 1 : 0); k < L - 1; ++k)   // T buffer of the MPO push at core k (uses R of core k+1)
                P.t_elems = std::max(P.t_elems, (ll)t.xr[k] * t.op->n_in[k] * t.op->ranks[k + 1] * P.q[k + 1]);
    P.v_elems = voff; P.tb_panels = (int)(tboff / (QR_NB * QR_NB)); P.ts_blocks = (int)(tsoff / 64); P.tau_len = (int)tauoff;
    for (int k = 0; k < L; ++k) {
        int s = (A.mode == MODE_ROUND) ? P.sb[k] : 0;
        P.x_elems = std::max(P.x_elems, (ll)P.a[k] * s);
    }
    for (int k = 0; k < L - 1; ++k) {
        P.g_elems = std::max(P.g_elems, (ll)P.sb[k + 1] * P.q[k + 1]);
        P.nvec_max = std::max(P.nvec_max, std::min(P.sb[k] * A.n[k], P.q[k + 1]));
    }
    if (A.mode == MODE_ROUND && P.nvec_max > 16384) { h->err = "SVD side exceeds the Jacobi kernel limit"; return QTT_ERR_UNSUPPORTED; }
    P.per_inst_bytes = 8 * (P.v_elems + tboff + tsoff + tauoff + P.m_elems + P.r_elems + P.t_elems + 2 * P.x_elems +
                            P.g_elems + (ll)P.nvec_max * P.nvec_max + (ll)GRAM_SPLITS * QR_NB * QR_NB);
    for (auto& t : A.terms) {                     // head-compression scratch (see compress_head)
        ll extra = 0;
        for (int k = 0; k < t.kc; ++k) extra += 4 * (ll)t.c[k] * A.n[k] * std::max(t.prod[k + 1], t.c[k + 1]);
        for (int k = 1; k <= t.kc; ++k) extra = std::max(extra, 2 * (ll)t.xr[k] * t.c[k] * t.op->n_out[k] * t.op->n_in[k] * t.op->ranks[k + 1]);
        if (t.kc > 0) extra += (ll)t.c[t.kc] * A.n[t.kc] * t.prod[t.kc + 1];
        P.per_inst_bytes += 8 * extra;
    }
    return QTT_OK;
}

struct SweepBufs {
    double *V, *Tb, *Ts, *tau, *M, *R, *T, *X0, *X1, *Gc, *JT, *Gram;
    ll sV, sTb, sTs, stau, sM, sR, sT, sX, sG, sJT, sGram;          // per-instance strides (elements)
};

struct QrBufs {
    double* M; ll sM; double* V; ll sV; double* tau; ll stau; double* Tb; ll sTb; double* Ts; ll sTs; double* Gram; ll sGram;
    bool ts;      // Tb is sized for tall-skinny panels (one T per row block), see ts_nblk
};
static QrBufs qr_bufs_for_core(const SweepPlan& P, const SweepBufs& Bf, int k) {
    QrBufs q;
    q.M = Bf.M; q.sM = Bf.sM; q.V = Bf.V + P.v_off[k]; q.sV = Bf.sV; q.tau = Bf.tau + P.tau_off[k]; q.stau = Bf.stau;
    q.Tb = Bf.Tb + P.tb_off[k]; q.sTb = Bf.sTb; q.Ts = Bf.Ts + P.ts_off[k]; q.sTs = Bf.sTs; q.Gram = Bf.Gram; q.sGram = Bf.sGram;
    q.ts = true;
    return q;
}

// Blocked Householder QR of the working matrices (a x b column-major in B.M) of every instance of the group.
static int qr_factor(qtt_handle_s* h, cudaStream_t st, const QrBufs& B, int G, int a, int b, bool need_last_T) {
    const int kmax = std::min(a, b);
    double* V = B.V;
    double* tau = B.tau;
    double* Tb = B.Tb;
    double* Ts = B.Ts;
    const int wsub = leaf_width(a);
    for (int pi = 0; pi < qr_npan(kmax); ++pi) {
        const int j0 = qr_pstart(kmax, pi), jb = qr_pwidth(kmax, pi);
        // inner sub-panels; QR_NB is a multiple of every admissible wsub only if wsub divides 32:
        // use the largest power of two <= wsub so sub-panels tile the outer panel exactly
        int w2 = 1; while (w2 * 2 <= wsub) w2 *= 2;
        const int rows = a - j0;
        const double* Vp = V + (ll)j0 * a + j0;
        double* Tp = Tb + tb_units_before(B.ts, a, kmax, pi) * QR_NB * QR_NB;
        const int nc = b - (j0 + jb), c0 = j0 + jb;
        const int nblk = ts_nblk(B.ts, rows, jb);
        h->leaf_flops += (2.0 * rows * jb * jb + 2.0 * rows * jb * std::min(w2, jb)) * G;   // left-looking sub-panels of width w2 (see qtt_b200.h)
        if (nblk >= 2) {
            // tall-skinny panel: row blocks factorised in registers, one T per block; trailing update in the same form
            TsLeafParams lp;
            lp.M = B.M; lp.m_stride = B.sM; lp.ld = a; lp.V = V; lp.v_stride = B.sV; lp.ldv = a;
            lp.Tq = Tp; lp.t_stride = B.sTb; lp.a = a; lp.j0 = j0; lp.nblk = nblk;
            if (g_ts_leaf_gen == 1) tsqr_leaf_kernel<<<G, TS_THREADS, tsqr_leaf_smem(), st>>>(lp);
            else tsqr2_leaf_kernel<<<G, TS2_THREADS, tsqr2_leaf_smem(), st>>>(lp);
            QTT_CHECK(check_launch(h, "qr_leaf_ts"));
            if (nc > 0) {
                LarfbTsParams tp; memset(&tp, 0, sizeof tp);
                tp.V = Vp; tp.v_stride = B.sV; tp.ldv = a;
                tp.Tq = Tp; tp.t_stride = B.sTb;
                tp.C = B.M + (ll)c0 * a + j0; tp.c_stride = B.sM; tp.ldc = a; tp.c_idx = 0;
                tp.rows = rows; tp.nblk = nblk; tp.nc = nc; tp.trans_t = 1;
                QTT_CHECK(larfb_ts_launch(h, st, tp, G, nc));
            }
            continue;
        }
        const bool force_cl = g_force_cluster && rows >= 64;
        const bool leaf2 = (w2 == QR_W) && !QTT_ENV_FLAG("QTT_OLD_LEAF") && !force_cl &&
                           qr_leaf2_smem(rows, rows > g_leaf_small_rows ? QR_THREADS_BIG : QR_THREADS_SMALL) <= 220 * 1024;
        // tall panels: rows split over C CTAs per instance - a thread-block cluster (distributed shared memory, any batch size;
        // default) or, for A/B, the cooperative launch of round 1 (QTT_COOP_LEAF=1, needs C * G resident CTAs)
        const int slice_cap = qr_leaf3_max_slice(QR_THREADS_BIG) - 16;      // room for the rounding of the slice to 4 (mod 16)
        int Cc = 1; while ((ll)Cc * slice_cap < rows) Cc *= 2;
        const bool tall = !leaf2 && (w2 < QR_W || force_cl);
        const bool coop = tall && QTT_ENV_FLAG("QTT_COOP_LEAF") && !QTT_ENV_FLAG("QTT_NO_COOP") && (ll)Cc * G <= h->sm_count;
        const bool leaf3 = tall && !QTT_ENV_FLAG("QTT_NO_CLUSTER_LEAF") && (coop || Cc <= h->max_cluster);
        if (leaf3) {
            if (!coop && Cc == 1) Cc = 2;                 // forced on a short panel (tests): still a real cluster
            // rows per CTA = column stride of its shared-memory slice: 4 (mod 16), so that the fragment-shaped reads of the DMMA passes
            // (lane (g, t4) -> column g, row t4) are bank-conflict free (a multiple of 16 puts all 8 columns into the same banks)
            const int slice0 = (int)ceil_div(rows, Cc);
            const int slice = slice0 + ((20 - (slice0 & 15)) & 15);
            const int cs = (st == h->st2) ? 1 : 0;
            if (coop) {
                const size_t need = (size_t)G * 2 * (Cc + 1) * 1024 * sizeof(double);
                // one scratch per stream: the head compression (secondary stream) and the main sweep may both run cooperative leaves
                if (h->coop_cap[cs] < need) {
                    if (h->coop_scratch[cs]) { cudaStreamSynchronize(st); cudaFree(h->coop_scratch[cs]); h->coop_scratch[cs] = nullptr; h->coop_cap[cs] = 0; }
                    QTT_CUDA(cudaMalloc(&h->coop_scratch[cs], need));
                    h->coop_cap[cs] = need;
                }
            }
            for (int i0 = j0; i0 < j0 + jb; i0 += QR_W) {
                Leaf3Params lp;
                lp.M = B.M; lp.m_stride = B.sM; lp.ld = a; lp.V = V; lp.v_stride = B.sV; lp.ldv = a;
                lp.tau = tau; lp.tau_stride = B.stau; lp.Tp = Tp; lp.t_stride = B.sTb;
                lp.scratch = coop ? h->coop_scratch[cs] : nullptr; lp.scratch_stride = (ll)2 * (Cc + 1) * 1024;
                lp.a = a; lp.j0 = j0; lp.i0 = i0; lp.w = std::min(QR_W, j0 + jb - i0); lp.last = (i0 + QR_W >= j0 + jb) ? 1 : 0;
                lp.C = Cc; lp.slice = slice;
                if (coop) {
                    void* args[] = {&lp};
                    QTT_CUDA(cudaLaunchCooperativeKernel((void*)qr_leaf3_kernel<QR_THREADS_BIG, false>, dim3(Cc, G), dim3(QR_THREADS_BIG), args,
                                                         qr_leaf3_smem(slice, QR_THREADS_BIG), st));
                } else {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(Cc, G); cfg.blockDim = dim3(QR_THREADS_BIG);
                    cfg.dynamicSmemBytes = qr_leaf3_smem(slice, QR_THREADS_BIG); cfg.stream = st;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = Cc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    QTT_CUDA(cudaLaunchKernelEx(&cfg, qr_leaf3_kernel<QR_THREADS_BIG, true>, lp));
                }
                QTT_CHECK(check_launch(h, "qr_leaf3"));
            }
        } else if (leaf2) {
            // second-generation leaf: combined block reflector of the earlier sub-panels, T of the panel built in place
            // one launch walks all sub-panels of the outer panel (the CTA of an instance depends only on its own earlier sub-panels):
            // 8x fewer launches; QTT_LEAF_LOOP=0 restores one launch per sub-panel
            const int step = g_leaf_loop ? jb : QR_W;
            for (int i0 = j0; i0 < j0 + jb; i0 += step) {
                Leaf2Params lp;
                lp.M = B.M; lp.m_stride = B.sM; lp.ld = a; lp.V = V; lp.v_stride = B.sV; lp.ldv = a;
                lp.tau = tau; lp.tau_stride = B.stau; lp.Tp = Tp; lp.t_stride = B.sTb;
                lp.a = a; lp.j0 = j0; lp.i0 = i0; lp.w = std::min(QR_W, j0 + jb - i0); lp.last = (i0 + QR_W >= j0 + jb) ? 1 : 0;
                lp.nsub = g_leaf_loop ? (int)ceil_div(jb, QR_W) : 1; lp.jend = j0 + jb;
                const int nt_leaf = (rows > g_leaf_small_rows) ? ((g_leaf_nt == 512) ? 512 : QR_THREADS_BIG) : QR_THREADS_SMALL;
                lp.vres = (lp.nsub > 1 && g_leaf_vres && qr_leaf2_smem(rows, nt_leaf, true) <= 220 * 1024) ? 1 : 0;
                const size_t leaf_smem = qr_leaf2_smem(rows, nt_leaf, lp.vres != 0);
                if (nt_leaf == 512) qr_leaf2_kernel<512><<<G, 512, leaf_smem, st>>>(lp);
                else if (nt_leaf == QR_THREADS_BIG) qr_leaf2_kernel<QR_THREADS_BIG><<<G, QR_THREADS_BIG, leaf_smem, st>>>(lp);
                else qr_leaf2_kernel<QR_THREADS_SMALL><<<G, QR_THREADS_SMALL, leaf_smem, st>>>(lp);
                QTT_CHECK(check_launch(h, "qr_leaf2"));
            }
        } else {
        for (int i0 = j0; i0 < j0 + jb; i0 += w2) {
            LeafParams lp;
            lp.M = B.M; lp.m_stride = B.sM; lp.ld = a;
            lp.V = V; lp.v_stride = B.sV; lp.ldv = a;
            lp.tau = tau; lp.tau_stride = B.stau;
            lp.Ts = Ts; lp.ts_stride = B.sTs;
            lp.a = a; lp.j0 = j0; lp.i0 = i0; lp.w = std::min(w2, j0 + jb - i0); lp.ws = w2;
            if (a - j0 > 768) qr_leaf_kernel<QR_THREADS_BIG><<<G, QR_THREADS_BIG, qr_leaf_smem(a - j0, w2, QR_THREADS_BIG), st>>>(lp);
            else qr_leaf_kernel<QR_THREADS_SMALL><<<G, QR_THREADS_SMALL, qr_leaf_smem(a - j0, w2, QR_THREADS_SMALL), st>>>(lp);
            QTT_CHECK(check_launch(h, "qr_leaf"));
        }
        const bool need_T = (nc > 0) || need_last_T;
        if (need_T) {   // T of the outer panel: split-row Gram partials, then the row-parallel dlarft recurrence
            const int nsplit = std::max(1, std::min(GRAM_SPLITS, (rows + 63) / 64));
            gram_kernel<<<dim3(nsplit, G), LARFB_THREADS, 0, st>>>(Vp, B.sV, a, rows, jb, B.Gram, B.sGram, nsplit);
            QTT_CHECK(check_launch(h, "gram"));
            h->gemm_flops += 2.0 * rows * 64.0 * 64.0 * G;
            larft_kernel<<<G, 256, (2 * QR_NB * QR_NB + 2 * QR_NB) * sizeof(double), st>>>(B.Gram, B.sGram, nsplit, tau, B.stau, j0, Tp, B.sTb, jb, QR_NB);
            QTT_CHECK(check_launch(h, "larft"));
        }
        }
        if (nc > 0) {
            LarfbParams lp; memset(&lp, 0, sizeof lp);
            lp.V = Vp; lp.v_stride = B.sV; lp.ldv = a;
            lp.T = Tp; lp.t_stride = B.sTb; lp.ldt = QR_NB;
            lp.C = B.M + (ll)c0 * a + j0; lp.c_stride = B.sM; lp.ldc = a; lp.c_idx = 0;
            lp.rows = rows; lp.jb = jb; lp.nc = nc; lp.trans_t = 1;
            QTT_CHECK(larfb_launch(h, st, lp, G, nc));
        }
    }
    return QTT_OK;
}

// X <- Q_k X for the stored reflectors of core k; X is (a x s) column-major, s per instance (ranks table) or fixed.
static int apply_q(qtt_handle_s* h, cudaStream_t st, const QrBufs& B, int G, int a, int kq,
                   double* X, ll x_stride, int x_idx, const int* d_idx, int s_bound, const int* ranks, int rstride, int kcol) {
    const double* V = B.V;
    const double* Tb = B.Tb;
    const int npan = qr_npan(kq);
    for (int pi = npan - 1; pi >= 0; --pi) {
        const int j0 = qr_pstart(kq, pi), jb = qr_pwidth(kq, pi), rows = a - j0;
        const double* Vp = V + (ll)j0 * a + j0;
        const double* Tp = Tb + tb_units_before(B.ts, a, kq, pi) * QR_NB * QR_NB;
        const int nblk = ts_nblk(B.ts, rows, jb);
        if (nblk >= 2) {
            LarfbTsParams tp; memset(&tp, 0, sizeof tp);
            tp.V = Vp; tp.v_stride = B.sV; tp.ldv = a;
            tp.Tq = Tp; tp.t_stride = B.sTb;
            tp.C = X + j0; tp.c_stride = x_stride; tp.ldc = a; tp.c_idx = x_idx;
            tp.rows = rows; tp.nblk = nblk; tp.nc = s_bound; tp.trans_t = 0;
            tp.idx = d_idx; tp.n_dyn = ranks; tp.n_dyn_stride = rstride; tp.n_dyn_off = kcol;
            QTT_CHECK(larfb_ts_launch(h, st, tp, G, s_bound));
            continue;
        }
        LarfbParams lp; memset(&lp, 0, sizeof lp);
        lp.V = Vp; lp.v_stride = B.sV; lp.ldv = a;
        lp.T = Tp; lp.t_stride = B.sTb; lp.ldt = QR_NB;
        lp.C = X + j0; lp.c_stride = x_stride; lp.ldc = a; lp.c_idx = x_idx;
        lp.rows = rows; lp.jb = jb; lp.nc = s_bound; lp.trans_t = 0;
        lp.idx = d_idx; lp.n_dyn = ranks; lp.n_dyn_stride = rstride; lp.n_dyn_off = kcol;
        QTT_CHECK(larfb_launch(h, st, lp, G, s_bound));
    }
    return QTT_OK;
}

// Column block of M_k contributed by one term, given R of core k+1 (or directly for the last core).
static int form_block(qtt_handle_s* h, cudaStream_t st, const SweepArgs& A, const SweepPlan& P, const SweepBufs& B,
                      int G, const int* d_idx, const Term& t, int k, bool accumulate, bool r_in_place = false) {
    const int L = A.L;
    const int a = P.a[k];
    const int q = P.q[k + 1];                       // rows of R_{k+1} (= 1 for the last core)
    double* Mblk = B.M + (ll)t.off[k] * a;          // column offset of the term
    const double* coef = (k == 0) ? t.coef : nullptr;
    const double cs = (k == 0) ? t.cs : 1.0;
    if (k == L - 1) {
        // no factor to push: the block is the term's own last core, (p, n, 1) C-order == column-major n x p
        if (!t.op) {
            CopyParams cp; memset(&cp, 0, sizeof cp);
            cp.src = t.x->cores[k]; cp.src_stride = t.x->slot[k]; cp.src_idx = 1;
            cp.dst = Mblk; cp.dst_stride = B.sM; cp.dst_idx = 0; cp.idx = d_idx;
            cp.count = (ll)t.pr[k] * A.n[k]; cp.coef = coef; cp.coef_scale = cs;
            dim3 grid(grid1d(cp.count, 256, 256), G);
            copy_kernel<<<grid, 256, 0, st>>>(cp);
            return check_launch(h, "copy last core");
        }
        ContractParams p; memset(&p, 0, sizeof p);
        p.A = t.op->cores[k]; p.a_shared = 1; p.pa = t.op->ranks[k]; p.qa = t.op->ranks[k + 1];
        p.X = t.x->cores[k]; p.x_slot = t.x->slot[k]; p.x_ranks = t.x->ranks;
        p.Y = Mblk; p.y_slot = B.sM; p.y_idx = 0;
        p.rstride = L + 1; p.k = k; p.na = t.op->n_out[k]; p.m = t.op->n_in[k]; p.nx = 1;
        p.idx = d_idx; p.coef = coef; p.coef_scale = cs;
        dim3 grid(grid1d((ll)t.pr[k] * A.n[k], 256, 256), G);
        contract_kernel<<<grid, 256, 0, st>>>(p);
        return check_launch(h, "contract last core");
    }
    const double* Rblk = B.R + (ll)t.off[k + 1] * q;    // R is q x p_{k+1} column-major
    if (t.kc > 0 && k <= t.kc) {
        GemmDesc d = gemm_desc_zero();
        d.M = q; d.N = t.pr[k] * A.n[k]; d.K = t.pr[k + 1];
        d.A = Rblk; d.a_m = 1; d.a_k = q; d.a_s[0] = B.sR;
        if (k == t.kc) { d.B = t.Z; d.b_k = (ll)t.pr[k] * A.n[k]; d.b_n = 1; d.b_s[0] = t.z_stride; }       // dense Z, column-major
        else { d.B = t.head[k]; d.b_k = 1; d.b_n = t.pr[k + 1]; d.b_s[0] = t.head_stride[k]; }             // head core, C order
        d.C = Mblk; d.c_m = 1; d.c_n = q; d.c_s[0] = B.sM;
        d.alpha = 1.0; d.beta = accumulate ? 1.0 : 0.0;        // the coefficient was folded into head core 0
        d.nb[0] = G; d.idx = d_idx;
        d.tri_div = 1; d.tri_kperiod = 1 << 30; d.tri_mperiod = 1 << 30; d.tri_off = t.off[k + 1];
        return gemm_launch(h, st, d);
    }
    if (!t.op) {
        // M_k[(n, j), P] = sum_beta core[P, n, beta] R[j, beta]
        GemmDesc d = gemm_desc_zero();
        d.M = q; d.N = t.pr[k] * A.n[k]; d.K = t.pr[k + 1];
        d.A = Rblk; d.a_m = 1; d.a_k = q; d.a_s[0] = B.sR;
        d.B = t.x->cores[k]; d.b_k = 1; d.b_n = t.pr[k + 1]; d.b_s[0] = t.x->slot[k]; d.b_idx = 1;
        d.C = Mblk; d.c_m = 1; d.c_n = q; d.c_s[0] = B.sM;
        d.alpha = cs; d.alpha_vec = coef; d.beta = accumulate ? 1.0 : 0.0;
        d.nb[0] = G; d.idx = d_idx;
        d.tri_div = 1; d.tri_kperiod = 1 << 30; d.tri_mperiod = 1 << 30; d.tri_off = t.off[k + 1];   // R[j, off + beta] = 0 for off + beta < j
        return gemm_launch(h, st, d);
    }
    // operator term: T[r, m, R', j] = sum_r' x[r, m, r'] R[j, R' rr + r'];  M_k[(no, j), (Rl, r)] = sum_{m, R'} Op[Rl, no, m, R'] T[r, m, R', j]
    const int rl = t.xr[k], rr = t.xr[k + 1], Rl = t.op->ranks[k], Rr = t.op->ranks[k + 1];
    const int mi = t.op->n_in[k], no = t.op->n_out[k];
    {   // T[(m, R'), r, j] = sum_r' x[r, m, r'] R[j, R' rr + r']   (bandwidth bound, dedicated kernel)
        PushTParams tp; memset(&tp, 0, sizeof tp);
        tp.R = Rblk; tp.r_stride = B.sR; tp.q = q; tp.ldr = q; tp.from_m = 0;
        if (r_in_place) { tp.R = B.M; tp.r_stride = B.sM; tp.ldr = P.a[k + 1]; tp.from_m = 1; }   // single operator term: see sweep_group
        tp.X = t.x->cores[k]; tp.x_slot = t.x->slot[k];
        tp.T = B.T; tp.t_stride = B.sT;
        tp.rl = rl; tp.mi = mi; tp.rr = rr; tp.Rr = Rr; tp.idx = d_idx;
        dim3 grid(grid1d((ll)Rr * q, 256, 1024), G);
        QTT_REQUIRE((size_t)rl * mi * rr * sizeof(double) <= XCORE_SMEM_MAX, QTT_ERR_UNSUPPORTED,
                    "operator term: vector core larger than 200 KB (rank envelope: 79 at mode size 4, 111 at mode size 2)");
        push_t_kernel<<<grid, 256, (size_t)rl * mi * rr * sizeof(double), st>>>(tp);
        QTT_CHECK(check_launch(h, "push_t"));
    }
    // rows (r, j) merged: M_k[(no, j), (Rl, r)] = sum_{(m, R')} T[(m, R'), (r, j)] Op[Rl, no, (m, R')]
    GemmDesc d = gemm_desc_zero();
    d.M = rl * q; d.N = Rl; d.K = mi * Rr;
    d.A = B.T; d.a_m = 1; d.a_k = (ll)rl * q; d.a_s[0] = B.sT;
    d.B = t.op->cores[k]; d.b_k = 1; d.b_n = (ll)no * mi * Rr; d.b_s[1] = (ll)mi * Rr;
    d.C = Mblk; d.c_m = 1; d.c_n = (ll)rl * a; d.c_s[0] = B.sM; d.c_s[1] = q;
    d.c_msplit = q; d.c_mhi = a;
    d.alpha = cs; d.alpha_vec = coef; d.beta = accumulate ? 1.0 : 0.0;
    d.nb[0] = G; d.nb[1] = no; d.idx = d_idx;
    d.tri_div = rr; d.tri_kperiod = Rr; d.tri_mperiod = q; d.tri_off = t.off[k + 1];   // T[(m, R'), r, j] = 0 for off + R' rr + rr - 1 < j
    return gemm_launch(h, st, d);
}

// Exact left-to-right compression of the head of an operator term (see Term).  Uses its own scratch.
static int compress_head(qtt_handle_s* h, cudaStream_t st, WsScope& ws, const SweepArgs& A, Term& t, int G, const int* d_idx) {
    const int L = A.L, kc = t.kc;
    if (kc <= 0) return QTT_OK;
    // scratch sizes
    ll m_el = 1, v_el = 1, r_el = 1, s_el = 1; int kq_max = 1;
    for (int k = 0; k < kc; ++k) {
        const ll arows = (ll)t.c[k] * A.n[k];
        m_el = std::max(m_el, arows * t.prod[k + 1]);
        v_el = std::max(v_el, arows * t.c[k + 1]);
        r_el = std::max(r_el, (ll)t.c[k + 1] * t.prod[k + 1]);
        kq_max = std::max(kq_max, t.c[k + 1]);
    }
    for (int k = 1; k <= kc; ++k)
        s_el = std::max(s_el, (ll)t.xr[k] * t.c[k] * t.op->n_out[k] * t.op->n_in[k] * t.op->ranks[k + 1]);
    QrBufs Q; memset(&Q, 0, sizeof Q);
    double *Rh, *Sh, *Xh;
    const int panels = (int)ceil_div(kq_max, QR_NB);
    Q.sM = m_el; Q.sV = v_el; Q.stau = kq_max; Q.sTb = (ll)panels * QR_NB * QR_NB; Q.sTs = (ll)kq_max * 64; Q.sGram = (ll)GRAM_SPLITS * QR_NB * QR_NB;
    QTT_CHECK(ws.alloc_t(&Q.M, (size_t)Q.sM * G)); QTT_CHECK(ws.alloc_t(&Q.V, (size_t)Q.sV * G));
    QTT_CHECK(ws.alloc_t(&Q.tau, (size_t)Q.stau * G)); QTT_CHECK(ws.alloc_t(&Q.Tb, (size_t)Q.sTb * G));
    QTT_CHECK(ws.alloc_t(&Q.Ts, (size_t)Q.sTs * G)); QTT_CHECK(ws.alloc_t(&Q.Gram, (size_t)Q.sGram * G));
    QTT_CHECK(ws.alloc_t(&Rh, (size_t)r_el * G)); QTT_CHECK(ws.alloc_t(&Sh, (size_t)s_el * G)); QTT_CHECK(ws.alloc_t(&Xh, (size_t)v_el * G));
    t.head.assign(kc, nullptr); t.head_stride.assign(kc, 0);
    for (int k = 0; k < kc; ++k) {
        t.head_stride[k] = (ll)t.c[k] * A.n[k] * t.c[k + 1];
        QTT_CHECK(ws.alloc_t(&t.head[k], (size_t)t.head_stride[k] * G));
    }
    t.z_stride = (ll)t.c[kc] * A.n[kc] * t.prod[kc + 1];
    QTT_CHECK(ws.alloc_t(&t.Z, (size_t)t.z_stride * G));

    for (int k = 0; k <= kc; ++k) {
        const int cdim = t.c[k], no = t.op->n_out[k], mi = t.op->n_in[k], Rl = t.op->ranks[k], Rr = t.op->ranks[k + 1];
        const int rl = t.xr[k], rr = t.xr[k + 1];
        const int arows = cdim * no, bcols = t.prod[k + 1];
        double* Zdst = (k == kc) ? t.Z : Q.M;
        const ll zs = (k == kc) ? t.z_stride : Q.sM;
        if (k == 0) {
            ContractParams p; memset(&p, 0, sizeof p);          // Z_0 = coef * (Op_0 (x) x_0), column-major n_0 x prod[1]
            p.A = t.op->cores[0]; p.a_shared = 1; p.pa = 1; p.qa = Rr;
            p.X = t.x->cores[0]; p.x_slot = t.x->slot[0]; p.x_ranks = t.x->ranks;
            p.Y = Zdst; p.y_slot = zs; p.y_idx = 0; p.rstride = L + 1; p.k = 0; p.na = no; p.m = mi; p.nx = 1;
            p.idx = d_idx; p.coef = t.coef; p.coef_scale = t.cs; p.col_major_out = 1;
            contract_kernel<<<dim3(grid1d((ll)arows * bcols, 256, 256), G), 256, 0, st>>>(p);
            QTT_CHECK(check_launch(h, "head contract"));
        } else {
            // S[rl][c][(no mi Rr)] = sum_Rl R_{k-1}[c, (Rl, rl)] Op[Rl, (no mi Rr)]
            GemmDesc d = gemm_desc_zero();
            d.M = cdim; d.N = no * mi * Rr; d.K = Rl;
            d.A = Rh; d.a_m = 1; d.a_k = (ll)rl * cdim; d.a_s[0] = r_el; d.a_s[1] = cdim;
            d.B = t.op->cores[k]; d.b_k = (ll)no * mi * Rr; d.b_n = 1;
            d.C = Sh; d.c_m = (ll)no * mi * Rr; d.c_n = 1; d.c_s[0] = s_el; d.c_s[1] = (ll)cdim * no * mi * Rr;
            d.nb[0] = G; d.nb[1] = rl;
            QTT_CHECK(gemm_launch(h, st, d));
            HeadZParams zp; memset(&zp, 0, sizeof zp);
            zp.S = Sh; zp.s_stride = s_el; zp.X = t.x->cores[k]; zp.x_slot = t.x->slot[k]; zp.Z = Zdst; zp.z_stride = zs;
            zp.cdim = cdim; zp.no = no; zp.mi = mi; zp.Rr = Rr; zp.rl = rl; zp.rr = rr; zp.idx = d_idx;
            const size_t z_smem = ((size_t)rl * mi * rr + (size_t)rr * 32 * 33) * sizeof(double);
            if (rr <= HEADZ_MAX_RR && z_smem <= 100 * 1024) {
                const ll ztiles = ceil_div(arows, 32) * ceil_div(Rr, 32);
                head_z_kernel<<<dim3((unsigned)std::min<ll>(ztiles, 1024), G), 256, z_smem, st>>>(zp);
            } else {
                QTT_REQUIRE((size_t)rl * mi * rr * sizeof(double) <= XCORE_SMEM_MAX, QTT_ERR_UNSUPPORTED,
                            "operator term: vector core larger than 200 KB (rank envelope: 79 at mode size 4, 111 at mode size 2)");
                head_z_generic_kernel<<<dim3(grid1d((ll)arows * Rr, 256, 1024), G), 256, (size_t)rl * mi * rr * sizeof(double), st>>>(zp);
            }
            QTT_CHECK(check_launch(h, "head_z"));
        }
        if (k == kc) break;
        // Z_k = Q_k R_k: R_k (c[k+1] x prod[k+1]) feeds the next core, Q_k becomes head core k
        const int kq = t.c[k + 1];
        QTT_CHECK(qr_factor(h, st, Q, G, arows, bcols, true));
        extract_r_kernel<<<dim3(grid1d((ll)kq * bcols, 256, 512), G), 256, 0, st>>>(Q.M, Q.sM, arows, Rh, r_el, kq, bcols);
        QTT_CHECK(check_launch(h, "head extract_r"));
        InitXParams ip; memset(&ip, 0, sizeof ip);
        ip.G = nullptr; ip.qc = kq; ip.X = Xh; ip.x_stride = v_el; ip.x_idx = 0; ip.a = arows; ip.ranks = nullptr; ip.s_fixed = kq; ip.idx = d_idx;
        init_x_kernel<<<dim3(grid1d((ll)arows * kq, 256, 512), G), 256, 0, st>>>(ip);
        QTT_CHECK(check_launch(h, "head init_x"));
        QTT_CHECK(apply_q(h, st, Q, G, arows, kq, Xh, v_el, 0, d_idx, kq, nullptr, 0, 0));
        transpose_kernel<<<dim3(grid1d((ll)arows * kq, 256, 512), G), 256, 0, st>>>(Xh, v_el, t.head[k], t.head_stride[k], arows, kq);
        QTT_CHECK(check_launch(h, "head transpose"));
    }
    return QTT_OK;
}

#define BIG_SVD_MIN_VEC 160     // above this many vectors the SVD of a carry core runs as a multi-CTA Jacobi

// SVD of the carry unfolding for large short sides (operator rounding): multi-CTA one-sided Jacobi on the rows of the
// unfolding (wide case) or, for a tall unfolding, on R^T after a Householder QR (so the rotated vectors stay short and
// contiguous); U = Q [U_r; 0] is recovered through the stored reflectors.  Requires the same left rank for every instance
// of the group (true for a single operator).  Synchronises once per Jacobi sweep.
static int big_svd(qtt_handle_s* h, cudaStream_t st, WsScope& ws, const SweepArgs& A, const SweepPlan& P, const SweepBufs& B, int G,
                   const int* d_idx, const int* h_idx, int k, double* cur, ll cur_stride, int cap) {
    qtt_batch* y = A.y;
    const int L = A.L, rs = L + 1, qc = P.q[k + 1], n = A.n[k];
    std::vector<int> hr;
    QTT_CHECK(read_ranks_host(h, st, y, hr));
    const int s_left = hr[(size_t)h_idx[0] * rs + k];
    bool uniform = true;
    for (int z = 1; z < G; ++z) if (hr[(size_t)h_idx[z] * rs + k] != s_left) uniform = false;
    if (!uniform) {
        // data-dependent left ranks (an eps-truncated bond) differ inside the group: the multi-CTA Jacobi works on one shape per
        // launch, so such a group is processed instance by instance (slow but correct; uniform groups - a single operator, capped
        // bonds - take the batched path below)
        for (int z = 0; z < G; ++z) {
            SweepBufs B1 = B;
            B1.JT = B.JT + (ll)z * B.sJT; B1.Gc = B.Gc + (ll)z * B.sG;
            QTT_CHECK(big_svd(h, st, ws, A, P, B1, 1, d_idx + z, h_idx + z, k, cur + (ll)z * cur_stride, cur_stride, cap));
        }
        return QTT_OK;
    }
    const int mr = s_left * n;
    const bool tall = mr > qc;
    const int nvec = std::min(mr, qc), np = (nvec + 1) & ~1;
    double *nrm, *X = cur; ll x_stride = cur_stride; int* order; int* rot;
    QTT_CHECK(ws.alloc_t(&nrm, (size_t)nvec * G)); QTT_CHECK(ws.alloc_t(&order, (size_t)nvec * G)); QTT_CHECK(ws.alloc_t(&rot, (size_t)G));
    QrBufs Q; memset(&Q, 0, sizeof Q);
    double *Rt = nullptr, *X0 = nullptr;
    if (tall) {
        const int panels = (int)ceil_div(qc, QR_NB);
        Q.sM = (ll)mr * qc; Q.sV = (ll)mr * qc; Q.stau = qc; Q.sTb = (ll)panels * QR_NB * QR_NB; Q.sTs = (ll)qc * 64; Q.sGram = (ll)GRAM_SPLITS * QR_NB * QR_NB;
        QTT_CHECK(ws.alloc_t(&Q.M, (size_t)Q.sM * G)); QTT_CHECK(ws.alloc_t(&Q.V, (size_t)Q.sV * G)); QTT_CHECK(ws.alloc_t(&Q.tau, (size_t)Q.stau * G));
        QTT_CHECK(ws.alloc_t(&Q.Tb, (size_t)Q.sTb * G)); QTT_CHECK(ws.alloc_t(&Q.Ts, (size_t)Q.sTs * G)); QTT_CHECK(ws.alloc_t(&Q.Gram, (size_t)Q.sGram * G));
        QTT_CHECK(ws.alloc_t(&Rt, (size_t)qc * qc * G));
        QTT_CHECK(ws.alloc_t(&X0, (size_t)mr * std::min(nvec, std::max(cap, 1)) * G));
        QTT_REQUIRE(leaf_supports(h, mr), QTT_ERR_UNSUPPORTED, "tall SVD: panel too tall for the QR leaf");
        // column-major copy of the unfolding, QR, R^T as the rows to rotate
        transpose_kernel<<<dim3(grid1d((ll)mr * qc, 256, 1024), G), 256, 0, st>>>(cur, cur_stride, Q.M, Q.sM, qc, mr);
        QTT_CHECK(check_launch(h, "big_svd transpose"));
        QTT_CHECK(qr_factor(h, st, Q, G, mr, qc, true));
        extract_r_kernel<<<dim3(grid1d((ll)qc * qc, 256, 512), G), 256, 0, st>>>(Q.M, Q.sM, mr, Rt, (ll)qc * qc, qc, qc);
        QTT_CHECK(check_launch(h, "big_svd extract_r"));
        X = Rt; x_stride = (ll)qc * qc;
    }
    big_jacobi_init_kernel<<<dim3(grid1d((ll)nvec * nvec, 256, 512), G), 256, 0, st>>>(B.JT, B.sJT, nvec);
    QTT_CHECK(check_launch(h, "big_jacobi_init"));
    std::vector<int> h_rot(G);
    bool converged = (nvec <= 1);
    // a whole sweep per cooperative launch (grid-wide barrier between the rounds) when the grid fits on the GPU
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, big_jacobi_sweep_kernel, 256, 0);
    const dim3 jgrid((unsigned)ceil_div(np / 2, 8), G);
    const bool coop = !QTT_ENV_FLAG("QTT_NO_COOP") && (ll)jgrid.x * jgrid.y <= (ll)per_sm * h->sm_count;
    for (int sweep = 0; sweep < 40 && !converged; ++sweep) {
        QTT_CUDA(cudaMemsetAsync(rot, 0, sizeof(int) * G, st));
        BigJacobiParams jp; memset(&jp, 0, sizeof jp);
        jp.X = X; jp.x_stride = x_stride; jp.ld = qc; jp.len = qc; jp.JT = B.JT; jp.jt_stride = B.sJT;
        jp.nvec = nvec; jp.np = np; jp.rot_count = rot; jp.tol = 1e-15;
        if (coop) {
            void* args[] = {&jp};
            QTT_CUDA(cudaLaunchCooperativeKernel((void*)big_jacobi_sweep_kernel, jgrid, dim3(256), args, 0, st));
            QTT_CHECK(check_launch(h, "jacobi sweep"));
        } else {
            for (int rho = 0; rho < np - 1; ++rho) {
                jp.rho = rho;
                big_jacobi_round_kernel<<<jgrid, 256, 0, st>>>(jp);
                QTT_CHECK(check_launch(h, "jacobi round"));
            }
        }
        QTT_CUDA(cudaMemcpyAsync(h_rot.data(), rot, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
        QTT_CUDA(cudaStreamSynchronize(st));
        converged = true;
        for (int v : h_rot) if (v) converged = false;
    }
    big_jacobi_norms_kernel<<<dim3((unsigned)ceil_div(nvec, 8), G), 256, 0, st>>>(X, x_stride, qc, qc, nvec, nrm, nvec);
    QTT_CHECK(check_launch(h, "jacobi norms"));
    BigRankParams rp; memset(&rp, 0, sizeof rp);
    rp.nrm = nrm; rp.n_stride = nvec; rp.nvec = nvec; rp.order = order; rp.o_stride = nvec; rp.ranks = y->ranks; rp.rstride = rs; rp.kcol = k;
    rp.idx = d_idx; rp.mr = mr; rp.qc = qc; rp.cap = cap; rp.eps2 = A.eps * A.eps; rp.status = A.status; rp.noconv = converged ? 0 : 1;
    big_jacobi_rank_kernel<<<G, 256, 0, st>>>(rp);
    QTT_CHECK(check_launch(h, "jacobi rank"));
    const int s_bound = std::min(nvec, std::max(cap, 0));
    BigOutParams op; memset(&op, 0, sizeof op);
    op.X = X; op.x_stride = x_stride; op.ld = qc; op.len = qc; op.JT = B.JT; op.jt_stride = B.sJT; op.nvec = nvec;
    op.nrm = nrm; op.n_stride = nvec; op.order = order; op.o_stride = nvec; op.ranks = y->ranks; op.rstride = rs; op.kcol = k; op.idx = d_idx;
    op.G = B.Gc; op.g_stride = B.sG; op.qc = qc; op.mr = mr; op.tall = tall ? 1 : 0; op.a_rows = mr;
    if (!tall) { op.U = y->cores[k]; op.u_stride = y->slot[k]; op.u_idx = 1; }
    else { op.U = X0; op.u_stride = (ll)mr * s_bound; op.u_idx = 0; }
    big_jacobi_out_kernel<<<dim3(grid1d((ll)std::max(mr, qc) * std::max(s_bound, 1), 256, 1024), G), 256, 0, st>>>(op);
    QTT_CHECK(check_launch(h, "jacobi out"));
    if (tall) {
        QTT_CHECK(apply_q(h, st, Q, G, mr, qc, X0, (ll)mr * s_bound, 0, d_idx, s_bound, y->ranks, rs, k + 1));
        transpose_dyn_kernel<<<dim3(grid1d((ll)mr * s_bound, 256, 1024), G), 256, 0, st>>>(X0, (ll)mr * s_bound, y->cores[k], y->slot[k], mr,
                                                                                            y->ranks, rs, k + 1, d_idx);
        QTT_CHECK(check_launch(h, "big_svd transpose out"));
    }
    return QTT_OK;
}

static int sweep_group_body(qtt_handle_s* h, cudaStream_t st, SweepArgs& A, int G, const int* d_idx, const int* h_idx);

// Everything the launch sequence of a sweep depends on: buffers, shapes, ranks, group membership.  Two calls with equal keys
// enqueue identical kernels with identical arguments (given the same workspace blocks, which the replay checks separately).
static void graph_key(const SweepArgs& A, int G, const int* d_idx, const int* h_idx, std::vector<ll>& key) {
    auto put_batch = [&](const qtt_batch* b) {
        if (!b) { key.push_back(-1); return; }
        key.push_back(b->L); key.push_back(b->N); key.push_back((ll)(size_t)b->ranks);
        for (int k = 0; k < b->L; ++k) { key.push_back((ll)(size_t)b->cores[k]); key.push_back(b->slot[k]); key.push_back(b->n[k]); }
    };
    key.clear();
    key.push_back((ll)A.mode); key.push_back(A.L); key.push_back(G); key.push_back((ll)(size_t)d_idx);
    key.push_back((ll)A.terms.size());
    for (const Term& t : A.terms) {
        key.push_back((ll)(size_t)t.op);
        if (t.op) for (int k = 0; k < A.L; ++k) { key.push_back((ll)(size_t)t.op->cores[k]); key.push_back(t.op->ranks[k]); key.push_back(t.op->n_out[k] * 65536 + t.op->n_in[k]); }
        put_batch(t.x);
        key.push_back((ll)(size_t)t.coef);
        ll csb; memcpy(&csb, &t.cs, sizeof csb); key.push_back(csb);
        for (int v : t.xr) key.push_back(v);
    }
    put_batch(A.y);
    for (int k = 0; k + 1 < A.L; ++k) key.push_back(A.caps ? A.caps[k] : -1);
    ll eb; memcpy(&eb, &A.eps, sizeof eb); key.push_back(eb);
    key.push_back((ll)(size_t)A.norm2); key.push_back((ll)(size_t)A.status);
    for (int i = 0; i < G; ++i) key.push_back(h_idx[i]);
}

#define GRAPH_CACHE_MAX 192

static void graph_cache_clear(qtt_handle_s* h) {
    for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
}

// One group of instances with a common rank signature: replay the captured launch sequence if this very sweep has been seen,
// else run it (capturing it on the way when it qualifies).
static int sweep_group(qtt_handle_s* h, cudaStream_t st, SweepArgs& A, int G, const int* d_idx, const int* h_idx) {
    bool ok = h->graphs_on && h->st_cap && !h->tl_on && !h->prof_on && !QTT_ENV_FLAG("QTT_COOP_LEAF");
    if (ok) {
        // sweeps that synchronise with the host inside (the multi-CTA Jacobi SVD polls its convergence flag) cannot be captured
        SweepPlan P;
        QTT_CHECK(make_plan(h, A, P));
        if (A.mode == MODE_ROUND)
            for (int k = 0; k < A.L - 1; ++k) if (std::min(P.sb[k] * A.n[k], P.q[k + 1]) > BIG_SVD_MIN_VEC) ok = false;
        // launch-bound sweeps only: with more than ~64 MB of workspace per sweep the kernels are long and a graph buys nothing
        if (P.per_inst_bytes * (ll)G > g_graph_max_bytes) ok = false;
    }
    if (!ok) return sweep_group_body(h, st, A, G, d_idx, h_idx);
    std::vector<ll> key;
    graph_key(A, G, d_idx, h_idx, key);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        // first sight of this sweep: run it plainly and remember it - one-off calls never pay for a capture
        if (h->graphs.size() >= GRAPH_CACHE_MAX) graph_cache_clear(h);
        h->graphs.emplace(std::move(key), qtt_handle_s::GraphEntry());
        return sweep_group_body(h, st, A, G, d_idx, h_idx);
    }
    if (it->second.exec) {
        qtt_handle_s::GraphEntry& e = it->second;
        WsScope ws(h);
        if (e.ws_gen == h->ws_gen && ws.retake(e.ws_idx, e.ws_ptr)) {
            QTT_CUDA(cudaGraphLaunch(e.exec, st));
            h->gemm_flops += e.gemm_flops; h->leaf_flops += e.leaf_flops; h->launches += e.launches;
            h->graph_replays += 1; e.stamp = ++h->graph_stamp;
            return QTT_OK;
        }
        cudaGraphExecDestroy(e.exec);
        e.exec = nullptr;
    }
    h->graphs.erase(it);
    // capture on the handle's own stream (the caller's may be the legacy default stream), then launch on the caller's
    const double f0 = h->gemm_flops, l0 = h->leaf_flops; const ll n0 = h->launches;
    if (cudaStreamBeginCapture(h->st_cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        return sweep_group_body(h, st, A, G, d_idx, h_idx);
    }
    h->ws_tracing = true; h->ws_trace.clear();
    cudaStream_t saved_tl = g_tl_stream; g_tl_stream = h->st_cap;
    const int rc = sweep_group_body(h, h->st_cap, A, G, d_idx, h_idx);
    g_tl_stream = saved_tl;
    h->ws_tracing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->st_cap, &graph);
    if (rc != QTT_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    if (ce != cudaSuccess || !graph) {
        cudaGetLastError();
        h->gemm_flops = f0; h->leaf_flops = l0; h->launches = n0;
        h->graphs_on = false;                                   // capture is not usable here: run uncaptured from now on
        return sweep_group_body(h, st, A, G, d_idx, h_idx);
    }
    qtt_handle_s::GraphEntry e;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        cudaGetLastError();
        h->gemm_flops = f0; h->leaf_flops = l0; h->launches = n0;
        h->graphs_on = false;
        return sweep_group_body(h, st, A, G, d_idx, h_idx);
    }
    e.ws_idx = h->ws_trace;
    for (int i : e.ws_idx) e.ws_ptr.push_back(h->blocks[i].ptr);
    e.gemm_flops = h->gemm_flops - f0; e.leaf_flops = h->leaf_flops - l0; e.launches = h->launches - n0;
    e.ws_gen = h->ws_gen; e.stamp = ++h->graph_stamp;
    h->graph_captures += 1;
    QTT_CUDA(cudaGraphLaunch(e.exec, st));
    h->graphs.emplace(std::move(key), std::move(e));
    return QTT_OK;
}

static int sweep_group_body(qtt_handle_s* h, cudaStream_t st, SweepArgs& A, int G, const int* d_idx, const int* h_idx) {
    const int L = A.L;
    SweepPlan P;
    QTT_CHECK(make_plan(h, A, P));
    NvtxRange nvtx_group("qtt:sweep_group");
    if (A.mode != MODE_NORM) {
        qtt_batch* y = A.y;
        for (int k = 0; k < L; ++k) {
            ll need = (A.mode == MODE_ROUND) ? (ll)P.sb[k] * A.n[k] * P.sb[k + 1] : (ll)P.q[k] * A.n[k] * P.q[k + 1];
            if (y->slot[k] < need) {
                char buf[256]; snprintf(buf, sizeof buf, "output slot %d too small: %lld < %lld", k, y->slot[k], need);
                h->err = buf; return QTT_ERR_SLOT;
            }
        }
    }
    WsScope ws(h);
    SweepBufs B; memset(&B, 0, sizeof B);
    B.sV = P.v_elems; B.sTb = (ll)std::max(P.tb_panels, 1) * QR_NB * QR_NB; B.sTs = (ll)std::max(P.ts_blocks, 1) * 64;
    B.stau = std::max(P.tau_len, 1); B.sM = P.m_elems; B.sR = std::max<ll>(P.r_elems, 1); B.sT = std::max<ll>(P.t_elems, 1);
    B.sX = std::max<ll>(P.x_elems, 1); B.sG = std::max<ll>(P.g_elems, 1); B.sJT = std::max<ll>((ll)P.nvec_max * P.nvec_max, 1);
    B.sGram = (ll)GRAM_SPLITS * QR_NB * QR_NB;
    QTT_CHECK(ws.alloc_t(&B.V, (size_t)std::max<ll>(B.sV, 1) * G));
    QTT_CHECK(ws.alloc_t(&B.Tb, (size_t)B.sTb * G));
    QTT_CHECK(ws.alloc_t(&B.Ts, (size_t)B.sTs * G));
    QTT_CHECK(ws.alloc_t(&B.tau, (size_t)B.stau * G));
    QTT_CHECK(ws.alloc_t(&B.M, (size_t)B.sM * G));
    QTT_CHECK(ws.alloc_t(&B.R, (size_t)B.sR * G));
    QTT_CHECK(ws.alloc_t(&B.T, (size_t)B.sT * G));
    QTT_CHECK(ws.alloc_t(&B.Gram, (size_t)B.sGram * G));
    if (A.mode == MODE_ROUND) {
        QTT_CHECK(ws.alloc_t(&B.X0, (size_t)B.sX * G));
        QTT_CHECK(ws.alloc_t(&B.X1, (size_t)B.sX * G));
        QTT_CHECK(ws.alloc_t(&B.Gc, (size_t)B.sG * G));
        QTT_CHECK(ws.alloc_t(&B.JT, (size_t)B.sJT * G));
    }

    // Head compression on the secondary stream: its result (head cores, dense core kc) is first needed at core kc_max of the sweep below.
    int kc_max = 0;
    for (auto& t : A.terms) kc_max = std::max(kc_max, t.kc);
    // (not while kernels are being timed: concurrent kernels would inflate the per-launch durations the roofline figures come from.
    //  Measured in round 2 with the overlap left on under qtt_set_profiling(1): bench value 16.44 -> 16.51 solves/s, but the summed
    //  event durations of the DMMA-class kernels then count the overlapped wall time twice - frac 0.71 -> 0.61, share 0.68 -> 0.81 -
    //  which says nothing about the kernels.  bench.py's e2e leg runs without per-launch events and has the overlap.)
    const bool fork = h->head_overlap && kc_max > 0 && kc_max < L - 1 && !h->tl_on && !h->prof_on && h->st2;
    cudaStream_t sth = fork ? h->st2 : st;
    if (fork) { QTT_CUDA(cudaEventRecord(h->ev_fork, st)); QTT_CUDA(cudaStreamWaitEvent(h->st2, h->ev_fork, 0)); }
    for (auto& t : A.terms) QTT_CHECK(compress_head(h, sth, ws, A, t, G, d_idx));
    if (fork) QTT_CUDA(cudaEventRecord(h->ev_join, h->st2));
    bool joined = !fork;

    // ---- right -> left: form M_k block by block, factorise ------------------------------------
    std::unique_ptr<NvtxRange> nvtx_qr(new NvtxRange("qtt:qr_sweep_right_to_left"));
    bool r_in_place = false;
    for (int k = L - 1; k >= 0; --k) {
        const int a = P.a[k], b = P.p[k];
        if (!joined && k <= kc_max) { QTT_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0)); joined = true; }
        for (size_t ti = 0; ti < A.terms.size(); ++ti)
            QTT_CHECK(form_block(h, st, A, P, B, G, d_idx, A.terms[ti], k, (k == 0) && ti > 0, r_in_place));
        if (k == 0) break;
        QTT_CHECK(qr_factor(h, st, qr_bufs_for_core(P, B, k), G, a, b, A.mode != MODE_NORM));
        // R_k feeds core k - 1.  A single operator term outside its compressed head consumes it through push_t only, which can read
        // the upper triangle in place from the factorised working matrix (push_t finishes before the GEMM of core k - 1 overwrites
        // it): no extracted copy (3.3 MB per instance and full-size core).  Everything else (plain terms, sums, head cores: the GEMM
        // reads R while it writes the working matrix) takes the copy.
        r_in_place = g_r_in_place && A.terms.size() == 1 && A.terms[0].op && (k - 1) > A.terms[0].kc && (k - 1) < L - 1;
        if (r_in_place) continue;
        dim3 grid(grid1d((ll)P.q[k] * b, 256, 512), G);
        extract_r_kernel<<<grid, 256, 0, st>>>(B.M, B.sM, a, B.R, B.sR, P.q[k], b);
        QTT_CHECK(check_launch(h, "extract_r"));
    }
    nvtx_qr.reset();
    // B.M now holds cur_0 = (1, n_0, q_1) for every instance
    if (A.norm2) {
        sumsq_kernel<<<G, 256, 0, st>>>(B.M, B.sM, (ll)P.a[0], d_idx, A.norm2);
        QTT_CHECK(check_launch(h, "sumsq"));
    }
    if (A.mode == MODE_NORM) return QTT_OK;

    qtt_batch* y = A.y;
    const int rs = L + 1;
    if (A.mode == MODE_ORTH) {
        for (int k = 0; k <= L; ++k) {
            set_rank_col_kernel<<<grid1d(G, 128), 128, 0, st>>>(y->ranks, rs, k, P.q[k], d_idx, G);
            QTT_CHECK(check_launch(h, "set_rank"));
        }
        CopyParams cp; memset(&cp, 0, sizeof cp);
        cp.src = B.M; cp.src_stride = B.sM; cp.src_idx = 0; cp.dst = y->cores[0]; cp.dst_stride = y->slot[0]; cp.dst_idx = 1;
        cp.idx = d_idx; cp.count = P.a[0]; cp.coef_scale = 1.0;
        copy_kernel<<<dim3(grid1d(cp.count, 256, 256), G), 256, 0, st>>>(cp);
        QTT_CHECK(check_launch(h, "copy core0"));
        for (int k = 1; k < L; ++k) {
            InitXParams ip; memset(&ip, 0, sizeof ip);
            ip.G = nullptr; ip.qc = P.q[k]; ip.X = y->cores[k]; ip.x_stride = y->slot[k]; ip.x_idx = 1; ip.a = P.a[k];
            ip.ranks = nullptr; ip.s_fixed = P.q[k]; ip.idx = d_idx;
            init_x_kernel<<<dim3(grid1d((ll)P.a[k] * P.q[k], 256, 512), G), 256, 0, st>>>(ip);
            QTT_CHECK(check_launch(h, "init_x"));
            QTT_CHECK(apply_q(h, st, qr_bufs_for_core(P, B, k), G, P.a[k], P.q[k], y->cores[k], y->slot[k], 1, d_idx, P.q[k], nullptr, 0, 0));
        }
        return QTT_OK;
    }

    // ---- left -> right: truncated SVD of the carry, push through the stored reflectors -------
    NvtxRange nvtx_svd("qtt:svd_sweep_left_to_right");
    set_rank_col_kernel<<<grid1d(G, 128), 128, 0, st>>>(y->ranks, rs, 0, 1, d_idx, G);
    QTT_CHECK(check_launch(h, "set_rank"));
    set_rank_col_kernel<<<grid1d(G, 128), 128, 0, st>>>(y->ranks, rs, L, 1, d_idx, G);
    QTT_CHECK(check_launch(h, "set_rank"));
    double* cur = B.M; ll cur_stride = B.sM;
    double* Xn = B.X0;
    for (int k = 0; k < L - 1; ++k) {
        // The multi-CTA Jacobi is for SVDs with more than BIG_SVD_MIN_VEC vectors.  The plan only knows BOUNDS (without a cap they grow
        // like n^k); when the bound is large the actual left ranks are read (one synchronisation) and the single-CTA kernel - which is
        // ragged-aware - still takes the step if the largest unfolding of the group is small.
        int mrb = P.sb[k] * A.n[k];                             // bound on the rows of the unfolding
        bool use_big = std::min(mrb, P.q[k + 1]) > BIG_SVD_MIN_VEC;
        if (use_big) {
            std::vector<int> hr;
            QTT_CHECK(read_ranks_host(h, st, y, hr));
            int mr_act = 0;
            for (int z = 0; z < G; ++z) mr_act = std::max(mr_act, hr[(size_t)h_idx[z] * rs + k] * A.n[k]);
            if (std::min(mr_act, P.q[k + 1]) <= BIG_SVD_MIN_VEC) { use_big = false; mrb = mr_act; }
        }
        if (use_big) {
            QTT_CHECK(big_svd(h, st, ws, A, P, B, G, d_idx, h_idx, k, cur, cur_stride, A.caps ? A.caps[k] : 2147483647));
        } else {
        SvdParams sp; memset(&sp, 0, sizeof sp);
        sp.A = cur; sp.a_stride = cur_stride; sp.qc = P.q[k + 1]; sp.n_mode = A.n[k];
        sp.ranks = y->ranks; sp.rstride = rs; sp.kcol = k; sp.idx = d_idx;
        sp.JT = B.JT; sp.jt_stride = B.sJT; sp.U = y->cores[k]; sp.u_slot = y->slot[k];
        sp.G = B.Gc; sp.g_stride = B.sG; sp.cap = A.caps ? A.caps[k] : 2147483647; sp.eps2 = A.eps * A.eps;
        sp.status = A.status; sp.max_sweeps = 60;
        const size_t svd_smem = g_svd_smem ? jacobi_svd_smem(std::min(mrb, P.q[k + 1]), std::max(mrb, P.q[k + 1])) : 0;
        sp.smem_vecs = svd_smem ? 1 : 0;
        const int svd_threads = std::max(SVD_THREADS, std::min(SVD_THREADS_MAX, 32 * ((std::min(mrb, P.q[k + 1]) + 1) / 2)));
        jacobi_svd_kernel<<<G, svd_threads, svd_smem, st>>>(sp);
        QTT_CHECK(check_launch(h, "jacobi_svd"));
        }
        InitXParams ip; memset(&ip, 0, sizeof ip);
        ip.G = B.Gc; ip.g_stride = B.sG; ip.qc = P.q[k + 1]; ip.X = Xn; ip.x_stride = B.sX; ip.x_idx = 0; ip.a = P.a[k + 1];
        ip.ranks = y->ranks; ip.rstride = rs; ip.kcol = k + 1; ip.idx = d_idx;
        init_x_kernel<<<dim3(grid1d((ll)P.a[k + 1] * P.sb[k + 1], 256, 512), G), 256, 0, st>>>(ip);
        QTT_CHECK(check_launch(h, "init_x"));
        QTT_CHECK(apply_q(h, st, qr_bufs_for_core(P, B, k + 1), G, P.a[k + 1], P.q[k + 1], Xn, B.sX, 0, d_idx, P.sb[k + 1], y->ranks, rs, k + 1));
        cur = Xn; cur_stride = B.sX;
        Xn = (Xn == B.X0) ? B.X1 : B.X0;
    }
    CopyParams cp; memset(&cp, 0, sizeof cp);
    cp.src = cur; cp.src_stride = cur_stride; cp.src_idx = 0; cp.dst = y->cores[L - 1]; cp.dst_stride = y->slot[L - 1]; cp.dst_idx = 1;
    cp.idx = d_idx; cp.coef_scale = 1.0;
    if (L == 1) { cp.count = P.a[0]; }
    else { cp.dyn = y->ranks; cp.dyn_stride = rs; cp.dyn_off = L - 1; cp.dyn_mul = P.a[L - 1]; }
    copy_kernel<<<dim3(grid1d((ll)P.a[L - 1] * P.sb[L - 1], 256, 256), G), 256, 0, st>>>(cp);
    QTT_CHECK(check_launch(h, "copy last"));
    return QTT_OK;
}

// Validate the terms, group instances by rank signature, run the sweeps chunk by chunk.
static int sweep_driver(qtt_handle_s* h, cudaStream_t st, int n_terms, const qtt_term* terms, qtt_batch* y, const int* caps,
                        double eps, double* norm2, int* status, SweepMode mode) {
    QTT_REQUIRE(n_terms >= 1 && terms, QTT_ERR_INVALID, "no terms");
    NvtxRange nvtx_sweep(mode == MODE_ROUND ? "qtt:round_sum" : (mode == MODE_ORTH ? "qtt:orthogonalize" : "qtt:norm2"));
    g_tl_stream = st;
    const qtt_batch* x0 = terms[0].x;
    QTT_REQUIRE(x0, QTT_ERR_INVALID, "null term");
    const int L = x0->L, N = x0->N;
    QTT_REQUIRE(L >= 2 && L <= QTT_MAX_L, QTT_ERR_UNSUPPORTED, "need 2 <= L <= 64 cores");
    QTT_REQUIRE(N >= 1 && N <= 65535, QTT_ERR_INVALID, "N out of range");
    if (mode != MODE_NORM) QTT_REQUIRE(y && y->L == L && y->N == N, QTT_ERR_INVALID, "output batch mismatch");
    QTT_CUDA(cudaSetDevice(h->device));
    std::vector<int> nmode(L);
    std::vector<std::vector<int>> hr(n_terms);
    for (int t = 0; t < n_terms; ++t) {
        const qtt_batch* x = terms[t].x;
        QTT_REQUIRE(x && x->L == L && x->N == N, QTT_ERR_INVALID, "term shape mismatch");
        if (terms[t].op) QTT_REQUIRE(terms[t].op->L == L, QTT_ERR_INVALID, "operator length mismatch");
        for (int k = 0; k < L; ++k) {
            int nout = terms[t].op ? terms[t].op->n_out[k] : x->n[k];
            if (terms[t].op) QTT_REQUIRE(terms[t].op->n_in[k] == x->n[k], QTT_ERR_INVALID, "operator n_in != vector mode size");
            if (t == 0) nmode[k] = nout; else QTT_REQUIRE(nmode[k] == nout, QTT_ERR_INVALID, "terms have different mode sizes");
        }
        QTT_CHECK(read_ranks_host(h, st, x, hr[t]));
    }
    if (mode != MODE_NORM) for (int k = 0; k < L; ++k) QTT_REQUIRE(y->n[k] == nmode[k], QTT_ERR_INVALID, "output mode size mismatch");
    if (status) QTT_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * N, st));

    // group by the concatenated rank signature
    std::map<std::vector<int>, std::vector<int>> groups;
    for (int i = 0; i < N; ++i) {
        std::vector<int> sig;
        sig.reserve((size_t)n_terms * (L + 1));
        for (int t = 0; t < n_terms; ++t) sig.insert(sig.end(), hr[t].begin() + (size_t)i * (L + 1), hr[t].begin() + (size_t)(i + 1) * (L + 1));
        groups[sig].push_back(i);
    }
    WsScope ws(h);
    int* d_idx_all = nullptr;
    QTT_CHECK(ws.alloc_t(&d_idx_all, N));
    std::vector<int> flat_host; flat_host.reserve(N);
    for (auto& g : groups) flat_host.insert(flat_host.end(), g.second.begin(), g.second.end());
    QTT_CUDA(cudaMemcpyAsync(d_idx_all, flat_host.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    QTT_CUDA(cudaStreamSynchronize(st));
    int pos = 0;
    for (auto& g : groups) {
        const std::vector<int>& sig = g.first;
        const int Gn = (int)g.second.size();
        SweepArgs A;
        A.L = L; A.n = nmode.data(); A.y = y; A.caps = caps; A.eps = eps; A.norm2 = norm2; A.status = status; A.mode = mode;
        A.terms.resize(n_terms);
        for (int t = 0; t < n_terms; ++t) {
            Term& T = A.terms[t];
            T.op = terms[t].op; T.x = terms[t].x; T.coef = terms[t].coef; T.cs = terms[t].coef_scale;
            T.xr.assign(sig.begin() + (size_t)t * (L + 1), sig.begin() + (size_t)(t + 1) * (L + 1));
            T.pr.resize(L + 1); T.off.assign(L + 1, 0);
            for (int k = 0; k <= L; ++k) T.pr[k] = T.xr[k] * (T.op ? T.op->ranks[k] : 1);
            QTT_REQUIRE(T.pr[0] == 1 && T.pr[L] == 1, QTT_ERR_INVALID, "boundary ranks must be 1");
            T.prod = T.pr; T.kc = 0;
            if (T.op && L >= 3 && !QTT_ENV_FLAG("QTT_NO_HEAD")) {
                std::vector<int> c(L + 1, 1);
                int kc = 0;
                for (int k = 0; k + 1 <= L - 2; ++k) {
                    c[k + 1] = (int)std::min<ll>((ll)c[k] * nmode[k], T.prod[k + 1]);
                    if (c[k + 1] < T.prod[k + 1]) kc = k + 1;
                }
                // worth it only when the compressed bond is at most half of the product bond somewhere
                bool useful = false;
                for (int k = 1; k <= kc; ++k) if (2 * c[k] <= T.prod[k]) useful = true;
                if (kc >= 1 && useful) {
                    T.kc = kc; T.c.assign(c.begin(), c.begin() + kc + 1);
                    for (int k = 1; k <= kc; ++k) T.pr[k] = T.c[k];
                }
            }
        }
        for (int k = 1; k < L; ++k) { int o = 0; for (int t = 0; t < n_terms; ++t) { A.terms[t].off[k] = o; o += A.terms[t].pr[k]; } }
        SweepPlan P;
        QTT_CHECK(make_plan(h, A, P));
        const ll budget = (ll)(h->ws_limit * 0.7);
        int chunk = (int)std::max<ll>(1, std::min<ll>(Gn, budget / std::max<ll>(P.per_inst_bytes, 1)));
        for (int c0 = 0; c0 < Gn; c0 += chunk) {
            const int Gc = std::min(chunk, Gn - c0);
            QTT_CHECK(sweep_group(h, st, A, Gc, d_idx_all + pos + c0, flat_host.data() + pos + c0));
        }
        pos += Gn;
    }
    return QTT_OK;
}

extern "C" int qtt_round_sum_batched(qtt_handle_t h, int n_terms, const qtt_term* terms, qtt_batch* y,
                                     const int* caps, double rel_eps, int* status, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    return sweep_driver(h, (cudaStream_t)stream, n_terms, terms, y, caps, rel_eps, nullptr, status, MODE_ROUND);
}

extern "C" int qtt_orthogonalize_batched(qtt_handle_t h, int n_terms, const qtt_term* terms, qtt_batch* y,
                                         double* norm2, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    return sweep_driver(h, (cudaStream_t)stream, n_terms, terms, y, nullptr, 0.0, norm2, nullptr, y ? MODE_ORTH : MODE_NORM);
}

// ------------------------------------------------------------------------------------------------
// truncated steepest descent, device resident                                  solver.py:251-275
// ------------------------------------------------------------------------------------------------
struct TmpBatch {
    qtt_batch b;
    std::vector<double*> cores;
    std::vector<ll> slot;
};

static int tmp_batch_alloc(WsScope& ws, TmpBatch& t, int L, int N, const int* n, const std::vector<int>& rb) {
    t.cores.resize(L); t.slot.resize(L);
    for (int k = 0; k < L; ++k) {
        t.slot[k] = (ll)rb[k] * n[k] * rb[k + 1];
        int rc = ws.alloc_t(&t.cores[k], (size_t)t.slot[k] * N);
        if (rc != QTT_OK) return rc;
    }
    int* ranks = nullptr;
    int rc = ws.alloc_t(&ranks, (size_t)N * (L + 1));
    if (rc != QTT_OK) return rc;
    t.b.L = L; t.b.N = N; t.b.n = n; t.b.ranks = ranks; t.b.cores = t.cores.data(); t.b.slot = t.slot.data();
    return QTT_OK;
}

__global__ void gd_gamma_kernel(const double* r2, const double* rAr, double lr, double* gamma, double* neg_gamma,
                                double* hist, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double g = (rAr[i] > 0.0) ? lr * r2[i] / rAr[i] : 0.0;    // zero right-hand side / converged instance: no step (not 0/0)
    gamma[i] = g; neg_gamma[i] = -g;
    if (hist) { hist[2 * i] = r2[i]; hist[2 * i + 1] = g; }
}

// done[i] becomes 1 when r2/x2 <= acc (first time: snapshot needed), steps[i] records the step
__global__ void gd_stop_kernel(const double* r2, const double* x2, double acc, int step, int* done, int* newly, int* steps, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    newly[i] = 0;
    if (!done[i] && (r2[i] / x2[i] <= acc)) { done[i] = 1; newly[i] = 1; steps[i] = step; }
}

// Conjugate gradients only (no reference counterpart): besides the tolerance test, an instance also stops when its relative
// residual has not halved over three consecutive checks (30 steps) - the tolerance a caller passes (e.g. AMEn's eps = 1e-15,
// squared) may lie below what FP64 with rank truncation can reach, and iterating on costs a full apply+round sweep per step.
// steps[i] records the step; stalled[i] = 1 tells the caller that the tolerance was NOT met.
__global__ void cg_stop_kernel(const double* r2, const double* x2, double acc, int step, int* done, int* newly, int* steps,
                               double* best, int* stall, int* stalled, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    newly[i] = 0;
    if (done[i]) return;
    const double ratio = r2[i] / x2[i];
    if (ratio <= acc) { done[i] = 1; newly[i] = 1; steps[i] = step; return; }
    if (ratio < 0.5 * best[i]) { best[i] = ratio; stall[i] = 0; return; }
    if (++stall[i] >= 3) { done[i] = 1; newly[i] = 1; steps[i] = step; if (stalled) stalled[i] = 1; }
}

__global__ void masked_copy_kernel(const double* src, double* dst, ll slot_src, ll slot_dst, const int* mask, int invert) {
    const int i = blockIdx.y;
    const bool on = mask ? (invert ? !mask[i] : mask[i] != 0) : true;
    if (!on) return;
    const ll cnt = slot_src < slot_dst ? slot_src : slot_dst;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (ll)gridDim.x * blockDim.x) dst[(ll)i * slot_dst + e] = src[(ll)i * slot_src + e];
}
__global__ void masked_copy_ranks_kernel(const int* src, int* dst, int rs, const int* mask, int invert, int N) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * rs) return;
    const int i = e / rs;
    const bool on = mask ? (invert ? !mask[i] : mask[i] != 0) : true;
    if (on) dst[e] = src[e];
}

static int batch_copy_masked(qtt_handle_s* h, cudaStream_t st, const qtt_batch* src, qtt_batch* dst, const int* mask, int invert) {
    const int L = src->L, N = src->N;
    for (int k = 0; k < L; ++k) {
        masked_copy_kernel<<<dim3(grid1d(std::min(src->slot[k], dst->slot[k]), 256, 64), N), 256, 0, st>>>(
            src->cores[k], dst->cores[k], src->slot[k], dst->slot[k], mask, invert);
        QTT_CHECK(check_launch(h, "masked_copy"));
    }
    masked_copy_ranks_kernel<<<grid1d((ll)N * (L + 1), 256), 256, 0, st>>>(src->ranks, dst->ranks, L + 1, mask, invert, N);
    return check_launch(h, "masked_copy_ranks");
}

extern "C" int qtt_gd_solve_batched(qtt_handle_t h, const qtt_mpo* A, const qtt_batch* b, qtt_batch* x, double* r2_out,
                                    int n_steps, double lr, int max_rank, int recalc_every, double rel_accuracy,
                                    double* hist, int* steps_done, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    NvtxRange nvtx_solve("qtt:gd_solve");
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(A && b && x && r2_out, QTT_ERR_INVALID, "null argument");
    const int L = b->L, N = b->N;
    QTT_REQUIRE(A->L == L && x->L == L && x->N == N && L >= 2 && L <= QTT_MAX_L, QTT_ERR_INVALID, "shape mismatch");
    QTT_REQUIRE(max_rank >= 1 && recalc_every >= 1 && n_steps >= 0, QTT_ERR_INVALID, "bad solver parameter");
    for (int k = 0; k < L; ++k) QTT_REQUIRE(A->n_out[k] == b->n[k] && A->n_in[k] == b->n[k] && x->n[k] == b->n[k], QTT_ERR_INVALID, "mode sizes");
    QTT_CUDA(cudaSetDevice(h->device));
    std::vector<int> caps(L - 1, max_rank), rb(L + 1, 1);
    for (int k = 1; k < L; ++k) rb[k] = max_rank;
    {   // structural bounds tighten the slots near the ends
        ll left = 1; for (int k = 1; k < L; ++k) { left = std::min<ll>(left * b->n[k - 1], max_rank); rb[k] = (int)std::min<ll>(rb[k], left); }
        ll right = 1; for (int k = L - 1; k >= 1; --k) { right = std::min<ll>(right * b->n[k], max_rank); rb[k] = (int)std::min<ll>(rb[k], right); }
    }
    for (int k = 0; k < L; ++k) QTT_REQUIRE(x->slot[k] >= (ll)rb[k] * b->n[k] * rb[k + 1], QTT_ERR_SLOT, "x slot too small for max_rank");
    WsScope ws(h);
    TmpBatch xa, xb, ra, rbt, Ar;
    QTT_CHECK(tmp_batch_alloc(ws, xa, L, N, b->n, rb)); QTT_CHECK(tmp_batch_alloc(ws, xb, L, N, b->n, rb));
    QTT_CHECK(tmp_batch_alloc(ws, ra, L, N, b->n, rb)); QTT_CHECK(tmp_batch_alloc(ws, rbt, L, N, b->n, rb));
    QTT_CHECK(tmp_batch_alloc(ws, Ar, L, N, b->n, rb));
    double *d_r2, *d_x2, *d_rAr, *d_g, *d_ng; int *d_done, *d_new;
    QTT_CHECK(ws.alloc_t(&d_r2, N)); QTT_CHECK(ws.alloc_t(&d_x2, N)); QTT_CHECK(ws.alloc_t(&d_rAr, N));
    QTT_CHECK(ws.alloc_t(&d_g, N)); QTT_CHECK(ws.alloc_t(&d_ng, N)); QTT_CHECK(ws.alloc_t(&d_done, N)); QTT_CHECK(ws.alloc_t(&d_new, N));
    QTT_CUDA(cudaMemsetAsync(d_done, 0, sizeof(int) * N, st));
    if (steps_done) { fill_int_kernel<<<grid1d(N, 256), 256, 0, st>>>(steps_done, N, n_steps); QTT_CHECK(check_launch(h, "fill")); }
    // x = 0 (rank-1 zero cores, tm.zeros)
    for (int k = 0; k < L; ++k) QTT_CUDA(cudaMemsetAsync(xa.cores[k], 0, sizeof(double) * xa.slot[k] * N, st));
    fill_int_kernel<<<grid1d((ll)N * (L + 1), 256), 256, 0, st>>>(xa.b.ranks, (ll)N * (L + 1), 1);
    QTT_CHECK(check_launch(h, "fill"));
    TmpBatch *xc = &xa, *xn = &xb, *rc = &ra, *rn = &rbt;
    bool any_done = false;
    std::vector<int> h_done(N, 0);
    auto residual = [&](TmpBatch* xcur, TmpBatch* rdst) -> int {
        qtt_term t[2];
        t[0].op = nullptr; t[0].x = b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
        t[1].op = A; t[1].x = &xcur->b; t[1].coef = nullptr; t[1].coef_scale = -1.0;
        return sweep_driver(h, st, 2, t, &rdst->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND);
    };
    int n = 0;
    for (; n < n_steps; ++n) {
        if (n % recalc_every == 0) QTT_CHECK(residual(xc, rc));
        {   // Ar = round(A @ r)
            qtt_term t; t.op = A; t.x = &rc->b; t.coef = nullptr; t.coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 1, &t, &Ar.b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
        }
        {   // r2 = ||r||^2 through the orthogonalisation sweep, as the reference does
            qtt_term t; t.op = nullptr; t.x = &rc->b; t.coef = nullptr; t.coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 1, &t, nullptr, nullptr, 0.0, d_r2, nullptr, MODE_NORM));
        }
        if (rel_accuracy > 0.0 && n > 10 && n % 10 == 0) {
            qtt_term t; t.op = nullptr; t.x = &xc->b; t.coef = nullptr; t.coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 1, &t, nullptr, nullptr, 0.0, d_x2, nullptr, MODE_NORM));
            gd_stop_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_r2, d_x2, rel_accuracy, n, d_done, d_new, steps_done ? steps_done : d_new, N);
            QTT_CHECK(check_launch(h, "gd_stop"));
            QTT_CHECK(batch_copy_masked(h, st, &xc->b, x, d_new, 0));      // snapshot the instances that just converged
            QTT_CUDA(cudaMemcpyAsync(h_done.data(), d_done, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
            QTT_CUDA(cudaStreamSynchronize(st));
            int cnt = 0; for (int v : h_done) cnt += v;
            any_done = cnt > 0;
            if (cnt == N) break;
        }
        QTT_CHECK(inner_plain(h, st, &rc->b, &Ar.b, d_rAr));
        gd_gamma_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_r2, d_rAr, lr, d_g, d_ng, hist ? hist + (ll)n * N * 2 : nullptr, N);
        QTT_CHECK(check_launch(h, "gd_gamma"));
        {   // x = round(x + gamma r);  r = round(r - gamma Ar)
            qtt_term t[2];
            t[0].op = nullptr; t[0].x = &xc->b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
            t[1].op = nullptr; t[1].x = &rc->b; t[1].coef = d_g; t[1].coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 2, t, &xn->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
            t[0].x = &rc->b;
            t[1].x = &Ar.b; t[1].coef = d_ng;
            QTT_CHECK(sweep_driver(h, st, 2, t, &rn->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
            std::swap(xc, xn); std::swap(rc, rn);
        }
    }
    // instances that never met the stopping test take the final iterate
    QTT_CHECK(batch_copy_masked(h, st, &xc->b, x, any_done ? d_done : nullptr, 1));
    {   // final residual of the returned x
        qtt_term t[2];
        t[0].op = nullptr; t[0].x = b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
        t[1].op = A; t[1].x = x; t[1].coef = nullptr; t[1].coef_scale = -1.0;
        QTT_CHECK(sweep_driver(h, st, 2, t, &rc->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
        qtt_term u; u.op = nullptr; u.x = &rc->b; u.coef = nullptr; u.coef_scale = 1.0;
        QTT_CHECK(sweep_driver(h, st, 1, &u, nullptr, nullptr, 0.0, r2_out, nullptr, MODE_NORM));
    }
    QTT_CUDA(cudaStreamSynchronize(st));
    return QTT_OK;
}

// ------------------------------------------------------------------------------------------------
// truncated conjugate gradients, device resident (the solver behind solve_with_amen).
// No reference counterpart (the reference calls ttpy's AMEn); restated on the CPU in oracle/tt_oracle.py
// (solve_with_cg) with the same conventions as the steepest-descent loop above.
// ------------------------------------------------------------------------------------------------
// alpha = <p, r> / <p, A p>  (exact line search along p; tolerant of the rounding perturbations of r and p)
__global__ void cg_alpha_kernel(const double* pr, const double* pAp, const double* rr, double* alpha, double* neg_alpha, double* hist, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double a = (pAp[i] > 0.0) ? pr[i] / pAp[i] : 0.0;
    alpha[i] = a; neg_alpha[i] = -a;
    if (hist) { hist[2 * i] = rr[i]; hist[2 * i + 1] = a; }
}
// beta = -<r_new, A p> / <p, A p>  (explicit A-orthogonalisation of the next direction)
__global__ void cg_beta_kernel(const double* rAp, const double* pAp, double* beta, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    beta[i] = (pAp[i] > 0.0) ? -rAp[i] / pAp[i] : 0.0;
}

extern "C" int qtt_cg_solve_batched(qtt_handle_t h, const qtt_mpo* A, const qtt_batch* b, qtt_batch* x, double* r2_out,
                                    int n_steps, int max_rank, int recalc_every, double rel_accuracy,
                                    double* hist, int* steps_done, void* stream) {
    if (!h) return QTT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    QTT_REQUIRE(A && b && x && r2_out, QTT_ERR_INVALID, "null argument");
    const int L = b->L, N = b->N;
    QTT_REQUIRE(A->L == L && x->L == L && x->N == N && L >= 2 && L <= QTT_MAX_L, QTT_ERR_INVALID, "shape mismatch");
    QTT_REQUIRE(max_rank >= 1 && recalc_every >= 1 && n_steps >= 0, QTT_ERR_INVALID, "bad solver parameter");
    for (int k = 0; k < L; ++k) QTT_REQUIRE(A->n_out[k] == b->n[k] && A->n_in[k] == b->n[k] && x->n[k] == b->n[k], QTT_ERR_INVALID, "mode sizes");
    QTT_CUDA(cudaSetDevice(h->device));
    std::vector<int> caps(L - 1, max_rank), rb(L + 1, 1);
    for (int k = 1; k < L; ++k) rb[k] = max_rank;
    {
        ll left = 1; for (int k = 1; k < L; ++k) { left = std::min<ll>(left * b->n[k - 1], max_rank); rb[k] = (int)std::min<ll>(rb[k], left); }
        ll right = 1; for (int k = L - 1; k >= 1; --k) { right = std::min<ll>(right * b->n[k], max_rank); rb[k] = (int)std::min<ll>(rb[k], right); }
    }
    for (int k = 0; k < L; ++k) QTT_REQUIRE(x->slot[k] >= (ll)rb[k] * b->n[k] * rb[k + 1], QTT_ERR_SLOT, "x slot too small for max_rank");
    WsScope ws(h);
    TmpBatch xa, xb, ra, rbt, pa, pb, Ap;
    QTT_CHECK(tmp_batch_alloc(ws, xa, L, N, b->n, rb)); QTT_CHECK(tmp_batch_alloc(ws, xb, L, N, b->n, rb));
    QTT_CHECK(tmp_batch_alloc(ws, ra, L, N, b->n, rb)); QTT_CHECK(tmp_batch_alloc(ws, rbt, L, N, b->n, rb));
    QTT_CHECK(tmp_batch_alloc(ws, pa, L, N, b->n, rb)); QTT_CHECK(tmp_batch_alloc(ws, pb, L, N, b->n, rb));
    QTT_CHECK(tmp_batch_alloc(ws, Ap, L, N, b->n, rb));
    double *d_rr, *d_x2, *d_pAp, *d_pr, *d_rAp, *d_a, *d_na, *d_beta; int *d_done, *d_new;
    QTT_CHECK(ws.alloc_t(&d_rr, N)); QTT_CHECK(ws.alloc_t(&d_x2, N)); QTT_CHECK(ws.alloc_t(&d_pAp, N));
    QTT_CHECK(ws.alloc_t(&d_pr, N)); QTT_CHECK(ws.alloc_t(&d_rAp, N));
    QTT_CHECK(ws.alloc_t(&d_a, N)); QTT_CHECK(ws.alloc_t(&d_na, N)); QTT_CHECK(ws.alloc_t(&d_beta, N));
    QTT_CHECK(ws.alloc_t(&d_done, N)); QTT_CHECK(ws.alloc_t(&d_new, N));
    double* d_best; int *d_stall, *d_steps_tmp;
    QTT_CHECK(ws.alloc_t(&d_best, N)); QTT_CHECK(ws.alloc_t(&d_stall, N)); QTT_CHECK(ws.alloc_t(&d_steps_tmp, N));
    QTT_CUDA(cudaMemsetAsync(d_done, 0, sizeof(int) * N, st));
    QTT_CUDA(cudaMemsetAsync(d_stall, 0, sizeof(int) * N, st));
    fill_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_best, N, 1e300);
    QTT_CHECK(check_launch(h, "fill"));
    if (steps_done) { fill_int_kernel<<<grid1d(N, 256), 256, 0, st>>>(steps_done, N, n_steps); QTT_CHECK(check_launch(h, "fill")); }
    for (int k = 0; k < L; ++k) QTT_CUDA(cudaMemsetAsync(xa.cores[k], 0, sizeof(double) * xa.slot[k] * N, st));
    fill_int_kernel<<<grid1d((ll)N * (L + 1), 256), 256, 0, st>>>(xa.b.ranks, (ll)N * (L + 1), 1);
    QTT_CHECK(check_launch(h, "fill"));
    TmpBatch *xc = &xa, *xn = &xb, *rc = &ra, *rn = &rbt, *pc = &pa, *pn = &pb;
    auto round1 = [&](const qtt_batch* src, TmpBatch* dst) -> int {
        qtt_term t; t.op = nullptr; t.x = src; t.coef = nullptr; t.coef_scale = 1.0;
        return sweep_driver(h, st, 1, &t, &dst->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND);
    };
    auto norm2 = [&](const qtt_batch* src, double* out) -> int {
        qtt_term t; t.op = nullptr; t.x = src; t.coef = nullptr; t.coef_scale = 1.0;
        return sweep_driver(h, st, 1, &t, nullptr, nullptr, 0.0, out, nullptr, MODE_NORM);
    };
    auto residual = [&](const qtt_batch* xcur, TmpBatch* rdst) -> int {
        qtt_term t[2];
        t[0].op = nullptr; t[0].x = b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
        t[1].op = A; t[1].x = xcur; t[1].coef = nullptr; t[1].coef_scale = -1.0;
        return sweep_driver(h, st, 2, t, &rdst->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND);
    };
    // r = round(b), p = r
    QTT_CHECK(round1(b, rc));
    QTT_CHECK(batch_copy_masked(h, st, &rc->b, &pc->b, nullptr, 0));
    bool any_done = false;
    std::vector<int> h_done(N, 0);
    for (int n = 0; n < n_steps; ++n) {
        {   // Ap = round(A @ p)
            qtt_term t; t.op = A; t.x = &pc->b; t.coef = nullptr; t.coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 1, &t, &Ap.b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
        }
        QTT_CHECK(inner_plain(h, st, &pc->b, &Ap.b, d_pAp));
        QTT_CHECK(inner_plain(h, st, &pc->b, &rc->b, d_pr));
        if (hist) QTT_CHECK(norm2(&rc->b, d_rr));
        cg_alpha_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_pr, d_pAp, d_rr, d_a, d_na, hist ? hist + (ll)n * N * 2 : nullptr, N);
        QTT_CHECK(check_launch(h, "cg_alpha"));
        {   // x = round(x + alpha p)
            qtt_term t[2];
            t[0].op = nullptr; t[0].x = &xc->b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
            t[1].op = nullptr; t[1].x = &pc->b; t[1].coef = d_a; t[1].coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 2, t, &xn->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
            std::swap(xc, xn);
        }
        if ((n + 1) % recalc_every == 0) {
            QTT_CHECK(residual(&xc->b, rn));
        } else {   // r = round(r - alpha Ap)
            qtt_term t[2];
            t[0].op = nullptr; t[0].x = &rc->b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
            t[1].op = nullptr; t[1].x = &Ap.b; t[1].coef = d_na; t[1].coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 2, t, &rn->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
        }
        std::swap(rc, rn);
        if (rel_accuracy > 0.0 && n + 1 > 10 && (n + 1) % 10 == 0) {
            QTT_CHECK(norm2(&rc->b, d_rr));
            QTT_CHECK(norm2(&xc->b, d_x2));
            cg_stop_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_rr, d_x2, rel_accuracy, n + 1, d_done, d_new, steps_done ? steps_done : d_steps_tmp,
                                                           d_best, d_stall, nullptr, N);
            QTT_CHECK(check_launch(h, "cg_stop"));
            QTT_CHECK(batch_copy_masked(h, st, &xc->b, x, d_new, 0));
            QTT_CUDA(cudaMemcpyAsync(h_done.data(), d_done, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
            QTT_CUDA(cudaStreamSynchronize(st));
            int cnt = 0; for (int v : h_done) cnt += v;
            any_done = cnt > 0;
            if (cnt == N) break;
        }
        QTT_CHECK(inner_plain(h, st, &rc->b, &Ap.b, d_rAp));
        cg_beta_kernel<<<grid1d(N, 256), 256, 0, st>>>(d_rAp, d_pAp, d_beta, N);
        QTT_CHECK(check_launch(h, "cg_beta"));
        {   // p = round(r + beta p)
            qtt_term t[2];
            t[0].op = nullptr; t[0].x = &rc->b; t[0].coef = nullptr; t[0].coef_scale = 1.0;
            t[1].op = nullptr; t[1].x = &pc->b; t[1].coef = d_beta; t[1].coef_scale = 1.0;
            QTT_CHECK(sweep_driver(h, st, 2, t, &pn->b, caps.data(), 1e-16, nullptr, nullptr, MODE_ROUND));
            std::swap(pc, pn);
        }
    }
    QTT_CHECK(batch_copy_masked(h, st, &xc->b, x, any_done ? d_done : nullptr, 1));
    {   // final residual of the returned x
        QTT_CHECK(residual(x, rc));
        QTT_CHECK(norm2(&rc->b, r2_out));
    }
    QTT_CUDA(cudaStreamSynchronize(st));
    return QTT_OK;
}
