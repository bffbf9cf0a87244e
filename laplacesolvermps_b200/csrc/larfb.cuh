// Fused application of a compact-WY block reflector to a block of columns (LAPACK dlarfb):
//
//     C <- (I - V op(T) V^T) C          V: rows x jb (jb <= 64), T: jb x jb upper triangular, C: rows x nc
//
// One CTA owns BN columns of one instance and makes two streaming passes over its rows in 64-row chunks:
// pass 1 accumulates W = V^T C in DMMA fragments, W2 = -op(T) W is formed in shared memory, pass 2
// updates C += V W2 in place.  This replaces three GEMM launches (and the W round trip through global
// memory) per QR panel; it is the trailing update of the blocked Householder QR (op(T) = T^T) and the
// application of the stored reflectors in the truncation sweep (op(T) = T).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

#define LARFB_NB 64
#define LARFB_THREADS 256
#define LARFB_LDS 68            // 64 + 4: conflict-free fragment reads along either index

__device__ int g_larfb_cp16 = 1;     // set from QTT_LARFB_CP16 at handle creation (A/B switch)

struct LarfbParams {
    const double* V; ll v_stride; int ldv;     // element (r, i) at V[i*ldv + r]
    const double* T; ll t_stride; int ldt;     // row-major jb x jb, upper triangular
    double* C; ll c_stride; int ldc; int c_idx; // element (r, c) at C[c*ldc + r]
    int rows, jb, nc, trans_t;
    const int* idx;
    const int* n_dyn; int n_dyn_stride, n_dyn_off;
};

template <int BN, int NSTAGE, int NTH = LARFB_THREADS>
__global__ void __launch_bounds__(NTH) larfb_kernel(const LarfbParams p) {
    extern __shared__ double sm[];
    constexpr int LDS = LARFB_LDS, NB = LARFB_NB, FN = BN / 8;
    constexpr int STAGE = (NB + BN) * LDS;  // one pipeline stage: a 64-row chunk of V and of C
    double* Vs = sm;                       // [NB][LDS]   Vs[i*LDS + r]   (also T staging: Ts[l*LDS + i])
    double* Cs = Vs + NB * LDS;            // [BN][LDS]   Cs[c*LDS + r]   (also W staging: Ws[c*LDS + i])
    double* W2s = sm + NSTAGE * STAGE;     // [BN][LDS]   W2s[c*LDS + i]  = -(op(T) W)[i][c]
    const int z = blockIdx.y, cb = blockIdx.x;
    const int inst = p.idx ? p.idx[z] : z;
    const int nc = p.n_dyn ? p.n_dyn[(ll)inst * p.n_dyn_stride + p.n_dyn_off] : p.nc;
    const int c0 = cb * BN;
    if (c0 >= nc) return;
    const int ncb = min(BN, nc - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const double* V = p.V + (ll)z * p.v_stride;
    const double* T = p.T + (ll)z * p.t_stride;
    double* C = p.C + (ll)(p.c_idx ? inst : z) * p.c_stride + (ll)c0 * p.ldc;
    const int rows = p.rows, jb = p.jb;
    const int kmax = (jb + 3) & ~3;

    // chunk loads go global -> shared with cp.async (8-byte, zero-filled out of range): no register staging,
    // all 32 loads of a thread are in flight at once (one CTA must keep ~50 KB in flight to cover L2 latency)
    // (Measured and dropped: staging the chunks with TMA bulk copies - cp.async.bulk, one 512-byte copy per column
    //  segment completing on an mbarrier, SASS UBLKCP - is slower here: 58 ms instead of 38 ms of larfb time per sweep.
    //  The copies are too small for the TMA unit; a tiled tensor map would need an unpadded, swizzled tile layout.)
    // Per-thread staging slots: element e = tid + 256 u has row r = tid & 63 (fixed) and column (tid >> 6) + 4 u, so
    // only base pointers are kept and every cp.async costs one 64-bit add (the index arithmetic used to be 3.5 integer
    // instructions per DMMA).
    constexpr int CPP = NTH / 64;                       // columns staged per pass of the CTA
    const int lr = tid & 63, lc0 = tid >> 6;
    const double* vsrc0 = V + (ll)lc0 * p.ldv + lr;
    const double* csrc0 = C + (ll)lc0 * p.ldc + lr;
    const unsigned vdst0 = (unsigned)__cvta_generic_to_shared(Vs + lc0 * LDS + lr);
    const unsigned cdst0 = (unsigned)__cvta_generic_to_shared(Cs + lc0 * LDS + lr);
    // 16-byte variant (half the copy instructions: the load-issue phases were 26 % of the kernel's stall samples):
    // a thread copies the row pair (lr2, lr2 + 1) of column (tid >> 5) + CPP2 u.  Needs 16-byte aligned column
    // starts (even leading dimensions, even row offset of the panel); otherwise the 8-byte path below is used.
    constexpr int CPP2 = NTH / 32;
    const int lr2 = (tid & 31) * 2, lc2 = tid >> 5;
    const bool al16 = g_larfb_cp16 && ((((unsigned long long)V | (unsigned long long)C) & 15ull) == 0) && !(p.ldv & 1) && !(p.ldc & 1);
    const double* vsrc2 = V + (ll)lc2 * p.ldv + lr2;
    const double* csrc2 = C + (ll)lc2 * p.ldc + lr2;
    const unsigned vdst2 = (unsigned)__cvta_generic_to_shared(Vs + lc2 * LDS + lr2);
    const unsigned cdst2 = (unsigned)__cvta_generic_to_shared(Cs + lc2 * LDS + lr2);
    auto issue_chunk = [&](int r0, int stage) {
        const unsigned soff = (unsigned)(stage * STAGE * sizeof(double));
        if (al16) {
            const int left = rows - (r0 + lr2);
            const int nb = left >= 2 ? 16 : (left == 1 ? 8 : 0);
#pragma unroll
            for (int u = 0; u < NB / CPP2; ++u) {
                const bool ok = (lc2 + CPP2 * u < jb);
                const double* src = (ok && nb) ? vsrc2 + (ll)(CPP2 * u) * p.ldv + r0 : V;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                             :: "r"(vdst2 + soff + (unsigned)(CPP2 * u * LDS * sizeof(double))), "l"(src), "r"(ok ? nb : 0));
            }
#pragma unroll
            for (int u = 0; u < (BN + CPP2 - 1) / CPP2; ++u) {
                const bool ok = (lc2 + CPP2 * u < ncb) && (lc2 + CPP2 * u < BN);
                const double* src = (ok && nb) ? csrc2 + (ll)(CPP2 * u) * p.ldc + r0 : C;
                if (lc2 + CPP2 * u < BN)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                                 :: "r"(cdst2 + soff + (unsigned)(CPP2 * u * LDS * sizeof(double))), "l"(src), "r"(ok ? nb : 0));
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            return;
        }
        const bool rok = (r0 + lr < rows);
#pragma unroll
        for (int u = 0; u < NB / CPP; ++u) {
            const bool ok = rok && (lc0 + CPP * u < jb);
            const double* src = ok ? vsrc0 + (ll)(CPP * u) * p.ldv + r0 : V;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                         :: "r"(vdst0 + soff + (unsigned)(CPP * u * LDS * sizeof(double))), "l"(src), "r"(ok ? 8 : 0));
        }
#pragma unroll
        for (int u = 0; u < (BN + CPP - 1) / CPP; ++u) {
            const bool ok = rok && (lc0 + CPP * u < ncb) && (lc0 + CPP * u < BN);
            const double* src = ok ? csrc0 + (ll)(CPP * u) * p.ldc + r0 : C;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                         :: "r"(cdst0 + soff + (unsigned)(CPP * u * LDS * sizeof(double))), "l"(src), "r"(ok ? 8 : 0));
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    // Streams the row chunks through `body(stage)`.  NSTAGE == 1: load, wait, compute.  NSTAGE == 2: the loads of
    // chunk c+1 are in flight while chunk c is computed.
    auto stream_rows = [&](auto body) {
        if (NSTAGE == 1) {
            for (int r0 = 0; r0 < rows; r0 += 64) {
                __syncthreads();
                issue_chunk(r0, 0);
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                __syncthreads();
                body(r0, 0);
            }
        } else {
            __syncthreads();
            issue_chunk(0, 0);
            int stage = 0;
            for (int r0 = 0; r0 < rows; r0 += 64) {
                if (r0 + 64 < rows) {
                    issue_chunk(r0 + 64, stage ^ 1);
                    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                }
                __syncthreads();
                body(r0, stage);
                __syncthreads();
                stage ^= 1;
            }
        }
    };

    // Warp tiling: BN == 64 -> 4 x 2 warps, each a 16 x 32 tile (2 x 4 fragments: 6 shared loads per 8 DMMA);
    // narrow blocks (BN <= 16) -> 8 x 1 warps, each 8 x BN.
    constexpr int NWARP = NTH / 32;
    constexpr bool WIDE = (BN >= 32);                     // 4 x (NWARP/4) warps of 16 x (BN / (NWARP/4)), else 8 x 1 warps of 8 x BN
    constexpr int NWN = WIDE ? NWARP / 4 : 1;             // warps along the columns
    constexpr int FMW = WIDE ? 2 : 1;                     // m-fragments per warp
    constexpr int FNW = WIDE ? BN / (8 * NWN) : FN;       // n-fragments per warp
    const int wm = WIDE ? (warp / NWN) * 16 : warp * 8;
    const int wn = WIDE ? (warp % NWN) * (BN / NWN) : 0;
    double acc[FMW][FNW][2];
#pragma unroll
    for (int i = 0; i < FMW; ++i)
#pragma unroll
        for (int j = 0; j < FNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // ---- pass 1: W = V^T C (rows of W = reflectors) -------------------------------------------------
    const bool warp_active = (wm < jb);
    stream_rows([&](int r0, int stage) {
        const double* Vc = Vs + stage * STAGE;
        const double* Cc = Cs + stage * STAGE;
        if (warp_active) {
#pragma unroll 4
            for (int kk = 0; kk < 64; kk += 4) {
                double a[FMW], b[FNW];
#pragma unroll
                for (int i = 0; i < FMW; ++i) a[i] = Vc[(wm + 8 * i + g) * LDS + kk + t4];
#pragma unroll
                for (int j = 0; j < FNW; ++j) b[j] = Cc[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
                for (int i = 0; i < FMW; ++i)
#pragma unroll
                    for (int j = 0; j < FNW; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
    });
    __syncthreads();
    // ---- W2 = -op(T) W ---------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < FMW; ++i)
#pragma unroll
        for (int j = 0; j < FNW; ++j) {
            Cs[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g] = acc[i][j][0];
            Cs[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g] = acc[i][j][1];
        }
    for (int e = tid; e < NB * NB; e += NTH) {
        const int l = e >> 6, i = e & 63;
        Vs[l * LDS + i] = (l < jb && i < jb) ? T[(ll)l * p.ldt + i] : 0.0;
    }
    __syncthreads();
    {   // W2 = -op(T) W on the DMMA pipe (T zero-padded to 64 x 64 in Vs, W in Cs): out[i][c] = sum_l op(T)[i][l] W[l][c]
        double o[FMW][FNW][2];
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) { o[i][j][0] = 0.0; o[i][j][1] = 0.0; }
#pragma unroll 4
        for (int kk = 0; kk < kmax; kk += 4) {
            double a[FMW], b[FNW];
#pragma unroll
            for (int i = 0; i < FMW; ++i)
                a[i] = p.trans_t ? Vs[(kk + t4) * LDS + wm + 8 * i + g] : Vs[(wm + 8 * i + g) * LDS + kk + t4];
#pragma unroll
            for (int j = 0; j < FNW; ++j) b[j] = Cs[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
            for (int i = 0; i < FMW; ++i)
#pragma unroll
                for (int j = 0; j < FNW; ++j) dmma_8x8x4(o[i][j][0], o[i][j][1], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                W2s[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g] = -o[i][j][0];
                W2s[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g] = -o[i][j][1];
            }
    }
    // ---- pass 2: C += V W2 (rows of the chunk x columns of the block) --------------------------------
    __syncthreads();
    stream_rows([&](int r0, int stage) {
        const double* Vc = Vs + stage * STAGE;
        const double* Cc = Cs + stage * STAGE;
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                acc[i][j][0] = Cc[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g];
                acc[i][j][1] = Cc[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g];
            }
        for (int kk = 0; kk < kmax; kk += 4) {
            double a[FMW], b[FNW];
#pragma unroll
            for (int i = 0; i < FMW; ++i) a[i] = Vc[(kk + t4) * LDS + wm + 8 * i + g];
#pragma unroll
            for (int j = 0; j < FNW; ++j) b[j] = W2s[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
            for (int i = 0; i < FMW; ++i)
#pragma unroll
                for (int j = 0; j < FNW; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < FMW; ++i) {
            const int rr = wm + 8 * i + g;
            if (r0 + rr < rows) {
#pragma unroll
                for (int j = 0; j < FNW; ++j) {
                    const int c = wn + 8 * j + 2 * t4;
                    if (c < ncb) C[(ll)c * p.ldc + r0 + rr] = acc[i][j][0];
                    if (c + 1 < ncb) C[(ll)(c + 1) * p.ldc + r0 + rr] = acc[i][j][1];
                }
            }
        }
    });
}

// Partial Gram matrices of a reflector panel, split over the rows: CTA (split, z) accumulates V^T V over the
// 64-row chunks split, split + nsplit, ... and writes its 64 x 64 partial to G[(z * nsplit + split) * 4096 + i * 64 + j]
// (larft_kernel sums the partials).  Same staging and DMMA tiling as pass 1 of larfb_kernel.
__global__ void __launch_bounds__(LARFB_THREADS) gram_kernel(const double* __restrict__ Vg, ll v_stride, int ldv, int rows, int jb,
                                                            double* __restrict__ G, ll g_stride, int nsplit) {
    __shared__ double Vs[LARFB_NB * LARFB_LDS];
    constexpr int LDS = LARFB_LDS, NB = LARFB_NB;
    const int split = blockIdx.x, z = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const double* V = Vg + (ll)z * v_stride;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int r0 = split * 64; r0 < rows; r0 += nsplit * 64) {
        __syncthreads();
#pragma unroll 4
        for (int e = tid; e < NB * 64; e += LARFB_THREADS) {
            const int r = e & 63, i = e >> 6;
            const bool ok = (r0 + r < rows && i < jb);
            const double* src = ok ? V + (ll)i * ldv + r0 + r : V;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(Vs + i * LDS + r);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(dst), "l"(src), "r"(ok ? 8 : 0));
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 64; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = Vs[(wm + 8 * i + g) * LDS + kk + t4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Vs[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double* out = G + (ll)z * g_stride + (ll)split * (NB * NB);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[(wm + 8 * i + g) * NB + wn + 8 * j + 2 * t4] = acc[i][j][0];
            out[(wm + 8 * i + g) * NB + wn + 8 * j + 2 * t4 + 1] = acc[i][j][1];
        }
}

// ------------------------------------------------------------------------------------------------------------------
// larfb2_kernel: the same update for 32-column blocks with a DOUBLE-BUFFERED 32-row chunk pipeline (one barrier per chunk,
// loads of chunk c+1 in flight while chunk c is multiplied) at the same shared-memory footprint (72.7 KB, 3 CTAs / SM) as the
// single-stage 64-row version above.  Pass 1 tiles W (64 reflectors x 32 columns) over 4 x 2 warps of 16 x 16; pass 2 tiles
// the 32 x 32 chunk of C over 4 x 2 warps of 8 x 16.  T (64 x 64) and W are staged across both chunk buffers at the pass boundary.
// ------------------------------------------------------------------------------------------------------------------
#define LARFB2_CH 32
#define LARFB2_LDC (LARFB2_CH + 4)

static size_t larfb2_smem() { return (size_t)(2 * (LARFB_NB + 32) * LARFB2_LDC + 32 * LARFB_LDS) * sizeof(double); }

__global__ void __launch_bounds__(LARFB_THREADS, 3) larfb2_kernel(const LarfbParams p) {
    extern __shared__ double sm[];
    constexpr int NTH = LARFB_THREADS, NB = LARFB_NB, BN = 32, CH = LARFB2_CH, LDC = LARFB2_LDC, LDS = LARFB_LDS;
    constexpr int STAGE = (NB + BN) * LDC;               // one stage: V chunk [NB][LDC] then C chunk [BN][LDC]
    double* W2s = sm + 2 * STAGE;                        // [BN][LDS]   W2s[c*LDS + i] = -(op(T) W)[i][c]
    double* Ts = sm;                                     // pass boundary: T [64][LDS] then W [BN][LDS] over the chunk buffers
    double* Ws = sm + NB * LDS;
    static_assert(NB * LARFB_LDS + 32 * LARFB_LDS <= 2 * (LARFB_NB + 32) * LARFB2_LDC, "T and W must fit the chunk buffers");
    const int z = blockIdx.y, cb = blockIdx.x;
    const int inst = p.idx ? p.idx[z] : z;
    const int nc = p.n_dyn ? p.n_dyn[(ll)inst * p.n_dyn_stride + p.n_dyn_off] : p.nc;
    const int c0 = cb * BN;
    if (c0 >= nc) return;
    const int ncb = min(BN, nc - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const double* V = p.V + (ll)z * p.v_stride;
    const double* T = p.T + (ll)z * p.t_stride;
    double* C = p.C + (ll)(p.c_idx ? inst : z) * p.c_stride + (ll)c0 * p.ldc;
    const int rows = p.rows, jb = p.jb;
    const int kmax = (jb + 3) & ~3;

    // staging: 16-byte path = row pair (lr2, lr2 + 1) of column lc2 + 16 u; 8-byte path = row lr of column lc0 + 8 u
    const unsigned sm_u = (unsigned)__cvta_generic_to_shared(sm);
    const int lr2 = (tid & 15) * 2, lc2 = tid >> 4;
    const int lr = tid & 31, lc0 = tid >> 5;
    const bool al16 = g_larfb_cp16 && ((((unsigned long long)V | (unsigned long long)C) & 15ull) == 0) && !(p.ldv & 1) && !(p.ldc & 1) && ((sm_u & 15u) == 0);
    // 16-byte path, slots precomputed: piece u of V is column lc2 + 16 u (source + u * 16 ldv, destination + u * 16 LDC doubles - a
    // compile-time offset), a chunk adds r0 to the source and the stage offset to the destination.  The loader used to rebuild every
    // address from column indices (64-bit multiplies, selects, bound tests): ~140 issued instructions per thread and chunk for six
    // copies, more than the 8 DMMA steps of the chunk they feed (SASS, profiles/r2_ncu_summary.md).  A copy that lies outside the
    // panel keeps its (unused) address and gets src-size 0: cp.async reads nothing then and zero-fills.
    const double* vsrc16 = V + (ll)lc2 * p.ldv + lr2;
    const double* csrc16 = C + (ll)lc2 * p.ldc + lr2;
    const ll vstep16 = 16LL * p.ldv, cstep16 = 16LL * p.ldc;
    const unsigned vdst16 = sm_u + (unsigned)((lc2 * LDC + lr2) * sizeof(double));
    const unsigned cdst16 = vdst16 + (unsigned)(NB * LDC * sizeof(double));
    unsigned okmask = 0;
#pragma unroll
    for (int u = 0; u < NB / 16; ++u) if (lc2 + 16 * u < jb) okmask |= 1u << u;
#pragma unroll
    for (int u = 0; u < BN / 16; ++u) if (lc2 + 16 * u < ncb) okmask |= 16u << u;
    auto issue_chunk = [&](int r0, int stage) {
        const unsigned vbase = sm_u + (unsigned)(stage * STAGE * sizeof(double));
        const unsigned cbase = vbase + (unsigned)(NB * LDC * sizeof(double));
        if (al16) {
            const int left = rows - (r0 + lr2);
            const int nb = left >= 2 ? 16 : (left == 1 ? 8 : 0);
            const unsigned so = (unsigned)(stage * STAGE * sizeof(double));
            const double* vs = vsrc16 + r0;
            const double* cs = csrc16 + r0;
#pragma unroll
            for (int u = 0; u < NB / 16; ++u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                             :: "r"(vdst16 + so + (unsigned)(u * 16 * LDC * sizeof(double))), "l"(vs + u * vstep16), "r"(((okmask >> u) & 1u) ? nb : 0));
#pragma unroll
            for (int u = 0; u < BN / 16; ++u)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                             :: "r"(cdst16 + so + (unsigned)(u * 16 * LDC * sizeof(double))), "l"(cs + u * cstep16), "r"(((okmask >> (4 + u)) & 1u) ? nb : 0));
        } else {
            const bool rok = (r0 + lr < rows);
#pragma unroll
            for (int u = 0; u < NB / 8; ++u) {
                const int col = lc0 + 8 * u;
                const bool ok = rok && (col < jb);
                const double* src = ok ? V + (ll)col * p.ldv + r0 + lr : V;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                             :: "r"(vbase + (unsigned)((col * LDC + lr) * sizeof(double))), "l"(src), "r"(ok ? 8 : 0));
            }
#pragma unroll
            for (int u = 0; u < BN / 8; ++u) {
                const int col = lc0 + 8 * u;
                const bool ok = rok && (col < ncb);
                const double* src = ok ? C + (ll)col * p.ldc + r0 + lr : C;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                             :: "r"(cbase + (unsigned)((col * LDC + lr) * sizeof(double))), "l"(src), "r"(ok ? 8 : 0));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    // chunk c is multiplied while chunk c + 1 loads; the single barrier of an iteration publishes stage s and retires stage s ^ 1
    auto stream_rows = [&](auto body) {
        __syncthreads();                                   // the chunk buffers are free (start of a pass)
        issue_chunk(0, 0);
        int stage = 0;
        for (int r0 = 0; r0 < rows; r0 += CH) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
            if (r0 + CH < rows) issue_chunk(r0 + CH, stage ^ 1);
            body(r0, stage);
            stage ^= 1;
        }
    };

    // ---- pass 1: W = V^T C, 4 x 2 warps of 16 reflectors x 16 columns --------------------------------
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
    double acc[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const bool warp_active = (wm < jb);
    stream_rows([&](int r0, int stage) {
        const double* Vc = sm + stage * STAGE;
        const double* Cc = Vc + NB * LDC;
        if (warp_active) {
#pragma unroll 4
            for (int kk = 0; kk < CH; kk += 4) {
                double a[2], b[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = Vc[(wm + 8 * i + g) * LDC + kk + t4];
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = Cc[(wn + 8 * j + g) * LDC + kk + t4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
    });
    __syncthreads();
    // ---- W2 = -op(T) W on the DMMA pipe --------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            Ws[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g] = acc[i][j][0];
            Ws[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g] = acc[i][j][1];
        }
    for (int e = tid; e < NB * NB; e += NTH) {
        const int l = e >> 6, i = e & 63;
        Ts[l * LDS + i] = (l < jb && i < jb) ? T[(ll)l * p.ldt + i] : 0.0;
    }
    __syncthreads();
    {
        double o[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) { o[i][j][0] = 0.0; o[i][j][1] = 0.0; }
#pragma unroll 4
        for (int kk = 0; kk < kmax; kk += 4) {
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                a[i] = p.trans_t ? Ts[(kk + t4) * LDS + wm + 8 * i + g] : Ts[(wm + 8 * i + g) * LDS + kk + t4];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Ws[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_8x8x4(o[i][j][0], o[i][j][1], a[i], b[j]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                W2s[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g] = -o[i][j][0];
                W2s[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g] = -o[i][j][1];
            }
    }
    // ---- pass 2: C += V W2, the 32 x 32 chunk of C over 4 x 2 warps of 8 rows x 16 columns ------------
    const int wm2 = (warp >> 1) * 8;
    // store slots of pass 2 (columns wn + 2 t4 + {0, 1, 8, 9}, row wm2 + g of the chunk): one base pointer, two strides, a 4-bit column mask
    double* cst = C + (ll)(wn + 2 * t4) * p.ldc + wm2 + g;
    const ll ldc1 = p.ldc, ldc8 = 8LL * p.ldc;
    const unsigned stmask = (wn + 2 * t4 < ncb ? 1u : 0u) | (wn + 2 * t4 + 1 < ncb ? 2u : 0u) | (wn + 8 + 2 * t4 < ncb ? 4u : 0u) | (wn + 9 + 2 * t4 < ncb ? 8u : 0u);
    stream_rows([&](int r0, int stage) {
        const double* Vc = sm + stage * STAGE;
        const double* Cc = Vc + NB * LDC;
        double c2[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            c2[j][0] = Cc[(wn + 8 * j + 2 * t4) * LDC + wm2 + g];
            c2[j][1] = Cc[(wn + 8 * j + 2 * t4 + 1) * LDC + wm2 + g];
        }
        for (int kk = 0; kk < kmax; kk += 4) {
            const double a = Vc[(kk + t4) * LDC + wm2 + g];
            double b[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = W2s[(wn + 8 * j + g) * LDS + kk + t4];
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma_8x8x4(c2[j][0], c2[j][1], a, b[j]);
        }
        if (r0 + wm2 + g < rows) {
            double* cw = cst + r0;
            if (stmask & 1u) cw[0] = c2[0][0];
            if (stmask & 2u) cw[ldc1] = c2[0][1];
            if (stmask & 4u) cw[ldc8] = c2[1][0];
            if (stmask & 8u) cw[ldc8 + ldc1] = c2[1][1];
        }
    });
}

// (Measured in round 2 and removed: a 64-column variant of this kernel - W2 held in registers as the B fragments of pass 2, warp
//  tiles 16 x 32 / 16 x 16, V staged once per 64 columns, 128 registers, 2 CTAs / SM - took 32.96 ms per sweep against 31.80 ms
//  for the 32-column kernel.  ncu on the largest launch of THIS kernel: FP64 tensor pipe 75.8 % busy, top stall
//  math_pipe_throttle - it is DMMA-bound; what separates it from 100 % is column padding (580 -> 608), wave quantisation
//  (2812 CTAs on 444 slots) and the pass boundary, not the shared-memory port.  profiles/r2_ncu_summary.md.)
static int g_larfb2 = 1;      // QTT_LARFB2=0: single-stage 64-row chunks (larfb_kernel<32,1>) for the wide case

template <int BN, int NSTAGE>
static size_t larfb_smem() { return (size_t)(NSTAGE * (LARFB_NB + BN) + BN) * LARFB_LDS * sizeof(double); }

// flops counted as the two GEMMs of the update (2 * 2 * rows * jb * nc) plus the small T product
static int larfb_launch(qtt_handle_s* h, cudaStream_t st, const LarfbParams& p, int G, int nc_bound) {
    if (p.rows <= 0 || p.jb <= 0 || nc_bound <= 0 || G <= 0) return QTT_OK;
    const double fl = (4.0 * p.rows * p.jb + 2.0 * p.jb * p.jb) * (double)nc_bound * G;
    h->gemm_flops += fl;
    h->launches += 1;
    size_t slot = 0;
    if (h->prof_on) {
        if (h->prof_used * 2 >= h->prof_ev.size()) prof_flush(h);
        slot = h->prof_used++;
        h->prof_fl[slot] = fl;
        cudaEventRecord(h->prof_ev[2 * slot], st);
    }
    if (nc_bound <= 8) {
        dim3 grid((unsigned)ceil_div(nc_bound, 8), G);
        larfb_kernel<8, 2><<<grid, LARFB_THREADS, larfb_smem<8, 2>(), st>>>(p);
    } else if (nc_bound <= 16) {
        dim3 grid((unsigned)ceil_div(nc_bound, 16), G);
        larfb_kernel<16, 2><<<grid, LARFB_THREADS, larfb_smem<16, 2>(), st>>>(p);
    } else {
        dim3 grid((unsigned)ceil_div(nc_bound, 64), G);
        // 64 columns, single stage: two CTAs per SM overlap each other's loads (measured faster than one
        // double-buffered CTA per SM: 112 vs 122 ms per sweep)
        if (g_larfb2) {
            dim3 grid32((unsigned)ceil_div(nc_bound, 32), G);
            larfb2_kernel<<<grid32, LARFB_THREADS, larfb2_smem(), st>>>(p);
        } else if (!QTT_ENV_FLAG("QTT_LARFB_BN64")) {
            dim3 grid32((unsigned)ceil_div(nc_bound, 32), G);
            larfb_kernel<32, 1><<<grid32, LARFB_THREADS, larfb_smem<32, 1>(), st>>>(p);
        } else larfb_kernel<64, 1, 512><<<grid, 512, larfb_smem<64, 1>(), st>>>(p);
    }
    if (h->prof_on) cudaEventRecord(h->prof_ev[2 * slot + 1], st);
    tl_mark(h, st, 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string("larfb launch: ") + cudaGetErrorString(e); return QTT_ERR_CUDA; }
    return QTT_OK;
}

