// Tall-skinny (flat-tree) Householder QR of a 64-column outer panel, and the matching block-reflector update.
//
// Why: the left-looking leaf (qr2.cuh) re-reads the earlier reflectors of the outer panel twice per 8-column sub-panel
// (640 column transfers per 64-column panel) and the reflector store of a whole batch (148 x 1.3 MB per panel) does not
// fit L2, so that leaf is HBM-bound.  Here the panel is factorised in ROW blocks that live in registers:
//
//     block 0:  S = P[0 : 64 + 256]                 -> Householder QR, V_0 = [L; V_0b], T_0, R_0
//     block b:  S = [R_{b-1}; P[64 + 256 b : ...]]  -> Householder QR, V_b = [I; V_bb], T_b, R_b     (R_{b-1} upper triangular)
//
// so Q_panel = Q_0 Q_1 ... Q_{nblk-1} with Q_b = I - Y_b T_b Y_b^T, Y_b = V_b embedded at the 64 top rows of the panel
// and the rows of block b.  The panel is read once and written once (192 column transfers), same flop count as the
// plain Householder QR.  The V store keeps its layout (a x q column-major): rows of block b hold V_bb, the 64 top rows
// hold L of block 0; only T changes (one 64 x 64 T per row block instead of one per panel).  (PLASMA-style TS kernels;
// numerics identical in kind to Householder QR - R can differ by row signs, Q R = M holds to round-off.)
//
// STATUS: correct (tests/test_gpu_parity.py::test_tall_skinny_qr_path_against_oracle) but OFF by default (QTT_TS_MIN_ROWS=0):
// measured on B200 at the C5 shapes the sweep takes 110 ms with it against 103 ms without.  The row-block factorisation
// multiplies the number of sequential column steps by the number of row blocks (10 x 64 per panel) and each step costs
// ~420 warp instructions per warp (FP64 shuffle reductions, redundant sqrt / divisions, masks), so the leaf is
// issue-bound at ~110 us per 256-row block - no faster than the HBM-bound left-looking leaf it replaces; larfb_ts itself
// runs at the same time per wave as larfb_kernel (73 % of the DMMA path).  Kept as an opt-in experiment.
//
// tsqr_leaf_kernel:  one CTA (512 threads) per instance and outer panel.  Warp w owns columns w, w+16, w+32, w+48; lane l
//   owns stacked rows l, l+32, ... (10 per column) - the whole 320 x 64 stacked block sits in registers.  Per column:
//   the owner publishes the updated, not yet normalised column x; after ONE barrier every warp takes the dots of x with
//   its columns and with x itself (one butterfly reduction), derives beta / tau redundantly and updates the columns
//   right of j or emits the Gram entries v_c^T v_j for the columns left of j, from which T is built
//   (blocked T12 = -T11 G12 T22, row-parallel, no further barriers).
// larfb_ts_kernel:   C <- Q^T C (trans_t = 1, blocks ascending, T^T) or C <- Q C (trans_t = 0, blocks descending, T) for
//   the TS representation; same staging / DMMA tiling as larfb_kernel.  The 64 top rows of the CTA's column block stay
//   in registers in accumulator layout between the row blocks (they enter W = C_top + V_b^T C_b as the initial value of
//   the accumulators and receive C_top -= W2), so the identity part of Y_b costs no DMMA.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "larfb.cuh"

#define TS_TOP 64
#define TS_BS 256
#define TS_SROWS (TS_TOP + TS_BS)
#define TS_THREADS 512
#define TS_NW (TS_THREADS / 32)
#define TS_CPW (64 / TS_NW)
#define TS_RPL (TS_SROWS / 32)
#define TS_GLD 65

struct TsLeafParams {
    double* M; ll m_stride; int ld;
    double* V; ll v_stride; int ldv;
    double* Tq; ll t_stride;               // T of row block 0 of this panel; block b at + b * 4096 (row-major 64 x 64)
    int a, j0, nblk;
};

static size_t tsqr_leaf_smem() { return (size_t)(2 * TS_SROWS + 64 * TS_GLD + 64 * 64 + 64 + 8) * sizeof(double); }

__global__ void __launch_bounds__(TS_THREADS, 1) tsqr_leaf_kernel(const TsLeafParams p) {
    extern __shared__ double sm[];
    double* vbuf = sm;                          // [2][TS_SROWS] reflector being applied (double buffered)
    double* Gt = vbuf + 2 * TS_SROWS;           // Gt[j * TS_GLD + c] = v_c^T v_j, c < j
    double* Tsm = Gt + 64 * TS_GLD;             // T of the current row block, row-major
    double* taus = Tsm + 64 * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.x;
    const int rows = p.a - p.j0;
    double* M = p.M + (ll)z * p.m_stride + (ll)p.j0 * p.ld + p.j0;      // panel origin: element (r, c) at M[c * ld + r]
    double* V = p.V + (ll)z * p.v_stride + (ll)p.j0 * p.ldv + p.j0;
    double* Tq = p.Tq + (ll)z * p.t_stride;
    const unsigned full = 0xffffffffu;

    double s[TS_CPW][TS_RPL];                   // s[t][i]: stacked row lane + 32 i of column warp + 16 t
    for (int e = tid; e < 64 * 64; e += TS_THREADS) Tsm[e] = 0.0;       // the strictly lower part stays zero
#pragma unroll
    for (int t = 0; t < TS_CPW; ++t) {
        const double* col = M + (ll)(warp + TS_NW * t) * p.ld;
        s[t][0] = col[lane];
        s[t][1] = col[lane + 32];
    }
    auto load_body = [&](int b) {
        const int rb0 = TS_TOP + b * TS_BS;
        const int nbody = min(TS_BS, rows - rb0);
#pragma unroll
        for (int t = 0; t < TS_CPW; ++t) {
            const double* col = M + (ll)(warp + TS_NW * t) * p.ld + rb0;
#pragma unroll
            for (int i = 2; i < TS_RPL; ++i) {
                const int rr = lane + 32 * (i - 2);
                s[t][i] = (rr < nbody) ? col[rr] : 0.0;
            }
        }
    };
    load_body(0);

    for (int b = 0; b < p.nblk; ++b) {
        const int rb0 = TS_TOP + b * TS_BS;
        const int nbody = min(TS_BS, rows - rb0);
        // ---- Householder QR of the stacked block: one barrier and one reduction round per column ------------------
        // The owner of column j publishes the (updated, not yet normalised) column x; every warp then takes the dots of
        // x below the diagonal with its own columns AND with x itself, derives beta / tau / 1/(alpha - beta) redundantly
        // (no second broadcast), and applies  v^T s_c = inv * x^T s_c + s_c[j]  to update columns right of j or to emit
        // the Gram entry v_c^T v_j for columns left of j.
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < TS_RPL; ++i) vbuf[lane + 32 * i] = s[0][i];
        }
        for (int j = 0; j < 64; ++j) {
            const double* xb = vbuf + (j & 1) * TS_SROWS;
            double* xn = vbuf + ((j + 1) & 1) * TS_SROWS;
            __syncthreads();
            double x[TS_RPL], d[TS_CPW], srow[TS_CPW], ss = 0.0;
            const double alpha = xb[j];
#pragma unroll
            for (int i = 0; i < TS_RPL; ++i) {
                const double xv = xb[lane + 32 * i];
                x[i] = (lane + 32 * i > j) ? xv : 0.0;
            }
#pragma unroll
            for (int t = 0; t < TS_CPW; ++t) {
                d[t] = 0.0;
                srow[t] = __shfl_sync(full, (j >> 5) ? s[t][1] : s[t][0], j & 31);
            }
#pragma unroll
            for (int i = 0; i < TS_RPL; ++i) {
                ss += x[i] * x[i];
#pragma unroll
                for (int t = 0; t < TS_CPW; ++t) d[t] += x[i] * s[t][i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ss += __shfl_xor_sync(full, ss, o);
#pragma unroll
                for (int t = 0; t < TS_CPW; ++t) d[t] += __shfl_xor_sync(full, d[t], o);
            }
            double tj = 0.0, beta = alpha, inv = 0.0;
            if (ss > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + ss), alpha);
                tj = (beta - alpha) / beta;
                inv = 1.0 / (alpha - beta);
            }
#pragma unroll
            for (int i = 0; i < TS_RPL; ++i) x[i] *= inv;                 // v below the diagonal (zero above)
#pragma unroll
            for (int t = 0; t < TS_CPW; ++t) {
                const int c = warp + TS_NW * t;
                if (c == j) {
#pragma unroll
                    for (int i = 0; i < TS_RPL; ++i) {
                        const int r = lane + 32 * i;
                        if (r > j) s[t][i] = x[i]; else if (r == j) s[t][i] = beta;
                    }
                } else {
                    const double e = inv * d[t] + srow[t];                // v_j^T s_c
                    if (c > j) {
                        const double w = tj * e;
#pragma unroll
                        for (int i = 0; i < TS_RPL; ++i) s[t][i] -= w * x[i];
                        if ((j & 31) == lane) { if (j >> 5) s[t][1] -= w; else s[t][0] -= w; }
                        if (c == j + 1) {
#pragma unroll
                            for (int i = 0; i < TS_RPL; ++i) xn[lane + 32 * i] = s[t][i];
                        }
                    } else if (lane == 0) {
                        Gt[j * TS_GLD + c] = e;
                    }
                }
            }
            if (tid == 0) taus[j] = tj;
        }
        // ---- reflectors out, R stays as the top of the next stacked block, next body in -------------------------
#pragma unroll
        for (int t = 0; t < TS_CPW; ++t) {
            const int c = warp + TS_NW * t;
            double* vcol = V + (ll)c * p.ldv;
#pragma unroll
            for (int i = 2; i < TS_RPL; ++i) {
                const int rr = lane + 32 * (i - 2);
                if (rr < nbody) vcol[rb0 + rr] = s[t][i];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = lane + 32 * i;
                if (b == 0) vcol[r] = (r < c) ? 0.0 : (r == c ? 1.0 : s[t][i]);
                if (r > c) s[t][i] = 0.0;
            }
        }
        if (b + 1 < p.nblk) load_body(b + 1);
        __syncthreads();
        // ---- T of the block from the Gram entries: diagonal 8 x 8 blocks, then T12 = -T11 G12 T22 row by row ----
        {
            const int i = tid >> 3, sub = tid & 7;
            if (sub == 0) {
                const int jend = i | 7;
                Tsm[i * 64 + i] = taus[i];
                for (int jj = i + 1; jj <= jend; ++jj) {
                    double x = 0.0;
                    for (int l = i; l < jj; ++l) x += Tsm[i * 64 + l] * Gt[jj * TS_GLD + l];
                    Tsm[i * 64 + jj] = -taus[jj] * x;
                }
            }
            __syncthreads();
            for (int J = 1; J < 8; ++J) {
                const int jb0 = 8 * J;
                double w = 0.0;
                if (i < jb0)
                    for (int l = i; l < jb0; ++l) w += Tsm[i * 64 + l] * Gt[(jb0 + sub) * TS_GLD + l];
                double out = 0.0;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const double wm = __shfl_sync(full, w, (lane & ~7) | m);
                    if (m <= sub) out += wm * Tsm[(jb0 + m) * 64 + jb0 + sub];
                }
                if (i < jb0) Tsm[i * 64 + jb0 + sub] = -out;
                __syncwarp();
            }
        }
        __syncthreads();
        double* Tb = Tq + (ll)b * 4096;
        for (int e = tid; e < 4096; e += TS_THREADS) Tb[e] = Tsm[e];
    }
    // ---- R into the working matrix (upper triangle of the top rows) ----------------------------------------------
#pragma unroll
    for (int t = 0; t < TS_CPW; ++t) {
        const int c = warp + TS_NW * t;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lane + 32 * i;
            if (r <= c) M[(ll)c * p.ld + r] = s[t][i];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------

struct LarfbTsParams {
    const double* V; ll v_stride; int ldv;     // panel origin (row j0, column j0): element (r, i) at V[i*ldv + r]
    const double* Tq; ll t_stride;             // T of row block 0; block b at + b * 4096
    double* C; ll c_stride; int ldc; int c_idx; // rows from j0: element (r, c) at C[c*ldc + r]
    int rows, nblk, nc, trans_t;
    const int* idx;
    const int* n_dyn; int n_dyn_stride, n_dyn_off;
};

template <int BN, int NSTAGE>
__global__ void __launch_bounds__(LARFB_THREADS, (BN >= 32 ? 3 : 2)) larfb_ts_kernel(const LarfbTsParams p) {
    extern __shared__ double sm[];
    constexpr int NTH = LARFB_THREADS;
    constexpr int LDS = LARFB_LDS, NB = LARFB_NB;
    constexpr int STAGE = (NB + BN) * LDS;
    double* Vs = sm;
    double* Cs = Vs + NB * LDS;
    double* W2s = sm + NSTAGE * STAGE;
    const int z = blockIdx.y, cb = blockIdx.x;
    const int inst = p.idx ? p.idx[z] : z;
    const int nc = p.n_dyn ? p.n_dyn[(ll)inst * p.n_dyn_stride + p.n_dyn_off] : p.nc;
    const int c0 = cb * BN;
    if (c0 >= nc) return;
    const int ncb = min(BN, nc - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const double* V = p.V + (ll)z * p.v_stride;
    const double* Tq = p.Tq + (ll)z * p.t_stride;
    double* C = p.C + (ll)(p.c_idx ? inst : z) * p.c_stride + (ll)c0 * p.ldc;
    const int rows = p.rows;

    // staging of a 64-row chunk [r0, r0 + 64) limited to rows < rend (see larfb_kernel)
    constexpr int CPP = NTH / 64;
    const int lr = tid & 63, lc0 = tid >> 6;
    const double* vsrc0 = V + (ll)lc0 * p.ldv + lr;
    const double* csrc0 = C + (ll)lc0 * p.ldc + lr;
    const unsigned vdst0 = (unsigned)__cvta_generic_to_shared(Vs + lc0 * LDS + lr);
    const unsigned cdst0 = (unsigned)__cvta_generic_to_shared(Cs + lc0 * LDS + lr);
    constexpr int CPP2 = NTH / 32;
    const int lr2 = (tid & 31) * 2, lc2 = tid >> 5;
    const unsigned vdst2 = (unsigned)__cvta_generic_to_shared(Vs + lc2 * LDS + lr2);
    const unsigned cdst2 = (unsigned)__cvta_generic_to_shared(Cs + lc2 * LDS + lr2);
    const bool al16 = g_larfb_cp16 && ((((unsigned long long)V | (unsigned long long)C) & 15ull) == 0) && !(p.ldv & 1) && !(p.ldc & 1) &&
                      ((vdst2 & 15u) == 0);
    const double* vsrc2 = V + (ll)lc2 * p.ldv + lr2;
    const double* csrc2 = C + (ll)lc2 * p.ldc + lr2;
    auto issue_chunk = [&](int r0, int rend, int stage) {
        const unsigned soff = (unsigned)(stage * STAGE * sizeof(double));
        if (al16) {
            const int left = rend - (r0 + lr2);
            const int nb = left >= 2 ? 16 : (left == 1 ? 8 : 0);
#pragma unroll
            for (int u = 0; u < NB / CPP2; ++u) {
                const double* src = nb ? vsrc2 + (ll)(CPP2 * u) * p.ldv + r0 : V;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                             :: "r"(vdst2 + soff + (unsigned)(CPP2 * u * LDS * sizeof(double))), "l"(src), "r"(nb));
            }
#pragma unroll
            for (int u = 0; u < (BN + CPP2 - 1) / CPP2; ++u) {
                const bool ok = (lc2 + CPP2 * u < ncb);
                const double* src = (ok && nb) ? csrc2 + (ll)(CPP2 * u) * p.ldc + r0 : C;
                if (lc2 + CPP2 * u < BN)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                                 :: "r"(cdst2 + soff + (unsigned)(CPP2 * u * LDS * sizeof(double))), "l"(src), "r"(ok ? nb : 0));
            }
        } else {
            const bool rok = (r0 + lr < rend);
#pragma unroll
            for (int u = 0; u < NB / CPP; ++u) {
                const double* src = rok ? vsrc0 + (ll)(CPP * u) * p.ldv + r0 : V;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                             :: "r"(vdst0 + soff + (unsigned)(CPP * u * LDS * sizeof(double))), "l"(src), "r"(rok ? 8 : 0));
            }
#pragma unroll
            for (int u = 0; u < (BN + CPP - 1) / CPP; ++u) {
                const bool ok = rok && (lc0 + CPP * u < ncb) && (lc0 + CPP * u < BN);
                const double* src = ok ? csrc0 + (ll)(CPP * u) * p.ldc + r0 : C;
                if (lc0 + CPP * u < BN)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n"
                                 :: "r"(cdst0 + soff + (unsigned)(CPP * u * LDS * sizeof(double))), "l"(src), "r"(ok ? 8 : 0));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto stream_rows = [&](int rbeg, int rend, auto body) {
        if (NSTAGE == 1) {
            for (int r0 = rbeg; r0 < rend; r0 += 64) {
                __syncthreads();
                issue_chunk(r0, rend, 0);
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                __syncthreads();
                body(r0, 0);
            }
        } else {
            __syncthreads();
            issue_chunk(rbeg, rend, 0);
            int stage = 0;
            for (int r0 = rbeg; r0 < rend; r0 += 64) {
                if (r0 + 64 < rend) {
                    issue_chunk(r0 + 64, rend, stage ^ 1);
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

    constexpr int NWARP = NTH / 32;
    constexpr bool WIDE = (BN >= 32);
    constexpr int NWN = WIDE ? NWARP / 4 : 1;
    constexpr int FMW = WIDE ? 2 : 1;
    constexpr int FNW = WIDE ? BN / (8 * NWN) : BN / 8;
    const int wm = WIDE ? (warp / NWN) * 16 : warp * 8;
    const int wn = WIDE ? (warp % NWN) * (BN / NWN) : 0;
    double acc[FMW][FNW][2], ctop[FMW][FNW][2];

    auto pass1 = [&](int rbeg, int rend) {                     // acc += V^T C over the rows [rbeg, rend)
        stream_rows(rbeg, rend, [&](int r0, int stage) {
            const double* Vc = Vs + stage * STAGE;
            const double* Cc = Cs + stage * STAGE;
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
        });
        __syncthreads();
    };
    auto tstep = [&](const double* T) {                        // W2s = -(op(T) W), W = acc
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                Cs[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g] = acc[i][j][0];
                Cs[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g] = acc[i][j][1];
            }
        for (int e = tid; e < NB * NB; e += NTH) {
            const int l = e >> 6, i = e & 63;
            Vs[l * LDS + i] = T[e];
        }
        __syncthreads();
        double o[FMW][FNW][2];
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) { o[i][j][0] = 0.0; o[i][j][1] = 0.0; }
#pragma unroll 4
        for (int kk = 0; kk < 64; kk += 4) {
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
        __syncthreads();
    };
    auto pass2 = [&](int rbeg, int rend, bool keep_top) {      // C += V W2 over the rows [rbeg, rend)
        stream_rows(rbeg, rend, [&](int r0, int stage) {
            const double* Vc = Vs + stage * STAGE;
            const double* Cc = Cs + stage * STAGE;
#pragma unroll
            for (int i = 0; i < FMW; ++i)
#pragma unroll
                for (int j = 0; j < FNW; ++j) {
                    acc[i][j][0] = Cc[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g];
                    acc[i][j][1] = Cc[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g];
                }
#pragma unroll 4
            for (int kk = 0; kk < 64; kk += 4) {
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
            if (keep_top && r0 == 0) {
#pragma unroll
                for (int i = 0; i < FMW; ++i)
#pragma unroll
                    for (int j = 0; j < FNW; ++j) { ctop[i][j][0] = acc[i][j][0]; ctop[i][j][1] = acc[i][j][1]; }
            }
#pragma unroll
            for (int i = 0; i < FMW; ++i) {
                const int rr = wm + 8 * i + g;
                if (r0 + rr < rend) {
#pragma unroll
                    for (int j = 0; j < FNW; ++j) {
                        const int c = wn + 8 * j + 2 * t4;
                        if (c < ncb) C[(ll)c * p.ldc + r0 + rr] = acc[i][j][0];
                        if (c + 1 < ncb) C[(ll)(c + 1) * p.ldc + r0 + rr] = acc[i][j][1];
                    }
                }
            }
        });
    };
    auto zero_acc = [&]() {
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    };
    auto acc_from_top = [&]() {
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) { acc[i][j][0] = ctop[i][j][0]; acc[i][j][1] = ctop[i][j][1]; }
    };
    auto top_add_w2 = [&]() {                                  // C_top -= W2  (W2s holds the negated product)
#pragma unroll
        for (int i = 0; i < FMW; ++i)
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                ctop[i][j][0] += W2s[(wn + 8 * j + 2 * t4) * LDS + wm + 8 * i + g];
                ctop[i][j][1] += W2s[(wn + 8 * j + 2 * t4 + 1) * LDS + wm + 8 * i + g];
            }
    };
    auto top_store = [&]() {
#pragma unroll
        for (int i = 0; i < FMW; ++i) {
            const int rr = wm + 8 * i + g;
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                const int c = wn + 8 * j + 2 * t4;
                if (c < ncb) C[(ll)c * p.ldc + rr] = ctop[i][j][0];
                if (c + 1 < ncb) C[(ll)(c + 1) * p.ldc + rr] = ctop[i][j][1];
            }
        }
    };
    auto top_load = [&]() {
#pragma unroll
        for (int i = 0; i < FMW; ++i) {
            const int rr = wm + 8 * i + g;
#pragma unroll
            for (int j = 0; j < FNW; ++j) {
                const int c = wn + 8 * j + 2 * t4;
                ctop[i][j][0] = (c < ncb) ? C[(ll)c * p.ldc + rr] : 0.0;
                ctop[i][j][1] = (c + 1 < ncb) ? C[(ll)(c + 1) * p.ldc + rr] : 0.0;
            }
        }
    };

    const int e0 = min(TS_TOP + TS_BS, rows);                  // block 0 = top rows + first body block, contiguous
    if (p.trans_t) {
        zero_acc();
        pass1(0, e0);
        tstep(Tq);
        pass2(0, e0, true);
        for (int b = 1; b < p.nblk; ++b) {
            const int rb = TS_TOP + b * TS_BS, re = min(rb + TS_BS, rows);
            acc_from_top();
            pass1(rb, re);
            tstep(Tq + (ll)b * 4096);
            top_add_w2();
            pass2(rb, re, false);
        }
        if (p.nblk > 1) top_store();
    } else {
        if (p.nblk > 1) {
            top_load();
            for (int b = p.nblk - 1; b >= 1; --b) {
                const int rb = TS_TOP + b * TS_BS, re = min(rb + TS_BS, rows);
                acc_from_top();
                pass1(rb, re);
                tstep(Tq + (ll)b * 4096);
                top_add_w2();
                pass2(rb, re, false);
            }
            top_store();
            __threadfence_block();
            __syncthreads();
        }
        zero_acc();
        pass1(0, e0);
        tstep(Tq);
        pass2(0, e0, false);
    }
}

template <int BN, int NSTAGE>
static size_t larfb_ts_smem() { return larfb_smem<BN, NSTAGE>(); }

static int larfb_ts_launch(qtt_handle_s* h, cudaStream_t st, const LarfbTsParams& p, int G, int nc_bound) {
    if (p.rows <= 0 || p.nblk <= 0 || nc_bound <= 0 || G <= 0) return QTT_OK;
    const double fl = (4.0 * p.rows * 64.0 + 2.0 * 64.0 * 64.0 * p.nblk) * (double)nc_bound * G;
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
        dim3 grid(1, G);
        larfb_ts_kernel<8, 2><<<grid, LARFB_THREADS, larfb_ts_smem<8, 2>(), st>>>(p);
    } else {
        dim3 grid((unsigned)ceil_div(nc_bound, 32), G);
        larfb_ts_kernel<32, 1><<<grid, LARFB_THREADS, larfb_ts_smem<32, 1>(), st>>>(p);
    }
    if (h->prof_on) cudaEventRecord(h->prof_ev[2 * slot + 1], st);
    tl_mark(h, st, 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string("larfb_ts launch: ") + cudaGetErrorString(e); return QTT_ERR_CUDA; }
    return QTT_OK;
}
