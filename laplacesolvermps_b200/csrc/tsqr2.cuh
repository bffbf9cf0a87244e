// Second-generation tall-skinny panel leaf: the flat-tree factorisation of tsqr.cuh with every row block factorised in
// SHARED MEMORY by the left-looking scheme of qr2.cuh (8-column sub-panels, one thread per stacked row).
//
// Why: the left-looking full-height leaf (qr2.cuh) re-reads the earlier reflectors of the 64-column outer panel twice per
// sub-panel - ~14 panel volumes from HBM per panel, because the reflector store of a batch (148 x 1.3 MB per panel) exceeds
// L2 - and ncu shows its two reflector passes running at the HBM rate.  The first tall-skinny leaf (tsqr_leaf_kernel) read
// the panel once but kept a row block in REGISTERS, which made every column step ~420 warp instructions (issue bound).
// Here a stacked row block [R_{b-1}; 256 body rows] (320 x 64, 166 KB) sits in shared memory, is factorised in place and
// leaves as reflectors: the panel is read once and written once, the earlier sub-panels of the block are re-read from
// shared memory, and a column step is one row-local rank-1 update in registers + one barrier.
//
// Same contract as tsqr_leaf_kernel (TsLeafParams; representation consumed by larfb_ts_kernel):
//     block 0:  S = P[0 : 320]                       -> V_0 = [L; V_0b], T_0, R_0
//     block b:  S = [R_{b-1}; P[64 + 256 b : ...]]   -> V_b = [I; V_bb], T_b, R_b        (R_{b-1} upper triangular)
// Q_panel = Q_0 Q_1 ... Q_{nblk-1}, Q_b = I - Y_b T_b Y_b^T.  For b >= 1 the strictly lower part of the stacked top rows is
// exactly zero before and after every step (0 * x = 0), so the identity top of Y_b needs no special code.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "tsqr.cuh"

#define TS2_THREADS TS_SROWS            // one stacked row per thread (320 = 10 warps)
#define TS2_NW (TS2_THREADS / 32)
#define TS2_LD (TS_SROWS + 4)           // column stride of the block in shared memory; == 4 (mod 16): DMMA fragment reads are conflict free
#define TS2_CR (TS2_NW * 8 + 8)         // per-column reduction buffer: 8 partial dots per warp + the pivot row

static size_t tsqr2_leaf_smem() {
    return (size_t)(64 * TS2_LD + 64 * 64 + TS2_NW * 4 * 64 + 2 * 512 + 64 + 64 + 2 * TS2_CR) * sizeof(double);
}

// out[(8 t + i) * 8 + c] = sum_r v(r, 8 t + i) * B(r, cB0 + c)   for the reflector groups t < ng,
// v = clean reflectors of the block held in S (unit diagonal, zero above it inside the 64 top rows); B = columns cB0 .. cB0 + 7
// of S, raw (CLEANB = false: the sub-panel being updated) or clean (CLEANB = true: Gram blocks for T).
template <bool CLEANB>
__device__ __forceinline__ void ts2_vt_times(const double* S, int ng, int cB0, int ksteps, double* red, double* out, int tid) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int t = warp % ng, part = warp / ng, nparts = (TS2_NW - 1 - t) / ng + 1;
    const int i = 8 * t + g, cB = cB0 + g;
    const double* At = S + i * TS2_LD + t4;
    const double* Bc = S + cB * TS2_LD + t4;
    double acc[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u][0] = 0.0; acc[u][1] = 0.0; }
    // the 64 top rows (k < 16) carry the triangular structure; four independent accumulator pairs break the DMMA dependency chain
    int k = part;
    for (; k < 16 && k < ksteps; k += nparts) {
        const int r = 4 * k + t4;
        double a = At[4 * k], b = Bc[4 * k];
        a = r < i ? 0.0 : (r == i ? 1.0 : a);
        if (CLEANB) b = r < cB ? 0.0 : (r == cB ? 1.0 : b);
        dmma_8x8x4(acc[0][0], acc[0][1], a, b);
    }
    for (; k + 3 * nparts < ksteps; k += 4 * nparts) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = At[4 * (k + u * nparts)]; b[u] = Bc[4 * (k + u * nparts)]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) dmma_8x8x4(acc[u][0], acc[u][1], a[u], b[u]);
    }
    for (; k < ksteps; k += nparts) dmma_8x8x4(acc[1][0], acc[1][1], At[4 * k], Bc[4 * k]);
    red[warp * 64 + g * 8 + 2 * t4] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    red[warp * 64 + g * 8 + 2 * t4 + 1] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    __syncthreads();
    for (int e = tid; e < ng * 64; e += TS2_THREADS) {
        const int tt = e >> 6, idx = e & 63;
        double x = 0.0;
        for (int wq = tt; wq < TS2_NW; wq += ng) x += red[wq * 64 + idx];
        out[e] = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TS2_THREADS, 1) tsqr2_leaf_kernel(const TsLeafParams p) {
    extern __shared__ double sm[];
    double* S = sm;                              // [64][TS2_LD] stacked block, column-major, factorised in place
    double* Tsm = S + 64 * TS2_LD;               // T of the current row block, 64 x 64 row-major
    double* red = Tsm + 64 * 64;                 // TS2_NW * 64 (x4: room for the widest pass)
    double* Wa = red + TS2_NW * 4 * 64;          // 64 x 8: V^T C, then the Gram blocks
    double* W2 = Wa + 512;                       // 64 x 8
    double* Tl = W2 + 512;                       // 8 x 8 T of the sub-panel
    double* taus = Tl + 64;                      // 64
    double* colred = taus + 64;                  // 2 x TS2_CR, alternating by column parity
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.x;
    const int rows = p.a - p.j0;
    double* M = p.M + (ll)z * p.m_stride + (ll)p.j0 * p.ld + p.j0;      // panel origin: element (r, c) at M[c * ld + r]
    double* V = p.V + (ll)z * p.v_stride + (ll)p.j0 * p.ldv + p.j0;
    double* Tq = p.Tq + (ll)z * p.t_stride;
    const unsigned full = 0xffffffffu;
    double* Srow = S + tid;                      // this thread's stacked row: column c at Srow[c * TS2_LD]

    for (int e = tid; e < 64 * 64; e += TS2_THREADS) Tsm[e] = 0.0;      // the strictly lower part stays zero
    {
        const bool ok = tid < rows;
#pragma unroll 16
        for (int c = 0; c < 64; ++c) Srow[c * TS2_LD] = ok ? M[(ll)c * p.ld + tid] : 0.0;
    }
    __syncthreads();

    for (int b = 0; b < p.nblk; ++b) {
        const int rb0 = TS_TOP + b * TS_BS;
        const int nbody = min(TS_BS, rows - rb0);
        const int ksteps = (TS_TOP + nbody + 3) >> 2;              // 4-row DMMA steps that hold data (rows beyond are zero)
        for (int s = 0; s < 8; ++s) {
            const int c0 = 8 * s, P = c0;
            double x[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) x[c] = Srow[(c0 + c) * TS2_LD];
            // ---- block reflector of the earlier sub-panels of this row block: C -= V_prev (T_prev^T (V_prev^T C)) -------------
            if (s > 0) {
                ts2_vt_times<false>(S, s, c0, ksteps, red, Wa, tid);
                for (int e = tid; e < P * 8; e += TS2_THREADS) {
                    const int i = e >> 3, c = e & 7;
                    double w = 0.0;
                    for (int l = 0; l <= i; ++l) w += Tsm[l * 64 + i] * Wa[l * 8 + c];
                    W2[e] = w;
                }
                __syncthreads();
                const int ilim = tid < TS_TOP ? min(P, tid) : P;   // reflectors i > row are zero at this row, i == row is the unit diagonal
                for (int i = 0; i < ilim; ++i) {
                    const double v = Srow[i * TS2_LD];
                    const double2* w2 = reinterpret_cast<const double2*>(W2 + i * 8);
                    const double2 w01 = w2[0], w23 = w2[1], w45 = w2[2], w67 = w2[3];
                    x[0] -= v * w01.x; x[1] -= v * w01.y; x[2] -= v * w23.x; x[3] -= v * w23.y;
                    x[4] -= v * w45.x; x[5] -= v * w45.y; x[6] -= v * w67.x; x[7] -= v * w67.y;
                }
                if (tid < P) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[c] -= W2[tid * 8 + c];
                }
            }
            // ---- 8 Householder columns: the row lives in registers, one barrier per column ---------------------------------
            double part[8];
            {
                const bool below = tid > c0;
#pragma unroll
                for (int c = 0; c < 8; ++c) part[c] = below ? x[0] * x[c] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pr = c0 + j;
                double* cb = colred + (j & 1) * TS2_CR;
                {   // warp sums of the 8 partials by a transposing butterfly (see qr2.cuh)
                    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
                    double a4[4], a2[2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double mine = h4 ? part[4 + k] : part[k], theirs = h4 ? part[k] : part[4 + k];
                        a4[k] = mine + __shfl_xor_sync(full, theirs, 16);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const double mine = h3 ? a4[2 + k] : a4[k], theirs = h3 ? a4[k] : a4[2 + k];
                        a2[k] = mine + __shfl_xor_sync(full, theirs, 8);
                    }
                    double tot = (h2 ? a2[1] : a2[0]) + __shfl_xor_sync(full, h2 ? a2[0] : a2[1], 4);
                    tot += __shfl_xor_sync(full, tot, 2);
                    tot += __shfl_xor_sync(full, tot, 1);
                    if ((lane & 3) == 0) cb[warp * 8 + (h4 ? 4 : 0) + (h3 ? 2 : 0) + (h2 ? 1 : 0)] = tot;
                }
                if (tid == pr) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) cb[TS2_NW * 8 + c] = (j + c < 8) ? x[(j + c) & 7] : 0.0;
                }
                __syncthreads();
                // every warp finishes the sums and derives the scalars redundantly: no second barrier
                double d = 0.0, pv = 0.0;
                if (lane < 8) {
#pragma unroll
                    for (int wq = 0; wq < TS2_NW; ++wq) d += cb[wq * 8 + lane];
                    pv = cb[TS2_NW * 8 + lane];
                }
                const double xnorm2 = __shfl_sync(full, d, 0);
                const double alpha = __shfl_sync(full, pv, 0);
                double tj = 0.0, beta = alpha, inv = 0.0;
                if (xnorm2 > 0.0) {
                    const double n2 = alpha * alpha + xnorm2;
                    const double rs = rsqrt(n2);
                    const double nrm = n2 * rs;
                    beta = -copysign(nrm, alpha);
                    tj = fma(fabs(alpha), rs, 1.0);
                    inv = copysign(__drcp_rn(fabs(alpha) + nrm), alpha);
                }
                const double fl = tj * (pv + d * inv);              // lane c: coefficient of column j + c
                if (tid == 0) taus[pr] = tj;
                double f[8];
#pragma unroll
                for (int c = 1; c < 8; ++c) f[c] = __shfl_sync(full, fl, c);
                if (tid == pr) {
                    if (tj != 0.0) x[j] = beta;
#pragma unroll
                    for (int c = 1; c < 8; ++c) if (j + c < 8) x[(j + c) & 7] -= f[c];
                } else if (tid > pr) {
                    const double v = (tj != 0.0) ? x[j] * inv : 0.0;
                    x[j] = v;
#pragma unroll
                    for (int c = 1; c < 8; ++c) if (j + c < 8) x[(j + c) & 7] -= f[c] * v;
                }
                if (j < 7) {
                    const bool below = tid > pr + 1;
#pragma unroll
                    for (int c = 0; c < 8; ++c) part[c] = (below && j + 1 + c < 8) ? x[(j + 1) & 7] * x[(j + 1 + c) & 7] : 0.0;
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) Srow[(c0 + c) * TS2_LD] = x[c];
            __syncthreads();
            // ---- T of the row block: Gram blocks [V_prev V_s]^T V_s, 8 x 8 recurrence, T12 = -T11 (G T22) ---------------------
            ts2_vt_times<true>(S, s + 1, c0, ksteps, red, Wa, tid);
            if (tid < 8) {
                const double* Gk = Wa + P * 8;
                const int r = tid;
                double trow[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) trow[i] = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double ti = taus[c0 + i];
                    if (i == r) trow[i] = ti;
                    else if (i > r) {
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c >= r && c < i) acc += trow[c] * Gk[c * 8 + i];
                        trow[i] = -ti * acc;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { Tl[r * 8 + i] = trow[i]; Tsm[(P + r) * 64 + P + i] = trow[i]; }
            }
            __syncthreads();
            if (s > 0) {
                for (int e = tid; e < P * 8; e += TS2_THREADS) {
                    const int i = e >> 3, c = e & 7;
                    double w = 0.0;
#pragma unroll
                    for (int l = 0; l < 8; ++l) w += Wa[i * 8 + l] * Tl[l * 8 + c];
                    W2[e] = w;
                }
                __syncthreads();
                for (int e = tid; e < P * 8; e += TS2_THREADS) {
                    const int i = e >> 3, c = e & 7;
                    double w = 0.0;
                    for (int l = i; l < P; ++l) w += Tsm[i * 64 + l] * W2[l * 8 + c];
                    Tsm[i * 64 + P + c] = -w;
                }
                __syncthreads();
            }
        }
        // ---- T and the reflectors of the block out; R stays as the top of the next stacked block, next body in ------------
        {
            double* Tb = Tq + (ll)b * 4096;
            for (int e = tid; e < 4096; e += TS2_THREADS) Tb[e] = Tsm[e];
        }
        if (tid < TS_TOP) {
            if (b == 0) {
#pragma unroll 16
                for (int c = 0; c < 64; ++c) V[(ll)c * p.ldv + tid] = tid < c ? 0.0 : (tid == c ? 1.0 : Srow[c * TS2_LD]);
            }
            if (b + 1 < p.nblk) for (int c = 0; c < tid; ++c) Srow[c * TS2_LD] = 0.0;
        } else {
            const int rr = tid - TS_TOP;
            if (rr < nbody) {
#pragma unroll 16
                for (int c = 0; c < 64; ++c) V[(ll)c * p.ldv + rb0 + rr] = Srow[c * TS2_LD];
            }
            if (b + 1 < p.nblk) {
                const int rn0 = rb0 + TS_BS;
                const bool ok = rr < min(TS_BS, rows - rn0);
#pragma unroll 16
                for (int c = 0; c < 64; ++c) Srow[c * TS2_LD] = ok ? M[(ll)c * p.ld + rn0 + rr] : 0.0;
            }
        }
        __syncthreads();
    }
    // ---- R into the working matrix (upper triangle of the top rows) ----------------------------------------------------------
    if (tid < TS_TOP)
        for (int c = tid; c < 64; ++c) M[(ll)c * p.ld + tid] = Srow[c * TS2_LD];
}
