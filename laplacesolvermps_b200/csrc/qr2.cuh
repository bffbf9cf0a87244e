// Second-generation panel leaf of the blocked Householder QR.
//
// Same contract as qr_leaf_kernel (qr.cuh) - factorise one 8-column sub-panel of the current 64-column outer
// panel, left-looking - but built for latency: one CTA per instance is alone on its SM, so what matters is the
// number of dependent phases, not flops.
//   * the s earlier sub-panels of the outer panel are applied as ONE block reflector with the panel's compact-WY
//     factor T built so far (one DMMA pass  W = V_prev^T C, one update pass), instead of s sequential ones;
//   * during the factorisation of the 8 columns every thread owns fixed rows (tid, tid+NT, ...) of the shared-memory
//     sub-panel, so the rank-1 update with reflector j and the dot products for reflector j+1 are one pass with
//     two barriers per column;
//   * the kernel extends the outer panel's T in place ( T12 = -T11 (V1^T V2) T22 ), so no Gram / dlarft launch is
//     needed once the panel is complete.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "qr.cuh"


// column stride of the shared-memory sub-panel: the smallest value >= rows that is 8 (mod 16)
__host__ __device__ static inline int qr_leaf2_ldc(int rows) { return rows + ((24 - (rows & 15)) & 15); }

struct Leaf2Params {
    double* M; ll m_stride; int ld;
    double* V; ll v_stride; int ldv;
    double* tau; ll tau_stride;
    double* Tp; ll t_stride;               // T of the current outer panel, 64 x 64 row-major
    int a, j0, i0, w;
    int last;                              // 1: last sub-panel of its outer panel (finishes the panel's T itself)
    int vres;                              // 1 (only with nsub > 1): the panel's reflectors are also kept in shared memory (short panels)
    int nsub, jend;                        // nsub > 1: the launch walks nsub consecutive sub-panels i0, i0 + 8, .. of the panel ending at
                                           // column jend (w and last are then derived per sub-panel) - one launch per outer panel
};

// out[t*64 + i*8 + c] = sum_r V_t[r][i] * B(r, c)  for the s previous sub-panels t (8 reflectors each).
// Warps are split into s groups; group t streams the rows of V_t (zero above row 8t).
// If out2 != nullptr the same pass also produces out2[t*64 + i*8 + c] = sum_r V_t[r][i] * V_{s-1}[r][c] for t < s-1
// (the Gram block that completes column block s-1 of the panel's T), reading V_{s-1} from the reflector store.
template <int NT, typename BL, typename BL2>
__device__ __forceinline__ void vprev_t_times(int s, const double* Vj0, int ldv, int rows, BL bload, BL2 bload2, double* red, double* out,
                                              double* out2, int tid) {
    constexpr int NW = NT / 32, U = 4;
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int t = warp % s, part = warp / s, nparts = (NW - 1 - t) / s + 1;
    const double* Vt = Vj0 + (ll)(8 * t) * ldv;             // column 8t of the panel, row index relative to j0
    const bool second = (out2 != nullptr) && (t < s - 1);
    const double* Vl = Vj0 + (ll)(8 * (s - 1)) * ldv;       // previous sub-panel (zero above row 8(s-1))
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    // 16-byte loads (two consecutive rows per lane, two DMMA k-steps per load) halve the number of strided global load
    // instructions of this pass - its L1TEX work - when the reflector columns are 16-byte aligned
    const bool vec2 = ((((unsigned long long)Vj0) & 15ull) == 0) && !(ldv & 1);
    if (vec2) {
        constexpr int U2 = 2;
        // L2 prefetch PF iterations ahead: the pass is bound by the latency of its strided reflector loads out of HBM (14 GB/s per SM
        // measured, a third of the SM's share of the HBM rate) and the 64-register budget of the 1024-thread CTA leaves no room for
        // deeper register prefetch; a prefetch instruction keeps requests in flight without holding registers.  One lane per 64-byte
        // column segment issues it (t4 == 0).
        constexpr int PF = 6;
        const int pf_step = 8 * nparts * U2 * PF;
        for (int r = 8 * t + 8 * part; r < rows; r += 8 * nparts * U2) {
            if (t4 == 0) {
#pragma unroll
                for (int u = 0; u < U2; ++u) {
                    const int rp = r + u * 8 * nparts + pf_step;
                    if (rp < rows) {
                        asm volatile("prefetch.global.L2 [%0];\n" :: "l"(Vt + (ll)g * ldv + rp));
                        if (second) asm volatile("prefetch.global.L2 [%0];\n" :: "l"(Vl + (ll)g * ldv + rp));
                    }
                }
            }
            double2 av[U2], bv[U2], gv[U2];
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int rr = r + u * 8 * nparts + 2 * t4;
                av[u] = make_double2(0.0, 0.0); bv[u] = av[u]; gv[u] = av[u];
                if (rr + 1 < rows) {
                    av[u] = *reinterpret_cast<const double2*>(Vt + (ll)g * ldv + rr);
                    bv[u] = bload2(rr, g);                       // rows rr, rr + 1 (rr is even): one 16-byte shared load
                    if (second) gv[u] = *reinterpret_cast<const double2*>(Vl + (ll)g * ldv + rr);
                } else if (rr < rows) {
                    av[u].x = Vt[(ll)g * ldv + rr]; bv[u].x = bload(rr, g);
                    if (second) gv[u].x = Vl[(ll)g * ldv + rr];
                }
            }
#pragma unroll
            for (int u = 0; u < U2; ++u) { dmma_8x8x4(c0, c1, av[u].x, bv[u].x); dmma_8x8x4(c0, c1, av[u].y, bv[u].y); }
            if (second) {
#pragma unroll
                for (int u = 0; u < U2; ++u) { dmma_8x8x4(d0, d1, av[u].x, gv[u].x); dmma_8x8x4(d0, d1, av[u].y, gv[u].y); }
            }
        }
    } else
    for (int r = 8 * t + 4 * part; r < rows; r += 4 * nparts * U) {
        double av[U], bv[U], gv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = r + u * 4 * nparts + t4;
            av[u] = 0.0; bv[u] = 0.0; gv[u] = 0.0;
            if (rr < rows) {
                av[u] = Vt[(ll)g * ldv + rr]; bv[u] = bload(rr, g);
                if (second) gv[u] = Vl[(ll)g * ldv + rr];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dmma_8x8x4(c0, c1, av[u], bv[u]);
        if (second) {
#pragma unroll
            for (int u = 0; u < U; ++u) dmma_8x8x4(d0, d1, av[u], gv[u]);
        }
    }
    red[warp * 128 + g * 8 + 2 * t4] = c0;
    red[warp * 128 + g * 8 + 2 * t4 + 1] = c1;
    red[warp * 128 + 64 + g * 8 + 2 * t4] = d0;
    red[warp * 128 + 64 + g * 8 + 2 * t4 + 1] = d1;
    __syncthreads();
    for (int e = tid; e < s * 64; e += NT) {
        const int tt = e >> 6, idx = e & 63;
        double x = 0.0, y = 0.0;
        for (int wq = tt; wq < NW; wq += s) { x += red[wq * 128 + idx]; y += red[wq * 128 + 64 + idx]; }
        out[e] = x;
        if (out2) out2[e] = y;
    }
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT) qr_leaf2_kernel(const Leaf2Params p) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int z = blockIdx.x;
    const int rows = p.a - p.j0;
    // Column stride of the sub-panel in shared memory: rows padded to 8 (mod 16) doubles.  The DMMA passes read it fragment-shaped -
    // lane (g, t4) takes rows (2 t4, 2 t4 + 1) of column g as one 16-byte load - and every panel of the full-size cores has a multiple
    // of 16 rows (2576 - 64 k): with stride = rows the 8 columns of a fragment fell into the SAME banks (ncu: 31 % of the kernel's
    // shared wavefronts were conflict replays, the W = V_prev^T C pass was bound by them, not by HBM); with the padded stride the
    // 16-byte loads of a quarter-warp (2 columns x 4 row pairs) cover the 8 bank groups exactly once.
    const int ldc = qr_leaf2_ldc(rows);
    double* Cs = sm;                             // [8][ldc] column-major
    double* red = Cs + (size_t)QR_W * ldc;       // NW * 64
    constexpr int RED = (NW * 128 > 4096) ? NW * 128 : 4096;   // also stages the panel's T (64 x 64) between the passes
    double* Wa = red + RED;                      // 64 * 8: W = V_prev^T C, later G = V_prev^T V_s
    double* W2 = Wa + 512;                       // 64 * 8
    double* Tl = W2 + 512;                       // 8 x 8 T of this sub-panel
    double* Wk = Tl + 64;                        // 8 x 8 scratch
    double* sc = Wk + 64;                        // [0] tau_j [1] inv [2] beta, [8+c] f_c, [16+j] taus
    double* piv = sc + 32;                       // pivot row, columns j..w-1
    double* Ga = piv + 8;                        // 64 * 8: Gram block V_{0..s-2}^T V_{s-1}
    double* M = p.M + (ll)z * p.m_stride;
    double* V = p.V + (ll)z * p.v_stride;
    double* tau = p.tau + (ll)z * p.tau_stride;
    double* Tp = p.Tp + (ll)z * p.t_stride;
    const double* Vj0 = V + (ll)p.j0 * p.ldv + p.j0;     // panel reflectors: element (r, i) at Vj0[i*ldv + r]
    // short panels: the earlier sub-panels are read back from a shared-memory copy (column stride rows + 4: 16-byte aligned, conflict free)
    double* Vsm = Ga + 512;
    const int ldvs = rows + 4;
    const double* Vp = p.vres ? Vsm : Vj0;
    const int ldp = p.vres ? ldvs : p.ldv;

    const int nsub = p.nsub > 1 ? p.nsub : 1;
    for (int it = 0; it < nsub; ++it) {
    const int i0 = p.i0 + it * QR_W;
    const int w = (nsub > 1) ? min(QR_W, p.jend - i0) : p.w;
    const int last = (nsub > 1) ? (i0 + QR_W >= p.jend ? 1 : 0) : p.last;
    const int s = (i0 - p.j0) / QR_W;            // earlier sub-panels of this outer panel
    const int P = s * QR_W;
    // 1. load the sub-panel (thread-owned rows, all 8 column loads of a row in flight; columns >= w are zero)
    {
        const double* Msub = M + (ll)i0 * p.ld + p.j0;
        for (int r = tid; r < rows; r += NT) {
            double tmp[QR_W];
#pragma unroll
            for (int c = 0; c < QR_W; ++c) tmp[c] = (c < w) ? Msub[(ll)c * p.ld + r] : 0.0;
#pragma unroll
            for (int c = 0; c < QR_W; ++c) Cs[c * ldc + r] = tmp[c];
        }
    }
    __syncthreads();

    // 2. W2 = T_prev^T (V_prev^T C): one block reflector for all earlier sub-panels.  The same pass over V_prev yields
    //    G = V_{0..s-2}^T V_{s-1}, which completes column block s-1 of the panel's T (T12 = -T11 G T22) - the previous
    //    leaf left only its diagonal block - so no leaf but the last of a panel needs a pass of its own for that.
    if (s > 0) {
        vprev_t_times<NT>(s, Vp, ldp, rows, [&](int r, int c) { return Cs[c * ldc + r]; },
                          [&](int r, int c) { return *reinterpret_cast<const double2*>(Cs + c * ldc + r); }, red, Wa, s > 1 ? Ga : nullptr, tid);
        // the panel's T built so far, staged in shared memory (the reduction scratch is free here): the small triangular
        // products below used to walk T in global memory at a stride of a row - a latency chain of several microseconds
        double* Tsm = red;
        for (int e = tid; e < P * 64; e += NT) Tsm[e] = Tp[e];
        __syncthreads();
        if (s > 1) {
            const int P1 = P - 8;                            // reflectors before sub-panel s-1
            for (int e = tid; e < P1 * 8; e += NT) {         // W2 = G T22
                const int i = e >> 3, c = e & 7;
                double x = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l) x += Ga[i * 8 + l] * Tsm[(P1 + l) * 64 + P1 + c];
                W2[e] = x;
            }
            __syncthreads();
            for (int e = tid; e < P1 * 8; e += NT) {
                const int i = e >> 3, c = e & 7;
                double x = 0.0;
                for (int l = i; l < P1; ++l) x += Tsm[i * 64 + l] * W2[l * 8 + c];
                Tsm[i * 64 + P1 + c] = -x;
                Tp[i * 64 + P1 + c] = -x;
            }
            __syncthreads();
        }
        for (int e = tid; e < P * 8; e += NT) {
            const int i = e >> 3, c = e & 7;
            double x = 0.0;
            for (int l = 0; l <= i; ++l) x += Tsm[l * 64 + i] * Wa[l * 8 + c];
            W2[e] = x;
        }
        __syncthreads();
    }

    // 3. update with the earlier reflectors.  From here on every thread owns the rows tid, tid + NT, ... of the
    //    sub-panel (its shared-memory words are private to it), so consecutive passes need no barrier.
    if (s > 0) {
        for (int r = tid; r < rows; r += NT) {
            double c8[QR_W];
#pragma unroll
            for (int c = 0; c < QR_W; ++c) c8[c] = Cs[c * ldc + r];
            for (int g0 = 0; g0 < P; g0 += 8) {
                double vv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = Vp[(ll)(g0 + i) * ldp + r];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < QR_W; ++c) c8[c] -= vv[i] * W2[(g0 + i) * 8 + c];
            }
#pragma unroll
            for (int c = 0; c < QR_W; ++c) Cs[c * ldc + r] = c8[c];
        }
    }

    // (while the 8 columns below run out of shared memory the memory system idles: pull the NEXT sub-panel of the working matrix towards
    //  L2 now - its load at the top of the next iteration is 9 % of the kernel's stall samples)
    if (it + 1 < nsub) {
        const double* Mnext = M + (ll)(i0 + QR_W) * p.ld + p.j0;
        const int nlines = (rows + 15) >> 4;
        const int wn = min(QR_W, p.jend - (i0 + QR_W));
        for (int e = tid; e < wn * nlines; e += NT) {
            const int c = e / nlines, l = e - c * nlines;
            asm volatile("prefetch.global.L2 [%0];\n" :: "l"(Mnext + (ll)c * p.ld + 16 * l));
        }
    }

    // 4. Householder on the 8 columns: pass j = rank-1 update with reflector j fused with the dot products of
    //    reflector j + 1; two barriers per column (reduction in, scalars out)
    const int pr0 = i0 - p.j0;
    double part[QR_W];
    auto dots_for = [&](int j) {                         // partial sums over this thread's rows below pivot j
#pragma unroll
        for (int c = 0; c < QR_W; ++c) part[c] = 0.0;
        for (int r = tid; r < rows; r += NT) {
            if (r <= pr0 + j) continue;
            const double x = Cs[j * ldc + r];
            part[0] += x * x;
#pragma unroll
            for (int c = 1; c < QR_W; ++c)
                if (j + c < w) part[c] += x * Cs[(j + c) * ldc + r];
        }
    };
    dots_for(0);
    for (int j = 0; j < w; ++j) {
        const int pr = pr0 + j;
        {   // warp sums of the 8 partials by a transposing butterfly: each round halves the values a lane carries (4 + 2 + 1 exchanges),
            // two plain rounds finish - 9 shuffles instead of 40; lane l ends with the total of column 4 b4 + 2 b3 + b2 of its index
            const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
            double a4[4], a2[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double mine = h4 ? part[4 + k] : part[k], theirs = h4 ? part[k] : part[4 + k];
                a4[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double mine = h3 ? a4[2 + k] : a4[k], theirs = h3 ? a4[k] : a4[2 + k];
                a2[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, 8);
            }
            double tot = (h2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? a2[0] : a2[1], 4);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            if ((lane & 3) == 0) red[warp * QR_W + (h4 ? 4 : 0) + (h3 ? 2 : 0) + (h2 ? 1 : 0)] = tot;
        }
        __syncthreads();
        if (warp == 0) {
            // the pivot row: lane c < 8 reads column j + c (nobody writes the panel between the two barriers)
            const double pv = (lane < QR_W && j + lane < w) ? Cs[(j + lane) * ldc + pr] : 0.0;
            // lane l sums column (l & 7) over a quarter of the warps' partials, two shuffle rounds finish the sums
            double d = 0.0;
            {
                const int cq = lane & 7, part = lane >> 3;
#pragma unroll
                for (int u = 0; u < NW / 4; ++u) d += red[(part * (NW / 4) + u) * QR_W + cq];
                d += __shfl_xor_sync(0xffffffffu, d, 8);
                d += __shfl_xor_sync(0xffffffffu, d, 16);
            }
            const double xnorm2 = __shfl_sync(0xffffffffu, d, 0);
            const double alpha = __shfl_sync(0xffffffffu, pv, 0);
            double tj = 0.0, beta = alpha, inv = 0.0;
            if (xnorm2 > 0.0) {
                // one reciprocal square root and one reciprocal instead of sqrt + two divisions on the serial path of the column:
                // beta = -sign(alpha) nrm,  tau = (beta - alpha) / beta = 1 + |alpha| / nrm,  1 / (alpha - beta) = sign(alpha) / (|alpha| + nrm)
                const double n2 = alpha * alpha + xnorm2;
                const double rs = rsqrt(n2);
                const double nrm = n2 * rs;
                beta = -copysign(nrm, alpha);
                tj = fma(fabs(alpha), rs, 1.0);
                inv = copysign(__drcp_rn(fabs(alpha) + nrm), alpha);
            }
            if (lane >= 1 && lane < QR_W) sc[8 + lane] = (j + lane < w) ? tj * (pv + d * inv) : 0.0;
            if (lane == 0) { sc[0] = tj; sc[1] = inv; sc[2] = beta; sc[16 + j] = tj; tau[i0 + j] = tj; }
        }
        __syncthreads();
        const double tj = sc[0], inv = sc[1], beta = sc[2];
        double f[QR_W];
#pragma unroll
        for (int c = 1; c < QR_W; ++c) f[c] = sc[8 + c];
        // the dot products of reflector j + 1 are taken from the updated values while they are in registers
#pragma unroll
        for (int c = 0; c < QR_W; ++c) part[c] = 0.0;
        for (int r = tid; r < rows; r += NT) {
            if (r < pr) continue;
            if (r == pr) {
                if (tj != 0.0) Cs[j * ldc + r] = beta;
#pragma unroll
                for (int c = 1; c < QR_W; ++c)
                    if (j + c < w) Cs[(j + c) * ldc + r] -= f[c];
            } else {
                const double v = (tj != 0.0) ? Cs[j * ldc + r] * inv : 0.0;
                Cs[j * ldc + r] = v;
                double nv[QR_W];
#pragma unroll
                for (int c = 1; c < QR_W; ++c) {
                    nv[c] = 0.0;
                    if (j + c < w) { nv[c] = Cs[(j + c) * ldc + r] - f[c] * v; Cs[(j + c) * ldc + r] = nv[c]; }
                }
                if (r > pr + 1) {
                    const double x = nv[1];
#pragma unroll
                    for (int c = 1; c < QR_W; ++c) part[c - 1] += x * nv[c];
                }
            }
        }
    }
    __syncthreads();

    // 5. T of the sub-panel and the extension of the outer panel's T
    auto vclean = [&](int r, int i) -> double {          // r relative to j0; reflector i of this sub-panel
        if (i >= w) return 0.0;
        const int pv = pr0 + i;
        if (r < pv) return 0.0;
        if (r == pv) return 1.0;
        return Cs[i * ldc + r];
    };
    // T of the sub-panel (8 x 8)
    gram8<NT>(rows - pr0, [&](int r, int i) { return vclean(pr0 + r, i); }, [&](int r, int i) { return vclean(pr0 + r, i); }, red, Wk, tid);
    if (tid < 8) {                                       // row r of T depends on row r only: 8 independent recurrences
        const int r = tid;
        double trow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) trow[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < w) {
                const double ti = sc[16 + i];
                if (i == r) trow[i] = ti;
                else if (i > r) {
                    double x = 0.0;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c >= r && c < i) x += trow[c] * Wk[c * 8 + i];
                    trow[i] = -ti * x;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Tl[r * 8 + i] = trow[i];
    }
    __syncthreads();
    // extend the outer panel's T:  T12 = -T11 (V_prev^T V_s) T22,  T22 = Tl,  T21 = 0
    if (s > 0 && last) {
        vprev_t_times<NT>(s, Vp, ldp, rows, vclean, [&](int r, int c) { return make_double2(vclean(r, c), vclean(r + 1, c)); },
                          red, Wa, nullptr, tid);  // Wa = G (P x 8)
        for (int e = tid; e < P * 8; e += NT) {                                 // W2 = G Tl
            const int i = e >> 3, c = e & 7;
            double x = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) x += Wa[i * 8 + l] * Tl[l * 8 + c];
            W2[e] = x;
        }
        __syncthreads();
        double* Tsm = red;                               // vprev_t_times is done with the scratch
        for (int e = tid; e < P * 64; e += NT) Tsm[e] = Tp[e];
        __syncthreads();
        for (int e = tid; e < P * 8; e += NT) {
            const int i = e >> 3, c = e & 7;
            double x = 0.0;
            for (int l = i; l < P; ++l) x += Tsm[i * 64 + l] * W2[l * 8 + c];
            Tp[i * 64 + P + c] = -x;
        }
    }
    for (int e = tid; e < 8 * 64; e += NT) {                                    // rows P..P+7 of T
        const int i = e >> 6, l = e & 63;
        Tp[(P + i) * 64 + l] = (l >= P && l < P + 8) ? Tl[i * 8 + (l - P)] : 0.0;
    }

    // 6. write back: R part into M (rows <= pivot), clean V into the reflector store
    for (int c = 0; c < w; ++c) {
        const int prc = pr0 + c;
        for (int r = tid; r < rows; r += NT) {
            const double x = Cs[c * ldc + r];
            if (r <= prc) M[(ll)(i0 + c) * p.ld + p.j0 + r] = x;
            const double vx = (r < prc) ? 0.0 : (r == prc ? 1.0 : x);
            V[(ll)(i0 + c) * p.ldv + p.j0 + r] = vx;
            if (p.vres) Vsm[(ll)(i0 - p.j0 + c) * ldvs + r] = vx;
        }
    }
    if (it + 1 < nsub) { __threadfence_block(); __syncthreads(); }    // the next sub-panel reads this one's reflectors, T and R from global memory
    }
}

static size_t qr_leaf2_smem(int rows, int nt, bool vres = false) {
    return ((size_t)QR_W * qr_leaf2_ldc(rows) + std::max((nt / 32) * 128, 4096) + 512 * 3 + 64 * 2 + 32 + 8 + (vres ? (size_t)64 * (rows + 4) : 0)) * sizeof(double);
}
