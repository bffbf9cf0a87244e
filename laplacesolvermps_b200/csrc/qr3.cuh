// Multi-CTA panel leaf for TALL panels, in two flavours of the same code:
//   CL = true   thread-block CLUSTER (sm_90+/sm_100): the C CTAs of an instance form one cluster, every CTA keeps its row
//               slice of the sub-panel in its own shared memory and the per-phase reductions read the partial sums of the
//               peer CTAs through distributed shared memory (cluster.map_shared_rank) between hardware cluster barriers.
//               No global scratch, no co-residency requirement on the whole grid: any number of instances, panels of up to
//               16 x ~2500 rows (rank caps 5..60 at 2D L = 12: 5152 .. 38640 rows).
//   CL = false  cooperative launch (the round-1 path, kept for A/B): reductions through a global scratch and grid.sync();
//               needs C * G <= resident CTAs.
//
// qr_leaf2_kernel keeps an 8-column sub-panel of one instance in the shared memory of ONE CTA, which limits it to
// ~3400 rows; beyond that the first-generation leaf falls back to single-column sub-panels (0.6 ms per column at 18432
// rows).  Here the rows of the sub-panel are split over C CTAs of a cooperative launch: every CTA runs the phases of the
// second-generation leaf on its row slice and the per-phase reductions (V_prev^T C, the dot products of each
// Householder column, the Gram blocks for T) go through a double-buffered global scratch and a grid-wide barrier.
// Same outputs as qr_leaf2_kernel: clean V, tau, R part of the working matrix, the panel's T extended in place.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "gemm.cuh"
#include "qr.cuh"
#include "qr2.cuh"

struct Leaf3Params {
    double* M; ll m_stride; int ld;
    double* V; ll v_stride; int ldv;
    double* tau; ll tau_stride;
    double* Tp; ll t_stride;
    double* scratch; ll scratch_stride;        // per instance: 2 * (C + 1) * 1024 doubles
    int a, j0, i0, w, last;
    int C, slice;                              // CTAs per instance, rows per CTA
};

template <int NT, bool CL>
__global__ void __launch_bounds__(NT) qr_leaf3_kernel(const Leaf3Params p) {
    extern __shared__ double sm[];
    namespace cg = cooperative_groups;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int cta = blockIdx.x, z = blockIdx.y, C = p.C;
    const int rows_all = p.a - p.j0;
    const int r_lo = cta * p.slice;                                  // first row of this CTA, relative to j0
    const int nloc = max(0, min(rows_all, r_lo + p.slice) - r_lo);   // rows held here
    const int w = p.w;
    const int s = (p.i0 - p.j0) / QR_W, P = s * QR_W;
    const int pr0 = p.i0 - p.j0;
    double* Cs = sm;                                  // [8][slice]
    double* red = Cs + (size_t)QR_W * p.slice;        // NW * 128
    double* Wa = red + NW * 128;                      // 512
    double* W2 = Wa + 512;                            // 512
    double* Ga = W2 + 512;                            // 512
    double* Tl = Ga + 512;                            // 64
    double* Wk = Tl + 64;                             // 64
    double* sc = Wk + 64;                             // 32
    double* part0 = sc + 32;                          // 1024: CTA-local partial of the reduction in flight
    double* part = part0;                             // (cluster flavour: two buffers, 2048 doubles, used alternately - see grid_reduce)
    double* M = p.M + (ll)z * p.m_stride;
    double* V = p.V + (ll)z * p.v_stride;
    double* tau = p.tau + (ll)z * p.tau_stride;
    double* Tp = p.Tp + (ll)z * p.t_stride;
    double* scratch = p.scratch + (ll)z * p.scratch_stride;
    const double* Vj0 = V + (ll)p.j0 * p.ldv + p.j0;
    int buf = 0;

    // sum over the CTAs of an instance of `n` (<= 1024) values held in part[]; result in out[] (shared) on every CTA
    auto grid_reduce = [&](int n, double* out) {
        if constexpr (CL) {
            // part (= part0 + buf * 1024) holds this CTA's partial.  One hardware cluster barrier publishes the partials of all
            // CTAs; every CTA then sums its peers' values out of their shared memory.  The buffers alternate, so a peer that is
            // still reading this CTA's buffer of reduction k cannot be overtaken: the buffer is rewritten only for reduction
            // k + 2, i.e. after barrier k + 1, at which that peer arrives only when its reads of reduction k are done.
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            for (int e = tid; e < n; e += NT) {
                double x = 0.0;
                for (int c = 0; c < C; ++c) x += cluster.map_shared_rank(part, c)[e];
                out[e] = x;
            }
            __syncthreads();
            buf ^= 1;
            part = part0 + buf * 1024;
            return;
        }
        cg::grid_group grid = cg::this_grid();
        double* mine = scratch + ((ll)buf * (C + 1) + cta) * 1024;
        for (int e = tid; e < n; e += NT) mine[e] = part[e];
        __threadfence();
        grid.sync();
        const double* base = scratch + (ll)buf * (C + 1) * 1024;
        for (int e = tid; e < n; e += NT) {
            double x = 0.0;
            for (int c = 0; c < C; ++c) x += base[(ll)c * 1024 + e];
            out[e] = x;
        }
        __syncthreads();
        buf ^= 1;
    };
    // CTA-local part[t*64 + i*8 + c] = sum over local rows of V_t[r][i] * B(r, c) (and, with second, part[512 + ..] with V_{s-1})
    auto local_vprev_t = [&](auto bload, bool want_second) {
        constexpr int U = 4;
        const int t = warp % s, pw = warp / s, nparts = (NW - 1 - t) / s + 1;
        const double* Vt = Vj0 + (ll)(8 * t) * p.ldv;
        const double* Vl = Vj0 + (ll)(8 * (s - 1)) * p.ldv;
        const bool second = want_second && (t < s - 1);
        const int lstart = max(8 * t - r_lo, 0);                      // V_t is zero above panel row 8t
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        for (int r = lstart + 4 * pw; r < nloc; r += 4 * nparts * U) {
            double av[U], bv[U], gv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + u * 4 * nparts + t4;
                av[u] = 0.0; bv[u] = 0.0; gv[u] = 0.0;
                if (rr < nloc) {
                    av[u] = Vt[(ll)g * p.ldv + r_lo + rr]; bv[u] = bload(rr, g);
                    if (second) gv[u] = Vl[(ll)g * p.ldv + r_lo + rr];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) dmma_8x8x4(c0, c1, av[u], bv[u]);
            if (second) {
#pragma unroll
                for (int u = 0; u < U; ++u) dmma_8x8x4(d0, d1, av[u], gv[u]);
            }
        }
        red[warp * 128 + g * 8 + 2 * t4] = c0; red[warp * 128 + g * 8 + 2 * t4 + 1] = c1;
        red[warp * 128 + 64 + g * 8 + 2 * t4] = d0; red[warp * 128 + 64 + g * 8 + 2 * t4 + 1] = d1;
        __syncthreads();
        for (int e = tid; e < s * 64; e += NT) {
            const int tt = e >> 6, idx = e & 63;
            double x = 0.0, y = 0.0;
            for (int wq = tt; wq < NW; wq += s) { x += red[wq * 128 + idx]; y += red[wq * 128 + 64 + idx]; }
            part[e] = x; part[512 + e] = y;
        }
        __syncthreads();
    };

    // 1. load the row slice of the sub-panel
    for (int c = 0; c < QR_W; ++c)
        for (int r = tid; r < p.slice; r += NT)
            Cs[c * p.slice + r] = (c < w && r < nloc) ? M[(ll)(p.i0 + c) * p.ld + p.j0 + r_lo + r] : 0.0;
    __syncthreads();

    // 2. block reflector of the earlier sub-panels; completes column block s-1 of the panel's T on the way
    if (s > 0) {
        local_vprev_t([&](int r, int c) { return Cs[c * p.slice + r]; }, s > 1);
        // one reduction for W (s*64 values) and G (s*64 values at offset 512): gather them contiguously
        grid_reduce(1024, Wa);                         // Wa[0..511] = W, Ga (= Wa + 1024)? no: see below
    }
    // NOTE: grid_reduce(1024, Wa) wrote 1024 values starting at Wa: Wa[0..511] = W and W2[0..511] = G (W2 follows Wa in
    // shared memory); move G to Ga before W2 is reused.
    if (s > 0) {
        for (int e = tid; e < 512; e += NT) Ga[e] = W2[e];
        __syncthreads();
        if (s > 1) {
            const int P1 = P - 8;
            for (int e = tid; e < P1 * 8; e += NT) {     // W2 = G T22
                const int i = e >> 3, c = e & 7;
                double x = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l) x += Ga[i * 8 + l] * Tp[(P1 + l) * 64 + P1 + c];
                W2[e] = x;
            }
            __syncthreads();
            // T12 = -T11 W2: every CTA needs it (kept in Ga), CTA 0 stores it
            for (int e = tid; e < P1 * 8; e += NT) {
                const int i = e >> 3, c = e & 7;
                double x = 0.0;
                for (int l = i; l < P1; ++l) x += Tp[i * 64 + l] * W2[l * 8 + c];
                Ga[e] = -x;
            }
            __syncthreads();
            if (cta == 0) for (int e = tid; e < P1 * 8; e += NT) Tp[(e >> 3) * 64 + P1 + (e & 7)] = Ga[e];
        }
        // W2 = T_prev^T W with T_prev = [T11 T12; 0 T22]: columns < P1 from global (earlier launches), block column s-1 from Ga / global diag
        for (int e = tid; e < P * 8; e += NT) {
            const int i = e >> 3, c = e & 7;
            const int P1 = P - 8;
            double x = 0.0;
            for (int l = 0; l <= i; ++l) {
                double tli;
                if (s > 1 && i >= P1 && l < P1) tli = Ga[l * 8 + (i - P1)];       // freshly computed T12 (not yet visible from global)
                else tli = Tp[l * 64 + i];
                x += tli * Wa[l * 8 + c];
            }
            W2[e] = x;
        }
        __syncthreads();
        // 3. update the local rows
        for (int r = tid; r < nloc; r += NT) {
            double c8[QR_W];
#pragma unroll
            for (int c = 0; c < QR_W; ++c) c8[c] = Cs[c * p.slice + r];
            for (int i0 = 0; i0 < P; i0 += 8) {
                double vv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = Vj0[(ll)(i0 + i) * p.ldv + r_lo + r];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < QR_W; ++c) c8[c] -= vv[i] * W2[(i0 + i) * 8 + c];
            }
#pragma unroll
            for (int c = 0; c < QR_W; ++c) Cs[c * p.slice + r] = c8[c];
        }
    }

    // 4. Householder on the 8 columns; rows are thread-owned, one grid reduction per column
    double pt[QR_W];
    auto dots_for = [&](int j) {
#pragma unroll
        for (int c = 0; c < QR_W; ++c) pt[c] = 0.0;
        for (int r = tid; r < nloc; r += NT) {
            if (r_lo + r <= pr0 + j) continue;
            const double x = Cs[j * p.slice + r];
            pt[0] += x * x;
#pragma unroll
            for (int c = 1; c < QR_W; ++c)
                if (j + c < w) pt[c] += x * Cs[(j + c) * p.slice + r];
        }
    };
    dots_for(0);
    for (int j = 0; j < w; ++j) {
        const int pr = pr0 + j;                         // pivot row relative to j0
#pragma unroll
        for (int c = 0; c < QR_W; ++c) {
            double x = pt[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            pt[c] = x;
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < QR_W; ++c) red[warp * QR_W + c] = pt[c];
        }
        __syncthreads();
        if (tid < 16) part[tid] = 0.0;
        __syncthreads();
        if (tid < QR_W) { double x = 0.0; for (int wq = 0; wq < NW; ++wq) x += red[wq * QR_W + tid]; part[tid] = x; }
        // the owner of the pivot row adds it to the reduction (others contribute zeros): part[8 + c] = C[pr][j + c]
        if (pr >= r_lo && pr < r_lo + nloc && tid < QR_W) part[8 + tid] = (j + tid < w) ? Cs[(j + tid) * p.slice + (pr - r_lo)] : 0.0;
        __syncthreads();
        grid_reduce(16, sc);                            // sc[0..7] dots, sc[8..15] pivot row
        const double xnorm2 = sc[0], alpha = sc[8];
        double tj = 0.0, beta = alpha, inv = 0.0;
        if (xnorm2 > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tj = (beta - alpha) / beta;
            inv = 1.0 / (alpha - beta);
        }
        double f[QR_W];
#pragma unroll
        for (int c = 1; c < QR_W; ++c) f[c] = (j + c < w) ? tj * (sc[8 + c] + sc[c] * inv) : 0.0;
        if (cta == 0 && tid == 0) tau[p.i0 + j] = tj;
        if (tid == 0) sc[16 + j] = tj;
        for (int r = tid; r < nloc; r += NT) {
            const int rg = r_lo + r;
            if (rg < pr) continue;
            if (rg == pr) {
                if (tj != 0.0) Cs[j * p.slice + r] = beta;
#pragma unroll
                for (int c = 1; c < QR_W; ++c)
                    if (j + c < w) Cs[(j + c) * p.slice + r] -= f[c];
            } else {
                const double v = (tj != 0.0) ? Cs[j * p.slice + r] * inv : 0.0;
                Cs[j * p.slice + r] = v;
#pragma unroll
                for (int c = 1; c < QR_W; ++c)
                    if (j + c < w) Cs[(j + c) * p.slice + r] -= f[c] * v;
            }
        }
        if (j + 1 < w) dots_for(j + 1);
    }
    __syncthreads();

    // 5. T of the sub-panel (Gram of the new reflectors over all rows), extension of the panel's T by the last sub-panel
    auto vclean = [&](int r, int i) -> double {          // r local, reflector i of this sub-panel
        if (i >= w) return 0.0;
        const int pv = pr0 + i, rg = r_lo + r;
        if (rg < pv) return 0.0;
        if (rg == pv) return 1.0;
        return Cs[i * p.slice + r];
    };
    {
        double c0 = 0.0, c1 = 0.0;
        constexpr int U = 4;
        for (int r = 4 * warp; r < nloc; r += 4 * NW * U) {
            double av[U], bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int rr = r + u * 4 * NW + t4; av[u] = rr < nloc ? vclean(rr, g) : 0.0; bv[u] = av[u]; }
#pragma unroll
            for (int u = 0; u < U; ++u) dmma_8x8x4(c0, c1, av[u], bv[u]);
        }
        __syncthreads();
        red[warp * 64 + g * 8 + 2 * t4] = c0; red[warp * 64 + g * 8 + 2 * t4 + 1] = c1;
        __syncthreads();
        if (tid < 64) { double x = 0.0; for (int wq = 0; wq < NW; ++wq) x += red[wq * 64 + tid]; part[tid] = x; }
        __syncthreads();
        grid_reduce(64, Wk);
    }
    if (tid == 0) {
        for (int i = 0; i < 64; ++i) Tl[i] = 0.0;
        for (int i = 0; i < w; ++i) {
            const double ti = sc[16 + i];
            Tl[i * 8 + i] = ti;
            for (int r = 0; r < i; ++r) {
                double x = 0.0;
                for (int c = r; c < i; ++c) x += Tl[r * 8 + c] * Wk[c * 8 + i];
                Tl[r * 8 + i] = -ti * x;
            }
        }
    }
    __syncthreads();
    if (s > 0 && p.last) {
        local_vprev_t(vclean, false);
        grid_reduce(512, Wa);                            // Wa = G (P x 8)
        for (int e = tid; e < P * 8; e += NT) {
            const int i = e >> 3, c = e & 7;
            double x = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) x += Wa[i * 8 + l] * Tl[l * 8 + c];
            W2[e] = x;
        }
        __syncthreads();
        if (cta == 0) {
            const int P1 = P - 8;
            for (int e = tid; e < P * 8; e += NT) {
                const int i = e >> 3, c = e & 7;
                double x = 0.0;
                for (int l = i; l < P; ++l) {
                    double til;
                    if (s > 1 && l >= P1 && i < P1) til = Ga[i * 8 + (l - P1)];    // block column s-1 written in this launch
                    else til = Tp[i * 64 + l];
                    x += til * W2[l * 8 + c];
                }
                Tp[i * 64 + P + c] = -x;
            }
        }
    }
    if (cta == 0) {
        for (int e = tid; e < 8 * 64; e += NT) {
            const int i = e >> 6, l = e & 63;
            Tp[(P + i) * 64 + l] = (l >= P && l < P + 8) ? Tl[i * 8 + (l - P)] : 0.0;
        }
    }
    // 6. write back the local rows
    for (int c = 0; c < w; ++c) {
        const int prc = pr0 + c;
        for (int r = tid; r < nloc; r += NT) {
            const int rg = r_lo + r;
            const double x = Cs[c * p.slice + r];
            if (rg <= prc) M[(ll)(p.i0 + c) * p.ld + p.j0 + rg] = x;
            V[(ll)(p.i0 + c) * p.ldv + p.j0 + rg] = (rg < prc) ? 0.0 : (rg == prc ? 1.0 : x);
        }
    }
    if constexpr (CL) cg::this_cluster().sync();      // no CTA may exit while a peer can still read its shared memory
}

static size_t qr_leaf3_smem(int slice, int nt) { return ((size_t)QR_W * slice + (nt / 32) * 128 + 512 * 3 + 64 * 2 + 32 + 2048 + 8) * sizeof(double); }
// rows per CTA that fit the shared-memory budget of the multi-CTA leaf
static int qr_leaf3_max_slice(int nt) {
    const size_t fixed = qr_leaf3_smem(0, nt);
    return (int)((((size_t)225 * 1024 - fixed) / (QR_W * sizeof(double))) / 4 * 4);    // 2592 rows at 1024 threads: a 5152-row panel (rank cap 8) splits in two
}
