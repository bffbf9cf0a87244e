// Batched blocked Householder QR building blocks (replaces np.linalg.qr = dgeqrf + dorgqr at
// tensormethods.py:94).  The a x b matrix of an instance lives column-major in global memory (L2);
// one CTA per instance factorises a sub-panel of <= 8 columns held in shared memory (left-looking
// inside the outer panel), the outer-panel compact-WY factor T is built from a Gram matrix, and the
// trailing matrix is updated by the DMMA GEMM kernel.  Reflectors are kept (clean V, tau, T): the
// orthogonal factor is never formed - the truncation sweep applies the stored reflectors to its short
// and wide carry instead, which halves the flops of the LAPACK route the reference takes.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

#define QR_W 8          // sub-panel width (shared-memory leaf)
#define QR_THREADS_SMALL 256
#define QR_THREADS_BIG 1024

struct LeafParams {
    double* M; ll m_stride; int ld;        // working matrix (col-major a x b), R ends up in its upper triangle
    double* V; ll v_stride; int ldv;       // clean reflectors (unit diagonal, zeros above), col-major a x kmax
    double* tau; ll tau_stride;            // [kmax]
    double* Ts; ll ts_stride;              // per sub-panel QR_W x QR_W T factors, sub-panel sp at Ts + sp*QR_W*QR_W
    int a;                                 // rows
    int j0;                                // first column of the enclosing outer panel
    int i0;                                // first column of this sub-panel
    int w;                                 // its width (<= ws)
    int ws;                                // sub-panel stride of this matrix (power of two <= QR_W)
};

// G[i][c] = sum_r X[r][i] * Y[r][c] over r in [0, rows): both operands 8 columns wide (zero padded),
// computed with DMMA (k = 4 rows per instruction), result (64 doubles) left in `out` (shared).
// xload(r, col) / yload(r, col) return the element or 0.
template <int NT, typename XL, typename YL>
__device__ __forceinline__ void gram8(int rows, XL xload, YL yload, double* red, double* out, int tid) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    constexpr int NW = NT / 32, U = 4;
    double c0 = 0.0, c1 = 0.0;
    for (int r = warp * 4; r < rows; r += 4 * NW * U) {
        double av[U], bv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {                     // independent loads first: the loop is latency bound
            const int rr = r + u * 4 * NW + t4;
            av[u] = 0.0; bv[u] = 0.0;
            if (rr < rows) { av[u] = xload(rr, g); bv[u] = yload(rr, g); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) dmma_8x8x4(c0, c1, av[u], bv[u]);
    }
    __syncthreads();
    red[warp * 64 + g * 8 + 2 * t4] = c0;
    red[warp * 64 + g * 8 + 2 * t4 + 1] = c1;
    __syncthreads();
    if (tid < 64) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * 64 + tid];
        out[tid] = s;
    }
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT) qr_leaf_kernel(const LeafParams p) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int z = blockIdx.x;
    const int rows = p.a - p.j0;            // rows j0 .. a-1 take part
    const int w = p.w;
    double* Cs = sm;                        // [w][rows] column-major
    double* red = Cs + (size_t)p.ws * rows; // (NT/32) warps * 64
    double* Wk = red + (NT / 32) * 64;      // 64
    double* Wk2 = Wk + 64;                  // 64
    double* Tl = Wk2 + 64;                  // 64: T of this sub-panel
    double* sc = Tl + 64;                   // scalars: [0] tau_j, [1] 1/(alpha-beta), [8+c] update coefficients, [16+j] taus
    double* M = p.M + (ll)z * p.m_stride;
    double* V = p.V + (ll)z * p.v_stride;
    double* tau = p.tau + (ll)z * p.tau_stride;
    double* Ts = p.Ts + (ll)z * p.ts_stride;

    // 1. load the sub-panel rows j0..a (loads batched eight deep: one CTA per SM is latency bound)
    {
        const int total = w * rows;
        for (int e0 = tid; e0 < total; e0 += NT * 8) {
            double tmp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * NT;
                tmp[u] = (e < total) ? M[(ll)(p.i0 + e / rows) * p.ld + p.j0 + e % rows] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * NT;
                if (e < total) Cs[e] = tmp[u];
            }
        }
    }
    __syncthreads();

    // 2. left-looking: apply the block reflectors of the earlier sub-panels of this outer panel
    for (int ct = p.j0; ct < p.i0; ct += p.ws) {
        const int off = ct - p.j0;                       // V_t is zero above row ct
        const double* Vt = V + (ll)ct * p.ldv + p.j0;    // element (r, i) at Vt[i*ldv + r], r relative to j0
        const double* Tt = Ts + (ll)(ct / p.ws) * QR_W * QR_W;
        gram8<NT>(rows - off,
              [&](int r, int i) { return i < p.ws ? Vt[(ll)i * p.ldv + off + r] : 0.0; },
              [&](int r, int c) { return c < w ? Cs[c * rows + off + r] : 0.0; }, red, Wk, tid);
        if (tid < 64) {                                  // Wk2 = T^T Wk
            const int i = tid >> 3, c = tid & 7;
            double s = 0.0;
            for (int l = 0; l <= i; ++l) s += Tt[l * QR_W + i] * Wk[l * 8 + c];
            Wk2[tid] = s;
        }
        __syncthreads();
        for (int r = tid + off; r < rows; r += NT) {
            double vr[QR_W];
#pragma unroll
            for (int i = 0; i < QR_W; ++i) vr[i] = (i < p.ws) ? Vt[(ll)i * p.ldv + r] : 0.0;
            for (int c = 0; c < w; ++c) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < QR_W; ++i) s += vr[i] * Wk2[i * 8 + c];
                Cs[c * rows + r] -= s;
            }
        }
        __syncthreads();
    }

    // 3. unblocked Householder on the w columns (dlarfg / dlarf semantics).  Per column: one pass of partial
    //    dot products, a shuffle + shared-memory reduction folded by warp 0 - which alone does the FP64
    //    sqrt / divisions and fixes up the pivot row - then one rank-1 update pass.
    for (int j = 0; j < w; ++j) {
        const int pr = p.i0 + j - p.j0;                  // pivot row, relative to j0
        double part[QR_W];
#pragma unroll
        for (int c = 0; c < QR_W; ++c) part[c] = 0.0;
        for (int r = pr + 1 + tid; r < rows; r += NT) {
            const double x = Cs[j * rows + r];
            part[0] += x * x;
#pragma unroll
            for (int c = 1; c < QR_W; ++c)
                if (j + c < w) part[c] += x * Cs[(j + c) * rows + r];
        }
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int c = 0; c < QR_W; ++c) {
            double x = part[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            part[c] = x;
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < QR_W; ++c) red[warp * QR_W + c] = part[c];
        }
        __syncthreads();
        if (warp == 0) {
            double d = 0.0;                              // lane c < 8 folds value c over the warps
            if (lane < QR_W)
                for (int wq = 0; wq < NT / 32; ++wq) d += red[wq * QR_W + lane];
            const double xnorm2 = __shfl_sync(0xffffffffu, d, 0);
            const double alpha = Cs[j * rows + pr];
            double tj = 0.0, beta = alpha, inv = 0.0;
            if (xnorm2 > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
                tj = (beta - alpha) / beta;
                inv = 1.0 / (alpha - beta);
            }
            if (lane >= 1 && lane < QR_W) {
                double f = 0.0;
                if (j + lane < w) {
                    f = tj * (Cs[(j + lane) * rows + pr] + d * inv);
                    Cs[(j + lane) * rows + pr] -= f;     // pivot row of the later columns
                }
                sc[8 + lane] = f;
            }
            if (lane == 0) {
                if (tj != 0.0) Cs[j * rows + pr] = beta;
                sc[0] = tj; sc[1] = inv; sc[16 + j] = tj;
                tau[p.i0 + j] = tj;
            }
        }
        __syncthreads();
        const double tj = sc[0], inv = sc[1];
        double f[QR_W];
#pragma unroll
        for (int c = 1; c < QR_W; ++c) f[c] = sc[8 + c];
        if (tj != 0.0) {
            for (int r = pr + 1 + tid; r < rows; r += NT) {
                const double v = Cs[j * rows + r] * inv;
                Cs[j * rows + r] = v;
#pragma unroll
                for (int c = 1; c < QR_W; ++c)
                    if (j + c < w) Cs[(j + c) * rows + r] -= f[c] * v;
            }
        } else {
            for (int r = pr + 1 + tid; r < rows; r += NT) Cs[j * rows + r] = 0.0;   // x == 0 anyway
        }
        __syncthreads();
    }

    // 4. T of this sub-panel from the Gram matrix of its reflectors
    const int pr0 = p.i0 - p.j0;
    auto vclean = [&](int r, int i) -> double {          // r relative to pr0
        if (i >= w) return 0.0;
        if (r < i) return 0.0;
        if (r == i) return 1.0;
        return Cs[i * rows + pr0 + r];
    };
    gram8<NT>(rows - pr0, vclean, vclean, red, Wk, tid);
    if (tid == 0) {
        for (int i = 0; i < QR_W * QR_W; ++i) Tl[i] = 0.0;
        for (int i = 0; i < w; ++i) {
            const double ti = sc[16 + i];
            Tl[i * QR_W + i] = ti;
            for (int r = 0; r < i; ++r) {
                double s = 0.0;
                for (int c = r; c < i; ++c) s += Tl[r * QR_W + c] * Wk[c * 8 + i];
                Tl[r * QR_W + i] = -ti * s;
            }
        }
    }
    __syncthreads();
    if (tid < 64) Ts[(ll)(p.i0 / p.ws) * QR_W * QR_W + tid] = Tl[tid];

    // 5. write back: R part into M (rows <= pivot), clean V into the reflector store
    for (int c = 0; c < w; ++c) {
        const int prc = pr0 + c;
        for (int r = tid; r < rows; r += NT) {
            const double x = Cs[c * rows + r];
            if (r <= prc) M[(ll)(p.i0 + c) * p.ld + p.j0 + r] = x;
            V[(ll)(p.i0 + c) * p.ldv + p.j0 + r] = (r < prc) ? 0.0 : (r == prc ? 1.0 : x);
        }
    }
}

static size_t qr_leaf_smem(int rows, int ws, int nt) { return ((size_t)ws * rows + (nt / 32) * 64 + 64 * 3 + 32) * sizeof(double); }

// T of an outer panel (jb x jb, upper triangular, row-major ld = nb) from G = V^T V and tau (dlarft, forward,
// columnwise):  T[i][i] = tau_i,  T[0:i, i] = -tau_i * T[0:i, 0:i] * G[0:i, i].  Row r of T depends only on
// its own earlier entries, T[r][i] = -tau_i * sum_{c=r}^{i-1} T[r][c] G[c][i], so one thread per row runs the
// whole recurrence without a barrier.
__global__ void larft_kernel(const double* __restrict__ G, ll g_stride, int nsplit, const double* __restrict__ tau, ll tau_stride,
                             int col0, double* __restrict__ T, ll t_stride, int jb, int nb) {
    extern __shared__ double sm[];
    double* Ts = sm;                 // nb*(nb+1): row r at Ts + r*(nb+1) (padded: rows are walked by different threads)
    double* Gs = Ts + nb * (nb + 1); // nb*nb, transposed: Gs[i*nb + c] = G[c][i] so a thread walks its column contiguously
    double* tv_s = Gs + nb * nb;     // nb
    const int z = blockIdx.x, tid = threadIdx.x;
    const double* g = G + (ll)z * g_stride;          // nsplit partial Gram matrices, nb*nb each
    const double* tv = tau + (ll)z * tau_stride + col0;
    for (int i = tid; i < nb * nb; i += blockDim.x) {
        double s = 0.0;
        if (i / nb < jb && i % nb < jb)
            for (int sp = 0; sp < nsplit; ++sp) s += g[(ll)sp * nb * nb + i];
        Gs[(i % nb) * nb + i / nb] = s;
    }
    for (int i = tid; i < nb * (nb + 1); i += blockDim.x) Ts[i] = 0.0;
    for (int i = tid; i < nb; i += blockDim.x) tv_s[i] = i < jb ? tv[i] : 0.0;
    __syncthreads();
    if (tid < jb) {
        double* row = Ts + tid * (nb + 1);
        row[tid] = tv_s[tid];
        for (int i = tid + 1; i < jb; ++i) {
            const double* gc = Gs + i * nb;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four chains: the sum is latency bound
            int c = tid;
            for (; c + 3 < i; c += 4) { s0 += row[c] * gc[c]; s1 += row[c + 1] * gc[c + 1]; s2 += row[c + 2] * gc[c + 2]; s3 += row[c + 3] * gc[c + 3]; }
            for (; c < i; ++c) s0 += row[c] * gc[c];
            row[i] = -tv_s[i] * ((s0 + s1) + (s2 + s3));
        }
    }
    __syncthreads();
    double* t = T + (ll)z * t_stride;
    for (int i = tid; i < nb * nb; i += blockDim.x) t[i] = Ts[(i / nb) * (nb + 1) + i % nb];
}

// R (q x b, column-major ld = q, zero below the diagonal) out of the factorised working matrix.
__global__ void extract_r_kernel(const double* __restrict__ M, ll m_stride, int ld, double* __restrict__ R, ll r_stride,
                                 int q, int b) {
    const int z = blockIdx.y;
    const ll total = (ll)q * b;
    const double* m = M + (ll)z * m_stride;
    double* r = R + (ll)z * r_stride;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (ll)gridDim.x * blockDim.x) {
        const int j = (int)(e % q), c = (int)(e / q);
        r[e] = (j <= c) ? m[(ll)c * ld + j] : 0.0;
    }
}
