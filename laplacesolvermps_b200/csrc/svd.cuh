// Batched one-sided (Hestenes) Jacobi SVD with the reference's rank rule, one CTA per instance
// (replaces np.linalg.svd = dgesdd at tensormethods.py:121 and _get_required_ranks_for_accuracy :5-12).
//
// The unfolding A (mr x qc, row-major) of the carry core is orthogonalised along its SHORT side:
//   mr <= qc : rows are rotated,   J^T A = diag(S) V^T  -> U = J,           S V^T = rotated rows
//   mr >  qc : columns are rotated, A W  = U diag(S)    -> U = cols / S,    S V^T = S W^T
// so the "push" factor diag(S) V^T of the reference falls out without a second product.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

#define SVD_THREADS 256
#define SVD_THREADS_MAX 1024
#define SVD_MAX_VEC 2048

struct SvdParams {
    double* A; ll a_stride;           // per group-local instance, row-major mr x qc (ld = qc)
    int qc;
    int n_mode;
    int* ranks; int rstride; int kcol; // mr = ranks[inst*rstride + kcol] * n_mode; writes ranks[.. + kcol + 1]
    const int* idx;
    double* JT; ll jt_stride;          // accumulated rotations, row i = column i of J (or W)
    double* U; ll u_slot;              // output core (instance indexed), row-major mr x s_new
    double* G; ll g_stride;            // output carry, row-major s_new x qc (ld = qc)
    int cap;
    double eps2;
    int* status;                       // instance indexed, may be null
    int max_sweeps;
    double* sigma; ll sigma_stride;    // optional singular values (group-local), may be null
    int smem_vecs;                     // 1: the launch carries dynamic shared memory for the vectors and J (see jacobi_svd_smem)
};

// dynamic shared memory of jacobi_svd_kernel for at most nvec_b vectors of at most len_b elements; 0 = does not fit (global-memory path)
static size_t jacobi_svd_smem(int nvec_b, int len_b) {
    const size_t need = ((size_t)nvec_b * (len_b + 1) + (size_t)nvec_b * nvec_b) * sizeof(double);
    return need <= (size_t)190 * 1024 ? need : 0;
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Launched with 32 threads per pair of vectors (256 .. 1024): a round of the tournament then costs ONE pair per warp instead of
// ceil(pairs / 8) in sequence (40 vectors of config C1: 20 pairs).
__global__ void __launch_bounds__(SVD_THREADS_MAX) jacobi_svd_kernel(const SvdParams p) {
    __shared__ double nrm[SVD_MAX_VEC];
    __shared__ int order[SVD_MAX_VEC];
    __shared__ int s_rot, s_keep;
    const int z = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = blockDim.x, NW = NT >> 5;
    const int inst = p.idx ? p.idx[z] : z;
    const int mr = p.ranks[(ll)inst * p.rstride + p.kcol] * p.n_mode;
    const int qc = p.qc;
    const bool wide = (mr <= qc);
    const int nvec = wide ? mr : qc, len = wide ? qc : mr;
    double* Ag = p.A + (ll)z * p.a_stride;
    double* JTg = p.JT + (ll)z * p.jt_stride;
    // The vectors (and the accumulated rotations) live in SHARED memory whenever they fit (p.smem_vecs: the launch reserved
    // nvec_bound * (len_bound + 1) + nvec_bound^2 doubles): every round of the tournament is three dependent passes over a pair of
    // vectors (dot products, rotation, rotation of J) behind one barrier, i.e. a latency chain - out of L2 a round costs 3-8 us,
    // out of shared memory about 1 us.  Vectors are stored contiguously (the columns of a tall unfolding are gathered), stride len + 1.
    extern __shared__ double svd_sm[];
    const bool in_smem = p.smem_vecs != 0;
    const int lds = len + 1;
    double* A = in_smem ? svd_sm : Ag;
    double* JT = in_smem ? svd_sm + (size_t)nvec * lds : JTg;
    const ll vs = in_smem ? lds : (wide ? qc : 1), es = in_smem ? 1 : (wide ? 1 : qc);
    if (in_smem)
        for (int e = tid; e < nvec * len; e += NT) {
            const int v = e / len, x = e % len;
            A[v * lds + x] = wide ? Ag[(ll)v * qc + x] : Ag[(ll)x * qc + v];
        }
    for (int e = tid; e < nvec * nvec; e += NT) JT[e] = (e / nvec == e % nvec) ? 1.0 : 0.0;
    __syncthreads();

    const int np = (nvec + 1) & ~1;            // players, padded to even
    const double tol = 1e-15;
    int sweeps = 0;
    bool converged = (nvec <= 1);
    while (!converged && sweeps < p.max_sweeps) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int rho = 0; rho < np - 1; ++rho) {
            for (int k = warp; k < np / 2; k += NW) {
                int i, j;
                if (k == 0) { i = np - 1; j = rho; }
                else { i = (rho + k) % (np - 1); j = (rho - k + (np - 1)) % (np - 1); }
                if (i > j) { int t = i; i = j; j = t; }
                if (j >= nvec) continue;
                double* vi = A + i * vs;
                double* vj = A + j * vs;
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int e = lane; e < len; e += 32) {
                    const double x = vi[e * es], y = vj[e * es];
                    al += x * x; be += y * y; ga += x * y;
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                // the serial scalar chain of a round: the convergence test without square roots where the product of the norms is
                // in range, the rotation from two reciprocals and two reciprocal square roots (was: three sqrt, three divisions
                // and two more sqrt for the test)
                const double prod = al * be;
                bool rotate;
                if (prod > 1e-280 && prod < 1e280) rotate = ga * ga > (tol * tol) * prod;
                else { const double lim = sqrt(al) * sqrt(be); rotate = (fabs(ga) > tol * lim) && !(lim < 1e-300); }
                if (!rotate) continue;
                const double zeta = (be - al) * __drcp_rn(2.0 * ga);
                if (zeta != zeta) continue;                       // 0 * inf: equal norms and a denormal overlap - nothing to rotate
                const double w2 = fma(zeta, zeta, 1.0);
                // sqrt(1 + zeta^2); for |zeta| > 1e150 (null vectors shrink by ~1e-16 per sweep: their norms do reach 1e-300) it is |zeta|
                const double w = (w2 < 1e300) ? w2 * rsqrt(w2) : fabs(zeta);
                const double t = copysign(__drcp_rn(fabs(zeta) + w), zeta);
                const double c = rsqrt(fma(t, t, 1.0)), s = c * t;
                for (int e = lane; e < len; e += 32) {
                    const double x = vi[e * es], y = vj[e * es];
                    vi[e * es] = c * x - s * y;
                    vj[e * es] = s * x + c * y;
                }
                double* ji = JT + (ll)i * nvec;
                double* jj = JT + (ll)j * nvec;
                for (int e = lane; e < nvec; e += 32) {
                    const double x = ji[e], y = jj[e];
                    ji[e] = c * x - s * y;
                    jj[e] = s * x + c * y;
                }
                if (lane == 0) atomicAdd(&s_rot, 1);
            }
            __syncthreads();
        }
        ++sweeps;
        converged = (s_rot == 0);
        __syncthreads();
    }

    // squared norms, descending order (stable), rank rule
    for (int i = warp; i < nvec; i += NW) {
        const double* vi = A + i * vs;
        double al = 0.0;
        for (int e = lane; e < len; e += 32) { const double x = vi[e * es]; al += x * x; }
        al = warp_sum(al);
        if (lane == 0) nrm[i] = al;
    }
    __syncthreads();
    // (a NaN norm - a poisoned input - must still yield a permutation: NaN sorts last, ties by index; the instance is flagged below)
    for (int i = tid; i < nvec; i += NT) {
        const double ni = nrm[i];
        const bool ni_nan = !(ni == ni);
        int r = 0;
        for (int j = 0; j < nvec; ++j) {
            const double nj = nrm[j];
            const bool nj_nan = !(nj == nj);
            r += ni_nan ? (!nj_nan || j < i) : (!nj_nan && ((nj > ni) || (nj == ni && j < i)));
        }
        order[r] = i;
    }
    __syncthreads();
    if (tid == 0) {
        const double s0 = nrm[order[0]];
        int removed = 0;
        for (int i = 0; i < nvec; ++i) removed += ((nrm[i] / s0) < p.eps2) ? 1 : 0;   // nan -> kept (reference)
        int keep = nvec - removed;
        if (keep > p.cap) keep = p.cap;
        if (keep < 0) keep = 0;
        s_keep = keep;
        p.ranks[(ll)inst * p.rstride + p.kcol + 1] = keep;
        if (p.status && (!converged || !(s0 == s0))) atomicAdd(&p.status[inst], 1);
    }
    __syncthreads();
    const int sn = s_keep;
    double* U = p.U + (ll)inst * p.u_slot;
    double* G = p.G + (ll)z * p.g_stride;
    if (p.sigma) for (int c = tid; c < nvec; c += NT) p.sigma[(ll)z * p.sigma_stride + c] = sqrt(nrm[order[c]]);
    if (wide) {
        for (int e = tid; e < sn * qc; e += NT) { const int c = e / qc, x = e % qc; G[e] = A[(ll)order[c] * vs + x * es]; }
        for (int e = tid; e < mr * sn; e += NT) { const int r = e / sn, c = e % sn; U[e] = JT[(ll)order[c] * nvec + r]; }
    } else {
        for (int e = tid; e < sn * qc; e += NT) {
            const int c = e / qc, x = e % qc;
            G[e] = sqrt(nrm[order[c]]) * JT[(ll)order[c] * nvec + x];
        }
        for (int e = tid; e < mr * sn; e += NT) {
            const int r = e / sn, c = e % sn;
            const double sg = sqrt(nrm[order[c]]);
            U[e] = sg > 0.0 ? A[(ll)order[c] * vs + r * es] / sg : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Large SVDs (operator rounding: up to ~2000 vectors): the same one-sided Jacobi, spread over the GPU.  One launch per
// round of the round-robin tournament, one warp per pair, rows of X (nvec x len, row-major) rotated in global memory
// (L2 resident: 1152 x 1152 doubles = 10 MB); convergence is read back once per sweep.
// ------------------------------------------------------------------------------------------------------------------
struct BigJacobiParams {
    double* X; ll x_stride; int ld; int len;     // rows to orthogonalise
    double* JT; ll jt_stride;                    // accumulated rotations (nvec x nvec), row i = column i of J
    int nvec; int np; int rho;
    int* rot_count;                              // [G] rotations done in this sweep
    double tol;
};

__device__ __forceinline__ void big_jacobi_pair(const BigJacobiParams& p, int z, int k, int rho, int lane) {
    if (k >= p.np / 2) return;
    int i, j;
    if (k == 0) { i = p.np - 1; j = rho; }
    else { i = (rho + k) % (p.np - 1); j = (rho - k + (p.np - 1)) % (p.np - 1); }
    if (i > j) { const int t = i; i = j; j = t; }
    if (j >= p.nvec) return;
    double* X = p.X + (ll)z * p.x_stride;
    double* vi = X + (ll)i * p.ld;
    double* vj = X + (ll)j * p.ld;
    // every pass keeps eight loads per vector in flight: the loops are bound by L2 latency, not bandwidth.  Vectors of up
    // to 1280 elements stay in registers between the dot products and the rotation (one pass over global memory less).
    double al = 0.0, be = 0.0, ga = 0.0;
    if (p.len <= 1280) {
        double x[40], y[40];
#pragma unroll
        for (int u = 0; u < 40; ++u) { const int e = lane + 32 * u; x[u] = e < p.len ? vi[e] : 0.0; y[u] = e < p.len ? vj[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 40; ++u) { al += x[u] * x[u]; be += y[u] * y[u]; ga += x[u] * y[u]; }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        const double lim = sqrt(al) * sqrt(be);
        if (!(fabs(ga) > p.tol * lim) || lim < 1e-300) return;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int u = 0; u < 40; ++u) { const int e = lane + 32 * u; if (e < p.len) { vi[e] = c * x[u] - s * y[u]; vj[e] = s * x[u] + c * y[u]; } }
        double* JTr = p.JT + (ll)z * p.jt_stride;
        double* ji = JTr + (ll)i * p.nvec;
        double* jj = JTr + (ll)j * p.nvec;
        for (int e0 = lane; e0 < p.nvec; e0 += 256) {
            double a8[8], b8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; a8[u] = e < p.nvec ? ji[e] : 0.0; b8[u] = e < p.nvec ? jj[e] : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; if (e < p.nvec) { ji[e] = c * a8[u] - s * b8[u]; jj[e] = s * a8[u] + c * b8[u]; } }
        }
        if (lane == 0) atomicAdd(&p.rot_count[z], 1);
        return;
    }
    for (int e0 = lane; e0 < p.len; e0 += 256) {
        double x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; x[u] = e < p.len ? vi[e] : 0.0; y[u] = e < p.len ? vj[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { al += x[u] * x[u]; be += y[u] * y[u]; ga += x[u] * y[u]; }
    }
    al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
    const double lim = sqrt(al) * sqrt(be);
    if (!(fabs(ga) > p.tol * lim) || lim < 1e-300) return;
    const double zeta = (be - al) / (2.0 * ga);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
    for (int e0 = lane; e0 < p.len; e0 += 256) {
        double x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; x[u] = e < p.len ? vi[e] : 0.0; y[u] = e < p.len ? vj[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; if (e < p.len) { vi[e] = c * x[u] - s * y[u]; vj[e] = s * x[u] + c * y[u]; } }
    }
    double* JT = p.JT + (ll)z * p.jt_stride;
    double* ji = JT + (ll)i * p.nvec;
    double* jj = JT + (ll)j * p.nvec;
    for (int e0 = lane; e0 < p.nvec; e0 += 256) {
        double x[8], y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; x[u] = e < p.nvec ? ji[e] : 0.0; y[u] = e < p.nvec ? jj[e] : 0.0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; if (e < p.nvec) { ji[e] = c * x[u] - s * y[u]; jj[e] = s * x[u] + c * y[u]; } }
    }
    if (lane == 0) atomicAdd(&p.rot_count[z], 1);
}

// one round per launch (fallback when the grid cannot be co-resident)
__global__ void __launch_bounds__(256) big_jacobi_round_kernel(const BigJacobiParams p) {
    big_jacobi_pair(p, blockIdx.y, blockIdx.x * 8 + (threadIdx.x >> 5), p.rho, threadIdx.x & 31);
}

// a whole sweep (np - 1 rounds) in one cooperative launch, grid-wide barrier between rounds
__global__ void __launch_bounds__(256) big_jacobi_sweep_kernel(const BigJacobiParams p) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    for (int rho = 0; rho < p.np - 1; ++rho) {
        big_jacobi_pair(p, blockIdx.y, blockIdx.x * 8 + (threadIdx.x >> 5), rho, threadIdx.x & 31);
        grid.sync();
    }
}

__global__ void big_jacobi_init_kernel(double* JT, ll jt_stride, int nvec) {
    double* j = JT + (ll)blockIdx.y * jt_stride;
    const ll total = (ll)nvec * nvec;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (ll)gridDim.x * blockDim.x)
        j[e] = (e / nvec == e % nvec) ? 1.0 : 0.0;
}

// squared row norms, one warp per row
__global__ void big_jacobi_norms_kernel(const double* X, ll x_stride, int ld, int len, int nvec, double* nrm, ll n_stride) {
    const int z = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nvec) return;
    const double* v = X + (ll)z * x_stride + (ll)i * ld;
    double al = 0.0;
    for (int e = lane; e < len; e += 32) al += v[e] * v[e];
    al = warp_sum(al);
    if (lane == 0) nrm[(ll)z * n_stride + i] = al;
}

// descending order (stable), rank rule; order / sigma to global, rank into the table.  One CTA per instance.
struct BigRankParams {
    const double* nrm; ll n_stride; int nvec;
    int* order; ll o_stride;
    int* ranks; int rstride; int kcol; const int* idx;
    int mr, qc, cap; double eps2;
    int* status; int noconv;      // noconv: the sweep limit was hit (host decision)
};

__global__ void __launch_bounds__(256) big_jacobi_rank_kernel(const BigRankParams p) {
    __shared__ int s_removed;
    const int z = blockIdx.x, tid = threadIdx.x;
    const int inst = p.idx ? p.idx[z] : z;
    const double* nrm = p.nrm + (ll)z * p.n_stride;
    int* order = p.order + (ll)z * p.o_stride;
    if (tid == 0) s_removed = 0;
    __syncthreads();
    for (int i = tid; i < p.nvec; i += 256) {      // NaN sorts last (see jacobi_svd_kernel): the order is always a permutation
        const double ni = nrm[i];
        const bool ni_nan = !(ni == ni);
        int r = 0;
        for (int j = 0; j < p.nvec; ++j) {
            const double nj = nrm[j];
            const bool nj_nan = !(nj == nj);
            r += ni_nan ? (!nj_nan || j < i) : (!nj_nan && ((nj > ni) || (nj == ni && j < i)));
        }
        order[r] = i;
    }
    __syncthreads();
    const double s0 = nrm[order[0]];
    int removed = 0;
    for (int i = tid; i < p.nvec; i += 256) removed += ((nrm[i] / s0) < p.eps2) ? 1 : 0;      // nan -> kept (reference)
    atomicAdd(&s_removed, removed);
    __syncthreads();
    if (tid == 0) {
        int keep = p.nvec - s_removed;
        keep = min(keep, min(p.cap, min(p.mr, p.qc)));
        if (keep < 0) keep = 0;
        p.ranks[(ll)inst * p.rstride + p.kcol + 1] = keep;
        if (p.status && (p.noconv || !(s0 == s0))) atomicAdd(&p.status[inst], 1);
    }
}

// outputs of the wide case (rows of the unfolding were rotated): U[r][c] = JT[order[c]][r], G[c][:] = X[order[c]][:]
// and of the tall case after QR (X = R^T rotated): Ur[:, c] = X[order[c]][:] / sigma_c (column-major qc x s, for the
// reflector application), G[c][:] = sigma_c * JT[order[c]][:]
struct BigOutParams {
    const double* X; ll x_stride; int ld; int len;
    const double* JT; ll jt_stride; int nvec;
    const double* nrm; ll n_stride; const int* order; ll o_stride;
    const int* ranks; int rstride; int kcol; const int* idx;
    double* U; ll u_stride; int u_idx; int mr;       // wide: out core (instance indexed, row-major mr x s); tall: X0 (group-local, column-major a x s)
    double* G; ll g_stride; int qc;
    int tall; int a_rows;                            // tall: leading dimension of X0
};

__global__ void big_jacobi_out_kernel(const BigOutParams p) {
    const int z = blockIdx.y;
    const int inst = p.idx ? p.idx[z] : z;
    const int s = p.ranks[(ll)inst * p.rstride + p.kcol + 1];
    const double* X = p.X + (ll)z * p.x_stride;
    const double* JT = p.JT + (ll)z * p.jt_stride;
    const double* nrm = p.nrm + (ll)z * p.n_stride;
    const int* order = p.order + (ll)z * p.o_stride;
    double* U = p.U + (ll)(p.u_idx ? inst : z) * p.u_stride;
    double* G = p.G + (ll)z * p.g_stride;
    const ll stride = (ll)gridDim.x * blockDim.x, t0 = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (!p.tall) {
        for (ll e = t0; e < (ll)s * p.qc; e += stride) { const int c = (int)(e / p.qc), x = (int)(e % p.qc); G[e] = X[(ll)order[c] * p.ld + x]; }
        for (ll e = t0; e < (ll)p.mr * s; e += stride) { const int r = (int)(e / s), c = (int)(e % s); U[e] = JT[(ll)order[c] * p.nvec + r]; }
    } else {
        for (ll e = t0; e < (ll)s * p.qc; e += stride) {
            const int c = (int)(e / p.qc), x = (int)(e % p.qc);
            G[e] = sqrt(nrm[order[c]]) * JT[(ll)order[c] * p.nvec + x];
        }
        for (ll e = t0; e < (ll)p.a_rows * s; e += stride) {            // X0 = [Ur ; 0], column-major a_rows x s
            const int r = (int)(e % p.a_rows), c = (int)(e / p.a_rows);
            double v = 0.0;
            if (r < p.len) { const double sg = sqrt(nrm[order[c]]); v = sg > 0.0 ? X[(ll)order[c] * p.ld + r] / sg : 0.0; }
            U[e] = v;
        }
    }
}
