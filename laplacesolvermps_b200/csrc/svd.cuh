// Batched one-sided (Hestenes) Jacobi SVD with the reference's rank rule, one CTA per instance
// (replaces np.linalg.svd = dgesdd at tensormethods.py:121 and _get_required_ranks_for_accuracy :5-12).
//
// The unfolding A (mr x qc, row-major) of the carry core is orthogonalised along its SHORT side:
//   mr <= qc : rows are rotated,   J^T A = diag(S) V^T  -> U = J,           S V^T = rotated rows
//   mr >  qc : columns are rotated, A W  = U diag(S)    -> U = cols / S,    S V^T = S W^T
// so the "push" factor diag(S) V^T of the reference falls out without a second product.
#pragma once
#include "common.cuh"

#define SVD_THREADS 256
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
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(SVD_THREADS) jacobi_svd_kernel(const SvdParams p) {
    __shared__ double nrm[SVD_MAX_VEC];
    __shared__ int order[SVD_MAX_VEC];
    __shared__ int s_rot, s_keep;
    const int z = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = SVD_THREADS / 32;
    const int inst = p.idx ? p.idx[z] : z;
    const int mr = p.ranks[(ll)inst * p.rstride + p.kcol] * p.n_mode;
    const int qc = p.qc;
    const bool wide = (mr <= qc);
    const int nvec = wide ? mr : qc, len = wide ? qc : mr;
    const ll vs = wide ? qc : 1, es = wide ? 1 : qc;
    double* A = p.A + (ll)z * p.a_stride;
    double* JT = p.JT + (ll)z * p.jt_stride;

    for (int e = tid; e < nvec * nvec; e += SVD_THREADS) JT[e] = (e / nvec == e % nvec) ? 1.0 : 0.0;
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
                const double lim = sqrt(al) * sqrt(be);
                if (!(fabs(ga) > tol * lim) || lim < 1e-300) continue;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
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
    for (int i = tid; i < nvec; i += SVD_THREADS) {
        const double ni = nrm[i];
        int r = 0;
        for (int j = 0; j < nvec; ++j) { const double nj = nrm[j]; r += (nj > ni) || (nj == ni && j < i); }
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
        if (p.status && !converged) atomicAdd(&p.status[inst], 1);
    }
    __syncthreads();
    const int sn = s_keep;
    double* U = p.U + (ll)inst * p.u_slot;
    double* G = p.G + (ll)z * p.g_stride;
    if (p.sigma) for (int c = tid; c < nvec; c += SVD_THREADS) p.sigma[(ll)z * p.sigma_stride + c] = sqrt(nrm[order[c]]);
    if (wide) {
        for (int e = tid; e < sn * qc; e += SVD_THREADS) { const int c = e / qc, x = e % qc; G[e] = A[(ll)order[c] * qc + x]; }
        for (int e = tid; e < mr * sn; e += SVD_THREADS) { const int r = e / sn, c = e % sn; U[e] = JT[(ll)order[c] * nvec + r]; }
    } else {
        for (int e = tid; e < sn * qc; e += SVD_THREADS) {
            const int c = e / qc, x = e % qc;
            G[e] = sqrt(nrm[order[c]]) * JT[(ll)order[c] * nvec + x];
        }
        for (int e = tid; e < mr * sn; e += SVD_THREADS) {
            const int r = e / sn, c = e % sn;
            const double sg = sqrt(nrm[order[c]]);
            U[e] = sg > 0.0 ? A[(ll)r * qc + order[c]] / sg : 0.0;
        }
    }
}
