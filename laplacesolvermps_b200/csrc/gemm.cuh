// Generic strided, batched FP64 GEMM on the DMMA pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
//   C[m, n] = alpha * alpha_vec[inst] * sum_k A[m, k] * B[k, n] + beta * C[m, n]
//
// Every operand is addressed by two element strides, so one kernel serves every contraction of the
// rounding sweeps (push of the triangular factor through an MPO core, compact-WY trailing updates,
// application of stored reflectors) without transposed copies.  tcgen05 has no FP64 kind; on B200 the
// DMMA and DFMA pipes peak at the same 37.1 TFLOP/s (tools/fp64_peak.cu), DMMA needs 8x fewer issue
// slots, which leaves room for the shared-memory traffic.
#pragma once
#include "common.cuh"

struct GemmDesc {
    int M, N, K;
    const double* A; ll a_m, a_k;
    const double* B; ll b_k, b_n;
    double* C; ll c_m, c_n;
    double alpha, beta;
    const double* alpha_vec;      // optional, indexed by instance id
    int nb[3];                    // batch dims; z = (i0*nb[1] + i1)*nb[2] + i2
    ll a_s[3], b_s[3], c_s[3];    // batch strides (elements)
    const int* idx;               // optional map: group-local i0 -> instance id
    int a_idx, b_idx, c_idx;      // 1: batch stride 0 multiplies the instance id, 0: multiplies i0
    const int* n_dyn;             // optional per-instance N: n_dyn[inst*n_dyn_stride + n_dyn_off]
    int n_dyn_stride, n_dyn_off;
    // optional two-level row addressing of C: row m lives at (m / c_msplit) * c_mhi + (m % c_msplit) * c_m
    int c_msplit; ll c_mhi;
    // optional structural zeros of A (the pushed triangular factor): with j = m % tri_mperiod,
    // A[m, k] == 0 whenever tri_off + (k % tri_kperiod) * tri_div + tri_div - 1 < j.  tri_div == 0: dense.
    int tri_div, tri_kperiod, tri_mperiod, tri_off;   // tri_off: column offset of the block inside the factor
};

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(128) gemm_dmma_kernel(const GemmDesc d, int tiles_m, int tiles_n) {
    constexpr int KC = 16, LDS = KC + 4;                 // +4 doubles: conflict-free fragment reads
    constexpr int WARPS_N = BN / WN;
    static_assert((BM / WM) * (BN / WN) == 4, "4 warps per CTA");
    constexpr int FM = WM / 8, FN = WN / 8;
    constexpr int LA = BM * KC / 128, LB = BN * KC / 128;
    __shared__ double As[2][BM * LDS];
    __shared__ double Bs[2][BN * LDS];

    const int tiles = tiles_m * tiles_n;
    const ll bid = blockIdx.x;
    const int z = (int)(bid / tiles);
    const int tile = (int)(bid % tiles);
    const int tm = tile % tiles_m, tn = tile / tiles_m;
    const int i2 = z % d.nb[2];
    const int i1 = (z / d.nb[2]) % d.nb[1];
    const int i0 = z / (d.nb[2] * d.nb[1]);
    const int inst = d.idx ? d.idx[i0] : i0;
    const int N = d.n_dyn ? d.n_dyn[(ll)inst * d.n_dyn_stride + d.n_dyn_off] : d.N;
    const int M = d.M, K = d.K;
    const int m0 = tm * BM, n0 = tn * BN;
    if (m0 >= M || n0 >= N) return;

    const double* A = d.A + (d.a_idx ? inst : i0) * d.a_s[0] + i1 * d.a_s[1] + i2 * d.a_s[2];
    const double* B = d.B + (d.b_idx ? inst : i0) * d.b_s[0] + i1 * d.b_s[1] + i2 * d.b_s[2];
    double* C = d.C + (d.c_idx ? inst : i0) * d.c_s[0] + i1 * d.c_s[1] + i2 * d.c_s[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    const bool a_kc = (d.a_k == 1), b_kc = (d.b_k == 1);

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    double ra[LA], rb[LB];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int e = tid + 128 * i, m, k;
            if (a_kc) { k = e % KC; m = e / KC; } else { m = e % BM; k = e / BM; }
            int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < K) ? __ldg(A + gm * d.a_m + gk * d.a_k) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int e = tid + 128 * i, n, k;
            if (b_kc) { k = e % KC; n = e / KC; } else { n = e % BN; k = e / BN; }
            int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < N && gk < K) ? __ldg(B + gk * d.b_k + gn * d.b_n) : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int e = tid + 128 * i, m, k;
            if (a_kc) { k = e % KC; m = e / KC; } else { m = e % BM; k = e / BM; }
            As[buf][m * LDS + k] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int e = tid + 128 * i, n, k;
            if (b_kc) { k = e % KC; n = e / KC; } else { n = e % BN; k = e / BN; }
            Bs[buf][n * LDS + k] = rb[i];
        }
    };

    const int nchunks = (K + KC - 1) / KC;
    int k_lo = 0;                                  // k % tri_kperiod below k_lo is structurally zero for this row tile
    if (d.tri_div > 0) {
        const int jlo = m0 % d.tri_mperiod, jhi = min(m0 + BM, M) - 1;
        const bool wraps = (jhi / d.tri_mperiod) != (m0 / d.tri_mperiod);
        k_lo = (wraps || jlo <= d.tri_off) ? 0 : (jlo - d.tri_off) / d.tri_div;
    }
    auto next_active = [&](int c) {
        if (d.tri_div > 0)
            while (c < nchunks && ((c * KC) % d.tri_kperiod) + KC - 1 < k_lo && ((c * KC) % d.tri_kperiod) + KC <= d.tri_kperiod) ++c;
        return c;
    };
    int c = next_active(0);
    if (c < nchunks) { gload(c * KC); sstore(0); }
    __syncthreads();
    int buf = 0;
    while (c < nchunks) {
        const int cn = next_active(c + 1);
        if (cn < nchunks) gload(cn * KC);
        const double* as = As[buf];
        const double* bs = Bs[buf];
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) af[i] = as[(wm0 + i * 8 + g) * LDS + kk + t4];
#pragma unroll
            for (int j = 0; j < FN; ++j) bf[j] = bs[(wn0 + j * 8 + g) * LDS + kk + t4];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (cn < nchunks) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
        c = cn;
    }

    double alpha = d.alpha;
    if (d.alpha_vec) alpha *= d.alpha_vec[inst];
    const double beta = d.beta;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int gm = m0 + wm0 + i * 8 + g;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn0 + j * 8 + 2 * t4 + e;
                if (gn >= N) continue;
                double* p = C + (d.c_msplit ? (gm / d.c_msplit) * d.c_mhi + (gm % d.c_msplit) * d.c_m : gm * d.c_m) + gn * d.c_n;
                double v = alpha * acc[i][j][e];
                if (beta != 0.0) v += beta * (*p);
                *p = v;
            }
        }
    }
}

static inline GemmDesc gemm_desc_zero() {
    GemmDesc d;
    memset(&d, 0, sizeof d);
    d.alpha = 1.0; d.beta = 0.0;
    d.nb[0] = d.nb[1] = d.nb[2] = 1;
    return d;
}

// Launch; dims that are zero make it a no-op (but beta == 0 still requires C to be defined: callers
// never rely on a K == 0 launch to clear C).
static int gemm_launch(qtt_handle_s* h, cudaStream_t st, const GemmDesc& d) {
    if (d.M <= 0 || d.N <= 0 || d.nb[0] <= 0 || d.nb[1] <= 0 || d.nb[2] <= 0) return QTT_OK;
    const ll nbatch = (ll)d.nb[0] * d.nb[1] * d.nb[2];
    auto launch = [&](auto kern, int BM, int BN) -> int {
        int tmn = (int)ceil_div(d.M, BM), tnn = (int)ceil_div(d.N, BN);
        ll blocks = nbatch * tmn * tnn;
        if (blocks > 2147483647LL) { h->err = "gemm grid too large"; return QTT_ERR_UNSUPPORTED; }
        // algorithmic flops actually executed: structurally-zero K chunks that the kernel skips are not counted
        double kexec = (double)d.K * d.M;
        if (d.tri_div > 0) {
            const int KC = 16, nchunks = (d.K + KC - 1) / KC;
            kexec = 0.0;
            for (int tm = 0; tm < tmn; ++tm) {
                const int m0 = tm * BM, mrows = std::min(BM, d.M - m0);
                const int jlo = m0 % d.tri_mperiod, jhi = m0 + mrows - 1;
                const bool wraps = (jhi / d.tri_mperiod) != (m0 / d.tri_mperiod);
                const int k_lo = (wraps || jlo <= d.tri_off) ? 0 : (jlo - d.tri_off) / d.tri_div;
                int active = 0;
                for (int c = 0; c < nchunks; ++c) {
                    const int kp = (c * KC) % d.tri_kperiod;
                    if (!(kp + KC - 1 < k_lo && kp + KC <= d.tri_kperiod)) active += std::min(KC, d.K - c * KC);
                }
                kexec += (double)active * mrows;
            }
        }
        const double fl = 2.0 * kexec * (double)d.N * (double)nbatch;
        h->gemm_flops += fl;
        h->launches += 1;
        size_t slot = 0;
        if (h->prof_on) {
            if (h->prof_used * 2 >= h->prof_ev.size()) prof_flush(h);
            slot = h->prof_used++;
            h->prof_fl[slot] = fl;
            cudaEventRecord(h->prof_ev[2 * slot], st);
        }
        kern<<<(unsigned)blocks, 128, 0, st>>>(d, tmn, tnn);
        if (h->prof_on) cudaEventRecord(h->prof_ev[2 * slot + 1], st);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("gemm launch: ") + cudaGetErrorString(e); return QTT_ERR_CUDA; }
        return QTT_OK;
    };
    if (d.N <= 16) return launch(gemm_dmma_kernel<128, 16, 32, 16>, 128, 16);
    if (d.M <= 32) return launch(gemm_dmma_kernel<32, 64, 16, 32>, 32, 64);
    if (ceil_div(d.N, 56) * 56 < ceil_div(d.N, 64) * 64) return launch(gemm_dmma_kernel<64, 56, 16, 56>, 64, 56);
    return launch(gemm_dmma_kernel<64, 64, 32, 32>, 64, 64);
}
