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
__global__ void __launch_bounds__(128, 3) gemm_dmma_kernel(const GemmDesc d, int tiles_m, int tiles_n) {
    constexpr int KC = 16;
    // Two shared-memory layouts per operand, chosen by which global index is contiguous, so that both the
    // staging stores (lanes along the contiguous index) and the DMMA fragment reads are bank-conflict free:
    //   K-contiguous operand:  S[row * (KC+4) + k]      row stride 20 doubles
    //   row-contiguous operand: S[k * (ROWS+4) + row]   k stride ROWS + 4 doubles (== 4 mod 16)
    constexpr int LDK = KC + 4, LDMA = BM + 4, LDNB = BN + 4;
    constexpr int A_ELEMS = (BM * LDK > KC * LDMA) ? BM * LDK : KC * LDMA;
    constexpr int B_ELEMS = (BN * LDK > KC * LDNB) ? BN * LDK : KC * LDNB;
    constexpr int WARPS_N = BN / WN;
    static_assert((BM / WM) * (BN / WN) == 4, "4 warps per CTA");
    constexpr int FM = WM / 8, FN = WN / 8;
    constexpr int LA = BM * KC / 128, LB = BN * KC / 128;
    __shared__ double As[2][A_ELEMS];
    __shared__ double Bs[2][B_ELEMS];

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

    // Staging slots of a thread, element e = tid + 128 i.  Because 128 is a multiple of KC (and of BM / BN
    // except for the 56-wide tile) row and k index advance by constant steps with i, so only the i = 0 slot
    // is kept in registers; the odd tile width falls back to recomputing the slot.
    constexpr bool A_REG_M = (128 % BM == 0), B_REG_N = (128 % BN == 0);
    const int a_r0 = a_kc ? tid / KC : tid % BM, a_k0 = a_kc ? tid % KC : tid / BM;
    const int a_rs = a_kc ? 128 / KC : 0, a_ks = a_kc ? 0 : (A_REG_M ? 128 / BM : 0);
    const int b_r0 = b_kc ? tid / KC : tid % BN, b_k0 = b_kc ? tid % KC : tid / BN;
    const int b_rs = b_kc ? 128 / KC : 0, b_ks = b_kc ? 0 : (B_REG_N ? 128 / BN : 0);
    const bool a_reg = a_kc || A_REG_M, b_reg = b_kc || B_REG_N;
    const ll a_base = (ll)(m0 + a_r0) * d.a_m + (ll)a_k0 * d.a_k, a_step = (ll)a_rs * d.a_m + (ll)a_ks * d.a_k;
    const ll b_base = (ll)(n0 + b_r0) * d.b_n + (ll)b_k0 * d.b_k, b_step = (ll)b_rs * d.b_n + (ll)b_ks * d.b_k;
    const int a_sm0 = a_kc ? a_r0 * LDK + a_k0 : a_k0 * LDMA + a_r0, a_sms = a_kc ? a_rs * LDK : a_ks * LDMA;
    const int b_sm0 = b_kc ? b_r0 * LDK + b_k0 : b_k0 * LDNB + b_r0, b_sms = b_kc ? b_rs * LDK : b_ks * LDNB;
    double ra[LA], rb[LB];
    auto gload = [&](int c) {
        const int kleft = K - c * KC;
        const double* Ac = A + (ll)c * KC * d.a_k;
        const double* Bc = B + (ll)c * KC * d.b_k;
        if (a_reg) {
#pragma unroll
            for (int i = 0; i < LA; ++i)
                ra[i] = (m0 + a_r0 + i * a_rs < M && a_k0 + i * a_ks < kleft) ? __ldg(Ac + a_base + i * a_step) : 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < LA; ++i) {
                const int e = tid + 128 * i, m = e % BM, k = e / BM;
                ra[i] = (m0 + m < M && k < kleft) ? __ldg(Ac + (ll)(m0 + m) * d.a_m + (ll)k * d.a_k) : 0.0;
            }
        }
        if (b_reg) {
#pragma unroll
            for (int i = 0; i < LB; ++i)
                rb[i] = (n0 + b_r0 + i * b_rs < N && b_k0 + i * b_ks < kleft) ? __ldg(Bc + b_base + i * b_step) : 0.0;
        } else {
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                const int e = tid + 128 * i, n = e % BN, k = e / BN;
                rb[i] = (n0 + n < N && k < kleft) ? __ldg(Bc + (ll)k * d.b_k + (ll)(n0 + n) * d.b_n) : 0.0;
            }
        }
    };
    auto sstore = [&](int buf) {
        if (a_reg) {
#pragma unroll
            for (int i = 0; i < LA; ++i) As[buf][a_sm0 + i * a_sms] = ra[i];
        } else {
#pragma unroll
            for (int i = 0; i < LA; ++i) { const int e = tid + 128 * i; As[buf][(e / BM) * LDMA + e % BM] = ra[i]; }
        }
        if (b_reg) {
#pragma unroll
            for (int i = 0; i < LB; ++i) Bs[buf][b_sm0 + i * b_sms] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < LB; ++i) { const int e = tid + 128 * i; Bs[buf][(e / BN) * LDNB + e % BN] = rb[i]; }
        }
    };

    const int nchunks = (K + KC - 1) / KC;
    int k_lo = 0;                                  // k % tri_kperiod below k_lo is structurally zero for this row tile
    if (d.tri_div > 0) {
        const int jlo = m0 % d.tri_mperiod, jhi = min(m0 + BM, M) - 1;
        const bool wraps = (jhi / d.tri_mperiod) != (m0 / d.tri_mperiod);
        k_lo = (wraps || jlo <= d.tri_off) ? 0 : (jlo - d.tri_off) / d.tri_div;
    }
    // Chunk skipping walks (chunk, position of the chunk's first column inside its period) together: the position used to be
    // recomputed as (c * KC) % tri_kperiod - a runtime integer modulo, ~25 instructions, twice per test - and that single line held
    // 12 % of the kernel's stall samples and 7 % of its executed instructions (ncu source page, profiles/r2_ncu_summary.md).
    const bool tri = d.tri_div > 0 && k_lo > 0;
    const int kperiod = d.tri_kperiod;
    auto skip_zero_chunks = [&](int& c, int& kp) {
        if (tri)
            while (c < nchunks && kp + KC - 1 < k_lo && kp + KC <= kperiod) {
                ++c; kp += KC;
                while (kp >= kperiod) kp -= kperiod;
            }
    };
    // fragment read offsets (elements) for kk = 0; advancing kk by 4 adds 4 (K-contiguous) or 4 * LD (row-contiguous)
    const int a_f0 = a_kc ? (wm0 + g) * LDK + t4 : t4 * LDMA + wm0 + g;
    const int a_fi = a_kc ? 8 * LDK : 8;            // next m-fragment
    const int a_fk = a_kc ? 4 : 4 * LDMA;           // next k-step
    const int b_f0 = b_kc ? (wn0 + g) * LDK + t4 : t4 * LDNB + wn0 + g;
    const int b_fj = b_kc ? 8 * LDK : 8;
    const int b_fk = b_kc ? 4 : 4 * LDNB;

    int c = 0, kp = 0;
    skip_zero_chunks(c, kp);
    if (c < nchunks) { gload(c); sstore(0); }
    __syncthreads();
    int buf = 0;
    while (c < nchunks) {
        int cn = c + 1, kpn = kp + KC;
        if (tri) { while (kpn >= kperiod) kpn -= kperiod; }
        skip_zero_chunks(cn, kpn);
        if (cn < nchunks) gload(cn);
        const double* as = As[buf] + a_f0;
        const double* bs = Bs[buf] + b_f0;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; ++i) af[i] = as[ks * a_fk + i * a_fi];
#pragma unroll
            for (int j = 0; j < FN; ++j) bf[j] = bs[ks * b_fk + j * b_fj];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (cn < nchunks) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
        c = cn; kp = kpn;
    }

    double alpha = d.alpha;
    if (d.alpha_vec) alpha *= d.alpha_vec[inst];
    const double beta = d.beta;
#pragma unroll
    for (int i = 0; i < FM; ++i) {
        const int gm = m0 + wm0 + i * 8 + g;
        if (gm >= M) continue;
        double* crow = C + (d.c_msplit ? (ll)(gm / d.c_msplit) * d.c_mhi + (ll)(gm % d.c_msplit) * d.c_m : (ll)gm * d.c_m);
#pragma unroll
        for (int j = 0; j < FN; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + wn0 + j * 8 + 2 * t4 + e;
                if (gn >= N) continue;
                double* p = crow + (ll)gn * d.c_n;
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
