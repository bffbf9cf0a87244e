// Push GEMM: the specialised form of gemm_dmma_kernel for the contractions that dominate a rounding sweep,
//     M_k[(no, j), (Rl, r)] = sum_{(m, R')} T[(m, R'), (r, j)] * Op[Rl, no, (m, R')]          (form_block in qtt_b200.cu)
// and the head-compression / dense-core products of the same shape class: A is M-contiguous (a_m == 1), B is K-contiguous
// (b_k == 1, BNC = false) or N-contiguous (b_n == 1, BNC = true), rows of both operands start 16-byte aligned.
//
// Why a second kernel: ncu on the generic kernel at this shape (profiles/r2_ncu_summary.md) shows the DMMA pipe 58 % busy at
// 3 warps per scheduler (166 registers: register-staged global loads) with 12.8 issued instructions per DMMA - run-time
// strides and layouts put 64-bit address arithmetic and bound predicates around every load.  Here
//   * both operands go global -> shared with 16-byte cp.async (LDGSTS, zero-filled at the edges) through a 3-stage ring:
//     no staging registers, one barrier per 16-deep K chunk, loads two chunks ahead;
//   * layouts and strides inside the tile are compile-time, fragment reads use immediate offsets;
//   * 4 CTAs (16 warps) per SM;
//   * the last K chunk runs only the 4-deep DMMA steps that hold data (K = 161 is 40.25 steps, not 44).
// Same tile shape (64 x BN, four warps of 16 x BN), same structural-zero chunk skipping, same epilogue as the generic kernel.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

#define PUSH_BM 64
#define PUSH_KC 16
#define PUSH_NST 3

template <int BN, bool BNC>
static constexpr size_t gemm_push_smem() {
    return (size_t)PUSH_NST * (PUSH_KC * (PUSH_BM + 4) + (BNC ? PUSH_KC * (BN + 4) : BN * (PUSH_KC + 4))) * sizeof(double);
}

__device__ __forceinline__ void cp_async16_zfill(unsigned dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(bytes));
}

template <int BN, bool BNC>
__global__ void __launch_bounds__(128, 4) gemm_push_kernel(const GemmDesc d, int tiles_m, int tiles_n) {
    constexpr int BM = PUSH_BM, KC = PUSH_KC, NST = PUSH_NST;
    constexpr int LDA = BM + 4;                 // A stage: S[k * LDA + m]   (== 4 mod 16: fragment reads conflict free)
    constexpr int LDB = BNC ? BN + 4 : KC + 4;  // B stage: S[k * LDB + n] (N-contiguous) or S[n * LDB + k] (K-contiguous)
    constexpr int A_ST = KC * LDA, B_ST = BNC ? KC * LDB : BN * LDB, STAGE = A_ST + B_ST;
    constexpr int FN = BN / 8;
    constexpr int NB_IT = BNC ? 4 : (BN * 8 + 127) / 128;   // 16-byte pieces of a B stage per thread
    extern __shared__ __align__(16) double smem[];

    const int tiles = tiles_m * tiles_n;
    const ll bid = blockIdx.x;
    const int z = (int)(bid / tiles);
    const int tile = (int)(bid % tiles);
    const int tm = tile % tiles_m, tn = tile / tiles_m;
    const int i2 = z % d.nb[2];
    const int i1 = (z / d.nb[2]) % d.nb[1];
    const int i0 = z / (d.nb[2] * d.nb[1]);
    const int inst = d.idx ? d.idx[i0] : i0;
    const int M = d.M, N = d.N, K = d.K;
    const int m0 = tm * BM, n0 = tn * BN;

    const double* A = d.A + (d.a_idx ? inst : i0) * d.a_s[0] + i1 * d.a_s[1] + i2 * d.a_s[2];
    const double* B = d.B + (d.b_idx ? inst : i0) * d.b_s[0] + i1 * d.b_s[1] + i2 * d.b_s[2];
    double* C = d.C + (d.c_idx ? inst : i0) * d.c_s[0] + i1 * d.c_s[1] + i2 * d.c_s[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm0 = warp * 16;

    double acc[2][FN][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // Loader slots of this thread (16-byte pieces = two consecutive elements along the contiguous index):
    //   A            piece i -> k = (tid >> 5) + 4 i, rows m0 + 2 (tid & 31) .. + 1
    //   B, K-contig  piece i -> column n0 + (tid >> 3) + 16 i, k = 2 (tid & 7) .. + 1
    //   B, N-contig  piece i -> k = (tid >> 5) + 4 i, columns n0 + 2 (tid & 31) .. + 1   (lanes beyond BN idle)
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(smem);
    const int a_k = tid >> 5, a_m = (tid & 31) * 2;
    const int a_bytes = max(0, min(2, M - (m0 + a_m))) * 8;     // an odd edge keeps one element, the rest of the piece is zero-filled
    const double* a_src = A + (ll)a_k * d.a_k + m0 + a_m;
    const ll a_step = 4 * d.a_k;
    const unsigned a_dst = s_base + (unsigned)((a_k * LDA + a_m) * sizeof(double));
    const int b_r = BNC ? (tid >> 5) : (tid >> 3);               // slow index inside the stage: k (N-contig) or n (K-contig)
    const int b_c = BNC ? (tid & 31) * 2 : (tid & 7) * 2;        // contiguous index: n (N-contig) or k (K-contig)
    const double* b_src = BNC ? B + (ll)b_r * d.b_k + n0 + b_c : B + (ll)(n0 + b_r) * d.b_n + b_c;
    const ll b_step = BNC ? 4 * d.b_k : 16 * d.b_n;
    const unsigned b_dst = s_base + (unsigned)((A_ST + b_r * LDB + b_c) * sizeof(double));
    const int bn_bytes = BNC ? ((b_c < BN) ? max(0, min(2, N - (n0 + b_c))) * 8 : 0) : 0;
    unsigned b_mask = 0;                                          // K-contig: which of the thread's columns exist
    if (!BNC) {
#pragma unroll
        for (int i = 0; i < NB_IT; ++i)
            if (tid + 128 * i < BN * 8 && n0 + b_r + 16 * i < N) b_mask |= 1u << i;
    }

    auto issue = [&](int k0, int stage) {
        const unsigned so = (unsigned)(stage * STAGE * sizeof(double));
        const double* as = a_src + (ll)k0 * d.a_k;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int nb = (k0 + a_k + 4 * i < K) ? a_bytes : 0;
            cp_async16_zfill(a_dst + so + (unsigned)(4 * i * LDA * sizeof(double)), nb ? (const void*)(as + i * a_step) : (const void*)A, nb);
        }
        if (BNC) {
            const double* bs = b_src + (ll)k0 * d.b_k;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nb = (k0 + b_r + 4 * i < K) ? bn_bytes : 0;
                if (b_c < BN)
                    cp_async16_zfill(b_dst + so + (unsigned)(4 * i * LDB * sizeof(double)), nb ? (const void*)(bs + i * b_step) : (const void*)B, nb);
            }
        } else {
            const int kb = max(0, min(2, K - (k0 + b_c))) * 8;
            const double* bs = b_src + k0;
#pragma unroll
            for (int i = 0; i < NB_IT; ++i) {
                if (tid + 128 * i < BN * 8) {
                    const int nb = ((b_mask >> i) & 1u) ? kb : 0;
                    cp_async16_zfill(b_dst + so + (unsigned)(16 * i * LDB * sizeof(double)), nb ? (const void*)(bs + i * b_step) : (const void*)B, nb);
                }
            }
        }
    };

    // Structural zeros of A (see GemmDesc): inside every period of tri_kperiod columns the first k_lo columns are zero for all rows
    // of this tile.  The K loop walks the ACTIVE SEGMENTS [p * kperiod + k_lo, (p + 1) * kperiod) only, each in 16-deep chunks
    // that start at the segment (rounded down to a multiple of 4: DMMA step and 16-byte alignment), so the zero columns of a
    // chunk are at most 3 in front and the overshoot into the next period's zero gap behind (needs k_lo >= 18: the overshoot
    // of at most 15 columns must not reach the next segment - tiles with a smaller k_lo run dense).  Chunks aligned to multiples
    // of 16 of the global column index, as in the generic kernel, straddle the period boundaries (161 = 10.06 chunks) and leave
    // 65 % of the chunks of the main push active where 50 % of the factor is non-zero; segments bring that to 57 %.
    int k_lo = 0;
    if (d.tri_div > 0) {
        const int jlo = m0 % d.tri_mperiod, jhi = min(m0 + BM, M) - 1;
        const bool wraps = (jhi / d.tri_mperiod) != (m0 / d.tri_mperiod);
        k_lo = (wraps || jlo <= d.tri_off) ? 0 : (jlo - d.tri_off) / d.tri_div;
    }
    const bool seg = k_lo >= 18;
    const int kper = seg ? d.tri_kperiod : K;
    const int klo = seg ? k_lo : 0;
    constexpr int K_END = 1 << 30;
    auto next_seg = [&](int& k0, int& ep, int& pb) {     // first chunk of the first non-empty segment at or after period base pb
        while (pb < K) {
            ep = (kper >= K - pb) ? K : pb + kper;
            if (pb + klo < ep) { k0 = (pb + klo) & ~3; return; }
            pb = ep;
        }
        k0 = K_END;
    };
    auto advance = [&](int& k0, int& ep, int& pb) {
        k0 += KC;
        if (k0 >= ep) { pb = ep; next_seg(k0, ep, pb); }
    };

    int kl = 0, epl = 0, pbl = 0;                // loader cursor: chunk start, end of its segment, base of its period
    next_seg(kl, epl, pbl);
    int kc = kl, epc = epl, pbc = pbl;           // consumer cursor (same walk, NST - 1 chunks behind)
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (kl < K_END) { issue(kl, s); advance(kl, epl, pbl); }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    const double* a_frag = smem + t4 * LDA + wm0 + g;
    const double* b_frag = smem + A_ST + (BNC ? t4 * LDB + g : g * LDB + t4);
    constexpr int B_FJ = BNC ? 8 : 8 * LDB;      // next n-fragment
    constexpr int B_FK = BNC ? 4 * LDB : 4;      // next k-step
    int stage = 0, lstage = NST - 1;
    while (kc < K_END) {
        asm volatile("cp.async.wait_group %0;\n" :: "n"(NST - 2) : "memory");
        __syncthreads();                         // chunk `stage` has landed for every thread; stage `lstage` was consumed one iteration ago
        if (kl < K_END) { issue(kl, lstage); advance(kl, epl, pbl); }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        const double* as = a_frag + stage * STAGE;
        const double* bs = b_frag + stage * STAGE;
        const int kleft = epc - kc;              // columns of this chunk inside its segment (what lies beyond is zero or past K)
        if (kleft >= KC) {
#pragma unroll
            for (int ks = 0; ks < KC / 4; ++ks) {
                double af[2], bf[FN];
#pragma unroll
                for (int i = 0; i < 2; ++i) af[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
                for (int j = 0; j < FN; ++j) bf[j] = bs[j * B_FJ + ks * B_FK];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        } else {
            const int nks = (kleft + 3) >> 2;    // the last chunk of a segment: only the steps that hold data
            for (int ks = 0; ks < nks; ++ks) {
                double af[2], bf[FN];
#pragma unroll
                for (int i = 0; i < 2; ++i) af[i] = as[ks * 4 * LDA + i * 8];
#pragma unroll
                for (int j = 0; j < FN; ++j) bf[j] = bs[j * B_FJ + ks * B_FK];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        advance(kc, epc, pbc);
        stage = (stage + 1 == NST) ? 0 : stage + 1;
        lstage = (lstage + 1 == NST) ? 0 : lstage + 1;
    }

    double alpha = d.alpha;
    if (d.alpha_vec) alpha *= d.alpha_vec[inst];
    const double beta = d.beta;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gm = m0 + wm0 + i * 8 + g;
        if (gm >= M) continue;
        double* crow = C + (d.c_msplit ? (ll)(gm / d.c_msplit) * d.c_mhi + (ll)(gm % d.c_msplit) * d.c_m : (ll)gm * d.c_m);
#pragma unroll
        for (int j = 0; j < FN; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gn = n0 + j * 8 + 2 * t4 + e;
                if (gn >= N) continue;
                double* p = crow + (ll)gn * d.c_n;
                double v = alpha * acc[i][j][e];
                if (beta != 0.0) v += beta * (*p);
                *p = v;
            }
        }
    }
}

// Can this launch take the specialised kernel (and with which B layout)?  0: no (generic gemm_dmma_kernel), 1: B K-contiguous, 2: B N-contiguous
static inline int gemm_push_eligible(const GemmDesc& d) {
    if (d.a_m != 1 || d.n_dyn) return 0;
    if (d.M <= 32 || d.N <= 16) return 0;
    if (d.a_k & 1) return 0;
    for (int i = 0; i < 3; ++i) if ((d.a_s[i] & 1) || (d.b_s[i] & 1)) return 0;
    if ((((unsigned long long)d.A) | ((unsigned long long)d.B)) & 15ull) return 0;
    if (d.b_k == 1 && !(d.b_n & 1)) return 1;
    if (d.b_n == 1 && !(d.b_k & 1)) return 2;
    return 0;
}

static int g_gemm_push = 1;      // QTT_GEMM_PUSH=0 (read at qtt_create): every GEMM through the generic kernel (A/B)

// Launch; dims that are zero make it a no-op (but beta == 0 still requires C to be defined: callers
// never rely on a K == 0 launch to clear C).
static int gemm_launch(qtt_handle_s* h, cudaStream_t st, const GemmDesc& d) {
    if (d.M <= 0 || d.N <= 0 || d.nb[0] <= 0 || d.nb[1] <= 0 || d.nb[2] <= 0) return QTT_OK;
    const ll nbatch = (ll)d.nb[0] * d.nb[1] * d.nb[2];
    auto launch = [&](auto kern, int BM, int BN, size_t smem, bool segments) -> int {
        int tmn = (int)ceil_div(d.M, BM), tnn = (int)ceil_div(d.N, BN);
        ll blocks = nbatch * tmn * tnn;
        if (blocks > 2147483647LL) { h->err = "gemm grid too large"; return QTT_ERR_UNSUPPORTED; }
        // algorithmic flops actually executed: structurally-zero columns that the kernel skips are not counted (the generic
        // kernel skips whole 16-column chunks of the global index, the push kernel walks the active segments - see there).
        // (Skipping 4-deep steps per warp inside active chunks as well removes 15 % of the DMMAs of the main push and not one
        //  microsecond: the warp with the lowest rows skips nothing and the CTA waits for it at the chunk barrier - measured, removed.)
        double kexec = (double)d.K * d.M;
        if (d.tri_div > 0) {
            const int KC = 16, nchunks = (d.K + KC - 1) / KC;
            kexec = 0.0;
            for (int tm = 0; tm < tmn; ++tm) {
                const int m0 = tm * BM, mrows = std::min(BM, d.M - m0);
                const int jlo = m0 % d.tri_mperiod, jhi = m0 + mrows - 1;
                const bool wraps = (jhi / d.tri_mperiod) != (m0 / d.tri_mperiod);
                const int k_lo = (wraps || jlo <= d.tri_off) ? 0 : (jlo - d.tri_off) / d.tri_div;
                if (segments) {
                    const bool seg = k_lo >= 18;
                    const ll kper = seg ? d.tri_kperiod : d.K;
                    const int klo = seg ? k_lo : 0;
                    for (ll pb = 0; pb < d.K; pb += kper) {
                        const ll ep = std::min<ll>(pb + kper, d.K);
                        if (pb + klo >= ep) continue;
                        const ll k0 = (pb + klo) & ~3LL;
                        kexec += (double)(((ep - k0 + 3) / 4) * 4) * mrows;       // executed 4-deep steps of the segment
                    }
                    continue;
                }
                for (int c = 0; c < nchunks; ++c) {
                    const int kp = (c * KC) % d.tri_kperiod;
                    if (kp + KC - 1 < k_lo && kp + KC <= d.tri_kperiod) continue;
                    kexec += (double)std::min(KC, d.K - c * KC) * mrows;
                }
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
        kern<<<(unsigned)blocks, 128, smem, st>>>(d, tmn, tnn);
        if (h->prof_on) cudaEventRecord(h->prof_ev[2 * slot + 1], st);
        tl_mark(h, st, 0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { h->err = std::string("gemm launch: ") + cudaGetErrorString(e); return QTT_ERR_CUDA; }
        return QTT_OK;
    };
    if (d.N <= 16) return launch(gemm_dmma_kernel<128, 16, 32, 16>, 128, 16, 0, false);
    if (d.M <= 32) return launch(gemm_dmma_kernel<32, 64, 16, 32>, 32, 64, 0, false);
    const bool n56 = ceil_div(d.N, 56) * 56 < ceil_div(d.N, 64) * 64;
    const int push = g_gemm_push ? gemm_push_eligible(d) : 0;     // A M-contiguous, rows 16-byte aligned: cp.async ring, 4 CTAs / SM
    if (push == 1) {
        if (n56) return launch(gemm_push_kernel<56, false>, 64, 56, gemm_push_smem<56, false>(), true);
        return launch(gemm_push_kernel<64, false>, 64, 64, gemm_push_smem<64, false>(), true);
    }
    if (push == 2) {
        if (n56) return launch(gemm_push_kernel<56, true>, 64, 56, gemm_push_smem<56, true>(), true);
        return launch(gemm_push_kernel<64, true>, 64, 64, gemm_push_smem<64, true>(), true);
    }
    if (n56) return launch(gemm_dmma_kernel<64, 56, 16, 56>, 64, 56, 0, false);
    return launch(gemm_dmma_kernel<64, 64, 32, 32>, 64, 64, 0, false);
}
