/*
 * qtt_b200.h - C ABI of the B200-native (sm_100a, FP64) tensor-train kernels that replace the numpy
 * hot path of MScherbela/LaplaceSolverMPS (`laplace_mps.tensormethods.TensorTrain`).
 *
 * The reference has no FFI of its own: its boundary is the Python class
 * `TensorTrain` and `solver.solve_with_grad_descent`.  Each entry point below names the reference
 * lines whose arithmetic it replaces; `laplacesolvermps_b200/_lib.py` is the ctypes binding and
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns an int status, 0 == QTT_OK; no exceptions cross the ABI;
 *   - the caller owns every buffer; the library owns only the opaque handle and its workspace;
 *   - all device work is enqueued on the cudaStream_t passed in (as void*); calls that must read
 *     rank tables back synchronise that stream internally and say so below;
 *   - a handle is not thread-safe: one per host thread / stream.
 *
 * Batched TT layout ("qtt_batch"): N independent instances of an L-core tensor train.
 *   ranks  : device int32 [N][L+1], ranks[i][0] = ranks[i][L] = 1, ranks[i][k] = bond left of core k
 *   cores  : host array of L device pointers; cores[k] holds N slots of slot[k] doubles; instance i's
 *            core k is the C-ordered array (r_k, n_k, r_{k+1}) - exactly numpy's layout, compact with
 *            the instance's own ranks - at cores[k] + i*slot[k]
 *   n      : host int [L], mode size of every core (all mode dimensions flattened)
 */
#ifndef QTT_B200_H
#define QTT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTT_OK 0
#define QTT_ERR_INVALID 1      /* bad argument / shape mismatch */
#define QTT_ERR_CUDA 2         /* a CUDA runtime call failed (see qtt_last_error) */
#define QTT_ERR_SLOT 3         /* an output slot is too small for the rank bound */
#define QTT_ERR_UNSUPPORTED 4  /* size outside what the kernels cover */
#define QTT_ERR_NOCONV 5       /* a Jacobi SVD did not converge (status array has details) */

typedef struct qtt_handle_s* qtt_handle_t;

typedef struct qtt_batch {
    int L;
    int N;
    const int* n;            /* host  [L]   */
    int* ranks;              /* device [N*(L+1)] */
    double* const* cores;    /* host  [L] device pointers */
    const long long* slot;   /* host  [L] doubles per instance slot */
} qtt_batch;

/* One operator (MPO) shared by all instances: core k is the C-ordered (R_k, n_out_k, n_in_k, R_{k+1}). */
typedef struct qtt_mpo {
    int L;
    const int* n_out;        /* host [L] */
    const int* n_in;         /* host [L] */
    const int* ranks;        /* host [L+1], ranks[0] = ranks[L] = 1 */
    double* const* cores;    /* host [L] device pointers */
} qtt_mpo;

/* One term of a sum that is rounded without being materialised:
 *   coef * x               (op == NULL)      or      coef * (op @ x)    (op != NULL)
 * coef is a device array [N] of per-instance scalars (NULL means 1.0), scaled by coef_scale. */
typedef struct qtt_term {
    const qtt_mpo* op;
    const qtt_batch* x;
    const double* coef;
    double coef_scale;
} qtt_term;

int qtt_create(qtt_handle_t* handle, int device);
int qtt_destroy(qtt_handle_t handle);
const char* qtt_last_error(qtt_handle_t handle);
int qtt_version(void);
/* Upper bound, in bytes, on the workspace the handle may keep cached (default 64 GiB). */
int qtt_set_workspace_limit(qtt_handle_t handle, size_t bytes);
/* Cumulative EXECUTED FP64 flops of the DMMA-class kernels (GEMM, block-reflector update, Gram; structural-zero chunks
 * that are skipped are not counted) and number of kernel launches since the last reset. */
int qtt_get_counters(qtt_handle_t handle, double* gemm_flops, long long* launches);
/* The same counter split by kernel class: dmma_flops as above; leaf_flops = executed flops of the panel QR leaves
 * (per outer panel of `rows` rows and jb columns in sub-panels of w: 2 rows jb^2 + 2 rows jb w).  Their sum is what
 * bench.py's whole-step roofline fraction is computed from. */
int qtt_get_flop_counters(qtt_handle_t handle, double* dmma_flops, double* leaf_flops);
int qtt_reset_counters(qtt_handle_t handle);
/* CUDA-graph cache of launch-bound sweeps (small cores: the solver's own roundings, the 1D configurations): number of sweeps
 * captured into a graph and number of sweeps replayed from one since the handle was created.  QTT_GRAPHS=0 turns the cache off. */
int qtt_get_graph_counters(qtt_handle_t handle, long long* captures, long long* replays);
/* Per-launch timing of the dominant (GEMM) kernel with CUDA events on the launching stream: switch on,
 * run, then read the summed duration (ms), the executed flops of those launches (2*M*N*K minus skipped structural-zero
 * chunks) and their count. */
int qtt_set_profiling(qtt_handle_t handle, int on);
int qtt_get_profile(qtt_handle_t handle, double* gemm_ms, double* gemm_flops, long long* gemm_launches, int reset);
/* qtt_set_profiling(h, 2) switches on a timeline mode instead: an event after every launch, the interval since the
 * previous event is attributed to the launch that finished.  ms8[0..5] = gemm, larfb, QR leaf, gram/larft, Jacobi SVD, other. */
int qtt_get_class_profile(qtt_handle_t handle, double* ms8, int reset);

/* Core-wise product y = a @ x, materialised.                      tensormethods.py:176-194
 * `a` per instance (a_shared == NULL) or one shared operator.  Core k of `a` is treated as
 * (p, na_k, m_k, q) and core k of `x` as (P, m_k, nx_k, Q); y core k is (pP, na_k*nx_k, qQ).
 * This covers all four ndim pairings of the reference (MPO@MPS, MPO@MPO, MPS@MPO, MPS@MPS).
 * m[k] = contracted mode size; a.n[k] = na_k*m_k, x.n[k] = m_k*nx_k, y.n[k] = na_k*nx_k.
 * m[k] == 1 gives the Kronecker product of the mode indices (level-interleaved 2D products, utils.py:182-185);
 * m[k] == 0 (per-instance `a` only) the Hadamard core product with equal mode index on both sides
 * (`TensorTrain.__mul__`, tensormethods.py:154-167): y core k = (pP, n_k, qQ), y[(p P), i, (q Q)] = a[p,i,q] x[P,i,Q].
 * Writes y.ranks.  Asynchronous. */
int qtt_matmul_batched(qtt_handle_t h, const qtt_batch* a, const qtt_mpo* a_shared, const qtt_batch* x,
                       const int* m, qtt_batch* y, void* stream);

/* Batched function encoders: the trains that enter the hot path (right-hand sides, exact solutions) built on the device.
 *
 * qtt_encode_linear_batched                                        utils.py:188-205 (get_polynomial_as_tt)
 *   A train whose cores are shared by all instances except for a linear dependence of core 0 on nc per-instance
 *   coefficients: y core 0 = sum_j coef[i][j] * shared_cores[0][j] (shared_cores[0] holds nc blocks of n_0 * ranks[1]
 *   doubles), y core k >= 1 = shared_cores[k] ((ranks[k], n_k, ranks[k+1]), device pointers in a host array).
 *   For polynomials: coef = monomial coefficients, shared_cores[0][j] = Legendre expansion of x^j pushed through the
 *   first zoom core, shared_cores[k] = the zoom core, last = corner evaluation.
 * qtt_encode_trig_batched                                          utils.py:208-225 (get_trig_function_as_tt)
 *   abnu: device [N][3] = (a, b, nu) of a cos(2 pi nu x) + b sin(2 pi nu x); y has L + 1 cores of mode size 2, bonds 2.
 * Both write y.ranks and synchronise the stream (a host-side rank table is uploaded). */
int qtt_encode_linear_batched(qtt_handle_t h, int nc, const double* coef, const double* const* shared_cores,
                              const int* ranks, qtt_batch* y, void* stream);
int qtt_encode_trig_batched(qtt_handle_t h, const double* abnu, qtt_batch* y, void* stream);

/* TT rounding of a sum of terms, y = round(sum_t coef_t * term_t), without materialising it:
 * right-to-left blocked Householder QR sweep (tensormethods.py:90-99), then left-to-right
 * truncated-SVD sweep (one-sided Jacobi) with the reference's rank rule
 *   s_k = min(rows, q_k, caps[k], #{sigma_i : (sigma_i/sigma_0)^2 >= rel_eps^2})    (tensormethods.py:5-12,101-133).
 * caps: host int [L-1] (NULL = no cap).  status: device int [N] (may be NULL), 0 = ok, else the
 * number of Jacobi SVDs that hit the sweep limit.  Synchronises `stream` (reads the rank tables). */
int qtt_round_sum_batched(qtt_handle_t h, int n_terms, const qtt_term* terms, qtt_batch* y,
                          const int* caps, double rel_eps, int* status, void* stream);

/* Orthogonalisation sweep alone (tensormethods.py:90-99): y = the same tensor with cores 1..L-1
 * row-orthonormal; norm2 (device [N], may be NULL) receives sum(core_0^2) = ||x||^2
 * (tensormethods.py:280-283).  If y == NULL only norm2 is produced.  Synchronises `stream`. */
int qtt_orthogonalize_batched(qtt_handle_t h, int n_terms, const qtt_term* terms, qtt_batch* y,
                              double* norm2, void* stream);

/* Inner products <a_i, b_i> as left-to-right transfer-matrix contractions (replaces the
 * `(a @ b).squeeze().eval()` idiom, tensormethods.py:187-188, 233-264).  out: device [N].
 * If op != NULL computes the bilinear form <a_i, op @ b_i>.  Asynchronous. */
int qtt_inner_batched(qtt_handle_t h, const qtt_batch* a, const qtt_mpo* op, const qtt_batch* b,
                      double* out, void* stream);

/* Block-diagonal sum y = alpha*a + beta*b, materialised (tensormethods.py:135-174).
 * alpha, beta: device [N] or NULL (= 1).  Asynchronous. */
int qtt_axpby_batched(qtt_handle_t h, const double* alpha, const qtt_batch* a, const double* beta,
                      const qtt_batch* b, qtt_batch* y, void* stream);

/* Truncated steepest descent, device resident (solver.py:251-275), for N right-hand sides b and one
 * shared operator A.  x: output batch (slots sized for max_rank), r2: device [N] final ||b - A x||^2.
 * hist: optional device [n_steps][N][2] (r2, gamma) per step.  Every instance runs exactly n_steps
 * steps when rel_accuracy <= 0 (the benchmark setting); otherwise instances that meet the reference's
 * stopping test at a multiple of 10 steps are frozen.  Synchronises `stream`. */
int qtt_gd_solve_batched(qtt_handle_t h, const qtt_mpo* A, const qtt_batch* b, qtt_batch* x, double* r2,
                         int n_steps, double lr, int max_rank, int recalc_every, double rel_accuracy,
                         double* hist, int* steps_done, void* stream);

/* Truncated conjugate gradients for a symmetric positive definite operator, device resident: the iterative solver
 * offered behind `solve_with_amen` (solver.py:236-248 delegates to the third-party ttpy AMEn, which is not vendored;
 * SURVEY 8(f).3).  Same arguments and conventions as qtt_gd_solve_batched without the step length; hist: optional
 * device [n_steps][N][2] (||r||^2, alpha) per step.  Synchronises `stream`. */
int qtt_cg_solve_batched(qtt_handle_t h, const qtt_mpo* A, const qtt_batch* b, qtt_batch* x, double* r2,
                         int n_steps, int max_rank, int recalc_every, double rel_accuracy,
                         double* hist, int* steps_done, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QTT_B200_H */
