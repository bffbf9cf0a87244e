"""Operators, PDE drivers and solvers (API of the reference's `laplace_mps.solver`,
reference file src/laplace_mps/solver.py).

Operator assembly is construction-time host code on cores of rank <= 32 (see _assembly.py).  The solver
`solve_with_grad_descent` (solver.py:251-275) and the pre/post-processing of the `solve_PDE_*` drivers
(`R @ f`, `C @ v`, `Q' @ v`, the roundings and the norms; solver.py:192-233, 420-441) are the hot path
and run on the GPU: every `@`, `.reapprox()` and `.norm_squared()` on a `TensorTrain` below is a call
into csrc/libqtt_b200.so.  `solve_with_grad_descent_batched` runs many right-hand sides at once,
device resident (qtt_gd_solve_batched).

`solve_with_amen` keeps the reference's signature; the reference delegates it to the third-party
Fortran package ttpy (un-vendored, un-pinned: parity unpinned).  If ttpy is importable it is used exactly
as the reference does; otherwise the call is served by the device steepest-descent solver.
"""
import numpy as np

from . import tensormethods as tm
from . import _assembly as _asm
from .bpx2D import (get_laplace_BPX_2D, get_BPX_preconditioner_2D, get_rhs_matrix_BPX_2D, get_BPX_Qp_2D,
                    _get_U_hat, _get_Ub_hat, _get_X_hat, _get_Z1_hat, mode_product, transpose_core, strong_kronecker)
from .utils import kronecker_prod_2D


def _compose(*factors):
    """Construction-time operator product of TensorTrains (host)."""
    cores = factors[0].tensors
    for f in factors[1:]:
        cores = _asm.compose(cores, f.tensors)
    return tm.TensorTrain(cores)


def _rounded(tt, **kw):
    _asm.round_(tt.tensors, **kw)
    return tt


def _scale_cores(tt, factor, count):
    for i in range(count):
        tt.tensors[i] = tt.tensors[i] * factor
    return tt


# ---- 1D nodal operators --------------------------------------------------------------------------

def _get_gram_matrix_legendre():
    return np.diag([2, 2 / 3])


_CORNER_TO_LEGENDRE = np.array([[1, 1], [-1, 1]])


def _get_refinement_tensor():
    """[[I, J^T], [0, J]] as a (2, 2, 2, 2) core: carries "same cell" / "neighbouring cell" through the levels."""
    return _get_U_hat()


def _chain_from_first_row(core, L):
    cores = [core.copy() for _ in range(L)]
    cores[0] = cores[0][:1]                     # start in state 0
    return cores


def get_overlap_matrix_as_tt(L):
    """M: nodal hat coefficients -> values at the two corners of every cell (as mean / half difference)."""
    cores = _chain_from_first_row(_get_refinement_tensor(), L)
    cores.append(np.array([[1, 1], [1, -1]]).reshape(2, 2, 1, 1) / 2)
    return tm.TensorTrain(cores)


def get_derivative_matrix_as_tt(L):
    """M': nodal hat coefficients -> slope in every cell."""
    cores = _chain_from_first_row(_get_refinement_tensor() * 2, L)
    cores[-1] = np.tensordot(cores[-1], np.array([[1.0], [-1.0]]), axes=[-1, 0])
    return tm.TensorTrain(cores)


def get_laplace_matrix_as_tt(L):
    D = get_derivative_matrix_as_tt(L)
    A = _compose(D.copy().transpose(), D)
    return _rounded(_scale_cores(A, 0.5, L), rel_error=1e-16)


def get_gram_matrix_as_tt(L, basis='legendre'):
    G = _get_gram_matrix_legendre()
    if basis == 'corner':
        G = G @ _CORNER_TO_LEGENDRE
    elif basis != 'legendre':
        raise ValueError("Unknown basis")
    cores = [(np.eye(2) / 2)[None, ..., None] for _ in range(L)]
    cores.append(G[None, ..., None] / 2)
    return tm.TensorTrain(cores)


def get_rhs_matrix_as_tt(L, basis="corner"):
    R = get_overlap_matrix_as_tt(L).transpose() * 0.5
    return _compose(R, get_gram_matrix_as_tt(L, basis))


def to_legendre_basis(t: tm.TensorTrain):
    t = t.copy()
    t.tensors[-1] = (t.tensors[-1].squeeze(-1) @ (np.array([[1, -1], [1, 1]]) / 2))[..., None]
    return t


def build_mass_matrix_in_nodal_basis(L):
    M = get_overlap_matrix_as_tt(L)
    return _rounded(_compose(M.copy().transpose(), get_gram_matrix_as_tt(L), M).squeeze())


# ---- 1D BPX operators ------------------------------------------------------------------------------

def _get_bpx_factors():
    """The four level cores of the 1D BPX factors (report section 2):
    W = U U^T (both prolongations carried), U' = W |><| (+-) (derivative on the finest carried level),
    V = X X^T (hat prolongation on both sides), Y = (1/2) X^T (averaging)."""
    W = transpose_core(_get_Ub_hat())
    U = strong_kronecker(W, np.kron(np.eye(2), np.array([[1.0], [-1.0]])).reshape(4, 1, 1, 2))
    X = _get_X_hat()
    V = transpose_core(mode_product(X, transpose_core(X)))
    Y = _get_Z1_hat()
    return W.copy(), U, V.copy(), Y.copy()


def get_bpx_preconditioner(L):
    """C = sum_l 2^-l P_l P_l^T, rank 8, block cores [[W/2, W/2], [0, V/2]]."""
    W, _, V, _ = _get_bpx_factors()
    core = np.zeros((8, 2, 2, 8))
    core[:4, :, :, :4] = core[:4, :, :, 4:] = W / 2
    core[4:, :, :, 4:] = V / 2
    first = np.zeros((1, 1, 1, 8)); first[..., 0] = first[..., 4] = 1
    last = np.zeros((8, 1, 1, 1)); last[4] = 1
    return tm.TensorTrain([first] + [core.copy() for _ in range(L)] + [last]).squeeze()


def get_bpx_Qp_term(L, l):
    """Level-l term of Q' = M' C (derivative of the level-l hat functions)."""
    W, U, _, Y = _get_bpx_factors()
    first = np.eye(2 if l == 0 else 4)[:1].reshape(1, 1, 1, -1)
    last = np.array([1.0, 0.0]).reshape(2, 1, 1, 1)
    middle = [Y] * L if l == 0 else [W] * (l - 1) + [U] + [Y] * (L - l)
    return tm.TensorTrain([first] + [c.copy() for c in middle] + [last]).squeeze()


def _sum_terms(terms, zero):
    acc = zero
    for t in terms:
        acc = _rounded(acc + t, rel_error=1e-16)
    return acc


def get_bpx_Qp(L):
    return _sum_terms((get_bpx_Qp_term(L, l) for l in range(L + 1)), tm.zeros([[2, 2] for _ in range(L)]))


def get_bpx_Qt_term(L, l):
    """Level-l term of Q^T = C M^T (L2 projection onto the level-l hat functions), trailing core in the
    (mean, half difference) basis."""
    W, _, V, _ = _get_bpx_factors()
    first = np.eye(4)[:1].reshape(1, 1, 1, 4)
    last = np.zeros((2, 2, 2))
    last[:, 0, :] = np.array([[1, 1], [1, -1]]) / 2
    cores = [first] + [W / 2] * l + [V / 2] * (L - l) + [last.reshape(4, 1, 2, 1)]
    return tm.TensorTrain([c.copy() for c in cores]).squeeze()


def get_bpx_Qt(L):
    return _sum_terms((get_bpx_Qt_term(L, l) for l in range(L + 1)), tm.zeros([[2, 2] for _ in range(L)] + [[1, 2]]))


def get_laplace_bpx(L):
    """B = Q'^T Q' (x 2^-L through the per-core halving), rounded at 1e-15."""
    Qp = get_bpx_Qp(L)
    B = _compose(Qp.copy().transpose(), Qp)
    return _rounded(_scale_cores(B, 0.5, L), rel_error=1e-15)


def get_rhs_matrix_bpx(L):
    R = get_bpx_Qt(L)
    G = (_get_gram_matrix_legendre() / 4) @ _CORNER_TO_LEGENDRE
    R.tensors[-1] = (R.tensors[-1].reshape(-1, 2) @ G)[..., None]
    return R


def get_level_mapping_matrix_as_tt(L, l):
    """P_l: prolongation of the level-l hat coefficients to the finest grid."""
    half = np.zeros((2, 2, 1, 2))
    half[0, :, 0, 0], half[0, :, 0, 1] = [0.5, 1], [0, 0.5]
    half[1, :, 0, 0], half[1, :, 0, 1] = [0.5, 0], [1, 0.5]
    half /= np.sqrt(2)
    e0 = np.array([1.0, 0.0])
    cores = [e0[None, None, :]] + [_get_refinement_tensor()] * l + [half] * (L - l) + [e0[:, None, None]]
    return tm.TensorTrain([c.copy() for c in cores]).squeeze()


def get_bpx_preconditioner_by_sum(L):
    C = None
    for l in range(L + 1):
        P = get_level_mapping_matrix_as_tt(L, l)
        PP = _compose(P, P.copy().transpose())
        C = PP if C is None else _rounded(C + 2 ** (-l) * PP, ranks_new=8)
    return C


# ---- 2D nodal operators ----------------------------------------------------------------------------

def _kron_xy(A, B):
    """A acting on x times B acting on y, modes kept separate: (p P, ax, ay, bx, by, q Q)."""
    return A.copy().expand_dims([1, 3]) * B.copy().expand_dims([0, 2])


def build_2D_mass_matrix(L):
    m = build_mass_matrix_in_nodal_basis(L)
    return _kron_xy(m, m).reshape_mode_indices([4, 4])


def build_laplace_matrix_2D(L):
    A = get_laplace_matrix_as_tt(L)
    M = build_mass_matrix_in_nodal_basis(L)
    return _rounded((_kron_xy(A, M) + _kron_xy(M, A)).reshape_mode_indices([4, 4]), rel_error=1e-15)


def get_rhs_matrix_as_tt_2D(L):
    R = get_rhs_matrix_as_tt(L).squeeze()
    R2 = _kron_xy(R, R).reshape_mode_indices([[4, 4]] * L + [[2, 2]])
    _rounded(R2, rel_error=1e-15)
    R2.tensors[-1] = R2.tensors[-1].reshape(-1, 4, 1)
    return R2


def get_rhs_matrix_bpx_by_sum_2D(L):
    modes = [(4, 4) for _ in range(L)] + [(4,)]
    R = tm.zeros(modes)
    G = get_gram_matrix_as_tt(L, basis='corner') * 0.5
    for l in range(L + 1):
        term = _compose(get_bpx_Qt_term(L, l), G)
        term = _kron_xy(term, term).reshape_mode_indices(modes)
        R = _rounded(R + (2 ** l) * term, ranks_new=80, rel_error=1e-14)
    R.tensors[-1] = R.tensors[-1][:, None, :, :]
    return R.squeeze()


def get_bpx_preconditioner_by_sum_2D(L):
    C = tm.zeros([[4, 4] for _ in range(L)])
    for l in range(L + 1):
        P = get_level_mapping_matrix_as_tt(L, l)
        PP = _compose(P, P.copy().transpose())
        PP = _kron_xy(PP, PP).reshape_mode_indices([4, 4])
        C = _rounded(C + (0.5 ** l) * PP, ranks_new=64)
    return C


def get_laplace_bpx_2D_by_sum(L, max_rank=70):
    Qp = [get_bpx_Qp_term(L, l) for l in range(L + 1)]
    Qt = [get_bpx_Qt_term(L, l) for l in range(L + 1)]
    for l in range(1, L + 1):
        _scale_cores(Qt[l], 2, l)
    G = get_gram_matrix_as_tt(L)
    B = tm.zeros([[2, 2, 2, 2] for _ in range(L)])
    for lm in range(2 * L + 1):
        for l in range(max(lm - L, 0), min(lm, L) + 1):
            m = lm - l
            if m > l:
                continue
            stiff = _compose(Qp[l].copy().transpose(), Qp[m])
            mass = _compose(Qt[l], G, Qt[m].copy().transpose()).squeeze()
            term = _kron_xy(stiff, mass) + _kron_xy(mass, stiff)
            if m == l:
                term = 0.5 * term
            B = _rounded(B + term, ranks_new=max_rank, rel_error=1e-15)
    B.reshape_mode_indices([4, 4])
    B = B + B.copy().transpose()
    return _scale_cores(B, 0.5, len(B))


# ---- norms (hot-path callers) ----------------------------------------------------------------------

def _sqrt_gram_times_overlap(L):
    G = get_gram_matrix_as_tt(L)
    G.tensors[-1] = np.sqrt(G.tensors[-1])
    return _compose(G, get_overlap_matrix_as_tt(L))


def get_L2_norm_1D(u):
    """||u||^2_L2 of a nodal (hat-basis) train: || G^{1/2} M u ||^2 * 2^L  (solver.py:420-428)."""
    L = len(u)
    u = u.copy()
    u.tensors.append(np.ones([1, 1, 1]))
    v = _sqrt_gram_times_overlap(L) @ u
    return v.norm_squared() * (2 ** L)


def get_L2_norm_2D(u):
    L = len(u)
    u = u.copy().reshape_mode_indices([4])
    u.tensors.append(np.ones([1, 1, 1]))
    GM = _sqrt_gram_times_overlap(L)
    GM = _kron_xy(GM, GM).reshape_mode_indices([(4, 4) for _ in range(L)] + [(4, 1)])
    v = GM @ u
    return v.norm_squared() * (4 ** L)


# ---- solvers ---------------------------------------------------------------------------------------

def solve_with_grad_descent(A, b, n_steps_max=200, lr=0.8, max_rank=20, print_steps=False, recalc_residual_every_n=10,
                            rel_accuracy=1e-14):
    """Truncated steepest descent (solver.py:251-275) for one right-hand side, on the GPU.
    Returns (x, r2) like the reference."""
    xs, r2, _ = solve_with_grad_descent_batched(A, [b], n_steps_max=n_steps_max, lr=lr, max_rank=max_rank,
                                                recalc_residual_every_n=recalc_residual_every_n, rel_accuracy=rel_accuracy)
    if print_steps:
        print(f"Final r2: {r2[0]:.2e}")
    return xs[0], float(r2[0])


def solve_with_grad_descent_batched(A, bs, n_steps_max=200, lr=0.8, max_rank=20, recalc_residual_every_n=10,
                                    rel_accuracy=1e-14, device=None, return_device=False):
    """Same iteration for a list of right-hand sides sharing the operator A; one launch sequence for all.
    Returns (list of TensorTrain, r2 array, steps array)."""
    from .batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine(device)
    op = A if isinstance(A, DeviceMPO) else DeviceMPO(A.tensors, eng.device)
    B = bs if isinstance(bs, BatchedTT) else BatchedTT.from_host([b.tensors for b in bs], eng.device)
    x, r2, steps, _ = eng.gd_solve(op, B, n_steps=n_steps_max, lr=lr, max_rank=max_rank, recalc_every=recalc_residual_every_n,
                                   rel_accuracy=rel_accuracy)
    if return_device:
        return x, r2, steps
    return [tm.TensorTrain(c) for c in x.to_host()], r2.cpu().numpy(), steps.cpu().numpy()


def solve_with_cg(A, b, n_steps_max=100, max_rank=20, print_steps=False, recalc_residual_every_n=10, rel_accuracy=1e-14):
    """Truncated TT conjugate gradients for a symmetric positive definite MPO, on the GPU.  Not part of the reference (it
    hands the linear solve to ttpy's AMEn, solver.py:236-248): same calling convention and stopping rule as
    `solve_with_grad_descent`, about half the iterations on the BPX-preconditioned systems.  Returns (x, r2)."""
    xs, r2, _ = solve_with_cg_batched(A, [b], n_steps_max=n_steps_max, max_rank=max_rank,
                                      recalc_residual_every_n=recalc_residual_every_n, rel_accuracy=rel_accuracy)
    if print_steps:
        print(f"Final r2: {r2[0]:.2e}")
    return xs[0], float(r2[0])


def solve_with_cg_batched(A, bs, n_steps_max=100, max_rank=20, recalc_residual_every_n=10, rel_accuracy=1e-14, device=None,
                          return_device=False):
    """`solve_with_cg` for a list of right-hand sides sharing the operator A.  Returns (list of TensorTrain, r2, steps)."""
    from .batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine(device)
    op = A if isinstance(A, DeviceMPO) else DeviceMPO(A.tensors, eng.device)
    B = bs if isinstance(bs, BatchedTT) else BatchedTT.from_host([b.tensors for b in bs], eng.device)
    x, r2, steps, _ = eng.cg_solve(op, B, n_steps=n_steps_max, max_rank=max_rank, recalc_every=recalc_residual_every_n,
                                   rel_accuracy=rel_accuracy)
    if return_device:
        return x, r2, steps
    return [tm.TensorTrain(c) for c in x.to_host()], r2.cpu().numpy(), steps.cpu().numpy()


def estimate_condition_number(A, max_rank=16, n_iter=300, rel_eps=1e-14, tol=1e-12, seed=0, return_trace=False):
    """Condition number lambda_max / lambda_min of a symmetric positive definite TT operator by TT power iteration,
    device resident (BASELINE.json config 4).  The reference has no such routine - it takes dense SVDs of
    `A.evalm()` for L <= 5 (2D) / L <= 11 (1D) (experiments/*_analyze_condition_number.py) - so parity exists only where
    the dense values exist; beyond that this is an estimate whose accuracy is limited by the rank cap.

    lambda_max: x <- round(A x) / ||.||, Rayleigh quotient <x, A x>.
    lambda_min: the same iteration on (lambda_max I - A), i.e. y <- round(lambda_max y - A y) / ||.||.
    Every product is a term of a fused rounding; the iterate never leaves the GPU."""
    import torch
    from .batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine()
    op = A if isinstance(A, DeviceMPO) else DeviceMPO(A.tensors, eng.device)
    L = op.L
    rng = np.random.default_rng(seed)
    modes = [(n,) for n in op.n_in]
    x0 = [rng.standard_normal((1 if k == 0 else 2, op.n_in[k], 1 if k == L - 1 else 2)) for k in range(L)]
    trace = []

    def normalise(x):
        n2 = eng.norm2(x)
        x.cores[0].mul_((1.0 / torch.sqrt(n2)).repeat_interleave(x.slot[0]))
        return x

    def iterate(shift):
        x = normalise(BatchedTT.from_host([x0], eng.device))
        lam_prev, lam = None, None
        for it in range(n_iter):
            terms = [(op, x)] if shift is None else [(None, x, None, shift), (op, x, None, -1.0)]
            y = eng.round_sum(terms, caps=max_rank, rel_eps=rel_eps)
            lam = float(eng.inner(x, y).cpu()[0])                 # Rayleigh quotient of the current unit iterate
            trace.append((shift is not None, it, lam))
            x = normalise(y)
            if lam_prev is not None and abs(lam - lam_prev) <= tol * abs(lam):
                break
            lam_prev = lam
        return lam

    lam_max = iterate(None)
    lam_min = lam_max - iterate(lam_max)
    cond = lam_max / lam_min
    if return_trace:
        return cond, lam_max, lam_min, trace
    return cond


def estimate_condition_number_krylov(A, max_rank=8, n_vectors=30, product_rank=None, rel_eps=1e-12, seed=0, return_trace=False):
    """Condition number of a symmetric positive definite TT operator by Rayleigh-Ritz on a truncated TT Krylov space
    (BASELINE.json config 4, "TT power / Lanczos iteration"; no reference counterpart beyond dense SVDs at L <= 5).

    q_0 random, then for k = 0, 1, ..: w_k = round(A q_k) at cap `product_rank` (default 4 x max_rank),
    q_{k+1} = round(w_k - sum_i <q_i, w_k> q_i) / ||.|| at cap `max_rank` (full orthogonalisation: the roundings break the
    three-term recurrence).  The extreme eigenvalues of the pencil (H, G), H_ij = <q_i, w_j> symmetrised, G_ij = <q_i, q_j>,
    are Ritz values: lambda_max is approached from below and lambda_min from above, so the returned ratio is a lower bound
    up to the truncation of the w_k.  Converges like conjugate gradients (tens of vectors) where the plain power
    iteration of `estimate_condition_number` needs hundreds of steps.  Every product / rounding / inner product runs on
    the GPU; only the n x n pencil is diagonalised on the host."""
    import torch
    from scipy.linalg import eigh
    from .batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine()
    op = A if isinstance(A, DeviceMPO) else DeviceMPO(A.tensors, eng.device)
    L = op.L
    product_rank = product_rank or 4 * max_rank
    rng = np.random.default_rng(seed)
    x0 = [rng.standard_normal((1 if k == 0 else 2, op.n_in[k], 1 if k == L - 1 else 2)) for k in range(L)]

    def normalised(x):
        n2 = eng.norm2(x)
        x.cores[0].mul_((1.0 / torch.sqrt(n2)).repeat_interleave(x.slot[0]))
        return x

    def dot(a, b):
        return float(eng.inner(a, b).cpu()[0])

    Q = [normalised(BatchedTT.from_host([x0], eng.device))]
    W, trace = [], []
    H = np.zeros((n_vectors, n_vectors))
    G = np.zeros((n_vectors, n_vectors))

    def ritz(m):
        g, h = G[:m, :m], 0.5 * (H[:m, :m] + H[:m, :m].T)
        ev, U = np.linalg.eigh(g)
        keep = ev > 1e-10 * ev.max()                       # directions the roundings made (nearly) dependent are dropped
        P = U[:, keep] / np.sqrt(ev[keep])
        lam = np.linalg.eigvalsh(P.T @ h @ P)
        return float(lam[-1]), float(lam[0])

    for k in range(n_vectors):
        w = eng.round_sum([(op, Q[k])], caps=product_rank, rel_eps=rel_eps)
        W.append(w)
        for i in range(k + 1):
            H[i, k] = dot(Q[i], w)
            if i < k:
                H[k, i] = dot(Q[k], W[i])
            G[i, k] = G[k, i] = dot(Q[i], Q[k])
        lmax, lmin = ritz(k + 1)
        trace.append((k + 1, lmax, lmin))
        if k + 1 == n_vectors:
            break
        coefs = [torch.full((1,), -H[i, k], dtype=torch.float64, device=eng.device) for i in range(k + 1)]
        terms = [(None, w)] + [(None, Q[i], coefs[i]) for i in range(k + 1)]
        nxt = eng.round_sum(terms, caps=max_rank, rel_eps=rel_eps)
        if float(eng.norm2(nxt).cpu()[0]) <= 1e-24 * max(abs(H[k, k]), 1e-300) ** 2:
            break                                           # invariant subspace reached
        Q.append(normalised(nxt))
    lmax, lmin = trace[-1][1], trace[-1][2]
    if return_trace:
        return lmax / lmin, lmax, lmin, trace
    return lmax / lmin


def _stack_instances(batches, picks):
    """One BatchedTT whose instance i is instance picks[i][1] of batches[picks[i][0]] (device-side gather; slots padded to the
    widest source)."""
    import torch
    from .batched import BatchedTT
    b0 = batches[0]
    L, n = b0.L, len(picks)
    slot = [max(b.slot[k] for b in batches) for k in range(L)]
    out = BatchedTT(b0.mode_shapes, n, slot, b0.device)
    for k in range(L):
        dst = out.cores[k].view(n, -1)
        for i, (bi, j) in enumerate(picks):
            src = batches[bi]
            dst[i, :src.slot[k]] = src.cores[k].view(src.N, -1)[j]
    out.ranks.copy_(torch.stack([batches[bi].ranks[j] for bi, j in picks]))
    return out


def estimate_condition_number_lobpcg(A, max_rank=16, n_iter=60, product_rank=None, rel_eps=1e-13, tol=1e-8, seed=0,
                                     return_trace=False):
    """Condition number of a symmetric positive definite TT operator by a locally optimal block iteration (LOBPCG without a
    preconditioner) for BOTH ends of the spectrum at once - BASELINE.json config 4; no reference counterpart beyond the dense
    SVDs at L <= 5 (experiments/2D_analyze_condition_number.py), which pin it there.

    State: the current Ritz vectors x_max, x_min and the previous search directions p_max, p_min, all rank-capped trains.
    One iteration: A x (a fused apply + round of a 2-instance batch), residuals r = A x - theta x, A [r; p] (one 4-instance
    batch), then Rayleigh-Ritz on span{x_max, x_min, r_max, r_min, p_max, p_min}: the 6 x 6 pencil (S^T A S, S^T S) is assembled
    from batched device inner products and diagonalised on the host; the new x and p are rounded linear combinations.  Unlike a
    Krylov basis, every vector the iteration keeps is close to an eigenvector or to its residual, so the rank cap binds far
    less (at L = 15, cap 8, the truncated Krylov space of `estimate_condition_number_krylov` stalls at 7.56).  Both Ritz
    values are Rayleigh quotients of actual vectors: lambda_max is approached from below and lambda_min from above, i.e. the
    returned ratio is a lower bound that rises with the cap.  Returns cond (and lambda_max, lambda_min, trace)."""
    import torch
    from .batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine()
    op = A if isinstance(A, DeviceMPO) else DeviceMPO(A.tensors, eng.device)
    L = op.L
    product_rank = product_rank or 2 * max_rank           # the products are rounded at twice the cap of the iterates: nearly free (the
    rng = np.random.default_rng(seed)                     # cost of an apply + round is set by the rank of its input), halves their bias

    def random_tt():
        return [rng.standard_normal((1 if k == 0 else 2, op.n_in[k], 1 if k == L - 1 else 2)) for k in range(L)]

    def scaled(x, fac):                                    # x_i * fac_i (device vector), in place on core 0
        x.cores[0].mul_(fac.repeat_interleave(x.slot[0]))
        return x

    def normalised(x):
        return scaled(x, 1.0 / torch.sqrt(eng.norm2(x)))

    def gram(S, AS):
        m = S.N
        pairs = [(i, j) for i in range(m) for j in range(i, m)]
        left = _stack_instances([S], [(0, i) for i, _ in pairs])
        g = eng.inner(left, _stack_instances([S], [(0, j) for _, j in pairs])).cpu().numpy()
        h = eng.inner(left, _stack_instances([AS], [(0, j) for _, j in pairs])).cpu().numpy()
        h2 = eng.inner(_stack_instances([AS], [(0, i) for i, _ in pairs]), _stack_instances([S], [(0, j) for _, j in pairs])).cpu().numpy()
        G, H = np.zeros((m, m)), np.zeros((m, m))
        for (i, j), gv, hv, hw in zip(pairs, g, h, h2):
            G[i, j] = G[j, i] = gv
            H[i, j] = H[j, i] = 0.5 * (hv + hw)                # <s_i, A s_j> symmetrised over the two rounded products
        return G, H

    X = normalised(BatchedTT.from_host([random_tt(), random_tt()], eng.device))
    P, trace = None, []
    lam_prev = None
    for it in range(n_iter):
        AX = eng.round_sum([(op, X)], caps=product_rank, rel_eps=rel_eps)
        theta = eng.inner(X, AX) / eng.norm2(X)
        R = eng.round_sum([(None, AX), (None, X, -theta)], caps=max_rank, rel_eps=rel_eps)
        rn = torch.sqrt(eng.norm2(R))
        R = scaled(R, 1.0 / torch.clamp(rn, min=1e-300))
        rest = R if P is None else _stack_instances([R, P], [(0, 0), (0, 1), (1, 0), (1, 1)])
        A_rest = eng.round_sum([(op, rest)], caps=product_rank, rel_eps=rel_eps)
        S = _stack_instances([X, rest], [(0, 0), (0, 1)] + [(1, i) for i in range(rest.N)])
        AS = _stack_instances([AX, A_rest], [(0, 0), (0, 1)] + [(1, i) for i in range(rest.N)])
        G, H = gram(S, AS)
        ev, U = np.linalg.eigh(G)
        keep = ev > 1e-12 * ev.max()                           # directions the iteration made (nearly) dependent are dropped
        W = U[:, keep] / np.sqrt(ev[keep])
        lam, Y = np.linalg.eigh(W.T @ H @ W)
        C = W @ Y                                              # columns: coefficient vectors of the Ritz vectors over S
        lam_max, lam_min = float(lam[-1]), float(lam[0])
        trace.append((it, lam_max, lam_min, float(rn[0]) / max(abs(lam_max), 1e-300), float(rn[1]) / max(abs(lam_min), 1e-300)))
        m = S.N
        newX, newP = [], []
        for col in (C[:, -1], C[:, 0]):
            coef = [torch.full((1,), float(c), dtype=torch.float64, device=eng.device) for c in col]
            singles = [_stack_instances([S], [(0, i)]) for i in range(m)]
            newX.append(eng.round_sum([(None, singles[i], coef[i]) for i in range(m)], caps=max_rank, rel_eps=rel_eps))
            newP.append(eng.round_sum([(None, singles[i], coef[i]) for i in range(2, m)], caps=max_rank, rel_eps=rel_eps))
        X = normalised(_stack_instances(newX, [(0, 0), (1, 0)]))
        P = normalised(_stack_instances(newP, [(0, 0), (1, 0)]))
        if lam_prev is not None and abs(lam_max - lam_prev[0]) <= tol * abs(lam_max) and abs(lam_min - lam_prev[1]) <= tol * abs(lam_min):
            break
        lam_prev = (lam_max, lam_min)
    cond = lam_max / lam_min
    if return_trace:
        return cond, lam_max, lam_min, trace
    return cond


def solve_with_amen(A, b, **solver_options):
    """Signature of solver.py:236-248.  With ttpy installed: the reference's AMEn call, unchanged in meaning.  Without it
    (this image) the linear solve runs on the GPU to the same `eps`: truncated conjugate gradients by default (every system
    the reference passes here is symmetric positive definite), `method='gd'` selects the reference's own steepest descent;
    at most `nswp` x 50 steps, rank cap `max_rank` (default 60, the drivers' default).  ttpy-specific options are ignored."""
    try:
        import tt
        from tt.amen import amen_solve
    except ImportError:
        eps = solver_options.get('eps', 1e-15)
        steps = int(solver_options.get('nswp', 40)) * 50
        solve = solve_with_grad_descent if solver_options.get('method', 'cg') == 'gd' else solve_with_cg
        x, r2 = solve(A, b, n_steps_max=steps, max_rank=int(solver_options.get('max_rank', 60)), rel_accuracy=max(eps, 1e-15) ** 2)
        try:                                        # tell the caller when the requested accuracy was not reached (AMEn would sweep on)
            rel = float(r2) / max(float(b.norm_squared()), 1e-300)
            if rel > max(eps, 1e-15) ** 2 * 1e4:
                import warnings
                warnings.warn(f"solve_with_amen (device {'GD' if solve is solve_with_grad_descent else 'CG'} in place of ttpy): relative "
                              f"residual ||b - A x||^2 / ||b||^2 = {rel:.3e} after at most {steps} steps at rank cap "
                              f"{int(solver_options.get('max_rank', 60))}; the requested eps^2 = {max(eps, 1e-15) ** 2:.1e} is below what FP64 "
                              "with rank truncation reaches for this system", RuntimeWarning)
        except Exception:
            pass
        return x
    opts = dict(eps=1e-15, nswp=40, local_iters=2, verb=1)
    opts.update(solver_options)
    opts.pop('max_rank', None)
    opts.pop('method', None)                        # selects the device solver above; not a ttpy option
    eps = opts.pop('eps')
    x = amen_solve(tt.matrix.from_list(A.to_ttpylist()), tt.vector.from_list(b.to_ttpylist()), tt.ones(2, len(A)), eps, **opts)
    return tm.TensorTrain.from_ttpylist(tt.vector.to_list(x))


def solve_PDE_1D(f, **solver_options):
    L = len(f) - 1
    g = (get_rhs_matrix_as_tt(L) @ f).squeeze()
    u = solve_with_amen(get_laplace_matrix_as_tt(L), g, **solver_options)
    Du = (get_derivative_matrix_as_tt(L) @ u).reapprox(rel_error=1e-15)
    return u, Du


def solve_PDE_2D(f: tm.TensorTrain, max_rank=60, **solver_options):
    f = f.copy().flatten_mode_indices()
    L = len(f) - 1
    g = (get_rhs_matrix_as_tt_2D(L) @ f).squeeze().reapprox(max_rank)
    u = solve_with_amen(build_laplace_matrix_2D(L), g, **solver_options)
    u.reapprox(rel_error=1e-14)
    ue = u.copy()
    ue.tensors.append(np.ones([1, 1, 1, 1]))
    D = get_derivative_matrix_as_tt(L)
    D.tensors.append(np.ones([1, 1, 1, 1]))
    M = get_overlap_matrix_as_tt(L)
    DuDx = (kronecker_prod_2D(D, M) @ ue).flatten_mode_indices()
    DuDy = (kronecker_prod_2D(M, D) @ ue).flatten_mode_indices()
    return u, DuDx, DuDy


def solve_PDE_1D_with_preconditioner(f, max_rank=60, **solver_options):
    eps = 1e-15
    L = len(f) - 1
    b = (get_rhs_matrix_bpx(L) @ f).squeeze()
    _scale_cores(b, 0.5, L)
    B = get_laplace_bpx(L)
    v = solve_with_amen(B, b, **solver_options)
    v.reapprox(rel_error=eps)
    u = (get_bpx_preconditioner(L) @ v).reapprox(rel_error=eps)
    Du = (get_bpx_Qp(L) @ v).reapprox(rel_error=eps)
    return u, Du


_operator_cache = {}


def _cached_laplace_BPX_2D(L, rel_error, on_device=True):
    """`get_laplace_BPX_2D(L).reapprox(rel_error)` (solver.py:218-220), the one expensive rounding of the pipeline
    (rank 1152 -> 161 at L = 12), cached per (L, eps).  By default it runs on the GPU like every other `reapprox`
    (cooperative QR leaf + multi-CTA Jacobi: 4.7 s at L = 12 against 12-22 s for numpy); `on_device=False` is the
    construction-time host route used by the CPU reference arm of bench.py and the CPU tests."""
    key = (L, float(rel_error), bool(on_device))
    if key not in _operator_cache:
        B = get_laplace_BPX_2D(L)
        _operator_cache[key] = B.reapprox(rel_error=rel_error) if on_device else _rounded(B, rel_error=rel_error)
    return _operator_cache[key].copy()


def solve_PDE_2D_with_preconditioner(f, max_rank=60, **solver_options):
    L = len(f) - 1
    f = f.copy().flatten_mode_indices()
    eps = solver_options.get('eps', 1e-14)
    B = _cached_laplace_BPX_2D(L, eps)
    b = (get_rhs_matrix_BPX_2D(L) @ f).squeeze().reapprox(ranks_new=max_rank, rel_error=eps)
    v = solve_with_amen(B, b, **solver_options)
    v.reapprox(ranks_new=max_rank, rel_error=eps)
    u = (get_BPX_preconditioner_2D(L) @ v).reapprox(ranks_new=max_rank, rel_error=eps)
    v.tensors.append(np.ones([1, 1, 1, 1]))
    Dx = (get_BPX_Qp_2D(L, 'x') @ v).flatten_mode_indices()
    Dy = (get_BPX_Qp_2D(L, 'y') @ v).flatten_mode_indices()
    return u, Dx, Dy
