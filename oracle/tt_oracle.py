"""CPU oracle for the tensor-train hot path.  TEST INFRASTRUCTURE ONLY.

A numpy restatement of the algorithms of the reference's `laplace_mps.tensormethods` and
`laplace_mps.solver.solve_with_grad_descent`, written as free functions over plain lists of float64
cores `(r_left, *modes, r_right)`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it - as the checker or the timed CPU baseline,
never as part of the product (`laplacesolvermps_b200` does not import `oracle`).

Pinning: tests/test_oracle_golden.py checks every function below against fixtures generated from the
unmodified reference (oracle/make_golden.py -> tests/golden/*.npz) and, when /root/reference is
mounted, against the live reference.  Parity is therefore PINNED for everything except the AMEn
solver (third-party `ttpy`, un-vendored and un-pinned: "parity unpinned").

Each function cites the reference lines whose behaviour it restates.
"""
import numpy as np

# ----------------------------------------------------------------------------------------------
# a1  container helpers                                              tensormethods.py:17-48,77-78
# ----------------------------------------------------------------------------------------------

def bond_ranks(cores):
    """Right bond of every core but the last (tensormethods.py:30-32)."""
    return [int(c.shape[-1]) for c in cores[:-1]]


def clone(cores):
    return [np.array(c, dtype=np.float64, copy=True) for c in cores]


def zeros_like_modes(mode_sizes):
    """Rank-1 all-zero TT (tensormethods.py:14-15)."""
    return [np.zeros((1,) + tuple(m) + (1,)) for m in mode_sizes]


# ----------------------------------------------------------------------------------------------
# a2  core-wise product (MPO x MPS and the three other ndim pairings)   tensormethods.py:176-194
# ----------------------------------------------------------------------------------------------

def matmul(a, b):
    """W_k[(pP), row.., col.., (qQ)] = sum_m U_k[p, row.., m, q] V_k[P, m, col.., Q].

    The contracted index is the last mode of U and the first mode of V; the left operand's bond is
    the slow half of the fused bond.  Only 3-D / 4-D cores are defined
    (NotImplementedError otherwise, tensormethods.py:190).
    """
    if len(a) != len(b):
        raise AssertionError("length mismatch")
    out = []
    for U, V in zip(a, b):
        if U.ndim not in (3, 4) or V.ndim not in (3, 4):
            raise NotImplementedError("Currently only supports 1D or 2D mode dimensions")
        p, q = U.shape[0], U.shape[-1]
        P, Q = V.shape[0], V.shape[-1]
        m = U.shape[-2]
        if V.shape[1] != m:
            raise ValueError("contracted mode sizes differ")
        rows = U.shape[1:-2]            # () for a 3-D U, (n_out,) for a 4-D U
        cols = V.shape[2:-1]            # () for a 3-D V, (n_in,)  for a 4-D V
        U2 = U.reshape(p, int(np.prod(rows, dtype=int)), m, q)
        V2 = V.reshape(P, m, int(np.prod(cols, dtype=int)), Q)
        W = np.einsum("pamq,PmcQ->pPacqQ", U2, V2, optimize=True)
        out.append(np.ascontiguousarray(W).reshape((p * P,) + rows + cols + (q * Q,)))
    return out


# ----------------------------------------------------------------------------------------------
# a3  orthogonalisation sweep, right -> left                              tensormethods.py:90-99
# ----------------------------------------------------------------------------------------------

def left_orthogonalize(cores):
    """In place.  For k = L-1 .. 1: thin QR of core_k viewed as (n*q_{k+1}) x p_k, orthonormal rows
    stay in core_k, the triangular factor is pushed into core_{k-1}; the non-orthogonal core ends at
    position 0 (the reference's method name notwithstanding)."""
    cur = cores[-1]
    for k in range(len(cores) - 1, 0, -1):
        left = cur.shape[0]
        Qm, Rm = np.linalg.qr(cur.reshape(left, -1).T)          # (n q) x p  =  Qm Rm
        cores[k] = Qm.T.reshape((Qm.shape[1],) + cur.shape[1:])
        cur = np.tensordot(cores[k - 1], Rm.T, axes=[-1, 0])
    cores[0] = cur
    return cores


def right_orthogonalize(cores):
    """Mirror sweep (tensormethods.py:81-88; unused by the hot path, kept for API parity)."""
    cur = cores[0]
    for k in range(len(cores) - 1):
        r = cur.shape[-1]
        Qm, Rm = np.linalg.qr(cur.reshape(-1, r))
        cores[k] = Qm.reshape(cur.shape[:-1] + (Qm.shape[1],))
        cur = np.tensordot(Rm, cores[k + 1], axes=[1, 0])
    cores[-1] = cur
    return cores


# ----------------------------------------------------------------------------------------------
# a4  truncation sweep, left -> right, and the rank rule         tensormethods.py:5-12, 101-133
# ----------------------------------------------------------------------------------------------

def required_rank(sigma, eps):
    """Number of singular values kept: every sigma_i with (sigma_i / sigma_0)^2 >= eps^2.
    NOT a cumulative tail bound (tensormethods.py:5-12).  sigma_0 == 0 gives nan ratios, every
    comparison is False and nothing is removed - same as the reference."""
    sigma = np.asarray(sigma, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = sigma ** 2 / sigma[0] ** 2
    return int(len(sigma) - np.count_nonzero(rel < eps ** 2))


def reapprox(cores, ranks_new=None, rel_error=1e-16, orthogonalize=True, return_sigma=False):
    """In place TT rounding: a3, then for k = 0 .. L-2 an SVD of the (s_{k-1} n) x q_k unfolding,
    s_k = min(rows, q_k, cap_k, required_rank), core_k <- U[:, :s_k], diag(S) V^T pushed right."""
    if orthogonalize:
        left_orthogonalize(cores)
    L = len(cores)
    if ranks_new is None:
        ranks_new = bond_ranks(cores)
    if isinstance(ranks_new, (int, np.integer)):
        ranks_new = [int(ranks_new)] * (L - 1)
    sigmas = []
    cur = cores[0]
    for k in range(L - 1):
        q = cur.shape[-1]
        mat = cur.reshape(-1, q)
        Um, S, Vt = np.linalg.svd(mat, full_matrices=False)
        sigmas.append(S.copy())
        keep = min(mat.shape[0], q, int(ranks_new[k]), required_rank(S, rel_error))
        cores[k] = Um[:, :keep].reshape(cur.shape[:-1] + (keep,))
        carry = S[:keep, None] * Vt[:keep, :]
        cur = np.tensordot(carry, cores[k + 1], axes=[1, 0])
    cores[-1] = cur
    if return_sigma:
        return cores, sigmas
    return cores


# ----------------------------------------------------------------------------------------------
# a5 / a6  norm and inner product                         tensormethods.py:280-283, 233-264
# ----------------------------------------------------------------------------------------------

def norm_squared(cores):
    """Copy, orthogonalise, sum of squares of core 0 (tensormethods.py:280-283)."""
    c = left_orthogonalize(clone(cores))
    return float(np.sum(c[0] ** 2))


def squeeze(cores):
    """Fold every core whose modes all have size 1 into its neighbour (tensormethods.py:251-264)."""
    out = [cores[0]]
    for U in cores[1:]:
        last = out[-1]
        if int(np.prod(last.shape[1:-1], dtype=int)) == 1:
            out[-1] = np.tensordot(last.reshape(last.shape[0], last.shape[-1]), U, axes=[-1, 0])
        elif int(np.prod(U.shape[1:-1], dtype=int)) == 1:
            out[-1] = np.tensordot(last, U.reshape(U.shape[0], U.shape[-1]), axes=[-1, 0])
        else:
            out.append(U)
    return out


def dense(cores, squeeze_result=True, reshape=None):
    """Contract the chain left to right (tensormethods.py:233-249)."""
    A = cores[0]
    for U in cores[1:]:
        A = np.tensordot(A, U, axes=(-1, 0))
    A = A.reshape(A.shape[1:-1])
    if squeeze_result:
        A = A.squeeze()
    if reshape == "vector":
        return A.flatten()
    if reshape == "matrix":
        L = len(cores)
        A = A.transpose(list(range(0, 2 * L, 2)) + list(range(1, 2 * L, 2)))
        return A.reshape(int(np.prod(A.shape[:L], dtype=int)), -1)
    return A


def inner(a, b):
    """<a, b> through the reference's idiom `(a @ b).squeeze().eval()`: explicit (r s) x (r s)
    transfer matrices multiplied left to right."""
    return float(dense(squeeze(matmul(a, b))))


# ----------------------------------------------------------------------------------------------
# a7  sums and scalings                                   tensormethods.py:135-174, 197-200
# ----------------------------------------------------------------------------------------------

def add(a, b):
    """Block-diagonal cores; first core concatenated along its right bond, last along its left."""
    if len(a) != len(b):
        raise AssertionError("length mismatch")
    L = len(a)
    out = []
    for k, (U, V) in enumerate(zip(a, b)):
        if k == 0:
            out.append(np.concatenate([U, V], axis=-1))
        elif k == L - 1:
            out.append(np.concatenate([U, V], axis=0))
        else:
            W = np.zeros((U.shape[0] + V.shape[0],) + U.shape[1:-1] + (U.shape[-1] + V.shape[-1],))
            W[:U.shape[0], ..., :U.shape[-1]] = U
            W[U.shape[0]:, ..., U.shape[-1]:] = V
            out.append(W)
    return out


def scale(a, alpha):
    """Scalar folded into core 0 (tensormethods.py:156-159, 169-174)."""
    out = clone(a)
    out[0] *= alpha
    return out


def sub(a, b):
    return add(a, scale(b, -1))


# ----------------------------------------------------------------------------------------------
# a8  truncated steepest descent                                           solver.py:251-275
# ----------------------------------------------------------------------------------------------

def solve_with_grad_descent(A, b, n_steps_max=200, lr=0.8, max_rank=20, recalc_residual_every_n=10,
                            rel_accuracy=1e-14, history=None):
    """Returns (x, r2).  Every rounding uses rel_error = 1e-16, i.e. the cap is the only effective
    truncation (solver.py:255-270 pass `ranks_new` only).  `history`, if a list, receives
    (step, r2, gamma) tuples - an oracle-only convenience for the parity tests."""
    x = zeros_like_modes([c.shape[1:-1] for c in b])
    r = None
    for n in range(n_steps_max):
        if n % recalc_residual_every_n == 0:
            r = reapprox(sub(b, matmul(A, x)), ranks_new=max_rank)
        Ar = reapprox(matmul(A, r), ranks_new=max_rank)
        r2 = norm_squared(r)
        if n > 10 and n % 10 == 0:
            x2 = norm_squared(x)
            if r2 / x2 <= rel_accuracy:
                break
        gamma = float(lr * r2 / inner(r, Ar))
        if history is not None:
            history.append((n, r2, gamma))
        x = reapprox(add(x, scale(r, gamma)), ranks_new=max_rank)
        r = reapprox(sub(r, scale(Ar, gamma)), ranks_new=max_rank)
    r = reapprox(sub(b, matmul(A, x)), ranks_new=max_rank)
    return x, norm_squared(r)


# ----------------------------------------------------------------------------------------------
# (f)3  truncated conjugate gradients - NOT in the reference ("parity unpinned")
# ----------------------------------------------------------------------------------------------

def solve_with_cg(A, b, n_steps_max=100, max_rank=20, recalc_residual_every_n=10, rel_accuracy=1e-14, history=None):
    """Truncated TT conjugate gradients for a symmetric positive definite MPO: the solver offered behind
    `solve_with_amen` as SURVEY section 8(f).3 asks.  There is NO reference implementation of it (the reference
    delegates to the third-party `ttpy` AMEn, un-vendored), so this restatement pins only that the CUDA path and the
    CPU path run the same algorithm; the answers are judged against dense / manufactured solutions in the tests.
    Built from the same three kernels as `solve_with_grad_descent` (apply + round, inner product, norm), with the
    same conventions: every rounding uses the cap only (rel_error 1e-16), the residual is recomputed from x every
    `recalc_residual_every_n` steps, the stopping test r2 / x2 <= rel_accuracy is evaluated at multiples of 10.
    Because the roundings perturb r and p, the coefficients are the perturbation-tolerant ones - exact line search
    alpha = <p, r> / <p, A p> and explicit A-orthogonalisation beta = -<r_new, A p> / <p, A p> - which keep every
    step an energy descent (the textbook r.r ratios stagnate once the cap binds).  Returns (x, ||b - A x||^2)."""
    x = zeros_like_modes([c.shape[1:-1] for c in b])
    r = reapprox(clone(b), ranks_new=max_rank)
    p = clone(r)
    for n in range(n_steps_max):
        Ap = reapprox(matmul(A, p), ranks_new=max_rank)
        pAp = inner(p, Ap)
        alpha = float(inner(p, r) / pAp) if pAp > 0.0 else 0.0
        if history is not None:
            history.append((n, norm_squared(r), alpha))
        x = reapprox(add(x, scale(p, alpha)), ranks_new=max_rank)
        if (n + 1) % recalc_residual_every_n == 0:
            r = reapprox(sub(b, matmul(A, x)), ranks_new=max_rank)
        else:
            r = reapprox(sub(r, scale(Ap, alpha)), ranks_new=max_rank)
        if n + 1 > 10 and (n + 1) % 10 == 0:
            if norm_squared(r) / norm_squared(x) <= rel_accuracy:
                break
        beta = float(-inner(r, Ap) / pAp) if pAp > 0.0 else 0.0
        p = reapprox(add(r, scale(p, beta)), ranks_new=max_rank)
    r = reapprox(sub(b, matmul(A, x)), ranks_new=max_rank)
    return x, norm_squared(r)
