"""Assembly-time host algebra for operator construction (NOT the hot path).

The reference builds its operators (stiffness, BPX preconditioner, right-hand-side maps) once per level
count L on the host from cores of rank <= 32, using core-wise products and a few roundings
(solver.py:84-98, 334-396; bpx2D.py:91-182).  SURVEY.md section 2 marks those builders "API to keep,
out of scope for CUDA" and section 8(f).1 lists moving their one large rounding to the device as a
later step.  This module is what the builders in solver.py / bpx2D.py / utils.py use for that
construction-time arithmetic.  It is deliberately private and is never reached from
`TensorTrain.__matmul__`, `.reapprox`, `.left_orthogonalize`, `.norm_squared` or the batched solver:
those always run on the GPU through csrc/libqtt_b200.so and raise if it is unavailable.
"""
import numpy as np
import scipy.linalg as sla


def compose(a_cores, b_cores):
    """Core-wise operator composition  W_k[(pP), i.., j.., (qQ)] = sum_m U_k[p, i.., m, q] V_k[P, m, j.., Q]."""
    assert len(a_cores) == len(b_cores)
    out = []
    for U, V in zip(a_cores, b_cores):
        if U.ndim not in (3, 4) or V.ndim not in (3, 4):
            raise NotImplementedError("Currently only supports 1D or 2D mode dimensions")
        lead, trail = U.shape[1:-2], V.shape[2:-1]
        W = np.tensordot(U, V, axes=[U.ndim - 2, 1])                # p, lead.., q, P, trail.., Q
        nl, nt = len(lead), len(trail)
        perm = [0, 1 + nl + 1] + list(range(1, 1 + nl)) + list(range(3 + nl, 3 + nl + nt)) + [1 + nl, 3 + nl + nt]
        W = W.transpose(perm)
        out.append(np.ascontiguousarray(W).reshape((U.shape[0] * V.shape[0],) + lead + trail + (U.shape[-1] * V.shape[-1],)))
    return out


def direct_sum(a_cores, b_cores):
    L = len(a_cores)
    assert len(b_cores) == L
    out = []
    for k, (U, V) in enumerate(zip(a_cores, b_cores)):
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


def kept_rank(sigma, eps):
    """The reference's rule (tensormethods.py:5-12): keep sigma_i while (sigma_i/sigma_0)^2 >= eps^2."""
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = (sigma * sigma) / (sigma[0] * sigma[0])
    return int(sigma.size - np.count_nonzero(ratio < eps * eps))


def sweep_right_to_left(cores):
    """Householder orthogonalisation sweep (rows of cores 1..L-1 orthonormal), in place."""
    carry = cores[-1]
    for k in range(len(cores) - 1, 0, -1):
        r = carry.shape[0]
        Q, R = sla.qr(carry.reshape(r, -1).T, mode="economic")
        cores[k] = np.ascontiguousarray(Q.T).reshape((Q.shape[1],) + carry.shape[1:])
        carry = np.tensordot(cores[k - 1], R.T, axes=[-1, 0])
    cores[0] = carry
    return cores


def round_(cores, ranks_new=None, rel_error=1e-16):
    """TT rounding with the reference's sweep directions and rank rule, in place."""
    sweep_right_to_left(cores)
    L = len(cores)
    if ranks_new is None:
        ranks_new = [c.shape[-1] for c in cores[:-1]]
    elif isinstance(ranks_new, (int, np.integer)):
        ranks_new = [int(ranks_new)] * (L - 1)
    carry = cores[0]
    for k in range(L - 1):
        q = carry.shape[-1]
        mat = carry.reshape(-1, q)
        Um, S, Vt = sla.svd(mat, full_matrices=False, lapack_driver="gesdd")
        keep = min(mat.shape[0], q, int(ranks_new[k]), kept_rank(S, rel_error))
        cores[k] = np.ascontiguousarray(Um[:, :keep]).reshape(carry.shape[:-1] + (keep,))
        carry = np.tensordot(S[:keep, None] * Vt[:keep], cores[k + 1], axes=[1, 0])
    cores[-1] = carry
    return cores


def norm_squared(cores):
    c = sweep_right_to_left([np.array(x, copy=True) for x in cores])
    return float(np.sum(c[0] ** 2))
