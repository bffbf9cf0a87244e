"""Explicit QTT cores of the 2D BPX-preconditioned Laplace system (API of the reference's
`laplace_mps.bpx2D`, reference file src/laplace_mps/bpx2D.py; maths: project report section 2.6).

Construction-time host code (cores of rank <= 32, built once per L).  All operators are assembled from
a handful of elementary 1D "hat" cores with three core-level products:

    mode_product      (A B)        core-wise matrix product, bonds fused        [(pP), a, c, (qQ)]
    tensor_product    A (x) B      Kronecker product of all four indices (x-dimension slow, y fast)
    strong_kronecker  A |><| B     contraction over the shared bond, modes Kronecker-multiplied

Block structure of every level core (level-interleaved, mode index = 2 * x_bit + y_bit):

    [ Ub   (2^-l) * Ub   ]            [ Ub   Ub |><| W ]
    [ 0    (1/4)  * Xb   ]   for C,   [ 0    Z / 2     ]   for Q and Q'.
"""
import numpy as np

from . import tensormethods as tm
from . import _assembly as _asm
from .utils import kronecker_prod_2D, _get_gram_matrix_tt


# ---- core-level products -----------------------------------------------------------------------

def mode_product(A, B):
    return _asm.compose([A], [B])[0]


def transpose_core(A):
    return np.swapaxes(A, 1, 2)


def tensor_product(A, B):
    assert A.ndim == B.ndim
    C = np.kron(A, B) if A.ndim <= 2 else None
    if C is None:
        idx_a = tuple(x for pair in zip([slice(None)] * A.ndim, [None] * A.ndim) for x in pair)
        idx_b = tuple(x for pair in zip([None] * B.ndim, [slice(None)] * B.ndim) for x in pair)
        C = (A[idx_a] * B[idx_b]).reshape([sa * sb for sa, sb in zip(A.shape, B.shape)])
    return C


def strong_kronecker(A, B):
    C = np.tensordot(A, B, axes=[3, 0])                       # p a b A B r
    C = C.transpose(0, 1, 3, 2, 4, 5)                         # p a A b B r
    return C.reshape(A.shape[0], A.shape[1] * B.shape[1], A.shape[2] * B.shape[2], B.shape[3])


# ---- elementary 1D cores (rank_left, row, col, rank_right) -------------------------------------

def _get_U_hat():
    """Refinement core [[I, J^T], [0, J]] of the overlap / derivative matrices."""
    J = np.array([[0.0, 1.0], [0.0, 0.0]])
    U = np.zeros((2, 2, 2, 2))
    U[0, :, :, 0], U[0, :, :, 1], U[1, :, :, 1] = np.eye(2), J.T, J
    return U


def _get_X_hat():
    """Hat-function prolongation: half values of the coarse hat at the fine nodes."""
    X = np.array([[[1, 0], [2, 1]], [[1, 2], [0, 1]]], dtype=float)      # [rank_in, row, rank_out]
    return X[:, :, None, :] / 2


def _get_Y0_hat():
    Y = np.array([[[2, 0], [2, 0]], [[-1, 1], [1, 1]]], dtype=float)
    return Y[:, :, None, :] / 2


def _get_Y1_hat():
    return np.full((1, 2, 1, 1), 0.5)


def _get_N0_hat():
    return np.eye(2).reshape(2, 2, 1, 1) / 2


def _get_N1_hat():
    return np.ones((1, 1, 1, 1))


def _get_P_hat():
    return np.array([1.0, 0.0]).reshape(2, 1, 1, 1)


def _get_Pb_hat():
    return np.eye(4)[0].reshape(4, 1, 1, 1)


def _get_W0_hat():
    return np.kron(np.array([[1, 1], [1, -1]]), np.eye(2)).reshape(4, 1, 1, 4)


def _get_W1_hat():
    return np.kron(np.array([[1], [-1]]), np.eye(2)).reshape(4, 1, 1, 2)


def _get_Ab_hat():
    return np.eye(4)[0].reshape(1, 1, 1, 4)


def _get_Ub_hat():
    U = _get_U_hat()
    return mode_product(U, transpose_core(U))


def _get_Z0_hat():
    return mode_product(_get_Y0_hat(), transpose_core(_get_X_hat()))


def _get_Z1_hat():
    return mode_product(_get_Y1_hat(), transpose_core(_get_X_hat()))


def _get_K0_hat():
    return mode_product(_get_N0_hat(), _get_P_hat())


def _get_K1_hat():
    return mode_product(_get_N1_hat(), _get_P_hat())


def _both(core_x, core_y=None):
    return tensor_product(core_x, core_x if core_y is None else core_y)


def _block_chain(first, levels, last):
    return tm.TensorTrain([first] + levels + [last]).squeeze()


# ---- operators ---------------------------------------------------------------------------------

def get_BPX_preconditioner_2D(L):
    """C = sum_l 2^-l P_l P_l^T (x) P_l P_l^T, rank 32 = 2 * 16.  (The reference ships `2**(4+21)` for
    this rank - a typo for 2**(4+1), see SURVEY.md section 9 - which cannot be allocated.)"""
    X = _get_X_hat()
    Ub, Xb = _both(_get_Ub_hat()), _both(mode_product(X, transpose_core(X)))
    Ab, Pb = _both(_get_Ab_hat()), _both(_get_Pb_hat())
    half = Ub.shape[0]
    levels = []
    for l in range(L):
        core = np.zeros((2 * half, 4, 4, 2 * half))
        core[:half, :, :, :half] = Ub
        core[:half, :, :, half:] = 0.5 ** (l + 1) * Ub
        core[half:, :, :, half:] = 0.25 * Xb
        levels.append(core)
    return _block_chain(np.concatenate([Ab, Ab], axis=-1), levels, np.concatenate([np.zeros_like(Pb), Pb], axis=0))


def _q_chain(L, W, Z, K, zero_rows):
    Ub, Ab = _both(_get_Ub_hat()), _both(_get_Ab_hat())
    UW = strong_kronecker(Ub, W)
    r0, r1 = Ub.shape[0], Z.shape[0]
    core = np.zeros((r0 + r1, 4, 4, r0 + r1))
    core[:r0, :, :, :r0] = Ub
    core[:r0, :, :, r0:] = UW
    core[r0:, :, :, r0:] = Z / 2
    first = np.concatenate([Ab, strong_kronecker(Ab, W)], axis=-1)
    last = np.concatenate([np.zeros((r0,) + K.shape[1:]), K], axis=0)
    return core, first, last


def get_BPX_Qp_2D(L, deriv='x'):
    """Q' = (D (x) M) C^{1/2}-type factor mapping BPX coefficients to the x- (or y-) derivative, rank 24."""
    core, first, last = _q_chain(L, _both(_get_W1_hat(), _get_W0_hat()), _both(_get_Z1_hat(), _get_Z0_hat()),
                                 _both(_get_K1_hat(), _get_K0_hat()), 16)
    Qp = _block_chain(first, [core.copy() for _ in range(L)], last)
    if deriv == 'x':
        return Qp
    if deriv != 'y':
        raise ValueError("Unknown axis")
    # swap the roles of the two dimensions inside every mode index
    Qp.reshape_mode_indices([(2, 2, 2, 2) for _ in range(L)] + [(1, 2, 1, 1)])
    for k, U in enumerate(Qp):
        Qp.tensors[k] = np.transpose(U, [0, 2, 1, 4, 3, 5])
    return Qp.reshape_mode_indices([(4, 4) for _ in range(L)] + [(2, 1)])


def get_BPX_Q_2D(L):
    """Q = (M (x) M) C-type factor (no derivative), rank 32."""
    core, first, last = _q_chain(L, _both(_get_W0_hat()), _both(_get_Z0_hat()), _both(_get_K0_hat()), 16)
    return _block_chain(first, [core / 2 for _ in range(L)], last)


def get_laplace_BPX_2D(L):
    """B = Q'_x^T G_y Q'_x + Q'_y^T G_x Q'_y (unrounded: rank 2 * 24 * 24 = 1152)."""
    Qx, Qy = get_BPX_Qp_2D(L, 'x'), get_BPX_Qp_2D(L, 'y')
    eye = tm.TensorTrain([np.eye(2)[None, :, :, None] for _ in range(L)] + [np.eye(1)[None, :, :, None]])
    G = _get_gram_matrix_tt(L) * 2
    Gy, Gx = kronecker_prod_2D(eye, G), kronecker_prod_2D(G, eye)

    def sandwich(Q, Gm):
        Qt = Q.copy().transpose()
        return _asm.compose(_asm.compose(Qt.tensors, Gm.tensors), Q.tensors)

    B = tm.TensorTrain(_asm.direct_sum(sandwich(Qx, Gy), sandwich(Qy, Gx)))
    for k in range(len(B)):
        B.tensors[k] = B.tensors[k] / 4
    return B.squeeze()


def get_rhs_matrix_BPX_2D(L):
    """R = Q^T (G (x) G): maps the corner-basis coefficients of f to the BPX right-hand side, rank 32."""
    Qt = get_BPX_Q_2D(L).transpose()
    to_legendre = np.diag([2, 2 / 3]) @ np.array([[1, 1], [-1, 1]]) * 0.5
    G = tm.TensorTrain([np.eye(2)[None, :, :, None] for _ in range(L)] + [to_legendre.reshape(1, 2, 2, 1)])
    R = tm.TensorTrain(_asm.compose(Qt.tensors, kronecker_prod_2D(G, G).tensors))
    for k in range(len(R)):
        R.tensors[k] = R.tensors[k] / 4
    return R
