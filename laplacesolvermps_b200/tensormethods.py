"""Drop-in replacement of `laplace_mps.tensormethods` (reference: src/laplace_mps/tensormethods.py).

`TensorTrain` keeps the reference's data model - `self.tensors` is a Python list of C-contiguous float64
numpy cores `(r_left, *modes, r_right)` that callers may mutate, most methods work in place and return
`self` - but the arithmetic of the hot path runs on the GPU through csrc/libqtt_b200.so:

    `a @ b`               -> qtt_matmul_batched            (tensormethods.py:176-194, all four pairings)
    `.left_orthogonalize` -> qtt_orthogonalize_batched     (tensormethods.py:90-99)
    `.reapprox`           -> qtt_round_sum_batched         (tensormethods.py:101-133)
    `.norm_squared`       -> qtt_orthogonalize_batched     (tensormethods.py:280-283)
    `.inner(other)`       -> qtt_inner_batched             (the `(a @ b).squeeze().eval()` idiom)

Each call copies the cores host -> device, runs the batched kernels with N = 1 and copies the result
back, because callers own and mutate the numpy cores.  Many instances at once (the throughput path) go
through `laplacesolvermps_b200.batched` / `solver.solve_with_grad_descent_batched` instead.  There is no
CPU fallback: without the CUDA library these methods raise `QttError`.

Shape bookkeeping (`+`, scalar `*`, Kronecker `*`, transpose, squeeze, eval, ...) is memcpy-class work on
a handful of small cores and stays in numpy, exactly like the container itself.
"""
from typing import List

import numpy as np


def _device():
    from . import _device_ops
    return _device_ops


def _get_required_ranks_for_accuracy(s, eps):
    """Rank rule of the reference (tensormethods.py:5-12): every singular value with
    (s_i/s_0)^2 >= eps^2 is kept; it is not a cumulative tail criterion."""
    s = np.asarray(s, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        ratio = s ** 2 / s[0] ** 2
    return int(len(s) - np.sum(ratio < eps ** 2))


def zeros(mode_sizes):
    return TensorTrain([np.zeros((1,) + tuple(m) + (1,)) for m in mode_sizes])


class TensorTrain:
    _warned_no_orth = False

    def __init__(self, tensors):
        if isinstance(tensors, TensorTrain):
            self.tensors = [np.array(U, copy=True) for U in tensors]
        else:
            self.tensors = tensors

    # ---- container ----------------------------------------------------------------------------
    def __len__(self):
        return len(self.tensors)

    def __getitem__(self, index):
        return self.tensors[index]

    def __iter__(self):
        return iter(self.tensors)

    @property
    def ranks(self):
        return [U.shape[-1] for U in self.tensors[:-1]]

    @property
    def shapes(self):
        return [U.shape for U in self.tensors]

    @property
    def mode_sizes(self):
        return [U.shape[1:-1] for U in self.tensors]

    def copy(self):
        return TensorTrain([np.array(U, copy=True) for U in self.tensors])

    @classmethod
    def from_ttpylist(cls, tensors):
        """ttpy stores cores in Fortran order and the chain reversed (tensormethods.py:42-45)."""
        return cls([np.transpose(U) for U in tensors][::-1])

    def to_ttpylist(self):
        return [np.transpose(U) for U in self.tensors][::-1]

    @classmethod
    def ttsvd(cls, A, ranks: List[int], use_sparse_svd=False):
        """TT-SVD of a dense array with prescribed ranks (tensormethods.py:52-75).  A host-side
        format conversion of a dense tensor, not part of the TT hot path."""
        A = np.asarray(A, dtype=np.float64)
        d = A.ndim
        assert len(ranks) == d - 1
        rk = [1] + list(ranks) + [1]
        cores = []
        rest = A
        for k in range(d - 1):
            mat = rest.reshape(rk[k] * A.shape[k], -1)
            if use_sparse_svd:
                from scipy.sparse.linalg import svds
                Um, S, Vt = svds(mat, rk[k + 1])
            else:
                Um, S, Vt = np.linalg.svd(mat, full_matrices=False)
                Um, S, Vt = Um[:, :rk[k + 1]], S[:rk[k + 1]], Vt[:rk[k + 1]]
            cores.append(Um.reshape(rk[k], A.shape[k], rk[k + 1]))
            rest = S[:, None] * Vt
        cores.append(rest.reshape(rk[-2], -1, rk[-1]))
        return cls(cores)

    # ---- hot path: GPU ------------------------------------------------------------------------
    def __matmul__(self, other):
        assert len(other) == len(self)
        for U, V in zip(self.tensors, other.tensors):
            if U.ndim not in (3, 4) or V.ndim not in (3, 4):
                raise NotImplementedError("Currently only supports 1D or 2D mode dimensions")
        if len(self) >= 2 and all(U.ndim == 4 for U in self.tensors) and all(V.ndim == 3 for V in other.tensors):
            _device().require_engine()                                         # no GPU / no library: fail HERE, like the eager path
            for U, V in zip(self.tensors, other.tensors):
                if V.shape[1] != U.shape[2]:
                    raise ValueError(f"contracted mode sizes differ: {U.shape} @ {V.shape}")
            return _LazyProduct(list(self.tensors), list(other.tensors))       # operator @ vector: evaluated on first use
        return TensorTrain(_device().matmul(self.tensors, other.tensors))

    def left_orthogonalize(self):
        """Right-to-left QR sweep; the non-orthogonal core ends at position 0 (the reference's naming)."""
        self.tensors = _device().orthogonalize(self.tensors)
        return self

    def right_orthogonalize(self):
        """Mirror sweep (tensormethods.py:81-88): run the same kernels on the reversed chain."""
        rev = [np.ascontiguousarray(np.moveaxis(U, [0, -1], [-1, 0])) for U in self.tensors[::-1]]
        rev = _device().orthogonalize(rev)
        self.tensors = [np.ascontiguousarray(np.moveaxis(U, [0, -1], [-1, 0])) for U in rev[::-1]]
        return self

    def reapprox(self, ranks_new=None, rel_error=1e-16, orthogonalize=True):
        """TT rounding in place (tensormethods.py:101-133).  `orthogonalize=False` is accepted for API parity, but the device sweep
        always orthogonalises first: on an already orthogonalised chain (the only case in which the reference's own callers could
        skip the step - none of them does) the result is the same; on a non-orthogonal chain the reference would truncate without
        the orthogonality the rank rule assumes, here the truncation is the optimal one.  A warning says so once."""
        if not orthogonalize and not TensorTrain._warned_no_orth:
            import warnings
            TensorTrain._warned_no_orth = True
            warnings.warn("TensorTrain.reapprox(orthogonalize=False): the device sweep always orthogonalises first "
                          "(same result on an orthogonalised chain)", RuntimeWarning, stacklevel=2)
        L = len(self)
        if ranks_new is None:
            caps = None
        elif isinstance(ranks_new, (int, np.integer)):
            caps = [int(ranks_new)] * (L - 1)
        else:
            caps = [int(r) for r in ranks_new]
        self.tensors = _device().round(self.tensors, caps, float(rel_error))
        return self

    def norm_squared(self, orthogonalize=True):
        return _device().norm_squared(self.tensors)

    def inner(self, other):
        """<self, other> by transfer-matrix contraction on the device (same value as the reference's
        `(a @ b).squeeze().eval()` for vector-valued trains)."""
        assert len(other) == len(self)
        return _device().inner(self.tensors, other.tensors)

    # ---- sums, scalings, Kronecker products: host bookkeeping ----------------------------------
    def __add__(self, other):
        assert len(other) == len(self)
        L = len(self)
        out = []
        for k, (U, V) in enumerate(zip(self.tensors, other.tensors)):
            if k == 0:
                out.append(np.concatenate([U, V], axis=-1))
            elif k == L - 1:
                out.append(np.concatenate([U, V], axis=0))
            else:
                W = np.zeros((U.shape[0] + V.shape[0],) + U.shape[1:-1] + (U.shape[-1] + V.shape[-1],))
                W[:U.shape[0], ..., :U.shape[-1]] = U
                W[U.shape[0]:, ..., U.shape[-1]:] = V
                out.append(W)
        return TensorTrain(out)

    def __neg__(self):
        return self * (-1.0)

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, other):
        if isinstance(other, (float, int)):
            out = self.copy()
            out.tensors[0] = out.tensors[0] * other
            return out
        assert len(other) == len(self)
        out = []
        for U, V in zip(self.tensors, other.tensors):
            W = U[:, None, ..., :, None] * V[None, :, ..., None, :]      # p P modes.. q Q, modes broadcast
            out.append(W.reshape((W.shape[0] * W.shape[1],) + W.shape[2:-2] + (W.shape[-2] * W.shape[-1],)))
        return TensorTrain(out)

    def __rmul__(self, other):
        assert isinstance(other, (float, int)), "Only defined for scalar multiplication"
        return self * other

    # ---- mode bookkeeping ---------------------------------------------------------------------
    def transpose(self, perm=None):
        perm = perm or [1, 0]
        axes = [0] + [p + 1 for p in perm] + [-1]
        for k, U in enumerate(self.tensors):
            if U.ndim > 3:
                self.tensors[k] = np.transpose(U, axes)
        return self

    def flatten_mode_indices(self):
        for k, U in enumerate(self.tensors):
            self.tensors[k] = U.reshape(U.shape[0], -1, U.shape[-1])
        return self

    def unflatten_mode_indices(self, mode_shapes):
        for k, (U, s) in enumerate(zip(self.tensors, mode_shapes)):
            self.tensors[k] = U.reshape((U.shape[0],) + tuple(s) + (U.shape[-1],))
        return self

    def reshape_mode_indices(self, mode_shapes):
        if isinstance(mode_shapes[0], (int, np.integer)):
            mode_shapes = [mode_shapes] * len(self)
        for k, U in enumerate(self.tensors):
            self.tensors[k] = U.reshape((U.shape[0],) + tuple(mode_shapes[k]) + (U.shape[-1],))
        return self

    def expand_dims(self, axis=-1):
        for k, U in enumerate(self.tensors):
            if isinstance(axis, (int, np.integer)):
                pos = axis - 1 if axis < 0 else axis + 1
            else:
                pos = [a - 1 if a < 0 else a + 1 for a in axis]
            self.tensors[k] = np.expand_dims(U, pos)
        return self

    def squeeze(self):
        """Fold cores whose modes are all of size 1 into a neighbour (tensormethods.py:251-264)."""
        merged = [self.tensors[0]]
        for U in self.tensors[1:]:
            last = merged[-1]
            if int(np.prod(last.shape[1:-1], dtype=int)) == 1:
                merged[-1] = np.tensordot(last.reshape(last.shape[0], last.shape[-1]), U, axes=[-1, 0])
            elif int(np.prod(U.shape[1:-1], dtype=int)) == 1:
                merged[-1] = np.tensordot(last, U.reshape(U.shape[0], U.shape[-1]), axes=[-1, 0])
            else:
                merged.append(U)
        self.tensors = merged
        return self

    # ---- dense evaluation ---------------------------------------------------------------------
    def eval(self, squeeze=True, reshape=None):
        A = self.tensors[0]
        for U in self.tensors[1:]:
            A = np.tensordot(A, U, axes=(-1, 0))
        A = A.reshape(A.shape[1:-1])
        if squeeze:
            A = A.squeeze()
        if reshape == 'vector':
            return A.flatten()
        if reshape == 'matrix':
            L = len(self)
            A = A.transpose(list(range(0, 2 * L, 2)) + list(range(1, 2 * L, 2)))
            return A.reshape(int(np.prod(A.shape[:L], dtype=int)), -1)
        return A

    def evalm(self):
        return self.eval(reshape='matrix', squeeze=False)

    def evalv(self):
        return self.eval(reshape='vector', squeeze=False)


if __name__ == '__main__':
    np.random.seed(0)
    A = TensorTrain.ttsvd(np.random.normal(size=[2, 2, 2, 2, 2]), ranks=[2] * 4)
    print((A @ A).squeeze().eval(), A.norm_squared())



class _LazyProduct(TensorTrain):
    """`A @ x` for an operator (4-D cores) and a vector-valued train (3-D cores) whose cores are computed on first use.

    Every access to `.tensors` (and therefore every inherited method) materialises the product exactly as the eager path does
    (`contract_kernel`, tensormethods.py:176-194).  The two calls the reference's drivers make directly on such a product -
    `(A @ x).reapprox(...)` (solver.py:229-231, 268-270) and `.norm_squared()` - run the FUSED device sweeps on the still
    unevaluated product instead: nothing of bond R * r is materialised, copied to the host and uploaded again (at the 2D L = 12
    shapes: 25 ms instead of 140 ms per call, `tools/dev_single.py`).  The operands are captured by reference at the time of the
    product; like every TensorTrain result it does not alias their cores afterwards."""

    def __init__(self, op_cores, x_cores):
        self._pending = (op_cores, x_cores)
        self._cores = None

    def __len__(self):
        return len(self._pending[0]) if self._cores is None else len(self._cores)

    @property
    def tensors(self):
        if self._cores is None:
            op, x = self._pending
            self._cores = _device().matmul(op, x)
            self._pending = None
        return self._cores

    @tensors.setter
    def tensors(self, value):
        self._cores = value
        self._pending = None

    def squeeze(self):
        if self._cores is None and all(U.shape[1] != 1 for U in self._pending[0]):
            return self                                     # no mode of size 1: nothing to fold (tensormethods.py:251-264)
        return super().squeeze()

    def reapprox(self, ranks_new=None, rel_error=1e-16, orthogonalize=True):
        if self._cores is not None:
            return super().reapprox(ranks_new, rel_error, orthogonalize)
        op, x = self._pending
        L = len(op)
        if ranks_new is None:
            caps = None
        elif isinstance(ranks_new, (int, np.integer)):
            caps = [int(ranks_new)] * (L - 1)
        else:
            caps = [int(r) for r in ranks_new]
        try:
            self.tensors = _device().round_product(op, x, caps, float(rel_error))
        except _device().QttError as e:
            if "QTT_ERR_UNSUPPORTED" not in str(e):
                raise
            # outside the envelope of the fused sweep (e.g. vector cores above 200 KB): materialise - on the device - and round that
            return super().reapprox(ranks_new, rel_error, orthogonalize)
        return self

    def norm_squared(self, orthogonalize=True):
        if self._cores is not None:
            return super().norm_squared(orthogonalize)
        op, x = self._pending
        try:
            return _device().norm_squared_product(op, x)
        except _device().QttError as e:
            if "QTT_ERR_UNSUPPORTED" not in str(e):
                raise
            return super().norm_squared(orthogonalize)      # outside the fused sweep's envelope: materialised product, still on the device

