"""Single-instance entry points used by `TensorTrain`: numpy cores in, numpy cores out, arithmetic on
the GPU (N = 1 batches through the same kernels as the batched solver).  No CPU fallback."""
import numpy as np

from ._lib import QttError
from .batched import BatchedTT, DeviceMPO, default_engine


def _prod(shape):
    out = 1
    for s in shape:
        out *= int(s)
    return out


def require_engine():
    """Raise now (QttError) if there is no CUDA device or the library is missing."""
    default_engine()


def _as_f64(cores):
    return [np.ascontiguousarray(c, dtype=np.float64) for c in cores]


def matmul(a, b):
    """Core-wise product for the four ndim pairings of tensormethods.py:176-194."""
    eng = default_engine()
    a, b = _as_f64(a), _as_f64(b)
    m, out_modes, a3, b3 = [], [], [], []
    for U, V in zip(a, b):
        rows, cols = U.shape[1:-2], V.shape[2:-1]
        mk = U.shape[-2]
        if V.shape[1] != mk:
            raise ValueError(f"contracted mode sizes differ: {U.shape} @ {V.shape}")
        m.append(mk)
        out_modes.append(tuple(rows) + tuple(cols))
        a3.append(U.reshape(U.shape[0], -1, U.shape[-1]))
        b3.append(V.reshape(V.shape[0], -1, V.shape[-1]))
    A = BatchedTT.from_host([a3], eng.device, pin=False)
    X = BatchedTT.from_host([b3], eng.device, pin=False)
    Y = eng.matmul(A, X, m, out_modes)
    return Y.to_host()[0]


def orthogonalize(cores):
    eng = default_engine()
    X = BatchedTT.from_host([_as_f64(cores)], eng.device, pin=False)
    if X.L < 2:
        return X.to_host()[0]
    Y, _ = eng.orthogonalize([(None, X)])
    return Y.to_host()[0]


def round(cores, caps, rel_eps):
    eng = default_engine()
    X = BatchedTT.from_host([_as_f64(cores)], eng.device, pin=False)
    if X.L < 2:
        return X.to_host()[0]
    Y = eng.round_sum([(None, X)], caps=caps, rel_eps=rel_eps)
    return Y.to_host()[0]


def norm_squared(cores):
    eng = default_engine()
    cores = _as_f64(cores)
    if len(cores) < 2:
        return float(np.sum(cores[0] ** 2))
    X = BatchedTT.from_host([cores], eng.device, pin=False)
    return float(eng.norm2(X).cpu().numpy()[0])


def inner(a, b):
    eng = default_engine()
    a3 = [c.reshape(c.shape[0], -1, c.shape[-1]) for c in _as_f64(a)]
    b3 = [c.reshape(c.shape[0], -1, c.shape[-1]) for c in _as_f64(b)]
    A = BatchedTT.from_host([a3], eng.device, pin=False)
    B = BatchedTT.from_host([b3], eng.device, pin=False)
    return float(eng.inner(A, B).cpu().numpy()[0])


def round_product(op, x, caps, rel_eps):
    """round(op @ x) without materialising the product (operator cores 4-D, vector cores 3-D): the fused sweep of the solver."""
    eng = default_engine()
    A = DeviceMPO(_as_f64(op), eng.device)
    X = BatchedTT.from_host([_as_f64(x)], eng.device, pin=False)
    Y = eng.round_sum([(A, X)], caps=caps, rel_eps=rel_eps)
    return Y.to_host()[0]


def norm_squared_product(op, x):
    """||op @ x||^2 through the orthogonalisation sweep of the unevaluated product."""
    eng = default_engine()
    A = DeviceMPO(_as_f64(op), eng.device)
    X = BatchedTT.from_host([_as_f64(x)], eng.device, pin=False)
    return float(eng.orthogonalize([(A, X)], want_cores=False)[1].cpu().numpy()[0])
