"""Batched, device-resident function encoders: N polynomials / sinusoids -> one `BatchedTT`, with no per-instance
Python and no host numpy cores (reference: utils.py:188-225, `get_polynomial_as_tt`, `get_trig_function_as_tt`), plus
the core products the example problems are built from (`TensorTrain.__mul__`, tensormethods.py:154-167;
`kronecker_prod_2D`-style level interleaving, utils.py:182-185).  Every function here ends in a call into
csrc/libqtt_b200.so; an L / right-hand-side sweep stays on the GPU from the coefficients to the error norms.

    poly  = polynomial_batch(coeffs [N, n], L)                 # coeffs[i, k] multiplies x^k on [0, 1]
    trig  = trig_batch(abnu [N, 3], L)                         # a cos(2 pi nu x) + b sin(2 pi nu x)
    f1d   = hadamard(poly, trig)                               # pointwise product, bonds multiply
    f2d   = outer_2d(fx, fy)                                   # f(x) g(y), mode index 2 x_bit + y_bit (trailing core 2 x 2 -> 4)
"""
from ctypes import byref, c_int, c_void_p

import numpy as np
import torch

from . import _lib
from .batched import BatchedTT, default_engine
from .utils import _get_legendre_zoom_in_tensor

_poly_cache = {}


def _poly_shared_cores(n, L, device):
    """Instance-independent part of `get_polynomial_as_tt` for n coefficients: head[j] = (Legendre expansion on [0,1] of x^j) . Z,
    L - 1 zoom cores Z (n, 2, n), the corner evaluation (n, 2, 1)."""
    key = (n, L, str(device))
    if key not in _poly_cache:
        from numpy.polynomial import Polynomial
        Z = _get_legendre_zoom_in_tensor(n - 1)
        conv = np.zeros((n, n))
        for j in range(n):
            leg = Polynomial(np.eye(n)[j]).convert(kind=np.polynomial.Legendre, domain=[0, 1]).coef
            conv[j, :leg.size] = leg
        head = np.tensordot(conv, Z, axes=[1, 0])                    # (n, 2, n): block j = core 0 of the monomial x^j
        ends = np.stack([(-1.0) ** np.arange(n), np.ones(n)], axis=1)[..., None]
        host = [head] + [Z] * (L - 1) + [ends]
        dev = [torch.from_numpy(np.ascontiguousarray(c).reshape(-1)).to(device) for c in host]
        _poly_cache[key] = dev
    return _poly_cache[key]


def polynomial_batch(coeffs, L, device=None):
    """`get_polynomial_as_tt(coeffs[i], L)` for every row of coeffs ([N, n] array or device tensor): L level cores + the
    trailing corner core, bonds n."""
    eng = default_engine(device)
    c = torch.as_tensor(np.asarray(coeffs, dtype=np.float64) if not torch.is_tensor(coeffs) else coeffs,
                        dtype=torch.float64).to(eng.device).contiguous()
    if c.ndim == 1:
        c = c[None, :]
    N, n = c.shape
    assert L >= 1
    shared = _poly_shared_cores(n, L, eng.device)
    ranks = [1] + [n] * L + [1]
    y = BatchedTT([(2,)] * (L + 1), N, [ranks[k] * 2 * ranks[k + 1] for k in range(L + 1)], eng.device)
    ptrs = (c_void_p * (L + 1))(*[t.data_ptr() for t in shared])
    rk = (c_int * (L + 2))(*ranks)
    rc = eng.lib.qtt_encode_linear_batched(eng.h, n, c_void_p(c.data_ptr()), ptrs, rk, byref(y.c_struct()), eng._stream())
    _lib.check(eng.h, rc, "qtt_encode_linear_batched")
    return y


def trig_batch(abnu, L, device=None):
    """`get_trig_function_as_tt((a, b, nu), L)` for every row of abnu ([N, 3])."""
    eng = default_engine(device)
    c = torch.as_tensor(np.asarray(abnu, dtype=np.float64) if not torch.is_tensor(abnu) else abnu,
                        dtype=torch.float64).to(eng.device).contiguous()
    if c.ndim == 1:
        c = c[None, :]
    N = c.shape[0]
    assert c.shape[1] == 3 and L >= 1
    slot = [4] + [8] * (L - 1) + [4]
    y = BatchedTT([(2,)] * (L + 1), N, slot, eng.device)
    rc = eng.lib.qtt_encode_trig_batched(eng.h, c_void_p(c.data_ptr()), byref(y.c_struct()), eng._stream())
    _lib.check(eng.h, rc, "qtt_encode_trig_batched")
    return y


def hadamard(a, b):
    """Pointwise product of two batched trains with equal mode sizes (`a * b` of the reference)."""
    eng = default_engine(a.device)
    return eng.matmul(a, b, [0] * a.L, a.mode_shapes)


def outer_2d(fx, fy):
    """f(x) g(y) in level-interleaved 2D form: mode index n_y * x_index + y_index on every core."""
    eng = default_engine(fx.device)
    return eng.matmul(fx, fy, [1] * fx.L, [(fx.n[k] * fy.n[k],) for k in range(fx.L)])


def scale(x, factor):
    """x * factor[i] (device tensor [N] or float), in place on core 0."""
    if torch.is_tensor(factor):
        x.cores[0].view(x.N, -1).mul_(factor.view(-1, 1))
    else:
        x.cores[0].mul_(float(factor))
    return x
