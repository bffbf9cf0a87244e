"""Function encoders and manufactured example problems (API of the reference's `laplace_mps.utils`).

Host-side, construction-time code: the trains built here have ranks <= 24 and are the INPUTS of the
TT hot path.  Written from the mathematics of the project report (doc/Tensormethods_Project.pdf):
  * a polynomial is carried by its Legendre coefficients; halving the interval maps the coefficient
    vector linearly (the "zoom" core), the trailing core evaluates at the two cell corners
    (reference: utils.py:188-205, 260-271);
  * a sinusoid a cos(2 pi nu x) + b sin(2 pi nu x) is carried by a rotating 2-vector: every level adds
    the phase +-pi nu 2^-l (reference: utils.py:208-225).
matplotlib is imported lazily (the reference imports it at module level, which makes the whole package
unimportable on headless boxes).
"""
from typing import List

import numpy as np
from numpy.polynomial import Polynomial, polynomial as _P, legendre as _Lg

from . import tensormethods as tm
from . import _assembly as _asm

REL_ERROR = 1e-15


def _rounded(tt, **kw):
    """Construction-time rounding on the host (see _assembly.py)."""
    _asm.round_(tt.tensors, **kw)
    return tt


def eval_poly(x, coeffs):
    return _P.polyval(np.asarray(x, dtype=float), np.asarray(coeffs, dtype=float))


def draw_vertical_grid(L, ax=None):
    import matplotlib.pyplot as plt
    ax = ax or plt.gca()
    for i in range(2 ** min(L, 6) + 1):
        ax.axvline(i * 2.0 ** (-min(L, 6)), color='k', alpha=0.1)


class TrigFunction:
    """a cos(omega x) + b sin(omega x)."""

    def __init__(self, a, b, omega):
        self.a, self.b, self.omega = a, b, omega

    def eval(self, x):
        return self.a * np.cos(self.omega * x) + self.b * np.sin(self.omega * x)

    def derive(self):
        return TrigFunction(self.b * self.omega, -self.a * self.omega, self.omega)

    @property
    def coeffs(self):
        return self.a, self.b, self.omega / (2 * np.pi)


# ---- encoders ----------------------------------------------------------------------------------

def _get_legendre_zoom_in_tensor(degree):
    """Z[i, h, j]: P_i restricted to the left (h=0) / right (h=1) half of [-1,1], re-expanded in the
    Legendre polynomials of that half:  P_i((t -+ 1)/2) = sum_j Z[i,h,j] P_j(t)."""
    n = degree + 1
    Z = np.zeros((n, 2, n))
    for i in range(n):
        mono = _Lg.leg2poly(np.eye(n)[i])
        for h, shift in ((0, -0.5), (1, 0.5)):
            comp = np.zeros(1)
            for c in mono[::-1]:                      # Horner in the polynomial (t/2 + shift)
                comp = _P.polyadd(_P.polymul(comp, [shift, 0.5]), [c])
            leg = _Lg.poly2leg(comp)
            Z[i, h, :min(n, leg.size)] = leg[:n]
    return Z


def get_polynomial_as_tt(coeffs, L):
    """coeffs[k] multiplies x^k on [0,1].  L level cores plus one trailing core with the two corner
    values of each cell."""
    coeffs = np.atleast_1d(np.asarray(coeffs, dtype=float))
    n = coeffs.size
    leg = Polynomial(coeffs).convert(kind=np.polynomial.Legendre, domain=[0, 1]).coef
    leg = np.concatenate([leg, np.zeros(n - leg.size)])
    Z = _get_legendre_zoom_in_tensor(n - 1)
    cores = [Z.copy() for _ in range(L)]
    cores[0] = np.tensordot(leg[None, :], cores[0], axes=[-1, 0])
    ends = np.stack([(-1.0) ** np.arange(n), np.ones(n)], axis=1)      # P_j(-1), P_j(+1)
    cores.append(ends[..., None])
    return tm.TensorTrain(cores)


def _rot(phi):
    c, s = np.cos(phi), np.sin(phi)
    return np.array([[c, -s], [s, c]])


def get_trig_function_as_tt(coeffs, L):
    """coeffs = (a, b, nu): a cos(2 pi nu x) + b sin(2 pi nu x) on [0,1]."""
    a, b, nu = coeffs
    cores = [(np.array([a, b]) @ _rot(nu * np.pi)).reshape(1, 1, 2)]
    for l in range(1, L + 1):
        step = 0.5 ** l * np.pi * nu
        cores.append(np.stack([_rot(-step), _rot(step)], axis=1))
    c, s = np.cos(0.5 ** L * np.pi * nu), np.sin(0.5 ** L * np.pi * nu)
    cores.append(np.array([[c, c], [-s, s]])[..., None])
    return tm.TensorTrain(cores).squeeze()


def evaluate_nodal_basis(tt: tm.TensorTrain, s, basis='corner'):
    """Contract the trailing (within-cell) core with basis functions evaluated at the points s in [-1,1]."""
    s = np.array(s)
    out = tt.copy()
    last = out.tensors[-1]
    if s.ndim == 1:
        assert tt[-1].shape[1] == 2, "Tensor to be evaluated must be in nodal-basis, i.e. have 2 basis elements"
        if basis == 'corner':
            w = np.array([(1 - s) / 2, (1 + s) / 2])
        elif basis == 'average':
            w = np.array([(1 - s) / 2, (1 + s) / 2]) @ np.ones([2, 1]) * 0.5
        elif basis == 'legendre':
            w = np.array([np.ones_like(s), 2 * s - 1])
        else:
            raise NotImplementedError(f"Unknown basis {basis}")
        out.tensors[-1] = (last.squeeze(-1) @ w)[..., None]
        out.squeeze()
    elif s.ndim == 2 and s.shape[0] == 2:
        assert tt[-1].shape[1:3] == (2, 2)
        if basis not in ('corner', 'average'):
            raise NotImplementedError(f"Unknown basis {basis}")
        w = np.array([(1 - s[0]) * (1 - s[1]), (1 - s[0]) * (1 + s[1]), (1 + s[0]) * (1 - s[1]), (1 + s[0]) * (1 + s[1])]) / 4
        out.tensors[-1] = (last.reshape(-1, 4) @ w)[..., None]
        if basis == 'average':
            out.tensors[-1] = (out.tensors[-1].reshape(-1, 4) @ (np.ones(4) * 0.25)).reshape(-1, 4, 1)
            out.squeeze()
    return out


# ---- generic u -> f helpers (reference utils.py:48-106) ----------------------------------------

def build_u_with_correct_boundary_conditions(poly_coeffs, trig_coeffs):
    trig_functions = [TrigFunction(c[0], c[1], c[2] * 2 * np.pi) for c in trig_coeffs]
    t1 = np.sum([t.eval(1) for t in trig_functions])
    dt1 = np.sum([t.derive().eval(1) for t in trig_functions])
    ratio = dt1 / t1
    poly_coeffs[0] = 0
    poly_coeffs[1] = -np.sum([(ratio + i) * poly_coeffs[i] for i in range(2, len(poly_coeffs))]) / (1 + ratio)
    return np.polynomial.Polynomial(poly_coeffs), trig_functions


def get_f_from_u(poly: np.polynomial.Polynomial, trig: List[TrigFunction], L):
    acc = [tm.zeros([(2,) for _ in range(L + 1)]) for _ in range(3)]
    for t in trig:
        d1 = t.derive()
        for i, fn in enumerate((t, d1, d1.derive())):
            acc[i] = acc[i] + get_trig_function_as_tt(fn.coeffs, L)
    p0, p1, p2 = (get_polynomial_as_tt(poly.deriv(i).coef if i else poly.coef, L) for i in range(3))
    f = acc[0] * p2 + 2 * acc[1] * p1 + acc[2] * p0
    return -_rounded(f, rel_error=1e-16)


def evaluate_f_from_u(poly: np.polynomial.Polynomial, trig: List[TrigFunction], x):
    t0 = sum(t.eval(x) for t in trig)
    t1 = sum(t.derive().eval(x) for t in trig)
    t2 = sum(t.derive().derive().eval(x) for t in trig)
    return -(poly(x) * t2 + 2 * poly.deriv(1)(x) * t1 + poly.deriv(2)(x) * t0)


def eval_function(poly, trig_functions, x):
    return sum(t.eval(x) for t in trig_functions) * poly(x)


def get_u_function_as_tt(poly, trig: List[TrigFunction], L):
    u = get_trig_function_as_tt(trig[0].coeffs, L)
    for t in trig[1:]:
        u = u + get_trig_function_as_tt(t.coeffs, L)
    return _rounded(u * get_polynomial_as_tt(poly.coef, L), rel_error=1e-16)


# ---- the manufactured problems of the experiments ----------------------------------------------
# 1D: u = (1 - 3x + 2x^2) sin(2 pi x);  2D: u = (-x + 5x^2 - 3x^3) * (1 - 3y + 2y^2) sin(2 pi y)

_PI = np.pi
_U1_POLY = [1, -3, 2]
_U1_SIN = [0, 1, 1.0]
_U1_COS = [1, 0, 1.0]
_UX_POLY = [0, -1, 5, -3]


def _corner_points(basis, allow_average=False):
    if basis == 'corner' or (allow_average and basis == 'average'):
        return np.array([-1, 1])
    if basis == 'nodal':
        return np.array([1])
    raise ValueError(f"Unknown basis: {basis}")


def get_example_u_1D(L, basis='corner'):
    u = get_polynomial_as_tt(_U1_POLY, L) * get_trig_function_as_tt(_U1_SIN, L)
    return _rounded(evaluate_nodal_basis(u, _corner_points(basis)).squeeze(), rel_error=REL_ERROR)


def get_example_u_deriv_1D(L, basis='corner'):
    du = get_polynomial_as_tt([-3, 4], L) * get_trig_function_as_tt(_U1_SIN, L) + \
        get_polynomial_as_tt([2 * _PI, -6 * _PI, 4 * _PI], L) * get_trig_function_as_tt(_U1_COS, L)
    d = _rounded(evaluate_nodal_basis(du, _corner_points(basis, allow_average=True)).squeeze(), rel_error=REL_ERROR)
    if basis == 'average':
        d.tensors[-1] = (d.tensors[-1].reshape(-1, 2) @ np.array([0.5, 0.5])).reshape(-1, 1, 1)
        d.squeeze()
    return d


def _minus_upp_1d(L):
    """u'' for the 1D example as an (unrounded) train."""
    return get_polynomial_as_tt([4 - 4 * _PI ** 2, 12 * _PI ** 2, -8 * _PI ** 2], L) * get_trig_function_as_tt(_U1_SIN, L) + \
        get_polynomial_as_tt([-12 * _PI, 16 * _PI], L) * get_trig_function_as_tt(_U1_COS, L)


def get_example_f_1D(L):
    return -_rounded(_minus_upp_1d(L), rel_error=REL_ERROR)


def get_example_u_2D(L, basis='corner'):
    ux = get_polynomial_as_tt(_UX_POLY, L)
    uy = get_polynomial_as_tt(_U1_POLY, L) * get_trig_function_as_tt(_U1_SIN, L)
    u = ux.expand_dims(1) * uy.expand_dims(0)
    if basis == 'corner':
        s = np.array([[-1, -1, 1, 1], [-1, 1, -1, 1]])
    elif basis == 'nodal':
        s = np.ones([2, 1])
    else:
        raise ValueError(f"Unknown basis: {basis}")
    return evaluate_nodal_basis(u, s).squeeze()


def get_example_f_2D(L):
    ux = get_polynomial_as_tt(_UX_POLY, L)
    uy = get_polynomial_as_tt(_U1_POLY, L) * get_trig_function_as_tt(_U1_SIN, L)
    uxx = get_polynomial_as_tt([10, -18], L)
    uyy = _minus_upp_1d(L)
    f = ux.expand_dims(1) * uyy.expand_dims(0) + uxx.expand_dims(1) * uy.expand_dims(0)
    return -_rounded(f, rel_error=REL_ERROR)


def get_example_grad_u_2D(L):
    half_sum_diff = np.array([[0.5, -0.5], [0.5, 0.5]])           # corner values -> (mean, half difference)
    ux = get_polynomial_as_tt(_UX_POLY, L)
    ux.tensors[-1] = (ux.tensors[-1].reshape(-1, 2) @ half_sum_diff).reshape(-1, 2, 1)
    uy = get_polynomial_as_tt(_U1_POLY, L) * get_trig_function_as_tt(_U1_SIN, L)
    uy.tensors[-1] = (uy.tensors[-1].reshape(-1, 2) @ half_sum_diff).reshape(-1, 2, 1)
    dux = get_polynomial_as_tt([-1, 10, -9], L)
    duy = get_polynomial_as_tt([-3, 4], L) * get_trig_function_as_tt(_U1_SIN, L) + \
        get_polynomial_as_tt([2 * _PI, -6 * _PI, 4 * _PI], L) * get_trig_function_as_tt([1.0, 0.0, 1.0], L)
    dux = evaluate_nodal_basis(dux, [-1, 1], 'average').reshape_mode_indices([2, 1])
    dux.tensors.append(np.ones([1, 1, 1, 1]))
    duy = evaluate_nodal_basis(duy, [-1, 1], 'average').reshape_mode_indices([2, 1])
    duy.tensors.append(np.ones([1, 1, 1, 1]))
    Dx = kronecker_prod_2D(dux, uy.reshape_mode_indices([2, 1])).flatten_mode_indices()
    Dy = kronecker_prod_2D(ux.reshape_mode_indices([2, 1]), duy).flatten_mode_indices()
    return Dx, Dy


# ---- small operator helpers --------------------------------------------------------------------

def kronecker_prod_2D(A, B):
    """Level-interleaved Kronecker product of two 1D operator trains: mode index = 2 * x_bit + y_bit."""
    C = A.copy().expand_dims([1, 3]) * B.copy().expand_dims([0, 2])
    return C.reshape_mode_indices([(U.shape[1] * U.shape[2], U.shape[3] * U.shape[4]) for U in C])


def _get_gram_matrix_tt(L):
    cores = [np.eye(2)[None, :, :, None] for _ in range(L)] + [np.diag([2, 2 / 3]).reshape(1, 2, 2, 1)]
    return tm.TensorTrain(cores)


def _get_identy_as_tt(L, tailing_eye=False):
    cores = [np.eye(2)[None, :, :, None] for _ in range(L)]
    if tailing_eye:
        cores.append(np.eye(1)[None, :, :, None])
    return tm.TensorTrain(cores)
