"""Batched device function encoders (laplacesolvermps_b200/encoders.py) against the host encoders of the package, which the
reference goldens pin (tests/test_assembly_golden.py: `poly_*`, `trig_*`, `f2d_*` of tests/golden/functions.npz), and a
right-hand-side sweep that never leaves the GPU.  Reference: utils.py:188-225, tensormethods.py:154-167."""
import numpy as np
import pytest
import torch

from oracle import tt_oracle as O

pytestmark = pytest.mark.gpu


def _dense(cores):
    return O.dense([np.asarray(c).reshape(c.shape[0], -1, c.shape[-1]) for c in cores], False)


def test_polynomial_and_trig_batches_match_the_host_encoders(golden):
    import laplacesolvermps_b200.utils as utils
    from laplacesolvermps_b200 import encoders as E
    g = golden("functions")
    rng = np.random.default_rng(3)
    for L in (3, 5, 9):
        coeffs = np.vstack([[1.0, -3.0, 2.0], rng.standard_normal((4, 3))])
        P = E.polynomial_batch(coeffs, L).to_host()
        for i in range(coeffs.shape[0]):
            ref = utils.get_polynomial_as_tt(coeffs[i], L)
            assert [c.shape for c in P[i]] == [c.shape for c in ref.tensors]
            assert np.allclose(_dense(P[i]), _dense(ref.tensors), rtol=0, atol=1e-13 * np.abs(_dense(ref.tensors)).max())
        abnu = np.vstack([[0.3, 1.0, 1.0], np.column_stack([rng.standard_normal((4, 2)), rng.integers(1, 5, 4)])])
        Tg = E.trig_batch(abnu, L).to_host()
        for i in range(abnu.shape[0]):
            ref = utils.get_trig_function_as_tt(list(abnu[i]), L)
            assert [c.shape for c in Tg[i]] == [c.shape for c in ref.tensors]
            assert np.allclose(_dense(Tg[i]), _dense(ref.tensors), rtol=0, atol=1e-13 * np.abs(_dense(ref.tensors)).max())
        if f"poly_{L}" in g:                                            # the reference's own dense evaluations
            assert np.allclose(_dense(P[0]).reshape(-1), np.asarray(g[f"poly_{L}"]).reshape(-1), atol=1e-13)
            assert np.allclose(_dense(Tg[0]).reshape(-1), np.asarray(g[f"trig_{L}"]).reshape(-1), atol=1e-13)


def test_hadamard_and_outer_products(golden):
    import laplacesolvermps_b200.utils as utils
    from laplacesolvermps_b200 import encoders as E
    L = 4
    coeffs = np.array([[1.0, -3.0, 2.0], [0.5, 0.0, -1.0]])
    abnu = np.array([[0.0, 1.0, 1.0], [1.0, 0.5, 2.0]])
    P, Tg = E.polynomial_batch(coeffs, L), E.trig_batch(abnu, L)
    H = E.hadamard(P, Tg).to_host()
    for i in range(2):
        ref = utils.get_polynomial_as_tt(coeffs[i], L) * utils.get_trig_function_as_tt(list(abnu[i]), L)
        assert [c.shape for c in H[i]] == [c.shape for c in ref.tensors]
        assert np.allclose(_dense(H[i]), _dense(ref.tensors), atol=1e-13)
    # f(x) g(y), level interleaved: the same as expand_dims / `*` / flatten on the host (utils.get_example_u_2D's construction)
    X = E.polynomial_batch(np.array([[0.0, -1.0, 5.0, -3.0]] * 2), L)
    F2 = E.outer_2d(X, E.hadamard(P, Tg)).to_host()
    for i in range(2):
        ux = utils.get_polynomial_as_tt([0.0, -1.0, 5.0, -3.0], L)
        uy = utils.get_polynomial_as_tt(coeffs[i], L) * utils.get_trig_function_as_tt(list(abnu[i]), L)
        ref = (ux.expand_dims(1) * uy.expand_dims(0)).flatten_mode_indices()
        assert np.allclose(_dense(F2[i]), _dense(ref.tensors), atol=1e-13)


def test_right_hand_side_sweep_stays_on_the_device():
    """A batch of manufactured problems -Laplace(u) = f with u_i = X(x) * (p_i(y) sin(2 pi nu_i y)) variants: f_i encoded on the
    device from coefficient arrays, solved, and compared with the same instances pushed through the host encoders."""
    import laplacesolvermps_b200.utils as utils
    from laplacesolvermps_b200 import encoders as E
    from laplacesolvermps_b200.batched import BatchedTT, default_engine
    from laplacesolvermps_b200.pipeline import Solver2DBatched
    eng = default_engine()
    L, N = 5, 6
    rng = np.random.default_rng(9)
    cx, cy = rng.standard_normal((N, 3)), rng.standard_normal((N, 3))
    abnu = np.column_stack([rng.standard_normal((N, 2)), rng.integers(1, 3, N)]).astype(float)
    f_dev = E.outer_2d(E.polynomial_batch(cx, L), E.hadamard(E.polynomial_batch(cy, L), E.trig_batch(abnu, L)))
    host = []
    for i in range(N):
        fx = utils.get_polynomial_as_tt(cx[i], L)
        fy = utils.get_polynomial_as_tt(cy[i], L) * utils.get_trig_function_as_tt(list(abnu[i]), L)
        host.append((fx.expand_dims(1) * fy.expand_dims(0)).flatten_mode_indices().tensors)
    solver = Solver2DBatched(L, max_rank=12, n_steps=40, eps=1e-13)
    a = solver.solve(f_dev)
    b = solver.solve(BatchedTT.from_host(host, eng.device))
    assert torch.allclose(a["energy"], b["energy"], rtol=1e-10, atol=0)
    assert torch.allclose(a["l2"], b["l2"], rtol=1e-10, atol=0)


def test_batching_over_level_counts():
    """Right-hand sides with different numbers of levels in one call (bucketed by L), against per-L batched solves."""
    from laplacesolvermps_b200.batched import BatchedTT, default_engine
    from laplacesolvermps_b200.pipeline import Solver2DBatched, solve_2D_mixed_levels
    from oracle.make_golden import c5_rhs_cores
    eng = default_engine()
    levels = [3, 5, 4, 3, 5, 4, 4]
    fs = [c5_rhs_cores(10 + i, L) for i, L in enumerate(levels)]
    mixed = solve_2D_mixed_levels(fs, max_rank=6, n_steps=25, eps=1e-13)
    for L in sorted(set(levels)):
        idx = [i for i, l in enumerate(levels) if l == L]
        ref = Solver2DBatched(L, max_rank=6, n_steps=25, eps=1e-13).solve(BatchedTT.from_host([fs[i] for i in idx], eng.device))
        for j, i in enumerate(idx):
            assert mixed[i]["L"] == L
            assert mixed[i]["energy"] == float(ref["energy"][j]) and mixed[i]["l2"] == float(ref["l2"][j])
