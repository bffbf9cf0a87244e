"""Truncated TT conjugate gradients (the solver offered behind `solve_with_amen`).  The reference has no
implementation of it - it calls the third-party ttpy AMEn - so parity here means: (CPU) the numpy restatement
`oracle.tt_oracle.solve_with_cg` reaches the dense solution of the reference's own BPX systems, in fewer steps than the
reference's steepest descent; (GPU) the device loop reproduces the restatement's per-step history and end state."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import tt_oracle as O


def bpx_system_1d(L):
    """B v = b of solver.py:192-202 for f = 1, assembled by the package's host-side builders (CPU numpy)."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    f = U.get_polynomial_as_tt([1.0], L)
    B = O.reapprox([c.copy() for c in S.get_laplace_bpx(L).tensors], rel_error=1e-15)
    b = O.squeeze(O.matmul(S.get_rhs_matrix_bpx(L).tensors, f.tensors))
    for k in range(L):
        b[k] = b[k] * 0.5
    return B, b


def test_oracle_cg_reaches_dense_solution_faster_than_descent():
    B, b = bpx_system_1d(8)
    xd = np.linalg.solve(O.dense(B, reshape='matrix'), O.dense(b, reshape='vector').reshape(-1))
    def err(x):
        return rel_err(O.dense(x, reshape='vector').reshape(-1), xd)
    # cap 20 does not bind (the middle bond is 16): 25 CG steps do what 80 descent steps do
    x_cg, r2_cg = O.solve_with_cg(B, b, n_steps_max=25, max_rank=20, rel_accuracy=0.0)
    x_gd, _ = O.solve_with_grad_descent(B, b, n_steps_max=25, max_rank=20, rel_accuracy=0.0)
    assert err(x_cg) < 1e-8 and err(x_cg) < 1e-2 * err(x_gd), (err(x_cg), err(x_gd))
    assert r2_cg < 1e-16 * O.norm_squared(b)
    # cap 12 binds: the roundings perturb r and p, the line-search coefficients must keep converging (no stagnation)
    x_cg, _ = O.solve_with_cg(B, b, n_steps_max=40, max_rank=12, rel_accuracy=0.0)
    x_gd, _ = O.solve_with_grad_descent(B, b, n_steps_max=40, max_rank=12, rel_accuracy=0.0)
    assert err(x_cg) < 1e-7 and err(x_cg) <= err(x_gd), (err(x_cg), err(x_gd))


def test_oracle_cg_stopping_rule():
    B, b = bpx_system_1d(8)
    h = []
    x, r2 = O.solve_with_cg(B, b, n_steps_max=100, max_rank=12, rel_accuracy=1e-14, history=h)
    assert len(h) <= 30 and len(h) % 10 == 0      # tested at multiples of 10 only, like solver.py:258
    assert r2 / O.norm_squared(x) <= 1e-13


@pytest.mark.gpu
def test_device_cg_matches_the_restatement():
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, default_engine
    eng = default_engine()
    B, b = bpx_system_1d(8)
    h_ref = []
    x_ref, r2_ref = O.solve_with_cg(B, b, n_steps_max=40, max_rank=12, rel_accuracy=0.0, history=h_ref)
    bs = [b, [b[0] * 0.37] + [c.copy() for c in b[1:]]]   # second instance: scaled right-hand side (same alphas, scaled residuals)
    x, r2, steps, hist = eng.cg_solve(DeviceMPO(B, eng.device), BatchedTT.from_host(bs, eng.device), n_steps=40, max_rank=12,
                                      rel_accuracy=0.0, want_history=True)
    h = hist.cpu().numpy()
    for n, rr, alpha in h_ref[:12]:             # compared while the residual is far above the round-off floor
        assert abs(h[n, 0, 0] - rr) < 1e-8 * rr and abs(h[n, 0, 1] - alpha) < 1e-8 * abs(alpha), n
        assert abs(h[n, 1, 0] - 0.37 ** 2 * rr) < 1e-8 * rr and abs(h[n, 1, 1] - alpha) < 1e-8 * abs(alpha), n
    xh = x.to_host()
    xd = np.linalg.solve(O.dense(B, reshape='matrix'), O.dense(b, reshape='vector').reshape(-1))
    e0 = rel_err(O.dense(xh[0], reshape='vector').reshape(-1), xd)
    e1 = rel_err(O.dense(xh[1], reshape='vector').reshape(-1), 0.37 * xd)
    eref = rel_err(O.dense(xh[0], reshape='vector'), O.dense(x_ref, reshape='vector'))
    assert e0 < 1e-7 and e1 < 1e-7, (e0, e1)        # 40 steps at a binding cap: 1e-8 in the CPU restatement
    assert eref < 1e-7, eref
    assert float(r2[0]) < 1e-14 * O.norm_squared(b), float(r2[0]) / O.norm_squared(b)


@pytest.mark.gpu
def test_device_cg_through_the_solver_api_and_amen_fallback():
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.tensormethods as tm
    import laplacesolvermps_b200.utils as U
    L = 10
    f = U.get_polynomial_as_tt([1.0], L)
    u, Du = S.solve_PDE_1D_with_preconditioner(f, max_rank=20, method='cg', nswp=1)     # <= 50 CG steps
    # f = 1: u = x - x^2/2 (natural condition at x = 1): ||u||^2_L2 = 2/15, |u|^2_H1 = 1/3
    assert abs(S.get_L2_norm_1D(u) - 2.0 / 15.0) < 1e-6          # get_L2_norm_1D returns the squared norm (solver.py:420-428)
    assert abs(Du.norm_squared() * 2.0 ** (-L) - 1.0 / 3.0) < 1e-5
    B, b = bpx_system_1d(L)
    x, r2 = S.solve_with_cg(tm.TensorTrain(B), tm.TensorTrain(b), n_steps_max=200, max_rank=20)
    assert r2 / x.norm_squared() <= 1e-13


@pytest.mark.gpu
def test_batched_pipeline_with_cg():
    """pipeline.Solver2DBatched(method='cg'): at caps that do not bind (2D L = 3) 30 CG steps and 120 descent steps reach
    the same solution of B v = b."""
    import torch
    from bench import c5_rhs_cores
    from laplacesolvermps_b200.batched import BatchedTT, default_engine
    from laplacesolvermps_b200.pipeline import Solver2DBatched, build_operators_2D
    eng = default_engine()
    L = 3
    ops = build_operators_2D(L, 1e-14, on_device=True)
    f = BatchedTT.from_host([c5_rhs_cores(i, L) for i in range(3)], eng.device)
    a = Solver2DBatched(L, max_rank=16, n_steps=120, eps=1e-14, host_ops=ops).solve(f)
    b = Solver2DBatched(L, max_rank=16, n_steps=30, eps=1e-14, host_ops=ops, method="cg").solve(f)
    assert torch.allclose(a["energy"], b["energy"], rtol=1e-9, atol=0) and torch.allclose(a["l2"], b["l2"], rtol=1e-9, atol=0)
    assert float((b["r2"] / eng.norm2(b["b"])).max()) < 1e-20
