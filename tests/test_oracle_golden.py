"""Pins oracle/tt_oracle.py (the CPU restatement) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py) and, when /root/reference is mounted, against the live reference."""
import numpy as np
import pytest

from conftest import cores_from, rel_err
from oracle import tt_oracle as O
from oracle.ref_loader import reference_available


def test_selftest_known_answer(golden):
    g = golden("tensor_algebra")
    A = cores_from(g, "selftest_cores")
    assert abs(O.inner(A, A) - 31.092539952213777) < 1e-12          # tensormethods.py:286-293
    assert abs(O.inner(A, A) - float(g["selftest_inner"])) < 1e-13
    assert abs(O.norm_squared(A) - float(g["selftest_norm2"])) < 1e-13


def test_matmul_all_pairings(golden):
    g = golden("tensor_algebra")
    a, b = cores_from(g, "mm_mpo_a"), cores_from(g, "mm_mpo_b")
    r, l, l2 = cores_from(g, "mm_mps_r"), cores_from(g, "mm_mps_l"), cores_from(g, "mm_mps_l2")
    assert rel_err(O.dense(O.matmul(a, b), False), g["mm_44_dense"]) < 1e-14
    assert rel_err(O.dense(O.matmul(a, r), False), g["mm_43_dense"]) < 1e-14
    assert rel_err(O.dense(O.matmul(l, a), False), g["mm_34_dense"]) < 1e-14
    m33 = O.matmul(l, l2)
    assert [list(c.shape) for c in m33] == g["mm_33_shapes"].tolist()
    assert abs(O.inner(l, l2) - float(g["mm_33_value"])) < 1e-13 * abs(float(g["mm_33_value"]))
    bil = float(O.dense(O.squeeze(O.matmul(O.matmul(l, a), r))))
    assert abs(bil - float(g["mm_bilinear"])) < 1e-12 * abs(float(g["mm_bilinear"]))
    assert O.bond_ranks(O.matmul(a, r)) == g["mm_43_ranks"].tolist()
    with pytest.raises(NotImplementedError):
        O.matmul([np.zeros((1, 2, 2, 2, 1))], [np.zeros((1, 2, 1))])


def test_orthogonalise_and_round(golden):
    g = golden("tensor_algebra")
    x, y = cores_from(g, "rd_x"), cores_from(g, "rd_y")
    xo = O.left_orthogonalize(O.clone(x))
    assert O.bond_ranks(xo) == g["rd_x_orth_ranks"].tolist()
    assert rel_err(O.dense(xo, False), g["rd_x_orth_dense"]) < 1e-13
    assert abs(O.norm_squared(x) - float(g["rd_x_norm2"])) < 1e-12 * float(g["rd_x_norm2"])
    s = O.add(x, y)
    assert O.bond_ranks(s) == g["rd_sum_ranks"].tolist()
    for tag, kw in [("cap3", dict(ranks_new=3)), ("eps", dict(rel_error=1e-10)),
                    ("caplist", dict(ranks_new=[2, 3, 5, 3, 1], rel_error=1e-12)), ("default", dict())]:
        t = O.reapprox(O.clone(s), **kw)
        assert O.bond_ranks(t) == g[f"rd_sum_{tag}_ranks"].tolist(), tag
        assert rel_err(O.dense(t, False), g[f"rd_sum_{tag}_dense"]) < 1e-12, tag
    d = O.reapprox(O.add(x, O.scale(x, 2.0)), rel_error=1e-10)
    assert O.bond_ranks(d) == g["rd_dup_ranks"].tolist()
    assert rel_err(O.dense(d, False), g["rd_dup_dense"]) < 1e-12
    mpo = cores_from(g, "rd_mpo")
    m2 = O.reapprox(O.add(mpo, mpo), rel_error=1e-10)
    assert O.bond_ranks(m2) == g["rd_mpo_ranks"].tolist()
    assert rel_err(O.dense(m2, False), g["rd_mpo_dense"]) < 1e-12


def test_rank_rule_is_not_cumulative():
    s = np.array([1.0, 0.6e-5, 0.6e-5, 0.6e-5, 0.6e-5])
    assert O.required_rank(s, 1e-5) == 1       # each value judged alone against sigma_0 (tensormethods.py:5-12)
    assert O.required_rank(s, 1e-6) == 5
    assert O.required_rank(np.zeros(3), 1e-3) == 3   # 0/0 -> nan -> nothing removed


def test_zero_train_keeps_ranks(golden):
    g = golden("tensor_algebra")
    rng = np.random.default_rng(0)
    z = O.add(O.zeros_like_modes([(2,)] * 4), [rng.standard_normal(s) for s in [(1, 2, 2), (2, 2, 2), (2, 2, 2), (2, 2, 1)]])
    z = O.sub(z, O.clone(z))
    assert O.bond_ranks(z) == g["rd_zero_ranks_before"].tolist()
    O.reapprox(z, ranks_new=3)
    assert O.bond_ranks(z) == g["rd_zero_ranks"].tolist()


def test_grad_descent_golden(golden):
    g = golden("solves")
    A, b = cores_from(g, "spd_A"), cores_from(g, "spd_b")
    x, r2 = O.solve_with_grad_descent(A, b, n_steps_max=60, max_rank=4)
    assert O.bond_ranks(x) == g["spd_x_ranks"].tolist()
    assert rel_err(O.dense(x, False), g["spd_x_dense"]) < 1e-9
    assert abs(r2 - float(g["spd_r2"])) <= 1e-6 * float(g["spd_r2"]) + 1e-30


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_against_live_reference():
    from oracle.ref_loader import load_reference
    from oracle.make_golden import random_tt
    tm, solver, bpx2D, utils = load_reference()
    rng = np.random.default_rng(42)
    L = 5
    op = random_tt(rng, [(2, 2)] * L, [3, 4, 4, 3])
    x = random_tt(rng, [(2,)] * L, [2, 3, 3, 2])
    ref = (tm.TensorTrain([c.copy() for c in op]) @ tm.TensorTrain([c.copy() for c in x])).reapprox(ranks_new=4, rel_error=1e-12)
    mine = O.reapprox(O.matmul(op, x), ranks_new=4, rel_error=1e-12)
    assert ref.ranks == O.bond_ranks(mine)
    assert rel_err(O.dense(mine, False), ref.eval(squeeze=False)) < 1e-12
    assert abs(O.norm_squared(x) - tm.TensorTrain([c.copy() for c in x]).norm_squared()) < 1e-12
    import io, contextlib
    A = tm.TensorTrain([c.copy() for c in op])
    A = (A.copy().transpose() @ A + utils._get_identy_as_tt(L)).reapprox(rel_error=1e-13)
    b = tm.TensorTrain([c.copy() for c in x])
    with contextlib.redirect_stdout(io.StringIO()):
        xr, r2r = solver.solve_with_grad_descent(A, b, n_steps_max=25, max_rank=5)
    xm, r2m = O.solve_with_grad_descent([c.copy() for c in A.tensors], [c.copy() for c in x], n_steps_max=25, max_rank=5)
    assert rel_err(O.dense(xm, False), xr.eval(squeeze=False)) < 1e-9
