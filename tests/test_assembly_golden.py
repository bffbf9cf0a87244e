"""Operator assembly and function encoders of the package (written fresh from the report's maths) against
dense evaluations of the reference's builders (tests/golden/operators.npz, functions.npz).  CPU only."""
import numpy as np
import pytest

from conftest import rel_err
import laplacesolvermps_b200.solver as S
import laplacesolvermps_b200.bpx2D as B2
import laplacesolvermps_b200.utils as U


def ev(tt):
    return tt.eval(squeeze=False)


@pytest.mark.parametrize("L", [2, 3, 4])
def test_1d_operators(golden, L):
    g = golden("operators")
    pairs = {
        "M": S.get_overlap_matrix_as_tt(L), "D": S.get_derivative_matrix_as_tt(L), "A": S.get_laplace_matrix_as_tt(L),
        "C": S.get_bpx_preconditioner(L), "Qp": S.get_bpx_Qp(L), "Qt": S.get_bpx_Qt(L), "B": S.get_laplace_bpx(L),
        "Rbpx": S.get_rhs_matrix_bpx(L), "Rnodal": S.get_rhs_matrix_as_tt(L), "G": S.get_gram_matrix_as_tt(L),
        "Gc": S.get_gram_matrix_as_tt(L, "corner"), "mass": S.build_mass_matrix_in_nodal_basis(L),
        "Cbysum": S.get_bpx_preconditioner_by_sum(L),
    }
    for name, tt in pairs.items():
        ref = g[f"{name}_{L}"]
        val = ev(tt)
        assert val.shape == ref.shape, name
        assert rel_err(val, ref) < 1e-12, name
    for l in range(L + 1):
        assert rel_err(ev(S.get_level_mapping_matrix_as_tt(L, l)), g[f"P_{L}_{l}"]) < 1e-13
    assert S.get_laplace_matrix_as_tt(L).ranks == g[f"A_{L}_ranks"].tolist()


@pytest.mark.parametrize("L", [2, 3])
def test_2d_operators(golden, L):
    g = golden("operators")
    pairs = {
        "A2D": S.build_laplace_matrix_2D(L), "mass2D": S.build_2D_mass_matrix(L), "R2D": S.get_rhs_matrix_as_tt_2D(L),
        "C2D": B2.get_BPX_preconditioner_2D(L), "Qpx2D": B2.get_BPX_Qp_2D(L, "x"), "Qpy2D": B2.get_BPX_Qp_2D(L, "y"),
        "Q2D": B2.get_BPX_Q_2D(L), "B2D": B2.get_laplace_BPX_2D(L), "Rbpx2D": B2.get_rhs_matrix_BPX_2D(L),
        "C2Dbysum": S.get_bpx_preconditioner_by_sum_2D(L),
    }
    for name, tt in pairs.items():
        ref = g[f"{name}_{L}"]
        val = ev(tt)
        assert val.shape == ref.shape, name
        assert rel_err(val, ref) < 1e-12, name


def test_explicit_vs_by_sum_cross_construction():
    """experiments/2D_validate_bpx_preconditioner.py: explicit cores against the level-sum builders."""
    L = 2
    assert rel_err(ev(B2.get_laplace_BPX_2D(L)), ev(S.get_laplace_bpx_2D_by_sum(L))) < 1e-12
    assert rel_err(ev(B2.get_rhs_matrix_BPX_2D(L)), ev(S.get_rhs_matrix_bpx_by_sum_2D(L))) < 1e-12


def test_rank_profiles(golden):
    g = golden("operators")
    B5 = B2.get_laplace_BPX_2D(5)
    assert B5.ranks == g["B2D_5_raw_ranks"].tolist()
    assert S._rounded(B5, rel_error=1e-12).ranks == g["B2D_5_ranks"].tolist()      # gap in the spectrum: well defined
    assert B2.get_BPX_preconditioner_2D(4).ranks == [32] * 3
    assert B2.get_BPX_Qp_2D(4).ranks == [24] * 3 + [8] or max(B2.get_BPX_Qp_2D(4).ranks) == 24


def test_condition_numbers(golden):
    """dense condition numbers of the BPX systems (1D L=2..8, 2D L=2..4)."""
    g = golden("operators")
    for i, L in enumerate(range(2, 9)):
        sv = np.linalg.svd(S.get_laplace_bpx(L).evalm(), compute_uv=False)
        assert abs(sv[0] / sv[-1] - g["cond_B1D_L2_8"][i]) < 1e-8 * g["cond_B1D_L2_8"][i]
    kat = [5.4127426440, 7.9349840729, 10.2732726902]
    for i, L in enumerate(range(2, 5)):
        sv = np.linalg.svd(B2.get_laplace_BPX_2D(L).evalm(), compute_uv=False)
        assert abs(sv[0] / sv[-1] - g["cond_B2D_L2_4"][i]) < 1e-8 * g["cond_B2D_L2_4"][i]
        assert abs(sv[0] / sv[-1] - kat[i]) < 1e-8


@pytest.mark.parametrize("L", [3, 5])
def test_function_encoders_1d(golden, L):
    g = golden("functions")
    assert rel_err(ev(U.get_polynomial_as_tt([1.0, -3.0, 2.0], L)), g[f"poly_{L}"]) < 1e-13
    assert rel_err(ev(U.get_trig_function_as_tt([0.3, 1.0, 1.0], L)), g[f"trig_{L}"]) < 1e-13
    assert rel_err(ev(U.get_example_u_1D(L, "corner")), g[f"u1d_corner_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_u_1D(L, "nodal")), g[f"u1d_nodal_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_f_1D(L)), g[f"f1d_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_u_deriv_1D(L, "corner")), g[f"du1d_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_u_deriv_1D(L, "average")), g[f"du1d_avg_{L}"]) < 1e-12


@pytest.mark.parametrize("L", [2, 3])
def test_function_encoders_2d(golden, L):
    g = golden("functions")
    assert rel_err(ev(U.get_example_u_2D(L, "corner")), g[f"u2d_corner_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_u_2D(L, "nodal")), g[f"u2d_nodal_{L}"]) < 1e-12
    assert rel_err(ev(U.get_example_f_2D(L)), g[f"f2d_{L}"]) < 1e-12
    Dx, Dy = U.get_example_grad_u_2D(L)
    assert rel_err(ev(Dx), g[f"gradx2d_{L}"]) < 1e-12
    assert rel_err(ev(Dy), g[f"grady2d_{L}"]) < 1e-12


def test_polynomial_samples_the_function():
    L = 6
    tt = U.evaluate_nodal_basis(U.get_polynomial_as_tt([0.5, -1.0, 0.25, 2.0], L), np.array([-1.0])).squeeze()
    x = np.arange(2 ** L) / 2 ** L
    assert np.allclose(tt.eval().reshape(-1), 0.5 - x + 0.25 * x ** 2 + 2 * x ** 3, atol=1e-13)


def test_factored_form_of_the_2d_bpx_operator():
    """B = S_x^T S_x + S_y^T S_y with S_d = 2^-L G^{1/2} Q'_d (the operators pipeline.FactoredSolver2D applies
    instead of the assembled rank-161 B) reproduces the assembled operator of bpx2D.py:159-171 densely."""
    from laplacesolvermps_b200.pipeline import build_operators_2D

    def dense_mat(cores):
        A = np.ones((1, 1, 1))
        for c in cores:
            r, no, ni, rr = c.shape
            A = np.einsum('abk,kijl->aibjl', A, c).reshape(A.shape[0] * no, A.shape[1] * ni, rr)
        return A[:, :, 0]

    for L in (2, 3):
        ops = build_operators_2D(L, 1e-14, on_device=False)
        Bd = dense_mat(ops["B"])
        tot = 0.0
        for key in ("GyQx", "GxQy"):
            cores = [c.copy() for c in ops[key]]
            cores[0] = cores[0] * 0.5 ** L
            Sd = dense_mat(cores)
            tot = tot + Sd.T @ Sd
        assert np.abs(Bd - tot).max() < 1e-13 * np.abs(Bd).max()
