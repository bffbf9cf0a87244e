"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_loader.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden            # small fixtures (seconds to a minute)
    python -m oracle.make_golden --heavy    # adds the 2D L=12 fixtures (tens of minutes of CPU)

The fixtures travel to the GPU box; /root/reference does not.  Only contracted quantities are stored
(dense evaluations at small L, inner products with seeded probe TTs, norms, rank lists) - never raw
cores of a result, because QR/SVD signs and the gauge between cores are arbitrary.
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def random_tt(rng, mode_shapes, ranks, scale=1.0):
    """Seeded random TT used as input / probe everywhere (also by the tests, which import this)."""
    rk = [1] + list(ranks) + [1]
    return [scale * rng.standard_normal((rk[k],) + tuple(mode_shapes[k]) + (rk[k + 1],))
            for k in range(len(mode_shapes))]


def c5_rhs_cores(i, L=12):
    """right-hand side of instance i - L+1 cores (1,4,2),(2,4,2)..,(2,4,1),
    entries default_rng(1234+i).standard_normal scaled by 1/sqrt(8) per core."""
    rng = np.random.default_rng(1234 + i)
    return random_tt(rng, [(4,)] * (L + 1), [2] * L, scale=1.0 / np.sqrt(8.0))


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


def pack_cores(prefix, cores):
    return {f"{prefix}_{k}": np.asarray(c) for k, c in enumerate(cores)}


def gen_tensor_algebra(tm):
    out = {}
    # (1) module self-test, tensormethods.py:286-293
    np.random.seed(0)
    A = tm.TensorTrain.ttsvd(np.random.normal(size=[2, 2, 2, 2, 2]), ranks=[2] * 4)
    out.update(pack_cores("selftest_cores", A.tensors))
    out["selftest_inner"] = np.float64((A @ A).squeeze().eval())
    out["selftest_norm2"] = np.float64(A.norm_squared())

    # (2) the four ndim pairings of `@`
    rng = np.random.default_rng(11)
    L = 4
    mpo_a = random_tt(rng, [(3, 2)] * L, [3, 4, 2])
    mpo_b = random_tt(rng, [(2, 3)] * L, [2, 3, 3])
    mps_r = random_tt(rng, [(2,)] * L, [3, 2, 4])
    mps_l = random_tt(rng, [(3,)] * L, [2, 2, 3])
    mps_l2 = random_tt(rng, [(3,)] * L, [4, 1, 2])
    out.update(pack_cores("mm_mpo_a", mpo_a)); out.update(pack_cores("mm_mpo_b", mpo_b))
    out.update(pack_cores("mm_mps_r", mps_r)); out.update(pack_cores("mm_mps_l", mps_l))
    out.update(pack_cores("mm_mps_l2", mps_l2))
    T = tm.TensorTrain
    r44 = T(mpo_a) @ T(mpo_b)
    r43 = T(mpo_a) @ T(mps_r)
    r34 = T(mps_l) @ T(mpo_a)
    r33 = T(mps_l) @ T(mps_l2)
    out["mm_44_dense"] = r44.eval(squeeze=False); out["mm_44_ranks"] = np.array(r44.ranks)
    out["mm_43_dense"] = r43.eval(squeeze=False); out["mm_43_ranks"] = np.array(r43.ranks)
    out["mm_34_dense"] = r34.eval(squeeze=False); out["mm_34_ranks"] = np.array(r34.ranks)
    out["mm_33_shapes"] = np.array([U.shape for U in r33.tensors])
    out["mm_33_value"] = np.float64(r33.squeeze().eval())
    out["mm_bilinear"] = np.float64((T(mps_l) @ T(mpo_a) @ T(mps_r)).squeeze().eval())

    # (3) orthogonalisation and rounding: cap-limited, eps-limited (with a spectral gap), mixed
    rng = np.random.default_rng(12)
    L = 6
    x = random_tt(rng, [(2,)] * L, [2, 4, 8, 4, 2])
    y = random_tt(rng, [(2,)] * L, [2, 3, 3, 3, 2])
    out.update(pack_cores("rd_x", x)); out.update(pack_cores("rd_y", y))
    xo = T([c.copy() for c in x]).left_orthogonalize()
    out["rd_x_orth_dense"] = xo.eval(squeeze=False)
    out["rd_x_orth_ranks"] = np.array(xo.ranks)
    out["rd_x_norm2"] = np.float64(T([c.copy() for c in x]).norm_squared())
    s = T(x) + T(y)                                   # exact rank <= rank(x) + rank(y)
    out["rd_sum_ranks"] = np.array(s.ranks)
    for tag, kw in [("cap3", dict(ranks_new=3)), ("eps", dict(rel_error=1e-10)),
                    ("caplist", dict(ranks_new=[2, 3, 5, 3, 1], rel_error=1e-12)), ("default", dict())]:
        t = s.copy().reapprox(**kw)
        out[f"rd_sum_{tag}_ranks"] = np.array(t.ranks)
        out[f"rd_sum_{tag}_dense"] = t.eval(squeeze=False)
    # a sum whose true rank is known: x + 2x has the ranks of x once rounded at 1e-10
    d = (T(x) + 2.0 * T(x)).reapprox(rel_error=1e-10)
    out["rd_dup_ranks"] = np.array(d.ranks)
    out["rd_dup_dense"] = d.eval(squeeze=False)
    # matrix-valued (4-D cores) rounding as used by operator assembly
    mpo = random_tt(rng, [(2, 2)] * 5, [3, 5, 5, 3])
    out.update(pack_cores("rd_mpo", mpo))
    m2 = (T(mpo) + T(mpo)).reapprox(rel_error=1e-10)
    out["rd_mpo_ranks"] = np.array(m2.ranks)
    out["rd_mpo_dense"] = m2.eval(squeeze=False)
    # zero TT through the sweep (sigma_0 = 0 -> nan rule keeps everything)
    z = tm.zeros([(2,)] * 4) + T(random_tt(rng, [(2,)] * 4, [2, 2, 2]))
    z = (z - z.copy())
    out["rd_zero_ranks_before"] = np.array(z.ranks)
    z.reapprox(ranks_new=3)
    out["rd_zero_ranks"] = np.array(z.ranks)
    out["rd_zero_norm"] = np.float64(np.abs(z.eval()).max())
    save("tensor_algebra", **out)


def gen_operators(tm, solver, bpx2D, utils):
    """Dense evaluations of every assembled operator at small L: pins the package's fresh assembly."""
    out = {}
    for L in (2, 3, 4):
        out[f"M_{L}"] = solver.get_overlap_matrix_as_tt(L).eval(squeeze=False)
        out[f"D_{L}"] = solver.get_derivative_matrix_as_tt(L).eval(squeeze=False)
        out[f"A_{L}"] = solver.get_laplace_matrix_as_tt(L).eval(squeeze=False)
        out[f"A_{L}_ranks"] = np.array(solver.get_laplace_matrix_as_tt(L).ranks)
        out[f"C_{L}"] = solver.get_bpx_preconditioner(L).eval(squeeze=False)
        out[f"Qp_{L}"] = solver.get_bpx_Qp(L).eval(squeeze=False)
        out[f"Qt_{L}"] = solver.get_bpx_Qt(L).eval(squeeze=False)
        out[f"B_{L}"] = solver.get_laplace_bpx(L).eval(squeeze=False)
        out[f"Rbpx_{L}"] = solver.get_rhs_matrix_bpx(L).eval(squeeze=False)
        out[f"Rnodal_{L}"] = solver.get_rhs_matrix_as_tt(L).eval(squeeze=False)
        out[f"G_{L}"] = solver.get_gram_matrix_as_tt(L).eval(squeeze=False)
        out[f"Gc_{L}"] = solver.get_gram_matrix_as_tt(L, "corner").eval(squeeze=False)
        out[f"mass_{L}"] = solver.build_mass_matrix_in_nodal_basis(L).eval(squeeze=False)
        for l in range(L + 1):
            out[f"P_{L}_{l}"] = solver.get_level_mapping_matrix_as_tt(L, l).eval(squeeze=False)
        out[f"Cbysum_{L}"] = solver.get_bpx_preconditioner_by_sum(L).eval(squeeze=False)
    for L in (2, 3):
        out[f"A2D_{L}"] = solver.build_laplace_matrix_2D(L).eval(squeeze=False)
        out[f"mass2D_{L}"] = solver.build_2D_mass_matrix(L).eval(squeeze=False)
        out[f"R2D_{L}"] = solver.get_rhs_matrix_as_tt_2D(L).eval(squeeze=False)
        out[f"C2D_{L}"] = bpx2D.get_BPX_preconditioner_2D(L).eval(squeeze=False)
        out[f"Qpx2D_{L}"] = bpx2D.get_BPX_Qp_2D(L, "x").eval(squeeze=False)
        out[f"Qpy2D_{L}"] = bpx2D.get_BPX_Qp_2D(L, "y").eval(squeeze=False)
        out[f"Q2D_{L}"] = bpx2D.get_BPX_Q_2D(L).eval(squeeze=False)
        out[f"B2D_{L}"] = bpx2D.get_laplace_BPX_2D(L).eval(squeeze=False)
        out[f"Rbpx2D_{L}"] = bpx2D.get_rhs_matrix_BPX_2D(L).eval(squeeze=False)
        out[f"C2Dbysum_{L}"] = solver.get_bpx_preconditioner_by_sum_2D(L).eval(squeeze=False)
    # rank profiles of the rounded system matrices at the config sizes that are cheap
    out["B1D_20_ranks"] = np.array(solver.get_laplace_bpx(20).reapprox(rel_error=1e-15).ranks)
    B5 = bpx2D.get_laplace_BPX_2D(5)
    out["B2D_5_raw_ranks"] = np.array(B5.ranks)
    out["B2D_5_ranks"] = np.array(B5.reapprox(rel_error=1e-12).ranks)
    # condition numbers: dense SVD of the dense evaluation
    conds = []
    for L in range(2, 9):
        sv = np.linalg.svd(solver.get_laplace_bpx(L).evalm(), compute_uv=False)
        conds.append(sv[0] / sv[-1])
    out["cond_B1D_L2_8"] = np.array(conds)
    conds = []
    for L in range(2, 5):
        sv = np.linalg.svd(bpx2D.get_laplace_BPX_2D(L).evalm(), compute_uv=False)
        conds.append(sv[0] / sv[-1])
    out["cond_B2D_L2_4"] = np.array(conds)
    save("operators", **out)


def gen_functions(tm, solver, bpx2D, utils):
    """Function encoders and the analytic KATs (experiments/1D_norm_of_interpolant.py, 2D_...)."""
    out = {}
    for L in (3, 5):
        out[f"poly_{L}"] = utils.get_polynomial_as_tt([1.0, -3.0, 2.0], L).eval(squeeze=False)
        out[f"trig_{L}"] = utils.get_trig_function_as_tt([0.3, 1.0, 1.0], L).eval(squeeze=False)
        out[f"u1d_corner_{L}"] = utils.get_example_u_1D(L, "corner").eval(squeeze=False)
        out[f"u1d_nodal_{L}"] = utils.get_example_u_1D(L, "nodal").eval(squeeze=False)
        out[f"f1d_{L}"] = utils.get_example_f_1D(L).eval(squeeze=False)
        out[f"du1d_{L}"] = utils.get_example_u_deriv_1D(L, "corner").eval(squeeze=False)
        out[f"du1d_avg_{L}"] = utils.get_example_u_deriv_1D(L, "average").eval(squeeze=False)
    for L in (2, 3):
        out[f"u2d_corner_{L}"] = utils.get_example_u_2D(L, "corner").eval(squeeze=False)
        out[f"u2d_nodal_{L}"] = utils.get_example_u_2D(L, "nodal").eval(squeeze=False)
        out[f"f2d_{L}"] = utils.get_example_f_2D(L).eval(squeeze=False)
        Dx, Dy = utils.get_example_grad_u_2D(L)
        out[f"gradx2d_{L}"] = Dx.eval(squeeze=False)
        out[f"grady2d_{L}"] = Dy.eval(squeeze=False)
    # 1D KAT: L2 norm and orthogonalised H1 norm of the interpolant of u = (1-3x+2x^2) sin(2 pi x)
    kat = []
    for L in (10, 12, 20, 30):
        u = utils.get_example_u_1D(L, basis="nodal")
        l2 = solver.get_L2_norm_1D(u)
        D = solver.get_derivative_matrix_as_tt(L)
        h1 = (D @ u).norm_squared() * 0.5 ** L
        kat.append([L, l2, h1])
    out["kat1d"] = np.array(kat)
    # 2D KAT at L = 6 (max_rank 40 as in experiments/2D_norm_of_interpolant.py)
    kat = []
    for L in (4, 6):
        u = utils.get_example_u_2D(L, basis="nodal").reapprox(ranks_new=40)
        l2 = solver.get_L2_norm_2D(u)
        kat.append([L, l2])
        out[f"u2d_nodal_ranks_{L}"] = np.array(u.ranks)
    out["kat2d"] = np.array(kat)
    save("functions", **out)


def gen_solves(tm, solver, bpx2D, utils):
    """Solve-level goldens for `solve_with_grad_descent` (the numpy-only solver, solver.py:251-275)."""
    out = {}
    import io, contextlib
    # scratchpad/test_amen.py:8-27 style random SPD system, L = 3
    np.random.seed(1)
    L = 3
    rng = np.random.default_rng(5)
    Wc = random_tt(rng, [(2, 2)] * L, [2, 2])
    W = tm.TensorTrain([c.copy() for c in Wc])
    A = (W.copy().transpose() @ W)
    for k in range(L):
        A.tensors[k] = A.tensors[k] + 0.0
    I = utils._get_identy_as_tt(L)
    A = (A + 0.5 * I).reapprox(rel_error=1e-14)
    b = tm.TensorTrain(random_tt(rng, [(2,)] * L, [2, 2]))
    out.update(pack_cores("spd_A", A.tensors)); out.update(pack_cores("spd_b", b.tensors))
    with contextlib.redirect_stdout(io.StringIO()):
        x, r2 = solver.solve_with_grad_descent(A, b, n_steps_max=60, max_rank=4)
    out["spd_x_dense"] = x.eval(squeeze=False); out["spd_r2"] = np.float64(r2)
    out["spd_x_ranks"] = np.array(x.ranks)

    # C1-like: 1D, BPX, f = 1 at L = 8 and L = 12 (cap binds), fixed number of steps
    for L, steps, cap in ((8, 40, 6), (12, 30, 8)):
        f = utils.get_polynomial_as_tt([1.0], L)
        B = solver.get_laplace_bpx(L).reapprox(rel_error=1e-15)
        b = (solver.get_rhs_matrix_bpx(L) @ f).squeeze()
        for i in range(L):
            b.tensors[i] /= 2
        with contextlib.redirect_stdout(io.StringIO()):
            v, r2 = solver.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
        C = solver.get_bpx_preconditioner(L)
        u = (C @ v).reapprox(rel_error=1e-12)
        Du = (solver.get_bpx_Qp(L) @ v).reapprox(rel_error=1e-12)
        out[f"c1_{L}_B_ranks"] = np.array(B.ranks)
        out[f"c1_{L}_v_ranks"] = np.array(v.ranks)
        out[f"c1_{L}_r2"] = np.float64(r2)
        out[f"c1_{L}_v_norm2"] = np.float64(v.norm_squared())
        out[f"c1_{L}_energy"] = np.float64((v.copy().transpose() @ B @ v).squeeze().eval()) if False else np.float64(
            (tm.TensorTrain([c.copy() for c in v.tensors]) @ (B @ v)).squeeze().eval())
        out[f"c1_{L}_l2"] = np.float64(solver.get_L2_norm_1D(u))
        out[f"c1_{L}_h1"] = np.float64(Du.norm_squared() * 0.5 ** L)
        if L <= 8:
            out[f"c1_{L}_v_dense"] = v.eval(squeeze=False)
            out[f"c1_{L}_u_dense"] = u.eval(squeeze=False)

    # C3-like: 2D BPX at L = 3 and 4, example f, cap 8
    for L, steps, cap in ((3, 30, 8), (4, 20, 6)):
        f = utils.get_example_f_2D(L).reapprox(ranks_new=cap)
        f = f.copy().flatten_mode_indices()
        B = bpx2D.get_laplace_BPX_2D(L).reapprox(rel_error=1e-12)
        b = (bpx2D.get_rhs_matrix_BPX_2D(L) @ f).squeeze().reapprox(ranks_new=cap, rel_error=1e-12)
        with contextlib.redirect_stdout(io.StringIO()):
            v, r2 = solver.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
        C = bpx2D.get_BPX_preconditioner_2D(L)
        u = (C @ v).reapprox(ranks_new=60, rel_error=1e-12)
        v2 = v.copy(); v2.tensors.append(np.ones([1, 1, 1, 1]))
        Dx = (bpx2D.get_BPX_Qp_2D(L, "x") @ v2).flatten_mode_indices()
        Dy = (bpx2D.get_BPX_Qp_2D(L, "y") @ v2).flatten_mode_indices()
        I = tm.TensorTrain([np.eye(2)[None, :, :, None] for _ in range(L)] + [np.eye(1)[None, :, :, None]])
        G = utils._get_gram_matrix_tt(L); G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)
        Gy = utils.kronecker_prod_2D(I, G); Gx = utils.kronecker_prod_2D(G, I)
        energy = 0.25 ** L * ((Gy @ Dx).norm_squared() + (Gx @ Dy).norm_squared())
        out[f"c3_{L}_B_ranks"] = np.array(B.ranks); out[f"c3_{L}_b_ranks"] = np.array(b.ranks)
        out[f"c3_{L}_v_ranks"] = np.array(v.ranks); out[f"c3_{L}_r2"] = np.float64(r2)
        out[f"c3_{L}_energy"] = np.float64(energy)
        out[f"c3_{L}_l2"] = np.float64(solver.get_L2_norm_2D(u))
        out[f"c3_{L}_v_dense"] = v.eval(squeeze=False)
        out[f"c3_{L}_vBv"] = np.float64((tm.TensorTrain([c.copy() for c in v.tensors]) @ (B @ v)).squeeze().eval())
    save("solves", **out)


def gen_heavy(tm, solver, bpx2D, utils, n_inst, steps):
    """2D L=12 (config C3/C5) fixtures: B rank profile, probes of B, one round(B@x), C5 solves."""
    import io, contextlib
    out = {}
    L = 12
    t0 = time.time()
    B = bpx2D.get_laplace_BPX_2D(L)
    out["B_raw_ranks"] = np.array(B.ranks)
    B.reapprox(rel_error=1e-12)
    out["B_ranks"] = np.array(B.ranks)
    print("B built", B.ranks, time.time() - t0, flush=True)
    rng = np.random.default_rng(77)
    probes = [random_tt(rng, [(4,)] * L, [2] * (L - 1), scale=0.5) for _ in range(3)]
    T = tm.TensorTrain
    # probes of the operator itself: p_i^T B p_j
    pBp = np.zeros((3, 3))
    for i in range(3):
        Bp = B @ T(probes[i])
        for j in range(3):
            pBp[j, i] = float((T(probes[j]) @ Bp).squeeze().eval())
    out["pBp"] = pBp
    R = bpx2D.get_rhs_matrix_BPX_2D(L)
    C = bpx2D.get_BPX_preconditioner_2D(L)
    # round(B @ x) at x rank 2, 4 and 8
    for r in (2, 4, 8):
        x = random_tt(np.random.default_rng(100 + r), [(4,)] * L, [r] * (L - 1), scale=0.5)
        y = (B @ T(x)).reapprox(ranks_new=r, rel_error=1e-10)
        out[f"round_r{r}_ranks"] = np.array(y.ranks)
        out[f"round_r{r}_norm2"] = np.float64(y.norm_squared())
        out[f"round_r{r}_probes"] = np.array([float((T(p) @ y).squeeze().eval()) for p in probes])
        print("round", r, y.ranks, time.time() - t0, flush=True)
    save("heavy_2d_L12_ops", **out)

    # C5 instances: rank-2 random f_i, b_i = round(R f_i, cap, 1e-12), `steps` GD steps at cap 4
    cap = 4
    res = {}
    for i in range(n_inst):
        f = T(c5_rhs_cores(i, L))
        b = (R @ f).squeeze().reapprox(ranks_new=cap, rel_error=1e-12)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            v, r2 = solver.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
        u = (C @ v).reapprox(ranks_new=cap, rel_error=1e-12)
        res[f"inst{i}_b_ranks"] = np.array(b.ranks)
        res[f"inst{i}_b_norm2"] = np.float64(b.norm_squared())
        res[f"inst{i}_v_ranks"] = np.array(v.ranks)
        res[f"inst{i}_r2"] = np.float64(r2)
        res[f"inst{i}_v_norm2"] = np.float64(v.norm_squared())
        res[f"inst{i}_v_probes"] = np.array([float((T(p) @ v).squeeze().eval()) for p in probes])
        res[f"inst{i}_u_l2"] = np.float64(solver.get_L2_norm_2D(u))
        print("instance", i, v.ranks, r2, time.time() - t0, flush=True)
        res["steps"] = np.int64(steps); res["n_inst"] = np.int64(i + 1)
        save(f"heavy_2d_L12_c5_steps{steps}", **res)


# ---------------------------------------------------------------------------------------------------------
# Round-2 fixtures: the BASELINE.json configs at their STATED sizes.
# Each record is written to its own small .npz so several generator processes can run side by side:
#     OPENBLAS_NUM_THREADS=2 python -m oracle.make_golden --config c5 --cap 4 --steps 100 --inst 0:6
#     python -m oracle.make_golden --merge c5_cap4_steps100
# ---------------------------------------------------------------------------------------------------------
def _quiet():
    import io, contextlib
    return contextlib.redirect_stdout(io.StringIO())


def _ops_2d(tm, solver, bpx2D, utils, L, eps):
    """B (rounded), R, C, Q'_x, Q'_y, G_x, G_y of solve_PDE_2D_with_preconditioner / 2D_sweep_L.py:44-49."""
    B = bpx2D.get_laplace_BPX_2D(L)
    B.reapprox(rel_error=eps)
    R = bpx2D.get_rhs_matrix_BPX_2D(L)
    C = bpx2D.get_BPX_preconditioner_2D(L)
    Qx, Qy = bpx2D.get_BPX_Qp_2D(L, "x"), bpx2D.get_BPX_Qp_2D(L, "y")
    G = utils._get_gram_matrix_tt(L)
    G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)
    I = utils._get_identy_as_tt(L, True)
    Gx, Gy = utils.kronecker_prod_2D(G, I), utils.kronecker_prod_2D(I, G)
    return dict(B=B, R=R, C=C, Qx=Qx, Qy=Qy, Gx=Gx, Gy=Gy)


def solve_record_2d(tm, solver, ops, f_flat, L, cap, steps, eps, probes, prefix):
    """One solve of config C3/C5 by the reference's own functions (solver.py:212-233 with solve_with_grad_descent,
    solver.py:251-275, in place of the absent ttpy), and the contracted quantities the parity tests compare."""
    T = tm.TensorTrain
    out = {}
    b = (ops["R"] @ f_flat).squeeze().reapprox(ranks_new=cap, rel_error=eps)
    with _quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v, r2 = solver.solve_with_grad_descent(ops["B"], b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
    u = (ops["C"] @ v).reapprox(ranks_new=cap, rel_error=eps)
    v_ext = v.copy()
    v_ext.tensors.append(np.ones([1, 1, 1, 1]))
    Dx = (ops["Qx"] @ v_ext).flatten_mode_indices()
    Dy = (ops["Qy"] @ v_ext).flatten_mode_indices()
    energy = 0.25 ** L * ((ops["Gy"] @ Dx).norm_squared() + (ops["Gx"] @ Dy).norm_squared())
    out[f"{prefix}_b_ranks"] = np.array(b.ranks)
    out[f"{prefix}_b_norm2"] = np.float64(b.norm_squared())
    out[f"{prefix}_v_ranks"] = np.array(v.ranks)
    out[f"{prefix}_u_ranks"] = np.array(u.ranks)
    out[f"{prefix}_r2"] = np.float64(r2)
    out[f"{prefix}_v_norm2"] = np.float64(v.norm_squared())
    out[f"{prefix}_v_probes"] = np.array([float((T(p) @ v).squeeze().eval()) for p in probes])
    out[f"{prefix}_u_l2"] = np.float64(solver.get_L2_norm_2D(u))
    out[f"{prefix}_energy"] = np.float64(energy)
    return out, (v, u, Dx, Dy)


def gen_config(tm, solver, bpx2D, utils, args):
    T = tm.TensorTrain
    t0 = time.time()
    cfg = args.config
    if cfg in ("c5", "c3"):
        L, eps, cap, steps = args.L, 1e-12, args.cap, args.steps
        ops = _ops_2d(tm, solver, bpx2D, utils, L, eps)
        print("operators built", ops["B"].ranks, time.time() - t0, flush=True)
        probes = [random_tt(np.random.default_rng(77), [(4,)] * L, [2] * (L - 1), scale=0.5) for _ in range(1)]
        rng = np.random.default_rng(77)
        probes = [random_tt(rng, [(4,)] * L, [2] * (L - 1), scale=0.5) for _ in range(3)]
    if cfg == "c5":
        lo, hi = [int(s) for s in args.inst.split(":")]
        for i in range(lo, hi):
            cores = c5_rhs_cores(i, L)
            if args.perturb > 0:       # conditioning probe: the same instance with every input entry moved by ~1 ulp
                prng = np.random.default_rng(99 + i)
                cores = [c * (1.0 + args.perturb * prng.standard_normal(c.shape)) for c in cores]
            f = T(cores)
            rec, _ = solve_record_2d(tm, solver, ops, f, L, cap, steps, eps, probes, f"inst{i}")
            rec["threads"] = np.int64(int(os.environ.get("OPENBLAS_NUM_THREADS", "0")))
            save(f"parts/c5_cap{cap}_steps{steps}{args.tag}_inst{i}", **rec)
            print("instance", i, rec[f"inst{i}_v_ranks"], rec[f"inst{i}_r2"], rec[f"inst{i}_energy"], time.time() - t0, flush=True)
    elif cfg == "c3":
        # 2D_sweep_L.py:38-73 at L = 12 with the example right-hand side; `max_rank` = cap for f, b, v and u alike
        f = utils.get_example_f_2D(L).reapprox(ranks_new=cap)
        f = f.copy().flatten_mode_indices()
        rec, (v, u, Dx, Dy) = solve_record_2d(tm, solver, ops, f, L, cap, steps, eps, probes, "c3")
        refnorm_L2 = (1 / 15 + 3 / (16 * np.pi ** 4) - 3 / (16 * np.pi ** 2)) * (67 / 210)
        refnorm_H1 = (2144 * np.pi ** 6 + 5926 * np.pi ** 4 + 7245 - 9255 * np.pi ** 2) / (25200 * np.pi ** 4)
        u_ref = utils.get_example_u_2D(L, basis="nodal").reshape_mode_indices([4])
        Dx_ref, Dy_ref = utils.get_example_grad_u_2D(L)
        max_rank = 60                                   # the rounding of the differences must not hide the error
        delta = (u - u_ref).reshape_mode_indices([4]).reapprox(rel_error=eps, ranks_new=max_rank)
        rec["c3_delta_ranks"] = np.array(delta.ranks)
        rec["c3_l2_err2"] = np.float64(solver.get_L2_norm_2D(delta))
        dDx = (Dx - Dx_ref).reapprox(rel_error=eps, ranks_new=max_rank)
        dDy = (Dy - Dy_ref).reapprox(rel_error=eps, ranks_new=max_rank)
        rec["c3_h1_err2"] = np.float64(0.25 ** L * ((ops["Gy"] @ dDx).norm_squared() + (ops["Gx"] @ dDy).norm_squared()))
        rec["c3_refnorm_L2"] = np.float64(refnorm_L2); rec["c3_refnorm_H1"] = np.float64(refnorm_H1)
        rec["c3_f_ranks"] = np.array(f.ranks)
        save(f"c3_L{L}_cap{cap}_steps{steps}", **rec)
        print("c3", cap, rec["c3_v_ranks"], rec["c3_r2"], rec["c3_energy"] / refnorm_H1 - 1, rec["c3_u_l2"] / refnorm_L2 - 1,
              rec["c3_l2_err2"], rec["c3_h1_err2"], time.time() - t0, flush=True)
    elif cfg == "c1":
        # configs C1 / C2: 1D, f = 1, BPX, steepest descent with the reference defaults (cap 20), solver.py:192-210
        rec = {}
        for L in [int(s) for s in args.levels.split(",")]:
            cap, steps = args.cap, args.steps
            f = utils.get_polynomial_as_tt([1.0], L)
            B = solver.get_laplace_bpx(L).reapprox(rel_error=1e-15)
            b = (solver.get_rhs_matrix_bpx(L) @ f).squeeze()
            for i in range(L):
                b.tensors[i] /= 2
            with _quiet(), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                v, r2 = solver.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
            Cm = solver.get_bpx_preconditioner(L)
            u = (Cm @ v).reapprox(rel_error=1e-15)
            Du = (solver.get_bpx_Qp(L) @ v).reapprox(rel_error=1e-15)
            p = f"c1_{L}"
            rec[f"{p}_B_ranks"] = np.array(B.ranks); rec[f"{p}_b_ranks"] = np.array(b.ranks)
            rec[f"{p}_v_ranks"] = np.array(v.ranks); rec[f"{p}_r2"] = np.float64(r2)
            rec[f"{p}_v_norm2"] = np.float64(v.norm_squared())
            rec[f"{p}_l2"] = np.float64(solver.get_L2_norm_1D(u))
            rec[f"{p}_h1"] = np.float64(Du.norm_squared() * 0.5 ** L)
            # rank profiles of u / Du at 1e-15 are decided by iteration noise; at 1e-10 they are the exact ones
            rec[f"{p}_u_ranks_1e10"] = np.array((Cm @ v).reapprox(rel_error=1e-10).ranks)
            print("c1", L, v.ranks, r2, rec[f"{p}_l2"] / (2 / 15) - 1, rec[f"{p}_h1"] * 3 - 1, time.time() - t0, flush=True)
            rec["cap"] = np.int64(cap); rec["steps"] = np.int64(steps)
            save(f"c1_cap{cap}_steps{steps}{args.tag}", **rec)


def gen_experiments(tm, solver, bpx2D, utils):
    """Output arrays of src/experiments/1D_norm_of_interpolant.py computed by the unmodified reference (the script's own loop,
    :16-25, without the plotting): the acceptance test compares the same script run on this repository's package against them."""
    refnorm_L2 = -3 / (16 * np.pi ** 2) + 3 / (16 * np.pi ** 4) + 1 / 15
    refnorm_H1 = -1 / (4 * np.pi ** 2) + 5 / 12 + 4 * np.pi ** 2 / 15
    Ls = np.arange(3, 45)
    l2, h1n, h1 = [], [], []
    for L in Ls:
        A = solver.get_laplace_matrix_as_tt(L)
        D = solver.get_derivative_matrix_as_tt(L)
        u = utils.evaluate_nodal_basis(utils.get_example_u_1D(L), [1.0], basis='corner').squeeze()
        Du = D @ u
        l2.append(solver.get_L2_norm_1D(u) - refnorm_L2)
        h1n.append(float((u @ A @ u).squeeze().eval()) - refnorm_H1)
        h1.append(Du.norm_squared() * 0.5 ** L - refnorm_H1)
    save("exp_1D_norm_of_interpolant", L_values=Ls, error_L2=np.array(l2), error_H1=np.array(h1n), error_H1_smart=np.array(h1))


def merge_parts(stem):
    import glob
    out = {}
    files = sorted(glob.glob(os.path.join(GOLDEN, "parts", stem + "_inst*.npz")))
    ids = []
    for fn in files:
        z = np.load(fn)
        for k in z.files:
            if k != "threads":
                out[k] = z[k]
        ids.append(int(fn.rsplit("_inst", 1)[1].split(".")[0]))
    out["instances"] = np.array(sorted(ids))
    save(stem, **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heavy", action="store_true")
    ap.add_argument("--heavy-instances", type=int, default=4)
    ap.add_argument("--heavy-steps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--config", default="", help="c5 | c3 | c1 (round-2 fixtures at the stated config sizes)")
    ap.add_argument("--cap", type=int, default=4)
    ap.add_argument("--L", type=int, default=12)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--inst", default="0:1")
    ap.add_argument("--levels", default="20")
    ap.add_argument("--tag", default="")
    ap.add_argument("--perturb", type=float, default=0.0, help="relative size of a random perturbation of the input cores (conditioning probe)")
    ap.add_argument("--merge", default="")
    args = ap.parse_args()
    if args.merge:
        merge_parts(args.merge)
        return
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.ref_loader import load_reference
    tm, solver, bpx2D, utils = load_reference()
    warnings.simplefilter("ignore")
    if args.config:
        gen_config(tm, solver, bpx2D, utils, args)
        return
    if args.heavy:
        gen_heavy(tm, solver, bpx2D, utils, args.heavy_instances, args.heavy_steps)
        return
    todo = args.only.split(",") if args.only else ["algebra", "operators", "functions", "solves", "experiments"]
    if "algebra" in todo:
        gen_tensor_algebra(tm)
    if "operators" in todo:
        gen_operators(tm, solver, bpx2D, utils)
    if "functions" in todo:
        gen_functions(tm, solver, bpx2D, utils)
    if "solves" in todo:
        gen_solves(tm, solver, bpx2D, utils)
    if "experiments" in todo:
        gen_experiments(tm, solver, bpx2D, utils)


if __name__ == "__main__":
    main()
