"""Parity of the CUDA path (through the C ABI, csrc/libqtt_b200.so) with the reference: golden fixtures
produced by the unmodified reference and the CPU oracle on seeded inputs.  Only contracted quantities are
compared (dense evaluations, inner products, norms, rank lists) - QR / SVD signs are arbitrary.

Tolerances: FP64 throughout; dense evaluations of single operations agree to 1e-12 relative, solve-level
quantities (which iterate 20-100 truncated steps) to 1e-10 / 1e-9 relative, as stated per assertion."""
import numpy as np
import pytest
import torch

from conftest import cores_from, rel_err
from oracle import tt_oracle as O
from oracle.make_golden import random_tt, c5_rhs_cores

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tm():
    import laplacesolvermps_b200.tensormethods as tm
    return tm


@pytest.fixture(scope="module")
def eng():
    from laplacesolvermps_b200.batched import default_engine
    return default_engine()


def T(tm, cores):
    return tm.TensorTrain([np.array(c, copy=True) for c in cores])


# ---- TensorTrain drop-in API against the reference's goldens ---------------------------------------

def test_selftest_known_answer(golden, tm):
    g = golden("tensor_algebra")
    A = T(tm, cores_from(g, "selftest_cores"))
    inner = float((A @ A).squeeze().eval())                       # the reference's idiom
    assert abs(inner - 31.092539952213777) < 1e-12
    assert abs(A.norm_squared() - float(g["selftest_norm2"])) < 1e-12
    assert abs(A.inner(A) - inner) < 1e-12


def test_matmul_all_pairings(golden, tm):
    g = golden("tensor_algebra")
    a, b = T(tm, cores_from(g, "mm_mpo_a")), T(tm, cores_from(g, "mm_mpo_b"))
    r, l, l2 = T(tm, cores_from(g, "mm_mps_r")), T(tm, cores_from(g, "mm_mps_l")), T(tm, cores_from(g, "mm_mps_l2"))
    assert rel_err((a @ b).eval(squeeze=False), g["mm_44_dense"]) < 1e-14
    assert (a @ b).ranks == g["mm_44_ranks"].tolist()
    assert rel_err((a @ r).eval(squeeze=False), g["mm_43_dense"]) < 1e-14
    assert rel_err((l @ a).eval(squeeze=False), g["mm_34_dense"]) < 1e-14
    m33 = l @ l2
    assert [list(s) for s in m33.shapes] == g["mm_33_shapes"].tolist()
    assert abs(float(m33.squeeze().eval()) - float(g["mm_33_value"])) < 1e-13 * abs(float(g["mm_33_value"]))
    assert abs(float((l @ a @ r).squeeze().eval()) - float(g["mm_bilinear"])) < 1e-12 * abs(float(g["mm_bilinear"]))
    with pytest.raises(NotImplementedError):
        tm.TensorTrain([np.zeros((1, 2, 2, 2, 1))]) @ tm.TensorTrain([np.zeros((1, 2, 1))])
    with pytest.raises(AssertionError):
        a @ tm.TensorTrain(cores_from(g, "mm_mps_r")[:2])


def test_orthogonalise_and_round(golden, tm):
    g = golden("tensor_algebra")
    x, y = cores_from(g, "rd_x"), cores_from(g, "rd_y")
    xo = T(tm, x).left_orthogonalize()
    assert xo.ranks == g["rd_x_orth_ranks"].tolist()
    assert rel_err(xo.eval(squeeze=False), g["rd_x_orth_dense"]) < 1e-13
    for c in xo.tensors[1:]:
        m = c.reshape(c.shape[0], -1)
        assert np.abs(m @ m.T - np.eye(m.shape[0])).max() < 1e-13
    ro = T(tm, x).right_orthogonalize()
    assert rel_err(ro.eval(squeeze=False), g["rd_x_orth_dense"]) < 1e-13
    for c in ro.tensors[:-1]:
        m = c.reshape(-1, c.shape[-1])
        assert np.abs(m.T @ m - np.eye(m.shape[1])).max() < 1e-13
    assert abs(T(tm, x).norm_squared() - float(g["rd_x_norm2"])) < 1e-12 * float(g["rd_x_norm2"])
    s = T(tm, x) + T(tm, y)
    for tag, kw in [("cap3", dict(ranks_new=3)), ("eps", dict(rel_error=1e-10)),
                    ("caplist", dict(ranks_new=[2, 3, 5, 3, 1], rel_error=1e-12)), ("default", dict())]:
        t = s.copy().reapprox(**kw)
        assert t.ranks == g[f"rd_sum_{tag}_ranks"].tolist(), tag
        assert rel_err(t.eval(squeeze=False), g[f"rd_sum_{tag}_dense"]) < 1e-12, tag
    d = (T(tm, x) + 2.0 * T(tm, x)).reapprox(rel_error=1e-10)
    assert d.ranks == g["rd_dup_ranks"].tolist()
    assert rel_err(d.eval(squeeze=False), g["rd_dup_dense"]) < 1e-12
    mpo = T(tm, cores_from(g, "rd_mpo"))
    m2 = (mpo + mpo).reapprox(rel_error=1e-10)
    assert m2.ranks == g["rd_mpo_ranks"].tolist() and m2.shapes[1][1:3] == (2, 2)
    assert rel_err(m2.eval(squeeze=False), g["rd_mpo_dense"]) < 1e-12


def test_zero_train(golden, tm):
    g = golden("tensor_algebra")
    rng = np.random.default_rng(0)
    z = tm.zeros([(2,)] * 4) + tm.TensorTrain(random_tt(rng, [(2,)] * 4, [2, 2, 2]))
    z = z - z.copy()
    z.reapprox(ranks_new=3)
    assert z.ranks == g["rd_zero_ranks"].tolist()            # sigma_0 == 0: nothing is removed, cap applies
    assert np.abs(z.eval()).max() < 1e-13
    zz = tm.zeros([(2,)] * 5)
    assert zz.norm_squared() == 0.0
    assert zz.copy().reapprox(ranks_new=4).ranks == [1, 1, 1, 1]


# ---- batched kernels against the oracle on seeded ragged batches -----------------------------------

def test_batched_ragged_against_oracle(eng):
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    rng = np.random.default_rng(3)
    L, N = 6, 9
    xs = [random_tt(rng, [(2,)] * L, [int(rng.integers(1, 7)) for _ in range(L - 1)]) for _ in range(N)]
    ys = [random_tt(rng, [(2,)] * L, [int(rng.integers(1, 4)) for _ in range(L - 1)]) for _ in range(N)]
    X, Y = BatchedTT.from_host(xs, eng.device), BatchedTT.from_host(ys, eng.device)
    got = eng.inner(X, Y).cpu().numpy()
    ref = np.array([O.inner(a, b) for a, b in zip(xs, ys)])
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12
    n2 = eng.norm2(X).cpu().numpy()
    assert np.max(np.abs(n2 - [O.norm_squared(t) for t in xs]) / n2) < 1e-12
    al = torch.tensor(rng.standard_normal(N), device=eng.device)
    for caps, eps in ((2, 1e-16), (None, 1e-10), (3, 1e-12)):
        R = eng.round_sum([(None, X), (None, Y, al, 1.0)], caps=caps, rel_eps=eps).to_host()
        for i in range(N):
            ref = O.reapprox(O.add(xs[i], O.scale(ys[i], float(al[i]))), ranks_new=caps, rel_error=eps)
            assert [c.shape for c in R[i]] == [c.shape for c in ref]
            assert rel_err(O.dense(R[i], False), O.dense(ref, False)) < 1e-12
    S = eng.axpby(al, X, None, Y).to_host()
    for i in range(N):
        assert rel_err(O.dense(S[i], False), O.dense(O.add(O.scale(xs[i], float(al[i])), ys[i]), False)) < 1e-14


def test_batched_operator_terms_against_oracle(eng):
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    rng = np.random.default_rng(4)
    L, N = 8, 3
    op = random_tt(rng, [(2, 2)] * L, [5] * (L - 1))
    xs = [random_tt(rng, [(2,)] * L, [2, 4, 8, 16, 8, 4, 2]) for _ in range(N)]     # > 32 columns: several QR panels
    bs = [random_tt(rng, [(2,)] * L, [2] * (L - 1)) for _ in range(N)]
    A, X, B = DeviceMPO(op, eng.device), BatchedTT.from_host(xs, eng.device), BatchedTT.from_host(bs, eng.device)
    Ym = eng.matmul(A, X, [2] * L, [(2,)] * L).to_host()
    for i in range(N):
        assert rel_err(O.dense(Ym[i], False), O.dense(O.matmul(op, xs[i]), False)) < 1e-14
    status = torch.zeros(N, dtype=torch.int32, device=eng.device)
    R = eng.round_sum([(A, X)], caps=12, rel_eps=1e-14, status=status).to_host()
    assert int(status.abs().sum()) == 0
    for i in range(N):
        ref = O.reapprox(O.matmul(op, xs[i]), ranks_new=12, rel_error=1e-14)
        assert [c.shape for c in R[i]] == [c.shape for c in ref]
        assert rel_err(O.dense(R[i], False), O.dense(ref, False)) < 1e-12
    R = eng.round_sum([(None, B), (A, X, None, -1.0)], caps=6, rel_eps=1e-16).to_host()
    for i in range(N):
        ref = O.reapprox(O.sub(bs[i], O.matmul(op, xs[i])), ranks_new=6)
        assert [c.shape for c in R[i]] == [c.shape for c in ref]
        assert rel_err(O.dense(R[i], False), O.dense(ref, False)) < 1e-12
    n2 = eng.orthogonalize([(A, X)], want_cores=False)[1].cpu().numpy()
    ref = np.array([O.norm_squared(O.matmul(op, x)) for x in xs])
    assert np.max(np.abs(n2 - ref) / ref) < 1e-12
    bil = eng.inner(B, X, op=A).cpu().numpy()                      # <b, A x>, the reference's (b @ A @ x).squeeze().eval()
    ref = np.array([O.inner(b, O.matmul(op, x)) for b, x in zip(bs, xs)])
    assert np.max(np.abs(bil - ref) / np.abs(ref)) < 1e-11


def test_randomised_shapes_against_oracle(eng):
    """Seeded sweep over odd shapes: L = 2..7, mode sizes 1..4 (mixed inside one train), ranks 1..7, plain / operator /
    mixed sums, cap-limited and eps-limited roundings, orthogonalisation, norms and inner products."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    rng = np.random.default_rng(2024)
    for case in range(24):
        L = int(rng.integers(2, 8))
        n_out = [int(rng.integers(1, 5)) for _ in range(L)]
        n_in = [int(rng.integers(1, 4)) for _ in range(L)]
        N = int(rng.integers(1, 5))
        op = random_tt(rng, [(n_out[k], n_in[k]) for k in range(L)], [int(rng.integers(1, 5)) for _ in range(L - 1)], scale=0.8)
        xs = [random_tt(rng, [(n_in[k],) for k in range(L)], [int(rng.integers(1, 6)) for _ in range(L - 1)]) for _ in range(N)]
        ys = [random_tt(rng, [(n_out[k],) for k in range(L)], [int(rng.integers(1, 8)) for _ in range(L - 1)]) for _ in range(N)]
        A, X, Y = DeviceMPO(op, eng.device), BatchedTT.from_host(xs, eng.device), BatchedTT.from_host(ys, eng.device)
        caps = [None, int(rng.integers(1, 6)), [int(rng.integers(1, 6)) for _ in range(L - 1)]][case % 3]
        eps = [1e-16, 1e-10, 1e-12][(case // 3) % 3]
        coef = torch.tensor(rng.standard_normal(N), device=eng.device)
        R = eng.round_sum([(None, Y), (A, X, coef, 0.5)], caps=caps, rel_eps=eps).to_host()
        for i in range(N):
            full = O.add(ys[i], O.scale(O.matmul(op, xs[i]), 0.5 * float(coef[i])))
            ref, sig = O.reapprox(O.clone(full), ranks_new=caps, rel_error=eps, return_sigma=True)
            gap_ok = all(np.all(np.abs(np.log10(np.maximum(s_[1:] / s_[0], 1e-300)) - np.log10(eps)) > 0.5) for s_ in sig if len(s_) > 1)
            if gap_ok:                                                # rank decision away from the threshold
                assert [c.shape for c in R[i]] == [c.shape for c in ref], (case, i)
                assert rel_err(O.dense(R[i], False), O.dense(ref, False)) < 1e-11, (case, i)
            else:
                assert rel_err(O.dense(R[i], False), O.dense(full, False)) < 50 * max(eps, 1e-15) * L or caps is not None
        Q, n2 = eng.orthogonalize([(None, Y)])
        Qh = Q.to_host()
        for i in range(N):
            ref = O.left_orthogonalize(O.clone(ys[i]))
            assert [c.shape for c in Qh[i]] == [c.shape for c in ref], (case, i)
            assert rel_err(O.dense(Qh[i], False), O.dense(ys[i], False)) < 1e-12
            assert abs(float(n2[i]) - O.norm_squared(ys[i])) < 1e-12 * O.norm_squared(ys[i])
        n2p = eng.orthogonalize([(A, X)], want_cores=False)[1].cpu().numpy()
        ip = eng.inner(Y, X, op=A).cpu().numpy()
        for i in range(N):
            ax = O.matmul(op, xs[i])
            assert abs(n2p[i] - O.norm_squared(ax)) < 1e-11 * max(O.norm_squared(ax), 1e-300)
            assert abs(ip[i] - O.inner(ys[i], ax)) < 1e-11 * np.sqrt(O.norm_squared(ys[i]) * O.norm_squared(ax)) + 1e-300


def test_odd_leading_dimensions_take_the_unaligned_paths(eng):
    """Mode size 3 with odd ranks gives working matrices with odd row counts / leading dimensions: the 16-byte staging of
    larfb and the 16-byte reflector loads of the leaf must fall back to their 8-byte paths (multi-sub-panel panels, shared-memory
    resident reflectors with an odd column stride) and still match the oracle."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    rng = np.random.default_rng(11)
    L = 5
    for rk in (7, 25, 45):
        xs = [random_tt(rng, [(3,)] * L, [rk] * (L - 1), scale=0.4) for _ in range(2)]
        X = BatchedTT.from_host(xs, eng.device)
        n2 = eng.norm2(X).cpu().numpy()
        Y = eng.round_sum([(None, X)], caps=5, rel_eps=1e-14).to_host()
        for i in range(2):
            assert abs(n2[i] - O.norm_squared(xs[i])) < 1e-12 * n2[i]
            ref = O.reapprox(O.clone(xs[i]), ranks_new=5, rel_error=1e-14)
            assert [c.shape for c in Y[i]] == [c.shape for c in ref]
            assert rel_err(O.dense(Y[i], False), O.dense(ref, False)) < 1e-11
    op = random_tt(rng, [(3, 3)] * L, [9] * (L - 1), scale=0.3)
    vs = [random_tt(rng, [(3,)] * L, [5] * (L - 1), scale=0.5) for _ in range(2)]
    Y = eng.round_sum([(DeviceMPO(op, eng.device), BatchedTT.from_host(vs, eng.device))], caps=6, rel_eps=1e-14).to_host()
    for i in range(2):
        ref = O.reapprox(O.matmul(op, vs[i]), ranks_new=6, rel_error=1e-14)
        assert [c.shape for c in Y[i]] == [c.shape for c in ref]
        assert rel_err(O.dense(Y[i], False), O.dense(ref, False)) < 1e-11


def _padded(cores, r):
    """The same tensor written with rank-r cores (zero padding)."""
    L, out = len(cores), []
    for k, c in enumerate(cores):
        rl, rr = (c.shape[0] if k == 0 else r), (c.shape[-1] if k == L - 1 else r)
        p = np.zeros((rl,) + c.shape[1:-1] + (rr,))
        p[:c.shape[0], ..., :c.shape[-1]] = c
        out.append(p)
    return out


@pytest.mark.parametrize("case", ["plan_bound_large_actual_small", "actual_large_ranks_differ"])
def test_uncapped_rounding_with_data_dependent_ranks_inside_a_group(eng, case):
    """Rounding by eps only (no cap) of `op @ x` with a product bond of 180 at L = 18 (mode size 2): in the PLAN the SVD sides
    exceed the single-CTA kernel's 160 vectors (bounds grow like 2^k without a cap).  Instances share a rank signature but differ
    in numerical rank (one is a rank-1 tensor written with rank-3 cores), so the left ranks differ inside the group after the first
    truncations.  (a) operator of true rank 2 padded to rank 60: the actual unfoldings are small - the run-time check keeps the
    ragged-aware single-CTA kernel (the path `bench.py --config c1` takes in its post-processing); (b) random rank-60 operator: the
    unfoldings are really large (bonds 180 vs 60) - the multi-CTA Jacobi runs instance by instance.  Both used to fail with
    QTT_ERR_UNSUPPORTED ("large SVD needs uniform left ranks")."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    rng = np.random.default_rng(11)
    L = 18
    if case == "plan_bound_large_actual_small":
        op = _padded(random_tt(rng, [(2, 2)] * L, [2] * (L - 1), scale=0.7), 60)
    else:
        op = random_tt(rng, [(2, 2)] * L, [60] * (L - 1), scale=0.2)
    xs = [random_tt(rng, [(2,)] * L, [3] * (L - 1)), _padded(random_tt(rng, [(2,)] * L, [1] * (L - 1)), 3)]
    A, X = DeviceMPO(op, eng.device), BatchedTT.from_host(xs, eng.device)
    R = eng.round_sum([(A, X)], caps=None, rel_eps=1e-10).to_host()
    profiles = []
    for i in range(len(xs)):
        ref = O.reapprox(O.matmul(op, xs[i]), rel_error=1e-10)
        profiles.append([c.shape[-1] for c in ref])
        assert [c.shape for c in R[i]] == [c.shape for c in ref]
        assert abs(O.inner(R[i], ref) / O.norm_squared(ref) - 1) < 1e-9 and abs(O.norm_squared(R[i]) / O.norm_squared(ref) - 1) < 1e-9
    assert profiles[0] != profiles[1]                       # the point of the test: ranks differ inside one group


def test_graph_cache_replays_bit_identically(monkeypatch):
    """Opt-in CUDA-graph cache of launch-bound sweeps (QTT_GRAPHS=1, csrc/qtt_b200.cu: sweep_group): inside a solver the same sweep
    (same buffers, same ranks) recurs every other iteration; it is captured on second sight and replayed afterwards.  A 1D
    steepest-descent solve with the cache on must replay sweeps and return exactly the iterate of the uncached path."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    from laplacesolvermps_b200 import _assembly as A
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
    L, N = 10, 3
    B = S.get_laplace_bpx(L)
    bs = []
    for i in range(N):
        f = U.get_polynomial_as_tt([1.0, 0.5 * i, -0.25 * (i + 1)], L)
        bs.append(O.squeeze(A.compose(S.get_rhs_matrix_bpx(L).tensors, f.tensors)))
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("QTT_GRAPHS", flag)
        e = Engine("cuda:0")
        monkeypatch.delenv("QTT_GRAPHS")
        x, r2, st, _ = e.gd_solve(DeviceMPO(B.tensors, e.device), BatchedTT.from_host(bs, e.device), n_steps=24, max_rank=6, rel_accuracy=0.0)
        out[flag] = ([c.cpu().numpy().copy() for c in x.cores], r2.cpu().numpy().copy(), e.graph_counters())
        e.close()
    assert out["0"][2] == (0, 0)
    captures, replays = out["1"][2]
    assert captures > 0 and replays > captures, (captures, replays)
    assert np.array_equal(out["0"][1], out["1"][1])
    for a, b in zip(out["0"][0], out["1"][0]):
        assert np.array_equal(a, b)


def test_tall_skinny_qr_path_against_oracle(monkeypatch):
    """Opt-in flat-tree (tall-skinny) panel QR (csrc/tsqr.cuh; QTT_TS_MIN_ROWS read at handle creation): orthogonalisation,
    norms and roundings of rank-100 trains (400-row panels -> two row blocks per 64-column panel) against the oracle, plus a
    fused round(A @ x) whose product bond is 120."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
    monkeypatch.setenv("QTT_TS_MIN_ROWS", "330")
    e2 = Engine("cuda:0")
    monkeypatch.delenv("QTT_TS_MIN_ROWS")
    rng = np.random.default_rng(5)
    L, N = 7, 2
    xs = [random_tt(rng, [(4,)] * L, [100] * (L - 1), scale=0.3) for _ in range(N)]
    X = BatchedTT.from_host(xs, e2.device)
    n2 = e2.norm2(X).cpu().numpy()
    for i in range(N):
        assert abs(n2[i] - O.norm_squared(xs[i])) < 1e-12 * n2[i]
    Q, _ = e2.orthogonalize([(None, X)])
    Qh = Q.to_host()
    for i in range(N):
        assert rel_err(O.dense(Qh[i], False), O.dense(xs[i], False)) < 1e-12
        for c in Qh[i][1:]:
            m = c.reshape(c.shape[0], -1)
            assert np.abs(m @ m.T - np.eye(m.shape[0])).max() < 1e-12
    for kw in (dict(caps=5, rel_eps=1e-16), dict(caps=None, rel_eps=1e-8)):
        Rh = e2.round_sum([(None, X)], **kw).to_host()
        for i in range(N):
            ref = O.reapprox(O.clone(xs[i]), ranks_new=kw["caps"], rel_error=kw["rel_eps"])
            assert [c.shape for c in Rh[i]] == [c.shape for c in ref]
            assert rel_err(O.dense(Rh[i], False), O.dense(ref, False)) < 1e-11
    op = random_tt(rng, [(4, 4)] * L, [30] * (L - 1), scale=0.2)
    vs = [random_tt(rng, [(4,)] * L, [4] * (L - 1), scale=0.5) for _ in range(N)]
    Y = e2.round_sum([(DeviceMPO(op, e2.device), BatchedTT.from_host(vs, e2.device))], caps=4, rel_eps=1e-14).to_host()
    for i in range(N):
        ref = O.reapprox(O.matmul(op, vs[i]), ranks_new=4, rel_error=1e-14)
        assert rel_err(O.dense(Y[i], False), O.dense(ref, False)) < 1e-11
    e2.close()


def test_integration_stub_from_the_docs_runs():
    """INTEGRATION.md section 2 shows the ctypes stub a maintainer of the reference would add; execute that very code
    block against the built library so the document cannot rot."""
    import os, re
    from conftest import ROOT
    from laplacesolvermps_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# laplace_mps/_qtt.py.*?```", text, flags=re.S).group(0)
    code = block[len("```python\n"):-3].replace('C.CDLL("libqtt_b200.so")', f'C.CDLL(r"{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    rng = np.random.default_rng(1)
    x = random_tt(rng, [(2,)] * 5, [3, 6, 6, 3])
    y = O.add(x, O.scale(x, 0.5))
    out = ns["reapprox"]([c.copy() for c in y], [4] * 4, 1e-12)
    ref = O.reapprox(O.clone(y), ranks_new=4, rel_error=1e-12)
    assert [c.shape for c in out] == [c.shape for c in ref]
    assert rel_err(O.dense(out, False), O.dense(ref, False)) < 1e-12


def test_sharding_does_not_change_arithmetic(eng):
    """Multi-GPU = same batch split over ranks: per-instance results must be bit-identical."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    from laplacesolvermps_b200.parallel import shard_range
    rng = np.random.default_rng(5)
    L, N = 6, 8
    op = random_tt(rng, [(2, 2)] * L, [4] * (L - 1))
    xs = [random_tt(rng, [(2,)] * L, [3] * (L - 1)) for _ in range(N)]
    A = DeviceMPO(op, eng.device)
    full = eng.round_sum([(A, BatchedTT.from_host(xs, eng.device))], caps=3).to_host()
    for world in (2, 4):
        for r in range(world):
            lo, hi = shard_range(N, r, world)
            part = eng.round_sum([(A, BatchedTT.from_host(xs[lo:hi], eng.device))], caps=3).to_host()
            for i in range(lo, hi):
                for a, b in zip(full[i], part[i - lo]):
                    assert np.array_equal(a, b)


# ---- solve-level parity ---------------------------------------------------------------------------

def test_grad_descent_spd_golden(golden, tm, eng):
    """scratchpad/test_amen.py-style random SPD system (cond ~ 1e3).  Steepest descent with lr = 0.8 is not
    contractive on it: round-off differences grow by ~10x every 4 steps (the oracle shows the same against
    itself on another BLAS), so the per-step history is compared while it is still meaningful (first 12
    steps, 1e-9) and the end state only loosely."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    import laplacesolvermps_b200.solver as S
    g = golden("solves")
    Ac, bc = cores_from(g, "spd_A"), cores_from(g, "spd_b")
    hist_ref = []
    O.solve_with_grad_descent(O.clone(Ac), O.clone(bc), n_steps_max=12, max_rank=4, history=hist_ref)
    _, _, _, hist = eng.gd_solve(DeviceMPO(Ac, eng.device), BatchedTT.from_host([bc], eng.device), n_steps=12, max_rank=4,
                                 rel_accuracy=1e-14, want_history=True)
    h = hist.cpu().numpy()[:, 0, :]
    for n, r2_ref, gamma_ref in hist_ref:
        assert abs(h[n, 0] - r2_ref) < 1e-9 * r2_ref and abs(h[n, 1] - gamma_ref) < 1e-9 * abs(gamma_ref), n
    x, r2 = S.solve_with_grad_descent(T(tm, Ac), T(tm, bc), n_steps_max=60, max_rank=4)
    assert x.ranks == g["spd_x_ranks"].tolist()
    assert rel_err(x.eval(squeeze=False), g["spd_x_dense"]) < 1e-3


@pytest.mark.parametrize("L,steps,cap", [(8, 40, 6), (12, 30, 8)])
def test_c1_1d_bpx_solve(golden, tm, L, steps, cap):
    """Config C1 at oracle-tractable size: 1D, BPX, f = 1 - solver.py:192-210 with the numpy-only solver."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    g = golden("solves")
    f = U.get_polynomial_as_tt([1.0], L)
    B = S.get_laplace_bpx(L)
    assert B.ranks == g[f"c1_{L}_B_ranks"].tolist()
    b = (S.get_rhs_matrix_bpx(L) @ f).squeeze()
    for i in range(L):
        b.tensors[i] = b.tensors[i] / 2
    v, r2 = S.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
    assert v.ranks == g[f"c1_{L}_v_ranks"].tolist()
    u = (S.get_bpx_preconditioner(L) @ v).reapprox(rel_error=1e-12)
    Du = (S.get_bpx_Qp(L) @ v).reapprox(rel_error=1e-12)
    assert abs(v.norm_squared() - float(g[f"c1_{L}_v_norm2"])) < 1e-10 * float(g[f"c1_{L}_v_norm2"])
    assert abs(v.inner(B @ v) - float(g[f"c1_{L}_energy"])) < 1e-10 * float(g[f"c1_{L}_energy"])
    assert abs(S.get_L2_norm_1D(u) - float(g[f"c1_{L}_l2"])) < 1e-10 * float(g[f"c1_{L}_l2"])
    assert abs(Du.norm_squared() * 0.5 ** L - float(g[f"c1_{L}_h1"])) < 1e-10 * float(g[f"c1_{L}_h1"])
    if L <= 8:
        assert rel_err(v.eval(squeeze=False), g[f"c1_{L}_v_dense"]) < 1e-10
        assert rel_err(u.eval(squeeze=False), g[f"c1_{L}_u_dense"]) < 1e-10


@pytest.mark.parametrize("L,steps,cap", [(3, 30, 8), (4, 20, 6)])
def test_c3_2d_bpx_solve(golden, tm, L, steps, cap):
    """Config C3 at small L: the 2D BPX pipeline of solver.py:212-233 with the numpy-only solver."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.bpx2D as B2
    import laplacesolvermps_b200.utils as U
    g = golden("solves")
    f = U.get_example_f_2D(L).reapprox(ranks_new=cap).flatten_mode_indices()
    B = B2.get_laplace_BPX_2D(L).reapprox(rel_error=1e-12)             # device rounding of the operator
    assert B.ranks == g[f"c3_{L}_B_ranks"].tolist()
    b = (B2.get_rhs_matrix_BPX_2D(L) @ f).squeeze().reapprox(ranks_new=cap, rel_error=1e-12)
    assert b.ranks == g[f"c3_{L}_b_ranks"].tolist()
    v, r2 = S.solve_with_grad_descent(B, b, n_steps_max=steps, max_rank=cap, rel_accuracy=0.0)
    assert v.ranks == g[f"c3_{L}_v_ranks"].tolist()
    assert rel_err(v.eval(squeeze=False), g[f"c3_{L}_v_dense"]) < 1e-9
    u = (B2.get_BPX_preconditioner_2D(L) @ v).reapprox(ranks_new=60, rel_error=1e-12)
    v2 = v.copy(); v2.tensors.append(np.ones([1, 1, 1, 1]))
    Dx = (B2.get_BPX_Qp_2D(L, "x") @ v2).flatten_mode_indices()
    Dy = (B2.get_BPX_Qp_2D(L, "y") @ v2).flatten_mode_indices()
    I = U._get_identy_as_tt(L, True)
    G = U._get_gram_matrix_tt(L); G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)
    Gy, Gx = U.kronecker_prod_2D(I, G), U.kronecker_prod_2D(G, I)
    energy = 0.25 ** L * ((Gy @ Dx).norm_squared() + (Gx @ Dy).norm_squared())
    assert abs(energy - float(g[f"c3_{L}_energy"])) < 1e-10 * float(g[f"c3_{L}_energy"])
    assert abs(S.get_L2_norm_2D(u) - float(g[f"c3_{L}_l2"])) < 1e-10 * float(g[f"c3_{L}_l2"])
    assert abs(v.inner(B @ v) - float(g[f"c3_{L}_vBv"])) < 1e-10 * float(g[f"c3_{L}_vBv"])


@pytest.mark.parametrize("L,cap,steps", [(20, 8, 6), (30, 6, 5)])
def test_c1_c2_config_sizes_against_oracle(tm, eng, L, cap, steps):
    """Configs C1 (1D, L = 20, rank cap 20) and C2 (deep levels, L = 30): the same truncated descent on the GPU and
    in the oracle, compared step by step (residual history, 1e-9) and on the final iterate (norm, probes)."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    f = U.get_polynomial_as_tt([1.0], L)
    B = S.get_laplace_bpx(L)
    b = (S.get_rhs_matrix_bpx(L) @ f).squeeze()
    for i in range(L):
        b.tensors[i] = b.tensors[i] / 2
    hist_ref = []
    x_ref, r2_ref = O.solve_with_grad_descent(O.clone(B.tensors), O.clone(b.tensors), n_steps_max=steps, max_rank=cap,
                                              rel_accuracy=0.0, history=hist_ref)
    x, r2, _, hist = eng.gd_solve(DeviceMPO(B.tensors, eng.device), BatchedTT.from_host([b.tensors], eng.device), n_steps=steps,
                                  max_rank=cap, rel_accuracy=0.0, want_history=True)
    h = hist.cpu().numpy()[:, 0, :]
    for n, r2n, gam in hist_ref:
        assert abs(h[n, 0] - r2n) < 1e-9 * r2n and abs(h[n, 1] - gam) < 1e-9 * abs(gam), n
    xh = x.to_host()[0]
    assert [c.shape for c in xh] == [c.shape for c in x_ref]
    assert abs(O.norm_squared(xh) - O.norm_squared(x_ref)) < 1e-10 * O.norm_squared(x_ref)
    rng = np.random.default_rng(9)
    for _ in range(3):
        p = random_tt(rng, [(2,)] * L, [2] * (L - 1), scale=0.7)
        assert abs(O.inner(p, xh) - O.inner(p, x_ref)) < 1e-9 * np.sqrt(O.norm_squared(p) * O.norm_squared(x_ref))
    assert abs(float(r2[0]) - r2_ref) < 1e-6 * r2_ref + 1e-12 * O.norm_squared(b.tensors)


def test_grad_descent_stopping_rule_and_ragged_batch(tm, eng):
    """rel_accuracy > 0: the reference stops at the first multiple of 10 steps (> 10) where ||r||^2/||x||^2 <= acc
    (solver.py:258-264).  Two different right-hand sides in one batch must each stop where the oracle stops."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    L, cap = 8, 10
    B = S.get_laplace_bpx(L)
    bs = []
    for f in (U.get_polynomial_as_tt([1.0], L), U.get_example_f_1D(L)):
        b = (S.get_rhs_matrix_bpx(L) @ f).squeeze()
        for i in range(L):
            b.tensors[i] = b.tensors[i] / 2
        bs.append(b.reapprox(rel_error=1e-14))                      # different ranks: a ragged batch
    refs = [O.solve_with_grad_descent(O.clone(B.tensors), O.clone(b.tensors), n_steps_max=200, max_rank=cap, rel_accuracy=1e-12,
                                      history=h) for b, h in zip(bs, ([], []))]
    hists = []
    for b in bs:
        h = []
        O.solve_with_grad_descent(O.clone(B.tensors), O.clone(b.tensors), n_steps_max=200, max_rank=cap, rel_accuracy=1e-12, history=h)
        hists.append(len(h))
    x, r2, steps, _ = eng.gd_solve(DeviceMPO(B.tensors, eng.device), BatchedTT.from_host([b.tensors for b in bs], eng.device),
                                   n_steps=200, max_rank=cap, rel_accuracy=1e-12)
    xs = x.to_host()
    assert steps.cpu().numpy().tolist() == hists                     # the step at which each instance broke out
    assert hists[0] < 200 and hists[1] < 200
    for i in range(2):
        xr, r2r = refs[i]
        assert [c.shape for c in xs[i]] == [c.shape for c in xr]
        assert rel_err(O.dense(xs[i], False), O.dense(xr, False)) < 1e-9
        assert abs(float(r2[i]) - r2r) < 1e-3 * r2r + 1e-16 * O.norm_squared(bs[i].tensors)


def test_pde_drivers_reach_the_manufactured_solution(tm):
    """solve_PDE_1D_with_preconditioner / solve_PDE_2D_with_preconditioner (solver.py:192-233) end to end on the GPU.
    `solve_with_amen` is served by the device iterative solver (ttpy is absent: parity unpinned; CG by default, the 2D call
    below asks for the reference's steepest descent), so the check is
    against the exact solution the experiments use (1D_solve_PDE.py, 2D_solve_PDE.py)."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    L = 8
    u, Du = S.solve_PDE_1D_with_preconditioner(U.get_example_f_1D(L), eps=1e-10, nswp=8, max_rank=20)
    u_ref = U.get_example_u_1D(L, basis="nodal")
    err = np.linalg.norm(u.eval().reshape(-1) - u_ref.eval().reshape(-1)) / np.linalg.norm(u_ref.eval().reshape(-1))
    assert err < 5e-4                                                # O(h^2) discretisation error at L = 8
    ref_h1 = -1 / (4 * np.pi ** 2) + 5 / 12 + 4 * np.pi ** 2 / 15
    assert abs(Du.norm_squared() * 0.5 ** L / ref_h1 - 1) < 2e-3
    L2 = 4
    f2 = U.get_example_f_2D(L2).reapprox(ranks_new=30)
    u2, Dx, Dy = S.solve_PDE_2D_with_preconditioner(f2, max_rank=30, eps=1e-10, nswp=6, method='gd')
    u2_ref = U.get_example_u_2D(L2, basis="nodal").flatten_mode_indices()
    a, b = u2.eval().reshape(-1), u2_ref.eval().reshape(-1)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 5e-2         # against the EXACT solution: discretisation error at L = 4
    # parity-level check: u solves the discrete system of solve_PDE_2D (solver.py:599-612), A u = g with the UNPRECONDITIONED
    # operator - the preconditioned driver must reproduce it to the solver's accuracy, independent of the mesh width
    A2 = S.build_laplace_matrix_2D(L2)
    g2 = (S.get_rhs_matrix_as_tt_2D(L2) @ f2.copy().flatten_mode_indices()).squeeze()
    res = (A2 @ u2).squeeze() - g2
    assert res.norm_squared() / g2.norm_squared() < 1e-16, res.norm_squared() / g2.norm_squared()
    assert Dx.shapes[-1][1] == 2 and Dy.shapes[-1][1] == 2          # trailing core carries the within-cell basis


def test_condition_number_by_tt_power_iteration(golden):
    """Config C4 where a reference value exists: cond(B) of the 2D BPX system at L = 2, 3 (dense SVD in the reference,
    5.4127426440 and 7.9349840729) and of the 1D system at L = 6, reproduced by TT power iteration on the GPU."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.bpx2D as B2
    g = golden("operators")
    for i, L in enumerate((2, 3)):
        B = S._rounded(B2.get_laplace_BPX_2D(L), rel_error=1e-14)
        cond = S.estimate_condition_number(B, max_rank=64, n_iter=600, tol=1e-14)
        assert abs(cond - g["cond_B2D_L2_4"][i]) < 1e-8 * cond
    # 1D: the low end of the spectrum is clustered, plain power iteration converges slowly - an estimate (1 %)
    cond = S.estimate_condition_number(S.get_laplace_bpx(6), max_rank=64, n_iter=800, tol=0.0)
    assert abs(cond - g["cond_B1D_L2_8"][4]) < 1e-2 * cond


def test_condition_number_by_tt_krylov_rayleigh_ritz(golden):
    """Config C4 with the Krylov estimator: the dense values of the reference (2D L = 2, 3, 4: 5.4127.., 7.9349.., 10.2732..) are
    reached to 2e-5 with 40 vectors where the power iteration needs hundreds of steps (the Ritz value for lambda_min converges like CG);
    with a binding rank cap the result is a lower bound."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.bpx2D as B2
    g = golden("operators")
    for i, L in enumerate((2, 3, 4)):
        B = S._rounded(B2.get_laplace_BPX_2D(L), rel_error=1e-14)
        cond, lmax, lmin, trace = S.estimate_condition_number_krylov(B, max_rank=64, n_vectors=40, rel_eps=1e-14, return_trace=True)
        assert abs(cond - g["cond_B2D_L2_4"][i]) < 2e-5 * cond and cond <= g["cond_B2D_L2_4"][i] * (1 + 1e-9), (L, cond, trace[-3:])
    B = S._rounded(B2.get_laplace_BPX_2D(4), rel_error=1e-14)
    low = S.estimate_condition_number_krylov(B, max_rank=4, n_vectors=30, rel_eps=1e-14)          # cap binds: 4^2 = 16 is the full bond
    assert 1.0 < low <= g["cond_B2D_L2_4"][2] * (1 + 1e-6), low      # a valid but loose lower bound (6.65 of 10.27): truncation to rank 4 binds


def test_condition_number_by_locally_optimal_block_iteration(golden):
    """Config C4 with the round-2 estimator (LOBPCG on both ends of the spectrum, every product / rounding / inner product on the
    GPU): the dense values of the reference at 2D L = 2..5 (5.4127426440, 7.9349840729, 10.2732726902, 12.3454280058 -
    experiments/2D_analyze_condition_number.py) to 2e-8, from below (Ritz values).  Measured: -7e-10 at L = 3, -3e-9 at L = 5; the
    remaining distance is the stopping tolerance on lambda_min, whose Ritz value converges with the square of its residual."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.bpx2D as B2
    dense = {2: 5.4127426440, 3: 7.9349840729, 4: 10.2732726902, 5: 12.3454280058}
    for L in (2, 3, 4, 5):
        B = S._rounded(B2.get_laplace_BPX_2D(L), rel_error=1e-14)
        cond, lmax, lmin, trace = S.estimate_condition_number_lobpcg(B, max_rank=32, n_iter=300, rel_eps=1e-14, tol=1e-11, return_trace=True)
        assert abs(cond - dense[L]) < 2e-8 * cond and cond <= dense[L] * (1 + 1e-9), (L, cond, trace[-2:])
    # binding cap: still a lower bound
    low = S.estimate_condition_number_lobpcg(S._rounded(B2.get_laplace_BPX_2D(4), rel_error=1e-14), max_rank=4, n_iter=60, rel_eps=1e-14)
    assert 1.0 < low <= dense[4] * (1 + 1e-6), low


def test_kat_norm_of_interpolant_1d(golden, tm):
    """experiments/1D_norm_of_interpolant.py: analytic norms of u = (1-3x+2x^2) sin(2 pi x)."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.utils as U
    g = golden("functions")
    ref_l2 = -3 / (16 * np.pi ** 2) + 3 / (16 * np.pi ** 4) + 1 / 15
    ref_h1 = -1 / (4 * np.pi ** 2) + 5 / 12 + 4 * np.pi ** 2 / 15
    for L, l2_ref, h1_ref in g["kat1d"]:
        L = int(L)
        u = U.get_example_u_1D(L, basis="nodal")
        l2 = S.get_L2_norm_1D(u)
        h1 = (S.get_derivative_matrix_as_tt(L) @ u).norm_squared() * 0.5 ** L
        assert abs(l2 - l2_ref) < 1e-11 * l2_ref and abs(h1 - h1_ref) < 1e-11 * h1_ref
        if L >= 20:
            assert abs(l2 / ref_l2 - 1) < 1e-10 and abs(h1 / ref_h1 - 1) < 1e-10


# ---- config C3 / C5 sizes: 2D, L = 12, B at rank 161 ----------------------------------------------

@pytest.fixture(scope="module")
def ops12():
    """Operators of the 2D L = 12 system; the raw rank-1152 operator is rounded ON THE DEVICE (cooperative QR leaf,
    multi-CTA Jacobi SVD) - the reference's goldens below were produced with numpy's rounding of the same operator."""
    from laplacesolvermps_b200.pipeline import build_operators_2D
    return build_operators_2D(12, 1e-12, on_device=True)


def test_device_operator_rounding_matches_host_construction(tm):
    """`get_laplace_BPX_2D(L).reapprox(eps)` on the GPU against the host construction route: identical
    rank profile [16, 161, .., 121, 16] and the same operator to the truncation tolerance."""
    import laplacesolvermps_b200.solver as S
    import laplacesolvermps_b200.bpx2D as B2
    for L in (4, 6):
        raw = B2.get_laplace_BPX_2D(L)
        assert max(raw.ranks) == 1152
        dev = raw.copy().reapprox(rel_error=1e-12)
        host = S._rounded(raw.copy(), rel_error=1e-12)
        assert dev.ranks == host.ranks
        if L == 4:
            a, b = dev.evalm(), raw.evalm()
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 5e-12
        else:
            d = dev - host
            assert d.norm_squared() < 1e-22 * host.norm_squared()


def test_L12_operator_and_single_rounding(golden, eng, ops12):
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO
    g = golden("heavy_2d_L12_ops")
    L = 12
    assert [c.shape[-1] for c in ops12["B"][:-1]] == g["B_ranks"].tolist()           # [16, 161 x 8, 121, 16]
    rng = np.random.default_rng(77)
    probes = [random_tt(rng, [(4,)] * L, [2] * (L - 1), scale=0.5) for _ in range(3)]
    A = DeviceMPO(ops12["B"], eng.device)
    P = BatchedTT.from_host(probes, eng.device)
    BP = eng.round_sum([(A, P)], caps=None, rel_eps=1e-13)           # exact up to 1e-13: B p_i
    pBp = np.zeros((3, 3))
    bp_host = BP.to_host()
    for i in range(3):
        Bi = BatchedTT.from_host([bp_host[i]] * 3, eng.device)
        pBp[:, i] = eng.inner(P, Bi).cpu().numpy()
    assert rel_err(pBp, g["pBp"]) < 1e-10
    for r in (2, 4, 8):                       # r = 8: panels of 5152 rows (4-wide shared-memory leaf, Gram + dlarft path)
        x = random_tt(np.random.default_rng(100 + r), [(4,)] * L, [r] * (L - 1), scale=0.5)
        Y = eng.round_sum([(A, BatchedTT.from_host([x], eng.device))], caps=r, rel_eps=1e-10)
        assert Y.ranks_host()[0, 1:-1].tolist() == g[f"round_r{r}_ranks"].tolist()
        n2 = float(eng.norm2(Y).cpu()[0])
        assert abs(n2 - float(g[f"round_r{r}_norm2"])) < 1e-10 * float(g[f"round_r{r}_norm2"])
        yh = Y.to_host()[0]
        Yb = BatchedTT.from_host([yh] * 3, eng.device)
        pr = eng.inner(P, Yb).cpu().numpy()
        assert rel_err(pr, g[f"round_r{r}_probes"]) < 1e-10


def test_factored_operator_application_equals_the_assembled_one_without_truncation(eng):
    """SURVEY 8(f).2 (pipeline.FactoredSolver2D): with caps that do not bind (L = 3: bonds <= 8) the steepest descent on
    B = S_x^T S_x + S_y^T S_y applied in factored form walks the same iterates as on the assembled operator."""
    from laplacesolvermps_b200.batched import BatchedTT
    from laplacesolvermps_b200.pipeline import Solver2DBatched, FactoredSolver2D, build_operators_2D
    L = 3
    ops = build_operators_2D(L, 1e-14, on_device=True)
    f = BatchedTT.from_host([c5_rhs_cores(i, L) for i in range(3)], eng.device)
    a = Solver2DBatched(L, max_rank=16, n_steps=30, eps=1e-14, host_ops=ops).solve(f)
    b = FactoredSolver2D(L, max_rank=16, inner_rank=64, n_steps=30, eps=1e-14, host_ops=ops).solve(f)
    va, vb = a["v"].to_host(), b["v"].to_host()
    for i in range(3):
        assert rel_err(O.dense(vb[i], False), O.dense(va[i], False)) < 1e-9
    assert torch.allclose(a["energy"], b["energy"], rtol=1e-9, atol=0) and torch.allclose(a["l2"], b["l2"], rtol=1e-9, atol=0)


@pytest.mark.parametrize("steps", [10, 100])
def test_c5_batched_solves(golden, eng, ops12, steps):
    """Config C5: batched 2D L=12 solves at cap 4 against the reference's own runs of the same instances."""
    from laplacesolvermps_b200.batched import BatchedTT
    from laplacesolvermps_b200.pipeline import Solver2DBatched
    g = golden(f"heavy_2d_L12_c5_steps{steps}")
    n_inst = int(g["n_inst"])
    L = 12
    solver = Solver2DBatched(L, max_rank=4, n_steps=steps, eps=1e-12, host_ops=ops12)
    f = BatchedTT.from_host([c5_rhs_cores(i, L) for i in range(n_inst)], eng.device)
    res = solver.solve(f)
    probes_rng = np.random.default_rng(77)
    probes = [random_tt(probes_rng, [(4,)] * L, [2] * (L - 1), scale=0.5) for _ in range(3)]
    v_host = res["v"].to_host()
    b_n2 = eng.norm2(res["b"]).cpu().numpy()
    v_n2 = eng.norm2(res["v"]).cpu().numpy()
    r2 = res["r2"].cpu().numpy()
    l2 = res["l2"].cpu().numpy()
    for i in range(n_inst):
        assert res["b"].ranks_host()[i, 1:-1].tolist() == g[f"inst{i}_b_ranks"].tolist()
        assert res["v"].ranks_host()[i, 1:-1].tolist() == g[f"inst{i}_v_ranks"].tolist()
        assert abs(b_n2[i] - float(g[f"inst{i}_b_norm2"])) < 1e-10 * float(g[f"inst{i}_b_norm2"])
        assert abs(v_n2[i] - float(g[f"inst{i}_v_norm2"])) < 1e-9 * float(g[f"inst{i}_v_norm2"])
        pr = np.array([O.inner(p, v_host[i]) for p in probes])
        assert rel_err(pr, g[f"inst{i}_v_probes"]) < 1e-8
        assert abs(l2[i] - float(g[f"inst{i}_u_l2"])) < 1e-9 * float(g[f"inst{i}_u_l2"])
        assert abs(r2[i] - float(g[f"inst{i}_r2"])) < 1e-4 * float(g[f"inst{i}_r2"]) + 1e-3 * b_n2[i] * 1e-12


def test_cluster_leaf_against_oracle(monkeypatch):
    """The thread-block-cluster QR leaf (csrc/qr3.cuh, rows of a sub-panel split over the CTAs of a cluster, reductions through
    distributed shared memory) forced on small panels: orthogonalisation and rounding of products against the oracle.
    At full size the same kernel serves every panel taller than ~3400 rows (rank caps >= 5 at 2D L = 12)."""
    from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
    monkeypatch.setenv("QTT_FORCE_CLUSTER_LEAF", "1")
    eng = Engine("cuda:0")
    try:
        rng = np.random.default_rng(21)
        L, N = 5, 5
        op = random_tt(rng, [(4, 4)] * L, [9] * (L - 1), scale=0.3)
        xs = [random_tt(rng, [(4,)] * L, [5] * (L - 1)) for _ in range(N)]
        A, X = DeviceMPO(op, eng.device), BatchedTT.from_host(xs, eng.device)
        before = eng.counters()[1]
        Y = eng.round_sum([(A, X)], caps=6, rel_eps=1e-14)
        Yo, n2 = eng.orthogonalize([(A, X)])
        assert eng.counters()[1] > before
        yh, oh = Y.to_host(), Yo.to_host()
        for i in range(N):
            prod = O.matmul(op, xs[i])
            ref = O.reapprox([c.copy() for c in prod], ranks_new=6, rel_error=1e-14)
            assert [c.shape for c in yh[i]] == [c.shape for c in ref]
            assert rel_err(O.dense(yh[i], False), O.dense(ref, False)) < 1e-12
            assert rel_err(O.dense(oh[i], False), O.dense(prod, False)) < 1e-12
            assert abs(float(n2[i]) - O.norm_squared(prod)) < 1e-12 * O.norm_squared(prod)
    finally:
        eng.close()
        monkeypatch.delenv("QTT_FORCE_CLUSTER_LEAF")
        Engine("cuda:0").close()        # re-reads the environment: the switch is off again for the tests that follow
