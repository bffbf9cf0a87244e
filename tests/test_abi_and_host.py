"""CPU-only checks: the C-ABI library loads and exports every symbol include/qtt_b200.h declares, the
host-side packing / sharding logic, and the absence of any CPU fallback for the hot path."""
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from laplacesolvermps_b200 import _lib
from laplacesolvermps_b200.batched import BatchedTT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qtt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qtt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    _lib.build_library()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qtt_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.qtt_version() >= 100


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import laplacesolvermps_b200.tensormethods as tm
    a = tm.TensorTrain([np.ones((1, 2, 1)), np.ones((1, 2, 1))])
    for call in (lambda: a @ a, lambda: a.copy().reapprox(), lambda: a.norm_squared(), lambda: a.copy().left_orthogonalize(),
                 lambda: a.inner(a)):
        with pytest.raises(_lib.QttError):
            call()


def test_product_does_not_import_oracle():
    import laplacesolvermps_b200  # noqa: F401
    pkg = os.path.join(ROOT, "laplacesolvermps_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_batched_pack_unpack_roundtrip_ragged():
    rng = np.random.default_rng(0)
    insts = []
    for i in range(7):
        rk = [1] + [int(rng.integers(1, 5)) for _ in range(3)] + [1]
        insts.append([rng.standard_normal((rk[k], 2, 2, rk[k + 1])) for k in range(4)])
    B = BatchedTT.from_host(insts, "cpu")
    assert B.n == [4] * 4 and B.N == 7
    back = B.to_host()
    for a, b in zip(insts, back):
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)
    st = B.c_struct()
    assert st.L == 4 and st.N == 7 and st.slot[0] == B.slot[0]


def test_tensortrain_container_semantics():
    import laplacesolvermps_b200.tensormethods as tm
    cores = [np.arange(4.0).reshape(1, 2, 2), np.arange(8.0).reshape(2, 2, 2), np.ones((2, 2, 1))]
    t = tm.TensorTrain(cores)
    assert t.tensors is cores                          # list is aliased (tensormethods.py:21-25)
    c = tm.TensorTrain(t)
    assert c.tensors[0] is not cores[0] and np.array_equal(c.tensors[0], cores[0])
    assert t.ranks == [2, 2] and t.mode_sizes == [(2,), (2,), (2,)] and len(t) == 3
    s = t + t
    assert s.ranks == [4, 4]
    assert np.allclose(s.eval(), 2 * t.eval())
    assert np.allclose((2.0 * t).eval(), 2 * t.eval()) and np.allclose((-t).eval(), -t.eval())
    assert np.allclose((t - t).eval(), 0)
    k = t.copy().expand_dims(1) * t.copy().expand_dims(0)
    assert k.shapes[1] == (4, 2, 2, 4)
    assert np.allclose(k.eval(squeeze=False), np.einsum("abc,def->adbecf", t.eval(), t.eval()))
    z = tm.zeros([(2,), (3,)])
    assert z.shapes == [(1, 2, 1), (1, 3, 1)]
    with pytest.raises(AssertionError):
        t + tm.TensorTrain(cores[:2])
    m = tm.TensorTrain([np.arange(8.0).reshape(1, 2, 4, 1)])
    assert m.copy().transpose().shapes == [(1, 4, 2, 1)]
    assert np.array_equal(m.evalm(), np.arange(8.0).reshape(2, 4))
    sq = tm.TensorTrain([np.ones((1, 1, 2)), np.arange(8.0).reshape(2, 2, 2), np.ones((2, 1, 1))]).squeeze()
    assert sq.shapes == [(1, 2, 1)]
    ttpy = t.to_ttpylist()
    back = tm.TensorTrain.from_ttpylist(ttpy)
    assert all(np.array_equal(a, b) for a, b in zip(back.tensors, t.tensors))
    np.random.seed(0)
    dense = np.random.normal(size=[2, 2, 2, 2, 2])
    assert np.allclose(tm.TensorTrain.ttsvd(dense, ranks=[2, 4, 4, 2]).eval(), dense)


def test_alias_package_resolves_to_this_repo():
    for k in [k for k in sys.modules if k == "laplace_mps" or k.startswith("laplace_mps.")]:
        del sys.modules[k]
    import laplace_mps.tensormethods as tm
    import laplace_mps.solver as solver
    import laplacesolvermps_b200.tensormethods as mine
    assert tm is mine and solver.__name__ == "laplacesolvermps_b200.solver"
    from laplace_mps.bpx2D import _get_Z0_hat, _get_Z1_hat  # noqa: F401  (experiments/2D_validate_bpx_preconditioner.py)


def test_shard_ranges_cover_everything():
    from laplacesolvermps_b200.parallel import shard_range
    for n, w in [(4096, 8), (10, 4), (3, 8), (148, 1)]:
        seen = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            seen.extend(range(lo, hi))
        assert seen == list(range(n))
