"""Acceptance harness: the reference's own experiment scripts (src/experiments/*.py), UNMODIFIED, executed on top of the
`laplace_mps` alias package of this repository - every `@`, `reapprox`, `norm_squared` in them runs on the GPU.

The script sources come from /root/reference (build container) or from oracle/_ref/reference_src.zip (GPU box; made by
`python -m oracle.build_ref`, git-ignored).  matplotlib is replaced by an inert stub (the scripts only plot with it); the
printed / computed errors are compared with SURVEY.md section 4's table (values measured with the reference itself)."""
import contextlib
import io
import os
import re
import sys
import types
from unittest import mock

import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.gpu


def _run_script(name, tmp_path):
    try:
        src = ref_loader.experiment_source(name)
    except FileNotFoundError:
        pytest.skip("reference experiment scripts not available (run python -m oracle.build_ref in the build container)")
    # inert matplotlib (any attribute / call returns another mock; subplots() must unpack)
    plt = mock.MagicMock(name="matplotlib.pyplot")
    def subplots(nrows=1, ncols=1, *a, **k):          # axes must index and iterate like the numpy arrays of Axes matplotlib returns
        if nrows == 1 and ncols == 1:
            axes = mock.MagicMock()
        elif nrows == 1 or ncols == 1:
            axes = [mock.MagicMock() for _ in range(max(nrows, ncols))]
        else:
            axes = [[mock.MagicMock() for _ in range(ncols)] for _ in range(nrows)]
        return mock.MagicMock(), axes
    plt.subplots.side_effect = subplots
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    for k in [k for k in sys.modules if k == "laplace_mps" or k.startswith("laplace_mps.")]:
        if "reference" in (getattr(sys.modules[k], "__file__", "") or ""):
            del sys.modules[k]                       # a reference import left by the oracle tests must not serve the scripts
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if not hasattr(np, "float"):
        np.float = float
    cwd = os.getcwd()
    os.makedirs(tmp_path / "outputs", exist_ok=True)
    os.chdir(tmp_path)
    out = io.StringIO()
    glob = {"__name__": "__main__", "__file__": name}
    try:
        with contextlib.redirect_stdout(out):
            exec(compile(src, name, "exec"), glob)
    finally:
        os.chdir(cwd)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    import laplace_mps.tensormethods as tm
    import laplacesolvermps_b200.tensormethods as mine
    assert tm is mine, "the scripts must have run on this repository's package"
    return glob, out.getvalue()


def test_1D_norm_of_interpolant(tmp_path):
    g, _ = _run_script("1D_norm_of_interpolant.py", tmp_path)
    Ls = list(g["L_values"])
    rel_l2 = g["error_L2"] / g["refnorm_L2"]
    rel_h1 = g["error_H1_smart"] / g["refnorm_H1"]
    # SURVEY section 4 (reference values): L2 / H1-orth = -9.7e-6 / -9.3e-6 (L=10), -6.1e-7 / -5.8e-7 (12), -9.2e-12 / -8.9e-12 (20)
    for L, l2, h1 in ((10, -9.7e-6, -9.3e-6), (12, -6.1e-7, -5.8e-7), (20, -9.2e-12, -8.9e-12)):
        i = Ls.index(L)
        assert abs(rel_l2[i] / l2 - 1) < 0.05 and abs(rel_h1[i] / h1 - 1) < 0.05, (L, rel_l2[i], rel_h1[i])
    i = Ls.index(30)
    assert abs(rel_l2[i]) < 1e-13 and abs(rel_h1[i]) < 1e-12                   # reference: 1.7e-15 / 4.5e-14
    # every level against the output of the SAME script run on the unmodified reference (oracle/make_golden.py --only experiments).
    # Past L = 30 the orthogonalised H1 error of the reference itself grows again (1.2e-9 at L = 40, 2.9e-7 at L = 44: round-off of
    # the encoded function amplified by the 2^L scaling of the derivative); the GPU path reproduces that curve, not a flat line.
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exp_1D_norm_of_interpolant.npz"))
    assert list(ref["L_values"]) == Ls
    for key, norm in (("error_L2", g["refnorm_L2"]), ("error_H1_smart", g["refnorm_H1"])):
        mine, theirs = g[key] / norm, ref[key] / norm
        # (that tail is amplified round-off: two implementations agree on its size - measured here within 15 % - not on its digits)
        assert np.all(np.abs(mine - theirs) < 0.5 * np.abs(theirs) + 1e-12), (key, mine - theirs)
        lo = Ls.index(14)
        assert np.all(np.abs(mine[:lo] - theirs[:lo]) < 1e-5 * np.abs(theirs[:lo])), (key, mine[:lo] - theirs[:lo])   # discretisation-error regime


def test_2D_norm_of_interpolant(tmp_path):
    g, _ = _run_script("2D_norm_of_interpolant.py", tmp_path)
    Ls = list(g["L_values"])
    rel_l2 = g["error_L2"] / g["refnorm_L2"]
    rel_h1 = g["error_H1_ort"] / g["refnorm_H1"]
    for L, l2, h1 in ((6, -2.7e-3, -2.6e-3), (12, -6.5e-7, -6.3e-7), (20, -1.0e-11, -1.1e-11)):
        i = Ls.index(L)
        assert abs(rel_l2[i] / l2 - 1) < 0.06 and abs(rel_h1[i] / h1 - 1) < 0.12, (L, rel_l2[i], rel_h1[i])


def test_2D_validate_bpx_preconditioner(tmp_path):
    _, out = _run_script("2D_validate_bpx_preconditioner.py", tmp_path)
    res = [float(x) for x in re.findall(r"Norm. residual\s*:\s*([0-9.eE+-]+)", out)]
    ratios = [float(x) for x in re.findall(r"log2 ratio\s*:\s*([0-9.eE+-]+)", out)]
    assert len(res) == 5 and len(ratios) == 5, out
    assert max(res) < 1e-24, out                       # reference (with its rank typo fixed): 1.2e-28, ..., 7.1e-33
    assert max(abs(r) for r in ratios) < 1e-3, out


def test_1D_validate_bpx_preconditioner(tmp_path):
    _, out = _run_script("1D_validate_bpx_preconditioner.py", tmp_path)
    errs = [float(x) for x in re.findall(r"rel error [a-z ]+:\s*([0-9.eE+-]+)", out)]
    assert len(errs) == 2 and max(errs) < 1e-13, out   # reference: 1.07e-15
