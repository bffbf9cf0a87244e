"""Load the UNMODIFIED reference package from /root/reference for fixture generation.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by the CPU tests that pin
oracle/tt_oracle.py against the live reference when /root/reference is mounted (it is not
on the GPU box).  Nothing in the product package imports this.

The reference cannot be imported as shipped on numpy 2.x without three in-memory shims
(SURVEY.md section 8c); none of them edits the read-only tree:
  1. utils.py:5 imports matplotlib.pyplot at module level  -> stub module
  2. tensormethods.py:171 uses the removed alias np.float   -> np.float = float
  3. bpx2D.py:104 ships rank = 2**(4+21) (128 PiB array)    -> re-exec with 2**(4+1) = 32
"""
import os
import sys
import types
import inspect
import warnings

_LIVE_SRC = "/root/reference/src"
_COPY_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_src.zip")   # made by oracle/build_ref.py (git-ignored)
# the mounted tree in the build container; on the GPU box the byte-for-byte archive that travelled with the working tree
# (imported through zipimport, never unpacked)
REFERENCE_SRC = _LIVE_SRC if os.path.isdir(os.path.join(_LIVE_SRC, "laplace_mps")) else _COPY_SRC


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "laplace_mps")) or os.path.isfile(REFERENCE_SRC)


def experiment_source(name):
    """Source text of src/experiments/<name> (or src/scratchpad/<name>) of the reference."""
    import zipfile
    for sub in ("experiments", "scratchpad"):
        if os.path.isdir(REFERENCE_SRC):
            path = os.path.join(REFERENCE_SRC, sub, name)
            if os.path.exists(path):
                with open(path) as fh:
                    return fh.read()
        elif os.path.isfile(REFERENCE_SRC):
            with zipfile.ZipFile(REFERENCE_SRC) as z:
                if f"{sub}/{name}" in z.namelist():
                    return z.read(f"{sub}/{name}").decode()
    raise FileNotFoundError(name)


def reference_kind():
    """'reference' when the unmodified reference sources are importable (mounted or copied), else None."""
    return "reference" if reference_available() else None


def load_reference():
    """Returns (tensormethods, solver, bpx2D, utils) modules of the reference."""
    import numpy as np
    if not reference_available():
        raise RuntimeError("reference sources not found (neither /root/reference nor oracle/_ref: run python -m oracle.build_ref)")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m, p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            m.pyplot = p
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = p
    if not hasattr(np, "float"):
        np.float = float
    # the repo ships its own `laplace_mps` alias package; make sure the reference one wins here
    for name in [k for k in sys.modules if k == "laplace_mps" or k.startswith("laplace_mps.")]:
        if not getattr(sys.modules[name], "__file__", "").startswith(REFERENCE_SRC):
            del sys.modules[name]
    if REFERENCE_SRC in sys.path:
        sys.path.remove(REFERENCE_SRC)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import laplace_mps.tensormethods as tm
            import laplace_mps.utils as utils
            import laplace_mps.bpx2D as bpx2D
            src = inspect.getsource(bpx2D.get_BPX_preconditioner_2D).replace("2**(4+21)", "2**(4+1)")
            exec(compile(src, "bpx2D_rank_fix", "exec"), bpx2D.__dict__)
            import laplace_mps.solver as solver
            solver.get_BPX_preconditioner_2D = bpx2D.get_BPX_preconditioner_2D
    finally:
        sys.path.remove(REFERENCE_SRC)
    assert tm.__file__.startswith(REFERENCE_SRC)
    return tm, solver, bpx2D, utils
