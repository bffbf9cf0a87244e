"""Alias package: `import laplace_mps.solver`, `laplace_mps.tensormethods`, ... resolve to the modules of
`laplacesolvermps_b200`, so the reference's experiment scripts (src/experiments/*.py, which import
`laplace_mps.*`) run unchanged with this repository on PYTHONPATH."""
import importlib
import sys

for _name in ("tensormethods", "utils", "bpx2D", "solver"):
    _mod = importlib.import_module("laplacesolvermps_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
