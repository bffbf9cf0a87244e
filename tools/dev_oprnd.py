"""Developer check: device rounding of the raw 2D BPX operator (rank 1152) against the host construction path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laplacesolvermps_b200.bpx2D as B2, laplacesolvermps_b200.solver as S
import laplacesolvermps_b200.tensormethods as tm

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eps = 1e-12
raw = B2.get_laplace_BPX_2D(L)
print("raw ranks", raw.ranks, flush=True)
t0 = time.time(); host = S._rounded(raw.copy(), rel_error=eps); th = time.time() - t0
print("host ranks", host.ranks, f"{th:.2f} s", flush=True)
from laplacesolvermps_b200.batched import default_engine
eng = default_engine()
for it in range(2):
    eng.set_profiling(2); eng.class_profile()
    torch.cuda.synchronize(); t0 = time.time()
    dev = raw.copy().reapprox(rel_error=eps)
    torch.cuda.synchronize(); td = time.time() - t0
    print("device ranks", dev.ranks, f"{td:.2f} s", {k: round(v) for k, v in eng.class_profile().items()}, flush=True)
    eng.set_profiling(0)
if L <= 5:
    a, b = dev.evalm(), host.evalm()
    print("rel diff dense", np.linalg.norm(a - b) / np.linalg.norm(b), "vs raw", np.linalg.norm(a - raw.evalm()) / np.linalg.norm(b))
else:
    d = dev - host
    print("||dev - host||^2 / ||host||^2", d.norm_squared() / host.norm_squared())
