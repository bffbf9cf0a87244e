"""Developer timing of config C1 (1D, L = 20, BPX, rank cap 20) batched on the GPU: seconds per descent step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laplacesolvermps_b200.solver as S, laplacesolvermps_b200.utils as U
from laplacesolvermps_b200 import _assembly as A
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
from oracle.tt_oracle import squeeze

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    eng = Engine("cuda:0")
    B = S.get_laplace_bpx(L)
    f = U.get_polynomial_as_tt([1.0], L)
    b = squeeze(A.compose(S.get_rhs_matrix_bpx(L).tensors, f.tensors))
    op = DeviceMPO(B.tensors, eng.device)
    bb = BatchedTT.from_host([b] * N, eng.device)
    prof_mode = os.environ.get("DEV_PROFILE", "0") == "1"       # the timeline mode turns the graph cache off
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if prof_mode: eng.set_profiling(2); eng.class_profile()
        x, r2, st, _ = eng.gd_solve(op, bb, n_steps=steps, max_rank=cap, rel_accuracy=0.0)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        prof = eng.class_profile() if prof_mode else {}
        if prof_mode: eng.set_profiling(0)
        print(f"N={N} L={L} cap={cap}: {dt/steps*1e3:.1f} ms per descent step for the batch, {dt/steps/N*1e3:.3f} ms per instance-step; "
              f"classes {({k: round(v, 1) for k, v in prof.items()})}", flush=True)
    print("graphs (captured, replayed)", eng.graph_counters(), "launches", eng.counters()[1])
    print("B ranks", B.ranks, "x ranks", x.ranks_host()[0].tolist(), "r2", float(r2[0]), flush=True)

if __name__ == "__main__":
    main()
