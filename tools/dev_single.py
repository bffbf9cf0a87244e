"""Developer timing of the SINGLE-INSTANCE TensorTrain API (numpy cores in, numpy cores out, host<->device copies per call) at the
2D L=12 shapes: one `(B @ x).reapprox(ranks_new=4)` - the call the reference's own drivers make - against the batched engine with N = 1."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.make_golden import random_tt
import laplacesolvermps_b200.tensormethods as tm
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine


def main():
    rng = np.random.default_rng(0)
    L, r = 12, 4
    Rk = [16] + [161] * 8 + [121, 16]
    op = random_tt(rng, [(4, 4)] * L, Rk, scale=0.05)
    x = random_tt(rng, [(4,)] * L, [r] * (L - 1), scale=0.5)
    A, X = tm.TensorTrain(op), tm.TensorTrain(x)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Y = (A @ X).reapprox(ranks_new=r)                   # the reference's idiom: materialised product, then rounding
        torch.cuda.synchronize(); t1 = time.perf_counter()
        n2 = Y.norm_squared()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"TensorTrain API iter {it}: (A @ x).reapprox(4) {1e3 * (t1 - t0):.1f} ms, norm_squared {1e3 * (t2 - t1):.1f} ms, ranks {Y.ranks}", flush=True)
    eng = Engine("cuda:0")
    Ad, Xd = DeviceMPO(op, eng.device), BatchedTT.from_host([x], eng.device)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Yd = eng.round_sum([(Ad, Xd)], caps=r, rel_eps=1e-16)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"batched engine N=1 iter {it}: fused round(A @ x) {1e3 * (t1 - t0):.1f} ms (device-resident operands)", flush=True)


if __name__ == "__main__":
    main()
