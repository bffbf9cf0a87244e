"""Developer timing of one round(B @ x) at the 2D L=12 shapes with a random operator of the BPX ranks."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.make_golden import random_tt
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = "cuda:0"
    eng = Engine(dev)
    rng = np.random.default_rng(0)
    L = 12
    Rk = [16] + [161] * 8 + [121, 16]
    op = random_tt(rng, [(4, 4)] * L, Rk, scale=0.05)
    A = DeviceMPO(op, dev)
    x1 = random_tt(rng, [(4,)] * L, [r] * (L - 1), scale=0.5)
    X = BatchedTT.from_host([x1] * N, dev)
    torch.cuda.synchronize()
    for it in range(reps + 1):
        eng.reset_counters()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        Y = eng.round_sum([(A, X)], caps=r, rel_eps=1e-16)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        fl, nl = eng.counters()
        print(f"iter {it}: {ms:.1f} ms, gemm flops {fl:.3e} ({fl/ms*1e-9:.2f} TF/s incl. everything), launches {nl}, per instance {ms/N:.2f} ms", flush=True)
    eng.set_profiling(2); eng.class_profile()
    Y = eng.round_sum([(A, X)], caps=r, rel_eps=1e-16); torch.cuda.synchronize()
    print("class profile (ms):", {k: round(v, 2) for k, v in eng.class_profile().items()})
    eng.set_profiling(0)
    print("ranks", Y.ranks_host()[0])
    n2 = eng.norm2(Y).cpu().numpy()
    print("norm2 spread", n2.min(), n2.max())

if __name__ == "__main__":
    main()
