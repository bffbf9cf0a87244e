"""Developer timing of the solver's small operations (rank-4 vectors, 2D L=12) at batch size N: host-inclusive wall time per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.make_golden import random_tt
from laplacesolvermps_b200.batched import BatchedTT, Engine

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = "cuda:0"; eng = Engine(dev); rng = np.random.default_rng(0); L = 12
    xs = [random_tt(rng, [(4,)] * L, [r] * (L - 1), scale=0.5) for _ in range(N)]
    ys = [random_tt(rng, [(4,)] * L, [r] * (L - 1), scale=0.5) for _ in range(N)]
    X, Y = BatchedTT.from_host(xs, dev), BatchedTT.from_host(ys, dev)
    g = torch.full((N,), 0.3, device=dev, dtype=torch.float64)
    def timeit(name, fn, reps=20):
        fn(); torch.cuda.synchronize()
        eng.reset_counters()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name}: {dt*1e3:.3f} ms per call, launches per call {eng.counters()[1] / reps:.0f}", flush=True)
    timeit("round(x + g y) cap r", lambda: eng.round_sum([(None, X), (None, Y, g)], caps=r, rel_eps=1e-16))
    timeit("norm2(x)", lambda: eng.norm2(X))
    timeit("inner(x, y)", lambda: eng.inner(X, Y))

    print("graphs (captured, replayed)", eng.graph_counters())

if __name__ == "__main__":
    main()
