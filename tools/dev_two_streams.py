"""Developer experiment: the same 148-instance round(B @ x) as tools/dev_time.py, once as one batch and once as two half
batches driven by two host threads, each with its own handle and CUDA stream (kernels of the two halves interleave)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.make_golden import random_tt
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    L, r = 12, 4
    Rk = [16] + [161] * 8 + [121, 16]
    op = random_tt(rng, [(4, 4)] * L, Rk, scale=0.05)
    x1 = random_tt(rng, [(4,)] * L, [r] * (L - 1), scale=0.5)
    A = DeviceMPO(op, dev)

    def run(parts):
        engines = [Engine(dev) for _ in parts]
        streams = [torch.cuda.Stream(device=dev) for _ in parts]
        Xs = [BatchedTT.from_host([x1] * n, dev) for n in parts]
        torch.cuda.synchronize()
        def work(i, k):
            with torch.cuda.stream(streams[i]):
                for _ in range(k):
                    engines[i].round_sum([(A, Xs[i])], caps=r, rel_eps=1e-16)
        for k in (1, reps):                      # warm-up, then timed
            torch.cuda.synchronize(); t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(i, k)) for i in range(len(parts))]
            for t in th: t.start()
            for t in th: t.join()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        return dt / reps * 1e3
    print(f"one batch of {N}: {run([N]):.1f} ms per sweep of all instances")
    print(f"two halves on two streams / threads: {run([N // 2, N - N // 2]):.1f} ms")
    print(f"four quarters: {run([N // 4] * 3 + [N - 3 * (N // 4)]):.1f} ms")

if __name__ == "__main__":
    main()
