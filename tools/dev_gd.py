import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import tt_oracle as O
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
g = np.load("tests/golden/solves.npz")
def cores_from(npz, prefix):
    out, k = [], 0
    while f"{prefix}_{k}" in npz: out.append(np.array(npz[f"{prefix}_{k}"])); k += 1
    return out
A, b = cores_from(g, "spd_A"), cores_from(g, "spd_b")
eng = Engine("cuda:0")
hist_ref = []
x_ref, r2_ref = O.solve_with_grad_descent(A, b, n_steps_max=60, max_rank=4, history=hist_ref)
op = DeviceMPO(A, eng.device); B = BatchedTT.from_host([b], eng.device)
x, r2, steps, hist = eng.gd_solve(op, B, n_steps=60, max_rank=4, rel_accuracy=1e-14, want_history=True)
h = hist.cpu().numpy()[:, 0, :]
print("steps", steps.cpu().numpy(), "ref steps", len(hist_ref))
for n in range(0, min(len(hist_ref), 60), 3):
    print(n, hist_ref[n][1], h[n, 0], hist_ref[n][2], h[n, 1])
print(r2_ref, float(r2[0]))
