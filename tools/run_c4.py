"""Config C4 (BASELINE.json): condition number estimate of the 2D BPX-preconditioned stiffness operator at L levels by TT power
iteration on the GPU, with its convergence trace.  The reference computes this only densely for L <= 5 (2D_analyze_condition_number.py);
there is no reference value at L = 15 - the output is an estimate limited by the rank cap, recorded with its trace.

    python tools/run_c4.py [L=15] [max_rank=8] [n_iter=200] [out.json]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    out = sys.argv[4] if len(sys.argv) > 4 else None
    import laplacesolvermps_b200.solver as S
    t0 = time.perf_counter()
    B = S._cached_laplace_BPX_2D(L, 1e-12)              # rank-1152 operator rounded on the device
    torch.cuda.synchronize()
    t_op = time.perf_counter() - t0
    t0 = time.perf_counter()
    cond, lmax, lmin, trace = S.estimate_condition_number(B, max_rank=cap, n_iter=n_iter, rel_eps=1e-12, tol=1e-10, return_trace=True)
    torch.cuda.synchronize()
    t_it = time.perf_counter() - t0
    t0 = time.perf_counter()
    kc, kmax, kmin, ktrace = S.estimate_condition_number_krylov(B, max_rank=cap, n_vectors=40, rel_eps=1e-12, return_trace=True)
    torch.cuda.synchronize()
    t_kr = time.perf_counter() - t0
    rec = {"krylov": {"n_vectors": len(ktrace), "product_rank": 4 * cap, "seconds": t_kr, "lambda_max": kmax, "lambda_min": kmin, "cond": kc,
                      "trace": [list(t) for t in ktrace]},
           "config": "C4", "L": L, "max_rank": cap, "operator_ranks": B.ranks, "operator_rounding_s": t_op, "iteration_s": t_it,
           "iterations": len(trace), "lambda_max": lmax, "lambda_min": lmin, "cond": cond,
           "dense_reference_values_L2_5": [5.4127426440, 7.9349840729, 10.2732726902, 12.3454280058],
           "trace_lambda_max": [t[2] for t in trace if not t[0]][::5], "trace_shifted": [t[2] for t in trace if t[0]][::5]}
    print(json.dumps({k: v for k, v in rec.items() if not k.startswith("trace") and k != "krylov"}))
    print("krylov:", json.dumps({k: v for k, v in rec["krylov"].items() if k != "trace"}), "last trace rows", ktrace[-3:])
    if out:
        json.dump(rec, open(out, "w"))


if __name__ == "__main__":
    main()
