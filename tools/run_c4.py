"""Config C4 (BASELINE.json): condition number of the 2D BPX-preconditioned stiffness operator at L levels on the GPU.
The reference computes it only densely for L <= 5 (experiments/2D_analyze_condition_number.py: 5.4127426440, 7.9349840729,
10.2732726902, 12.3454280058 at L = 2..5); there is no reference value at L = 15.  Three device estimators are run and recorded
with their convergence traces: the locally optimal block iteration (round 2: `estimate_condition_number_lobpcg`) at several rank
caps, and - for comparison - the TT power iteration and the truncated Krylov Rayleigh-Ritz of round 1.  All three return Ritz
values (lambda_max from below, lambda_min from above): LOWER bounds of the condition number that rise with the cap.

    python tools/run_c4.py [L=15] [caps=8,16,24] [n_iter=80] [out.json] [with_round1_estimators=0]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

DENSE = {2: 5.4127426440, 3: 7.9349840729, 4: 10.2732726902, 5: 12.3454280058}


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    caps = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "8,16,24").split(",")]
    n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    out = sys.argv[4] if len(sys.argv) > 4 else None
    old = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    import laplacesolvermps_b200.solver as S
    t0 = time.perf_counter()
    B = S._cached_laplace_BPX_2D(L, 1e-12)              # rank-1152 operator rounded on the device
    torch.cuda.synchronize()
    rec = {"config": "C4", "L": L, "operator_ranks": B.ranks, "operator_rounding_s": time.perf_counter() - t0,
           "dense_reference_values_L2_5": [DENSE[k] for k in (2, 3, 4, 5)], "lobpcg": []}
    for cap in caps:
        t0 = time.perf_counter()
        cond, lmax, lmin, trace = S.estimate_condition_number_lobpcg(B, max_rank=cap, n_iter=n_iter, rel_eps=1e-13, tol=1e-9, return_trace=True)
        torch.cuda.synchronize()
        row = {"max_rank": cap, "iterations": len(trace), "seconds": time.perf_counter() - t0, "lambda_max": lmax, "lambda_min": lmin,
               "cond": cond, "trace": [list(t) for t in trace]}
        if L in DENSE:
            row["rel_err_vs_dense"] = cond / DENSE[L] - 1
        rec["lobpcg"].append(row)
        print(json.dumps({k: v for k, v in row.items() if k != "trace"}), flush=True)
        print("   last iterations (it, lambda_max, lambda_min, |r_max|/lambda_max, |r_min|/lambda_min):", [tuple(float(f"{v:.9g}") for v in t) for t in trace[-3:]], flush=True)
    if old:
        cap = caps[0]
        t0 = time.perf_counter()
        cond, lmax, lmin, trace = S.estimate_condition_number(B, max_rank=cap, n_iter=150, rel_eps=1e-12, tol=1e-10, return_trace=True)
        rec["power"] = {"max_rank": cap, "seconds": time.perf_counter() - t0, "iterations": len(trace), "lambda_max": lmax, "lambda_min": lmin, "cond": cond}
        t0 = time.perf_counter()
        kc, kmax, kmin, ktrace = S.estimate_condition_number_krylov(B, max_rank=cap, n_vectors=40, rel_eps=1e-12, return_trace=True)
        rec["krylov"] = {"max_rank": cap, "seconds": time.perf_counter() - t0, "n_vectors": len(ktrace), "lambda_max": kmax, "lambda_min": kmin, "cond": kc}
        print("power:", json.dumps(rec["power"]), "\nkrylov:", json.dumps(rec["krylov"]), flush=True)
    if out:
        json.dump(rec, open(out, "w"))


if __name__ == "__main__":
    main()
