"""Accuracy / cost study of the factored application of the 2D BPX operator (pipeline.FactoredSolver2D)
against the assembled rank-161 operator (the reference's arithmetic, pipeline.Solver2DBatched) on the C5 instances.

    python tools/study_factored.py [N=148] [L=12] [out.json]

For every inner rank: relative deviation of energy / L2 norm / solution from the assembled path, the residual of both solutions
measured with the ASSEMBLED operator at a generous cap (so neither path grades itself), and the time per batched solve."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import c5_rhs_cores, operators_cached
from laplacesolvermps_b200.batched import BatchedTT
from laplacesolvermps_b200.pipeline import Solver2DBatched, FactoredSolver2D


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    out = sys.argv[3] if len(sys.argv) > 3 else None
    cap, steps = 4, 100
    ops = operators_cached(L, 1e-12, 0, 1)
    std = Solver2DBatched(L, max_rank=cap, n_steps=steps, host_ops=ops)
    eng = std.engine
    f = BatchedTT.from_host([c5_rhs_cores(i, L) for i in range(N)], eng.device)

    def timed(solver):
        solver.solve(f); torch.cuda.synchronize()
        t0 = time.perf_counter(); res = solver.solve(f); torch.cuda.synchronize()
        return res, time.perf_counter() - t0

    def true_residual(res):
        # ||b - B v||^2 / ||b||^2 with the assembled operator, cap 32 (the product has bond 644; the solution rank is 4)
        r = eng.round_sum([(None, res["b"]), (std.ops["B"], res["v"], None, -1.0)], caps=32, rel_eps=1e-12)
        return (eng.norm2(r) / eng.norm2(res["b"])).cpu().numpy()

    res_s, t_s = timed(std)
    rec = {"L": L, "N": N, "max_rank": cap, "gd_steps": steps, "assembled": {"seconds_per_batched_solve": t_s, "solves_per_s": N / t_s,
           "true_rel_residual_median": float(np.median(true_residual(res_s)))}}
    sub = slice(0, min(N, 16))
    rs_sub = true_residual({"b": res_s["b"], "v": res_s["v"]})
    for inner in (4, 8, 16):
        fac = FactoredSolver2D(L, max_rank=cap, inner_rank=inner, n_steps=steps, host_ops=ops)
        res_f, t_f = timed(fac)
        dv = eng.round_sum([(None, res_f["v"]), (None, res_s["v"], None, -1.0)], caps=None, rel_eps=1e-14)
        dvn = torch.sqrt(eng.norm2(dv) / eng.norm2(res_s["v"])).cpu().numpy()
        de = (torch.abs(res_f["energy"] - res_s["energy"]) / res_s["energy"]).cpu().numpy()
        dl = (torch.abs(res_f["l2"] - res_s["l2"]) / res_s["l2"]).cpu().numpy()
        rf = true_residual(res_f)
        rec[f"factored_inner{inner}"] = {"seconds_per_batched_solve": t_f, "solves_per_s": N / t_f, "speedup": t_s / t_f,
                                         "rel_dev_solution_median": float(np.median(dvn)), "rel_dev_solution_max": float(dvn.max()),
                                         "rel_dev_energy_median": float(np.median(de)), "rel_dev_energy_max": float(de.max()),
                                         "rel_dev_l2_median": float(np.median(dl)), "rel_dev_l2_max": float(dl.max()),
                                         "true_rel_residual_median": float(np.median(rf)),
                                         "true_rel_residual_ratio_to_assembled_median": float(np.median(rf / rs_sub))}
        print(inner, json.dumps(rec[f"factored_inner{inner}"]), flush=True)
    print(json.dumps(rec["assembled"]))
    if out:
        json.dump(rec, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
