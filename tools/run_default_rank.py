"""The reference's 2D driver at its DEFAULT arguments on the GPU: `solve_PDE_2D_with_preconditioner(get_example_f_2D(L))`
(solver.py:212-233, max_rank = 60) followed by the post-processing of src/experiments/2D_sweep_L.py:38-73 - through the
drop-in `laplace_mps` API, i.e. every `@`, `reapprox`, `norm_squared` below runs in the CUDA library.

    python tools/run_default_rank.py [L=12] [out.json]

Reports the four quantities of the experiment (L2 / H1 error against the manufactured solution, error of the L2 / H1 norm
against the analytic values) next to the discretisation error of the interpolant at the same L (2D_norm_of_interpolant.py),
which is the floor any solver can reach at that level."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run(L=12, **solver_options):
    from laplace_mps.solver import solve_PDE_2D_with_preconditioner, get_L2_norm_2D
    from laplace_mps.utils import (get_example_u_2D, get_example_f_2D, _get_gram_matrix_tt, _get_identy_as_tt, kronecker_prod_2D,
                                   get_example_grad_u_2D)
    import laplace_mps.tensormethods as tm
    import laplacesolvermps_b200.tensormethods as mine
    assert tm is mine
    refnorm_L2 = (1 / 15 + 3 / (16 * np.pi ** 4) - 3 / (16 * np.pi ** 2)) * (67 / 210)
    refnorm_H1 = (2144 * np.pi ** 6 + 5926 * np.pi ** 4 + 7245 - 9255 * np.pi ** 2) / (25200 * np.pi ** 4)
    rel_error, max_rank = 1e-14, 60
    out = dict(L=L, max_rank=max_rank, solver_options=solver_options)
    t0 = time.perf_counter()
    u_ref = get_example_u_2D(L, basis='nodal').reshape_mode_indices([4])
    f = get_example_f_2D(L)
    DuDx_ref, DuDy_ref = get_example_grad_u_2D(L)
    t1 = time.perf_counter()
    u, DuDx, DuDy = solve_PDE_2D_with_preconditioner(f, **solver_options)          # default max_rank = 60
    t2 = time.perf_counter()
    G = _get_gram_matrix_tt(L)
    G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)
    I = _get_identy_as_tt(L, True)
    Gx, Gy = kronecker_prod_2D(G, I), kronecker_prod_2D(I, G)
    delta_u = (u - u_ref).reshape_mode_indices([4]).reapprox(rel_error=rel_error, ranks_new=max_rank)
    out["L2_delta_u"] = float(np.sqrt(get_L2_norm_2D(delta_u)))
    dDx = (DuDx - DuDx_ref).reapprox(rel_error=rel_error, ranks_new=max_rank)
    dDy = (DuDy - DuDy_ref).reapprox(rel_error=rel_error, ranks_new=max_rank)
    out["H1_delta_u"] = float(np.sqrt(0.25 ** L * ((Gy @ dDx).norm_squared() + (Gx @ dDy).norm_squared())))
    H1_norm = 0.25 ** L * ((Gy @ DuDx).norm_squared() + (Gx @ DuDy).norm_squared())
    out["rel_error_of_H1_norm"] = float(H1_norm / refnorm_H1 - 1)
    out["rel_error_of_L2_norm"] = float(get_L2_norm_2D(u) / refnorm_L2 - 1)
    # the floor at this level: the same norms of the interpolated exact solution (2D_norm_of_interpolant.py)
    H1_ip = 0.25 ** L * ((Gy @ DuDx_ref).norm_squared() + (Gx @ DuDy_ref).norm_squared())
    out["interpolant_rel_error_of_H1_norm"] = float(H1_ip / refnorm_H1 - 1)
    out["interpolant_rel_error_of_L2_norm"] = float(get_L2_norm_2D(u_ref.copy()) / refnorm_L2 - 1)
    t3 = time.perf_counter()
    out["u_ranks"] = [int(r) for r in u.ranks]
    out["seconds"] = dict(setup=t1 - t0, solve=t2 - t1, postprocessing=t3 - t2)
    return out


if __name__ == "__main__":
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    res = run(L)
    line = json.dumps(res)
    print(line)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as fh:
            fh.write(line + "\n")
