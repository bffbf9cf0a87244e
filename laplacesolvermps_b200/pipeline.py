"""Batched, device-resident version of `solve_PDE_2D_with_preconditioner` (solver.py:212-233) with the
gradient-descent solver (solver.py:251-275) behind it, plus the energy / L2 norms the experiments
compute from a solution (2D_sweep_L.py:52-64, solver.py:430-441).  This is benchmark config C5
(SURVEY.md section 8d): many independent right-hand sides, one shared set of operators per GPU.

One *solve* of instance i:
    b_i  = round((R @ f_i).squeeze(), max_rank, eps)                     RHS into the BPX basis
    v_i  = steepest descent on B v = b_i, fixed number of steps, rank cap max_rank
    u_i  = round(C @ v_i, max_rank, eps)                                  back-transform
    E_i  = 4^-L (||G_y Q'_x v_i||^2 + ||G_x Q'_y v_i||^2)                 energy norm (orthogonalised route)
    l2_i = 4^L ||(G^1/2 M (x) G^1/2 M) u_i||^2                            L2 norm
Nothing is materialised: every product above is a term of a fused rounding / orthogonalisation sweep.
"""
import numpy as np
import torch

from . import bpx2D, solver, tensormethods as tm
from . import _assembly as _asm
from .batched import BatchedTT, DeviceMPO, default_engine
from .utils import kronecker_prod_2D, _get_gram_matrix_tt


def build_operators_2D(L, eps=1e-12, on_device=True):
    """Host cores of every operator a batched 2D solve needs (construction time, cached by the caller).  The rounding of
    the raw rank-1152 system operator runs on the GPU unless on_device=False (CPU reference arm, CPU-only tests)."""
    B = solver._cached_laplace_BPX_2D(L, eps, on_device=on_device)
    R = bpx2D.get_rhs_matrix_BPX_2D(L)
    C = bpx2D.get_BPX_preconditioner_2D(L)
    eye = tm.TensorTrain([np.eye(2)[None, :, :, None] for _ in range(L)] + [np.eye(1)[None, :, :, None]])
    G = _get_gram_matrix_tt(L)
    G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)          # as in experiments/2D_sweep_L.py:44-45
    Gy, Gx = kronecker_prod_2D(eye, G), kronecker_prod_2D(G, eye)
    GyQx = tm.TensorTrain(_asm.compose(Gy.tensors, bpx2D.get_BPX_Qp_2D(L, 'x').tensors))
    GxQy = tm.TensorTrain(_asm.compose(Gx.tensors, bpx2D.get_BPX_Qp_2D(L, 'y').tensors))
    GM = solver._sqrt_gram_times_overlap(L)
    GM2 = solver._kron_xy(GM, GM).reshape_mode_indices([(4, 4) for _ in range(L)] + [(4, 1)])
    return dict(B=B.tensors, R=R.tensors, C=C.tensors, GyQx=GyQx.tensors, GxQy=GxQy.tensors, GM2=GM2.tensors,
                Qx=bpx2D.get_BPX_Qp_2D(L, 'x').tensors, Qy=bpx2D.get_BPX_Qp_2D(L, 'y').tensors, Gx=Gx.tensors, Gy=Gy.tensors)


class Solver2DBatched:
    """Operators resident on one GPU + the batched solve."""

    def __init__(self, L, max_rank=4, n_steps=100, eps=1e-12, lr=0.8, recalc_every=10, device=None, host_ops=None, method="gd"):
        self.L, self.max_rank, self.n_steps, self.eps, self.lr, self.recalc_every = L, max_rank, n_steps, eps, lr, recalc_every
        assert method in ("gd", "cg")
        self.method = method                  # "gd": the reference's steepest descent (benchmark); "cg": truncated conjugate gradients
        self.engine = default_engine(device)
        self.device = self.engine.device
        self.host_ops = host_ops if host_ops is not None else build_operators_2D(L, eps)
        self.upload_operators()

    def upload_operators(self):
        self.ops = {k: DeviceMPO(v, self.device) for k, v in self.host_ops.items()}

    def operator_bytes(self):
        return sum(op.nbytes() for op in self.ops.values())

    @staticmethod
    def _extend_with_ones(x):
        """x with a trailing (1,1,1) core of ones - `v.tensors.append(np.ones([1,1,1,1]))` of the reference."""
        ext = BatchedTT.__new__(BatchedTT)
        ext.mode_shapes = x.mode_shapes + [(1,)]
        ext.L, ext.N, ext.n, ext.slot, ext.device = x.L + 1, x.N, x.n + [1], x.slot + [1], x.device
        ext.cores = x.cores + [torch.ones(x.N, dtype=torch.float64, device=x.device)]
        ext.ranks = torch.cat([x.ranks, torch.ones((x.N, 1), dtype=torch.int32, device=x.device)], dim=1).contiguous()
        ext._c = None
        return ext

    @staticmethod
    def _fold_last(y):
        """(R @ f).squeeze(): the trailing core has mode size 1 and, after rounding, bond 1 - a scalar."""
        out = BatchedTT.__new__(BatchedTT)
        L = y.L - 1
        out.mode_shapes, out.L, out.N, out.n, out.slot, out.device = y.mode_shapes[:L], L, y.N, y.n[:L], y.slot[:L], y.device
        out.cores = list(y.cores[:L])
        out.cores[L - 1] = (y.cores[L - 1].view(y.N, -1) * y.cores[L].view(y.N, -1)[:, :1]).reshape(-1)
        out.ranks = y.ranks[:, :L + 1].contiguous()
        out._c = None
        return out

    def solve(self, f, want_u=True):
        """f: BatchedTT of L+1 cores (mode 4).  Returns dict of device tensors / batches."""
        eng, ops, cap, L = self.engine, self.ops, self.max_rank, self.L
        b_ext = eng.round_sum([(ops["R"], f)], caps=[cap] * L, rel_eps=self.eps)
        assert int(b_ext.ranks[:, L].max()) == 1
        b = self._fold_last(b_ext)
        if self.method == "cg":
            v, r2, steps, _ = eng.cg_solve(ops["B"], b, n_steps=self.n_steps, max_rank=cap, recalc_every=self.recalc_every, rel_accuracy=0.0)
        else:
            v, r2, steps, _ = eng.gd_solve(ops["B"], b, n_steps=self.n_steps, lr=self.lr, max_rank=cap,
                                           recalc_every=self.recalc_every, rel_accuracy=0.0)
        u = eng.round_sum([(ops["C"], v)], caps=cap, rel_eps=self.eps)
        v_ext = self._extend_with_ones(v)
        ex = eng.orthogonalize([(ops["GyQx"], v_ext)], want_cores=False)[1]
        ey = eng.orthogonalize([(ops["GxQy"], v_ext)], want_cores=False)[1]
        energy = (ex + ey) * 0.25 ** L
        l2 = eng.orthogonalize([(ops["GM2"], self._extend_with_ones(u))], want_cores=False)[1] * 4.0 ** L
        out = dict(energy=energy, l2=l2, r2=r2, v=v, b=b)
        if want_u:
            out["u"] = u
        return out

    def _broadcast(self, tt, N):
        """A single host train (TensorTrain or list of cores, any mode shapes) as an N-instance batch with flattened modes."""
        if isinstance(tt, BatchedTT):
            return tt
        cores = [np.ascontiguousarray(c).reshape(c.shape[0], -1, c.shape[-1]) for c in (tt.tensors if hasattr(tt, "tensors") else tt)]
        return BatchedTT.from_host([cores] * N, self.device)

    def errors(self, res, u_exact, Dx_exact=None, Dy_exact=None, max_rank=60, eps=None):
        """Error of a batch of solutions against exact solutions, device resident (2D_sweep_L.py:44-70):

            delta   = round(u - u_exact, max_rank, eps)                 l2_err2 = ||delta||^2_L2           (get_L2_norm_2D, solver.py:430-441)
            dDx     = round(Q'_x v - Dx_exact, max_rank, eps)  (dDy alike)
            h1_err2 = 4^-L (||G_y dDx||^2 + ||G_x dDy||^2)

        `res` is the dict returned by `solve`; the exact trains are `BatchedTT`s (one per instance) or single trains that are
        broadcast (e.g. `utils.get_example_u_2D(L, 'nodal')`, `utils.get_example_grad_u_2D(L)`).  Differences and products are terms
        of fused roundings / orthogonalisation sweeps; nothing returns to the host but the result scalars the caller reads."""
        eng, ops, L = self.engine, self.ops, self.L
        eps = self.eps if eps is None else eps
        u, v = res["u"], res["v"]
        N = u.N
        u_ref = self._broadcast(u_exact, N)
        delta = eng.round_sum([(None, u), (None, u_ref, None, -1.0)], caps=max_rank, rel_eps=eps)
        out = dict(delta=delta)
        out["l2_err2"] = eng.orthogonalize([(ops["GM2"], self._extend_with_ones(delta))], want_cores=False)[1] * 4.0 ** L
        if Dx_exact is not None and Dy_exact is not None:
            v_ext = self._extend_with_ones(v)
            h1 = None
            for q, g, d_exact in (("Qx", "Gy", Dx_exact), ("Qy", "Gx", Dy_exact)):
                d_ref = self._broadcast(d_exact, N)
                dd = eng.round_sum([(ops[q], v_ext), (None, d_ref, None, -1.0)], caps=max_rank, rel_eps=eps)
                e = eng.orthogonalize([(ops[g], dd)], want_cores=False)[1]
                h1 = e if h1 is None else h1 + e
            out["h1_err2"] = h1 * 0.25 ** L
        return out

    def solve_host(self, f_instances, pinned=True):
        """End-to-end call with HOST buffers: numpy cores in, numpy results out (the benchmark's e2e leg)."""
        f = BatchedTT.from_host(f_instances, self.device, pin=pinned)
        res = self.solve(f)
        scal = torch.stack([res["energy"], res["l2"], res["r2"]], dim=1).cpu().numpy()
        u_host = res["u"].to_host()
        return scal, u_host


class FactoredSolver2D(Solver2DBatched):
    """SURVEY 8(f).2 - the system operator applied in FACTORED form, B = S_x^T S_x + S_y^T S_y with S_d = 2^-L G_d'^{1/2} Q'_d
    (rank 24; bpx2D.py:159-171 assembles and rounds the same sum to rank 161):

        w_d   = round(S_d r, inner_rank)                      rank 24 r        -> inner_rank
        B r  ~= round(S_x^T w_x + S_y^T w_y, max_rank)        rank 48 inner_rank -> max_rank

    The bonds the orthogonalisation sweeps meet are 96 and 48 * inner_rank instead of 644 (at max_rank 4), so one
    application costs a fraction of the assembled one.  It is NOT the reference's arithmetic: the intermediate rounding
    moves the truncation points, so iterates differ from the assembled path at the level of the truncation error and rank
    profiles of intermediates are not comparable - this class is an extension with its own accuracy study
    (tools/study_factored.py, profiles/), not part of the parity path or of the headline benchmark.
    The energy <r, B r> = ||S_x r||^2 + ||S_y r||^2 is available from the first stage."""

    def __init__(self, L, max_rank=4, inner_rank=8, n_steps=100, eps=1e-12, lr=0.8, recalc_every=10, device=None, host_ops=None):
        super().__init__(L, max_rank=max_rank, n_steps=n_steps, eps=eps, lr=lr, recalc_every=recalc_every, device=device, host_ops=host_ops)
        self.inner_rank = inner_rank
        scale = 0.5 ** L
        self.S, self.St = {}, {}
        for d, key in (("x", "GyQx"), ("y", "GxQy")):
            cores = [np.array(c, copy=True) for c in self.host_ops[key]]
            cores[0] = cores[0] * scale
            self.S[d] = DeviceMPO(cores, self.device)
            self.St[d] = DeviceMPO([np.ascontiguousarray(np.transpose(c, (0, 2, 1, 3))) for c in cores], self.device)

    def apply_terms(self, x, coef=None, scale=1.0):
        """Terms of coef * scale * (B @ x) for a fused rounding; x has L cores.  Also returns <x, B x> per instance."""
        eng = self.engine
        x_ext = self._extend_with_ones(x)
        terms, energy = [], None
        for d in ("x", "y"):
            w = eng.round_sum([(self.S[d], x_ext)], caps=self.inner_rank, rel_eps=1e-16)
            e = eng.norm2(w)
            energy = e if energy is None else energy + e
            terms.append((self.St[d], w, coef, scale))
        return terms, energy

    def _round_ext(self, terms, caps):
        """Rounding of a sum of (L+1)-core terms whose last mode has size 1, returned as an L-core batch."""
        y = self.engine.round_sum(terms, caps=[caps] * self.L, rel_eps=1e-16)
        return self._fold_last(y)

    def gd_solve(self, b):
        """solver.py:251-275 with the factored application (same step rule, same recalculation period)."""
        eng, cap, L = self.engine, self.max_rank, self.L
        N = b.N
        zero = [np.zeros((1, 4, 1)) for _ in range(L)]
        x = BatchedTT.from_host([zero] * N, self.device)
        b_ext = self._extend_with_ones(b)
        r = None
        for n in range(self.n_steps):
            if n % self.recalc_every == 0:
                tb, _ = self.apply_terms(x, scale=-1.0)
                r = self._round_ext([(None, b_ext)] + tb, cap)
            tr, rBr = self.apply_terms(r)
            Ar = self._round_ext(tr, cap)
            r2 = eng.norm2(r)
            gamma = self.lr * r2 / eng.inner(r, Ar)
            x = eng.round_sum([(None, x), (None, r, gamma)], caps=cap, rel_eps=1e-16)
            r = eng.round_sum([(None, r), (None, Ar, -gamma)], caps=cap, rel_eps=1e-16)
        tb, _ = self.apply_terms(x, scale=-1.0)
        r = self._round_ext([(None, b_ext)] + tb, cap)
        return x, eng.norm2(r)

    def solve(self, f, want_u=True):
        eng, ops, cap, L = self.engine, self.ops, self.max_rank, self.L
        b_ext = eng.round_sum([(ops["R"], f)], caps=[cap] * L, rel_eps=self.eps)
        b = self._fold_last(b_ext)
        v, r2 = self.gd_solve(b)
        u = eng.round_sum([(ops["C"], v)], caps=cap, rel_eps=self.eps)
        v_ext = self._extend_with_ones(v)
        ex = eng.orthogonalize([(ops["GyQx"], v_ext)], want_cores=False)[1]
        ey = eng.orthogonalize([(ops["GxQy"], v_ext)], want_cores=False)[1]
        l2 = eng.orthogonalize([(ops["GM2"], self._extend_with_ones(u))], want_cores=False)[1] * 4.0 ** L
        out = dict(energy=(ex + ey) * 0.25 ** L, l2=l2, r2=r2, v=v, b=b)
        if want_u:
            out["u"] = u
        return out


def build_operators_1D(L):
    """Host cores of the operators of `solve_PDE_1D_with_preconditioner` (solver.py:192-210) and of the norms (solver.py:420-428)."""
    GM = solver._sqrt_gram_times_overlap(L)
    R = [np.asarray(c) for c in solver.get_rhs_matrix_bpx(L).tensors]
    R[-1] = R[-1].reshape(R[-1].shape[0], 1, 2, 1)          # (r, n_in = 2, 1) -> operator core (r, n_out = 1, n_in = 2, 1)
    return dict(B=solver.get_laplace_bpx(L).tensors, R=R, C=solver.get_bpx_preconditioner(L).tensors,
                Qp=solver.get_bpx_Qp(L).tensors, GM=GM.tensors)


class Solver1DBatched:
    """Batched, device-resident `solve_PDE_1D_with_preconditioner` (solver.py:192-210) with steepest descent
    (solver.py:251-275) behind it - benchmark configs C1 / C2: many right-hand sides f_i (L + 1 cores,
    mode 2), one shared set of rank <= 21 operators.

        b_i  = 2^-L (R @ f_i).squeeze()                       (exact product, as the reference: no rounding of b)
        v_i  = steepest descent on B v = b_i, rank cap max_rank
        u_i  = round(C @ v_i, eps),  Du_i = round(Q' @ v_i, eps)
        l2_i = 2^L ||G^1/2 M u_i||^2,  h1_i = 2^-L ||Du_i||^2"""

    def __init__(self, L, max_rank=20, n_steps=300, eps=1e-15, lr=0.8, recalc_every=10, rel_accuracy=0.0, device=None, host_ops=None):
        self.L, self.max_rank, self.n_steps, self.eps, self.lr = L, max_rank, n_steps, eps, lr
        self.recalc_every, self.rel_accuracy = recalc_every, rel_accuracy
        self.engine = default_engine(device)
        self.device = self.engine.device
        self.host_ops = host_ops if host_ops is not None else build_operators_1D(L)
        self.ops = {k: DeviceMPO(v, self.device) for k, v in self.host_ops.items()}

    def rhs(self, f):
        eng, L = self.engine, self.L
        R = self.ops["R"]
        y = eng.matmul(R, f, R.n_in, [(no,) for no in R.n_out])           # exact product, L + 1 cores, last mode size 1
        N = y.N
        rk = y.ranks_host()
        assert (rk == rk[0]).all(), "right-hand sides of one batch must share their rank profile"
        r_last = int(rk[0, L])
        out = BatchedTT.__new__(BatchedTT)
        out.mode_shapes, out.L, out.N, out.n, out.device = y.mode_shapes[:L], L, N, y.n[:L], y.device
        prev = y.cores[L - 1].view(N, -1)[:, :int(rk[0, L - 1]) * y.n[L - 1] * r_last].reshape(N, -1, r_last)
        tail = y.cores[L].view(N, -1)[:, :r_last].reshape(N, r_last, 1)
        out.cores = list(y.cores[:L - 1]) + [torch.bmm(prev, tail).reshape(-1).contiguous()]
        out.slot = y.slot[:L - 1] + [int(rk[0, L - 1]) * y.n[L - 1]]
        out.ranks = y.ranks[:, :L + 1].clone()
        out.ranks[:, L] = 1
        out._c = None
        out.cores[0] = out.cores[0] * 0.5 ** L                              # `b.tensors[i] /= 2` for every level (solver.py:200-202)
        return out

    def solve(self, f):
        eng, ops, L = self.engine, self.ops, self.L
        b = self.rhs(f)
        v, r2, steps, _ = eng.gd_solve(ops["B"], b, n_steps=self.n_steps, lr=self.lr, max_rank=self.max_rank,
                                       recalc_every=self.recalc_every, rel_accuracy=self.rel_accuracy)
        u = eng.round_sum([(ops["C"], v)], caps=None, rel_eps=self.eps)
        Du = eng.round_sum([(ops["Qp"], v)], caps=None, rel_eps=self.eps)
        l2 = eng.orthogonalize([(ops["GM"], Solver2DBatched._extend_with_ones(u))], want_cores=False)[1] * 2.0 ** L
        h1 = eng.norm2(Du) * 0.5 ** L
        return dict(l2=l2, h1=h1, r2=r2, v=v, u=u, Du=Du, b=b, steps=steps)


_solver_cache = {}


def solve_2D_mixed_levels(f_instances, max_rank=4, n_steps=100, eps=1e-12, device=None, want_u=False, **kw):
    """Batching over LEVEL COUNTS (north star: "right-hand sides, coefficient fields, level counts L"): right-hand sides with
    different numbers of cores are bucketed by L; every bucket runs as one batched solve on the operators of its L (built once
    and cached per (L, max_rank, n_steps, eps)).  Different L cannot share launches - the operators differ - but they share the
    handle, its workspace and the stream, so the buckets run back to back without host round trips in between.

    f_instances: list of host trains (lists of cores, L_i + 1 cores of mode size 4).  Returns a list of dicts (energy, l2, r2, L[, u])
    in the order of the input."""
    eng = default_engine(device)
    buckets = {}
    for pos, f in enumerate(f_instances):
        cores = f.tensors if hasattr(f, "tensors") else f
        buckets.setdefault(len(cores) - 1, []).append((pos, [np.ascontiguousarray(c).reshape(c.shape[0], -1, c.shape[-1]) for c in cores]))
    out = [None] * len(f_instances)
    pending = []
    for L, items in sorted(buckets.items()):
        key = (L, max_rank, n_steps, float(eps), str(eng.device), tuple(sorted(kw.items())))
        if key not in _solver_cache:
            _solver_cache[key] = Solver2DBatched(L, max_rank=max_rank, n_steps=n_steps, eps=eps, device=eng.device, **kw)
        res = _solver_cache[key].solve(BatchedTT.from_host([c for _, c in items], eng.device), want_u=want_u)
        pending.append((L, items, res))
    for L, items, res in pending:                       # one device->host read per bucket, after everything is enqueued
        scal = torch.stack([res["energy"], res["l2"], res["r2"]], dim=1).cpu().numpy()
        u_host = res["u"].to_host() if want_u else None
        for j, (pos, _) in enumerate(items):
            out[pos] = dict(energy=float(scal[j, 0]), l2=float(scal[j, 1]), r2=float(scal[j, 2]), L=L)
            if want_u:
                out[pos]["u"] = u_host[j]
    return out
