"""Device-resident batched tensor trains and the engine that drives the qtt_b200 kernels.

`BatchedTT` holds N independent instances of an L-core TT in the layout of include/qtt_b200.h (compact
C-ordered cores in fixed-size slots + an int32 rank table), `DeviceMPO` one operator shared by all
instances, `Engine` wraps a library handle bound to one GPU / one CUDA stream.  PyTorch is used for
device memory, streams and host<->device copies only.
"""
import ctypes
from ctypes import POINTER, byref, c_double, c_int, c_longlong, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import qtt_batch, qtt_mpo, qtt_term


def _prod(shape):
    out = 1
    for s in shape:
        out *= int(s)
    return out


class BatchedTT:
    """N instances x L cores on one device.  mode_shapes[k] is the tuple of mode dimensions of core k
    (kept for unpacking; the kernels see the flattened size n[k])."""

    def __init__(self, mode_shapes, N, slot, device):
        self.mode_shapes = [tuple(int(d) for d in m) for m in mode_shapes]
        self.L = len(self.mode_shapes)
        self.N = int(N)
        self.n = [_prod(m) for m in self.mode_shapes]
        self.slot = [int(s) for s in slot]
        self.device = torch.device(device)
        self.cores = [torch.zeros(self.N * max(s, 1), dtype=torch.float64, device=self.device) for s in self.slot]
        self.ranks = torch.ones((self.N, self.L + 1), dtype=torch.int32, device=self.device)
        self._c = None

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def for_rank_bound(cls, mode_shapes, N, rank_bound, device):
        """Empty batch whose slots fit bonds <= rank_bound[k] (int or list of L-1)."""
        L = len(mode_shapes)
        if isinstance(rank_bound, int):
            rank_bound = [rank_bound] * (L - 1)
        rb = [1] + [int(r) for r in rank_bound] + [1]
        slot = [rb[k] * _prod(mode_shapes[k]) * rb[k + 1] for k in range(L)]
        return cls(mode_shapes, N, slot, device)

    @classmethod
    def from_host(cls, instances, device, slot=None, pin=True):
        """instances: list (length N) of lists of numpy cores (r, *modes, r')."""
        N = len(instances)
        L = len(instances[0])
        mode_shapes = [tuple(c.shape[1:-1]) for c in instances[0]]
        ranks = np.ones((N, L + 1), dtype=np.int32)
        for i, tt in enumerate(instances):
            assert len(tt) == L, "all instances need the same number of cores"
            for k, c in enumerate(tt):
                assert tuple(c.shape[1:-1]) == mode_shapes[k], "mode shapes differ between instances"
                ranks[i, k] = c.shape[0]
                ranks[i, k + 1] = c.shape[-1]
        n = [_prod(m) for m in mode_shapes]
        need = [int(max(ranks[:, k] * n[k] * ranks[:, k + 1])) for k in range(L)]
        if slot is None:
            slot = need
        else:
            assert all(s >= q for s, q in zip(slot, need)), "slot too small"
        out = cls(mode_shapes, N, slot, device)
        for k in range(L):
            host = torch.zeros((N, max(slot[k], 1)), dtype=torch.float64)
            if pin and out.device.type == "cuda":
                host = host.pin_memory()
            hv = host.numpy()
            for i, tt in enumerate(instances):
                flat = np.ascontiguousarray(tt[k], dtype=np.float64).reshape(-1)
                hv[i, :flat.size] = flat
            out.cores[k].view(N, -1).copy_(host, non_blocking=True)
        out.ranks.copy_(torch.from_numpy(ranks), non_blocking=False)
        return out

    @staticmethod
    def pack_pinned(instances, slot=None):
        """Host-side packing into pinned buffers (done once, outside any timed region): returns
        (mode_shapes, slot, [pinned core tensors N x slot_k], pinned rank table)."""
        tmp = BatchedTT.from_host(instances, "cpu", slot=slot, pin=False)
        cores = [c.view(tmp.N, -1).pin_memory() for c in tmp.cores]
        return tmp.mode_shapes, tmp.slot, cores, tmp.ranks.pin_memory()

    @classmethod
    def from_pinned(cls, packed, device):
        """Host -> device copy of a pre-packed batch (the H2D leg of an end-to-end step)."""
        mode_shapes, slot, cores, ranks = packed
        out = cls(mode_shapes, ranks.shape[0], slot, device)
        for dst, src in zip(out.cores, cores):
            dst.view(ranks.shape[0], -1).copy_(src, non_blocking=True)
        out.ranks.copy_(ranks, non_blocking=True)
        return out

    def to_host(self):
        """list (N) of lists (L) of numpy cores with the stored mode shapes."""
        ranks = self.ranks.cpu().numpy()
        host = [c.view(self.N, -1).cpu().numpy() for c in self.cores]
        out = []
        for i in range(self.N):
            tt = []
            for k in range(self.L):
                r0, r1 = int(ranks[i, k]), int(ranks[i, k + 1])
                cnt = r0 * self.n[k] * r1
                tt.append(host[k][i, :cnt].reshape((r0,) + self.mode_shapes[k] + (r1,)).copy())
            out.append(tt)
        return out

    def ranks_host(self):
        return self.ranks.cpu().numpy()

    def clone(self):
        out = BatchedTT(self.mode_shapes, self.N, self.slot, self.device)
        for a, b in zip(out.cores, self.cores):
            a.copy_(b)
        out.ranks.copy_(self.ranks)
        return out

    # -- C view ---------------------------------------------------------------------------------
    def c_struct(self):
        if self._c is None:
            n_arr = (c_int * self.L)(*self.n)
            slot_arr = (c_longlong * self.L)(*self.slot)
            core_arr = (c_void_p * self.L)(*[c.data_ptr() for c in self.cores])
            st = qtt_batch(self.L, self.N, n_arr, c_void_p(self.ranks.data_ptr()), core_arr, slot_arr)
            self._c = (st, n_arr, slot_arr, core_arr)
        return self._c[0]


class DeviceMPO:
    """One operator, cores (R, n_out, n_in, R') resident on the device."""

    def __init__(self, cores, device):
        self.device = torch.device(device)
        self.L = len(cores)
        self.host_cores = [np.ascontiguousarray(c, dtype=np.float64) for c in cores]
        for c in self.host_cores:
            assert c.ndim == 4, "operator cores must be (R, n_out, n_in, R')"
        self.n_out = [int(c.shape[1]) for c in self.host_cores]
        self.n_in = [int(c.shape[2]) for c in self.host_cores]
        self.ranks = [int(self.host_cores[0].shape[0])] + [int(c.shape[3]) for c in self.host_cores]
        self.cores = [torch.from_numpy(c.reshape(-1)).to(self.device) for c in self.host_cores]
        self._no = (c_int * self.L)(*self.n_out)
        self._ni = (c_int * self.L)(*self.n_in)
        self._rk = (c_int * (self.L + 1))(*self.ranks)
        self._cp = (c_void_p * self.L)(*[c.data_ptr() for c in self.cores])
        self._c = qtt_mpo(self.L, self._no, self._ni, self._rk, self._cp)

    def c_struct(self):
        return self._c

    def nbytes(self):
        return sum(c.numel() * 8 for c in self.cores)


class Engine:
    """A qtt_b200 handle on one device.  All work goes to torch's current stream of that device."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.QttError("CUDA device required: the TT hot path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        h = c_void_p()
        rc = self.lib.qtt_create(byref(h), self.device.index or 0)
        if rc != 0:
            raise _lib.QttError(f"qtt_create failed with status {rc}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.qtt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_workspace_limit(self, nbytes):
        _lib.check(self.h, self.lib.qtt_set_workspace_limit(self.h, int(nbytes)), "qtt_set_workspace_limit")

    def counters(self):
        f, l = c_double(), c_longlong()
        self.lib.qtt_get_counters(self.h, byref(f), byref(l))
        return f.value, l.value

    def flop_counters(self):
        """(executed flops of the DMMA-class kernels, executed flops of the panel QR leaves) since the last reset."""
        d, l = c_double(), c_double()
        self.lib.qtt_get_flop_counters(self.h, byref(d), byref(l))
        return d.value, l.value

    def graph_counters(self):
        """(sweeps captured into a CUDA graph, sweeps replayed from one) since the handle was created."""
        c, r = c_longlong(), c_longlong()
        self.lib.qtt_get_graph_counters(self.h, byref(c), byref(r))
        return c.value, r.value

    def reset_counters(self):
        self.lib.qtt_reset_counters(self.h)

    def set_profiling(self, on):
        """0 off, 1 / True: GEMM-class event pairs, 2: timeline of every launch (see class_profile)."""
        _lib.check(self.h, self.lib.qtt_set_profiling(self.h, int(on)), "qtt_set_profiling")

    def class_profile(self, reset=True):
        """ms per kernel class since the last reset (timeline mode): gemm, larfb, leaf, gram_larft, jacobi, other."""
        arr = (c_double * 8)()
        _lib.check(self.h, self.lib.qtt_get_class_profile(self.h, arr, 1 if reset else 0), "qtt_get_class_profile")
        return dict(zip(["gemm", "larfb", "leaf", "gram_larft", "jacobi", "other"], list(arr)[:6]))

    def profile(self, reset=True):
        """(summed GEMM-kernel ms, algorithmic flops of those launches, launch count) since the last reset."""
        ms, fl, n = c_double(), c_double(), c_longlong()
        _lib.check(self.h, self.lib.qtt_get_profile(self.h, byref(ms), byref(fl), byref(n), 1 if reset else 0), "qtt_get_profile")
        return ms.value, fl.value, n.value

    # -- terms ----------------------------------------------------------------------------------
    @staticmethod
    def _terms(terms):
        """terms: list of (op | None, BatchedTT, coef tensor | None, scale)."""
        arr = (qtt_term * len(terms))()
        keep = []
        for i, t in enumerate(terms):
            op, x = t[0], t[1]
            coef = t[2] if len(t) > 2 else None
            scale = t[3] if len(t) > 3 else 1.0
            xs = x.c_struct()
            arr[i].op = ctypes.pointer(op.c_struct()) if op is not None else POINTER(qtt_mpo)()
            arr[i].x = ctypes.pointer(xs)
            arr[i].coef = c_void_p(coef.data_ptr()) if coef is not None else c_void_p()
            arr[i].coef_scale = float(scale)
            keep.append((op, x, coef, xs))
        return arr, keep

    @staticmethod
    def _out_modes(terms):
        op, x = terms[0][0], terms[0][1]
        if op is None:
            return x.mode_shapes
        return [(no,) for no in op.n_out]

    def rank_bounds_after_round(self, terms, caps):
        """Host-side bound on the output bonds (used to size output slots)."""
        L = terms[0][1].L
        modes = self._out_modes(terms)
        n = [_prod(m) for m in modes]
        p = [1] * (L + 1)
        for k in range(1, L):
            s = 0
            for t in terms:
                op, x = t[0], t[1]
                xr = x.slot[k - 1] // max(x.n[k - 1], 1)   # crude: right bond <= slot / n (left bond >= 1)
                s += xr * (op.ranks[k] if op is not None else 1)
            p[k] = s
        q = [1] * (L + 1)
        for k in range(L - 1, 0, -1):
            q[k] = min(n[k] * q[k + 1], p[k])
        sb = [1] * (L + 1)
        for k in range(L - 1):
            cap = caps[k] if caps is not None else 1 << 30
            sb[k + 1] = min(sb[k] * n[k], q[k + 1], cap)
        return sb, q

    # -- ops ------------------------------------------------------------------------------------
    def round_sum(self, terms, caps=None, rel_eps=1e-16, out=None, status=None):
        """y = round(sum_t coef_t * term_t) - see qtt_round_sum_batched."""
        x0 = terms[0][1]
        L, N = x0.L, x0.N
        if isinstance(caps, (int, np.integer)):
            caps = [int(caps)] * (L - 1)
        modes = self._out_modes(terms)
        if out is None:
            ranks_h = [t[1].ranks_host() for t in terms]
            n = [_prod(m) for m in modes]
            p = np.zeros(L + 1, dtype=np.int64)
            for t, rh in zip(terms, ranks_h):
                mult = np.array(t[0].ranks if t[0] is not None else [1] * (L + 1), dtype=np.int64)
                p += rh.max(axis=0).astype(np.int64) * mult
            p[0] = p[L] = 1
            q = [1] * (L + 1)
            for k in range(L - 1, 0, -1):
                q[k] = int(min(n[k] * q[k + 1], p[k]))
            sb = [1] * (L + 1)
            for k in range(L - 1):
                cap = caps[k] if caps is not None else 1 << 30
                sb[k + 1] = int(min(sb[k] * n[k], q[k + 1], cap))
            out = BatchedTT(modes, N, [sb[k] * n[k] * sb[k + 1] for k in range(L)], x0.device)
        arr, keep = self._terms(terms)
        caps_arr = (c_int * (L - 1))(*[int(c) for c in caps]) if caps is not None else None
        rc = self.lib.qtt_round_sum_batched(self.h, len(terms), arr, byref(out.c_struct()), caps_arr, float(rel_eps),
                                            c_void_p(status.data_ptr()) if status is not None else c_void_p(), self._stream())
        _lib.check(self.h, rc, "qtt_round_sum_batched")
        return out

    def orthogonalize(self, terms, want_cores=True):
        """Returns (y | None, norm2 tensor [N]) - see qtt_orthogonalize_batched."""
        x0 = terms[0][1]
        L, N = x0.L, x0.N
        modes = self._out_modes(terms)
        norm2 = torch.empty(N, dtype=torch.float64, device=x0.device)
        out = None
        if want_cores:
            ranks_h = [t[1].ranks_host() for t in terms]
            n = [_prod(m) for m in modes]
            p = np.zeros(L + 1, dtype=np.int64)
            for t, rh in zip(terms, ranks_h):
                mult = np.array(t[0].ranks if t[0] is not None else [1] * (L + 1), dtype=np.int64)
                p += rh.max(axis=0).astype(np.int64) * mult
            p[0] = p[L] = 1
            q = [1] * (L + 1)
            for k in range(L - 1, 0, -1):
                q[k] = int(min(n[k] * q[k + 1], p[k]))
            out = BatchedTT(modes, N, [q[k] * n[k] * q[k + 1] for k in range(L)], x0.device)
        arr, keep = self._terms(terms)
        rc = self.lib.qtt_orthogonalize_batched(self.h, len(terms), arr, byref(out.c_struct()) if out is not None else None,
                                                c_void_p(norm2.data_ptr()), self._stream())
        _lib.check(self.h, rc, "qtt_orthogonalize_batched")
        return out, norm2

    def norm2(self, x):
        return self.orthogonalize([(None, x)], want_cores=False)[1]

    def matmul(self, a, x, m, out_modes):
        """Materialised core-wise product; a is a DeviceMPO (shared) or a BatchedTT (per instance)."""
        L, N = x.L, x.N
        shared = isinstance(a, DeviceMPO)
        rx = x.ranks_host().max(axis=0).astype(np.int64)
        ra = np.array(a.ranks, dtype=np.int64) if shared else a.ranks_host().max(axis=0).astype(np.int64)
        n = [_prod(mo) for mo in out_modes]
        slot = [int(ra[k] * rx[k] * n[k] * ra[k + 1] * rx[k + 1]) for k in range(L)]
        y = BatchedTT(out_modes, N, slot, x.device)
        m_arr = (c_int * L)(*[int(v) for v in m])
        rc = self.lib.qtt_matmul_batched(self.h, None if shared else byref(a.c_struct()),
                                         byref(a.c_struct()) if shared else None, byref(x.c_struct()), m_arr,
                                         byref(y.c_struct()), self._stream())
        _lib.check(self.h, rc, "qtt_matmul_batched")
        return y

    def inner(self, a, b, op=None):
        """<a_i, b_i>, or the bilinear form <a_i, op @ b_i> when a DeviceMPO is given."""
        out = torch.empty(a.N, dtype=torch.float64, device=a.device)
        rc = self.lib.qtt_inner_batched(self.h, byref(a.c_struct()), byref(op.c_struct()) if op is not None else None,
                                        byref(b.c_struct()), c_void_p(out.data_ptr()), self._stream())
        _lib.check(self.h, rc, "qtt_inner_batched")
        return out

    def axpby(self, alpha, a, beta, b):
        L, N = a.L, a.N
        ra = a.ranks_host().max(axis=0).astype(np.int64)
        rb = b.ranks_host().max(axis=0).astype(np.int64)
        r = ra + rb
        r[0] = r[L] = 1
        y = BatchedTT(a.mode_shapes, N, [int(r[k] * a.n[k] * r[k + 1]) for k in range(L)], a.device)
        rc = self.lib.qtt_axpby_batched(self.h, c_void_p(alpha.data_ptr()) if alpha is not None else c_void_p(), byref(a.c_struct()),
                                        c_void_p(beta.data_ptr()) if beta is not None else c_void_p(), byref(b.c_struct()),
                                        byref(y.c_struct()), self._stream())
        _lib.check(self.h, rc, "qtt_axpby_batched")
        return y


    def gd_solve(self, A, b, n_steps=200, lr=0.8, max_rank=20, recalc_every=10, rel_accuracy=0.0, want_history=False):
        """Batched truncated steepest descent (solver.py:251-275) - see qtt_gd_solve_batched.
        Returns (x BatchedTT, r2 tensor [N], steps_done tensor [N], history tensor [n_steps, N, 2] | None)."""
        x = self._solution_batch(b, max_rank)
        N = b.N
        r2 = torch.empty(N, dtype=torch.float64, device=b.device)
        steps = torch.empty(N, dtype=torch.int32, device=b.device)
        hist = torch.zeros((max(n_steps, 1), N, 2), dtype=torch.float64, device=b.device) if want_history else None
        rc = self.lib.qtt_gd_solve_batched(self.h, byref(A.c_struct()), byref(b.c_struct()), byref(x.c_struct()),
                                           c_void_p(r2.data_ptr()), int(n_steps), float(lr), int(max_rank), int(recalc_every),
                                           float(rel_accuracy), c_void_p(hist.data_ptr()) if hist is not None else c_void_p(),
                                           c_void_p(steps.data_ptr()), self._stream())
        _lib.check(self.h, rc, "qtt_gd_solve_batched")
        return x, r2, steps, hist


    def cg_solve(self, A, b, n_steps=100, max_rank=20, recalc_every=10, rel_accuracy=0.0, want_history=False):
        """Batched truncated conjugate gradients for an SPD operator - see qtt_cg_solve_batched (no reference counterpart:
        the reference hands this step to ttpy's AMEn).  Returns like gd_solve; history rows are (||r||^2, alpha)."""
        x = self._solution_batch(b, max_rank)
        N = b.N
        r2 = torch.empty(N, dtype=torch.float64, device=b.device)
        steps = torch.empty(N, dtype=torch.int32, device=b.device)
        hist = torch.zeros((max(n_steps, 1), N, 2), dtype=torch.float64, device=b.device) if want_history else None
        rc = self.lib.qtt_cg_solve_batched(self.h, byref(A.c_struct()), byref(b.c_struct()), byref(x.c_struct()),
                                           c_void_p(r2.data_ptr()), int(n_steps), int(max_rank), int(recalc_every),
                                           float(rel_accuracy), c_void_p(hist.data_ptr()) if hist is not None else c_void_p(),
                                           c_void_p(steps.data_ptr()), self._stream())
        _lib.check(self.h, rc, "qtt_cg_solve_batched")
        return x, r2, steps, hist

    @staticmethod
    def _solution_batch(b, max_rank):
        L, N = b.L, b.N
        rb = [1] * (L + 1)
        left = 1
        for k in range(1, L):
            left = min(left * b.n[k - 1], max_rank)
            rb[k] = left
        right = 1
        for k in range(L - 1, 0, -1):
            right = min(right * b.n[k], max_rank)
            rb[k] = min(rb[k], right)
        return BatchedTT(b.mode_shapes, N, [rb[k] * b.n[k] * rb[k + 1] for k in range(L)], b.device)


_default_engines = {}


def default_engine(device=None):
    if not torch.cuda.is_available():
        raise _lib.QttError("CUDA device required: the TT hot path has no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    key = dev.index or 0
    if key not in _default_engines:
        _default_engines[key] = Engine(dev)
    return _default_engines[key]
