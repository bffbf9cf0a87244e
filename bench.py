#!/usr/bin/env python
"""Benchmark of the north-star metric: batched 2D QTT Laplace solves per second (BASELINE.json, config C5).

Workload: per GPU `--instances` independent right-hand sides (13 cores
(1,4,2),(2,4,2)..,(2,4,1), entries default_rng(1234+i).standard_normal / sqrt(8)), 2D, L = 12 levels,
shared BPX-preconditioned stiffness operator B (ranks [16,161x8,121,16]).  One *solve* = RHS assembly
b = round(R f) + 100 truncated steepest-descent steps at rank cap 4 (111 fused apply+round sweeps of B) +
back-transform u = round(C v) + energy and L2 norms.  One *step* = one batched solve of the whole per-GPU batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--instances I] [--instances-total T] [--impl reference]

Default: weak scaling, `--instances` per GPU.  `--instances-total 4096` is the literal C5 job (strong scaling: the fixed
total is sharded over the ranks and each rank walks its shard in chunks of `--instances`).

N > 1 is launched by torchrun (one rank per GPU); instances are sharded, results gathered once per step
(NCCL all_gather of an [I, 3] tensor); `value` is the whole-job aggregate.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_LEVELS = 12
MAX_RANK = 4
GD_STEPS = 100
EPS = 1e-12
METRIC = "QTT Laplace solves/sec (2D, L=12, batched)"
FP64_PEAK_FALLBACK_TFLOPS = 37.1      # tools/fp64_peak.cu on this pool's B200 (DMMA pipe), gpurun_out/fp64_peak.txt


def c5_rhs_cores(i, L=L_LEVELS):
    rng = np.random.default_rng(1234 + i)
    rk = [1] + [2] * L + [1]
    return [rng.standard_normal((rk[k], 4, rk[k + 1])) / np.sqrt(8.0) for k in range(L + 1)]


def operators_cached(L, eps, rank, world, on_device=True):
    """Operator assembly; rank 0 builds and caches to disk, the others load.  GPU arm: the rank-1152 system operator is
    rounded on the device; CPU reference arm: everything on the host."""
    import torch.distributed as dist
    from laplacesolvermps_b200.pipeline import build_operators_2D
    tag = "dev" if on_device else "host"
    cache = os.path.join(os.environ.get("QTT_CACHE_DIR", "/tmp"), f"qtt_ops2d_L{L}_eps{eps:g}_{tag}.npz")

    def load():
        z = np.load(cache)
        names = sorted(set(k.rsplit("_", 1)[0] for k in z.files))
        return {n: [z[f"{n}_{k}"] for k in range(len([f for f in z.files if f.startswith(n + "_")]))] for n in names}

    if rank == 0 and not os.path.exists(cache):
        ops = build_operators_2D(L, eps, on_device=on_device)
        tmp = cache + f".tmp{os.getpid()}.npz"
        np.savez(tmp, **{f"{n}_{k}": c for n, cores in ops.items() for k, c in enumerate(cores)})
        os.replace(tmp, cache)
    if world > 1:
        dist.barrier()
    return load()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak():
    """FP64 DMMA peak on this GPU from the in-tree microbenchmark (MEASURED_PEAKS.json has no FP64 entry)."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        vals = [float(l.split(":")[1].split("TFLOP/s")[0]) for l in out.splitlines() if l.startswith("DMMA")]
        if vals:
            return max(vals), "measured live by tools/fp64_peak (DMMA m8n8k4 f64)"
    except Exception:
        pass
    return FP64_PEAK_FALLBACK_TFLOPS, "fallback: tools/fp64_peak measured earlier on this pool (37.1)"


# ----------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline.  Preferred: the UNMODIFIED reference package (oracle/_ref/reference_src.zip, made by
# oracle/build_ref.py from /root/reference; `kind: "reference"`), driven through its own public functions exactly as
# solve_PDE_2D_with_preconditioner does (solver.py:212-233) with solve_with_grad_descent (solver.py:251-275) in place of the
# absent ttpy.  Fallback when the archive is missing: the numpy oracle port (`kind: "port"`).
# ----------------------------------------------------------------------------------------------------

def workload_config(instances_per_gpu, n_total, strong=False):
    """The ONE description of the workload both arms print (the driver compares the two `config` objects)."""
    return {"workload": f"C5: 2D L=12 BPX-preconditioned solves, rank cap {MAX_RANK}, 100 truncated steepest-descent steps "
                        "(111 fused apply+round sweeps of B, ranks [16,161x8,121,16]) + RHS assembly + back-transform + norms",
            "L": L_LEVELS, "max_rank": MAX_RANK, "gd_steps": GD_STEPS, "instances_per_gpu": instances_per_gpu,
            "instances_total": n_total, "eps": EPS, "sharding": "strong (fixed total)" if strong else "weak (fixed per GPU)",
            "l2_policy": "per-step working set (reflector store ~70 MB per instance) is far larger than the 126 MB L2"}


def cpu_threads_all():
    """Make BLAS use every host core (torchrun exports OMP_NUM_THREADS=1 to its children) and return the count in use."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=n)                 # process-wide, stays in force
        got = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
        return int(got)
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


class CpuArm:
    """The reference's CPU implementation of one C5 solve, on a bounded number of descent steps."""

    def __init__(self):
        from oracle import ref_loader
        self.kind = "port"
        self.ref = None
        if ref_loader.reference_available():
            try:
                self.ref = ref_loader.load_reference()
                self.kind = "reference"
            except Exception as e:      # noqa: BLE001 - fall back to the port, say why
                print(f"reference import failed ({e}); using the numpy port", file=sys.stderr)
        self.ops = None

    def build_operators(self):
        L = L_LEVELS
        if self.kind == "reference":
            tm, solver, bpx2D, utils = self.ref
            cache = os.path.join(os.environ.get("QTT_CACHE_DIR", "/tmp"), f"qtt_refB_L{L}_eps{EPS:g}.npz")
            if os.path.exists(cache):
                z = np.load(cache)
                B = tm.TensorTrain([z[f"B_{k}"] for k in range(L)])
            else:
                B = bpx2D.get_laplace_BPX_2D(L)
                B.reapprox(rel_error=EPS)                       # solver.py:218-220 (one-off per L: ~20 s)
                tmp = cache + f".tmp{os.getpid()}.npz"
                np.savez(tmp, **{f"B_{k}": c for k, c in enumerate(B.tensors)})
                os.replace(tmp, cache)
            G = utils._get_gram_matrix_tt(L)
            G.tensors[-1] = np.sqrt(G.tensors[-1] / 2)
            I = utils._get_identy_as_tt(L, True)
            self.ops = dict(B=B, R=bpx2D.get_rhs_matrix_BPX_2D(L), C=bpx2D.get_BPX_preconditioner_2D(L),
                            Qx=bpx2D.get_BPX_Qp_2D(L, "x"), Qy=bpx2D.get_BPX_Qp_2D(L, "y"),
                            Gx=utils.kronecker_prod_2D(G, I), Gy=utils.kronecker_prod_2D(I, G))
        else:
            self.ops = operators_cached(L_LEVELS, EPS, 0, 1, on_device=False)

    def sample(self, n_gd_steps, instance=0):
        """Full pipeline of ONE instance with n_gd_steps of the 100 descent steps.  Returns (seconds, sweeps of B done)."""
        import contextlib, io, warnings
        L = L_LEVELS
        ops = self.ops
        t0 = time.perf_counter()
        if self.kind == "reference":
            tm, solver, bpx2D, utils = self.ref
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                f = tm.TensorTrain(c5_rhs_cores(instance))
                b = (ops["R"] @ f).squeeze().reapprox(ranks_new=MAX_RANK, rel_error=EPS)
                v, r2 = solver.solve_with_grad_descent(ops["B"], b, n_steps_max=n_gd_steps, max_rank=MAX_RANK, rel_accuracy=0.0)
                u = (ops["C"] @ v).reapprox(ranks_new=MAX_RANK, rel_error=EPS)
                v.tensors.append(np.ones([1, 1, 1, 1]))
                Dx = (ops["Qx"] @ v).flatten_mode_indices()
                Dy = (ops["Qy"] @ v).flatten_mode_indices()
                energy = 0.25 ** L * ((ops["Gy"] @ Dx).norm_squared() + (ops["Gx"] @ Dy).norm_squared())
                l2 = solver.get_L2_norm_2D(u)
        else:
            from oracle import tt_oracle as O
            f = c5_rhs_cores(instance)
            b = O.reapprox(O.squeeze(O.matmul(ops["R"], f)), ranks_new=MAX_RANK, rel_error=EPS)
            v, r2 = O.solve_with_grad_descent(ops["B"], b, n_steps_max=n_gd_steps, max_rank=MAX_RANK, rel_accuracy=0.0)
            u = O.reapprox(O.matmul(ops["C"], v), ranks_new=MAX_RANK, rel_error=EPS)
            v_ext = v + [np.ones((1, 1, 1))]
            energy = 0.25 ** L * (O.norm_squared(O.matmul(ops["GyQx"], v_ext)) + O.norm_squared(O.matmul(ops["GxQy"], v_ext)))
            l2 = 4.0 ** L * O.norm_squared(O.matmul(ops["GM2"], u + [np.ones((1, 1, 1))]))
        dt = time.perf_counter() - t0
        assert np.isfinite(energy) and np.isfinite(l2) and np.isfinite(r2)
        heavy = (n_gd_steps + 9) // 10 + n_gd_steps + 1
        return dt, heavy

    def describe(self):
        return ("unmodified reference package (laplace_mps from oracle/_ref/reference_src.zip) through its own functions"
                if self.kind == "reference" else "numpy oracle port of laplace_mps (oracle/tt_oracle.py)")


FULL_HEAVY = GD_STEPS // 10 + GD_STEPS + 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads_all()
    arm = CpuArm()
    arm.build_operators()
    n_gd = args.ref_gd_steps
    for _ in range(args.warmup):
        arm.sample(min(n_gd, 3))                      # warm-up: BLAS threads, page cache; a few descent steps are enough
    times = []
    for s in range(args.steps):
        dt, heavy = arm.sample(n_gd, instance=s)
        times.append(dt)
    t_sample = float(np.mean(times))
    t_solve = t_sample * FULL_HEAVY / heavy           # extrapolated by the count of apply+round sweeps of B
    value = 1.0 / t_solve
    sample = (f"{arm.describe()}; 1 instance x {n_gd} of {GD_STEPS} descent steps ({heavy} of {FULL_HEAVY} apply+round sweeps of B) incl. RHS "
              f"assembly and post-processing; {t_sample:.2f} s per sample"
              + ("" if heavy == FULL_HEAVY else ", extrapolated linearly in the sweep count") + f"; {cores} BLAS threads")
    n_total = args.instances * args.gpus if not args.instances_total else args.instances_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_sample, "higher_is_better": True,
            "scaling": "strong" if args.instances_total else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.instances, n_total, strong=bool(args.instances_total)),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from laplacesolvermps_b200.batched import BatchedTT
    from laplacesolvermps_b200.parallel import shard_range, gather_results
    from laplacesolvermps_b200.pipeline import Solver2DBatched

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    strong = bool(args.instances_total)                # the literal C5: a FIXED total (4096) sharded over the ranks
    n_total = args.instances_total if strong else args.instances * world
    lo, hi = shard_range(n_total, rank, world)
    host_ops = operators_cached(L_LEVELS, EPS, rank, world)
    solver = Solver2DBatched(L_LEVELS, max_rank=MAX_RANK, n_steps=GD_STEPS, eps=EPS, device=dev, host_ops=host_ops)
    eng = solver.engine
    # a rank's shard is processed in chunks of at most --instances (one workspace chunk, two waves of the QR leaves)
    n_chunks = max(1, -(-(hi - lo) // args.instances))
    per_chunk = -(-(hi - lo) // n_chunks)                  # balanced chunks (512 per rank -> 2 x 256, not 296 + 216)
    chunks = [(c0, min(c0 + per_chunk, hi)) for c0 in range(lo, hi, per_chunk)]
    packed = [BatchedTT.pack_pinned([c5_rhs_cores(i) for i in range(c0, c1)]) for c0, c1 in chunks]
    f_dev = [BatchedTT.from_pinned(pk, dev) for pk in packed]      # inputs resident in HBM for the device-timed leg
    torch.cuda.synchronize()
    h2d_bytes = sum(sum(c.numel() * 8 for c in pk[2]) + pk[3].numel() * 4 for pk in packed)

    def record(res):
        return torch.stack([res["energy"], res["l2"], res["r2"]], dim=1)

    def step_device():
        recs = [record(solver.solve(f, want_u=True)) for f in f_dev]
        return gather_results(torch.cat(recs, dim=0), n_total)

    def step_e2e():
        recs, u_h, rk_h = [], [], []
        for pk in packed:
            f = BatchedTT.from_pinned(pk, dev)
            res = solver.solve(f)
            recs.append(record(res))
            u_h += [c.cpu() for c in res["u"].cores]
            rk_h.append(res["u"].ranks.cpu())
        rec_h = gather_results(torch.cat(recs, dim=0), n_total).cpu()
        return rec_h, u_h, rk_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.reset_counters()
    eng.set_profiling(True)
    eng.profile(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec = step_device()
    e1.record()
    barrier()
    gemm_ms, gemm_flops, gemm_launches = eng.profile(reset=True)
    eng.set_profiling(False)
    _, launches = eng.counters()
    dmma_flops, leaf_flops = eng.flop_counters()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # end-to-end leg: host buffers in, host results out, copies inside the timed region
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        rec_h, u_h, rk_h = step_e2e()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    d2h_bytes = rec_h.numel() * 8 + sum(c.numel() * 8 for c in u_h) + sum(r.numel() * 4 for r in rk_h)

    if rank == 0:
        value = n_total * args.steps / (ms_total * 1e-3)
        e2e_value = n_total * args.e2e_steps / float(t_e2e.item())
        peak, peak_how = measure_fp64_peak()
        achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
        # whole step, rank 0's shard: every executed FP64 flop the library counts (DMMA-class kernels + panel QR leaves;
        # skipped structural-zero chunks, SVDs and the bandwidth-bound glue kernels are NOT counted) over the wall time of the step
        step_tflops = (dmma_flops + leaf_flops) / (ms_total * 1e-3) * 1e-12
        finite = bool(torch.isfinite(rec).all())
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.instances, n_total, strong=strong),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "frac_step": step_tflops / peak if peak else None, "achieved_step": step_tflops,
                         "frac_note": "frac: DMMA-class kernels only, flops / their own event-timed duration.  frac_step: all executed flops the "
                                      "library counts (DMMA-class + QR leaves) / the whole step's duration incl. every latency-bound kernel and host gap",
                         "executed_flops_per_solve": (dmma_flops + leaf_flops) / max(args.steps * (hi - lo), 1),
                         "leaf_share_of_flops": leaf_flops / max(dmma_flops + leaf_flops, 1.0),
                         "traffic": None,
                         "traffic_note": "per-launch DRAM bytes of the current kernels are in profiles/r2_ncu_summary.md (ncu --set full); not repeated "
                                         "here as a constant",
                         "kernel": "gemm_push_kernel + gemm_dmma_kernel + larfb2_kernel + larfb_kernel (all FP64 DMMA: mma.sync.m8n8k4.f64)", "launches": gemm_launches,
                         "kernel_share_of_step": gemm_ms / ms_total if ms_total else None,
                         "peak_source": peak_how + "; MEASURED_PEAKS.json carries no FP64 figure"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "steps": args.e2e_steps},
            "gpu_launches": int(launches),
            "launches_per_step": int(launches) / max(args.steps, 1),
            "clocks": clocks,
            "results_finite": finite,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = cpu_threads_all()
            arm = CpuArm()
            arm.build_operators()
            n_gd = args.ref_gd_steps
            arm.sample(2)
            dt, heavy = arm.sample(n_gd)
            line["cpu_baseline"] = {"value": 1.0 / (dt * FULL_HEAVY / heavy), "unit": "solves/s", "cores": cores, "kind": arm.kind,
                                    "sample": f"{arm.describe()}; 1 instance x {n_gd} of {GD_STEPS} descent steps ({heavy} of {FULL_HEAVY} sweeps of B), "
                                              f"{dt:.2f} s" + ("" if heavy == FULL_HEAVY else ", extrapolated linearly in the sweep count")}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# Second line (only with --config c1): BASELINE.json configs[0] as a batch - 1D Poisson, QTT L = 20, BPX, rank cap 20,
# 300 steepest-descent steps (the reference's own CPU experiment), N right-hand sides per GPU.
# ----------------------------------------------------------------------------------------------------
C1_L, C1_CAP, C1_STEPS = 20, 20, 300
C1_METRIC = "QTT Laplace solves/sec (1D, L=20, batched)"


def c1_coeffs(lo, hi):
    """Right-hand sides f_i(x) = c0 + c1 x + c2 x^2 (instance 0 is the config's f = 1)."""
    out = np.zeros((hi - lo, 3))
    for j, i in enumerate(range(lo, hi)):
        out[j] = [1.0, 0.0, 0.0] if i == 0 else np.random.default_rng(4321 + i).standard_normal(3)
    return out


def c1_coeff_list(i):
    """Coefficients of instance i for the CPU arms, trailing zeros dropped: the reference's get_polynomial_as_tt sizes its cores by
    len(coeffs) but numpy trims the converted Legendre series (f = 1 given as [1, 0, 0] raises a shape mismatch in utils.py:203)."""
    co = list(c1_coeffs(i, i + 1)[0])
    while len(co) > 1 and co[-1] == 0.0:
        co.pop()
    return co


def c1_config(n_per_gpu, n_total):
    return {"workload": "C1: 1D L=20 BPX-preconditioned solves, rank cap 20, 300 truncated steepest-descent steps (331 fused apply+round "
                        "sweeps of B, ranks <= 21) + RHS assembly + back-transform + norms", "L": C1_L, "max_rank": C1_CAP,
            "gd_steps": C1_STEPS, "instances_per_gpu": n_per_gpu, "instances_total": n_total,
            "l2_policy": "working set of a batch (reflector stores) exceeds the 126 MB L2 from ~100 instances on"}


def c1_cpu_sample(n_gd, warmup, steps):
    """The reference's CPU implementation of one C1 solve on a bounded number of descent steps: (solves/s extrapolated to the full
    300 steps, seconds per sample, BLAS threads, kind, description)."""
    import contextlib, io, warnings
    from oracle import ref_loader
    cores = cpu_threads_all()
    kind = "reference" if ref_loader.reference_available() else "port"
    if kind == "reference":
        tm, solver, bpx2D, utils = ref_loader.load_reference()
        B = solver.get_laplace_bpx(C1_L).reapprox(rel_error=1e-15)
        R, C, Qp = solver.get_rhs_matrix_bpx(C1_L), solver.get_bpx_preconditioner(C1_L), solver.get_bpx_Qp(C1_L)

        def sample(i):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                f = utils.get_polynomial_as_tt(c1_coeff_list(i), C1_L)
                b = (R @ f).squeeze()
                for k in range(C1_L):
                    b.tensors[k] /= 2
                v, r2 = solver.solve_with_grad_descent(B, b, n_steps_max=n_gd, max_rank=C1_CAP, rel_accuracy=0.0)
                u = (C @ v).reapprox(rel_error=1e-15)
                Du = (Qp @ v).reapprox(rel_error=1e-15)
                l2, h1 = solver.get_L2_norm_1D(u), Du.norm_squared() * 0.5 ** C1_L
            assert np.isfinite(l2) and np.isfinite(h1)
            return time.perf_counter() - t0
    else:
        from oracle import tt_oracle as O
        from laplacesolvermps_b200.pipeline import build_operators_1D
        import laplacesolvermps_b200.utils as utils
        ops = build_operators_1D(C1_L)

        def sample(i):
            t0 = time.perf_counter()
            f = utils.get_polynomial_as_tt(c1_coeff_list(i), C1_L).tensors
            b = O.squeeze(O.matmul(ops["R"], f))
            b[0] = b[0] * 0.5 ** C1_L
            v, r2 = O.solve_with_grad_descent(ops["B"], b, n_steps_max=n_gd, max_rank=C1_CAP, rel_accuracy=0.0)
            O.reapprox(O.matmul(ops["C"], v), rel_error=1e-15)
            return time.perf_counter() - t0
    for _ in range(warmup):
        sample(0)
    times = [sample(s) for s in range(steps)]
    heavy, full = (n_gd + 9) // 10 + n_gd + 1, C1_STEPS // 10 + C1_STEPS + 1
    t_sample = float(np.mean(times))
    value = 1.0 / (t_sample * full / heavy)
    what = ("unmodified reference package (laplace_mps from oracle/_ref/reference_src.zip) through its own functions" if kind == "reference"
            else "numpy port of the reference (oracle/tt_oracle.py)")
    return value, t_sample, cores, kind, (f"{what}; 1 instance x {n_gd} of {C1_STEPS} descent steps, {t_sample:.2f} s per sample, "
                                          f"extrapolated linearly in the sweep count; {cores} BLAS threads")


def run_c1_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    value, t_sample, cores, kind, what = c1_cpu_sample(min(args.ref_gd_steps, C1_STEPS), args.warmup, max(args.steps, 1))
    n_total = args.instances * args.gpus
    emit({"impl": "reference", "metric": C1_METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": 1e3 * t_sample, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f64", "data": "synthetic", "config": c1_config(args.instances, n_total),
          "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": kind, "sample": what},
          "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_c1_gpu(args):
    import torch
    import torch.distributed as dist
    from laplacesolvermps_b200 import encoders
    from laplacesolvermps_b200.parallel import shard_range, gather_results
    from laplacesolvermps_b200.pipeline import Solver1DBatched
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n_total = args.instances * world
    lo, hi = shard_range(n_total, rank, world)
    solver = Solver1DBatched(C1_L, max_rank=C1_CAP, n_steps=C1_STEPS, rel_accuracy=0.0, device=dev)
    eng = solver.engine
    coef_host = torch.from_numpy(c1_coeffs(lo, hi)).pin_memory()
    coef_dev = coef_host.to(dev)

    def step(coef):
        f = encoders.polynomial_batch(coef, C1_L, device=dev)           # right-hand sides encoded on the device from [N, 3] coefficients
        res = solver.solve(f)
        return gather_results(torch.stack([res["l2"], res["h1"], res["r2"]], dim=1), n_total), res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(coef_dev)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        rec, res = step(coef_dev)
    e1.record()
    barrier()
    _, launches = eng.counters()
    dmma_flops, leaf_flops = eng.flop_counters()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        rec_e, res_e = step(coef_host.to(dev, non_blocking=True))
        rec_h = rec_e.cpu()
        u_h = [c.cpu() for c in res_e["u"].cores]
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        value = n_total * args.steps / (ms_total * 1e-3)
        # HBM roof: algorithmic bytes of a solve = every sweep reads its input trains once and writes its output once
        # (SURVEY 8d "minimum bytes for a rounding"): per descent step 4 roundings + 1 norm + 1 inner product on trains of
        # bond <= cap (20 cores of <= 20 x 2 x 20 doubles = 6.4 KB each) and one operator core set per batch.
        core_bytes = 8 * sum(min(C1_CAP, 2 ** min(k, C1_L - k)) * 2 * min(C1_CAP, 2 ** min(k + 1, C1_L - k - 1)) for k in range(C1_L))
        bytes_per_solve = C1_STEPS * (4 * 3 + 1 + 2) * core_bytes
        hbm_peak = 6540.5
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        achieved = bytes_per_solve * value / 1e9
        peak, peak_how = measure_fp64_peak()
        step_tflops = (dmma_flops + leaf_flops) / (ms_total * 1e-3) * 1e-12
        line = {"metric": C1_METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic", "config": c1_config(args.instances, n_total),
              "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                           "note": "cores of a few KB and ~2000 dependent launches per descent step: this config is launch / latency bound, "
                                   "the HBM fraction is reported as the contract asks; the executed-flop rate of the step is given beside it",
                           "fp64_frac_step": step_tflops / peak if peak else None, "fp64_achieved_step": step_tflops,
                           "algorithmic_bytes_per_solve": bytes_per_solve, "peak_source": "MEASURED_PEAKS.json hbm_gbs"},
              "e2e": {"value": n_total * args.e2e_steps / float(t_e2e.item()), "unit": "solves/s", "h2d_bytes_per_step": int(coef_host.numel() * 8),
                      "d2h_bytes_per_step": int(rec_h.numel() * 8 + sum(c.numel() * 8 for c in u_h)), "steps": args.e2e_steps},
              "gpu_launches": int(launches), "launches_per_step": int(launches) / max(args.steps, 1), "clocks": clocks,
              "results_finite": bool(torch.isfinite(rec).all()),
              "l2_norm_of_instance_0_vs_2_15": float(rec[0, 0]) / (2 / 15) - 1}
        if world == 1 and not args.no_cpu_baseline:
            cv, _, ccores, ckind, cwhat = c1_cpu_sample(min(args.ref_gd_steps, C1_STEPS), 0, 1)
            line["cpu_baseline"] = {"value": cv, "unit": "solves/s", "cores": ccores, "kind": ckind, "sample": cwhat}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _route_stdout_to_stderr():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is written to
    the saved descriptor by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _route_stdout_to_stderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--instances", type=int, default=296, help="instances per GPU per step (weak scaling); 296 = two waves of the one-CTA-per-instance QR leaves on 148 SMs, the batch size with the best measured time per instance")
    ap.add_argument("--instances-total", type=int, default=0, help="strong scaling: a FIXED total number of instances (the literal C5 is 4096) "
                    "sharded over the ranks and processed in chunks of --instances; one step = the whole job")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-gd-steps", type=int, default=50, help="descent steps (of 100) of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c5", choices=["c5", "c1"], help="c5 (default): the headline metric; c1: the 1D L=20 config as a batch")
    ap.add_argument("--max-rank", type=int, default=4, help="rank cap of config C5 (4 = the headline; 8 = the 'stress' cap of BASELINE.json: "
                    "5152-row panels through the thread-block-cluster leaf - use --instances 74: two cluster CTAs per instance fill the 148 SMs)")
    args = ap.parse_args()
    global MAX_RANK
    MAX_RANK = args.max_rank
    if args.config == "c1":
        if args.instances == 296:
            args.instances = 512
        (run_c1_reference if args.impl == "reference" else run_c1_gpu)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
