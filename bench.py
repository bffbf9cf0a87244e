#!/usr/bin/env python
"""Benchmark of the north-star metric: batched 2D QTT Laplace solves per second (BASELINE.json, config C5).

Workload: per GPU `--instances` independent right-hand sides (13 cores
(1,4,2),(2,4,2)..,(2,4,1), entries default_rng(1234+i).standard_normal / sqrt(8)), 2D, L = 12 levels,
shared BPX-preconditioned stiffness operator B (ranks [16,161x8,121,16]).  One *solve* = RHS assembly
b = round(R f) + 100 truncated steepest-descent steps at rank cap 4 (111 fused apply+round sweeps of B) +
back-transform u = round(C v) + energy and L2 norms.  One *step* = one batched solve of the whole per-GPU batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--instances I] [--impl reference]

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
# CPU reference arm / cpu_baseline: the numpy oracle (a port - the Python reference cannot travel to the box)
# ----------------------------------------------------------------------------------------------------

def cpu_sample(ops, n_gd_steps, instance=0):
    """One bounded sample of the workload on the host cores: the full pipeline of ONE instance with
    n_gd_steps of the 100 descent steps.  Returns (seconds, heavy sweeps done, r2)."""
    from oracle import tt_oracle as O
    L = L_LEVELS
    t0 = time.perf_counter()
    f = c5_rhs_cores(instance)
    b = O.reapprox(O.squeeze(O.matmul(ops["R"], f)), ranks_new=MAX_RANK, rel_error=EPS)
    v, r2 = O.solve_with_grad_descent(ops["B"], b, n_steps_max=n_gd_steps, max_rank=MAX_RANK, rel_accuracy=0.0)
    u = O.reapprox(O.matmul(ops["C"], v), ranks_new=MAX_RANK, rel_error=EPS)
    v_ext = v + [np.ones((1, 1, 1))]
    energy = 0.25 ** L * (O.norm_squared(O.matmul(ops["GyQx"], v_ext)) + O.norm_squared(O.matmul(ops["GxQy"], v_ext)))
    l2 = 4.0 ** L * O.norm_squared(O.matmul(ops["GM2"], u + [np.ones((1, 1, 1))]))
    dt = time.perf_counter() - t0
    heavy = (n_gd_steps + 9) // 10 + n_gd_steps + 1
    return dt, heavy, (energy, l2, r2)


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ops = operators_cached(L_LEVELS, EPS, 0, 1, on_device=False)
    n_gd = args.ref_gd_steps
    full_heavy = GD_STEPS // 10 + GD_STEPS + 1
    for _ in range(args.warmup):
        cpu_sample(ops, n_gd)
    times = []
    for s in range(args.steps):
        dt, heavy, _ = cpu_sample(ops, n_gd, instance=s)
        times.append(dt)
    t_sample = float(np.mean(times))
    t_solve = t_sample * full_heavy / heavy           # extrapolated by the count of apply+round sweeps of B
    value = 1.0 / t_solve
    sample = (f"1 instance x {n_gd} of {GD_STEPS} descent steps ({heavy} of {full_heavy} apply+round sweeps of B) incl. RHS assembly and "
              f"post-processing; {t_sample:.2f} s per sample, extrapolated linearly in the sweep count")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_sample, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5: 2D L=12 BPX solves, rank cap 4, 100 GD steps, numpy oracle port of laplace_mps on host cores",
                       "L": L_LEVELS, "max_rank": MAX_RANK, "gd_steps": GD_STEPS},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
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

    n_total = args.instances * world                  # weak scaling: fixed work per GPU
    lo, hi = shard_range(n_total, rank, world)
    host_ops = operators_cached(L_LEVELS, EPS, rank, world)
    solver = Solver2DBatched(L_LEVELS, max_rank=MAX_RANK, n_steps=GD_STEPS, eps=EPS, device=dev, host_ops=host_ops)
    eng = solver.engine
    instances = [c5_rhs_cores(i) for i in range(lo, hi)]
    packed = BatchedTT.pack_pinned(instances)
    f_dev = BatchedTT.from_pinned(packed, dev)          # inputs resident in HBM for the device-timed leg
    torch.cuda.synchronize()
    h2d_bytes = sum(c.numel() * 8 for c in packed[2]) + packed[3].numel() * 4

    def step_device():
        res = solver.solve(f_dev)
        rec = torch.stack([res["energy"], res["l2"], res["r2"]], dim=1)
        return gather_results(rec, n_total), res

    def step_e2e():
        f = BatchedTT.from_pinned(packed, dev)
        res = solver.solve(f)
        rec = gather_results(torch.stack([res["energy"], res["l2"], res["r2"]], dim=1), n_total)
        rec_h = rec.cpu()
        u_h = [c.cpu() for c in res["u"].cores]
        rk_h = res["u"].ranks.cpu()
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
        rec, res = step_device()
    e1.record()
    barrier()
    gemm_ms, gemm_flops, gemm_launches = eng.profile(reset=True)
    eng.set_profiling(False)
    _, launches = eng.counters()
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
    d2h_bytes = rec_h.numel() * 8 + sum(c.numel() * 8 for c in u_h) + rk_h.numel() * 4

    if rank == 0:
        value = n_total * args.steps / (ms_total * 1e-3)
        e2e_value = n_total * args.e2e_steps / float(t_e2e.item())
        peak, peak_how = measure_fp64_peak()
        achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
        finite = bool(torch.isfinite(rec).all())
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C5: 2D L=12 BPX-preconditioned solves, rank cap 4, 100 truncated steepest-descent steps "
                                   "(111 fused apply+round sweeps of B, ranks [16,161x8,121,16]) + RHS assembly + back-transform + norms",
                       "L": L_LEVELS, "max_rank": MAX_RANK, "gd_steps": GD_STEPS, "instances_per_gpu": args.instances,
                       "instances_total": n_total, "eps": EPS,
                       "l2_policy": "per-step working set (reflector store ~70 MB per instance) is far larger than the 126 MB L2"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": 3.2045e9 * args.instances / 148.0,
                         "traffic_note": "dram read+write bytes of one full-core push launch of gemm_dmma_kernel<64,56,16,56>: 3.2045e9 measured at 148 instances "
                                         "(ncu --set full, profiles/r1_ncu_summary.md; algorithmic bytes 3.9e9 = T in, M out), scaled linearly to this batch size",
                         "kernel": "gemm_dmma_kernel + larfb_kernel + gram_kernel (all FP64 DMMA: mma.sync.m8n8k4.f64)", "launches": gemm_launches,
                         "kernel_share_of_step": gemm_ms / ms_total if ms_total else None,
                         "peak_source": peak_how + "; MEASURED_PEAKS.json carries no FP64 figure"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "steps": args.e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "results_finite": finite,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_gd = args.ref_gd_steps
            dt, heavy, _ = cpu_sample(host_ops, n_gd)
            full_heavy = GD_STEPS // 10 + GD_STEPS + 1
            line["cpu_baseline"] = {"value": 1.0 / (dt * full_heavy / heavy), "unit": "solves/s", "cores": cpu_threads(), "kind": "port",
                                    "sample": f"numpy oracle, 1 instance x {n_gd} of {GD_STEPS} descent steps ({heavy} of {full_heavy} sweeps of B), "
                                              f"{dt:.2f} s, extrapolated linearly in the sweep count"}
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
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-gd-steps", type=int, default=20, help="descent steps of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
