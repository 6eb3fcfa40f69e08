#!/usr/bin/env python
"""bench.py -- dispatch LPs/sec on BASELINE.json's headline config (C2: wind+battery, 24 periods, 10 000
synthetic LMP scenarios per GPU), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, weak scaling: 10 000 LPs per rank,
                                                            one all_gather of the objectives per step)

A "step" = one pass of the hot path over one batch: parameter rows already in HBM -> one kernel launch ->
obj/status/iters in HBM (`value`), or through the host C-ABI call with H2D / D2H inside the timed region (`e2e`).
`--impl reference` times the CPU path the reference would take for these LPs (restated LP + HiGHS in a process
pool over all host cores; CBC/IPOPT/Pyomo are not installable here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T = 24
BATCH = 10000
METRIC = "dispatch LPs/sec (24-period price-taker batch); max |obj| rel-err vs IPOPT"
ALG_BYTES_PER_LP = 8 * T + 8 + 4 + 4           # lmp row in; obj, status, iters out (208 B)
ALG_FLOP_PER_LP = 0.30e6                       # banded-IPM algorithmic FP64 flops per LP
FP64_PEAK_TFLOPS_NOMINAL = 37.0                # HGX B200 spec sheet (296 TF / 8 GPUs); no measured FP64 peak file


def workload(rank):
    from dispatches_b200 import scenarios as SC, templates as TP
    lmp, cf, W, P = SC.c2(BATCH, seed=20240101 + rank)
    rp = TP.wind_battery_rparams(T, cf, W, P)[0]
    return lmp, cf, W, P, rp


def config(n_gpus):
    return {"workload": f"C2: renewables wind+battery 24-period price-taker, {BATCH} synthetic LMP scenarios per GPU "
                        f"(seed 20240101+rank), fixed design 847 MW wind / 211.75 MW 4-h battery",
            "T": T, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"scenario-shard x{n_gpus}",
            "l2": "flushed between timed steps (256 MiB write)", "template": "wind_battery_T24 (m=96, n=167, w=4)", "kernel": "dsp_ipm_stage_wb_kernel (warp per LP, lane per period, registers only)"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_cores():
    from oracle import highs as H
    return H.host_cores()


def cpu_reference_run(lmp, cf, W, P, n_sample, procs=None):
    """The oracle loop (HiGHS dual simplex, constraints pre-built, cost vector swapped per LP) on all host cores."""
    from oracle import highs as H
    procs = procs or host_cores()
    obj, dt, procs = H.solve_batch("wind_battery", lmp[:n_sample], kwargs=dict(cf=cf, wind_mw=W, batt_mw=P), procs=procs)
    return obj, dt, procs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lmp, cf, W, P, rp = workload(0)
    cores = host_cores()
    per_step = int(min(BATCH, max(200, 250 * cores)))          # ~1 s of wall clock per step on all cores
    for _ in range(args.warmup):
        cpu_reference_run(lmp, cf, W, P, min(per_step, 64 * cores))
    t_tot = 0.0
    for k in range(args.steps):
        lo = (k * per_step) % max(1, BATCH - per_step + 1)
        _, dt, procs = cpu_reference_run(lmp[lo:], cf, W, P, per_step)
        t_tot += dt
    value = per_step * args.steps / t_tot
    sample = f"{per_step} LPs of the C2 batch per step, HiGHS dual simplex, {procs} processes"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "LPs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "LPs/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "LPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    import torch
    import torch.distributed as dist
    from dispatches_b200 import solver as S, templates as TP
    from dispatches_b200.csrc import build
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lmp, cf, W, P, rp = workload(rank)
    t = TP.wind_battery(T)
    sol = S.BatchLPSolver(t)
    cp_d = torch.tensor(lmp, device=dev)
    rp_d = torch.tensor(rp, device=dev)
    out = sol.solve(cp_d, rp_d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty(BATCH * world, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        sol.solve(cp_d, rp_d, out=out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.obj)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    n0 = S.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                 # L2 flush, outside the per-step event pair
        ev[k][0].record()
        sol.solve(cp_d, rp_d, out=out)
        ev[k][2].record()                     # kernel-only stop (before the collective)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.obj)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = S.launch_count() - n0
    step_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    kern_ms = sum(a.elapsed_time(c) for a, _, c in ev)
    tt = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = float(tt[0]), float(tt[1])
    # ---- e2e: the host C-ABI call (pinned staging, H2D, kernel, D2H, sync), every step
    # host buffers are page-locked (the contract's "pinned host memory"): inputs are DMA'd straight from them
    lmp_pin = S.BatchLPSolver.pinned_empty(lmp.shape); lmp_pin[:] = lmp
    rp_pin = S.BatchLPSolver.pinned_empty(rp.shape); rp_pin[:] = rp
    r_host = S.LPResult(S.BatchLPSolver.pinned_empty(BATCH), S.BatchLPSolver.pinned_empty(BATCH, np.int32),
                        S.BatchLPSolver.pinned_empty(BATCH, np.int32))
    for _ in range(2):
        sol.solve_host(lmp_pin, rp_pin, out=r_host)
    barrier()
    e2e_wall = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()                     # the L2 flush stays outside the timed call
        t0 = time.perf_counter()
        sol.solve_host(lmp_pin, rp_pin, out=r_host)
        e2e_wall += time.perf_counter() - t0
    tt = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_wall = float(tt[0])
    clocks = sampler.stop() if rank == 0 else None
    status = out.status.cpu().numpy()
    iters = out.iters.cpu().numpy()
    stats = torch.tensor([float((status != 0).sum()), float(iters.sum()), float(iters.max())], dtype=torch.float64, device=dev)
    if world > 1:
        s2 = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(s2, op=dist.ReduceOp.MAX)
        stats[2] = s2[2]
    stats = stats.cpu()
    if world > 1:                              # all collectives are done: every rank leaves the group together
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return
    total = BATCH * world
    value = total * args.steps / (step_ms * 1e-3)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.load(open(pk))
    hbm_peak, which = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    kern_s = kern_ms * 1e-3 / args.steps
    ach = ALG_BYTES_PER_LP * BATCH / kern_s / 1e9
    fp64 = ALG_FLOP_PER_LP * BATCH / kern_s / 1e12
    fp64_peak = S.fp64_peak_tflops()
    if not fp64_peak > 0:
        fp64_peak = None
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), not measured live
    traffic = None
    try:
        cur = json.load(open(ROOT / "profiles" / "current.json"))["stage_profile"]
        traffic = float(json.load(open(ROOT / "profiles" / cur))["traffic_bytes_per_launch"])
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "LPs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config(world),
            "e2e": {"value": total * args.steps / e2e_wall, "unit": "LPs/s", "h2d_bytes_per_step": int(lmp.nbytes + rp.nbytes),
                    "d2h_bytes_per_step": int(BATCH * 16), "ms_per_step": 1e3 * e2e_wall / args.steps,
                    "api": "dsp_lp_solve_batch_host (C-ABI, host buffers)"},
            "gpu_launches": int(launches), "kernel_ms_per_step": kern_ms / args.steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": traffic, "alg_bytes_per_launch": ALG_BYTES_PER_LP * BATCH, "peak_source": which,
                         "note": "on-chip FP64 solve: HBM is not the binding resource (SURVEY.md §8d); see fp64",
                         "fp64": {"achieved_tflops": fp64, "peak_tflops": fp64_peak or FP64_PEAK_TFLOPS_NOMINAL,
                                  "frac": fp64 / (fp64_peak or FP64_PEAK_TFLOPS_NOMINAL),
                                  "peak_source": "measured DFMA micro-benchmark (dsp_lp_fp64_peak_tflops)" if fp64_peak else "nominal (HGX B200 spec)",
                                  "peak_tflops_nominal": FP64_PEAK_TFLOPS_NOMINAL,
                                  "alg_flop_per_lp": ALG_FLOP_PER_LP}},
            "solver": {"non_optimal": int(stats[0]), "iters_mean": float(stats[1]) / total, "iters_max": int(stats[2]),
                       "launch": S.last_launch()},
            "clocks": clocks, "wall_s_timed_loop": t_wall}
    if not args.no_cpu_baseline:
        # CPU baseline in a fresh interpreter (no fork of this CUDA process; nothing runs before the ranks rendezvous)
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            outp = os.path.join(td, "cpu.npz")
            code = ("import sys; sys.path.insert(0, %r); import numpy as np, bench; "
                    "lmp, cf, W, P, _ = bench.workload(0); c = bench.host_cores(); n = int(min(bench.BATCH, max(500, 400 * c))); "
                    "bench.cpu_reference_run(lmp, cf, W, P, min(n, 16 * c)); ref, dt, procs = bench.cpu_reference_run(lmp, cf, W, P, n); "
                    "np.savez(%r, ref=ref, dt=dt, procs=procs, n=n)") % (str(ROOT), outp)
            rc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
            if rc.returncode == 0:
                z = np.load(outp)
                cpu = dict(ref=z["ref"], dt=float(z["dt"]), procs=int(z["procs"]), n=int(z["n"]))
            else:
                line["cpu_baseline"] = {"error": rc.stderr[-300:]}
    if cpu is not None:
        ref, dt, procs, n_sample = cpu["ref"], cpu["dt"], cpu["procs"], cpu["n"]
        err = np.abs(r_host.obj[:n_sample] - ref) / np.maximum(1.0, np.abs(ref))
        line["cpu_baseline"] = {"value": n_sample / dt, "unit": "LPs/s", "cores": procs, "kind": "port",
                                "sample": f"first {n_sample} LPs of rank 0's batch, restated LP + HiGHS dual simplex, {procs} processes"}
        line["max_rel_err_vs_oracle"] = float(err.max())
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
