#!/usr/bin/env python
"""bench.py — simulated fish-passages/s of the B200 passage engine (BASELINE.json metric) and of the CPU baseline.

  python bench.py [--gpus N --steps K --warmup W] [--workload c2|c3|c5] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A *step* is one pass of the hot path (count -> scan -> passage kernels) over the workload's cells with a fresh Philox
seed.  Default workload ``c2`` = BASELINE.json configs[1]: synthetic 3-unit Kaplan+Kaplan+Francis facility with bypass
and spill, 1e6 fish/iteration x 1000 iterations (x 1 day) on one B200 = 1e9 fish-passages per step.  With N ranks every
rank simulates its own 1000 iterations (weak scaling; iterations are independent, the Philox counter is keyed by the
global iteration id) and the tally arrays are summed with ONE NCCL all-reduce per step.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# Algorithmic lane-ops per fish-passage = instructions of the best known schedule of the algorithm (DESIGN.md §4
# "Roofline"): what the specialised kernel executes per fish (ncu, profiles/r01_inst_per_fish.json).  SURVEY.md §8d
# estimated 500 (C2) / 1900 (C5) assuming 4 Philox blocks and atan/sin/cos per fish; the kernels need fewer blocks and no
# trigonometry (C2: ONE block), so the survey figure would put the fraction above 1.
ALGO_LANE_OPS = {"c2": 187.0, "c3": 187.0, "c5": 777.0, "c4": 225.0}


def workload_project(name, n_fish, iterations, days):
    from oracle import fixtures
    if name == "c2":
        return fixtures.c2_facility(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0)
    if name == "c5":
        return fixtures.c5_cascade(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0)
    if name == "c3":
        return fixtures.c3_population(n_species=20, iterations=iterations, days=days, curr_Q=9000.0)
    if name == "c4":
        return fixtures.c4_barotrauma(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0)
    raise SystemExit(f"unknown workload {name}")


def build_plan(prj, name):
    os.environ.setdefault("STRYKE_MAX_DAILY_FISH", "2000000000")
    os.environ.setdefault("STRYKE_MAX_DAILY_STATE_ROWS", "2000000000000")
    import gpu_util as G
    sim, plan = G.make_plan(prj, name)
    return sim, plan


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line): NVML polled every 2 ms
    (the same counters ``nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`` prints), with the
    nvidia-smi command itself as the fallback when the NVML binding is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvml"
        self.nv = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists plain ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.nv = pynvml
        except Exception:
            self.nv, self.source = None, "nvidia-smi"

    def _nvml_sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        for nm, bit in zip(self.NAMES, (nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                                        nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(nm)

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")] if out else []
        if r and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if len(r) > 1 and r[1].replace(".", "").isdigit():
            self.mx.append(float(r[1]))
        for i, nm in enumerate(self.NAMES):
            if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._nvml_sample()
                else:
                    self._smi_sample()
            except Exception:
                pass
            time.sleep(0.002 if self.nv is not None else 0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


CPU_SAMPLE = {"c2": dict(n_fish=200_000, iterations=24, days=1), "c5": dict(n_fish=100_000, iterations=12, days=1),
              "c4": dict(n_fish=200_000, iterations=24, days=1),
              "c3": dict(n_fish=0, iterations=2, days=40)}    # per worker process and per step


def cpu_port_rate(name, procs, seed=0, scale=1):
    """The oracle (NumPy port of the reference path, oracle/stryke_oracle.run_project) on the host cores, fanned out
    over ``procs`` processes (the reference itself is single-threaded): a bounded sample of the same workload."""
    import multiprocessing as mp
    cfg = dict(CPU_SAMPLE[name])
    cfg["iterations"] = max(1, cfg["iterations"] * scale)
    t0 = time.perf_counter()
    if procs == 1:
        parts = [_cpu_worker((name, cfg, seed))]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_cpu_worker, [(name, cfg, seed + i) for i in range(procs)])
    dt = time.perf_counter() - t0
    done, cells = sum(p[0] for p in parts), sum(p[1] for p in parts)
    desc = (f"{done} fish-passages ({cells} cells) in {dt:.1f} s on {procs} processes: per process "
            f"{cfg['iterations']} iterations x {cfg['days']} day(s)" + (f" x {cfg['n_fish']} fish" if cfg["n_fish"] else " x 20 species, EPRI counts"))
    return done / dt, done, dt, desc


def _cpu_worker(args):
    name, cfg, seed = args
    from oracle import stryke_oracle as O
    prj = workload_project(name, cfg["n_fish"], cfg["iterations"], cfg["days"])
    out = O.run_project(prj, seed=seed)
    return int(sum(int(r["daily"][0]) for r in out)), len(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--fish", type=int, default=None, help="fish per cell (fixed-count workloads)")
    ap.add_argument("--iterations", type=int, default=None)
    ap.add_argument("--days", type=int, default=None)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel-variant", type=int, default=0, help="0 auto, 1 general, 2 generic throughput, 3 NVRTC-specialised")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    defaults = {"c2": (1_000_000, 1000, 1), "c5": (250_000, 400, 1), "c3": (0, 1000, 365), "c4": (1_000_000, 500, 1)}[args.workload]
    n_fish = args.fish if args.fish is not None else defaults[0]
    iterations = args.iterations if args.iterations is not None else defaults[1]
    days = args.days if args.days is not None else defaults[2]
    nproc = os.cpu_count() or 1
    workload_desc = {"c2": f"C2: 3-unit Kaplan+Kaplan+Francis facility, bypass+spill, 9 nodes, 6 moves; {n_fish} fish/iteration x "
                           f"{iterations} iterations x {days} day(s) per GPU",
                     "c5": f"C5: 5-dam cascade, 40 nodes, 22 moves; {n_fish} fish x {iterations} iterations x {days} day(s) per GPU",
                     "c4": f"C4: 6-unit propeller / Francis station with the Pflugrath barotrauma response, 11 nodes, 6 moves; {n_fish} "
                           f"fish x {iterations} iterations x {days} day(s) per GPU",
                     "c3": f"C3: population mode, 20 species x {days} days x {iterations} iterations per GPU, EPRI-fitted entrainment"}[args.workload]

    # ---------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # The reference is pure Python and cannot travel to the GPU box; its path is timed through the NumPy port
        # (oracle/stryke_oracle.py) on all host cores, on a bounded sample of the same workload.
        if rank != 0:
            return
        vals, desc = [], ""
        for i in range(args.warmup + args.steps):
            rate, done, dt, desc = cpu_port_rate(args.workload, nproc, seed=100 + i)
            if i >= args.warmup:
                vals.append((rate, done, dt))
        rate = float(sum(v[1] for v in vals) / sum(v[2] for v in vals))
        line = {"impl": "reference", "metric": "simulated fish-passages/sec (route+strike+survival)", "value": rate,
                "unit": "fish-passages/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([v[2] for v in vals]) * 1e3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_desc},
                "cpu_baseline": {"value": rate, "unit": "fish-passages/s", "cores": nproc, "kind": "port",
                                 "sample": f"per step: {desc} (NumPy port of Stryke/stryke.py:2897-3383; the unmodified "
                                           "reference measured 2087 fish-passages/s on one core, BASELINE.md §2)"},
                "e2e": {"value": rate, "unit": "fish-passages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from stryke_b200 import _abi, engine
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prj = workload_project(args.workload, n_fish, iterations, days)
    with contextlib.redirect_stdout(io.StringIO()):
        sim, plan_h = build_plan(prj, f"bench_{args.workload}_{rank}")
    # partition-independent Philox keys: this rank's iterations are global iterations [rank*I, (rank+1)*I)
    cell_id = plan_h.cell_id + np.uint64(rank) * np.uint64(1 << 40)
    species = plan_h.species_table
    plan = engine.Plan(plan_h.fac, plan_h.days, species, plan_h.cell_day, plan_h.cell_species, cell_id,
                       device=local_rank, precision=args.precision, seed=0x5EED, max_daily_fish=2_000_000_000,
                       kernel_variant=args.kernel_variant)
    t_n, t_daily, t_sd, t_all = plan.allocate(with_base=True)      # three views of one flat buffer
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    stream = torch.cuda.current_stream()

    def step(seed):
        # fresh Philox key per step
        plan.set_seed(seed)
        plan.launch(t_n, t_daily, t_sd)
        if world > 1:
            dist.all_reduce(t_all)       # ONE NCCL all-reduce for cell_n + Daily + State_Daily

    for i in range(args.warmup):
        step(1000 + i)
        flush_buf.fill_(i & 0xff)
    torch.cuda.synchronize()
    fish_per_step = plan.status()
    sampler = ClockSampler(local_rank)      # NVML set-up takes a few ms and differs from process to process: before the barrier
    sampler.start()
    plan.time_kernel(True)       # CUDA events around the passage kernel of every timed launch, on the launch stream
    align = torch.zeros(1, device=f"cuda:{local_rank}")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.sm.clear()           # keep only the samples taken during the timed region
    launches0 = engine.launch_count()
    if world > 1:
        # device-side start line (untimed): the ranks' host threads leave the barrier a few ms apart; without this the
        # earliest rank's first step would contain that host skew as time spent waiting in the first all-reduce
        dist.all_reduce(align)
    evs = []
    for i in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step(2000 + i)
        e1.record(stream)
        evs.append((e0, e1))
        flush_buf.fill_((i + 1) & 0xff)     # L2 flush between timed steps (256 MiB > 126 MB L2), outside the event pairs
    torch.cuda.synchronize()
    launches = engine.launch_count() - launches0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms, kernel_n = plan.kernel_ms()
    plan.time_kernel(False)
    plan.status()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    t_fish = torch.tensor([float(fish_per_step)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_fish, op=dist.ReduceOp.SUM)
    ms_total, fish_all = float(t_ms.item()), float(t_fish.item())
    value = fish_all * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (passage_kernel): instruction-issue bound ----------------------------
    peak_ops, cal_ms = engine.calibrate_issue(local_rank, stream.cuda_stream)
    algo = ALGO_LANE_OPS[args.workload]
    per_gpu_rate = fish_per_step * args.steps / (ms_total * 1e-3)
    if kernel_n <= 0 or kernel_ms <= 0.0:      # (cannot happen with the in-tree library; keeps the line printable)
        kernel_ms, kernel_n = ms, args.steps
    achieved = fish_per_step * kernel_n / (kernel_ms * 1e-3) * algo      # this rank's passage kernel alone
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_c3_passage_metrics.json" if args.workload == "c3" else "r01_passage_metrics.json")
    if os.path.isfile(prof):
        with open(prof) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    roofline = {"bound": "issue", "achieved": achieved / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                "frac": achieved / peak_ops, "traffic": traffic,
                "kernel_ms_per_launch": kernel_ms / max(kernel_n, 1), "kernel_share_of_step": kernel_ms / ms,
                "frac_of_whole_step": per_gpu_rate * algo / peak_ops,
                "note": "neither HBM nor tensor binds this path (SURVEY.md §8d): bound = SM instruction issue (INT32 Philox on the "
                        "half-rate ALU pipe + FP32/MUFU). "
                        f"achieved = {algo:.0f} algorithmic lane-ops per fish-passage (DESIGN.md §4: the instruction count of the best known "
                        "schedule, = what the specialised kernel executes per fish under ncu) x fish-passages per launch / the passage "
                        "kernel's own launch duration (CUDA events recorded around it on the launch stream inside the timed region, "
                        "stryke_b200_plan_time_kernel; kernel_share_of_step = its share of the step's CUDA-event time, the count / scan "
                        "kernels and the tally memsets are the rest; frac_of_whole_step uses the step time instead); peak = lane-ops/s of "
                        "16 independent FFMA chains per thread at full occupancy (1 warp instruction per scheduler per cycle) measured live by "
                        "stryke_b200_calibrate_issue on this GPU; traffic = DRAM bytes per launch (ncu)",
                "hbm": {"achieved_GBps": (plan.n_cells * (4 + 20 + plan.n_pairs * 8) * 2) / (ms_total / args.steps * 1e-3) / 1e9,
                        "peak_GBps": _measured_peak()}}

    # ---- e2e: the C-ABI call with HOST buffers (tables H2D, kernels, tallies D2H inside the timed region) ---------
    e2e = None
    if rank == 0 or world > 1:
        h2d = sum(a.nbytes for a in (plan_h.fac.node_kind, plan_h.fac.node_apriori, plan_h.fac.node_unit, plan_h.fac.edge_off,
                                     plan_h.fac.edge_dst, plan_h.fac.unit_static, plan_h.days.edge_cdf, plan_h.days.node_stay,
                                     plan_h.days.unit_coef, plan_h.days.curr_Q, species, plan_h.cell_day, plan_h.cell_species, cell_id))
        d2h = plan.n_cells * (4 + _abi.DAILY_W * 4 + plan.n_pairs * 8)
        e2e_steps = max(2, min(args.steps, 5))

        def pinned(a):      # host buffers of the call live in pinned memory (full PCIe rate for both directions)
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t.numpy()
        h_day, h_spc, h_id = pinned(plan_h.cell_day), pinned(plan_h.cell_species), pinned(cell_id)
        out = (pinned(np.zeros(plan.n_cells, np.int32)), pinned(np.zeros((plan.n_cells, _abi.DAILY_W), np.uint32)),
               pinned(np.zeros((plan.n_cells, plan.n_pairs, 2), np.uint32)))
        dev = f"cuda:{local_rank}"
        engine.run_cells(plan_h.fac, plan_h.days, species, h_day, h_spc, h_id, device=local_rank, precision=args.precision,
                         seed=1, max_daily_fish=2_000_000_000, out=out, kernel_variant=args.kernel_variant)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            r = engine.run_cells(plan_h.fac, plan_h.days, species, h_day, h_spc, h_id, device=local_rank,
                                 precision=args.precision, seed=3000 + i, max_daily_fish=2_000_000_000, out=out,
                                 kernel_variant=args.kernel_variant)
            if world > 1:   # the public multi-GPU call ends with the tally all-reduce
                td = torch.from_numpy(r.daily.view(np.int32)).to(dev, non_blocking=True)
                ts = torch.from_numpy(r.state_daily.view(np.int32)).to(dev, non_blocking=True)
                dist.all_reduce(td)
                dist.all_reduce(ts)
                td.cpu(), ts.cpu()
        dt = time.perf_counter() - t0
        t_dt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
        e2e = {"value": fish_all * e2e_steps / float(t_dt.item()), "unit": "fish-passages/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "how": "stryke_b200_run through ctypes: host tables in, count+scan+passage kernels, host tallies out; wall clock"}

    # ---- every rank's own figures (a straggling GPU shows up here; the headline uses the max over ranks) -----------
    clocks = sampler.summary()
    per_rank = None
    if world > 1:
        names = ClockSampler.NAMES
        mine = torch.tensor([ms / args.steps, kernel_ms / max(kernel_n, 1), clocks["sm_mhz"] or 0.0,
                             float(sum(1 << i for i, nm in enumerate(names) if nm in clocks["reasons"]))],
                            dtype=torch.float64, device=f"cuda:{local_rank}")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rows = [[float(x) for x in t.tolist()] for t in allr]
        per_rank = {"step_ms": [r[0] for r in rows], "kernel_ms": [r[1] for r in rows], "sm_mhz": [r[2] for r in rows]}
        bits = 0
        for r in rows:
            bits |= int(r[3])
        clocks = dict(clocks, sm_mhz=min(r[2] for r in rows), reasons=[nm for i, nm in enumerate(names) if bits >> i & 1],
                      note="sm_mhz = lowest per-rank median, reasons = union over ranks")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, done, dt, desc = cpu_port_rate(args.workload, nproc, seed=5, scale=2)
        cpu = {"value": rate, "unit": "fish-passages/s", "cores": nproc, "kind": "port",
               "sample": f"{desc} (NumPy port oracle/stryke_oracle.py; unmodified reference: 2087 fish-passages/s/core, "
                         "BASELINE.md §2)"}
    if rank == 0:
        line = {"metric": "simulated fish-passages/sec (route+strike+survival)", "value": value, "unit": "fish-passages/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
                "config": {"workload": workload_desc, "rng": "Philox4x32-10 keyed (seed; cell id, fish, slot)",
                           "l2": "256 MiB flush write between timed steps; kernel inputs are < 64 KiB of tables",
                           "fish_passages_per_step": fish_all, "cells_per_gpu": plan.n_cells, "kernel": plan.kernel(),
                           "collective": "one NCCL all_reduce(sum) of the flat tally buffer (cell_n + Daily + State_Daily) per step" if world > 1 else "none"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks}
        if per_rank:
            line["per_rank"] = per_rank
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs")
    return 6650.0


if __name__ == "__main__":
    main()
