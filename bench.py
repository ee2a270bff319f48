#!/usr/bin/env python
"""bench.py — simulated fish-passages/s of the B200 passage engine (BASELINE.json metric) and of the CPU baseline.

  python bench.py [--gpus N --steps K --warmup W] [--workload c2|c3|c4|c5] [--also c3,c5] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A *step* is one pass of the hot path (count -> passage kernels, compact tally rows) over the workload's cells with a fresh
Philox seed, followed — on N > 1 ranks — by the exchange of the ranks' rows over NVLink.  The headline workload ``c2`` is
BASELINE.json configs[1]: the synthetic 3-unit Kaplan+Kaplan+Francis facility with bypass and spill, 1e6 fish/iteration x
1000 iterations over its 30-day hydrograph = 3e10 fish-passages per step on one B200 (weak scaling: every rank simulates
its own 1000 iterations; the Philox counter is keyed by the global iteration).  The same run also measures, and reports
under ``workloads`` in the same JSON line, the two configurations BASELINE.json shards over 8 GPUs:

  c3  population mode, 20 species x 365 days x 1000 iterations x 5 flow scenarios = 3.65e7 cells — STRONG scaling, the
      iterations of every (scenario, species) group are split over the ranks (stryke.py:2897);
  c5  the 5-dam cascade, 1.25e9 fish per GPU (weak scaling; 1e10 fish on 8 GPUs).
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

SCENARIOS5 = ["Q050", "Q075", "Q100", "Q150", "Q200"]


def workload_project(name, n_fish, iterations, days):
    from stryke_b200 import workloads as W
    if name == "c2":
        return W.c2_facility(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0)
    if name == "c5":
        return W.c5_cascade(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0)
    if name == "c3":
        return W.c3_population(n_species=20, iterations=iterations, days=days, curr_Q=9000.0, scenarios=SCENARIOS5,
                               facility_scenario_column=False)
    if name == "c4":      # BASELINE.json configs[3] as named: propeller / pump station, friction-loss pressures
        return W.c4_barotrauma(n_fish=n_fish, iterations=iterations, days=days, curr_Q=9000.0, pump=True, friction=True)
    raise SystemExit(f"unknown workload {name}")


DEFAULTS = {"c2": (1_000_000, 1000, 30), "c5": (1_250_000, 1000, 1), "c3": (0, 1000, 365), "c4": (1_000_000, 1000, 10)}
SCALING = {"c2": "weak", "c4": "weak", "c5": "weak", "c3": "strong"}


def describe(name, n_fish, iterations, days, world):
    per = "per GPU" if SCALING[name] == "weak" else f"in total, iterations split over {world} GPU(s)"
    return {"c2": f"C2: 3-unit Kaplan+Kaplan+Francis facility, bypass+spill, 9 nodes, 6 moves; {n_fish} fish/iteration x "
                  f"{iterations} iterations x {days} day(s) {per}",
            "c5": f"C5: 5-dam cascade, 40 nodes, 22 moves; {n_fish} fish x {iterations} iterations x {days} day(s) {per}",
            "c4": f"C4: 6-unit propeller / pump station, Pflugrath barotrauma response with friction-loss pressures, 12 nodes, 6 moves; {n_fish} "
                  f"fish x {iterations} iterations x {days} day(s) {per}",
            "c3": f"C3: population mode, 20 species x {days} days x {iterations} iterations x 5 flow scenarios {per}, "
                  "EPRI-fitted entrainment"}[name]


def op_budget(fac, species_row):
    """Kernel-independent budget of lane-ops per fish-passage: what ANY one-thread-per-fish schedule of the
    algorithm has to issue, counted from the facility's structure — not read back from a profile of our kernel.
      Philox4x32-10: 47 per block of four 32-bit words (40 IMAD.WIDE / LOP3 + 7 for counter and key), one word per draw
        class per level that can need it (r/R + cause share a word, survival dice, routing), + the length word;
      uniform conversion 3 per draw; length (Box-Muller + lognormal + width) 15;
      turbine level 40 (impingement, unified Franke strike formula, cause roll) + 25 with barotrauma;
      a-priori level with survival < 1: 4; routing draw 15 (cdf row compare + successor load);
      per level: tallies 4 (arrivals + survivors into the cell's counters), bookkeeping 5 (state, first-'U' rule, loop);
      per fish: cell / tile bookkeeping 8."""
    M = fac.n_moves
    kinds = fac.node_kind
    level = [[int(fac.pair_node[p]) for p in range(fac.level_off[k], fac.level_off[k + 1])] for k in range(M)]
    deg = np.diff(fac.edge_off)
    words, ops = 1, 15.0 + 3.0 + 8.0
    for k, nodes in enumerate(level):
        turb = any(kinds[v] != 0 for v in nodes)
        dice = turb or any(not (np.float32(fac.node_apriori[v]) >= 1.0) for v in nodes)
        route = k < M - 1 and any(deg[v] > 1 for v in nodes)
        if turb:
            # the share of fish that stand at a turbine node of this level is workload dependent; count the level in full
            words += 1
            ops += 40.0 + 6.0 + (25.0 if fac.barotrauma else 0.0)
        if dice:
            words += 1
            ops += 3.0 + (0.0 if turb else 4.0)
        if route:
            words += 1
            ops += 3.0 + 15.0
        ops += 4.0 + 5.0
    blocks = -(-words // 4)
    return ops + 47.0 * blocks, {"moves": M, "philox_blocks": blocks, "words": words}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line): NVML polled every 2 ms
    (the same counters ``nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`` prints), with the
    nvidia-smi command itself as the fallback when the NVML binding is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.paused = index, False, True
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
                if not self.paused:
                    if self.nv is not None:
                        self._nvml_sample()
                    else:
                        self._smi_sample()
            except Exception:
                pass
            time.sleep(0.002 if self.nv is not None else 0.05)

    def window(self):
        """Start a sampling window (the timed region of one workload)."""
        self.sm, self.reasons, self.paused = [], set(), False

    def summary(self):
        self.paused = True
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
            f"{cfg['iterations']} iterations x {cfg['days']} day(s)" + (f" x {cfg['n_fish']} fish" if cfg["n_fish"] else " x 20 species x 5 scenarios, EPRI counts"))
    return done / dt, done, dt, desc


def _cpu_worker(args):
    name, cfg, seed = args
    from oracle import stryke_oracle as O          # the CPU baseline leg: the one place bench.py executes oracle/
    prj = workload_project(name, cfg["n_fish"], cfg["iterations"], cfg["days"])
    out = O.run_project(prj, seed=seed)
    return int(sum(int(r["daily"][0]) for r in out)), len(out)


REF_SAMPLE = dict(n_fish=2000, days=3)      # per worker process and per step (~3 s of reference code + its import)


def reference_rate(procs, n_fish=REF_SAMPLE["n_fish"], days=REF_SAMPLE["days"]):
    """The UNMODIFIED reference (oracle/_ref, a copy of Stryke/stryke.py made by oracle/make_ref.py where the reference
    tree exists) through its own run() under the import stubs of oracle/harness.py, one process per host core."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_ref_worker, [(n_fish, days, 100 + i) for i in range(procs)])
    dt = time.perf_counter() - t0
    done = sum(parts)
    return done / dt, done, dt


def _ref_worker(args):
    n_fish, days, seed = args
    os.environ["STRYKE_REFERENCE_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
    from oracle import harness
    from stryke_b200 import workloads as W
    prj = W.c2_facility(n_fish=n_fish, iterations=1, days=days)
    sim, store, rec = harness.run_reference(prj, seed=seed, traces=False)
    return int(store["Daily"]["pop_size"].sum())


# ------------------------------------------------------------------------------------------------------------------
def measure(name, args, rank, local_rank, world, sampler, main):
    """One workload at ``world`` ranks.  Returns the fields of its JSON object (rank 0) — value, ms_per_step, roofline,
    e2e, gpu_launches, clocks, per_rank."""
    import torch
    import torch.distributed as dist
    from stryke_b200 import engine, workloads as W
    from stryke_b200.distributed import SLAB_SHARES, ShardedRun, slab_plan
    n_fish, iterations, days = DEFAULTS[name]
    if main:
        n_fish = args.fish if args.fish is not None else n_fish
        iterations = args.iterations if args.iterations is not None else iterations
        days = args.days if args.days is not None else days
    os.environ.setdefault("STRYKE_MAX_DAILY_FISH", "2000000000")
    os.environ.setdefault("STRYKE_MAX_DAILY_STATE_ROWS", "2000000000000")
    prj = workload_project(name, n_fish, iterations, days)
    t_host0 = time.perf_counter()
    sim, plan_h = W.make_plan(prj, f"bench_{name}_{rank}")
    host_plan_s = time.perf_counter() - t_host0
    species = plan_h.species_table
    strong = SCALING[name] == "strong"
    if strong:
        grids = [plan_h.grid.iterations_slice(r / world, (r + 1) / world) for r in range(world)]
    else:   # weak: every rank owns `iterations` iterations of its own — global iterations [rank * I, (rank + 1) * I)
        grids = []
        for r in range(world):
            g = engine.CellGrid(plan_h.grid.groups, plan_h.grid.active_days, plan_h.grid.id_days, r << 40)
            grids.append(g)
    grid = grids[rank]
    n_max = max(g.n_cells for g in grids)
    dev = f"cuda:{local_rank}"
    cap_scale, run = 1.0, None
    while True:
        run = ShardedRun(plan_h.fac, plan_h.days, species, grid, rank, world, local_rank, precision=args.precision, seed=0x5EED,
                         max_daily_fish=2_000_000_000, kernel_variant=args.kernel_variant, n_cells_max=n_max,
                         slabs=slab_plan(grids, species, n_max, shares=SLAB_SHARES if world > 1 else (1.0,), scale=cap_scale))
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream()
        for i in range(args.warmup):
            run.step(1000 + i)
            flush_buf.fill_(i & 0xff)
        torch.cuda.synchronize()
        over = torch.tensor([1 if run.slab_overflow() else 0], device=dev)
        if world > 1:
            dist.all_reduce(over)
        if int(over.item()) == 0:
            break
        run.close()
        cap_scale *= 2.0
    run.plan.status()
    _, fish_per_step = run.plan.compact_totals()
    sampler.window()
    run.plan.time_kernel(True)       # CUDA events around the passage kernel of every timed launch, on the launch stream
    align = torch.zeros(1, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.window()                 # keep only the samples taken during the timed region
    launches0 = engine.launch_count()
    if world > 1:
        # device-side start line (untimed): the ranks' host threads leave the barrier a few ms apart; without this the
        # earliest rank's first step would contain that host skew as time spent waiting in the first exchange
        dist.all_reduce(align)
    evs = []
    for i in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run.step(2000 + i)
        e1.record(stream)
        evs.append((e0, e1))
        flush_buf.fill_((i + 1) & 0xff)     # L2 flush between timed steps (256 MiB > 126 MB L2), outside the event pairs
    torch.cuda.synchronize()
    launches = engine.launch_count() - launches0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.summary()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms, kernel_n = run.plan.kernel_ms()
    run.plan.time_kernel(False)
    run.plan.status()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    t_fish = torch.tensor([float(fish_per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_fish, op=dist.ReduceOp.SUM)
    ms_total, fish_all = float(t_ms.item()), float(t_fish.item())
    value = fish_all * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (passage kernel): instruction-issue bound ----------------------------------
    peak_ops, cal_ms = engine.calibrate_issue(local_rank, stream.cuda_stream)
    budget, budget_info = op_budget(plan_h.fac, species[0])
    if kernel_n <= 0 or kernel_ms <= 0.0:      # (cannot happen with the in-tree library; keeps the line printable)
        kernel_ms, kernel_n = ms, args.steps
    kernel_ms_step = kernel_ms / args.steps                     # all of a step's passage launches (one per slab of cells)
    achieved = fish_per_step / (kernel_ms_step * 1e-3) * budget           # this rank's passage kernels alone
    prof = {}
    pf = os.path.join(ROOT, "profiles", f"r02_{name}_passage_metrics.json")
    if os.path.isfile(pf):
        with open(pf) as fh:
            prof = json.load(fh)
    executed = prof.get("warp_inst_per_fish")      # executed warp instructions per simulated fish under ncu (x 32 lanes)
    rows_bytes = None
    slots_used = run.slots_used()
    rows_bytes = slots_used * run.L.W * 4
    algo_bytes = run.n_cells * 4 + rows_bytes + ((run.n_cells + 31) // 32) * 12
    roofline = {"bound": "issue", "achieved": achieved / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                "frac": achieved / peak_ops,
                "issue_util": (executed * 32.0 * fish_per_step / (kernel_ms_step * 1e-3) / peak_ops) if executed else None,
                "traffic": prof.get("dram_bytes_per_launch"),
                "op_budget_per_fish": budget, "op_budget_terms": budget_info,
                "kernel_ms_per_step": kernel_ms_step, "kernel_launches_per_step": kernel_n / args.steps,
                "kernel_share_of_step": kernel_ms / ms, "frac_of_whole_step": (fish_per_step * args.steps / (ms * 1e-3)) * budget / peak_ops,
                "note": "neither HBM nor tensor binds this path (SURVEY.md §8d): bound = SM instruction issue (INT32 Philox on the "
                        "half-rate ALU pipe + FP32/MUFU).  frac = op budget x fish-passages / passage-kernel time / peak, where the op "
                        "budget is counted from the facility's structure (bench.op_budget, DESIGN.md §4) and does not depend on our "
                        "kernel; issue_util = warp instructions our kernel actually executed per fish (ncu, profiles/r02_*_passage_"
                        "metrics.json) x 32 x fish-passages / time / peak = the issue-slot utilisation ncu reports.  Kernel time: CUDA "
                        "events around every passage launch on the launch stream inside the timed region; peak: lane-ops/s of 16 "
                        "independent FFMA chains per thread at full occupancy, measured live (stryke_b200_calibrate_issue); traffic: "
                        "DRAM bytes per step of the passage + count kernels (ncu)",
                "hbm": {"algorithmic_bytes_per_step": int(algo_bytes),
                        "achieved_GBps": algo_bytes / (ms / args.steps * 1e-3) / 1e9, "peak_GBps": _measured_peak()}}

    # ---- e2e: host buffers in the timed region ---------------------------------------------------------------------------
    e2e_steps = max(2, min(args.steps, 5))
    h2d = int(sum(np.asarray(a).nbytes for a in (plan_h.fac.node_kind, plan_h.fac.node_apriori, plan_h.fac.node_unit, plan_h.fac.edge_off,
                                                 plan_h.fac.edge_dst, plan_h.fac.unit_static, plan_h.days.edge_cdf, plan_h.days.node_stay,
                                                 plan_h.days.unit_coef, plan_h.days.curr_Q, species, grid.groups, grid.active_days)))
    d2h = int(rows_bytes + ((run.n_cells + 31) // 32) * 12)
    if world == 1:
        # the C-ABI call a binding inside the reference's run() makes: stryke_b200_run_compact with HOST buffers — tables
        # and the cell grid go up, count + passage kernels run slab by slab, every slab's rows come down under the next slab's
        # kernels; pinned result buffers, wall clock around the calls
        run.close()
        del run
        torch.cuda.empty_cache()
        bufs = engine.compact_buffers(grid.n_cells, plan_h.fac.n_pairs, int(slots_used * 1.02) + 4096)
        kw = dict(device=local_rank, precision=args.precision, max_daily_fish=2_000_000_000, kernel_variant=args.kernel_variant,
                  rows_cap=bufs[1].shape[0], buffers=bufs)
        engine.run_compact(plan_h.fac, plan_h.days, species, grid, seed=1, **kw)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            r = engine.run_compact(plan_h.fac, plan_h.days, species, grid, seed=3000 + i, **kw)
        dt = time.perf_counter() - t0
        e2e = {"value": fish_all * e2e_steps / dt, "unit": "fish-passages/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(r.n_slots * run_w(plan_h) * 4 + ((grid.n_cells + 31) // 32) * 12),
               "how": "stryke_b200_run_compact through ctypes: host tables + cell grid in, count + passage kernels per slab of cells, "
                      "compact tally rows out to pinned host memory (copy of slab k under the kernels of slab k + 1); wall clock"}
    else:
        # the repo's multi-GPU call (distributed.ShardedRun.step): kernels, the exchange of the ranks' rows over NVLink AND the
        # copy of this rank's own rows to its pinned host buffer (one PCIe link per rank), wall clock, max over ranks
        run.step(2999, to_host=True)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            run.step(3000 + i, to_host=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_dt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t_dt, op=dist.ReduceOp.MAX)
        e2e = {"value": fish_all * e2e_steps / float(t_dt.item()), "unit": "fish-passages/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": d2h,
               "how": f"distributed.ShardedRun.step(to_host=True) on every rank: tables and cell grid resident, kernels + exchange "
                      f"({run.mode}) + this rank's rows to its pinned host buffer; bytes are per rank; wall clock, max over ranks"}
        exchange_mode = run.mode
        run.close()
        del run
        torch.cuda.empty_cache()

    per_rank = None
    if world > 1:
        names = ClockSampler.NAMES
        mine = torch.tensor([ms / args.steps, kernel_ms_step, clocks["sm_mhz"] or 0.0,
                             float(sum(1 << i for i, nm in enumerate(names) if nm in clocks["reasons"])), float(fish_per_step)],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rows = [[float(x) for x in t.tolist()] for t in allr]
        per_rank = {"step_ms": [r[0] for r in rows], "kernel_ms": [r[1] for r in rows], "sm_mhz": [r[2] for r in rows],
                    "fish": [r[4] for r in rows]}
        bits = 0
        for r in rows:
            bits |= int(r[3])
        clocks = dict(clocks, sm_mhz=min(r[2] for r in rows), reasons=[nm for i, nm in enumerate(names) if bits >> i & 1],
                      note="sm_mhz = lowest per-rank median, reasons = union over ranks")
    out = {"value": value, "unit": "fish-passages/s", "ms_per_step": ms_total / args.steps, "scaling": SCALING[name],
           "config": {"workload": describe(name, n_fish, iterations, days, world),
                      "rng": "Philox4x32-10 keyed (seed; cell id, fish, slot)",
                      "l2": "256 MiB flush write between timed steps; kernel inputs are < 1 MiB of tables",
                      "fish_passages_per_step": fish_all, "cells_per_gpu": int(grid.n_cells),
                      "tally_rows_per_gpu": int(slots_used), "host_plan_seconds": host_plan_s,
                      "kernel": kernel_name(plan_h, grid, local_rank, args),
                      "collective": ("none" if world == 1 else
                                     ("every rank's rows pushed slab by slab into rank 0's buffer over NVLink (peer copies on side streams, torch "
                                      "symmetric memory) under the next slab's kernels + one 4-byte all-reduce as the fence" if exchange_mode == "push" else
                                      "one NCCL all_gather_into_tensor of the ranks' flat row buffers per step"))},
           "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if per_rank:
        out["per_rank"] = per_rank
    return out


def run_w(plan_h):
    return plan_h.fac.n_pairs + 3


_KNAME = {}


def kernel_name(plan_h, grid, device, args):
    from stryke_b200 import engine
    p = engine.Plan(plan_h.fac, plan_h.days, plan_h.species_table, grid, None, None, device=device, precision=args.precision,
                    seed=1, max_daily_fish=2_000_000_000, kernel_variant=args.kernel_variant)
    nm = p.kernel()
    p.close()
    return nm


def api_timing(device):
    """The class API end to end (what drops into webapp/app.py:4530-4539): simulation.run() and summary() wall clock on a
    C3-shaped project (20 species x 365 days x 100 iterations = 730 000 cells) and on a C2 project."""
    import tempfile
    from stryke_b200 import workloads as W
    from stryke_b200.store import open_store
    out = {}
    for tag, prj in (("c3_730k_cells", W.c3_population(n_species=20, iterations=100, days=365)),
                     ("c2_30M_fish", W.c2_facility(n_fish=1_000_000, iterations=30, days=1))):
        tmp = tempfile.mkdtemp(prefix="bench_api_")
        sim = W.make_sim(prj, "api", proj_dir=tmp)
        sim.device = device
        inputs = (("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
                  ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges), ("Unit_Parameters", sim.unit_params),
                  ("Facilities", sim.facility_params), ("Hydrograph", sim.input_hydrograph_df))
        with open_store(sim.hdf_path, mode="w") as st:
            for key, df in inputs:
                st[key] = df
        with contextlib.redirect_stdout(io.StringIO()):
            tw = time.perf_counter()
            sim.run()                       # first call in the process: includes the NVRTC specialisation for this facility
            first = time.perf_counter() - tw
            with open_store(sim.hdf_path, mode="w") as st:      # start over: run() appends
                for key, df in inputs:
                    st[key] = df
            t0 = time.perf_counter()
            sim.run()
            t1 = time.perf_counter()
            sim.summary()
            t2 = time.perf_counter()
        fish = int(sum(int(t.n.sum()) for t in sim._last_tallies.values()))
        out[tag] = {"fish_passages": fish, "run_s": t1 - t0, "first_run_s": first, "summary_s": t2 - t1,
                    "run_fish_passages_per_s": fish / (t1 - t0), "run_plus_summary_fish_passages_per_s": fish / (t2 - t0)}
        import shutil
        from stryke_b200.store import wait_for_exports
        wait_for_exports()                  # (the .h5 export of the last close runs on a worker thread)
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--also", default="c3,c5", help="further workloads measured in the same run and reported under 'workloads'")
    ap.add_argument("--fish", type=int, default=None, help="fish per cell (fixed-count workloads)")
    ap.add_argument("--iterations", type=int, default=None)
    ap.add_argument("--days", type=int, default=None)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-api", action="store_true")
    ap.add_argument("--kernel-variant", type=int, default=0, help="0 auto, 1 general, 2 generic throughput, 3 NVRTC-specialised")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nproc = os.cpu_count() or 1
    n_fish, iterations, days = DEFAULTS[args.workload]
    n_fish = args.fish if args.fish is not None else n_fish
    iterations = args.iterations if args.iterations is not None else iterations
    days = args.days if args.days is not None else days
    workload_desc = describe(args.workload, n_fish, iterations, days, world)

    # ---------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        # The reference's own CPU implementation of the path on the box's host cores: the unmodified reference
        # (oracle/_ref) when that copy exists, else the NumPy port (oracle/stryke_oracle.py); bounded sample per step.
        if rank != 0:
            return
        have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "Stryke", "stryke.py"))
        vals, desc = [], ""
        for i in range(args.warmup + args.steps):
            if have_ref:
                rate, done, dt = reference_rate(nproc)
                desc = (f"{done} fish-passages in {dt:.1f} s on {nproc} processes: the UNMODIFIED reference (oracle/_ref = copy of "
                        f"Stryke/stryke.py) through its own run() under oracle/harness.py, C2 facility, {REF_SAMPLE['n_fish']} fish x "
                        f"{REF_SAMPLE['days']} days per process (wall clock includes each process's import of the reference)")
            else:
                rate, done, dt, desc = cpu_port_rate(args.workload, nproc, seed=100 + i)
            if i >= args.warmup:
                vals.append((rate, done, dt))
        rate = float(sum(v[1] for v in vals) / sum(v[2] for v in vals))
        kind = "reference" if have_ref else "port"
        line = {"impl": "reference", "metric": "simulated fish-passages/sec (route+strike+survival)", "value": rate,
                "unit": "fish-passages/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([v[2] for v in vals]) * 1e3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_desc},
                "cpu_baseline": {"value": rate, "unit": "fish-passages/s", "cores": nproc, "kind": kind,
                                 "sample": f"per step: {desc}" + ("" if have_ref else " (NumPy port of Stryke/stryke.py:2897-3383; the "
                                           "unmodified reference measured 2087 fish-passages/s on one core, BASELINE.md §2)")},
                "e2e": {"value": rate, "unit": "fish-passages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)      # NVML set-up takes a few ms and differs from process to process: before any barrier
    sampler.start()
    main_out = measure(args.workload, args, rank, local_rank, world, sampler, main=True)
    others = {}
    for nm in [x for x in args.also.split(",") if x and x != args.workload]:
        others[nm] = measure(nm, args, rank, local_rank, world, sampler, main=False)
    sampler.stop_flag = True

    cpu = api = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, done, dt, desc = cpu_port_rate(args.workload, nproc, seed=5, scale=2)
        cpu = {"value": rate, "unit": "fish-passages/s", "cores": nproc, "kind": "port",
               "sample": f"{desc} (NumPy port oracle/stryke_oracle.py; unmodified reference: 2087 fish-passages/s/core, "
                         "BASELINE.md §2)"}
        if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "Stryke", "stryke.py")):
            # the unmodified reference itself, where its files travelled (oracle/make_ref.py); the port stays beside it
            r_rate, r_done, r_dt = reference_rate(nproc)
            cpu = {"value": r_rate, "unit": "fish-passages/s", "cores": nproc, "kind": "reference",
                   "sample": f"{r_done} fish-passages in {r_dt:.1f} s on {nproc} processes: the UNMODIFIED reference (oracle/_ref) through "
                             f"its own run(), C2 facility, {REF_SAMPLE['n_fish']} fish x {REF_SAMPLE['days']} days per process",
                   "port": cpu}
    if rank == 0 and world == 1 and not args.no_api:
        api = api_timing(local_rank)
    if rank == 0:
        line = {"metric": "simulated fish-passages/sec (route+strike+survival)", "value": main_out["value"], "unit": "fish-passages/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_out["ms_per_step"],
                "higher_is_better": True, "scaling": main_out["scaling"], "vs_baseline": None,
                "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
                "config": main_out["config"], "roofline": main_out["roofline"], "cpu_baseline": cpu, "e2e": main_out["e2e"],
                "gpu_launches": main_out["gpu_launches"], "clocks": main_out["clocks"]}
        if "per_rank" in main_out:
            line["per_rank"] = main_out["per_rank"]
        if others:
            line["workloads"] = others
        if api:
            line["e2e_api"] = api
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
