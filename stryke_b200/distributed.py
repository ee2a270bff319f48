"""Multi-GPU layer: one process per GPU, cells sharded over ranks, ONE all-reduce of the tally arrays.

Scenario x species x iteration x day cells are independent (the reference's loops share nothing but RNG streams and
the HDF append, Stryke/stryke.py:2806-2902) and the Philox counter is keyed by the logical cell id, so any partition
of the cell list gives the same tallies (tests/test_gpu_native_rng.py).  Each rank simulates a contiguous block of
the engine-order cell list, writes its block into zero-initialised full-size tally tensors and the ranks sum them
with ``torch.distributed.all_reduce`` (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests) — there is no
exchange inside the simulation, so no fused compute+collective kernel exists or is needed on this path.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_cells, rank, world):
    """Contiguous block of cells of ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n_cells), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_weight(weights, world):
    """Contiguous blocks with (nearly) equal total weight — weights = expected fish per cell.  Returns world+1 bounds."""
    w = np.asarray(weights, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)


def allreduce_tallies(tensors, group=None):
    """Sum the tally tensors over ranks in place (each rank contributed zeros outside its own block)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def run_sharded(plan_host, rank, world, device, *, precision=32, seed=0, max_daily_fish=100000, group=None):
    """Simulate this rank's block of ``plan_host`` (a ``runner.RunPlan``) on ``device`` and all-reduce.

    Returns full-size (cell_n, daily, state_daily) torch tensors (int32, summed over ranks) on ``device``."""
    import torch
    from . import _abi, engine
    lo, hi = shard_bounds(plan_host.n_cells, rank, world)
    dev = torch.device("cuda", device)
    C, P = plan_host.n_cells, plan_host.fac.n_pairs
    full_n, full_d, full_s, base = engine.tally_tensors(C, P, device, with_base=True)     # views of one flat buffer
    if hi > lo:
        plan = engine.Plan(plan_host.fac, plan_host.days, plan_host.species_table, plan_host.cell_day[lo:hi],
                           plan_host.cell_species[lo:hi], plan_host.cell_id[lo:hi], device=device, precision=precision,
                           seed=seed, max_daily_fish=max_daily_fish)
        # the block is a contiguous slice of the full tensors: the kernels write in place, no copy
        plan.launch(full_n[lo:hi], full_d[lo:hi], full_s[lo:hi])
        plan.status()
        plan.close()
    allreduce_tallies([base], group)          # ONE collective for the three tensors
    return full_n, full_d, full_s
