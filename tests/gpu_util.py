"""Shared helpers for the -m gpu parity tests: build a ``stryke_b200.simulation`` from a fixture project and turn
golden / oracle draws into the engine's injected-draw arrays."""
import contextlib
import io

import numpy as np

from oracle import fixtures, stryke_oracle as O
from stryke_b200 import simulation
from stryke_b200.engine import Injected, TraceBuffers, run_cells
from stryke_b200.runner import RunPlan


from stryke_b200.workloads import make_plan, make_sim  # noqa: E402,F401


def reference_cell_index(plan):
    """For every engine-order cell: its index in the reference's loop order (scenario -> species -> iteration -> day)."""
    out = np.zeros(plan.n_cells, np.int64)
    base = 0
    for g in plan.groups:
        D, I = g.n_active_days, g.iterations
        dd = np.repeat(np.arange(D), I)
        ii = np.tile(np.arange(I), D)
        out[g.cell0:g.cell0 + D * I] = base + ii * D + dd
        base += D * I
    return out


def injected_from_cells(plan, ref_idx, cell_draws, presence, count_draw, n_moves):
    """``cell_draws[r]`` is an oracle ``Draws`` (or None for an absent cell) of reference cell r, plus its
    ``length_z``; returns the engine-order ``Injected`` block."""
    C = plan.n_cells
    ns = np.array([0 if cell_draws[r] is None else cell_draws[r][0].dice.shape[1] for r in ref_idx], np.int64)
    off = np.concatenate([[0], np.cumsum(ns)])
    NF, M = int(off[-1]), n_moves
    z = np.zeros(NF)
    dice, rR, depth, route = (np.full((M, NF), np.nan) for _ in range(4))
    cause = np.full((M, 4, NF), np.nan)
    g_rR, g_depth, g_cause = np.full((M, C), np.nan), np.full((M, C), np.nan), np.full((M, 4, C), np.nan)
    for c, r in enumerate(ref_idx):
        if cell_draws[r] is None:
            continue
        d, lz = cell_draws[r]
        s = slice(off[c], off[c + 1])
        z[s] = lz
        dice[:, s], rR[:, s], depth[:, s], route[:, s] = d.dice, d.rR, d.depth_u, d.route
        cause[:, :, s] = np.transpose(d.cause, (0, 2, 1))
        if d.ghost_cause is not None:
            g_rR[:, c], g_depth[:, c], g_cause[:, :, c] = d.ghost_rR, d.ghost_depth, d.ghost_cause
    pres = np.array([presence[r] for r in ref_idx], float)
    cnt = np.array([np.nan if count_draw[r] is None else count_draw[r] for r in ref_idx], float)
    return Injected(off, pres, cnt, z, dice, rR, depth, cause, route, g_rR, g_depth, g_cause), off


def golden_injected(g, plan):
    ref_idx = reference_cell_index(plan)
    draws, pres, cnt = [], [], []
    for ci, cm in enumerate(g.cells):
        pres.append(cm["presence"])
        cnt.append(cm["count_draw"])
        draws.append(None if cm["n"] == 0 else (g.draws(ci), g.ref(ci, "length_z")))
    inj, off = injected_from_cells(plan, ref_idx, draws, pres, cnt, g.meta["moves"])
    return inj, off, ref_idx


def run_injected(plan, inj, precision=64, trace=True, device=0):
    NF = int(inj.fish_off[-1])
    tr = TraceBuffers.allocate(plan.fac.n_moves, NF, probs=True) if trace else None
    res = run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id,
                    device=device, precision=precision, seed=1, max_daily_fish=10 ** 9, injected=inj, trace=tr)
    return res


def state_daily_dict(plan, sd_cell):
    """(n_pairs, 2) tally slice -> {(move, node name): (successes, count)} for populated pairs."""
    fac = plan.fac
    out = {}
    for p in range(fac.n_pairs):
        if sd_cell[p, 1] > 0:
            out[(int(fac.pair_move[p]), fac.node_names[int(fac.pair_node[p])])] = (float(sd_cell[p, 0]), float(sd_cell[p, 1]))
    return out


def full_daily(n, d5):
    """Engine daily[5] -> the reference's 7 Daily counters (with the entrained > 0 gate, stryke.py:3359)."""
    ent, sur, imp, bld, bar = (int(x) for x in d5)
    if ent == 0:
        imp = bld = bar = 0
    return np.array([int(n), ent, sur, ent - sur, imp, bld, bar], np.int64)
