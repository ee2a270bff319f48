"""Turn a reference recording (oracle/harness.py ``Recorder``) into dense slot-tagged ``Draws`` and replay it
through the oracle.  TEST INFRASTRUCTURE (see oracle/stryke_oracle.py header)."""
from __future__ import annotations

import numpy as np

from . import stryke_oracle as O


def cell_normals(cell):
    """Standard normals behind the fish lengths: the last PCG64 ``standard_normal`` batch (stryke.py:3070)."""
    if not cell["moves"]:
        return np.zeros(0)
    n = cell["moves"][0]["n"]
    if "legacy_normal" in cell:      # Length_mean / Length_sd mode: np.random.normal(mean, sd, n), stryke.py:3090
        loc, scale, x = cell["legacy_normal"]
        z = (x - loc) / scale
    else:
        normals = [p for p in cell["pcg"] if p[0] == "standard_normal"]
        z = normals[-1][-1]
    assert z.shape[0] == n, (z.shape, n)
    return z


def cell_lengths(cell, pop_row):
    """Fish lengths (ft) of a recorded cell."""
    if "legacy_normal" in cell:
        return np.abs(cell["legacy_normal"][2]) / 12.0
    return O.lengths_from_normals(cell_normals(cell), pop_row["length shape"], pop_row["length location"],
                                  pop_row["length scale"])


def cell_count_draw(cell):
    """The PCG64 draw ``population_sim`` consumed (None in anadromous mode)."""
    pre = cell["pcg"][:-1] if (cell["moves"] and "legacy_normal" not in cell) else cell["pcg"]
    if not pre:
        return None
    return float(pre[0][-1][0])


def draws_from_recording(cell, pop_row):
    moves = cell["moves"]
    M = len(moves)
    n = moves[0]["n"]
    dice = np.stack([m["dice"] for m in moves])
    rR = np.full((M, n), np.nan)
    depth = np.full((M, n), np.nan)
    cause = np.full((M, n, 4), np.nan)
    route = np.full((M, n), np.nan)
    g_rR, g_depth, g_cause = np.full(M, np.nan), np.full(M, np.nan), np.full((M, 4), np.nan)
    for k, m in enumerate(moves):
        assert m["surv_calls"] == n + 1
        for e, v in m["rR"].items():
            if e == 0:
                g_rR[k] = v
            else:
                rR[k, e - 1] = v
        for e, v in m["depth"].items():
            if e == 0:
                g_depth[k] = v
            else:
                depth[k, e - 1] = v
        for e, v in m["cause"].items():
            if e == 0:
                g_cause[k, :len(v)] = v
            else:
                cause[k, e - 1, :len(v)] = v
        for e, v in m["route"].items():
            if e > 0:
                route[k, e - 1] = v
    return O.Draws(lengths_ft=cell_lengths(cell, pop_row), dice=dice, rR=rR, depth_u=depth, cause=cause, route=route,
                   ghost_rR=g_rR, ghost_depth=g_depth, ghost_cause=g_cause)


def pack_cell(cell, pop_row):
    """Flat dict of arrays for ``np.savez`` (tests/golden/*.npz)."""
    d = draws_from_recording(cell, pop_row)
    return dict(lengths_ft=d.lengths_ft, length_z=cell_normals(cell), dice=d.dice, rR=d.rR, depth=d.depth_u, cause=d.cause, route=d.route,
                ghost_rR=d.ghost_rR, ghost_depth=d.ghost_depth, ghost_cause=d.ghost_cause)


def unpack_cell(z, prefix=""):
    g = lambda k: z[prefix + k]
    return O.Draws(lengths_ft=g("lengths_ft"), dice=g("dice"), rR=g("rR"), depth_u=g("depth"), cause=g("cause"),
                   route=g("route"), ghost_rR=g("ghost_rR"), ghost_depth=g("ghost_depth"),
                   ghost_cause=g("ghost_cause"))
