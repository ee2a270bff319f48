"""Host-side flow allocation and routing-probability rules.

These are the per-(scenario, day) table builders of the passage engine: the reference recomputes them for every
fish inside ``movement`` (82 % of its run time); here they run once per day on the host and the
resulting cumulative transition probabilities are staged in shared memory by the CUDA kernel.

Semantics follow the reference exactly (file:line relative to /root/reference):
  * ``unit_flow_allocations``  — ``simulation.compute_unit_flow_allocations`` Stryke/stryke.py:1657-1821
  * ``route_table``            — the probability half of ``simulation.movement`` Stryke/stryke.py:1841-2056
  * ``run_of_river_hours`` / ``peaking_hours`` — ``simulation.daily_hours`` Stryke/stryke.py:2309-2503
Node roles are decided by substrings of the node NAME ('U' = unit, 'env', 'bypass', 'spill'), as the reference does
(SURVEY.md Appendix B, E13).
"""
from __future__ import annotations

import math

import numpy as np


def _num(val, default=0.0):
    """``safe_float`` of stryke.py:1668-1674."""
    try:
        if val is None or (isinstance(val, float) and math.isnan(val)):
            return default
        return float(val)
    except (TypeError, ValueError):
        return default


def _missing(x):
    return x is None or (isinstance(x, float) and math.isnan(x))


def penstock_order(pen_id):
    """Sort key for penstock ids: numeric ids first, then ids with digits, then names (stryke.py:1676-1686)."""
    text = "" if pen_id is None else str(pen_id).strip()
    if text.isdigit():
        return (0, int(text))
    try:
        return (0, float(text))
    except (TypeError, ValueError):
        digits = "".join(ch for ch in text if ch.isdigit())
        return (1, int(digits)) if digits else (2, text.lower())


def fill_units_in_order(flow, ordered_units, qopt_of, qcap_of, min_start):
    """Sequential unit loading inside one penstock (``allocate_units_sequential``, stryke.py:1688-1731):
    each unit in operating order takes min(remaining, min(Qopt or Qcap, Qcap)) subject to the minimum start flow;
    whatever is left then tops running units up to Qcap."""
    given = dict.fromkeys(ordered_units, 0.0)
    left = float(flow or 0.0)
    if left <= 0.0:
        return given
    start = _num(min_start, 0.0)
    gated = start > 0.0
    for u in ordered_units:
        if left <= 0.0 or (gated and left < start):
            break
        cap = _num(qcap_of.get(u, 0.0))
        if cap <= 0.0 or (gated and cap < start):
            continue
        opt = _num(qopt_of.get(u, 0.0))
        want = min(opt if opt > 0.0 else cap, cap)
        if gated and want < start:
            want = start
        take = min(left, want)
        if gated and take < start:
            break
        given[u] = take
        left -= take
    if left > 0.0:
        for u in ordered_units:
            if left <= 0.0:
                break
            cap = _num(qcap_of.get(u, 0.0))
            have = given.get(u, 0.0)
            if cap <= 0.0 or have <= 0.0:
                continue
            room = max(cap - have, 0.0)
            if room <= 0.0:
                continue
            extra = min(left, room)
            given[u] = have + extra
            left -= extra
    return given


def unit_meta(unit_params):
    """Per-unit lookups of ``unit_flow_allocations`` taken out of the frame once: {unit: (Facility, op_order | None,
    Qopt | None)} — the tables of a run are built day by day, and cell-by-cell ``DataFrame.at`` was half their time."""
    if unit_params is None:
        return None
    has_order, has_qopt = "op_order" in unit_params.columns, "Qopt" in unit_params.columns
    out = {}
    for unit in unit_params.index:
        order = None
        if has_order:
            val = unit_params.at[unit, "op_order"]
            if not (isinstance(val, float) and math.isnan(val)):
                order = int(val)
        out[unit] = (unit_params.at[unit, "Facility"] if "Facility" in unit_params.columns else None, order,
                     _num(unit_params.at[unit, "Qopt"]) if has_qopt else None)
    return out


def unit_flow_allocations(curr_Q, unit_nodes, min_Q_dict, env_Q_dict, bypass_Q_dict, sta_cap_dict, unit_fac_dict,
                          penstock_id_dict, penstock_cap_dict, cap_dict, unit_params=None, meta=None):
    """Flow through every unit in ``unit_nodes`` at river discharge ``curr_Q``.

    Returns ``(allocations, prod_Q_by_facility, total_env_Q, total_bypass_Q)`` exactly like
    ``simulation.compute_unit_flow_allocations`` (stryke.py:1657-1821).  ``meta`` = ``unit_meta(unit_params)`` when the
    caller has it (same values, no frame lookups)."""
    if meta is None:
        meta = unit_meta(unit_params)

    def facility_of(unit):
        fac = unit_fac_dict.get(unit)
        if fac is None and meta is not None and unit in meta:
            fac = meta[unit][0]
        return fac

    facilities = []
    for unit in unit_nodes:
        fac = facility_of(unit)
        if fac is not None and fac not in facilities:
            facilities.append(fac)
    total_env = sum(env_Q_dict.get(f, 0.0) for f in set(facilities))
    total_bypass = sum(bypass_Q_dict.get(f, 0.0) for f in set(facilities))

    production = {}
    for fac in facilities:
        if curr_Q > min_Q_dict.get(fac, 0.0):
            usable = curr_Q - env_Q_dict.get(fac, 0.0) - bypass_Q_dict.get(fac, 0.0)
            production[fac] = min(max(usable, 0.0), sta_cap_dict.get(fac, 0.0))
        else:
            production[fac] = 0.0

    order_of, qopt_of = {}, {}
    if meta is not None:
        for unit in unit_nodes:
            if unit not in meta:
                continue
            _, order, qopt = meta[unit]
            if order is not None:
                order_of[unit] = order
            if qopt is not None:
                qopt_of[unit] = qopt

    by_facility = {}
    for unit in unit_nodes:
        fac = facility_of(unit)
        if fac is None:
            continue
        pen = penstock_id_dict.get(unit) if penstock_id_dict else None
        if not pen:
            pen = f"{fac}|{unit}"
        by_facility.setdefault(fac, {}).setdefault(pen, []).append(unit)

    allocations = dict.fromkeys(unit_nodes, 0.0)
    for fac, penstocks in by_facility.items():
        left = production.get(fac, 0.0)
        start = _num(min_Q_dict.get(fac, 0.0))
        for pen in sorted(penstocks, key=penstock_order):
            if left <= 0.0:
                break
            members = penstocks[pen]
            cap = penstock_cap_dict.get(pen) if penstock_cap_dict is not None else None
            if _missing(cap):
                cap = sum(_num(cap_dict.get(u, 0.0)) for u in members)
            cap = _num(cap)
            through = min(left, cap) if cap > 0.0 else left
            left -= through
            ranked = sorted(members, key=lambda u: (order_of.get(u, 999), u))
            allocations.update(fill_units_in_order(through, ranked, qopt_of, cap_dict, start))
    return allocations, production, total_env, total_bypass


def _role(name):
    """Neighbour role used by the units branch of ``movement`` (stryke.py:1887-1897, :1923-1953)."""
    low = str(name).lower()
    if "U" in name:
        return "unit"
    if "env" in low:
        return "env"
    if "bypass" in low:
        return "bypass"
    if "spill" in low:
        return "spill"
    return "other"


_ROLE_RANK = {"unit": 0, "env": 1, "bypass": 2, "other": 3, "spill": 4}


def route_table(location, neighbours, edge_weight, Q_dict, op_order, cap_dict, unit_fac_dict, penstock_id_dict=None,
                penstock_cap_dict=None, unit_params=None, facility_params=None, meta=None, served_cache=None):
    """Routing decision table of a live fish standing at ``location``.

    ``neighbours`` is ``list(graph.neighbors(location))`` (insertion order) and ``edge_weight(n)`` the edge's
    ``weight`` attribute (default 1.0).  Returns ``None`` when the fish stays put (stryke.py:1849-1855, :2018-2019)
    or ``(locs, probs, route_flows)`` exactly as handed to ``np.random.choice`` at stryke.py:2058."""
    curr_Q = Q_dict["curr_Q"]
    if curr_Q <= 0.0 or len(neighbours) == 0:
        return None
    neighbours = [str(n) for n in neighbours]
    min_Q, sta_cap = Q_dict["min_Q"], Q_dict["sta_cap"]
    env_Q, bypass_Q = Q_dict["env_Q"], Q_dict["bypass_Q"]
    roles = [_role(n) for n in neighbours]
    locs, flows = [], []

    if "unit" in roles:   # units branch, stryke.py:1867-1953
        units_here = [n for n, r in zip(neighbours, roles) if r == "unit"]
        alloc, _, env_total, bypass_total = unit_flow_allocations(
            curr_Q, units_here, min_Q, env_Q, bypass_Q, sta_cap, unit_fac_dict, penstock_id_dict, penstock_cap_dict,
            cap_dict, unit_params=unit_params, meta=meta)
        ranked = sorted(neighbours, key=lambda n: (_ROLE_RANK[_role(n)], op_order.get(n, 999) if _role(n) == "unit" else 0))
        n_env = sum(1 for n in ranked if "env" in n.lower())
        n_bypass = sum(1 for n in ranked if "bypass" in n.lower())
        env_share = env_total if n_env else 0.0
        bypass_share = bypass_total if n_bypass else 0.0
        through_units = sum(alloc.values())
        for n in ranked:
            role = _role(n)
            if role == "unit":
                q = float(alloc.get(n, 0.0) or 0.0)
            elif role == "env":
                q = env_share / n_env if n_env else 0.0
            elif role == "bypass":
                q = bypass_share / n_bypass if n_bypass else 0.0
            elif role == "spill":
                q = max(0.0, curr_Q - through_units - bypass_share - env_share)
            else:
                q = bypass_Q.get(unit_fac_dict.get(n, None), 0.0)
            locs.append(n)
            flows.append(q)
    elif any("spill" in n.lower() for n in neighbours):   # spill branch, stryke.py:1955-2000
        key = tuple(neighbours)
        if served_cache is not None and key in served_cache:
            served = served_cache[key]
        else:
            served = list(facility_params[facility_params.Spillway.isin(neighbours)].index) if facility_params is not None else []
            if served_cache is not None:
                served_cache[key] = served
        cap_total = sum(sta_cap.get(f, 0.0) for f in served)
        env_total = sum(env_Q.get(f, 0.0) for f in served)
        bypass_total = sum(bypass_Q.get(f, 0.0) for f in served)
        n_env = sum(1 for n in neighbours if "env" in n.lower())
        n_bypass = sum(1 for n in neighbours if "bypass" in n.lower())
        env_share = env_total if n_env else 0.0
        bypass_share = bypass_total if n_bypass else 0.0
        for n in neighbours:
            low = n.lower()
            if "env" in low:
                q = env_share / n_env if n_env else 0.0
            elif "bypass" in low:
                q = bypass_share / n_bypass if n_bypass else 0.0
            elif "spill" in low:
                if curr_Q >= cap_total + env_share + bypass_share:
                    q = curr_Q - cap_total - env_share - bypass_share
                else:
                    q = max(0.0, curr_Q - env_share - bypass_share)
                q = max(q, 0.0)
            else:
                q = 1.0 * curr_Q     # "unknown" neighbour: p = 1, flow = Q (stryke.py:1994-1996)
            locs.append(n)
            flows.append(q)
    else:   # edge weights, stryke.py:2002-2008
        for n in neighbours:
            locs.append(n)
            flows.append(edge_weight(n) * curr_Q)

    flows = np.asarray(flows, dtype=float)
    routed = float(flows.sum())
    if routed <= 0.0:
        return None
    if routed < curr_Q:   # deficit goes to the first neighbour whose name contains 'spill' (case-sensitive, :2022)
        for i, n in enumerate(locs):
            if "spill" in n:
                flows[i] += curr_Q - routed
                routed = float(flows.sum())
                break
    probs = flows / curr_Q if abs(routed - curr_Q) <= 1e-6 else flows / routed
    return locs, probs, flows


def normalised_cdf(probs):
    """``np.random.choice`` (legacy RandomState): cdf = cumsum(p); cdf /= cdf[-1]; idx = searchsorted(cdf, u, 'right')."""
    total = float(np.sum(probs))
    if abs(total - 1.0) > math.sqrt(np.finfo(np.float64).eps):
        raise ValueError(f"probabilities do not sum to 1 (sum={total})")
    cdf = np.cumsum(np.asarray(probs, dtype=float))
    return cdf / cdf[-1]
