"""Host table builder: turns a ``simulation`` object's DataFrames into the flat arrays of include/stryke_b200.h.

Built once per run (facility, species) and once per (scenario, day) (routing CDFs, unit coefficients); the CUDA
kernel only ever sees these tables.  Reference semantics (file:line relative to /root/reference):
  create_route Stryke/stryke.py:1581-1655, unit dictionaries :2714-2800, per-day Q_dict :2930-2977,
  movement :1823-2065, Kaplan/Propeller/Francis :846-1094, species rows :2858-2890.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import pandas as pd

from . import _abi, routing

G_FT = 32.2   # stryke.py:894


def _isnan(x):
    return x is None or (isinstance(x, float) and math.isnan(x)) or (not isinstance(x, str) and pd.isna(x))


@dataclass
class FacilityTables:
    node_names: list
    unit_names: list
    n_moves: int
    start_node: int
    barotrauma: bool
    node_kind: np.ndarray
    node_apriori: np.ndarray
    node_unit: np.ndarray
    node_has_U: np.ndarray
    edge_off: np.ndarray
    edge_dst: np.ndarray
    move_lex_rank: np.ndarray
    level_off: np.ndarray
    pair_node: np.ndarray
    pair_move: np.ndarray
    unit_static: np.ndarray
    succ_order: list = field(default_factory=list)    # per node: successor names in routing order

    @property
    def n_nodes(self):
        return len(self.node_names)

    @property
    def n_units(self):
        return len(self.unit_names)

    @property
    def n_edges(self):
        return int(self.edge_dst.shape[0])

    @property
    def n_pairs(self):
        return int(self.pair_node.shape[0])

    def c_struct(self):
        f = _abi.Facility()
        f.n_nodes, f.n_units, f.n_moves, f.n_edges = self.n_nodes, self.n_units, self.n_moves, self.n_edges
        f.n_pairs, f.start_node, f.barotrauma = self.n_pairs, self.start_node, int(self.barotrauma)
        for k in ("node_kind", "node_apriori", "node_unit", "node_has_U", "edge_off", "edge_dst", "move_lex_rank",
                  "level_off", "pair_node", "unit_static"):
            setattr(f, k, _abi.ptr(getattr(self, k)))
        return f


@dataclass
class DayTables:
    curr_Q: np.ndarray          # (n_days,)
    edge_cdf: np.ndarray        # (n_days, n_edges)
    node_stay: np.ndarray       # (n_days, n_nodes) uint8
    unit_coef: np.ndarray       # (n_days, n_units, 8)
    unit_Q: np.ndarray          # (n_days, n_units) allocated flow (Unit_Hours / Route_Flows)
    edge_flow: np.ndarray       # (n_days, n_edges) routed flow per successor (Route_Flows)

    def c_struct(self):
        d = _abi.DayTables()
        d.n_days = int(self.curr_Q.shape[0])
        for k in ("curr_Q", "edge_cdf", "node_stay", "unit_coef"):
            setattr(d, k, _abi.ptr(getattr(self, k)))
        return d


class RunContext:
    """The per-unit dictionaries ``run()`` builds before its loops (stryke.py:2661-2800)."""

    def __init__(self, sim):
        self.sim = sim
        up = sim.unit_params
        self.units = list(up.index)
        self.unit_fac = {}
        self.q_cap = {}
        self.op_order = {}
        self.intake_vel = {}
        self.penstock_id = {}
        pen_units, pen_caps = {}, {}
        for unit, row in up.iterrows():
            fac = row["Facility"]
            self.unit_fac[unit] = fac
            self.q_cap[unit] = row["Qcap"]
            self.op_order[unit] = row["op_order"]
            self.intake_vel[unit] = row["intake_vel"]
            pen = row.get("Penstock_ID", None)
            default_pen = f"{fac}|{row['Unit'] if 'Unit' in row.index else unit}"
            if _isnan(pen):
                pen = default_pen
            pen = str(pen).strip() or default_pen
            self.penstock_id[unit] = pen
            pen_units.setdefault(pen, []).append(unit)
            pq = row.get("Penstock_Qcap", None)
            if not _isnan(pq):
                try:
                    pen_caps.setdefault(pen, []).append(float(pq))
                except (TypeError, ValueError):
                    pass
            ada = float(row["ada"])
            if ada > 1.0:
                raise ValueError(f"Unit {unit} efficiency {ada:.3f} must be a fraction (0-1), not a percent.")
        self.penstock_cap = {pen: float(max(pen_caps[pen]) if pen_caps.get(pen) else
                                        sum(float(self.q_cap.get(u, 0.0) or 0.0) for u in members))
                             for pen, members in pen_units.items()}
        self.sta_cap = {}
        for unit in self.units:
            self.sta_cap[self.unit_fac[unit]] = self.sta_cap.get(self.unit_fac[unit], 0) + self.q_cap[unit]

    def q_dict(self, curr_Q, scenario):
        """stryke.py:2930-2962."""
        cache = self.__dict__.setdefault("_q_static", {})
        if scenario not in cache:          # everything but curr_Q depends on the scenario only
            fp = self.sim.facility_params
            rows = fp[fp.Scenario == scenario] if "Scenario" in fp.columns else fp
            base = {"min_Q": {}, "env_Q": {}, "bypass_Q": {}}
            for fac, row in rows.iterrows():
                base["min_Q"][fac] = row["Min_Op_Flow"]
                base["env_Q"][fac] = row["Env_Flow"]
                base["bypass_Q"][fac] = row["Bypass_Flow"]
            cache[scenario] = base
        base = cache[scenario]
        Q = {"curr_Q": curr_Q, "min_Q": dict(base["min_Q"]), "env_Q": dict(base["env_Q"]), "bypass_Q": dict(base["bypass_Q"]),
             "sta_cap": dict(self.sta_cap)}
        for u in self.units:
            Q[u] = self.q_cap[u]
        return Q

    def allocations(self, Q):
        """All-units allocation that sets ``u_param_dict[u]['Q']`` (stryke.py:2964-2977)."""
        alloc, _, _, _ = routing.unit_flow_allocations(
            Q["curr_Q"], self.units, Q["min_Q"], Q["env_Q"], Q["bypass_Q"], Q["sta_cap"], self.unit_fac,
            self.penstock_id, self.penstock_cap, self.q_cap, unit_params=self.sim.unit_params, meta=self._meta())
        return {u: float(alloc.get(u, 0.0) or 0.0) for u in self.units}

    def _meta(self):
        if "_unit_meta" not in self.__dict__:
            self._unit_meta = routing.unit_meta(self.sim.unit_params)
            self._served = {}
        return self._unit_meta

    def route(self, location, Q):
        g = self.sim.graph
        nb = list(g.neighbors(location)) if location in g else []
        meta = self._meta()
        return routing.route_table(location, nb, lambda n: g[location][n].get("weight", 1.0), Q, self.op_order,
                                   self.q_cap, self.unit_fac, self.penstock_id, self.penstock_cap,
                                   unit_params=self.sim.unit_params, facility_params=self.sim.facility_params,
                                   meta=meta, served_cache=self._served)


def barotrauma_required(unit_params):
    """stryke.py:2672-2710 incl. the fail-loud validation messages."""
    if unit_params is None:
        return False
    cols = [c for c in ("ps_D", "ps_length", "fb_depth", "submergence_depth", "elevation_head") if c in unit_params.columns]
    if not (cols and unit_params[cols].notna().any().any()):
        return False
    if "fb_depth" not in unit_params.columns:
        raise ValueError("Barotrauma requires columns ['fb_depth'] in Unit_Parameters.")
    tail_cols = [c for c in ("elevation_head", "submergence_depth") if c in unit_params.columns]
    if not any(unit_params[c].notna().any() for c in tail_cols):
        raise ValueError("Barotrauma requires elevation_head or submergence_depth in Unit_Parameters.")
    missing = unit_params[unit_params[["fb_depth"]].isna().any(axis=1)].index.tolist()
    if missing:
        raise ValueError(f"Barotrauma requires fb_depth for all units. Missing: {missing}")
    missing = unit_params[unit_params[tail_cols].isna().all(axis=1)].index.tolist()
    if missing:
        raise ValueError("Barotrauma requires elevation_head or submergence_depth for all units. "
                         f"Missing: {missing}")
    return True


def routing_order(neighbours, op_order):
    """Order in which ``movement`` lists the successors (stryke.py:1887-1899 when a unit is among them, insertion
    order otherwise)."""
    neighbours = [str(n) for n in neighbours]
    if any("U" in n for n in neighbours):
        return sorted(neighbours, key=lambda n: (routing._ROLE_RANK[routing._role(n)],
                                                 op_order.get(n, 999) if routing._role(n) == "unit" else 0))
    return neighbours


def build_facility(sim, ctx: RunContext) -> FacilityTables:
    names = [str(x) for x in sim.nodes.Location.values]
    idx = {n: i for i, n in enumerate(names)}
    n_nodes = len(names)
    if n_nodes > _abi.MAX_NODES or len(ctx.units) > _abi.MAX_UNITS:
        raise ValueError(f"facility too large for the compiled kernel limits ({n_nodes} nodes, {len(ctx.units)} units)")
    unit_idx = {u: i for i, u in enumerate(ctx.units)}
    kind = np.zeros(n_nodes, np.uint8)
    apriori = np.zeros(n_nodes, np.float64)
    node_unit = np.full(n_nodes, -1, np.int16)
    surv = dict(zip(names, sim.nodes.Survival.values)) if "Survival" in sim.nodes.columns else {}
    for i, n in enumerate(names):
        fun = sim.surv_fun_dict[n]["Surv_Fun"]
        if fun == "a priori":
            val = surv.get(n)
            if _isnan(val):
                raise ValueError(f"A priori survival for route '{n}' is empty or NaN.")
            apriori[i] = float(val)
        else:
            if fun not in _abi.KIND:
                raise KeyError(f"Unknown survival function '{fun}' at node '{n}'")
            if n not in unit_idx:
                raise KeyError(n)   # u_param_dict[route], stryke.py:1410
            kind[i] = _abi.KIND[fun]
            node_unit[i] = unit_idx[n]
    for u in ctx.units:   # stryke.py:2749-2787 builds no param_dict for other runner types (Appendix E8)
        rt = sim.unit_params.at[u, "Runner Type"]
        pump_ok = rt == "Pump" and _has(sim.unit_params.loc[u], "Q_p") and _has(sim.unit_params.loc[u], "gamma")      # opt-in, see unit_coefficients
        if u in idx and kind[idx[u]] and rt not in ("Kaplan", "Propeller", "Francis") and not pump_ok:
            raise KeyError(u)
    has_U = np.array([1 if "U" in n else 0 for n in names], np.uint8)

    g = sim.graph
    edge_off = np.zeros(n_nodes + 1, np.int32)
    edge_dst, succ_order = [], []
    for i, n in enumerate(names):
        nb = routing_order(list(g.neighbors(n)) if (n_nodes > 1 and n in g) else [], ctx.op_order)
        for m in nb:
            if m not in idx:
                raise KeyError(f"Missing survival functions for states: ['{m}']")   # stryke.py:3178-3185
        succ_order.append(nb)
        edge_dst += [idx[m] for m in nb]
        edge_off[i + 1] = len(edge_dst)
    n_moves = int(len(sim.moves))
    if n_moves > _abi.MAX_MOVES or len(edge_dst) > _abi.MAX_EDGES:
        raise ValueError("facility too large for the compiled kernel limits (moves/edges)")
    start = idx["river_node_0"] if n_nodes > 1 else 0
    order = sorted(range(n_moves), key=lambda k: f"state_{k}")     # stryke.py:3272
    lex = np.zeros(n_moves, np.uint8)
    for rank, k in enumerate(order):
        lex[k] = rank
    baro = barotrauma_required(sim.unit_params)
    ustat = np.zeros((len(ctx.units), _abi.UNIT_STATIC_W), np.float64)
    up = sim.unit_params
    for j, u in enumerate(ctx.units):
        rack = sim.facility_params.at[ctx.unit_fac[u], "Rack Spacing"]
        if isinstance(rack, pd.Series):
            rack = rack.iloc[0]
        ustat[j, 0] = float(rack)
        ustat[j, 1] = float(up.at[u, "intake_vel"])
        if baro:
            ustat[j, 2] = float(up.at[u, "fb_depth"])
            tail = up.at[u, "elevation_head"] if "elevation_head" in up.columns else None
            if (_isnan(tail) or getattr(sim, "barotrauma_pressure", "hydrostatic") == "friction") and "submergence_depth" in up.columns:
                tail = up.at[u, "submergence_depth"]      # (friction-loss mode: P2 = atm + rho g h_D, h_D the draft-tube submergence)
            if _isnan(tail):
                raise ValueError(f"Barotrauma requires elevation_head or submergence_depth for route {u}.")
            ustat[j, 3] = float(tail)
    return FacilityTables(names, list(ctx.units), n_moves, start, baro, kind, apriori, node_unit, has_U, edge_off,
                          np.array(edge_dst, np.int16), lex, np.zeros(n_moves + 1, np.int32),
                          np.zeros(0, np.int16), np.zeros(0, np.int16), ustat, succ_order)


def finalize_levels(fac: FacilityTables, days: "DayTables"):
    """Reachable (move, node) pairs = the dense slots of the State_Daily tally.  A live fish at node v is at one of
    v's successors at the next move, or still at v when it "stays put" that day (no successors, curr_Q <= 0 or zero
    routed flow: stryke.py:1849-1855, :2018).  Level sets are propagated per distinct stay pattern and united."""
    patterns = np.unique(days.node_stay, axis=0) if days.node_stay.shape[0] else np.zeros((1, fac.n_nodes), np.uint8)
    levels = [set() for _ in range(fac.n_moves)]
    for stay in patterns:
        cur = {fac.start_node}
        for k in range(fac.n_moves):
            levels[k] |= cur
            nxt = set()
            for v in cur:
                succ = fac.edge_dst[fac.edge_off[v]:fac.edge_off[v + 1]]
                if stay[v] or len(succ) == 0:
                    nxt.add(v)
                else:
                    nxt.update(int(x) for x in succ)
            cur = nxt
    level_off, pair_node, pair_move = [0], [], []
    for k in range(fac.n_moves):
        for v in sorted(levels[k]):
            pair_node.append(v)
            pair_move.append(k)
        level_off.append(len(pair_node))
    if len(pair_node) > _abi.MAX_PAIRS:
        raise ValueError(f"{len(pair_node)} reachable (move, node) pairs exceed the compiled kernel limit {_abi.MAX_PAIRS}")
    fac.level_off = np.array(level_off, np.int32)
    fac.pair_node = np.array(pair_node, np.int16)
    fac.pair_move = np.array(pair_move, np.int16)
    return fac


def _has(row, col):
    return col in row.index and not _isnan(row[col])


RHO, G_SI, P_ATM = 998.2, 9.80665, 101325.0      # stryke.py:1471-1473 (scipy.constants.g, atm)


def friction_pressure_offset(row, Q):
    """Friction-loss pressure mode (opt-in, ``simulation.barotrauma_pressure = "friction"``): the day's offset c0 of
    P1 / P2 = c0 + c1 * fish_depth_ft for one unit.  Restates the composition of Stryke/stryke_barotrauma.py:565-601 over
    the helpers of Stryke/barotrauma.py:11-137 (the reference does not wire it into run()): imperial inputs are converted
    to SI, v = Q / A in the penstock, Darcy-Weisbach head loss with the friction factor of ``calc_friction``,
    P2 = atm + rho g h_D (h_D = submergence_depth), P1 = P2 + rho g (h_fish - h_2) + rho g h_loss (h_2 = elevation_head,
    0 when empty; the velocity heads cancel, v_1 = v_2).  So c1 = rho g 0.3048 / P2 and c0 = 1 + rho g (h_loss - h_2) / P2."""
    for c in ("ps_D", "ps_length", "roughness", "submergence_depth"):
        if not _has(row, c):
            raise ValueError(f"Friction-loss barotrauma pressure requires {c} for unit {row.name}.")
    ps_d = float(row["ps_D"]) * 0.3048
    ps_l = float(row["ps_length"]) * 0.3048
    a = np.pi * (float(row["ps_D"]) / 2.0) ** 2 * 0.092903
    q = float(Q) * 0.02831683199881
    h_D = float(row["submergence_depth"]) * 0.3048
    h_2 = (float(row["elevation_head"]) if _has(row, "elevation_head") else 0.0) * 0.3048
    v = np.float64(q / a)
    k_v = 0.0010016 / RHO
    with np.errstate(divide="ignore", invalid="ignore"):
        Re = np.float64(v * ps_d / k_v)          # (NumPy scalars: a unit without flow gives Re = 0 -> f = 0, no exception)
        f = 0.25 / (np.log((float(row["roughness"]) / ps_d) / (3.7 * ps_d) + (5.74 / Re ** 0.9))) ** 2
    h_loss = f * (ps_l / ps_d) * ((v * v) / (2 * G_SI))
    if not np.isfinite(h_loss):
        h_loss = 0.0
    p2 = P_ATM + RHO * G_SI * h_D
    return 1.0 + RHO * G_SI * (h_loss - h_2) / p2


def unit_coefficients(row, Q):
    """Per-(day, unit) coefficient row (layout: include/stryke_b200.h).  The day's flow ``Q`` is the all-units
    allocation (stryke.py:2977); formulas follow Kaplan/Propeller/Francis (stryke.py:907-918, :993-1006, :1078-1091)."""
    out = np.zeros(_abi.UNIT_COEF_W)
    rt = row["Runner Type"]
    H, RPM, D = float(row["H"]), float(row["RPM"]), float(row["D"])
    ada, N, lam = float(row["ada"]), float(row["N"]), float(row["lambda"])
    omega = RPM * ((2 * np.pi) / 60)
    Ewd = (G_FT * H) / (omega * D) ** 2
    out[0], out[1], out[2] = lam, N, D
    out[6] = 1.0 if Q > 0 else 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        Qwd = np.float64(Q) / (omega * D ** 3)
        if rt == "Kaplan":
            out[3] = np.pi * ada * Ewd
            out[4] = 2 * Qwd
            out[5] = 8 * Qwd
        elif rt == "Propeller":
            rR = 0.75
            Qwd_opt = float(row["Qopt"]) / (omega * D ** 3)
            beta = np.arctan((np.pi / 8 * rR) / Qwd_opt)
            a_a = np.arctan((np.pi * Ewd * ada) / (2 * Qwd * rR) + (np.pi / 8 * rR) / Qwd - np.tan(beta))
            out[3] = (np.cos(a_a) / (8 * Qwd)) + np.sin(a_a) / (np.pi * rR)
        elif rt == "Francis":
            B, D1, D2, iota = float(row["B"]), float(row["D1"]), float(row["D2"]), float(row["iota"])
            Qper = float(row["Qopt"] / row["Qcap"])
            beta = np.arctan((0.707 * (np.pi / 8)) / (iota * Qper * (D1 / D2) ** 3))
            alpha = np.arctan((2 * np.pi * Ewd * ada) / Qwd * (B / D1)
                              + (np.pi * 0.707 ** 2) / (2 * Qwd) * (B / D1) * (D2 / D1) ** 2
                              - 4 * 0.707 * np.tan(beta) * (B / D1) * (D1 / D2))
            out[3] = (np.sin(alpha) * (B / D1)) / (2 * Qwd) + np.cos(alpha) / np.pi
        elif rt == "Pump" and _has(row, "Q_p") and _has(row, "gamma"):
            # opt-in: the reference's Pump formula (stryke.py:1096-1127) needs Q_p and gamma, which no sheet of the reference
            # carries — its run() stops with KeyError at a Pump node (:2749-2787, :2977).  With both columns present the node
            # runs: p = gamma * (N L / (0.707 D2)) * (sin(beta_p) (B / D1) / (2 Qpwd) + cos(beta_p) / pi); the pumped
            # discharge Q_p, not the day's allocation, enters.
            B, D1, D2 = float(row["B"]), float(row["D1"]), float(row["D2"])
            Qpwd = float(row["Q_p"]) / (omega * D ** 3)
            beta_p = np.arctan((0.707 * np.pi / 8) / (Qpwd * np.power(D1 / D2, 3)))
            out[0], out[2], out[6] = float(row["gamma"]), 0.707 * D2, 1.0
            out[3] = ((np.sin(beta_p) * (B / D1)) / (2 * Qpwd)) + (np.cos(beta_p) / np.pi)
        else:
            raise KeyError(rt)
    if rt not in ("Kaplan", "Pump") and not np.isfinite(out[3]):
        out[3] = np.inf      # reference: ZeroDivisionError for a fish at a unit with no flow (Appendix E10)
    return out


def build_days(sim, ctx: RunContext, fac: FacilityTables, day_list):
    """``day_list`` = [(scenario, curr_Q), ...] -> DayTables (one row per entry; identical discharges share work)."""
    n_days = len(day_list)
    E, U, Nn = fac.n_edges, fac.n_units, fac.n_nodes
    curr = np.zeros(n_days)
    cdf = np.zeros((n_days, E))
    stay = np.zeros((n_days, Nn), np.uint8)
    coef = np.zeros((n_days, U, _abi.UNIT_COEF_W))
    uQ = np.zeros((n_days, U))
    eflow = np.zeros((n_days, E))
    cache = {}
    rows = [sim.unit_params.loc[u] for u in fac.unit_names]
    for d, (scenario, Qd) in enumerate(day_list):
        Qd = float(Qd)
        key = (scenario, Qd)
        if key not in cache:
            Q = ctx.q_dict(Qd, scenario)
            alloc = ctx.allocations(Q)
            c_coef = np.stack([unit_coefficients(rows[j], alloc[u]) for j, u in enumerate(fac.unit_names)]) if U else np.zeros((0, _abi.UNIT_COEF_W))
            if U and fac.barotrauma and getattr(sim, "barotrauma_pressure", "hydrostatic") == "friction":
                for j, u in enumerate(fac.unit_names):
                    c_coef[j, 7] = friction_pressure_offset(rows[j], alloc[u])
            c_cdf, c_flow, c_stay = np.ones(E), np.zeros(E), np.zeros(Nn, np.uint8)
            for i, n in enumerate(fac.node_names):
                lo, hi = fac.edge_off[i], fac.edge_off[i + 1]
                rt = ctx.route(n, Q) if Nn > 1 else None
                if rt is None:
                    c_stay[i] = 1
                    continue
                locs, probs, flows = rt
                assert list(locs) == fac.succ_order[i], (n, locs, fac.succ_order[i])
                c_cdf[lo:hi] = routing.normalised_cdf(probs)
                c_flow[lo:hi] = flows
            cache[key] = (c_coef, c_cdf, c_flow, c_stay, np.array([alloc[u] for u in fac.unit_names]))
        c_coef, c_cdf, c_flow, c_stay, c_q = cache[key]
        curr[d] = Qd
        coef[d], cdf[d], eflow[d], stay[d], uQ[d] = c_coef, c_cdf, c_flow, c_stay, c_q
    return DayTables(curr, cdf, stay, coef, uQ, eflow)


HABITAT = {"Pelagic": (0.01, 0.33), "Benthic": (0.8, 1.0)}   # stryke.py:1476-1484


def species_row(sim, spc_dat):
    """One Population row -> params[STRYKE_SPECIES_W] (stryke.py:2865-2890, :3028-3029, :3068-3090)."""
    g = lambda c: spc_dat.iat[0, spc_dat.columns.get_loc(c)] if c in spc_dat.columns else float("nan")
    p = np.zeros(_abi.SPECIES_W)
    s, lloc, lscale = g("length shape"), g("length location"), g("length scale")
    mean_len, sd_len = g("Length_mean"), g("Length_sd")
    if not _isnan(s):
        p[0], p[1], p[2], p[5] = float(s), float(lloc), float(lscale), 0.0
    else:
        if _isnan(mean_len) or _isnan(sd_len):
            raise ValueError(f"Normal population draw requires mean_len/sd_len for species '{g('Species')}'.")
        if sd_len < 0:
            raise ValueError(f"Normal population draw has negative sd_len={sd_len} for species '{g('Species')}'.")
        p[3], p[4], p[5] = float(mean_len), float(sd_len), 1.0
    common = g("Common Name") if "Common Name" in spc_dat.columns else None
    family = g("Family") if "Family" in spc_dat.columns else None
    p[6] = sim._resolve_fish_width_ratio(species_name=g("Species"), common_name=None if _isnan(common) else common,
                                         family_name=None if _isnan(family) else family)
    if "U_crit" not in spc_dat.columns:
        raise KeyError(f"U_crit missing in population data for species '{g('Species')}'.")
    if _isnan(g("U_crit")):
        raise ValueError(f"U_crit is NaN for species '{g('Species')}'.")
    p[7] = float(g("U_crit"))
    d1, d2 = HABITAT.get(g("vertical_habitat"), (0.01, 1.0))
    p[8], p[9] = d1, d2
    b0, b1 = g("beta_0"), g("beta_1")
    p[10] = 0.0 if _isnan(b0) else float(b0)
    p[11] = 0.0 if _isnan(b1) else float(b1)
    shape = g("shape")
    if _isnan(shape):
        p[12] = _abi.ENT["fixed"]
        p[18] = float(int(g("Fish")))
    else:
        dist = g("dist")
        p[12] = _abi.ENT.get(dist, _abi.ENT["Weibull"])    # any other label falls to weibull_min (stryke.py:2564)
        p[13], p[14], p[15] = float(shape), float(g("location")), float(g("scale"))
    mr = g("max_ent_rate")
    try:
        p[16] = float(mr)
    except (TypeError, ValueError):
        p[16] = float("nan")
    occ = g("occur_prob")
    p[17] = 1.0 if _isnan(occ) else float(occ)     # stryke.py:2888-2890
    return p
