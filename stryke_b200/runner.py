"""``simulation.run()`` — orchestration of one Monte Carlo run on the GPU (reference: Stryke/stryke.py:2620-3517).

Host work: validation, per-unit dictionaries, hydrographs, daily operating hours, one routing/strike table per
(scenario, day); then ONE call into the CUDA engine for every (scenario, species, iteration, day) cell, and the
assembly of the reference's result frames (Daily, State_Daily, Route_Flows, Unit_Hours) from the tally arrays.
"""
from __future__ import annotations

import logging
import math
import os
from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import _abi, engine, tables
from .store import open_store

logger = logging.getLogger(__name__)

CFS_TO_CMS = 0.02831683199881    # stryke.py:3253


def _dist_rank_world():
    """(rank, world size) of an initialised torch.distributed job, else (0, 1)."""
    import sys
    if "torch" not in sys.modules:       # a job that initialised a process group has imported torch; do not pay for the
        return 0, 1                      # import (seconds) in single-process runs
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _flag(name, default="0"):
    return str(os.environ.get(name, default)).strip().lower() in ("1", "true", "yes", "on")


@dataclass
class Group:
    """All cells of one (scenario, species): ``iterations`` x ``days`` of them."""
    scen_index: int
    scenario: str
    scen_num: object
    season: object
    species: str
    species_index: int
    iterations: int
    day0: int            # first row of this scenario in the day tables
    n_days: int
    cell0: int = 0       # first cell of the group in engine order (day-major, iteration-minor)
    n_active_days: int = 0


class RunPlan:
    """Everything ``run()`` derives on the host before the device call."""

    def __init__(self, sim):
        self.sim = sim
        sim._validate_unit_params_required_fields()
        sim._route_flow_logged_keys = set()
        sim._mortality_components = {"impingement": 0, "blade_strike": 0, "barotrauma": 0, "latent": 0}
        sim.create_route()
        if not hasattr(sim, "moves") or sim.moves is None:
            raise RuntimeError("moves not initialized; create_route() must set moves before run.")
        self.ctx = tables.RunContext(sim)
        self.fac = tables.build_facility(sim, self.ctx)
        sim.barotrauma_required = self.fac.barotrauma
        self.day_rows = []          # (scenario, curr_Q)
        self.day_dates = []
        self.groups = []
        self.species_params = []
        self.scen_days = {}
        self._enumerate()
        self.days = tables.build_days(sim, self.ctx, self.fac, self.day_rows)
        tables.finalize_levels(self.fac, self.days)
        self._hours()
        self._cells()

    # ---- scenario / species enumeration (stryke.py:2806-2897) -----------------------------------------------
    def _enumerate(self):
        sim = self.sim
        total = len(sim.flow_scenarios)
        for si, scen in enumerate(sim.flow_scenarios):
            print(f"[INFO] Processing scenario {si + 1} of {total}: {scen}", flush=True)
            if not hasattr(sim, "flow_scenarios_df") or sim.flow_scenarios_df is None:
                raise ValueError("flow_scenarios_df is missing; cannot process scenarios.")
            sdf = sim.flow_scenarios_df[sim.flow_scenarios_df["Scenario"] == scen]
            if sdf.empty:
                raise ValueError(f"Scenario '{scen}' not found in flow_scenarios_df.")
            col = lambda c: sdf.iat[0, sdf.columns.get_loc(c)]
            scen_num, season, scenario, months = col("Scenario Number"), col("Season"), col("Scenario"), col("Months")
            sim.discharge_type = "hydrograph" if col("Flow") == "hydrograph" else "fixed"
            if isinstance(months, (int, np.integer)):
                months = [int(months)]
            else:
                months = [int(m) for m in str(months).split(",")]
            if not hasattr(sim, "operating_scenarios_df") or sim.operating_scenarios_df is None:
                raise ValueError("operating_scenarios_df is missing; cannot process operating scenarios.")
            if sim.operating_scenarios_df[sim.operating_scenarios_df["Scenario"] == scenario].empty:
                raise ValueError(f"Scenario '{scenario}' not found in operating_scenarios_df.")
            if sim.discharge_type == "hydrograph":
                flow_df = sim.create_hydrograph("hydrograph", scen, months, sim.flow_scenarios_df)
            else:
                flow_df = sim.create_hydrograph("fixed", scen, months, sim.flow_scenarios_df, fixed_discharge=col("Flow"))
            n_days = len(flow_df)
            print(f"[INFO] Scenario '{scenario}' will simulate {n_days} days", flush=True)
            day0 = len(self.day_rows)
            for q, day in zip(flow_df["DAvgFlow_prorate"].values, flow_df["datetimeUTC"].values):
                self.day_rows.append((scenario, float(q)))
                self.day_dates.append(pd.to_datetime(day))
            self.scen_days[scenario] = (day0, n_days)
            for spc in sim.pop.Species.unique():
                dat = sim.pop[(sim.pop["Scenario"] == scenario) & (sim.pop.Species == spc)]
                if dat.empty:
                    continue
                iterations = int(dat.iat[0, dat.columns.get_loc("Iterations")])
                self.species_params.append(tables.species_row(sim, dat))
                self.groups.append(Group(si, scenario, scen_num, season, str(dat.iat[0, dat.columns.get_loc("Species")]),
                                         len(self.species_params) - 1, iterations, day0, n_days))
                print(f"[INFO] Simulating species: {self.groups[-1].species} with {iterations} iterations", flush=True)

    # ---- daily operating hours (stryke.py:2979, :3025) ------------------------------------------------------
    def _hours(self):
        sim = self.sim
        n = len(self.day_rows)
        self.tot_hours = np.zeros(n)
        self.hours = [None] * n
        self.peaking = False
        fp = sim.facility_params
        ops = fp["Operations"].astype(str).values if "Operations" in fp.columns else []
        self.peaking = any(o != "run-of-river" for o in ops)
        cache, ror = {}, {}
        for d, (scenario, q) in enumerate(self.day_rows):
            key = (scenario, q)
            if key not in cache or self.peaking:
                if self.peaking:
                    th, _, hd, _ = sim.daily_hours(self.ctx.q_dict(q, scenario), scenario)
                    cache[key] = (float(np.sum(th)), dict(hd))
                else:
                    if scenario not in ror:
                        ror[scenario] = self._run_of_river_units(scenario)
                    cache[key] = self._run_of_river_hours(ror[scenario], q)
            self.tot_hours[d], self.hours[d] = cache[key]

    # Run-of-river plants (stryke.py:2432-2470): which units get an hours entry depends on the day's flow only through
    # the point where the cumulative capacity passes min(production flow, station capacity).  The unit order, keys and
    # hours come from ONE call of the API mirror per scenario; the per-day part is then a walk over a few floats
    # (identical results: tests/test_host_logic.py compares with simulation.daily_hours day by day).
    def _run_of_river_units(self, scenario):
        sim = self.sim
        probe = self.ctx.q_dict(1.0e30, scenario)          # flow above every capacity: every unit gets its entry
        _, _, hours_dict, _ = sim.daily_hours(probe, scenario)
        ops_df = sim.operating_scenarios_df[sim.operating_scenarios_df.Scenario == scenario]
        order = list(hours_dict.keys())                      # facility by facility, units by op_order
        key_col = sim.unit_params.index.name or "index"
        per_fac = []
        for facility in ops_df.Facility.unique():
            fac_units = sim.unit_params[sim.unit_params.Facility == facility]
            if not fac_units.index.equals(pd.RangeIndex(len(fac_units))):
                fac_units = fac_units.reset_index(drop=fac_units.index.name in fac_units.columns)
            fac_units = fac_units.sort_values(by="op_order")
            units = []
            for i, row in fac_units.iterrows():
                if ops_df[ops_df.Unit == row["Unit"]].empty:
                    continue
                key = row.get(key_col, None) if key_col in fac_units.columns else None
                if key is None or (isinstance(key, float) and np.isnan(key)):
                    key = row.get("Unit", None)
                if key is None or (isinstance(key, float) and np.isnan(key)):
                    key = i
                units.append((str(key), float(row["Qcap"])))
            per_fac.append((facility, probe["env_Q"][facility], probe["bypass_Q"][facility], probe["sta_cap"][facility], units))
        return per_fac, {k: hours_dict[k] for k in order}

    @staticmethod
    def _run_of_river_hours(prepared, q):
        per_fac, all_hours = prepared
        hours = {}
        for facility, env_q, byp_q, sta_cap, units in per_fac:
            limit = min(q - env_q - byp_q, sta_cap)
            cum = 0.0
            for key, cap in units:
                hours[key] = all_hours[key]
                if cum + cap <= limit:
                    cum += cap
                else:
                    break
        return float(sum(hours.values())), hours

    # ---- the cell list, engine order: group -> day -> iteration ----------------------------------------------
    def _cells(self):
        """Cells are described per (scenario, species) group (``engine.CellGrid``: a few numbers per group); the explicit
        per-cell arrays are materialised only when something asks for them (injected replays, tests)."""
        c0, act_off = 0, 0
        max_it = max([g.iterations for g in self.groups] + [1])
        max_days = max([g.n_days for g in self.groups] + [1])
        n_sp = max(len(self.species_params), 1)
        recs = np.zeros(len(self.groups), _abi.GROUP_DTYPE)
        active = []
        for j, g in enumerate(self.groups):
            g.cell0 = c0
            if self.peaking:
                # hours are random per (species, iteration, day) in a peaking plant; drawn on the host like the
                # reference does (SURVEY.md §8 f-4 "next" row) — every cell keeps its own draw
                act = np.ones(g.n_days, bool)
            else:
                act = self.tot_hours[g.day0:g.day0 + g.n_days] > 0
            days = np.nonzero(act)[0]
            g.active_days = days
            g.n_active_days = len(days)
            recs[j] = (c0, g.iterations, len(days), g.day0, g.species_index, act_off, 0,
                       (g.scen_index * n_sp + g.species_index) * max_it)
            active.append(days.astype(np.int32))
            act_off += len(days)
            c0 += len(days) * g.iterations
        keep = (recs["n_active_days"] > 0) & (recs["iterations"] > 0)
        self.grid_group_index = np.nonzero(keep)[0]        # grid group -> index into self.groups
        self.grid = engine.CellGrid(recs[keep], np.concatenate(active).astype(np.int32) if active else np.zeros(0, np.int32),
                                    int(max_days))
        self.n_cells = c0
        self._explicit = None
        self.cell_hours = {}
        if self.peaking:
            self._peaking_hours()

    def _explicit_cells(self):
        if self._explicit is None:
            self._explicit = self.grid.explicit()
        return self._explicit

    cell_day = property(lambda self: self._explicit_cells()[0])
    cell_species = property(lambda self: self._explicit_cells()[1])
    cell_id = property(lambda self: self._explicit_cells()[2])

    def _peaking_hours(self):
        """Peaking / pumped-storage plants (stryke.py:2405-2431): per (species, iteration, day) every unit flips a
        not-operating coin and draws lognormal hours clipped to [0, 24]; a cell with no operating unit is skipped
        altogether (stryke.py:3025).  The draws are made on the device from the cell's Philox stream (count_kernel,
        include/stryke_b200.h ``StrykePeaking``); here only the per-(scenario, unit) parameter table is built.  A recorded
        run is replayed by injecting its hours (``sim._injected_hours``)."""
        sim = self.sim
        given = getattr(sim, "_injected_hours", None)      # replay of a recorded run: {(scenario, species): (keys, H[I, D, units])}
        self.peak_params, self.peak_keys = None, {}
        tables_ = []
        for gi, g in enumerate(self.groups):
            if given is not None:
                keys, H = given[(str(g.scenario), str(g.species))]
                self.cell_hours[gi] = (list(keys), np.ascontiguousarray(np.transpose(np.asarray(H, dtype=float), (1, 0, 2))))
                continue
            ops = sim.operating_scenarios_df[sim.operating_scenarios_df.Scenario == g.scenario].reset_index(drop=True)
            keys, rows = [], []
            for fac in ops.Facility.unique():
                fac_type = sim.facility_params.at[fac, "Operations"]
                if isinstance(fac_type, pd.Series):
                    fac_type = fac_type.iloc[0]
                fu = sim.unit_params[sim.unit_params.Facility == fac].sort_values(by="op_order")
                for uname, row in fu.iterrows():
                    orow = ops[(ops.Facility == fac) & (ops.Unit == row["Unit"])]
                    if orow.empty:
                        continue
                    o = orow.iloc[0]
                    keys.append(str(uname))
                    if fac_type == "run-of-river":
                        rows.append([0.0, np.nan, 0.0, 0.0, float(o["Hours"])])
                    else:
                        rows.append([float(o["Prob_Not_Op"]), float(o["shape"]), float(o["location"]), float(o["scale"]), 0.0])
            self.peak_keys[gi] = keys
            tables_.append((g.species_index, np.array(rows, dtype=np.float64).reshape(-1, 5)))
        if tables_:
            U = max(t.shape[0] for _, t in tables_)
            pk = np.zeros((len(self.species_params), U, 5))
            pk[:, :, 1] = np.nan                    # units a scenario does not list: fixed 0 hours
            for si, t in tables_:
                pk[si, :t.shape[0]] = t
            self.peak_params = pk

    def set_device_hours(self, hours):
        """Hours per (cell, unit) drawn on the device (engine order) -> ``cell_hours`` per group, [days, iterations, units]."""
        for gi, g in enumerate(self.groups):
            keys = self.peak_keys.get(gi)
            if keys is None:
                continue
            D, I = g.n_active_days, g.iterations
            H = np.asarray(hours[g.cell0:g.cell0 + D * I], dtype=np.float64).reshape(D, I, -1)[:, :, :len(keys)]
            self.cell_hours[gi] = (keys, H)

    @property
    def species_table(self):
        return np.stack(self.species_params) if self.species_params else np.zeros((0, _abi.SPECIES_W))

    def fish_cap(self):
        """Effective per-cell population cap: STRYKE_MAX_DAILY_FISH (stryke.py:2610, :3052) and
        STRYKE_MAX_DAILY_STATE_ROWS / len(moves) (stryke.py:3059-3067)."""
        max_fish = int(os.environ.get("STRYKE_MAX_DAILY_FISH", "100000"))
        max_rows = int(os.environ.get("STRYKE_MAX_DAILY_STATE_ROWS", "2000000"))
        return max_fish, max_rows, max(1, len(self.sim.moves))

    def check_fixed_counts(self):
        max_fish, max_rows, M = self.fish_cap()
        for g in self.groups:
            p = self.species_params[g.species_index]
            if int(p[12]) != _abi.ENT["fixed"]:
                continue
            n = int(p[18]) or 1
            if n > max_fish:
                raise ValueError(f"Population size n={n} exceeds STRYKE_MAX_DAILY_FISH={max_fish} for species "
                                 f"'{g.species}', scenario '{g.scenario}'. Reduce entrainment parameters or increase "
                                 "STRYKE_MAX_DAILY_FISH deliberately.")
            if n * M > max_rows:
                raise ValueError(f"Estimated daily state rows {n * M} exceed STRYKE_MAX_DAILY_STATE_ROWS={max_rows} "
                                 f"(n={n}, moves={M}, species='{g.species}', scenario='{g.scenario}').")


def run_simulation(sim, injected=None, trace=None):
    """Body of ``simulation.run()``."""
    print("[DEBUG RUN START] ========== SIMULATION BEGINNING ==========", flush=True)
    print(f"[DEBUG RUN START] Project: {getattr(sim, 'project_name', 'UNKNOWN')}", flush=True)
    print(f"[DEBUG RUN START] Population records: {len(sim.pop) if hasattr(sim, 'pop') else 'NO POP DATA'}", flush=True)
    print(f"[DEBUG RUN START] Unit params: {len(sim.unit_params) if hasattr(sim, 'unit_params') else 'NO UNITS'} units", flush=True)
    plan = RunPlan(sim)
    plan.check_fixed_counts()
    max_fish, max_rows, M = plan.fish_cap()
    cap = min(max_fish, max_rows // M)
    rank, world = _dist_rank_world()
    if world > 1 and "device" not in sim.__dict__:       # one process per GPU: this process's GPU unless the caller chose one
        sim.device = int(os.environ.get("LOCAL_RANK", rank))
    res = None
    try:
        if injected is not None or trace is not None:
            # replay of recorded draws / per-fish trace: dense tallies, explicit cell arrays (parity and debugging path)
            res = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id,
                                   device=sim.device, precision=sim.precision, seed=sim.seed, max_daily_fish=cap,
                                   injected=injected, trace=trace)
            pieces = [(plan.grid, res)]
        elif plan.peaking and plan.peak_params is not None:
            # peaking plants: hours per (cell, unit) are drawn on the device with the counts (small projects: one GPU, rank 0)
            if rank != 0:
                return None
            cres = run_compact_retry(plan, sim.device, sim.precision, sim.seed, cap, peaking=plan.peak_params)
            plan.set_device_hours(cres.hours)
            pieces = [(plan.grid, cres)]
        elif world > 1:
            # one process per GPU (torchrun): every rank simulates its iterations of every group, the ranks exchange their
            # compact rows; rank 0 writes the frames (Philox is keyed by the cell, so the result does not depend on world)
            from .distributed import run_sharded
            pieces = run_sharded(plan, rank, world, device=sim.device, precision=sim.precision, seed=sim.seed, max_daily_fish=cap)
            if pieces is None:          # the rows went to rank 0, which writes the frames
                return None
        else:
            pieces = [(plan.grid, run_compact_retry(plan, sim.device, sim.precision, sim.seed, cap))]
    except _abi.StrykeError as exc:
        if exc.code == _abi.E_MAX_FISH:
            raise ValueError(f"{exc}. Adjust entrainment inputs or raise STRYKE_MAX_DAILY_FISH / "
                             "STRYKE_MAX_DAILY_STATE_ROWS deliberately.") from exc
        raise
    tallies = group_tallies(plan, pieces)
    sim._last_plan, sim._last_result, sim._last_tallies = plan, res, tallies
    warn = int(os.environ.get("STRYKE_WARN_DAILY_FISH", "50000"))        # stryke.py:3041-3051 warns cell by cell
    n_max = max([int(t.n.max()) for t in tallies.values() if t.n.size] + [0])
    if n_max >= warn:
        n_big = int(sum(int((t.n >= warn).sum()) for t in tallies.values()))
        logger.warning("Large daily fish population: %d cell(s) with n >= %d (largest n=%d; STRYKE_WARN_DAILY_FISH)",
                       n_big, warn, n_max)
    if rank != 0:          # the frames are written once
        return None
    frames = assemble_frames(sim, plan, tallies)
    # ---- optional per-fish table (STRYKE_STORE_SIM_TABLE, stryke.py:116, :3243-3249) ---------------------------------
    fish_tables = {}
    if _flag("STRYKE_STORE_SIM_TABLE"):
        tr = trace
        if res is None:
            res = dense_from_tallies(plan, tallies)
        if tr is None:      # second pass with the general kernel: same seed, same fish, now with the per-fish columns
            n_total = int(res.cell_n.astype(np.int64).sum())
            tr = engine.TraceBuffers.allocate(plan.fac.n_moves, n_total, probs=False)
            tr.n_fish = n_total
            again = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id,
                                     device=sim.device, precision=sim.precision, seed=sim.seed, max_daily_fish=cap, trace=tr)
            if not (np.array_equal(again.cell_n, res.cell_n) and np.array_equal(again.daily, res.daily)):
                raise RuntimeError("per-fish pass disagrees with the tally pass")       # same Philox words: cannot happen
        fish_tables = assemble_fish_tables(sim, plan, res, tr, rates=_flag("STRYKE_STORE_AGENT_TRACES"))
    with open_store(sim.hdf_path, mode="a") as hdf:
        for key, df in fish_tables.items():
            hdf.append(key, df, min_itemsize=256)
        for key, df in frames.items():
            if df is not None and len(df):
                kw = {}
                if key == "State_Daily":
                    kw = dict(min_itemsize={"scenario": 128, "species": 128, "state": 256})
                elif key in ("Route_Flows", "Unit_Hours"):
                    kw = dict(min_itemsize={"scenario": 128, "route": 256})
                hdf.append(key, df, **kw)
    all_iters = _flag("STRYKE_DAY_PROGRESS_ALL_ITERS", "1")
    for g in plan.groups:
        # the day/iteration progress lines the web UI parses for its progress bar (stryke.py:2922-2923): every
        # max(1, days // 10)-th day, all iterations unless STRYKE_DAY_PROGRESS_ALL_ITERS=0 (then the first three)
        step = max(1, int(g.n_days / 10))
        days = [d for d in range(1, g.n_days + 1) if d % step == 0]
        its = range(1, g.iterations + 1) if all_iters else range(1, min(g.iterations, 3) + 1)
        if days:
            print("\n".join(f"[INFO] Day {d} of {g.n_days} (Iteration {i})" for i in its for d in days), flush=True)
        print(f"[INFO] ✅ Completed scenario '{g.scenario}' for species {g.species}", flush=True)
    print("[INFO] ✅ All scenarios complete! Finalizing results...", flush=True)
    print("[INFO] 💾 Simulation data saved successfully", flush=True)
    return None


def assemble_fish_tables(sim, plan: RunPlan, res: engine.CellResult, tr, rates=False):
    """The reference's per-fish ``fishes`` table (stryke.py:3105-3127, :3230-3236), one frame per
    ``simulations/{scenario}/{species}`` key, rows in the reference's order (iteration -> day -> fish).  ``draw_k`` of
    STRYKE_STORE_AGENT_TRACES is not kept (the dice are Philox words consumed in registers); ``rates_k`` is."""
    fac = plan.fac
    M = fac.n_moves
    names = np.array([str(x) for x in fac.node_names], dtype=object)
    dates = np.array(plan.day_dates, dtype="datetime64[ns]")
    n_all = res.cell_n.astype(np.int64)
    off = np.concatenate([[0], np.cumsum(n_all)])
    out = {}
    for g in plan.groups:
        D, I = g.n_active_days, g.iterations
        if D == 0 or I == 0:
            continue
        # engine cell of (iteration i, active day d) = cell0 + d * I + i; the reference loops i outer, d inner
        cells = (g.cell0 + np.arange(D)[None, :] * I + np.arange(I)[:, None]).ravel()
        n = n_all[cells]
        total = int(n.sum())
        if total == 0:
            continue
        first = np.repeat(off[cells], n)
        within = np.arange(total) - np.repeat(np.cumsum(n) - n, n)
        fish = first + within                                   # index into the trace arrays
        day_rows = g.day0 + g.active_days
        it_col = np.repeat(np.repeat(np.arange(I, dtype=np.int64), D), n)
        d_idx = np.repeat(np.tile(day_rows, I), n)
        cols = {"scenario_num": np.repeat(g.scen_num, total), "species": np.repeat(g.species, total),
                "flow_scenario": np.repeat(g.scenario, total), "season": np.repeat(g.season, total),
                "iteration": it_col, "day": dates[d_idx], "flow": plan.days.curr_Q[d_idx].astype(np.float64),
                "population": tr.length_ft[fish].astype(np.float32)}
        for k in range(M):
            cols[f"state_{k}"] = names[tr.state[k, fish]].astype(str)
            if rates:
                cols[f"rates_{k}"] = tr.rate[k, fish].astype(np.float32)
            cols[f"survival_{k}"] = tr.survival[k, fish].astype(np.float32)
        out[f"simulations/{g.scenario}/{g.species}"] = pd.DataFrame(cols)
    return out


def run_compact_retry(plan, device, precision, seed, cap, peaking=None):
    """``engine.run_compact`` on the plan's cell grid with an estimated row capacity, once more with the exact need when the
    estimate was too small (the engine names it)."""
    import re
    from .distributed import expected_slots
    rows_cap = min(2 * plan.n_cells, expected_slots(plan.grid, plan.species_table)) if plan.n_cells > 4096 else 2 * plan.n_cells
    try:
        return engine.run_compact(plan.fac, plan.days, plan.species_table, plan.grid, device=device, precision=precision, seed=seed,
                                  max_daily_fish=cap, rows_cap=rows_cap, peaking=peaking)
    except _abi.StrykeError as exc:
        m = re.search(r"rows buffer too small: (\d+) slots needed", str(exc))
        if not m:
            raise
        return engine.run_compact(plan.fac, plan.days, plan.species_table, plan.grid, device=device, precision=precision, seed=seed,
                                  max_daily_fish=cap, rows_cap=int(m.group(1)), peaking=peaking)


@dataclass
class GroupTallies:
    """Tallies of one (scenario, species) group in the reference's order (iteration-major, day-minor)."""
    n: np.ndarray          # int64[I, D]  pop_size (0 = species absent that day)
    daily: np.ndarray      # int64[I, D, 5]
    it: np.ndarray         # present cells, sorted by (iteration, day): iteration
    dd: np.ndarray         #                                            index of the active day
    state: np.ndarray      # [len(it), n_pairs, 2] successes, count


def group_tallies(plan, pieces):
    """{index into plan.groups: GroupTallies} from the device results: ``pieces`` = [(engine.CellGrid, engine.CompactResult
    or dense engine.CellResult)], one per rank (iterations of a group may be spread over the pieces, in rank order)."""
    P = plan.fac.n_pairs
    parts = {}
    for grid_p, r in pieces:
        kept, ilo = getattr(grid_p, "kept", None), getattr(grid_p, "iter_lo", None)
        if isinstance(r, engine.CompactResult):
            cells, n, daily, state = r.columns()
        else:
            cells = np.flatnonzero(r.cell_n > 0)
            n, daily, state = r.cell_n[cells].astype(np.int64), r.daily[cells].astype(np.int64), r.state_daily[cells]
        for j, g in enumerate(grid_p.groups):
            gi = int(plan.grid_group_index[kept[j] if kept is not None else j])
            Ip, D, c0 = int(g["iterations"]), int(g["n_active_days"]), int(g["cell0"])
            a, b = np.searchsorted(cells, [c0, c0 + Ip * D])
            loc = cells[a:b] - c0
            parts.setdefault(gi, []).append((int(ilo[j]) if ilo is not None else 0, loc % Ip, loc // Ip, n[a:b], daily[a:b], state[a:b]))
    out = {}
    for gi, g in enumerate(plan.groups):
        I, D = g.iterations, g.n_active_days
        N, DL = np.zeros((I, D), np.int64), np.zeros((I, D, _abi.DAILY_W), np.int64)
        its, dds, sts = [], [], []
        for i_lo, ii, dd, n, daily, state in parts.get(gi, []):
            N[i_lo + ii, dd] = n
            DL[i_lo + ii, dd] = daily
            its.append(i_lo + ii); dds.append(dd); sts.append(state)
        it = np.concatenate(its) if its else np.zeros(0, np.int64)
        dd = np.concatenate(dds) if dds else np.zeros(0, np.int64)
        st = np.concatenate(sts) if sts else np.zeros((0, P, 2), np.uint16)
        order = np.argsort(it * max(D, 1) + dd, kind="stable")
        out[gi] = GroupTallies(N, DL, it[order], dd[order], st[order])
    return out


def dense_from_tallies(plan, tallies):
    """Engine-order dense ``CellResult`` (day-major, iteration-minor per group) from the grouped tallies."""
    P = plan.fac.n_pairs
    out_n = np.zeros(plan.n_cells, np.int32)
    out_d = np.zeros((plan.n_cells, _abi.DAILY_W), np.uint32)
    out_s = np.zeros((plan.n_cells, P, 2), np.uint32)
    for gi, g in enumerate(plan.groups):
        t = tallies[gi]
        I, D = g.iterations, g.n_active_days
        if I == 0 or D == 0:
            continue
        sl = slice(g.cell0, g.cell0 + I * D)
        out_n[sl] = t.n.T.ravel()
        out_d[sl] = t.daily.transpose(1, 0, 2).reshape(I * D, -1)
        out_s[g.cell0 + t.dd * I + t.it] = t.state
    return engine.CellResult(out_n, out_d, out_s, None)


def assemble_frames(sim, plan: RunPlan, tallies):
    """Grouped tallies -> the reference's result frames, in the reference's row order
    (scenario -> species -> iteration -> day; stryke.py:3256-3383, :3433-3502).  Columns are gathered as arrays over all
    groups and every frame is built once; the label columns (species, scenario, season, state) take their strings from a
    small table by code — object columns like the reference's for ordinary projects, categoricals above
    STRYKE_B200_CATEGORICAL_ROWS rows (default 50 000), where one Python string per row would cost more than the simulation
    (the result store writes them dictionary-encoded and hands them back as plain object columns)."""
    fac = plan.fac
    P = fac.n_pairs
    metric = sim.output_units == "metric"
    dates = np.array(plan.day_dates, dtype="datetime64[ns]")
    pair_name = np.array([str(fac.node_names[v]) for v in fac.pair_node], dtype=object)
    # State_Daily rows of one cell come out of a groupby per move: moves ascending, states alphabetical (:3319-3324)
    pair_order = np.array(sorted(range(P), key=lambda p: (int(fac.pair_move[p]), str(pair_name[p]))), dtype=np.int64)
    move_of, name_of = fac.pair_move[pair_order].astype(np.int64), pair_name[pair_order]
    dcols = {k: [] for k in ("grp", "iteration", "day", "flow", "pop_size", "num_entrained", "num_survived", "num_mortality",
                             "mortality_impingement", "mortality_blade_strike", "mortality_barotrauma")}
    scols = {k: [] for k in ("grp", "iteration", "day", "move", "state", "successes", "count")}
    rf_parts, uh_parts = [], []
    logged = {}     # scenario -> bool[iteration, day row]: Route_Flows / Unit_Hours already written (stryke.py:3437)
    for gi, g in enumerate(plan.groups):
        D, I = g.n_active_days, g.iterations
        if D == 0 or I == 0:
            continue
        t = tallies[gi]
        keep = None                           # peaking plants: cells whose units all sat idle produce no rows (:3025)
        if gi in plan.cell_hours:
            k2 = (plan.cell_hours[gi][1].sum(axis=-1) > 0).T
            keep = None if k2.all() else k2
        dl = t.daily
        day_rows = g.day0 + g.active_days
        flow = plan.days.curr_Q[day_rows] * (CFS_TO_CMS if metric else 1.0)
        sel = slice(None) if keep is None else keep.ravel()
        ent, sur = dl[..., 0].ravel()[sel], dl[..., 1].ravel()[sel]
        gate = ent > 0                                              # stryke.py:3359
        dcols["grp"].append(np.full(ent.shape[0], gi, np.int32))
        dcols["iteration"].append(np.repeat(np.arange(I, dtype=np.int64), D)[sel])
        dcols["day"].append(np.tile(dates[day_rows], I)[sel])
        dcols["flow"].append(np.tile(flow.astype(np.float64), I)[sel])
        dcols["pop_size"].append(t.n.ravel()[sel])
        dcols["num_entrained"].append(ent)
        dcols["num_survived"].append(sur)
        dcols["num_mortality"].append(ent - sur)
        for j, nm in ((2, "mortality_impingement"), (3, "mortality_blade_strike"), (4, "mortality_barotrauma")):
            dcols[nm].append(np.where(gate, dl[..., j].ravel()[sel], 0))
        kp = None if keep is None else keep[t.it, t.dd]
        st = t.state if kp is None else t.state[kp]
        s_it, s_dd = (t.it, t.dd) if kp is None else (t.it[kp], t.dd[kp])
        cnt = st[:, pair_order, 1]
        ci, pi = np.nonzero(cnt)
        if ci.size:
            scols["grp"].append(np.full(ci.size, gi, np.int32))
            scols["iteration"].append(s_it[ci].astype(np.int64))
            scols["day"].append(dates[day_rows][s_dd[ci]])
            scols["move"].append(move_of[pi])
            scols["state"].append(pi.astype(np.int32))
            scols["successes"].append(st[ci, pair_order[pi], 0].astype(np.float64))
            scols["count"].append(cnt[ci, pi].astype(np.float64))
        # Route_Flows / Unit_Hours: once per (scenario, iteration, day), by the first species that gets there (:3437)
        n_rows = len(plan.day_rows)
        done = logged.setdefault(g.scenario, np.zeros((0, n_rows), bool))
        if done.shape[0] < I:
            done = np.vstack([done, np.zeros((I - done.shape[0], n_rows), bool)])
        fresh = ~done[:I][:, day_rows]
        if fresh.any():
            hours = plan.cell_hours.get(gi)
            suc = np.zeros((I, D, P), np.int64)                    # successes of the group's cells (only its first species per scenario gets here)
            suc[s_it, s_dd] = st[:, pair_order, 0]
            rf, uh = _route_and_hours(plan, g, suc, pair_order, fresh, day_rows, dates, hours)
            rf_parts += rf
            uh_parts += uh
        done[:I, day_rows] = True
        logged[g.scenario] = done
    cat = lambda parts: pd.concat(parts, ignore_index=True) if parts else None
    join = lambda xs: np.concatenate(xs) if xs else np.zeros(0)
    daily = sd = None
    if dcols["grp"]:
        grp = join(dcols["grp"])
        daily = pd.DataFrame({
            "species": _labels(["{:50}".format(g.species) for g in plan.groups], grp),
            "scenario": _labels(["{:50}".format(g.scenario) for g in plan.groups], grp),
            "season": _labels(["{:50}".format(g.season) for g in plan.groups], grp),
            **{k: join(v) for k, v in dcols.items() if k != "grp"}}, copy=False)
    if scols["grp"]:
        grp = join(scols["grp"])
        sd = pd.DataFrame({
            "scenario": _labels([str(g.scenario) for g in plan.groups], grp), "species": _labels([str(g.species) for g in plan.groups], grp),
            "iteration": join(scols["iteration"]), "day": join(scols["day"]), "move": join(scols["move"]),
            "state": _labels([str(x) for x in name_of], join(scols["state"])), "successes": join(scols["successes"]),
            "count": join(scols["count"])}, copy=False)
    return {"Daily": daily, "State_Daily": sd, "Route_Flows": cat(rf_parts), "Unit_Hours": cat(uh_parts)}


def _labels(labels, codes):
    """Label column: ``labels[codes]`` as an object Series (what the reference writes) or, for very long frames, a
    categorical with the same values."""
    codes = np.asarray(codes)
    if codes.shape[0] > int(os.environ.get("STRYKE_B200_CATEGORICAL_ROWS", "50000")):
        cats, inv = np.unique(np.asarray(labels, dtype=object).astype(str), return_inverse=True)
        return pd.Categorical.from_codes(np.asarray(inv)[codes], categories=list(cats))
    return pd.Series(np.asarray(labels, dtype=object)[codes], dtype=object)


def _route_and_hours(plan, g, suc, pair_order, fresh, day_rows, dates, cell_hours):
    """Route_Flows (stryke.py:3433-3462): the flows ``movement`` recorded for the successors of every node a live
    fish stood at before the last move; Unit_Hours (stryke.py:3463-3502): hours_dict x the day's allocation.
    Cells of one day that saw live fish at the same set of nodes share their records, so the work is per (day,
    distinct live set), not per cell."""
    fac = plan.fac
    I, D, P = suc.shape
    movable = fac.pair_move[pair_order] < fac.n_moves - 1
    nodes = fac.pair_node[pair_order]
    rf = {"iteration": [], "day": [], "route": [], "discharge_cfs": []}       # column pieces, ONE frame per call at the end
    uh = {"iteration": [], "day": [], "route": [], "hours": [], "discharge_cfs": []}
    unit_pos = {u: j for j, u in enumerate(fac.unit_names)}
    for j, d in enumerate(day_rows):
        its = np.nonzero(fresh[:, j])[0]
        if its.size == 0:
            continue
        stay = plan.days.node_stay[d]
        alloc = {str(u): float(plan.days.unit_Q[d, unit_pos[u]]) for u in fac.unit_names}
        if cell_hours is not None:                 # peaking plant: every cell has its own hours
            names, H = cell_hours[0], cell_hours[1][j][its]
        else:
            names = list((plan.hours[d] or {}).keys())
            H = np.tile(np.array([float(plan.hours[d][k]) for k in names]), (its.size, 1)) if names else None
        if names:
            uh["iteration"].append(np.repeat(its, len(names)).astype(np.int64))
            uh["day"].append(np.full(its.size * len(names), dates[d]))
            uh["route"].append(np.tile(np.array([str(k) for k in names], dtype=object), its.size))
            uh["hours"].append(np.asarray(H, dtype=np.float64).ravel())
            uh["discharge_cfs"].append(np.tile(np.array([alloc.get(str(k), np.nan) for k in names]), its.size))
        live = (suc[its, j, :] > 0) & movable[None, :]              # a live fish stood at this pair before a move
        masks, inverse = np.unique(live, axis=0, return_inverse=True)
        inverse = np.asarray(inverse).ravel()
        for m, mask in enumerate(masks):
            rec = {}
            for p in np.nonzero(mask)[0]:
                v = int(nodes[p])
                if stay[v]:
                    continue
                lo, hi = fac.edge_off[v], fac.edge_off[v + 1]
                for name, q in zip(fac.succ_order[v], plan.days.edge_flow[d, lo:hi]):
                    rec[name] = float(q)
            if not rec:
                continue
            who = its[inverse == m]
            routes = np.array([str(k) for k in rec.keys()], dtype=object)
            rf["iteration"].append(np.repeat(who, len(routes)).astype(np.int64))
            rf["day"].append(np.full(who.size * len(routes), dates[d]))
            rf["route"].append(np.tile(routes, who.size))
            rf["discharge_cfs"].append(np.tile(np.array(list(rec.values()), dtype=np.float64), who.size))

    def frame(cols, order):
        if not cols["iteration"]:
            return []
        data = {"scenario": str(g.scenario)}
        data.update({k: np.concatenate(cols[k]) for k in order})
        df = pd.DataFrame(data)
        df["route"] = df["route"].astype(str)
        return [df]
    rf_rows = frame(rf, ("iteration", "day", "route", "discharge_cfs"))
    uh_rows = frame(uh, ("iteration", "day", "route", "hours", "discharge_cfs"))
    return rf_rows, uh_rows
