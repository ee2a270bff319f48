"""``simulation`` — the reference's class API (Stryke/stryke.py:249 ``class simulation``) on the B200 engine.

Same constructor, import methods, helper-method signatures, progress lines, result keys and error behaviour as the
reference; ``run()`` replaces the scenario/species/iteration/day Python loops (stryke.py:2806-3507) with one host
table build per (scenario, day) and one call into the CUDA engine for all cells.  There is no CPU fallback: the
per-fish work (entrainment count, lengths, routing draws, strike/barotrauma, survival rolls, tallies) only exists as
device code behind include/stryke_b200.h.

file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import contextlib

import json
import logging
import math
import os
from collections import defaultdict

import networkx as nx
import numpy as np
import pandas as pd
from networkx.readwrite import json_graph

from . import _abi, engine, routing, tables
from .store import open_store


def _is_writer():
    """Only rank 0 of a torch.distributed job creates / writes the project's result store (every rank runs the ingest)."""
    from .runner import _dist_rank_world
    return _dist_rank_world()[0] == 0

logger = logging.getLogger(__name__)


def _env_flag(name, default="0"):
    return str(os.environ.get(name, default)).strip().lower() in ("1", "true", "yes", "on")


def to_dataframe(data, numeric_cols=None, index_col=None):
    """stryke.py:164-173."""
    df = data if isinstance(data, pd.DataFrame) else pd.DataFrame(data)
    for col in numeric_cols or ():
        if col in df.columns:
            df[col] = pd.to_numeric(df[col], errors="coerce")
    if index_col and index_col in df.columns:
        df.set_index(index_col, inplace=True, drop=False)
    return df


def _read_csv(path, numeric_cols=None, index_col=None):
    """read_csv_if_exists (stryke.py:211-246): tolerant path handling + encoding fallback chain."""
    if not path or (isinstance(path, str) and not path.strip()):
        return None
    if not isinstance(path, (str, bytes, os.PathLike)):
        raise TypeError(f"read_csv_if_exists(file_path=...) expected a path, got {type(path).__name__}")
    if not os.path.exists(path):
        return None
    df, last = None, None
    for enc in ("utf-8", "utf-8-sig", "utf-16", "utf-16le", "utf-16be", "latin-1"):
        try:
            df = pd.read_csv(path, encoding=enc)
            break
        except UnicodeDecodeError as exc:
            last = exc
    if df is None:
        raise ValueError(f"Failed to decode CSV as UTF-8/UTF-16/Latin-1: {path}") from last
    for col in numeric_cols or ():
        if col in df.columns:
            df[col] = pd.to_numeric(df[col], errors="coerce")
    if index_col:
        if index_col in df.columns:
            df.set_index(index_col, inplace=True, drop=False)
        elif index_col == "Unit_Name" and {"Facility", "Unit"} <= set(df.columns):
            df["Unit_Name"] = df["Facility"].astype(str) + " - Unit " + df["Unit"].astype(str)
            df.set_index("Unit_Name", inplace=True, drop=False)
    return df


def _normalize_facility_flow_columns(df):
    """stryke.py:197-209."""
    if df is None or df.empty:
        return df
    # 'Exc_Rack_Spacing' is how the reference's example workbooks label the column stryke.py:447 reads as 'Rack Spacing'
    ren = {s: d for s, d in {"Min Op Flow": "Min_Op_Flow", "Env Flow": "Env_Flow", "Bypass Flow": "Bypass_Flow",
                             "Exc_Rack_Spacing": "Rack Spacing"}.items()
           if s in df.columns and d not in df.columns}
    return df.rename(columns=ren) if ren else df


FAMILY_RATIOS = {"Centrarchidae": 0.10, "Ictaluridae": 0.18, "Catostomidae": 0.18, "Acipenseridae": 0.125,
                 "Salmonidae": 0.20, "Percidae": 0.125, "Cyprinidae": 0.10}     # stryke.py:1270-1278

_PARAM15 = ("H", "RPM", "D", "Q", "ada", "N", "Qopt", "Qper", "_lambda", "iota", "D1", "D2", "B", "Q_p", "gamma")


class simulation:
    """Facility-specific individual-based entrainment simulation (API of stryke.py:249-4101)."""

    #: arithmetic width of the per-fish kernel: 32 (throughput) or 64 (validation, reference operation order)
    precision = 32
    #: Philox4x32-10 key of the native RNG
    seed = 0x5EED
    #: barotrauma pressures: "hydrostatic" (what the reference's run() does, stryke.py:1496-1503) or "friction" (the
    #: Darcy-Weisbach penstock head loss of Stryke/barotrauma.py composed as in Stryke/stryke_barotrauma.py:565-601; opt-in)
    barotrauma_pressure = "hydrostatic"
    #: CUDA ordinal; inside a torch.distributed job (one process per GPU) the default is the process's LOCAL_RANK
    device = 0

    def __init__(self, proj_dir, output_name, wks=None, existing=False):
        if existing is False and wks:
            self.worksheet_import(proj_dir, wks, output_name)
        elif existing is True and wks:
            self.existing_import(proj_dir, wks, output_name)

    # ----------------------------------------------------------------------------------------------------------
    # ingest
    # ----------------------------------------------------------------------------------------------------------
    def worksheet_import(self, proj_dir, wks, output_name):
        """Spreadsheet ingest (stryke.py:348-487): sheet geometry per SURVEY.md Appendix F."""
        from . import xlsx
        self.wks_dir = os.path.join(proj_dir, wks)
        book = xlsx.Workbook(self.wks_dir)
        self.nodes = book.frame("Nodes", "B:D", skiprows=9, expect=("Location", "Surv_Fun"))
        self.edges = book.frame("Edges", "B:D", skiprows=9, expect=("_from", "_to"))
        self.surv_fun_df = self.nodes[["Location", "Surv_Fun"]].set_index("Location")
        self.surv_fun_dict = self.surv_fun_df.to_dict("index")
        if len(self.nodes) > 1:
            self.max_river_node = self._max_river_node(self.nodes.Location.values)
        self.unit_params = book.frame("Unit Params", "B:Y", skiprows=4, expect=("Facility", "Unit"))
        # the reference indexes by 'Unit' and then reads row['Unit'] (stryke.py:389 vs :2724);
        # keep the column so run() works for spreadsheet projects
        self.unit_params.set_index("Unit", inplace=True, drop=False)
        self.facility_params = _normalize_facility_flow_columns(book.frame("Facilities", "B:I", skiprows=3, expect=("Facility", "Scenario")))
        self.facility_params.set_index("Facility", inplace=True)
        self.flow_cap = self.unit_params.groupby("Facility")["Qcap"].sum()
        self.flow_scenarios_df = book.frame("Flow Scenarios", "B:I", skiprows=5, dtype={"Gage": str}, expect=("Scenario", "Flow"))
        self.input_hydrograph_df = book.frame("Hydrology", "B:C", skiprows=3, dtype={"Date": str, "Discharge": float})
        meta = book.frame("Background and Metadata", None, skiprows=0)
        self.output_units = meta.iat[13, 1]
        up = self.unit_params
        if self.output_units == "metric":      # stryke.py:425-445
            self.input_hydrograph_df["Discharge"] = self.input_hydrograph_df["Discharge"] * 35.31469989
            for col, k in (("intake_vel", 3.28084), ("H", 3.28084), ("D", 3.28084), ("Qopt", 35.31469989),
                           ("Qcap", 35.31469989), ("Penstock_Qcap", 35.31469989), ("B", 3.28084), ("D1", 3.28084),
                           ("D2", 3.28084), ("fb_depth", 3.28084), ("ps_D", 3.28084), ("ps_length", 3.28084),
                           ("submergence_depth", 3.28084), ("elevation_head", 3.28084)):
                if col in up.columns:
                    up[col] = pd.to_numeric(up[col], errors="coerce") * k
            self.facility_params["Rack Spacing"] = self.facility_params["Rack Spacing"] * 0.0328084
        else:
            self.facility_params["Rack Spacing"] = self.facility_params["Rack Spacing"] / 12.0
        self.operating_scenarios_df = book.frame("Operating Scenarios", "B:I", skiprows=8, expect=("Scenario", "Facility", "Unit"))
        self.operating_scenarios_df["OpScenario"] = (self.operating_scenarios_df.Scenario.astype(str) + " "
                                                     + self.operating_scenarios_df.Unit.astype(str))
        self.ops_scens = None
        self.flow_scenarios = self.flow_scenarios_df["Scenario"].unique()
        self.op_scenarios = self.operating_scenarios_df["Scenario"].unique()
        self.pop = book.frame("Population", "B:V", skiprows=11, expect=("Species", "Scenario"))
        self.proj_dir, self.output_name = proj_dir, output_name
        self.hdf_path = os.path.join(self.proj_dir, "%s.h5" % self.output_name)
        with (open_store(self.hdf_path, mode="w") if _is_writer() else contextlib.nullcontext({})) as hdf:
            hdf["Flow Scenarios"] = self.flow_scenarios_df
            hdf["Operating Scenarios"] = self.operating_scenarios_df
            hdf["Population"] = self.pop
            hdf["Nodes"] = self.nodes
            hdf["Edges"] = self.edges
            hdf["Unit_Parameters"] = self.unit_params

    def existing_import(self, proj_dir, wks, output_name):
        """Re-open a project whose results already exist, e.g. to call summary() again (stryke.py:304-346): reads the
        network sheets only and leaves the result store untouched.  (The reference's version stops at
        ``create_route(self.wks_dir)`` — a one-argument call of a zero-argument method; the intent is kept.)"""
        from . import xlsx
        self.wks_dir = os.path.join(proj_dir, wks)
        self.proj_dir, self.output_name = proj_dir, output_name
        self.hdf_path = os.path.join(self.proj_dir, "%s.h5" % self.output_name)
        book = xlsx.Workbook(self.wks_dir)
        self.nodes = book.frame("Nodes", "B:D", skiprows=9, expect=("Location", "Surv_Fun"))
        self.edges = book.frame("Edges", "B:D", skiprows=9, expect=("_from", "_to"))
        self.surv_fun_df = self.nodes[["Location", "Surv_Fun"]].set_index("Location")
        self.surv_fun_dict = self.surv_fun_df.to_dict("index")
        if len(self.nodes) > 1:
            self.max_river_node = self._max_river_node(self.nodes.Location.values)
        self.create_route()

    @staticmethod
    def _max_river_node(locations):
        top = 0
        for name in locations:
            parts = str(name).split("_")
            if len(parts) > 1:
                try:
                    top = max(top, int(parts[2]))
                except (IndexError, ValueError):
                    pass
        return top

    def webapp_import(self, data_dict, output_name):
        """In-memory / web-app ingest (stryke.py:489-796; keys assembled at webapp/app.py:4680-4698)."""
        self.output_name = output_name
        self.output_units = data_dict.get("units_system")
        self.sim_mode = data_dict.get("simulation_mode")
        self.model_setup = data_dict.get("model_setup", "N/A")
        self.project_name = data_dict.get("project_name", "N/A")
        self.project_notes = data_dict.get("project_notes", "N/A")
        self.units_session = data_dict.get("units", "N/A")
        self.proj_dir = data_dict.get("proj_dir", os.getcwd())
        hdf_path = os.path.join(self.proj_dir, f"{output_name}.h5")
        self.wks_dir = hdf_path
        summary = data_dict.get("graph_summary", {})
        self.nodes = to_dataframe(summary.get("Nodes", []))
        edges = summary.get("Edges")
        self.edges = to_dataframe(edges) if edges else to_dataframe([])
        self.surv_fun_df = self.nodes[["Location", "Surv_Fun"]].set_index("Location")
        self.surv_fun_dict = self.surv_fun_df.to_dict("index")
        graph_json = data_dict.get("simulation_graph")
        if graph_json is not None:
            G = json_graph.node_link_graph(graph_json)
        else:
            G = nx.DiGraph()
            for _, row in self.nodes.iterrows():
                G.add_node(row.get("ID", row.get("Location")), **row.to_dict())
            for _, row in self.edges.iterrows():
                try:
                    w = float(row.get("weight", 1.0))
                except (ValueError, TypeError):
                    w = 1.0
                G.add_edge(row.get("_from"), row.get("_to"), Weight=w)   # capital W, as the reference (stryke.py:560)
        if "graph_data" in data_dict:
            G = json_graph.node_link_graph(data_dict["graph_data"])
        if len(self.nodes) > 1:
            try:
                paths = list(nx.all_shortest_paths(G, "river_node_0", "river_node_1"))
                self.moves = np.arange(0, max(len(p) for p in paths) + 1, 1)
            except nx.NetworkXNoPath:
                logger.warning("No path found between river_node_0 and river_node_1.")
                self.moves = np.zeros(1, dtype=np.int32)
        else:
            self.moves = np.zeros(1, dtype=np.int32)
        self.graph = G
        if "unit_parameters_file" in data_dict:
            self.unit_params = _read_csv(
                data_dict["unit_parameters_file"],
                numeric_cols=["B", "D", "D1", "D2", "H", "N", "Qcap", "Qopt", "RPM", "ada", "intake_vel", "iota", "lambda",
                              "op_order", "roughness", "fb_depth", "ps_D", "ps_length", "submergence_depth",
                              "elevation_head", "Penstock_Qcap"], index_col="Unit_Name")
            if self.unit_params is not None:
                self.unit_params["Unit_Name"] = self.unit_params.Facility + " - Unit " + self.unit_params.Unit.astype("str")
                self.unit_params.set_index("Unit_Name", inplace=True)
        elif "unit_parameters" in data_dict:
            self.unit_params = to_dataframe(data_dict["unit_parameters"])
        if "facilities" in data_dict:
            self.facility_params = _normalize_facility_flow_columns(to_dataframe(
                data_dict["facilities"], numeric_cols=["Bypass Flow", "Env Flow", "Min Op Flow", "Bypass_Flow", "Env_Flow",
                                                       "Min_Op_Flow", "Rack Spacing", "Units"], index_col="Facility"))
        if "flow_scenarios" in data_dict:
            self.flow_scenarios_df = to_dataframe(data_dict["flow_scenarios"], numeric_cols=["FlowYear", "Prorate"])
            self.flow_scenarios = self.flow_scenarios_df["Scenario"].unique()
        if "operating_scenarios_file" in data_dict:
            self.operating_scenarios_df = _read_csv(
                data_dict["operating_scenarios_file"],
                numeric_cols=["Hours", "Location", "Prob Not Operating", "Scale", "Shape", "Unit"])
        else:
            self.operating_scenarios_df = None
        if self.operating_scenarios_df is not None and "Scenario" in self.operating_scenarios_df.columns:
            self.op_scenarios = self.operating_scenarios_df["Scenario"].unique()
        else:
            self.op_scenarios = []
        if "population" in data_dict:
            pop = data_dict["population"]
            pop = [pop] if isinstance(pop, dict) else pop
            self.pop = to_dataframe(pop, numeric_cols=["Iterations", "Length_mean", "Length_sd", "U_crit", "length location",
                                                       "length scale", "length shape", "location", "max_ent_rate",
                                                       "occur_prob", "scale", "shape"])
        if "hydrograph_file" in data_dict:
            df = _read_csv(data_dict["hydrograph_file"])
            if df is None:
                raise ValueError(f"Failed to load hydrograph file: {data_dict['hydrograph_file']}")
            cols = set(df.columns)
            if {"datetimeUTC", "DAvgFlow_prorate"} <= cols:
                src = ("datetimeUTC", "DAvgFlow_prorate")
            elif {"Date", "Discharge"} <= cols:
                src = ("Date", "Discharge")
            else:
                raise KeyError("Hydrograph file must include either ['datetimeUTC', 'DAvgFlow_prorate'] or ['Date', 'Discharge'].")
            df["datetimeUTC"] = pd.to_datetime(df[src[0]], errors="coerce")
            df["DAvgFlow_prorate"] = pd.to_numeric(df[src[1]], errors="coerce")
            bad_d, bad_q = int(df["datetimeUTC"].isna().sum()), int(df["DAvgFlow_prorate"].isna().sum())
            if bad_d or bad_q:
                raise ValueError(f"Hydrograph contains invalid {src[0]}/{src[1]} values: invalid_dates={bad_d}, "
                                 f"invalid_flows={bad_q}. Fix or remove malformed rows in hydrograph.csv.")
            self.input_hydrograph_df = df
        with (open_store(hdf_path, mode="w") if _is_writer() else contextlib.nullcontext({})) as hdf:
            for key, df in (("Flow Scenarios", getattr(self, "flow_scenarios_df", None)),
                            ("Operating Scenarios", getattr(self, "operating_scenarios_df", None)),
                            ("Population", getattr(self, "pop", None)), ("Nodes", self.nodes), ("Edges", self.edges),
                            ("Unit_Parameters", getattr(self, "unit_params", None)),
                            ("Facilities", getattr(self, "facility_params", None)),
                            ("Hydrograph", getattr(self, "input_hydrograph_df", None))):
                if df is not None:
                    hdf[key] = df
        self.hdf_path = hdf_path
        up = getattr(self, "unit_params", None)
        if up is not None:
            self.flow_cap = up.groupby("Facility")["Qcap"].sum() if {"Qcap", "Facility"} <= set(up.columns) else None
        logger.info("Web app import completed. Data stored to %s", hdf_path)

    # ----------------------------------------------------------------------------------------------------------
    # graph
    # ----------------------------------------------------------------------------------------------------------
    def create_route(self):
        """Movement graph + number of moves (stryke.py:1581-1655)."""
        if getattr(self, "graph", None) is not None:
            return
        route = nx.DiGraph()
        route.add_nodes_from(self.nodes.Location.values)
        if len(self.nodes) > 1:
            for _, row in self.edges.iterrows():
                route.add_edge(row["_from"], row["_to"], weight=row["weight"])
            longest = 0
            for path in nx.all_shortest_paths(route, "river_node_0", "river_node_%s" % self.max_river_node):
                longest = max(longest, len(path))
            self.moves = np.arange(0, longest + 1, 1)
        else:
            self.moves = np.zeros(1, dtype=np.int32)
        self.graph = route

    # ----------------------------------------------------------------------------------------------------------
    # host-side operations helpers (same signatures as the reference; they feed the table builder)
    # ----------------------------------------------------------------------------------------------------------
    def compute_unit_flow_allocations(self, curr_Q, unit_nodes, min_Q_dict, env_Q_dict, bypass_Q_dict, sta_cap_dict,
                                      unit_fac_dict, penstock_id_dict, penstock_cap_dict, cap_dict):
        """stryke.py:1657-1821 -> (unit_flow_allocations, prod_Q_by_facility, total_env_Q, total_bypass_Q)."""
        return routing.unit_flow_allocations(curr_Q, unit_nodes, min_Q_dict, env_Q_dict, bypass_Q_dict, sta_cap_dict,
                                             unit_fac_dict, penstock_id_dict, penstock_cap_dict, cap_dict,
                                             unit_params=getattr(self, "unit_params", None))

    def movement(self, location, status, swim_speed, graph, intake_vel_dict, Q_dict, op_order, cap_dict, unit_fac_dict,
                 penstock_id_dict=None, penstock_cap_dict=None, route_flow_recorder=None):
        """One routing decision for one fish (stryke.py:1823-2065).  Kept for API parity (the reference's routing
        tests call it for its ``route_flow_recorder`` side effect); ``run()`` never calls it — the same table goes to
        the device once per day instead."""
        if status != 1:
            return location
        nbors = list(graph.neighbors(location)) if Q_dict["curr_Q"] > 0.0 else []
        table = routing.route_table(location, nbors, lambda n: graph[location][n].get("weight", 1.0), Q_dict, op_order,
                                    cap_dict, unit_fac_dict, penstock_id_dict, penstock_cap_dict,
                                    unit_params=getattr(self, "unit_params", None),
                                    facility_params=getattr(self, "facility_params", None))
        if table is None:
            return location
        locs, probs, flows = table
        if route_flow_recorder is not None:
            for name, q in zip(locs, flows):
                route_flow_recorder[name] = float(q)
        try:
            cdf = routing.normalised_cdf(probs)
        except ValueError as exc:
            raise ValueError(f"Routing choice failed at location '{location}': probs sum {float(np.sum(probs))}, "
                             f"probs={probs}, locs={locs}") from exc
        return str(locs[int(np.searchsorted(cdf, np.random.random_sample(), side="right"))])

    def daily_hours(self, Q_dict, scenario, operations="independent"):
        """Hours and flow per unit for one day (stryke.py:2309-2503) -> (tot_hours, tot_flow, hours_dict, flow_dict)."""
        ops_df = self.operating_scenarios_df[self.operating_scenarios_df.Scenario == scenario]
        if not hasattr(self, "facility_params") or self.facility_params is None:
            raise ValueError(f"facility_params is missing; cannot compute daily_hours for scenario '{scenario}'.")
        if not isinstance(self.facility_params, pd.DataFrame):
            raise TypeError(f"facility_params must be a DataFrame, got {type(self.facility_params).__name__}.")
        if "Scenario" in self.facility_params.columns:
            seasonal = self.facility_params[self.facility_params.Scenario == scenario]
            if seasonal.empty:
                raise ValueError(f"Scenario '{scenario}' not found in facility_params.")
        else:
            seasonal = self.facility_params
        key_col = self.unit_params.index.name or "index"
        hours_dict, flow_dict = {}, {}
        for facility in ops_df.Facility.unique():
            curr_Q = Q_dict["curr_Q"]
            sta_cap = Q_dict["sta_cap"][facility]
            prod_Q = curr_Q - Q_dict["env_Q"][facility] - Q_dict["bypass_Q"][facility]
            fac_type = seasonal.at[facility, "Operations"]
            if isinstance(fac_type, pd.Series):
                fac_type = fac_type.iloc[0]
            fac_units = self.unit_params[self.unit_params.Facility == facility]
            if not fac_units.index.equals(pd.RangeIndex(len(fac_units))):
                fac_units = fac_units.reset_index(drop=fac_units.index.name in fac_units.columns)
            fac_units = fac_units.sort_values(by="op_order")

            def unit_key(i, row):
                key = row.get(key_col, None) if key_col in fac_units.columns else None
                if key is None or (isinstance(key, float) and np.isnan(key)):
                    key = row.get("Unit", None)
                if key is None or (isinstance(key, float) and np.isnan(key)):
                    key = i
                return str(key)

            if fac_type != "run-of-river":      # peaking / pumped storage, stryke.py:2405-2431
                from scipy.stats import lognorm
                for i in fac_units.index:
                    row = fac_units.loc[i]
                    key = unit_key(i, row)
                    shape, location, scale = ops_df.at[i, "shape"], ops_df.at[i, "location"], ops_df.at[i, "scale"]
                    lognorm.rvs(shape, location, scale, 1000)    # drawn and discarded by the reference (:2412)
                    if np.random.uniform(0, 1, 1) <= ops_df.at[i, "Prob_Not_Op"]:
                        hours_dict[key] = 0.0
                        flow_dict[key] = 0.0
                    else:
                        hours = min(max(lognorm.rvs(shape, location, scale, 1)[0], 0.0), 24.0)
                        hours_dict[key] = hours
                        flow_dict[key] = fac_units.at[i, "Qcap"] * hours * 3600.0
            elif len(fac_units):
                cum_Q, limit = 0.0, min(prod_Q, sta_cap)
                for i, row in fac_units.iterrows():
                    unit_ops = ops_df[ops_df.Unit == row["Unit"]]
                    if unit_ops.empty:
                        continue
                    hours, u_cap = unit_ops.iloc[0]["Hours"], row["Qcap"]
                    key = unit_key(i, row)
                    hours_dict[key] = hours
                    if cum_Q + u_cap <= limit:
                        flow_dict[key] = u_cap * hours * 3600.0
                        cum_Q += u_cap
                    else:
                        flow_dict[key] = (u_cap - (cum_Q + u_cap - limit)) * hours * 3600.0
                        break
        tot_hours = sum(hours_dict.values())
        tot_flow = sum(flow_dict[u] for u in hours_dict)
        return tot_hours, tot_flow, hours_dict, flow_dict

    # ----------------------------------------------------------------------------------------------------------
    # per-fish probability functions: device code only
    # ----------------------------------------------------------------------------------------------------------
    def _strike(self, kind, length, param_dict, rR=None):
        p = np.array([float(param_dict.get(k, np.nan)) for k in _PARAM15])
        scalar = np.ndim(length) == 0
        out = engine.strike_survival(_abi.KIND[kind], p, np.atleast_1d(np.asarray(length, float)), rR,
                                     precision=64, device=self.device)
        return float(out[0]) if scalar else out

    def Kaplan(self, length, param_dict):
        """Franke et al. 1997 Kaplan strike survival with rR ~ U(0.3, 1) (stryke.py:846-922), evaluated on the GPU."""
        n = np.size(length)
        return self._strike("Kaplan", length, param_dict, rR=np.random.uniform(0.3, 1.0, n))

    def Propeller(self, length, param_dict):
        """stryke.py:924-1009 (rR = 0.75)."""
        return self._strike("Propeller", length, param_dict)

    def Francis(self, length, param_dict):
        """stryke.py:1011-1094."""
        return self._strike("Francis", length, param_dict)

    def Pump(self, length, param_dict):
        """stryke.py:1096-1127."""
        return self._strike("Pump", length, param_dict)

    def node_surv_rate(self, length, u_crit, status, surv_fun, route, surv_dict, u_param_dict, barotrauma=False,
                       width_ratio=None):
        """Survival probability of ONE fish at one node (stryke.py:1326-1578): API-parity helper that evaluates the
        strike / barotrauma models on the device and keeps the reference's mortality-cause counters."""
        def scalar(x):
            if isinstance(x, (np.ndarray, list)) and len(x) == 1:
                return x[0]
            if hasattr(x, "item") and np.ndim(x) == 0:
                return x.item()
            return x
        length, u_crit, status, surv_fun, route = map(scalar, (length, u_crit, status, surv_fun, route))
        width_ratio = scalar(width_ratio)
        if width_ratio is None or (isinstance(width_ratio, float) and np.isnan(width_ratio)):
            width_ratio = 0.10
        width_ratio = float(width_ratio)
        if status == 0:
            return 0.0
        if surv_fun == "a priori":
            if route not in surv_dict:
                raise KeyError(f"A priori survival missing for route '{route}'. Available routes: {list(surv_dict.keys())}")
            prob = surv_dict[route]
            if prob is None or (isinstance(prob, float) and np.isnan(prob)):
                raise ValueError(f"A priori survival for route '{route}' is empty or NaN.")
            return np.float32(prob)
        param_dict = u_param_dict[route]
        blocked = length * width_ratio > param_dict["rack_spacing"]
        imp = 0.0 if (blocked and u_crit < param_dict["intake_vel"]) else 1.0
        if blocked:
            strike = 1.0
        elif surv_fun in ("Kaplan", "Propeller", "Francis", "Pump"):
            strike = getattr(self, surv_fun)(length, param_dict)
        else:
            raise UnboundLocalError(f"no strike model for survival function '{surv_fun}'")
        if barotrauma is True and not blocked:
            if route not in self.unit_params.index:
                raise KeyError(f"Route {route} not found in unit_params for barotrauma.")
            if "fb_depth" not in self.unit_params.columns:
                raise ValueError("Barotrauma requires fb_depth in Unit_Parameters.")
            fb = self.unit_params.at[route, "fb_depth"]
            if pd.isna(fb):
                raise ValueError(f"Barotrauma requires fb_depth for route {route}.")
            tail = self.unit_params.at[route, "elevation_head"] if "elevation_head" in self.unit_params.columns else None
            if tail is None or pd.isna(tail):
                tail = self.unit_params.at[route, "submergence_depth"] if "submergence_depth" in self.unit_params.columns else None
            if tail is None or pd.isna(tail):
                raise ValueError(f"Barotrauma requires elevation_head or submergence_depth for route {route}.")
            d1, d2 = tables.HABITAT.get(self.pop["vertical_habitat"].item(), (0.01, 1.0))
            depth = np.random.uniform(fb * d1, fb * d2, 1)[0]
            baro = float(engine.baro_survival(depth, tail, self.pop["beta_0"].item(), self.pop["beta_1"].item(),
                                              precision=64, device=self.device)[0])
        else:
            baro = 1.0
        prob = scalar(imp * strike * baro * 1.0)
        if not hasattr(self, "_mortality_components"):
            self._mortality_components = {"impingement": 0, "blade_strike": 0, "barotrauma": 0, "latent": 0}
        if np.random.random() > imp:
            self._mortality_components["impingement"] += 1
        elif np.random.random() > strike:
            self._mortality_components["blade_strike"] += 1
        elif np.random.random() > baro:
            self._mortality_components["barotrauma"] += 1
        elif np.random.random() > 1.0:
            self._mortality_components["latent"] += 1
        return np.float32(prob)

    def population_sim(self, output_units, spc_df, curr_Q):
        """Daily entrainment count for one cell (stryke.py:2505-2618), drawn on the device."""
        col = lambda c: spc_df.iat[0, spc_df.columns.get_loc(c)]
        params = np.zeros(_abi.SPECIES_W)
        params[2], params[6] = 1.0, 0.10          # a dummy length distribution: only the count kernel's output is used
        params[12] = _abi.ENT.get(col("dist"), _abi.ENT["Weibull"])
        params[13], params[14], params[15] = float(col("shape")), float(col("location")), float(col("scale"))
        try:
            params[16] = float(spc_df.max_ent_rate.values[0])
        except Exception:
            params[16] = np.nan
        params[17] = 1.0
        fac = _single_node_tables()
        days = tables.DayTables(np.array([float(curr_Q)]), np.zeros((1, 0)), np.ones((1, 1), np.uint8),
                                np.zeros((1, 0, _abi.UNIT_COEF_W)), np.zeros((1, 0)), np.zeros((1, 0)))
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        try:
            res = engine.run_cells(fac, days, params[None, :], [0], [0], [0], device=self.device, precision=32, seed=seed,
                                   max_daily_fish=int(os.environ.get("STRYKE_MAX_DAILY_FISH", "100000")),
                                   flags=_abi.FLAG_NO_PROMOTE)      # the n == 0 -> 1 promotion belongs to run() (stryke.py:3032)
        except _abi.StrykeError as exc:
            if exc.code == _abi.E_MAX_FISH:
                raise ValueError(f"{exc} Inputs: curr_Q={curr_Q}. Adjust entrainment inputs or raise "
                                 "STRYKE_MAX_DAILY_FISH deliberately.") from exc
            raise
        return float(res.cell_n[0])

    # ----------------------------------------------------------------------------------------------------------
    # fish width ratio (stryke.py:1220-1294)
    # ----------------------------------------------------------------------------------------------------------
    def _load_fish_class_lookup(self):
        if getattr(self, "_fish_class_loaded", False):
            return
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fish_family.json")
        try:
            with open(path) as fh:
                data = json.load(fh)
        except (OSError, ValueError) as exc:
            raise RuntimeError(f"Failed to load fish family lookup for width ratios from {path}: {exc}") from exc
        self._fish_family_by_species = data["species"]
        self._fish_family_by_common = data["common"]
        self._fish_family_by_genus = data["genus"]
        self._fish_class_loaded = True

    def _resolve_fish_family(self, species_name=None, common_name=None, family_name=None):
        if family_name is not None and str(family_name).strip():
            return str(family_name).strip()
        self._load_fish_class_lookup()
        common = str(common_name).strip().lower() if common_name is not None else ""
        if common and common in self._fish_family_by_common:
            return self._fish_family_by_common[common]
        species = str(species_name).strip().lower() if species_name is not None else ""
        if species and species in self._fish_family_by_species:
            return self._fish_family_by_species[species]
        genus = species.split()[0] if species else ""
        return self._fish_family_by_genus.get(genus) if genus else None

    def _resolve_fish_width_ratio(self, species_name=None, common_name=None, family_name=None):
        if not hasattr(self, "_fish_width_ratio_cache"):
            self._fish_width_ratio_cache = {}
        s_key = str(species_name).strip().lower() if species_name is not None else ""
        c_key = str(common_name).strip().lower() if common_name is not None else ""
        f_key = str(family_name).strip() if family_name is not None else ""
        key = (s_key, c_key, f_key)
        if key in self._fish_width_ratio_cache:
            return self._fish_width_ratio_cache[key]
        family = self._resolve_fish_family(species_name=species_name, common_name=common_name, family_name=family_name)
        ratio = FAMILY_RATIOS.get(family, 0.10)
        hint = c_key or s_key
        if family == "Centrarchidae":
            if "bass" in hint or "micropterus" in hint:
                ratio = 0.135
            elif "sunfish" in hint or "bluegill" in hint or "lepomis" in hint:
                ratio = 0.09
        elif family == "Cyprinidae" and "carp" in hint:
            ratio = 0.11
        self._fish_width_ratio_cache[key] = ratio
        return ratio

    def _validate_unit_params_required_fields(self):
        """stryke.py:1296-1324."""
        up = getattr(self, "unit_params", None)
        if up is None or up.empty:
            raise ValueError("Unit parameters are required but missing or empty.")
        missing = [c for c in ("Qopt", "Qcap") if c not in up.columns]
        if missing:
            raise ValueError(f"Unit parameters missing required columns: {', '.join(missing)}.")
        qopt, qcap = pd.to_numeric(up["Qopt"], errors="coerce"), pd.to_numeric(up["Qcap"], errors="coerce")
        labels = up.index.astype(str)
        if "Facility" in up.columns and "Unit" in up.columns:
            labels = up["Facility"].astype(str) + " - Unit " + up["Unit"].astype(str)
        elif "Unit" in up.columns:
            labels = up["Unit"].astype(str)
        bad_opt, bad_cap = qopt.isna() | (qopt <= 0), qcap.isna() | (qcap <= 0)
        if bad_opt.any() or bad_cap.any():
            msgs = []
            if bad_opt.any():
                msgs.append("Qopt missing/invalid for units: " + ", ".join(sorted(set(np.asarray(labels)[bad_opt.values].tolist()))))
            if bad_cap.any():
                msgs.append("Qcap missing/invalid for units: " + ", ".join(sorted(set(np.asarray(labels)[bad_cap.values].tolist()))))
            raise ValueError("Invalid unit parameters. " + "; ".join(msgs))

    # ----------------------------------------------------------------------------------------------------------
    # hydrograph (stryke.py:2154-2307; USGS download is out of scope — no network)
    # ----------------------------------------------------------------------------------------------------------
    @staticmethod
    def speed(L, A, M):
        """Swimming speed after Sambilay 1990 (stryke.py:2067-2092): L fish length (ft), A caudal fin aspect ratio,
        M swimming mode (0 sustained, 1 burst); result in ft/s."""
        sa = 10 ** (-0.828 + 0.6196 * np.log10(np.asarray(L, dtype=float) * 30.48) + 0.3478 * np.log10(A) + 0.7621 * np.asarray(M))
        return sa * 0.911344

    def create_graph_from_app(self, model_setup, session):
        """Graph of a web-app project (stryke.py:798-843): one unit node for the single-unit setups, otherwise the
        node-link graph the editor stored under ``graph_data``."""
        import networkx as nx
        if model_setup in ("single_unit_survival_only", "single_unit_simulated_entrainment"):
            G = nx.DiGraph()
            up = getattr(self, "unit_params", None)
            if up is not None and not up.empty:
                row = up.iloc[0]
                number = row["Unit"] if "Unit" in row else row.name
                G.add_node(f"unit_{number}", label=str(number), type="unit", **row.to_dict())
            else:
                print("Warning: No unit parameters available for single unit model.")
            return G
        graph_json = session.get("graph_data")
        if graph_json:
            from networkx.readwrite import json_graph
            edges_key = "links" if "links" in graph_json else "edges"
            return json_graph.node_link_graph(graph_json, edges=edges_key)
        print("No graph data found in session; returning an empty graph.")
        return nx.DiGraph()

    def get_USGS_hydrograph(self, gage, prorate, flow_year):
        """Daily mean flow of a USGS gage for one year (stryke.py:2094-2152): columns datetimeUTC, DAvgFlow,
        DAvgFlow_prorate, year, month.  The reference downloads it (hydrofunctions NWIS 'dv' service); this build has no
        network client, so the series comes from ``<gage>_<year>.csv`` (columns ``datetimeUTC, DAvgFlow``) in the
        directory named by ``self.usgs_cache_dir`` / ``$STRYKE_USGS_CACHE``, or from ``self.usgs_provider(gage, year)``."""
        df = None
        provider = getattr(self, "usgs_provider", None)
        if provider is not None:
            df = provider(str(gage), int(flow_year))
        else:
            cache = getattr(self, "usgs_cache_dir", None) or os.environ.get("STRYKE_USGS_CACHE")
            path = os.path.join(cache, f"{gage}_{int(flow_year)}.csv") if cache else None
            if path and os.path.isfile(path):
                df = pd.read_csv(path)
        if df is None:
            raise RuntimeError(f"USGS gage download ({gage}, {flow_year}) needs network access; supply the hydrograph in "
                               "the Hydrology sheet / hydrograph_file, or a <gage>_<year>.csv under $STRYKE_USGS_CACHE.")
        df = df[["datetimeUTC", "DAvgFlow"]].copy()
        df["DAvgFlow_prorate"] = df.DAvgFlow * prorate
        df["datetimeUTC"] = pd.to_datetime(df.datetimeUTC)
        df["year"] = pd.DatetimeIndex(df["datetimeUTC"]).year
        df = df[df.year == flow_year].copy()
        df["month"] = pd.DatetimeIndex(df["datetimeUTC"]).month
        return df

    def create_hydrograph(self, discharge_type, scen, scen_months, flow_scenarios_df, fixed_discharge=None):
        flow_df = pd.DataFrame()
        scen_df = flow_scenarios_df[flow_scenarios_df.Scenario == scen]
        flow_year = None
        if discharge_type == "hydrograph":
            gage = str(scen_df.at[scen_df.index[0], "Gage"])
            prorate = scen_df.at[scen_df.index[0], "Prorate"]
            flow_year = scen_df.at[scen_df.index[0], "FlowYear"]
            if sum(ch.isdigit() for ch in gage) > 1:
                df = self.get_USGS_hydrograph(gage, prorate, flow_year)
                for m in scen_months:
                    flow_df = pd.concat([flow_df, df[df.month == m]])
                if flow_df.empty:
                    raise ValueError(f"ERROR: Hydrograph is empty after filtering for scenario '{scen}'.\n"
                                     f"Requested months: {scen_months}, year: {flow_year}")
                return flow_df
            df = self.input_hydrograph_df.copy()
            if "Discharge" in df.columns and "Date" in df.columns:
                df["DAvgFlow_prorate"] = df["Discharge"] * prorate
                df["datetimeUTC"] = pd.to_datetime(df["Date"])
            elif "DAvgFlow_prorate" in df.columns and "datetimeUTC" in df.columns:
                df["datetimeUTC"] = pd.to_datetime(df["datetimeUTC"], errors="coerce")
                df["DAvgFlow_prorate"] = pd.to_numeric(df["DAvgFlow_prorate"], errors="coerce")
                bad = df["datetimeUTC"].isna() | df["DAvgFlow_prorate"].isna()
                if bad.any():
                    raise ValueError("Webapp hydrograph contains malformed rows in ['datetimeUTC', 'DAvgFlow_prorate'] "
                                     f"(invalid rows={int(bad.sum())}).")
            else:
                raise KeyError("Hydrograph DataFrame must have either ['Discharge', 'Date'] or "
                               "['DAvgFlow_prorate', 'datetimeUTC'] columns.")
            df["year"] = df["datetimeUTC"].dt.year
            df = df[df["year"] == flow_year]
            df["month"] = df["datetimeUTC"].dt.month.astype(int)
            for m in [int(x) for x in scen_months]:
                flow_df = pd.concat([flow_df, df[df["month"] == m]])
        elif discharge_type == "fixed":
            days_in = {1: 31, 2: 28, 3: 31, 4: 30, 5: 31, 6: 30, 7: 31, 8: 31, 9: 30, 10: 31, 11: 30, 12: 31}
            print(f"[DIAG] create_hydrograph: fixed discharge={fixed_discharge}", flush=True)
            acc = {}
            for month in scen_months:    # the reference re-emits earlier months on every pass (stryke.py:2261-2278)
                for day in np.arange(1, days_in[month] + 1, 1):
                    acc["2023-" + str(month) + "-" + str(day)] = fixed_discharge
                df = pd.DataFrame({"datetimeUTC": list(acc.keys()), "DAvgFlow_prorate": list(acc.values())})
                df["month"] = pd.to_datetime(df.datetimeUTC).dt.month
                flow_df = pd.concat([df, flow_df])
        if flow_df.empty:
            raise ValueError(f"ERROR: Hydrograph is empty after filtering for scenario '{scen}'.\n"
                             f"Requested months: {scen_months}, year: {flow_year if discharge_type == 'hydrograph' else 'N/A'}\n"
                             "Please ensure your hydrograph data covers the requested time period.")
        return flow_df

    # ----------------------------------------------------------------------------------------------------------
    # run
    # ----------------------------------------------------------------------------------------------------------
    def run(self):
        """The Monte Carlo run (stryke.py:2620-3517) on the GPU.  Writes Daily, State_Daily, Route_Flows and
        Unit_Hours to the project store."""
        from .runner import run_simulation
        return run_simulation(self)

    def summary(self):
        """stryke.py:3520-4101."""
        from .summary import summarize
        return summarize(self)


def _single_node_tables():
    """A one-node, zero-move facility: lets ``population_sim`` reuse the count kernel alone."""
    return tables.FacilityTables(["n"], [], 1, 0, False, np.zeros(1, np.uint8), np.ones(1), np.full(1, -1, np.int16),
                                 np.zeros(1, np.uint8), np.zeros(2, np.int32), np.zeros(0, np.int16), np.zeros(1, np.uint8),
                                 np.array([0, 1], np.int32), np.zeros(1, np.int16), np.zeros(1, np.int16),
                                 np.zeros((0, _abi.UNIT_STATIC_W)), [[]])
