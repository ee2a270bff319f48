"""Ingest paths of the class API on the GPU: spreadsheet projects (``simulation(proj_dir, name, wks=...)`` ->
``worksheet_import``, stryke.py:348-487; sheet geometry SURVEY.md Appendix F) and web-app projects
(``webapp_import(data_dict, name)``, stryke.py:489-796; keys of webapp/app.py:4680-4698).  Each ingested project must
give exactly the tallies of the same project set up attribute-by-attribute (the way the reference's tests do)."""
import contextlib
import io
import json
import os

import networkx as nx
import numpy as np
import pandas as pd
import pytest
from networkx.readwrite import json_graph

from oracle import fixtures
from stryke_b200 import simulation, xlsx
from stryke_b200.store import open_store

pytestmark = pytest.mark.gpu

UNIT_COLS = ["Facility", "Unit", "Runner Type", "intake_vel", "op_order", "H", "RPM", "D", "fb_depth", "ps_D", "ps_length",
             "ada", "N", "Qopt", "Qcap", "Qper", "B", "iota", "D1", "D2", "lambda", "roughness", "submergence_depth",
             "elevation_head"]
POP_COLS = ["Common Name", "Species", "Scenario", "Fish", "shape", "location", "scale", "dist", "max_ent_rate", "occur_prob",
            "Iterations", "vertical_habitat", "Length_mean", "Length_sd", "U_crit", "beta_0", "beta_1", "length shape",
            "length location", "length scale", "Notes"]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def run_attr_project(prj, tmp_path, name, seed):
    sim = simulation.__new__(simulation)
    fixtures.apply_to(sim, prj, proj_dir=str(tmp_path), output_name=name)
    sim.seed = seed
    with quiet():
        sim.run()
    with open_store(sim.hdf_path, "r") as st:
        return st["Daily"], st["State_Daily"]


def write_workbook(prj, path):
    """A spreadsheet project with the geometry ``worksheet_import`` reads (header row = skiprows + 1, columns from B)."""
    units = pd.DataFrame(prj["unit_params"])
    units["Unit"] = units["Unit_Name"]          # spreadsheet projects name the unit nodes in the 'Unit' column
    for c in UNIT_COLS:
        if c not in units.columns:
            units[c] = np.nan
    fac = pd.DataFrame(prj["facility_params"])
    fac["Rack Spacing"] = fac["Rack Spacing"] * 12.0          # the sheet holds inches (stryke.py:447)
    ops = pd.DataFrame(prj["operating_scenarios"])
    ops["Unit"] = [units["Unit"].iloc[int(u) - 1] for u in ops["Unit"]]
    pop = pd.DataFrame(prj["population"])
    pop["Notes"] = "synthetic"
    meta = pd.DataFrame({"a": [f"row {i}" for i in range(16)], "b": ["x"] * 13 + [prj["output_units"]] + ["y"] * 2})
    frames = {"Background and Metadata": meta, "Nodes": pd.DataFrame(prj["nodes"]), "Edges": pd.DataFrame(prj["edges"]),
              "Unit Params": units[UNIT_COLS],
              "Facilities": fac[["Facility", "Scenario", "Operations", "Rack Spacing", "Min_Op_Flow", "Env_Flow", "Bypass_Flow", "Spillway"]],
              "Flow Scenarios": pd.DataFrame(prj["flow_scenarios"])[["Scenario", "Scenario Number", "Flow", "Gage", "FlowYear", "Prorate", "Season", "Months"]],
              "Hydrology": pd.DataFrame(prj["hydrograph"]),
              "Operating Scenarios": ops[["Scenario", "Facility", "Unit", "Hours", "Prob_Not_Op", "shape", "location", "scale"]],
              "Population": pop[POP_COLS]}
    rows = {"Background and Metadata": 0, "Nodes": 9, "Edges": 9, "Unit Params": 4, "Facilities": 3, "Flow Scenarios": 5,
            "Hydrology": 3, "Operating Scenarios": 8, "Population": 11}
    cols = {k: (0 if k == "Background and Metadata" else 1) for k in frames}
    xlsx.new_workbook(path)
    xlsx.append_sheets(path, frames, startrow=rows, startcol=cols, index=False)


def test_spreadsheet_project_runs_like_the_attribute_built_project(tmp_path):
    prj = fixtures.c2_facility(n_fish=4000, iterations=3, days=4, baro=True)
    want_daily, want_sd = run_attr_project(prj, tmp_path, "attr", seed=321)
    write_workbook(prj, str(tmp_path / "project.xlsx"))
    with quiet():
        sim = simulation(str(tmp_path), "sheet", wks="project.xlsx")
    assert list(sim.nodes.Location) == [n["Location"] for n in prj["nodes"]]
    assert list(sim.unit_params.index) == ["U1", "U2", "U3"] and sim.output_units == "imperial"
    assert sim.facility_params.at["FacA", "Rack Spacing"] == pytest.approx(2 / 12.0)
    assert sim.max_river_node == 1 and len(sim.pop) == 1
    sim.seed = 321
    with quiet():
        sim.run()
        np.random.seed(0)
        sim.summary()
    with open_store(sim.hdf_path, "r") as st:
        daily, sd = st["Daily"], st["State_Daily"]
        assert {"/Nodes", "/Edges", "/Unit_Parameters", "/Population", "/Flow Scenarios", "/Operating Scenarios"} <= set(st.keys())
    cols = ["iteration", "pop_size", "num_entrained", "num_survived", "num_mortality", "mortality_impingement",
            "mortality_blade_strike", "mortality_barotrauma"]
    assert daily[cols].reset_index(drop=True).equals(want_daily[cols].reset_index(drop=True))
    assert sd[["move", "state", "successes", "count"]].reset_index(drop=True).equals(
        want_sd[["move", "state", "successes", "count"]].reset_index(drop=True))
    # summary() appended its three sheets to the workbook (stryke.py:3895-3900)
    book = xlsx.Workbook(str(tmp_path / "project.xlsx"))
    assert {"beta fit", "daily summary", "yearly summary"} <= set(book.sheets)
    beta = book.frame("beta fit", None, 0)
    assert "survival rate" in beta.columns and len(beta) == len(sim.beta_df)


def webapp_project(prj, tmp_path):
    """Rename unit nodes the way the web app does ('<Facility> - Unit <n>', stryke.py:618) and write its CSV inputs."""
    ren = {u["Unit_Name"]: f"{u['Facility']} - Unit {u['Unit']}" for u in prj["unit_params"]}
    fix = lambda n: ren.get(n, n)
    nodes = [dict(n, Location=fix(n["Location"]), ID=fix(n["Location"])) for n in prj["nodes"]]
    edges = [dict(e, _from=fix(e["_from"]), _to=fix(e["_to"])) for e in prj["edges"]]
    units = pd.DataFrame(prj["unit_params"]).drop(columns=["Unit_Name"])
    units.to_csv(tmp_path / "unit_params.csv", index=False)
    pd.DataFrame(prj["operating_scenarios"]).to_csv(tmp_path / "operating_scenarios.csv", index=False)
    pd.DataFrame(prj["hydrograph"]).to_csv(tmp_path / "hydrograph.csv", index=False)
    G = nx.DiGraph()
    for n in nodes:
        G.add_node(n["Location"])
    for e in edges:
        G.add_edge(e["_from"], e["_to"], weight=e["weight"])
    data = dict(proj_dir=str(tmp_path), units_system="imperial", simulation_mode="multiple_powerhouses_simulated_entrainment_routing",
                model_setup="multiple_powerhouses_simulated_entrainment_routing", project_name="webapp test", project_notes="",
                units="imperial", graph_summary=dict(Nodes=nodes, Edges=edges), graph_data=json_graph.node_link_data(G),
                unit_parameters_file=str(tmp_path / "unit_params.csv"), facilities=prj["facility_params"],
                flow_scenarios=prj["flow_scenarios"], operating_scenarios_file=str(tmp_path / "operating_scenarios.csv"),
                population=prj["population"], hydrograph_file=str(tmp_path / "hydrograph.csv"))
    renamed = dict(prj, nodes=[{k: v for k, v in n.items() if k != "ID"} for n in nodes], edges=edges,
                   unit_params=[dict(u, Unit_Name=ren[u["Unit_Name"]]) for u in prj["unit_params"]])
    return json.loads(json.dumps(data)), renamed


def test_webapp_project_runs_like_the_attribute_built_project(tmp_path):
    prj = fixtures.c2_facility(n_fish=3000, iterations=2, days=3)
    data, renamed = webapp_project(prj, tmp_path)
    want_daily, want_sd = run_attr_project(renamed, tmp_path, "attr", seed=77)
    sim = simulation(str(tmp_path), "web", existing=False)
    with quiet():
        sim.webapp_import(data, "web")
    assert sim.hdf_path == os.path.join(str(tmp_path), "web.h5") and len(sim.moves) == 6
    assert list(sim.unit_params.index) == ["FacA - Unit 1", "FacA - Unit 2", "FacA - Unit 3"]
    sim.seed = 77
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        sim.run()
        sim.summary()
    assert "[INFO] Scenario 'Spring' will simulate 3 days" in out.getvalue()     # the UI parses these lines
    with open_store(sim.hdf_path, "r") as st:
        daily, sd = st["Daily"], st["State_Daily"]
        assert {"/Facilities", "/Hydrograph", "/Daily_Summary", "/Driver_Diagnostics"} <= set(st.keys())
    cols = ["iteration", "pop_size", "num_entrained", "num_survived", "num_mortality", "mortality_blade_strike"]
    assert daily[cols].reset_index(drop=True).equals(want_daily[cols].reset_index(drop=True))
    assert sd[["move", "state", "successes", "count"]].reset_index(drop=True).equals(
        want_sd[["move", "state", "successes", "count"]].reset_index(drop=True))
    assert set(sim.driver_diagnostics["route"]) == {"FacA - Unit 1", "FacA - Unit 2", "FacA - Unit 3"}


def test_reopening_an_existing_project_keeps_its_results(tmp_path):
    """simulation(proj_dir, name, wks=..., existing=True) (stryke.py:298-346): network sheets only, results untouched,
    summary() runs on what the earlier run wrote."""
    prj = fixtures.c2_facility(n_fish=2000, iterations=3, days=4)
    write_workbook(prj, str(tmp_path / "project.xlsx"))
    with quiet():
        sim = simulation(str(tmp_path), "again", wks="project.xlsx")
        sim.run()
    with open_store(sim.hdf_path, "r") as st:
        daily = st["Daily"]
    with quiet():
        re = simulation(str(tmp_path), "again", wks="project.xlsx", existing=True)
        assert list(re.moves) == list(sim.moves) and sorted(re.graph.edges) == sorted(sim.graph.edges)
        np.random.seed(1)
        re.summary()
    with open_store(re.hdf_path, "r") as st:
        assert st["Daily"].equals(daily)
        assert "/Beta_Distributions" in st.keys() and len(st["Yearly_Summary"]) == 1
