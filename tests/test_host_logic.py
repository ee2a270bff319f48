"""CPU tests of the host side of the drop-in class API.  They read like the reference's own unit tests
(/root/reference/tests/*.py): ``simulation.__new__(simulation)`` + hand-set attributes, same method
signatures, same pinned numbers."""
import math

import networkx as nx
import numpy as np
import pandas as pd
import pytest

from stryke_b200 import simulation


def _sim(unit_params=None, facility_params=None):
    sim = simulation.__new__(simulation)
    if unit_params is not None:
        sim.unit_params = unit_params
    if facility_params is not None:
        sim.facility_params = facility_params
    return sim


def _two_units(qopt=(40.0, 40.0)):
    return pd.DataFrame([{"Facility": "FacilityA", "op_order": 1, "Qopt": qopt[0]},
                         {"Facility": "FacilityA", "op_order": 2, "Qopt": qopt[1]}],
                        index=pd.Index(["UnitA", "UnitB"], name="Unit"))


def _alloc(sim, curr_Q, min_q, sta_cap, pen_cap, caps):
    return sim.compute_unit_flow_allocations(
        curr_Q=curr_Q, unit_nodes=["UnitA", "UnitB"], min_Q_dict={"FacilityA": min_q}, env_Q_dict={"FacilityA": 0.0},
        bypass_Q_dict={"FacilityA": 0.0}, sta_cap_dict={"FacilityA": sta_cap},
        unit_fac_dict={"UnitA": "FacilityA", "UnitB": "FacilityA"}, penstock_id_dict={"UnitA": "P1", "UnitB": "P1"},
        penstock_cap_dict={"P1": pen_cap}, cap_dict={"UnitA": caps[0], "UnitB": caps[1]})


# reference: tests/test_compute_unit_flow_allocations.py:12-119
def test_no_production_below_min_flow():
    alloc, prod, _, _ = _alloc(_sim(_two_units()), 20.0, 30.0, 80.0, 80.0, (40.0, 40.0))
    assert prod["FacilityA"] == 0.0 and alloc.get("UnitA", 0.0) == 0.0 and alloc.get("UnitB", 0.0) == 0.0


def test_skip_units_below_min_start():
    alloc, _, _, _ = _alloc(_sim(_two_units((10.0, 40.0))), 100.0, 30.0, 70.0, 70.0, (20.0, 50.0))
    assert alloc["UnitA"] == 0.0 and alloc["UnitB"] == 50.0


def test_penstock_cap_limits_allocations():
    alloc, _, _, _ = _alloc(_sim(_two_units((50.0, 50.0))), 120.0, 0.0, 120.0, 60.0, (60.0, 60.0))
    assert alloc["UnitA"] == 50.0 and alloc["UnitB"] == 10.0 and alloc["UnitA"] + alloc["UnitB"] == 60.0


# reference: tests/test_routing_env_bypass.py:16-46
def test_env_bypass_spill_routing_flows():
    sim = _sim(facility_params=pd.DataFrame([{"Spillway": "spillway"}], index=pd.Index(["FacilityA"], name="Facility")))
    G = nx.DiGraph()
    for n in ("env_flow", "bypass_flow", "spillway"):
        G.add_edge("river_node_0", n)
    Q = {"curr_Q": 100.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 0.0}, "env_Q": {"FacilityA": 20.0},
         "bypass_Q": {"FacilityA": 10.0}}
    rec = {}
    sim.movement("river_node_0", 1, 0.0, G, {}, Q, {}, {}, {}, route_flow_recorder=rec)
    assert rec == {"env_flow": 20.0, "bypass_flow": 10.0, "spillway": 70.0}


# reference: tests/test_routing_units_env_bypass.py:22-58
def test_unit_env_bypass_spill_routing_flows():
    sim = _sim(pd.DataFrame([{"Facility": "FacilityA", "op_order": 1, "Qopt": 60.0}], index=pd.Index(["UnitA"], name="Unit")))
    G = nx.DiGraph()
    for n in ("UnitA", "env_flow", "bypass_flow", "spillway"):
        G.add_edge("river_node_0", n)
    Q = {"curr_Q": 100.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 60.0}, "env_Q": {"FacilityA": 20.0},
         "bypass_Q": {"FacilityA": 10.0}}
    rec = {}
    nxt = sim.movement("river_node_0", 1, 0.0, G, {}, Q, {"UnitA": 1}, {"UnitA": 60.0}, {"UnitA": "FacilityA"},
                       route_flow_recorder=rec)
    assert rec == {"UnitA": 60.0, "env_flow": 20.0, "bypass_flow": 10.0, "spillway": 10.0}
    assert nxt in rec
    assert sim.movement("river_node_0", 0, 0.0, G, {}, Q, {}, {}, {}) == "river_node_0"      # dead fish stay put


# reference: tests/test_movement_multiple_routes.py:16-51
def test_multiple_env_bypass_nodes_split_flow():
    sim = _sim(facility_params=pd.DataFrame([{"Spillway": "spillway"}], index=pd.Index(["FacilityA"], name="Facility")))
    G = nx.DiGraph()
    for n in ("env_flow_a", "env_flow_b", "bypass_flow_a", "bypass_flow_b", "spillway"):
        G.add_edge("river_node_0", n)
    Q = {"curr_Q": 100.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 0.0}, "env_Q": {"FacilityA": 20.0},
         "bypass_Q": {"FacilityA": 10.0}}
    rec = {}
    sim.movement("river_node_0", 1, 0.0, G, {}, Q, {}, {}, {}, route_flow_recorder=rec)
    assert rec == {"env_flow_a": 10.0, "env_flow_b": 10.0, "bypass_flow_a": 5.0, "bypass_flow_b": 5.0, "spillway": 70.0}
    assert sum(rec.values()) == 100.0


# reference: tests/test_daily_hours.py:35-53
def test_daily_hours_run_of_river_capacity_cut():
    sim = _sim(pd.DataFrame([{"Facility": "FacilityA", "Unit": 1, "op_order": 1, "Qcap": 40.0},
                             {"Facility": "FacilityA", "Unit": 2, "op_order": 2, "Qcap": 40.0}],
                            index=pd.Index(["U1", "U2"], name="Unit_Name")),
               pd.DataFrame([{"Scenario": "S", "Operations": "run-of-river"}], index=pd.Index(["FacilityA"], name="Facility")))
    sim.operating_scenarios_df = pd.DataFrame([{"Scenario": "S", "Facility": "FacilityA", "Unit": 1, "Hours": 24.0},
                                               {"Scenario": "S", "Facility": "FacilityA", "Unit": 2, "Hours": 24.0}])
    Q = {"curr_Q": 50.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 80.0}, "env_Q": {"FacilityA": 0.0},
         "bypass_Q": {"FacilityA": 0.0}}
    tot_hours, tot_flow, hours, flow = sim.daily_hours(Q, "S")
    assert set(hours) == {"U1", "U2"} and tot_hours == 48.0
    assert flow["U1"] == 40.0 * 24 * 3600.0 and flow["U2"] == 10.0 * 24 * 3600.0
    Q["curr_Q"] = 40.0
    _, _, hours, flow = sim.daily_hours(Q, "S")
    assert set(hours) == {"U1"} or flow.get("U2", 0.0) == 0.0
    with pytest.raises(ValueError, match="not found in facility_params"):
        sim.daily_hours(Q, "Other")


# reference: tests/test_rack_spacing.py:45-57 (width ratios, stryke.py:1260-1294)
def test_width_ratio_lookup():
    sim = _sim()
    assert sim._resolve_fish_width_ratio(species_name="Micropterus salmoides", common_name="Largemouth bass") == 0.135
    assert sim._resolve_fish_width_ratio(species_name="Lepomis macrochirus", common_name="Bluegill") == 0.09
    assert sim._resolve_fish_width_ratio(species_name="Alosa sapidissima", common_name="American shad") == 0.10
    assert sim._resolve_fish_width_ratio(species_name="x", common_name="y", family_name="Salmonidae") == 0.20
    assert sim._resolve_fish_width_ratio(species_name="Cyprinus carpio", common_name="Common carp") == 0.11
    assert sim._resolve_fish_width_ratio(species_name="unknown fish") == 0.10


# reference: tests/test_unit_params_validation.py:13-28 and tests/test_barotrauma_validation.py:35-72
def test_fail_loud_validation_messages():
    sim = _sim(pd.DataFrame([{"Facility": "F", "Unit": 1, "Qopt": np.nan, "Qcap": 10.0}], index=["U1"]))
    with pytest.raises(ValueError, match="Qopt missing/invalid for units: F - Unit 1"):
        sim._validate_unit_params_required_fields()
    from stryke_b200.tables import barotrauma_required
    up = pd.DataFrame([{"fb_depth": np.nan, "submergence_depth": 5.0}, {"fb_depth": 10.0, "submergence_depth": 5.0}], index=["U1", "U2"])
    with pytest.raises(ValueError, match=r"Barotrauma requires fb_depth for all units. Missing: \['U1'\]"):
        barotrauma_required(up)
    up = pd.DataFrame([{"fb_depth": 10.0, "submergence_depth": np.nan, "elevation_head": np.nan}], index=["U1"])
    with pytest.raises(ValueError, match="Barotrauma requires elevation_head or submergence_depth in Unit_Parameters"):
        barotrauma_required(up)
    assert barotrauma_required(pd.DataFrame([{"Qcap": 1.0}], index=["U1"])) is False


def test_create_route_moves_and_graph():
    from oracle import fixtures
    for prj, moves, nodes in ((fixtures.c2_facility(), 6, 9), (fixtures.c5_cascade(), 22, 40), (fixtures.c1_single_unit(), 1, 1)):
        sim = _sim()
        fixtures.apply_to(sim, prj)
        sim.create_route()
        assert len(sim.moves) == moves and sim.graph.number_of_nodes() == nodes


def test_fixed_discharge_hydrograph_repeats_months_like_the_reference():
    """stryke.py:2261-2278 re-emits the accumulated dictionary on every month (quirk kept)."""
    sim = _sim()
    df = pd.DataFrame([{"Scenario": "S"}])
    flow = sim.create_hydrograph("fixed", "S", [1, 2], df, fixed_discharge=100.0)
    assert len(flow) == (31 + 28) + 31 and (flow["DAvgFlow_prorate"] == 100.0).all()


def test_hydrograph_sheet_filter_and_errors():
    sim = _sim()
    sim.input_hydrograph_df = pd.DataFrame({"Date": pd.date_range("2023-01-01", periods=120).astype(str), "Discharge": np.arange(120.0)})
    fs = pd.DataFrame([{"Scenario": "S", "Gage": "", "Prorate": 2.0, "FlowYear": 2023}])
    flow = sim.create_hydrograph("hydrograph", "S", [3, 1], fs)
    assert len(flow) == 62 and flow["DAvgFlow_prorate"].iloc[0] == 2.0 * 59 and flow["month"].iloc[-1] == 1
    with pytest.raises(ValueError, match="Hydrograph is empty"):
        sim.create_hydrograph("hydrograph", "S", [9], fs)
    with pytest.raises(RuntimeError, match="network"):
        sim.create_hydrograph("hydrograph", "S", [3], pd.DataFrame([{"Scenario": "S", "Gage": "01170500", "Prorate": 1.0, "FlowYear": 2023}]))


# ---- the host table builder against the reference's own routing decisions ----------------------------------------------
from golden_util import CASES as _GOLDEN_CASES, Golden as _Golden


@pytest.mark.parametrize("case", _GOLDEN_CASES)
def test_day_tables_reproduce_the_reference_routing_probabilities(case):
    """tables.build_days (what the kernels read) vs the (successors, p) pairs captured at the reference's
    ``np.random.choice`` call sites (stryke.py:2058) while the goldens were recorded: same successor order, the
    normalised cumulative probabilities np.random.choice would search, bit for bit."""
    import gpu_util as G
    g = _Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    fac, days = plan.fac, plan.days
    names = list(plan.node_names) if hasattr(plan, "node_names") else list(g.model.nodes)
    idx = {nm: i for i, nm in enumerate(names)}
    ref_idx = G.reference_cell_index(plan)
    engine_of = {int(r): c for c, r in enumerate(ref_idx)}
    checked = 0
    for rec in g.meta["routing"]:
        c = engine_of[rec["cell"]]
        d = int(plan.cell_day[c])
        v = idx[rec["location"]]
        e0, e1 = int(fac.edge_off[v]), int(fac.edge_off[v + 1])
        assert [names[int(x)] for x in fac.edge_dst[e0:e1]] == rec["locs"]
        assert not days.node_stay[d, v]
        p = np.array(rec["probs"], dtype=np.float64)
        want = np.cumsum(p)
        want = want / want[-1]
        assert np.array_equal(days.edge_cdf[d, e0:e1], want), (rec["location"], days.edge_cdf[d, e0:e1], want)
        checked += 1
    assert checked == len(g.meta["routing"])


@pytest.mark.parametrize("fixture,kw", [("c2_facility", dict(n_fish=10, iterations=1, days=12, curr_Q=3000.0)),
                                        ("c5_cascade", dict(n_fish=10, iterations=1, days=9, curr_Q=1500.0)),
                                        ("cabot_station", dict(iterations=1, days=20, seed=2, median_Q=4000.0)),
                                        ("random_facility", dict(seed=4, n_fish=10, days=9)),
                                        ("random_facility", dict(seed=11, n_fish=10, days=9))])
def test_run_plan_hours_equal_the_daily_hours_method(fixture, kw):
    """RunPlan's per-day walk for run-of-river plants vs the API mirror ``simulation.daily_hours`` (stryke.py:2309-2503)
    called day by day: same total hours, same per-unit dict (keys, order, values)."""
    import gpu_util as G
    from oracle import fixtures
    prj = getattr(fixtures, fixture)(**kw)
    sim, plan = G.make_plan(prj, "hours")
    assert not plan.peaking
    for d, (scenario, q) in enumerate(plan.day_rows):
        th, _, hd, _ = sim.daily_hours(plan.ctx.q_dict(q, scenario), scenario)
        assert float(np.sum(th)) == plan.tot_hours[d]
        assert list(hd.items()) == list(plan.hours[d].items())
