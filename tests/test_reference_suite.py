"""The reference's own helper-level tests, re-stated against ``stryke_b200.simulation`` (same scenarios, same expected
values; objects built with ``simulation.__new__`` exactly as the reference tests do).  Reference files:
tests/test_rack_spacing.py, test_mortality_component_counters.py, test_barotrauma_validation.py,
test_unit_params_validation.py, test_population_sim_caps.py, test_movement_multiple_routes.py, test_route_stats.py.
Scenarios that evaluate a strike / barotrauma / entrainment model call device code and carry the gpu marker; input
validation that must fail before any device work runs on the CPU suite too."""
import networkx as nx
import numpy as np
import pandas as pd
import pytest

from stryke_b200 import simulation


def bare_sim(unit_params=None, habitat="Pelagic"):
    sim = simulation.__new__(simulation)
    sim.unit_params = pd.DataFrame() if unit_params is None else unit_params
    sim.pop = pd.DataFrame([{"vertical_habitat": habitat, "beta_0": -4.0, "beta_1": 3.0}])
    return sim


KAPLAN_UNIT = {"H": 10.0, "RPM": 100.0, "D": 5.0, "ada": 0.9, "N": 4.0, "Qopt": 50.0, "Qper": 0.5, "_lambda": 0.2}
FRANCIS_UNIT = {"intake_vel": 1.0, "rack_spacing": 1.0, "H": 10.0, "RPM": 100.0, "D": 5.0, "Q": 50.0, "Qper": 1.0, "ada": 0.9,
                "N": 4.0, "iota": 1.0, "D1": 3.0, "D2": 4.0, "B": 0.5, "_lambda": 0.2}


# ---- tests/test_rack_spacing.py ---------------------------------------------------------------------------------------
def test_fish_wider_than_the_rack_and_slower_than_the_intake_is_impinged():
    sim = bare_sim()
    rate = sim.node_surv_rate(length=2.0, u_crit=1.0, status=1, surv_fun="Francis", route="unitA", surv_dict={},
                              u_param_dict={"unitA": {"intake_vel": 2.0, "rack_spacing": 0.1}}, barotrauma=True)
    assert rate == 0.0


def test_fish_wider_than_the_rack_but_faster_than_the_intake_survives():
    sim = bare_sim()
    rate = sim.node_surv_rate(length=2.0, u_crit=3.0, status=1, surv_fun="Francis", route="unitA", surv_dict={},
                              u_param_dict={"unitA": {"intake_vel": 2.0, "rack_spacing": 0.1}}, barotrauma=True)
    assert rate == 1.0


def test_width_ratio_of_a_bass_overrides_the_family_default():
    sim = bare_sim()
    assert sim._resolve_fish_width_ratio(species_name="Micropterus", common_name="smallmouth bass") == 0.135


# ---- tests/test_mortality_component_counters.py -----------------------------------------------------------------------
def _depth_params():
    return pd.DataFrame([{"fb_depth": 10.0, "submergence_depth": 5.0}], index=pd.Index(["UnitA"], name="Unit"))


def test_blocked_slow_fish_counts_as_impingement_every_time():
    sim = bare_sim(_depth_params())
    sim.Kaplan = lambda length, params: 1.0
    unit = dict(KAPLAN_UNIT, intake_vel=2.0, rack_spacing=0.1)
    for _ in range(3):      # width 2.0 * 0.2 = 0.4 > rack; u_crit < intake velocity -> impingement survival 0
        sim.node_surv_rate(length=2.0, u_crit=0.5, status=1, surv_fun="Kaplan", route="UnitA", surv_dict={},
                           u_param_dict={"UnitA": unit}, barotrauma=False, width_ratio=0.2)
    assert sim._mortality_components == {"impingement": 3, "blade_strike": 0, "barotrauma": 0, "latent": 0}


def test_certain_survival_leaves_every_counter_at_zero():
    sim = bare_sim(_depth_params())
    sim.Kaplan = lambda length, params: 1.0
    unit = dict(KAPLAN_UNIT, intake_vel=0.1, rack_spacing=1.0)
    for _ in range(5):
        sim.node_surv_rate(length=1.0, u_crit=2.0, status=1, surv_fun="Kaplan", route="UnitA", surv_dict={},
                           u_param_dict={"UnitA": unit}, barotrauma=False, width_ratio=0.1)
    assert sim._mortality_components == {"impingement": 0, "blade_strike": 0, "barotrauma": 0, "latent": 0}


# ---- tests/test_barotrauma_validation.py ------------------------------------------------------------------------------
@pytest.mark.gpu
def test_barotrauma_without_forebay_depth_is_an_error():
    sim = bare_sim(pd.DataFrame([{"Facility": "FacilityA"}], index=pd.Index(["UnitA"], name="Unit")))
    with pytest.raises(ValueError, match="fb_depth"):
        sim.node_surv_rate(length=1.0, u_crit=2.0, status=1, surv_fun="Francis", route="UnitA", surv_dict={},
                           u_param_dict={"UnitA": dict(FRANCIS_UNIT)}, barotrauma=True)


@pytest.mark.gpu
def test_barotrauma_without_tailrace_depth_is_an_error():
    sim = bare_sim(pd.DataFrame([{"Facility": "FacilityA", "fb_depth": 10.0}], index=pd.Index(["UnitA"], name="Unit")))
    with pytest.raises(ValueError, match="elevation_head|submergence_depth"):
        sim.node_surv_rate(length=1.0, u_crit=2.0, status=1, surv_fun="Francis", route="UnitA", surv_dict={},
                           u_param_dict={"UnitA": dict(FRANCIS_UNIT)}, barotrauma=True)


@pytest.mark.gpu
def test_unit_survival_with_barotrauma_is_the_product_of_the_device_functions():
    """No reference test pins this value (strike and barotrauma numerics unpinned in the reference's
    tests); the product rule itself (stryke.py:1517) is checked against the exported device functions."""
    from stryke_b200 import engine
    up = pd.DataFrame([{"fb_depth": 30.0, "submergence_depth": 6.0}], index=pd.Index(["UnitA"], name="Unit"))
    sim = bare_sim(up, habitat="Benthic")
    np.random.seed(4)
    rate = sim.node_surv_rate(length=0.8, u_crit=2.0, status=1, surv_fun="Francis", route="UnitA", surv_dict={},
                              u_param_dict={"UnitA": dict(FRANCIS_UNIT)}, barotrauma=True)
    np.random.seed(4)
    strike = sim.Francis(0.8, dict(FRANCIS_UNIT))
    depth = np.random.uniform(30.0 * 0.8, 30.0 * 1.0, 1)[0]
    baro = float(engine.baro_survival(depth, 6.0, -4.0, 3.0, precision=64)[0])
    assert 0.0 < float(rate) < 1.0
    assert float(rate) == pytest.approx(float(np.float32(strike * baro)), rel=1e-6)


# ---- tests/test_unit_params_validation.py -----------------------------------------------------------------------------
@pytest.mark.parametrize("present,missing", [("Qcap", "Qopt"), ("Qopt", "Qcap")])
def test_unit_parameters_need_qopt_and_qcap(present, missing):
    df = pd.DataFrame([{"Facility": "FacilityA", "Unit": 1, present: 100.0}]).set_index(
        pd.Index(["FacilityA - Unit 1"], name="Unit_Name"))
    sim = bare_sim(df)
    with pytest.raises(ValueError, match=missing):
        sim._validate_unit_params_required_fields()


# ---- tests/test_population_sim_caps.py --------------------------------------------------------------------------------
def _runaway_species(max_ent_rate):
    # the reference test patches lognorm.rvs to return 1e12; the draw happens on the device here, so the same runaway
    # rate comes from the scale parameter (rate = exp(0.5 z) * 1e12 for any z)
    return pd.DataFrame([{"shape": 0.5, "location": 0.0, "scale": 1.0e12, "dist": "Log Normal", "max_ent_rate": float(max_ent_rate)}])


@pytest.mark.gpu
def test_entrainment_rate_is_capped_one_order_of_magnitude_above_the_observed_maximum():
    sim = simulation.__new__(simulation)
    n = sim.population_sim(output_units="metric", spc_df=_runaway_species(17.0), curr_Q=100.0)
    daily_mft3 = (60 * 60 * 24 * 100.0) / 1_000_000.0
    assert n == float(np.round(daily_mft3 * 17.0 * 10.0, 0))


@pytest.mark.gpu
def test_runaway_population_fails_fast(monkeypatch):
    sim = simulation.__new__(simulation)
    monkeypatch.setenv("STRYKE_MAX_DAILY_FISH", "1000000")
    with pytest.raises(ValueError, match="STRYKE_MAX_DAILY_FISH"):
        sim.population_sim(output_units="metric", spc_df=_runaway_species(1000.0), curr_Q=100000.0)


# ---- tests/test_movement_multiple_routes.py ---------------------------------------------------------------------------
def test_env_and_bypass_flow_split_evenly_over_their_nodes_and_spill_takes_the_rest():
    sim = simulation.__new__(simulation)
    sim.facility_params = pd.DataFrame([{"Spillway": "spillway"}], index=pd.Index(["FacilityA"], name="Facility"))
    G = nx.DiGraph()
    for nb in ("env_flow_a", "env_flow_b", "bypass_flow_a", "bypass_flow_b", "spillway"):
        G.add_edge("river_node_0", nb)
    Q = {"curr_Q": 100.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 0.0}, "env_Q": {"FacilityA": 20.0},
         "bypass_Q": {"FacilityA": 10.0}}
    flows = {}
    nxt = sim.movement("river_node_0", 1, 0.0, G, {}, Q, {}, {}, {}, route_flow_recorder=flows)
    assert nxt in set(G.successors("river_node_0"))
    assert flows == {"env_flow_a": 10.0, "env_flow_b": 10.0, "bypass_flow_a": 5.0, "bypass_flow_b": 5.0, "spillway": 70.0}
    assert sum(flows.values()) == 100.0


def test_dead_fish_do_not_move():
    sim = simulation.__new__(simulation)
    sim.facility_params = pd.DataFrame([{"Spillway": "spillway"}], index=pd.Index(["FacilityA"], name="Facility"))
    G = nx.DiGraph()
    G.add_edge("river_node_0", "spillway")
    Q = {"curr_Q": 100.0, "min_Q": {"FacilityA": 0.0}, "sta_cap": {"FacilityA": 0.0}, "env_Q": {"FacilityA": 0.0},
         "bypass_Q": {"FacilityA": 0.0}}
    assert sim.movement("river_node_0", 0, 0.0, G, {}, Q, {}, {}, {}) == "river_node_0"     # stryke.py:1840-1842


# ---- tests/test_route_stats.py ----------------------------------------------------------------------------------------
def test_first_non_river_state_is_the_passage_route():
    """The web app's route statistics (webapp/app.py "Policy B") take the first state that is not a river node; the
    State_Daily tallies written by run() carry the same information per (move, node)."""
    df = pd.DataFrame([
        {"index": 0, "iteration": 0, "day": "2024-01-01", "state_0": "river_node_0", "state_1": "Unit A", "survival_0": 1, "survival_1": 1},
        {"index": 1, "iteration": 0, "day": "2024-01-01", "state_0": "river_node_0", "state_1": "Unit B", "survival_0": 1, "survival_1": 0},
        {"index": 0, "iteration": 0, "day": "2024-01-02", "state_0": "river_node_0", "state_1": "Unit A", "survival_0": 1, "survival_1": 0}])
    states = [c for c in df.columns if c.startswith("state_")]
    first = df[states].apply(lambda r: next(s for s in r if "river_node" not in s.lower()), axis=1)
    assert first.value_counts().to_dict() == {"Unit A": 2, "Unit B": 1}


# ---- remaining public helpers of the class (stryke.py:798-843, :2067-2092) ---------------------------------------------
def test_swimming_speed_formula():
    # Sambilay 1990 as the reference writes it: 10^(-0.828 + 0.6196 log10(L*30.48) + 0.3478 log10(A) + 0.7621 M) * 0.911344
    want = 10 ** (-0.828 + 0.6196 * np.log10(1.5 * 30.48) + 0.3478 * np.log10(2.2) + 0.7621 * 1) * 0.911344
    assert simulation.speed(1.5, 2.2, 1) == pytest.approx(want, rel=1e-15)
    assert simulation.speed(np.array([1.5, 1.5]), 2.2, np.array([0, 1]))[1] == pytest.approx(want, rel=1e-15)


def test_graph_from_the_web_app_session():
    sim = bare_sim(pd.DataFrame([{"Facility": "F", "Unit": 3, "Qcap": 10.0}], index=pd.Index(["F - Unit 3"], name="Unit_Name")))
    G = sim.create_graph_from_app("single_unit_survival_only", {})
    assert list(G.nodes) == ["unit_3"] and G.nodes["unit_3"]["type"] == "unit" and G.nodes["unit_3"]["Qcap"] == 10.0
    from networkx.readwrite import json_graph
    H = nx.DiGraph()
    H.add_edge("river_node_0", "spill", weight=1.0)
    G2 = sim.create_graph_from_app("multiple_powerhouses_simulated_entrainment_routing", {"graph_data": json_graph.node_link_data(H)})
    assert list(G2.edges(data="weight")) == [("river_node_0", "spill", 1.0)]
    assert len(sim.create_graph_from_app("multiple_powerhouses_simulated_entrainment_routing", {})) == 0
