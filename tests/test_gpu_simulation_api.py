"""The reference-facing class API on the GPU: ``simulation.run()`` / ``summary()`` and the result frames.

With the reference's recorded uniforms injected, the ``Daily`` and ``State_Daily`` frames written by ``run()`` must
equal the reference's own rows — values AND row order — and ``Route_Flows`` / ``Unit_Hours`` must hold the same
records (golden frames: tests/golden/*.npz meta, generated from the unmodified reference)."""
import contextlib
import io
import os

import numpy as np
import pandas as pd
import pytest

from oracle import fixtures, stryke_oracle as O
from golden_util import Golden
import gpu_util as G
from stryke_b200 import engine, runner, simulation
from stryke_b200.runner import RunPlan
from stryke_b200.store import open_store

pytestmark = pytest.mark.gpu
DAILY_COLS = ["pop_size", "num_entrained", "num_survived", "num_mortality", "mortality_impingement",
              "mortality_blade_strike", "mortality_barotrauma"]


def fresh_sim(prj, tmp_path, name):
    sim = simulation.__new__(simulation)
    fixtures.apply_to(sim, prj, proj_dir=str(tmp_path), output_name=name)
    with open_store(sim.hdf_path, mode="w") as st:
        st["Flow Scenarios"] = sim.flow_scenarios_df
        st["Operating Scenarios"] = sim.operating_scenarios_df
        st["Population"] = sim.pop
        st["Nodes"] = sim.nodes
        st["Edges"] = sim.edges
        st["Unit_Parameters"] = sim.unit_params
        st["Facilities"] = sim.facility_params
        st["Hydrograph"] = sim.input_hydrograph_df
    return sim


@pytest.mark.parametrize("case", ["c2_kaplan_francis", "c2_barotrauma", "c5_cascade", "c3_population", "c1_single_unit",
                                  "c2_low_flow", "c4_propeller_baro", "cabot_station", "cabot_station_fish", "example_v250304", "c3_extreme", "c2_normal_lengths", "c2_penstock", "c2_zero_flow",
                                  "c3_multi_scenario"])
def test_run_with_injected_uniforms_writes_the_reference_frames(case, tmp_path):
    g = Golden(case)
    sim = fresh_sim(g.prj, tmp_path, case)
    sim.precision = 64
    with contextlib.redirect_stdout(io.StringIO()):
        plan = RunPlan(sim)
    inj, off, ref_idx = G.golden_injected(g, plan)
    sim.graph = None
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        runner.run_simulation(sim, injected=inj)
    assert "will simulate" in out.getvalue()
    import re
    assert re.search(r"\[INFO\] Day \d+ of \d+ \(Iteration \d+\)", out.getvalue())      # the web UI's progress regex
    with open_store(sim.hdf_path, mode="r") as st:
        daily, sd = st["Daily"], st["State_Daily"] if "State_Daily" in st else None
        rf = st["Route_Flows"] if "Route_Flows" in st else None
        uh = st["Unit_Hours"]
    assert len(daily) == len(g.cells)
    ref_daily = np.stack([g.ref(ci, "daily") for ci in range(len(g.cells))])
    assert np.array_equal(daily[DAILY_COLS].to_numpy(dtype=np.int64), ref_daily)
    assert [s.strip() for s in daily["species"]] == [cm["species"] for cm in g.cells]
    assert [s.strip() for s in daily["scenario"]] == [cm["scenario"] for cm in g.cells]
    assert list(daily["iteration"]) == [cm["iteration"] for cm in g.cells]
    assert [str(d) for d in daily["day"]] == [cm["day"] for cm in g.cells]
    assert np.allclose(daily["flow"].to_numpy(), [cm["curr_Q"] for cm in g.cells], rtol=0, atol=0)
    # State_Daily: same rows in the same order as the reference's groupby output
    want_rows = []
    for ci, cm in enumerate(g.cells):
        if cm["n"] == 0:
            continue
        for m, s, v in zip(g.ref(ci, "sd_move"), g.ref(ci, "sd_state"), g.ref(ci, "sd_vals")):
            want_rows.append((cm["scenario"], cm["species"], cm["iteration"], cm["day"], int(m), g.model.nodes[int(s)], float(v[0]), float(v[1])))
    got_rows = [(r.scenario, r.species, int(r.iteration), str(r.day), int(r.move), r.state, float(r.successes), float(r["count"]))
                for _, r in sd.iterrows()] if sd is not None else []
    assert got_rows == want_rows
    # Route_Flows / Unit_Hours: same record sets, same values
    key = lambda d: (d["scenario"], int(d["iteration"]), str(pd.to_datetime(d["day"])), d["route"])
    want_rf = {key(d): d["discharge_cfs"] for d in g.meta["route_flows"]}
    got_rf = {key(d): d["discharge_cfs"] for d in (rf.to_dict("records") if rf is not None else [])}
    # (the golden frames went through DataFrame.to_json, which keeps 10 significant decimals)
    assert set(got_rf) == set(want_rf)
    assert all(np.isclose(got_rf[k], want_rf[k], rtol=1e-9, atol=1e-9) for k in want_rf)
    want_uh = {key(d): (d["hours"], d["discharge_cfs"]) for d in g.meta["unit_hours"]}
    got_uh = {key(d): (d["hours"], d["discharge_cfs"]) for d in uh.to_dict("records")}
    assert set(got_uh) == set(want_uh)
    assert all(np.allclose(got_uh[k], want_uh[k], rtol=1e-9, atol=1e-9) for k in want_uh)


def test_run_and_summary_native_rng(tmp_path):
    """Public call sequence of webapp/app.py:4530-4539: import -> run() -> summary() -> read the store."""
    prj = fixtures.c2_facility(n_fish=5000, iterations=4, days=5)
    sim = fresh_sim(prj, tmp_path, "native")
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
        np.random.seed(3)
        sim.summary()
    with open_store(sim.hdf_path, mode="r") as st:
        keys = set(st.keys())
        daily, beta, yearly, diag = st["Daily"], st["Beta_Distributions"], st["Yearly_Summary"], st["Driver_Diagnostics"]
    for k in ("/Daily", "/State_Daily", "/Route_Flows", "/Unit_Hours", "/Daily_Summary", "/Beta_Distributions",
              "/Beta_Distributions_Units", "/Yearly_Summary", "/Driver_Diagnostics"):
        assert k in keys, (k, keys)
    assert len(daily) == 20 and (daily["pop_size"] == 5000).all()
    assert list(daily.columns) == ["species", "scenario", "season", "iteration", "day", "flow"] + DAILY_COLS
    assert len(daily["species"].iloc[0]) == 50
    whole = beta[beta["state"] == "whole"].iloc[0]
    assert 0.85 < whole["survival rate"] < 1.0 and whole["ll"] <= whole["survival rate"] <= whole["ul"]
    assert {"U1", "U2", "U3", "bypass_flow", "spillway"} <= set(sim.beta_df_units_only["Passage Route"])
    assert abs(beta.loc[beta["state"] == "bypass_flow", "survival rate"].iloc[0] - 0.98) < 0.02
    assert yearly.iloc[0]["prob_entrainment"] > 0.5
    assert set(diag["route"]) == {"U1", "U2", "U3"} and "flow_share" in diag.columns
    assert sim.cum_sum is not None and sim.daily_summary is not None and sim.beta_caption
    # the result file itself: <name>.h5 in the pandas table layout (by PyTables where it exists, else by stryke_b200.h5lite)
    from stryke_b200 import store
    store.wait_for_exports()
    assert os.path.isfile(sim.hdf_path)
    if not store.pytables_available():
        import h5parse
        f = h5parse.File(sim.hdf_path)
        tables = {k for k, o in f.walk() if o["kind"] == "dataset"}
        assert {"/Daily/table", "/State_Daily/table", "/Route_Flows/table", "/Unit_Hours/table", "/Daily_Summary/table",
                "/Beta_Distributions/table", "/Yearly_Summary/table", "/Population/table"} <= tables
        tab = f.get("Daily/table")
        rec = f.read(tab)
        cols = {c: (name, j) for name in f.get("Daily")["attrs"]["values_cols"] for j, c in enumerate(tab["attrs"][f"{name}_kind"])}
        name, j = cols["pop_size"]
        assert len(rec) == 20 and np.array_equal(rec[name][:, j], daily["pop_size"].to_numpy())
        name, j = cols["species"]
        assert [x.decode() for x in rec[name][:, j]] == list(daily["species"])


def test_population_cap_raises_like_the_reference(tmp_path, monkeypatch):
    """STRYKE_MAX_DAILY_FISH fail-fast (reference tests/test_population_sim_caps.py:23-52, stryke.py:2610-2616)."""
    prj = fixtures.c3_population(n_species=1, iterations=30, days=4)
    sim = fresh_sim(prj, tmp_path, "cap")
    monkeypatch.setenv("STRYKE_MAX_DAILY_FISH", "5")
    with contextlib.redirect_stdout(io.StringIO()), pytest.raises(ValueError, match="STRYKE_MAX_DAILY_FISH"):
        sim.run()
    prj2 = fixtures.c2_facility(n_fish=500000, iterations=1, days=1)
    sim2 = fresh_sim(prj2, tmp_path, "cap2")
    monkeypatch.setenv("STRYKE_MAX_DAILY_FISH", "100000")
    with contextlib.redirect_stdout(io.StringIO()), pytest.raises(ValueError, match="exceeds STRYKE_MAX_DAILY_FISH"):
        sim2.run()


def test_peaking_plant_hours_gate_and_unit_hours(tmp_path):
    """Peaking / pumped-storage operations (stryke.py:2405-2431): per-cell not-operating coin + lognormal hours; cells
    whose units all sat idle write no Daily row; Unit_Hours carries the per-cell hours."""
    prj = fixtures.peaking_facility(n_fish=500, iterations=40, days=6)
    sim = fresh_sim(prj, tmp_path, "peaking")
    np.random.seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
    with open_store(sim.hdf_path, mode="r") as st:
        daily, uh = st["Daily"], st["Unit_Hours"]
    n_cells = 40 * 6
    p_idle = 0.35 * 0.55 * 0.75            # all three units not operating
    assert len(uh) == n_cells * 3 and uh["hours"].between(0, 24).all()
    tot = uh.groupby(["iteration", "day"])["hours"].sum()
    assert len(daily) == int((tot > 0).sum()) and len(daily) < n_cells
    assert abs((tot == 0).mean() - p_idle) < 4 * np.sqrt(p_idle * (1 - p_idle) / n_cells)
    assert (daily["pop_size"] == 500).all()
    idle_u1 = (uh[uh.route == "U1"]["hours"] == 0).mean()
    assert abs(idle_u1 - 0.35) < 0.13


@pytest.mark.parametrize("case", ["c2_kaplan_francis", "c5_cascade", "c3_population", "cabot_station_fish"])
def test_per_fish_table_matches_the_reference_table(case, tmp_path, monkeypatch):
    """STRYKE_STORE_SIM_TABLE=1 (stryke.py:116, :3243): ``simulations/{scenario}/{species}`` holds the reference's
    ``fishes`` rows — same fish order, lengths (fp32 ``population``), states and survival flags — when run() replays the
    reference's recorded draws."""
    monkeypatch.setenv("STRYKE_STORE_SIM_TABLE", "1")
    monkeypatch.setenv("STRYKE_STORE_AGENT_TRACES", "1")
    g = Golden(case)
    sim = fresh_sim(g.prj, tmp_path, case)
    sim.precision = 64
    with contextlib.redirect_stdout(io.StringIO()):
        plan = RunPlan(sim)
    inj, off, ref_idx = G.golden_injected(g, plan)
    tr = engine.TraceBuffers.allocate(plan.fac.n_moves, int(off[-1]), probs=False)
    with contextlib.redirect_stdout(io.StringIO()):
        runner.run_simulation(sim, injected=inj, trace=tr)
    M = g.meta["moves"]
    with open_store(sim.hdf_path, mode="r") as st:
        keys = [k for k in st.keys() if "simulations" in k]
        tabs = {k.strip("/"): st[k] for k in keys}
    assert tabs
    seen = {k: 0 for k in tabs}
    for ci, cm in enumerate(g.cells):
        if cm["n"] == 0:
            continue
        key = f"simulations/{cm['scenario']}/{cm['species']}"
        df = tabs[key]
        rows = df.iloc[seen[key]:seen[key] + cm["n"]]
        seen[key] += cm["n"]
        assert (rows["iteration"] == cm["iteration"]).all() and (pd.to_datetime(rows["day"]) == pd.Timestamp(cm["day"])).all()
        assert (rows["flow"] == cm["curr_Q"]).all()
        want_states, want_surv, want_rates = g.ref(ci, "states"), g.ref(ci, "survival"), g.ref(ci, "rates")
        for k in range(M):
            assert list(rows[f"state_{k}"]) == [g.model.nodes[int(s)] for s in want_states[k]]
            assert np.array_equal(rows[f"survival_{k}"].to_numpy(), want_surv[k].astype(np.float32))
            close = np.isclose(rows[f"rates_{k}"].to_numpy(), want_rates[k], rtol=2e-7, atol=0)
            assert close.all()
        assert np.allclose(rows["population"].to_numpy(), g.ref(ci, "lengths_ft").astype(np.float32), rtol=3e-7)
    assert all(seen[k] == len(tabs[k]) for k in tabs)


def test_per_fish_table_in_population_mode_uses_a_second_pass(tmp_path, monkeypatch):
    """Native RNG, entrainment counts drawn on the device: the fish count is only known after the tally pass, the
    per-fish pass repeats the run with the same Philox key and must reproduce every tally."""
    monkeypatch.setenv("STRYKE_STORE_SIM_TABLE", "1")
    prj = fixtures.c3_population(n_species=3, iterations=3, days=6)
    sim = fresh_sim(prj, tmp_path, "fish_tab")
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
    with open_store(sim.hdf_path, mode="r") as st:
        daily = st["Daily"]
        tabs = {k.strip("/"): st[k] for k in st.keys() if "simulations" in k}
    assert sum(len(t) for t in tabs.values()) == int(daily["pop_size"].sum()) > 0
    for key, t in tabs.items():
        _, scen, spc = key.split("/")
        d = daily[(daily.species.str.strip() == spc) & (daily.scenario.str.strip() == scen)]
        per_cell = t.groupby(["iteration", "day"]).size()
        want = d[d.pop_size > 0].set_index(["iteration", "day"])["pop_size"]
        assert per_cell.sort_index().to_dict() == want.sort_index().to_dict()
        # entrained fish = fish whose path contains a node with 'U' in its name; they sum to Daily.num_entrained
        states = t[[c for c in t.columns if c.startswith("state_")]].to_numpy(dtype=str)
        ent = (np.char.find(states, "U") >= 0).any(axis=1)
        assert int(ent.sum()) == int(d["num_entrained"].sum())


def test_peaking_plant_replays_the_reference_run(tmp_path):
    """A peaking / pumped-storage style plant (per-unit not-operating coin + lognormal hours, stryke.py:2405-2431) recorded
    from the reference: with the recorded hours and uniforms injected, run() writes the reference's Daily, State_Daily
    and Unit_Hours rows — cells whose units all sat idle produce no rows at all (stryke.py:3025)."""
    g = Golden("peaking_c2", subdir="peaking")
    sim = fresh_sim(g.prj, tmp_path, "peaking")
    sim.precision = 64
    live = [c for c in g.cells if not c.get("idle")]
    assert len(live) < len(g.cells)                     # the recording does contain idle cells
    prow = g.prj["population"][0]
    I, D = int(prow["Iterations"]), len(g.prj["hydrograph"])
    keys = list(g.cells[0]["hours"].keys())
    H = np.array([[c["hours"][k] for k in keys] for c in g.cells]).reshape(I, D, len(keys))      # reference order: iteration, day
    sim._injected_hours = {(prow["Scenario"], prow["Species"]): (keys, H)}
    with contextlib.redirect_stdout(io.StringIO()):
        plan = RunPlan(sim)
    inj, off, ref_idx = G.golden_injected(g, plan)
    sim.graph = None
    with contextlib.redirect_stdout(io.StringIO()):
        runner.run_simulation(sim, injected=inj)
    with open_store(sim.hdf_path, mode="r") as st:
        daily, sd, uh = st["Daily"], st["State_Daily"], st["Unit_Hours"]
    assert len(daily) == len(live)
    want = np.stack([g.ref(ci, "daily") for ci, c in enumerate(g.cells) if not c.get("idle")])
    assert np.array_equal(daily[DAILY_COLS].to_numpy(dtype=np.int64), want)
    assert list(daily["iteration"]) == [c["iteration"] for c in live] and [str(d) for d in daily["day"]] == [c["day"] for c in live]
    want_rows = []
    for ci, cm in enumerate(g.cells):
        if cm.get("idle") or cm["n"] == 0:
            continue
        for m, s, v in zip(g.ref(ci, "sd_move"), g.ref(ci, "sd_state"), g.ref(ci, "sd_vals")):
            want_rows.append((cm["iteration"], cm["day"], int(m), g.model.nodes[int(s)], float(v[0]), float(v[1])))
    got_rows = [(int(r.iteration), str(r.day), int(r.move), r.state, float(r.successes), float(r["count"])) for _, r in sd.iterrows()]
    assert got_rows == want_rows
    key = lambda d: (int(d["iteration"]), str(pd.to_datetime(d["day"])), d["route"])
    want_uh = {key(d): (d["hours"], d["discharge_cfs"]) for d in g.meta["unit_hours"]}
    got_uh = {key(d): (d["hours"], d["discharge_cfs"]) for d in uh.to_dict("records")}
    assert set(got_uh) == set(want_uh)
    assert all(np.allclose(got_uh[k], want_uh[k], rtol=1e-9, atol=1e-9) for k in want_uh)


def test_device_peaking_hours_match_the_reference_distribution():
    """f-4: the not-operating coin and the clipped lognormal hours of ``daily_hours`` (stryke.py:2405-2431) are drawn on the
    device (count_kernel) from the cell's Philox stream.  Against 20 000 calls of the UNMODIFIED reference's daily_hours
    (tests/golden/peaking/hours_samples.npz, oracle/make_peaking_samples.py): binomial test of every unit's idle frequency,
    two-sample KS of the hours of operating units, the idle-cell frequency; idle cells hold no fish."""
    from scipy import stats
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "peaking", "hours_samples.npz"))
    ref = g["hours"].astype(np.float64)
    prj = fixtures.peaking_facility(n_fish=40, iterations=6000, days=4)
    sim, plan = G.make_plan(prj, "pk_stat")
    assert plan.peaking and plan.peak_params is not None and plan.peak_params.shape[1:] == (3, 5)
    assert [str(k) for k in g["keys"]] == plan.peak_keys[0]
    res = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.grid, None, None, precision=32, seed=99,
                           max_daily_fish=10 ** 6, peaking=plan.peak_params)
    H = res.hours.astype(np.float64)
    assert H.shape == (plan.n_cells, 3) and H.min() >= 0 and H.max() <= 24.0
    tests = []
    for j in range(3):
        p_not = float(g["prob_not_op"][j])
        tests.append(stats.binomtest(int((H[:, j] == 0).sum()), len(H), p_not).pvalue)
        tests.append(stats.ks_2samp(H[H[:, j] > 0, j], ref[ref[:, j] > 0, j]).pvalue)
        # and against the law itself: exp(shape Z) scale + location, clipped at 24
        law = stats.lognorm(float(g["shape"][j]), float(g["location"][j]), float(g["scale"][j]))
        op = H[(H[:, j] > 0) & (H[:, j] < 24.0), j]
        tests.append(stats.kstest(op, lambda x: law.cdf(x) / law.cdf(24.0)).pvalue)
    idle = H.sum(axis=1) == 0
    p_idle = float(np.prod(g["prob_not_op"]))
    tests.append(stats.binomtest(int(idle.sum()), len(H), p_idle).pvalue)
    assert min(tests) > 0.01 / len(tests), tests
    assert (res.cell_n[idle] == 0).all() and (res.cell_n[~idle] == 40).all()
    assert res.state_daily[idle].sum() == 0
