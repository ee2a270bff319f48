"""Synthetic projects of the shapes BASELINE.json names (C1-C5) and the reference's example layouts — the workload
definitions shared by bench.py, ``__graft_entry__.smoke()``, the tests and the oracle harness.

A *project* is a plain dict of lists-of-rows (no pandas objects) so that the very same description can be turned into

* the attribute set the reference's unit tests put on ``simulation.__new__(simulation)``
  (reference: tests/test_compute_unit_flow_allocations.py:6-9), and
* the attribute set of ``stryke_b200.simulation`` (same names, same meaning) — ``apply_to`` / ``make_plan``.

Facility shapes follow BASELINE.json:configs / SURVEY.md §8(d):
  C1 single-node Francis (Input_Spreadsheet_v250304.xlsx example),
  C2 3-unit Kaplan+Kaplan+Francis with bypass and spill (9 nodes, 6 moves),
  C3 C2 in population mode (EPRI-fitted entrainment, many species, several flow scenarios),
  C4 6-unit propeller / pump station with barotrauma,
  C5 5-dam cascade (40 nodes, 22 moves).
"""
from __future__ import annotations

import math

import numpy as np

NAN = float("nan")

KAPLAN = dict(**{"Runner Type": "Kaplan"}, intake_vel=2.0, H=40.0, RPM=120.0, D=12.0, ada=0.9, N=5,
              Qopt=2500.0, Qcap=3000.0, B=NAN, iota=NAN, D1=NAN, D2=NAN)
FRANCIS = dict(**{"Runner Type": "Francis"}, intake_vel=2.0, H=40.0, RPM=97.3, D=10.76, ada=0.94, N=13,
               Qopt=2150.72, Qcap=2288.0, B=4.0, iota=1.1, D1=11.22, D2=10.76)
PROPELLER = dict(**{"Runner Type": "Propeller"}, intake_vel=2.5, H=25.0, RPM=150.0, D=9.0, ada=0.88, N=4,
                 Qopt=1500.0, Qcap=1800.0, B=NAN, iota=NAN, D1=NAN, D2=NAN)

SHAD = dict(Species="Alosa", **{"Common Name": "American Shad"}, U_crit=1.6, vertical_habitat="Pelagic",
            beta_0=-5.301, beta_1=4.825, **{"length shape": 0.5907, "length location": -2.1245,
                                           "length scale": 15.9345})

# A few pre-fitted entrainment entries in the style of the reference's species library
# (webapp/app.py:2870-4018 ``species_defaults``; values quoted in SURVEY.md §8(d) C1/C3).
EPRI_SPECIES = [
    dict(Species="Ambloplites", **{"Common Name": "Rock Bass"}, dist="Pareto", shape=0.5974, location=0.0,
         scale=0.0052, max_ent_rate=1.85, occur_prob=0.3279, vertical_habitat="Benthic",
         beta_0=-4.657, beta_1=3.41, U_crit=1.6,
         **{"length shape": 0.5907, "length location": -2.1245, "length scale": 15.9345}),
    dict(Species="Micropterus", **{"Common Name": "Smallmouth Bass"}, dist="Log Normal", shape=0.99, location=0.0,
         scale=0.0013, max_ent_rate=0.0413, occur_prob=0.4118, vertical_habitat="Benthic",
         beta_0=-4.657, beta_1=3.41, U_crit=1.6,
         **{"length shape": 0.5907, "length location": -2.1245, "length scale": 15.9345}),
    dict(Species="Alosa", **{"Common Name": "American Shad"}, dist="Weibull", shape=0.62, location=0.0,
         scale=0.011, max_ent_rate=2.4, occur_prob=0.55, vertical_habitat="Pelagic",
         beta_0=-5.301, beta_1=4.825, U_crit=2.1,
         **{"length shape": 0.41, "length location": 0.0, "length scale": 9.8}),
    dict(Species="Lepomis", **{"Common Name": "Bluegill"}, dist="Pareto", shape=0.81, location=0.0,
         scale=0.0071, max_ent_rate=3.2, occur_prob=0.62, vertical_habitat="Pelagic",
         beta_0=-3.308, beta_1=6.195, U_crit=1.1,
         **{"length shape": 0.35, "length location": 0.0, "length scale": 7.4}),
]


def _unit_row(facility, unit_no, name, op_order, runner, baro=False, **over):
    row = dict(Facility=facility, Unit=unit_no, Unit_Name=name, op_order=op_order, **runner)
    row["lambda"] = 0.2
    if baro:
        row.update(fb_depth=30.0, submergence_depth=6.0, ps_D=10.0, ps_length=40.0, roughness=0.025,
                   elevation_head=NAN)
    row.update(over)
    return row


def _pop_row(species, scenario, fish=NAN, iterations=1, occur_prob=1.0, **over):
    row = dict(Species=species["Species"], Scenario=scenario, Fish=fish, shape=NAN, location=NAN, scale=NAN,
               dist=NAN, max_ent_rate=NAN, occur_prob=occur_prob, Iterations=iterations,
               Length_mean=NAN, Length_sd=NAN)
    row.update({k: v for k, v in species.items() if k != "Species"})
    if not (isinstance(fish, float) and math.isnan(fish)):
        # anadromous ("Fish = n") mode: no entrainment-rate distribution
        row.update(shape=NAN, location=NAN, scale=NAN, dist=NAN, max_ent_rate=NAN, occur_prob=occur_prob)
    row.update(over)
    return row


def _scenario_rows(names, months="3", flow="hydrograph", year=2023):
    return [dict(Scenario=s, **{"Scenario Number": i + 1}, Flow=flow, Gage="", FlowYear=year, Prorate=1.0,
                 Season=s, Months=months) for i, s in enumerate(names)]


def _months_for(days, start="2023-03-01"):
    import datetime as _dt
    d0 = _dt.date.fromisoformat(start)
    seen = []
    for i in range(days):
        m = (d0 + _dt.timedelta(days=i)).month
        if m not in seen:
            seen.append(m)
    return ",".join(str(m) for m in seen)


def _hydrograph(days, lo, hi, start="2023-03-01"):
    import datetime as _dt
    d0 = _dt.date.fromisoformat(start)
    q = np.linspace(lo, hi, days)
    return [dict(Date=(d0 + _dt.timedelta(days=i)).isoformat(), Discharge=float(q[i])) for i in range(days)]


def c2_facility(n_fish=1000, iterations=1, days=1, curr_Q=9000.0, baro=False, scenario="Spring",
                species=SHAD, occur_prob=1.0, length_normal=None, penstock=False, flows=None):
    """river_node_0 -> forebay -> {U1 Kaplan, U2 Kaplan, U3 Francis, bypass_flow, spillway}
    -> tailrace -> river_node_1.  Anadromous mode (Fish = n)."""
    nodes = [("river_node_0", "a priori", 1.0), ("forebay", "a priori", 1.0), ("U1", "Kaplan", 0.0),
             ("U2", "Kaplan", 0.0), ("U3", "Francis", 0.0), ("bypass_flow", "a priori", 0.98),
             ("spillway", "a priori", 0.97), ("tailrace", "a priori", 1.0), ("river_node_1", "a priori", 1.0)]
    mid = ["U1", "U2", "U3", "bypass_flow", "spillway"]
    edges = ([("river_node_0", "forebay", 1.0)] + [("forebay", x, 1.0) for x in mid]
             + [(x, "tailrace", 1.0) for x in mid] + [("tailrace", "river_node_1", 1.0)])
    units = [_unit_row("FacA", 1, "U1", 1, KAPLAN, baro), _unit_row("FacA", 2, "U2", 2, KAPLAN, baro),
             _unit_row("FacA", 3, "U3", 3, FRANCIS, baro)]
    if penstock:       # U1 and U2 share a penstock that carries less than their capacities (stryke.py:2722-2794, :1700-1790)
        units[0].update(Penstock_ID="P1", Penstock_Qcap=4200.0)
        units[1].update(Penstock_ID="P1", Penstock_Qcap=4200.0)
        units[2].update(Penstock_ID="P2", Penstock_Qcap=NAN)
    return dict(
        name="c2", output_units="imperial", max_river_node=1,
        nodes=[dict(Location=a, Surv_Fun=b, Survival=c) for a, b, c in nodes],
        edges=[dict(_from=a, _to=b, weight=w) for a, b, w in edges],
        unit_params=units,
        facility_params=[dict(Facility="FacA", Scenario=scenario, Operations="run-of-river",
                              **{"Rack Spacing": 2 / 12.0}, Min_Op_Flow=500.0, Env_Flow=0.0, Bypass_Flow=100.0,
                              Spillway="spillway")],
        flow_scenarios=_scenario_rows([scenario], months=_months_for(days)),
        operating_scenarios=[dict(Scenario=scenario, Facility="FacA", Unit=u, Hours=24.0, Prob_Not_Op=NAN,
                                  shape=NAN, location=NAN, scale=NAN) for u in (1, 2, 3)],
        hydrograph=(_hydrograph(days, curr_Q * 0.6, curr_Q * 1.4) if flows is None else
                    [dict(h, Discharge=float(q)) for h, q in zip(_hydrograph(len(flows), 1.0, 1.0), flows)]),
        population=[_pop_row(species, scenario, fish=float(n_fish), iterations=iterations, occur_prob=occur_prob,
                             **({} if length_normal is None else {   # lengths ~ |N(mean, sd)| inches (stryke.py:3090)
                                 "Length_mean": float(length_normal[0]), "Length_sd": float(length_normal[1]),
                                 "length shape": NAN, "length location": NAN, "length scale": NAN}))],
    )


# A generalised-extreme-value entrainment rate (population_sim's 'Extreme' branch, stryke.py:2560; the web app's species
# library holds no such entry, epri.ExtremeFit produces them)
EXTREME_SPECIES = dict(Species="Perca", **{"Common Name": "Yellow Perch"}, dist="Extreme", shape=-0.35, location=0.004,
                       scale=0.006, max_ent_rate=1.4, occur_prob=0.58, vertical_habitat="Benthic", beta_0=-4.657,
                       beta_1=3.41, U_crit=1.3, **{"length shape": 0.45, "length location": 0.0, "length scale": 11.0})


def c3_population(n_species=4, iterations=2, days=5, curr_Q=9000.0, scenarios=("Spring",), baro=False, extreme=False,
                  facility_scenario_column=True):
    """C2's facility in population mode: daily counts drawn from EPRI-fitted Pareto / Log Normal / Weibull
    (``extreme=True`` appends a species with a generalised-extreme-value rate).  Several ``scenarios`` share one
    hydrograph scaled through ``Prorate``; with ``facility_scenario_column=False`` the Facilities table has ONE row per
    facility and no ``Scenario`` column — the only layout with which the reference itself runs more than one flow
    scenario (it takes the ``else`` branches at stryke.py:2359 and :2941; with one row per scenario it stops at :2759)."""
    prj = c2_facility(1, iterations, days, curr_Q, baro=baro, scenario=scenarios[0])
    prj["name"] = "c3"
    pop = []
    for scen in scenarios:
        for j in range(n_species):
            sp = dict(EPRI_SPECIES[j % len(EPRI_SPECIES)])
            if j >= len(EPRI_SPECIES):
                sp["Species"] = f"{sp['Species']}_{j}"
                sp["scale"] = sp["scale"] * (1.0 + 0.1 * j)
            pop.append(_pop_row(sp, scen, fish=NAN, iterations=iterations, occur_prob=sp["occur_prob"],
                                shape=sp["shape"], location=sp["location"], scale=sp["scale"], dist=sp["dist"],
                                max_ent_rate=sp["max_ent_rate"]))
    if extreme:
        for scen in scenarios:
            sp = EXTREME_SPECIES
            pop.append(_pop_row(sp, scen, fish=NAN, iterations=iterations, occur_prob=sp["occur_prob"], shape=sp["shape"],
                                location=sp["location"], scale=sp["scale"], dist=sp["dist"], max_ent_rate=sp["max_ent_rate"]))
    prj["population"] = pop
    start = "2023-03-01"
    if days > 300:   # a full synthetic year: lognormal-ish seasonal hydrograph around curr_Q
        start = "2023-01-01"
        rs = np.random.RandomState(1)
        q = curr_Q * np.exp(0.35 * np.sin(2 * np.pi * (np.arange(days) - 60) / 365.0) + 0.25 * rs.standard_normal(days))
        import datetime as _dt
        d0 = _dt.date.fromisoformat(start)
        prj["hydrograph"] = [dict(Date=(d0 + _dt.timedelta(days=i)).isoformat(), Discharge=float(q[i])) for i in range(days)]
    prj["flow_scenarios"] = _scenario_rows(list(scenarios), months=_months_for(days, start))
    if len(scenarios) > 1:   # one hydrograph, scaled per scenario through Prorate (x0.5 .. x2)
        for r, k in zip(prj["flow_scenarios"], np.linspace(0.5, 2.0, len(scenarios))):
            r["Prorate"] = float(k)
    if facility_scenario_column:
        prj["facility_params"] = [dict(prj["facility_params"][0], Scenario=s) for s in scenarios]
    else:
        prj["facility_params"] = [{k: v for k, v in prj["facility_params"][0].items() if k != "Scenario"}]
    prj["operating_scenarios"] = [dict(r, Scenario=s) for s in scenarios for r in prj["operating_scenarios"][:3]]
    return prj


PUMP = dict(**{"Runner Type": "Pump"}, intake_vel=2.0, H=60.0, RPM=97.3, D=10.76, ada=0.94, N=13, Qopt=2150.72, Qcap=2288.0,
            B=4.0, iota=1.1, D1=11.22, D2=10.76, Q_p=1800.0, gamma=0.153)


def c4_barotrauma(n_fish=1000, iterations=1, days=1, curr_Q=9000.0, habitat="Pelagic", pump=False, friction=False):
    """6-unit station with barotrauma columns populated.  Default: 4 Propeller + 2 Francis units and
    hydrostatic pressures — what the reference's ``run()`` can do (a Pump node stops it with KeyError, Appendix E8).
    ``pump=True``: BASELINE.json's C4 as named, 3 Propeller + 3 Pump units (``Q_p``, ``gamma`` columns: opt-in Pump runner);
    ``friction=True``: the friction-loss pressure mode (``barotrauma_pressure = "friction"``)."""
    names = [f"U{i}" for i in range(1, 7)]
    kind_of = (lambda i: "Propeller" if i < 3 else "Pump") if pump else (lambda i: "Propeller" if i < 4 else "Francis")
    runner_of = (lambda i: PROPELLER if i < 3 else PUMP) if pump else (lambda i: PROPELLER if i < 4 else FRANCIS)
    nodes = ([("river_node_0", "a priori", 1.0), ("forebay", "a priori", 1.0)]
             + [(u, kind_of(i), 0.0) for i, u in enumerate(names)]
             + [("bypass_flow", "a priori", 0.98), ("spillway", "a priori", 0.97), ("tailrace", "a priori", 1.0),
                ("river_node_1", "a priori", 1.0)])
    mid = names + ["bypass_flow", "spillway"]
    edges = ([("river_node_0", "forebay", 1.0)] + [("forebay", x, 1.0) for x in mid]
             + [(x, "tailrace", 1.0) for x in mid] + [("tailrace", "river_node_1", 1.0)])
    units = [_unit_row("Sta", i + 1, u, i + 1, runner_of(i), baro=True,
                       fb_depth=30.0 + 2 * i, submergence_depth=6.0 if (i % 2 == 0 or friction) else NAN,
                       elevation_head=NAN if i % 2 == 0 else 4.0 + i)
             for i, u in enumerate(names)]
    species = dict(SHAD, vertical_habitat=habitat)
    return dict(
        name="c4", output_units="imperial", max_river_node=1, barotrauma_pressure="friction" if friction else "hydrostatic",
        nodes=[dict(Location=a, Surv_Fun=b, Survival=c) for a, b, c in nodes],
        edges=[dict(_from=a, _to=b, weight=w) for a, b, w in edges],
        unit_params=units,
        facility_params=[dict(Facility="Sta", Scenario="Spring", Operations="run-of-river",
                              **{"Rack Spacing": 1.5 / 12.0}, Min_Op_Flow=800.0, Env_Flow=0.0, Bypass_Flow=150.0,
                              Spillway="spillway")],
        flow_scenarios=_scenario_rows(["Spring"], months=_months_for(days)),
        operating_scenarios=[dict(Scenario="Spring", Facility="Sta", Unit=i + 1, Hours=24.0, Prob_Not_Op=NAN,
                                  shape=NAN, location=NAN, scale=NAN) for i in range(6)],
        hydrograph=_hydrograph(days, curr_Q * 0.6, curr_Q * 1.4),
        population=[_pop_row(species, "Spring", fish=float(n_fish), iterations=iterations)],
    )


def c5_cascade(n_fish=1000, iterations=1, days=1, curr_Q=9000.0, dams=5):
    """Cumulative multi-facility cascade: ``dams`` facilities in series.  Dam i:
    river_node_i -> forebay_i -> {DiU1, DiU2, [env_i], bypass_i, spill_i} -> tailrace_i -> river_node_{i+1}.
    With 5 dams: 40 nodes, 22 moves (exercises the lexicographic first-'U' rule)."""
    nodes, edges, units, facs, ops = [], [], [], [], []
    for i in range(dams):
        fac = f"Dam{i + 1}"
        rn, fb, tr = f"river_node_{i}", f"forebay_{i}", f"tailrace_{i}"
        u1, u2 = f"D{i + 1}U1", f"D{i + 1}U2"
        mid = [u1, u2]
        nodes += [(rn, "a priori", 1.0), (fb, "a priori", 1.0),
                  (u1, "Kaplan" if i % 2 == 0 else "Francis", 0.0), (u2, "Propeller" if i % 2 == 0 else "Kaplan", 0.0)]
        has_env = i < dams - 1
        if has_env:
            nodes.append((f"env_{i}", "a priori", 0.99)); mid.append(f"env_{i}")
        nodes += [(f"bypass_{i}", "a priori", 0.98), (f"spill_{i}", "a priori", 0.97), (tr, "a priori", 1.0)]
        mid += [f"bypass_{i}", f"spill_{i}"]
        edges += [(rn, fb, 1.0)] + [(fb, x, 1.0) for x in mid] + [(x, tr, 1.0) for x in mid] + [(tr, f"river_node_{i + 1}", 1.0)]
        r1 = KAPLAN if i % 2 == 0 else FRANCIS
        r2 = PROPELLER if i % 2 == 0 else KAPLAN
        units += [_unit_row(fac, 1, u1, 1, r1), _unit_row(fac, 2, u2, 2, r2)]
        facs.append(dict(Facility=fac, Scenario="Spring", Operations="run-of-river", **{"Rack Spacing": 2 / 12.0},
                         Min_Op_Flow=400.0 + 50 * i, Env_Flow=200.0 if has_env else 0.0, Bypass_Flow=100.0 + 10 * i,
                         Spillway=f"spill_{i}"))
        ops += [dict(Scenario="Spring", Facility=fac, Unit=u, Hours=24.0, Prob_Not_Op=NAN, shape=NAN, location=NAN,
                     scale=NAN) for u in (1, 2)]
    nodes.append((f"river_node_{dams}", "a priori", 1.0))
    return dict(
        name="c5", output_units="imperial", max_river_node=dams,
        nodes=[dict(Location=a, Surv_Fun=b, Survival=c) for a, b, c in nodes],
        edges=[dict(_from=a, _to=b, weight=w) for a, b, w in edges],
        unit_params=units, facility_params=facs, flow_scenarios=_scenario_rows(["Spring"], months=_months_for(days)),
        operating_scenarios=ops, hydrograph=_hydrograph(days, curr_Q * 0.6, curr_Q * 1.4),
        population=[_pop_row(SHAD, "Spring", fish=float(n_fish), iterations=iterations)],
    )


def random_facility(seed=0, n_fish=1000, iterations=1, days=4, baro=False):
    """Randomised cascade for fuzzing the device path against the oracle: 1-3 dams in series, 1-7 units per dam with
    random runner types, capacities, operating orders and minimum flows, optional environmental-flow nodes, random
    a-priori survivals, and a hydrograph that spans from (nearly) no flow — nothing runs, fish may stay put — to far
    above the station capacities (everything spills)."""
    rng = np.random.default_rng(seed)
    dams = int(rng.integers(1, 4))
    nodes, edges, units, facs, ops = [], [], [], [], []
    runners = (KAPLAN, FRANCIS, PROPELLER)
    total_cap = 0.0
    for i in range(dams):
        fac = f"Dam{i + 1}"
        rn, fb, tr = f"river_node_{i}", f"forebay_{i}", f"tailrace_{i}"
        n_units = int(rng.integers(1, 8))
        names = [f"D{i + 1}U{j + 1}" for j in range(n_units)]
        order = rng.permutation(n_units) + 1
        mid = list(names)
        nodes += [(rn, "a priori", 1.0), (fb, "a priori", float(rng.choice([1.0, 1.0, 0.995])))]
        cap = 0.0
        for j, nm in enumerate(names):
            r = dict(runners[int(rng.integers(0, 3))])
            scale = float(rng.uniform(0.3, 1.6))
            r.update(Qcap=round(r["Qcap"] * scale, 1), Qopt=round(r["Qopt"] * scale, 1), H=float(rng.uniform(15, 70)))
            cap += r["Qcap"]
            nodes.append((nm, r["Runner Type"], 0.0))
            units.append(_unit_row(fac, j + 1, nm, int(order[j]), r, baro))
            ops.append(dict(Scenario="Spring", Facility=fac, Unit=j + 1, Hours=24.0, Prob_Not_Op=NAN, shape=NAN,
                            location=NAN, scale=NAN))
        has_env, has_byp = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        if has_env:
            nodes.append((f"env_{i}", "a priori", float(rng.uniform(0.95, 1.0)))); mid.append(f"env_{i}")
        if has_byp:
            nodes.append((f"bypass_{i}", "a priori", float(rng.choice([1.0, 0.98])))); mid.append(f"bypass_{i}")
        nodes += [(f"spill_{i}", "a priori", float(rng.uniform(0.9, 1.0))), (tr, "a priori", float(rng.choice([1.0, 0.99])))]
        mid.append(f"spill_{i}")
        edges += [(rn, fb, 1.0)] + [(fb, x, 1.0) for x in mid] + [(x, tr, 1.0) for x in mid] + [(tr, f"river_node_{i + 1}", 1.0)]
        facs.append(dict(Facility=fac, Scenario="Spring", Operations="run-of-river",
                         **{"Rack Spacing": float(rng.choice([1.0, 2.0, 0.35])) / 12.0},
                         Min_Op_Flow=float(rng.uniform(100, 900)), Env_Flow=float(rng.uniform(50, 400)) if has_env else 0.0,
                         Bypass_Flow=float(rng.uniform(20, 200)) if has_byp else 0.0, Spillway=f"spill_{i}"))
        total_cap = max(total_cap, cap)
    nodes.append((f"river_node_{dams}", "a priori", 1.0))
    import datetime as _dt
    d0 = _dt.date(2023, 3, 1)
    levels = np.concatenate([[float(rng.choice([0.0, 30.0, 150.0]))], rng.uniform(0.05, 2.5, max(days - 1, 0)) * total_cap])
    flows = levels[rng.permutation(days)]
    return dict(
        name=f"random{seed}", output_units="imperial", max_river_node=dams,
        nodes=[dict(Location=a, Surv_Fun=b, Survival=c) for a, b, c in nodes],
        edges=[dict(_from=a, _to=b, weight=w) for a, b, w in edges],
        unit_params=units, facility_params=facs, flow_scenarios=_scenario_rows(["Spring"], months=_months_for(days)),
        operating_scenarios=ops,
        hydrograph=[dict(Date=(d0 + _dt.timedelta(days=i)).isoformat(), Discharge=float(flows[i])) for i in range(days)],
        population=[_pop_row(SHAD, "Spring", fish=float(n_fish), iterations=iterations)],
    )


def c1_single_unit(iterations=5, days=92, seed=0, fish=NAN):
    """the v250304 example — one Francis unit node, Micropterus Log Normal entrainment,
    barotrauma on (fb_depth / submergence set), seeded synthetic hydrograph (the USGS gage needs network)."""
    rs = np.random.RandomState(seed)
    q = 3000.0 * np.exp(0.5 * rs.standard_normal(days))
    import datetime as _dt
    d0 = _dt.date(2020, 3, 1)
    hyd = [dict(Date=(d0 + _dt.timedelta(days=i)).isoformat(), Discharge=float(q[i])) for i in range(days)]
    unit = _unit_row("Cabot", 1, "U1 Cabot", 1, dict(FRANCIS, H=60.0), fb_depth=10.0, submergence_depth=2.0,
                     ps_D=NAN, ps_length=NAN, roughness=NAN, elevation_head=NAN)
    sp = EPRI_SPECIES[1]
    pop = _pop_row(sp, "Spring", fish=fish, iterations=iterations, occur_prob=sp["occur_prob"])
    if isinstance(fish, float) and math.isnan(fish):
        pop.update(shape=sp["shape"], location=sp["location"], scale=sp["scale"], dist=sp["dist"],
                   max_ent_rate=sp["max_ent_rate"])
    return dict(
        name="c1", output_units="imperial", max_river_node=0,
        nodes=[dict(Location="U1 Cabot", Surv_Fun="Francis", Survival=0.0)],
        edges=[],
        unit_params=[unit],
        facility_params=[dict(Facility="Cabot", Scenario="Spring", Operations="run-of-river",
                              **{"Rack Spacing": 2 / 12.0}, Min_Op_Flow=300.0, Env_Flow=0.0, Bypass_Flow=0.0,
                              Spillway="none")],
        flow_scenarios=_scenario_rows(["Spring"], months="3,4,5", year=2020),
        operating_scenarios=[dict(Scenario="Spring", Facility="Cabot", Unit=1, Hours=24.0, Prob_Not_Op=NAN,
                                  shape=NAN, location=NAN, scale=NAN)],
        hydrograph=hyd, population=[pop],
    )


def peaking_facility(n_fish=200, iterations=2, days=3, curr_Q=9000.0):
    """C2's facility operated as a peaking plant: per-unit not-operating coin + lognormal hours
    (reference: Stryke/stryke.py:2405-2431; Pumped_Storage_Bad_Creek example)."""
    prj = c2_facility(n_fish, iterations, days, curr_Q)
    prj["name"] = "peaking"
    prj["facility_params"][0]["Operations"] = "peaking"
    for i, r in enumerate(prj["operating_scenarios"]):
        r.update(Hours=NAN, Prob_Not_Op=0.35 + 0.2 * i, shape=0.4, location=0.0, scale=9.0 + i)
    return prj


def apply_to(sim, prj, proj_dir="/tmp/stryke_b200_project", output_name="project"):
    """Populate a ``simulation.__new__(simulation)`` object (reference's or ours) from a project dict, the way the
    reference's unit tests do (attribute names: SURVEY.md Appendix D)."""
    import os
    import pandas as pd
    sim.proj_dir, sim.output_name = proj_dir, output_name
    sim.hdf_path = os.path.join(proj_dir, output_name + ".h5")
    sim.wks_dir = None
    sim.output_units = prj["output_units"]
    sim.nodes = pd.DataFrame(prj["nodes"], columns=["Location", "Surv_Fun", "Survival"])
    sim.edges = pd.DataFrame(prj["edges"], columns=["_from", "_to", "weight"])
    sim.surv_fun_dict = sim.nodes[["Location", "Surv_Fun"]].set_index("Location").to_dict("index")
    sim.max_river_node = prj["max_river_node"]
    sim.unit_params = pd.DataFrame(prj["unit_params"]).set_index("Unit_Name")
    sim.facility_params = pd.DataFrame(prj["facility_params"]).set_index("Facility")
    sim.flow_scenarios_df = pd.DataFrame(prj["flow_scenarios"])
    sim.flow_scenarios = sim.flow_scenarios_df.Scenario.unique()
    sim.operating_scenarios_df = pd.DataFrame(prj["operating_scenarios"])
    sim.input_hydrograph_df = pd.DataFrame(prj["hydrograph"])
    sim.pop = pd.DataFrame(prj["population"])
    if prj.get("barotrauma_pressure", "hydrostatic") != "hydrostatic":      # (our class only; the reference has no such switch)
        sim.barotrauma_pressure = prj["barotrauma_pressure"]
    return sim


def make_sim(prj, name="case", proj_dir="/tmp/stryke_b200_workloads"):
    """A ``stryke_b200.simulation`` carrying the project's tables (no files are read or written)."""
    from .simulation import simulation
    sim = simulation.__new__(simulation)
    apply_to(sim, prj, proj_dir=proj_dir, output_name=name)
    return sim


def make_plan(prj, name="case", quiet=True):
    """(simulation, runner.RunPlan) of a project: routes, per-(scenario, day) routing / strike tables, the cell grid."""
    import contextlib
    import io
    from .runner import RunPlan
    sim = make_sim(prj, name)
    with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
        plan = RunPlan(sim)
    return sim, plan
