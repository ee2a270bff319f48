"""Project fixtures of the oracle harness and the tests.

TEST INFRASTRUCTURE.  The synthetic C1-C5 projects live in ``stryke_b200/workloads.py`` (bench.py and smoke() use them too)
and are re-exported here under the names the committed goldens refer to; the two fixtures below read the tables extracted
from the reference's own example workbooks (tests/golden/*_project.json, oracle/extract_example.py).
"""
from __future__ import annotations

import math

import numpy as np

from stryke_b200.workloads import (EPRI_SPECIES, EXTREME_SPECIES, FRANCIS, KAPLAN, NAN, PROPELLER, SHAD, _hydrograph,  # noqa: F401
                                   _months_for, _pop_row, _scenario_rows, _unit_row, apply_to, c1_single_unit, c2_facility,
                                   c3_population, c4_barotrauma, c5_cascade, peaking_facility, random_facility)


def example_v250304(iterations=2, days=30, seed=0, median_Q=3000.0, fish=NAN):
    """BASELINE.json configs[0]: the reference's example project Spreadsheet Interface/Input_Spreadsheet_v250304.xlsx —
    one node, one Francis unit (U1 Cabot), barotrauma on (fb_depth 10 ft, elevation_head 2 ft), Micropterus with a Log
    Normal entrainment rate — with the tables exactly as the workbook holds them (tests/golden/
    example_v250304_project.json, oracle/extract_example.py) and the synthetic hydrograph SURVEY.md §8(d) prescribes in
    place of the USGS gage (Q = 3000 exp(0.5 N(0,1)), seeded)."""
    return cabot_station(iterations, days, seed, median_Q, fish, source="example_v250304_project.json", name="example_v250304")


def cabot_station(iterations=2, days=6, seed=0, median_Q=9000.0, fish=NAN, source="cabot_station_project.json",
                  name="cabot_station"):
    """The reference's own complex-network example (Spreadsheet Interface/Examples/
    Run_of_River_Complex_Network_Cabot_Station.xlsx: 20 nodes, 31 edges, 11 Francis units in two facilities that share
    one spillway, a 0.05 / 0.95 edge-weight split at the canal, Micropterus with a Log Normal entrainment rate).  The
    tables come from tests/golden/cabot_station_project.json (written by oracle/extract_example.py); the USGS gage the
    workbook names needs the network, so the hydrograph is a seeded synthetic one (Q = median * exp(0.5 N(0,1)))."""
    import datetime as _dt
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", source)
    with open(path) as fh:
        raw = json.load(fh)
    nan = lambda v: NAN if v is None else v
    ordinal, units = {}, []
    for r in raw["unit_params"]:
        fac = r["Facility"]
        ordinal[fac] = ordinal.get(fac, 0) + 1
        row = {k: nan(v) for k, v in r.items()}
        row.update(Unit_Name=r["Unit"], Unit=ordinal[fac])
        units.append(row)
    number = {(u["Facility"], u["Unit_Name"]): u["Unit"] for u in units}
    ops = [dict({k: nan(v) for k, v in r.items() if k != "OpScenario"}, Unit=number[(r["Facility"], r["Unit"])])
           for r in raw["operating_scenarios"]]
    d0 = _dt.date(2020, 3, 1)
    q = median_Q * np.exp(0.5 * np.random.default_rng(seed).standard_normal(days))
    scen = raw["flow_scenarios"][0]
    pop = {k: nan(v) for k, v in raw["population"][0].items() if k not in ("Notes", "N=")}
    pop.update(Iterations=iterations)
    pop.setdefault("vertical_habitat", "Benthic")
    pop.setdefault("beta_0", NAN)
    pop.setdefault("beta_1", NAN)
    if not (isinstance(fish, float) and math.isnan(fish)):     # fixed count per day instead of the fitted entrainment rate
        pop.update(Fish=float(fish), shape=NAN, location=NAN, scale=NAN, dist=NAN, max_ent_rate=NAN, occur_prob=1.0)
    max_rn = max([int(n["Location"].split("_")[2]) for n in raw["nodes"] if n["Location"].startswith("river_node_")] + [0])
    return dict(
        name=name, output_units=raw["output_units"], max_river_node=max_rn,
        nodes=[dict(n) for n in raw["nodes"]],
        edges=[dict(e) for e in raw["edges"]],
        unit_params=units,
        facility_params=[{k: nan(v) for k, v in f.items()} for f in raw["facility_params"]],
        flow_scenarios=[dict({k: nan(v) for k, v in scen.items()}, Flow="hydrograph", Gage="", FlowYear=2020, Prorate=1.0,
                             Months=_months_for(days, "2020-03-01"))],
        operating_scenarios=ops,
        hydrograph=[dict(Date=(d0 + _dt.timedelta(days=i)).isoformat(), Discharge=float(q[i])) for i in range(days)],
        population=[pop],
    )


