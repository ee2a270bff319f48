"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference under ``oracle/harness.py``.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_goldens``
Each file holds, per (scenario, species, iteration, day) cell: the slot-tagged draws the reference consumed, and
the reference's own outputs (Daily row, State_Daily rows, per-fish states / fp32 rates / survival flags), plus the
routing tables captured from the reference's ``movement()`` and the Route_Flows / Unit_Hours frames.
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from . import fixtures, harness, replay, stryke_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (fixture function name, kwargs, seed, {species: family})
    "c2_kaplan_francis": ("c2_facility", dict(n_fish=250, iterations=2, days=2), 1, {}),
    "c2_barotrauma": ("c2_facility", dict(n_fish=250, iterations=1, days=2, baro=True), 2, {}),
    "c4_propeller_baro": ("c4_barotrauma", dict(n_fish=250, iterations=1, days=2, habitat="Benthic"), 3, {}),
    "c5_cascade": ("c5_cascade", dict(n_fish=200, iterations=1, days=2), 4, {}),
    "c1_single_unit": ("c1_single_unit", dict(iterations=2, days=60), 5, {"Micropterus": "Centrarchidae"}),
    "c3_population": ("c3_population", dict(n_species=4, iterations=2, days=8), 6,
                      {"Ambloplites": "Centrarchidae", "Micropterus": "Centrarchidae", "Lepomis": "Centrarchidae"}),
    "c3_extreme": ("c3_population", dict(n_species=1, iterations=3, days=10, extreme=True), 11,
                   {"Ambloplites": "Centrarchidae", "Perca": "Percidae"}),
    "c2_normal_lengths": ("c2_facility", dict(n_fish=200, iterations=1, days=2, length_normal=(11.0, 3.5)), 12, {}),
    "c2_penstock": ("c2_facility", dict(n_fish=200, iterations=1, days=3, curr_Q=6500.0, penstock=True), 13, {}),
    "c2_zero_flow": ("c2_facility", dict(n_fish=80, iterations=1, days=4, flows=[0.0, 250.0, 9000.0, 0.0]), 14, {}),
    # several flow scenarios (BASELINE.json configs[2] "x 5 flow scenarios"): the Facilities table has no Scenario column, the
    # layout with which the reference runs them (else-branches stryke.py:2359, :2941; close() neutralised, Appendix E1).  With
    # one Facilities row per scenario — the layout of its own Bad Creek example — it stops at stryke.py:2759 (float(Series)).
    "c3_multi_scenario": ("c3_population", dict(n_species=3, iterations=2, days=5, scenarios=["Spring", "Summer", "Fall"],
                                                facility_scenario_column=False), 21,
                          {"Ambloplites": "Centrarchidae", "Micropterus": "Centrarchidae"}),
    "c2_low_flow": ("c2_facility", dict(n_fish=120, iterations=1, days=3, curr_Q=600.0), 7, {}),
    # the reference's own complex-network example workbook (tables extracted by oracle/extract_example.py)
    "cabot_station": ("cabot_station", dict(iterations=2, days=24, seed=0), 8, {"Micropterus": "Centrarchidae"}),
    # BASELINE.json configs[0]: the example project workbook itself (one Francis unit, barotrauma on)
    "example_v250304": ("example_v250304", dict(iterations=2, days=45, seed=0), 10, {"Micropterus": "Centrarchidae"}),
    "cabot_station_fish": ("cabot_station", dict(iterations=1, days=4, seed=3, median_Q=7000.0, fish=300.0), 9,
                           {"Micropterus": "Centrarchidae"}),
}


PEAKING = {   # written to tests/golden/peaking/ (own test: the hours of every cell are injected too)
    "peaking_c2": ("peaking_facility", dict(n_fish=150, iterations=3, days=5), 31, {}),
}


def build_case(name):
    fn, kw, seed, fam = (CASES.get(name) or PEAKING[name])
    prj = getattr(fixtures, fn)(**kw)
    sim, store, rec = harness.run_reference(prj, seed=seed)
    model = O.Model.from_project(prj)
    node_idx = {nd: i for i, nd in enumerate(model.nodes)}
    daily = store["Daily"]
    sd = store["State_Daily"] if "State_Daily" in store else None
    simtabs = {k: store[k] for k in store.keys() if "simulations" in k}
    pops = {(r["Species"], r["Scenario"]): r for r in prj["population"]}
    M = len(model.moves)
    out = {}
    cells_meta = []
    # peaking plants: a (species, iteration, day) whose units all sat idle is skipped before anything is written
    # (stryke.py:3025): the cell is kept in the list (its hours are part of the recording) but has no Daily row
    idle = [c["tot_hours"] == 0 and not c["moves"] and c["presence"] is None for c in rec.cells]
    assert len(rec.cells) - sum(idle) == len(daily)
    di = -1
    last = None
    for ci, cell in enumerate(rec.cells):
        if idle[ci]:
            cells_meta.append(dict(last or {}, idle=True, curr_Q=cell["curr_Q"], presence=None, tot_hours=0.0, hours=cell["hours"],
                                   count_draw=None, n=0, iteration=None, day=None))
            continue
        di += 1
        drow = daily.iloc[di]
        spc, scen = drow["species"].strip(), drow["scenario"].strip()
        prow = pops[(spc, scen)]
        p = f"c{ci}_"
        cm = dict(scenario=scen, species=spc, iteration=int(drow["iteration"]), day=str(drow["day"]),
                  curr_Q=cell["curr_Q"], presence=cell["presence"], tot_hours=cell["tot_hours"], hours=cell["hours"],
                  count_draw=replay.cell_count_draw(cell), n=(cell["moves"][0]["n"] if cell["moves"] else 0))
        cells_meta.append(cm)
        last = dict(scenario=scen, species=spc)
        out[p + "daily"] = drow.iloc[6:].to_numpy(dtype=float).astype(np.int64)
        if not cell["moves"]:
            continue
        for k, v in replay.pack_cell(cell, prow).items():
            out[p + k] = v
        sub = sd[(sd.scenario == scen) & (sd.species == spc) & (sd.iteration == drow["iteration"]) & (sd.day == drow["day"])]
        out[p + "sd_move"] = sub["move"].to_numpy(dtype=np.int32)
        out[p + "sd_state"] = np.array([node_idx[s] for s in sub["state"]], dtype=np.int32)
        out[p + "sd_vals"] = sub[["successes", "count"]].to_numpy(dtype=np.float64)
        f = simtabs[f"/simulations/{scen}/{spc}"]
        f = f[(f.iteration == drow["iteration"]) & (f.day == drow["day"])]
        out[p + "states"] = np.stack([np.array([node_idx[s] for s in f[f"state_{k}"].astype(str)], dtype=np.int16) for k in range(M)])
        out[p + "rates"] = np.stack([f[f"rates_{k}"].to_numpy(dtype=np.float32) for k in range(M)])
        out[p + "survival"] = np.stack([f[f"survival_{k}"].to_numpy(dtype=np.uint8) for k in range(M)])
    routing = [dict(cell=c, location=loc, locs=locs, probs=[float(x) for x in p]) for (c, loc), (locs, p) in rec.routing.items()]

    def frame(key):
        if key not in store:
            return []
        df = store[key].copy()
        df["day"] = df["day"].astype(str)
        return json.loads(df.to_json(orient="records"))
    meta = dict(case=name, fixture=fn, kwargs=kw, seed=seed, families=fam, cells=cells_meta, routing=routing,
                route_flows=frame("Route_Flows"), unit_hours=frame("Unit_Hours"), nodes=model.nodes, moves=M,
                numpy=np.__version__, reference="knebiolo/stryke @ /root/reference (unmodified, oracle/harness.py)")
    out["meta"] = np.array(json.dumps(meta))
    out_dir = os.path.join(GOLDEN_DIR, "peaking") if name in PEAKING else GOLDEN_DIR
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **out)
    return path, len(rec.cells)


def main(argv):
    names = argv or list(CASES)
    for nm in names:
        path, nc = build_case(nm)
        print(f"{nm}: {nc} cells -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main(sys.argv[1:])
