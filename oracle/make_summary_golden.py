"""Golden for ``summary()``: the UNMODIFIED reference runs a small project under the harness, then — with the global
NumPy stream re-seeded — its own ``summary()`` (stryke.py:3520-4101).  Stored: the frames ``run()`` wrote (the input of
summary) and every frame ``summary()`` produced.  stryke_b200.summary must reproduce them from the same input under
the same seed (its NumPy bootstrap consumes the MT19937 stream exactly like the reference's loop).

Run in the build container only:  python -m oracle.make_summary_golden  ->  tests/golden/summary_*.json
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np

from . import fixtures, harness

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SUMMARY_SEED = 4242
CASES = {
    "summary_c2": ("c2_facility", dict(n_fish=150, iterations=4, days=5), 21),
    "summary_c3": ("c3_population", dict(n_species=3, iterations=5, days=12), 22),
}
RUN_KEYS = ("Daily", "State_Daily", "Route_Flows", "Unit_Hours")
SUMMARY_KEYS = ("Daily_Summary", "Beta_Distributions", "Beta_Distributions_Units", "Yearly_Summary", "Driver_Diagnostics")


def _frame(df):
    d = df.copy()
    for c in d.columns:
        if str(d[c].dtype).startswith("datetime"):
            d[c] = d[c].astype(str)
    return json.loads(d.to_json(orient="split", double_precision=15))


def build(name):
    fn, kw, seed = CASES[name]
    prj = getattr(fixtures, fn)(**kw)
    sim, store, rec = harness.run_reference(prj, seed=seed, traces=False, summary=False)
    run_frames = {k: _frame(store[k]) for k in RUN_KEYS if k in store}
    for key, df in [("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
                    ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges), ("Unit_Parameters", sim.unit_params),
                    ("Facilities", sim.facility_params), ("Hydrograph", sim.input_hydrograph_df)]:
        store[key] = df
    np.random.seed(SUMMARY_SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.summary()
    out = dict(fixture=fn, kwargs=kw, seed=seed, summary_seed=SUMMARY_SEED, run=run_frames,
               summary={k: _frame(store[k]) for k in SUMMARY_KEYS if k in store},
               cum_sum=_frame(sim.cum_sum) if hasattr(sim, "cum_sum") else None)
    path = os.path.join(GOLDEN_DIR, name + ".json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    return path


if __name__ == "__main__":
    for nm in CASES:
        p = build(nm)
        print(nm, os.path.getsize(p) // 1024, "KiB")
