"""Large-sample tallies of the UNMODIFIED reference (oracle/harness.py) for the statistical parity tests of the native-RNG
kernels ("reference samples from the harness, >= 1e5 fish"): per case the reference's own Daily and
State_Daily rows for a few cells with tens of thousands of fish each.  Only the aggregated counts are kept.

Run in the build container only (minutes of CPU: the reference simulates ~2 000 fish-passages/s):
    python -m oracle.make_reference_samples   ->  tests/golden/reference_samples.json
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

from . import harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_samples.json")

CASES = {
    # name: (fixture, kwargs, seed)
    "c2": ("c2_facility", dict(n_fish=60000, iterations=1, days=2), 101),
    "c2_barotrauma": ("c2_facility", dict(n_fish=40000, iterations=1, days=2, baro=True), 102),
    "c5_cascade": ("c5_cascade", dict(n_fish=30000, iterations=1, days=2), 103),
    "cabot_station": ("cabot_station", dict(iterations=1, days=2, seed=3, median_Q=7000.0, fish=40000.0), 104),
    "c4_barotrauma": ("c4_barotrauma", dict(n_fish=30000, iterations=1, days=2), 105),
}


def main(argv):
    from . import fixtures
    out = json.load(open(OUT)) if os.path.isfile(OUT) else {}
    for name in (argv or list(CASES)):
        fn, kw, seed = CASES[name]
        prj = getattr(fixtures, fn)(**kw)
        t0 = time.time()
        sim, store, rec = harness.run_reference(prj, seed=seed, traces=False)
        daily, sd = store["Daily"], store["State_Daily"]
        cells = []
        for ci in range(len(daily)):
            d = daily.iloc[ci]
            sub = sd[(sd.iteration == d["iteration"]) & (sd.day == d["day"])]
            cells.append(dict(iteration=int(d["iteration"]), day=str(d["day"]), flow=float(d["flow"]),
                              daily=[int(v) for v in d.iloc[6:].to_numpy(dtype=float)],
                              state_daily=[[int(r.move), str(r.state), int(r.successes), int(r["count"])] for _, r in sub.iterrows()]))
        out[name] = dict(fixture=fn, kwargs=kw, seed=seed, cells=cells, seconds=round(time.time() - t0, 1),
                         reference="knebiolo/stryke @ /root/reference (unmodified, oracle/harness.py), numpy " + np.__version__)
        print(name, len(cells), "cells", out[name]["seconds"], "s", flush=True)
        with open(OUT, "w") as fh:
            json.dump(out, fh)


if __name__ == "__main__":
    main(sys.argv[1:])
