"""Samples of the UNMODIFIED reference's ``daily_hours`` for a peaking plant (per-unit not-operating coin + clipped lognormal
hours, Stryke/stryke.py:2405-2431): ``python -m oracle.make_peaking_samples`` -> tests/golden/peaking/hours_samples.npz.
Build container only (needs /root/reference).  TEST INFRASTRUCTURE: the -m gpu tests compare the device's native peaking
draws with these samples (two-sample KS on the hours of operating units, binomial on the idle frequency)."""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np

from . import fixtures, harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "peaking", "hours_samples.npz")


def main(n=20000, seed=77):
    S = harness.load_reference()
    prj = fixtures.peaking_facility(n_fish=10, iterations=1, days=1)
    sim = S.simulation.__new__(S.simulation)
    fixtures.apply_to(sim, prj)
    Q_dict = {"curr_Q": 9000.0, "min_Q": {"FacA": 500.0}, "env_Q": {"FacA": 0.0}, "bypass_Q": {"FacA": 100.0}, "sta_cap": {"FacA": 8288.0}}
    np.random.seed(seed)
    keys, rows = None, []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(n):
            tot_hours, tot_flow, hours, flows = sim.daily_hours(dict(Q_dict), "Spring")
            keys = keys or list(hours.keys())
            rows.append([hours[k] for k in keys])
    H = np.array(rows, dtype=np.float64)
    ops = prj["operating_scenarios"]
    np.savez_compressed(OUT, hours=H.astype(np.float32), keys=np.array(keys), prob_not_op=np.array([o["Prob_Not_Op"] for o in ops]),
                        shape=np.array([o["shape"] for o in ops]), location=np.array([o["location"] for o in ops]),
                        scale=np.array([o["scale"] for o in ops]), seed=seed)
    print(OUT, H.shape, "idle cells", float((H.sum(axis=1) == 0).mean()), "not operating per unit", (H == 0).mean(axis=0))


if __name__ == "__main__":
    main()
