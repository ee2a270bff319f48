"""Known-answer vectors for the per-fish probability functions, computed by calling the UNMODIFIED reference functions
(Kaplan / Propeller / Francis / Pump, stryke.py:846-1127; baro_injury_prob, Stryke/barotrauma.py:139-156) on random
parameter sets.  The reference holds no such vectors itself ("strike and baro numerics: parity
unpinned"); these are the larger set the survey asks for.

Run in the build container only:  python -m oracle.make_function_goldens  ->  tests/golden/function_vectors.json
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "function_vectors.json")
P15 = ("H", "RPM", "D", "Q", "ada", "N", "Qopt", "Qper", "_lambda", "iota", "D1", "D2", "B", "Q_p", "gamma")


def main(n_sets=60, seed=7):
    S = harness.load_reference()
    sim = S.simulation.__new__(S.simulation)
    rng = np.random.default_rng(seed)
    strike = []
    for kind, fn in (("Kaplan", sim.Kaplan), ("Propeller", sim.Propeller), ("Francis", sim.Francis), ("Pump", sim.Pump)):
        for _ in range(n_sets):
            D = float(rng.uniform(2.5, 16.0))
            Qopt = float(rng.uniform(120.0, 3500.0))
            p = dict(H=float(rng.uniform(12.0, 90.0)), RPM=float(rng.uniform(70.0, 280.0)), D=D,
                     Q=float(Qopt * rng.uniform(0.35, 1.25)), ada=float(rng.uniform(0.8, 0.95)), N=int(rng.integers(3, 17)),
                     Qopt=Qopt, Qper=float(rng.uniform(0.6, 1.0)), _lambda=0.2, iota=float(rng.uniform(1.0, 1.2)),
                     D1=float(D * rng.uniform(1.0, 1.3)), D2=D, B=float(D * rng.uniform(0.25, 0.5)),
                     Q_p=float(rng.uniform(400.0, 2500.0)), gamma=float(rng.uniform(0.1, 0.2)))
            L = [float(x) for x in rng.uniform(0.05, 3.0, 6)]
            rRs = [float(x) for x in rng.uniform(0.3, 1.0, 6)]
            out = []
            orig = np.random.uniform
            for length, rR in zip(L, rRs):
                np.random.uniform = lambda low=0.0, high=1.0, size=None, _r=rR: np.array([_r])      # Kaplan's own draw (stryke.py:900)
                try:
                    out.append(float(np.ravel(fn(length, p))[0]))
                finally:
                    np.random.uniform = orig
            strike.append(dict(kind=kind, params=[float(p[k]) for k in P15], length=L, rR=rRs, survival=out))
    baro = []
    B = S.baro_injury_prob if hasattr(S, "baro_injury_prob") else None
    if B is None:
        import importlib
        B = importlib.import_module("Stryke.barotrauma").baro_injury_prob
    for _ in range(200):
        depth, tail = float(rng.uniform(0.05, 60.0)), float(rng.uniform(0.5, 15.0))
        b0, b1 = float(rng.uniform(-6.0, -2.5)), float(rng.uniform(2.0, 7.0))
        p1 = 101325.0 + 998.2 * 9.80665 * depth * 0.3048          # stryke.py:1496-1503
        p2 = 101325.0 + 998.2 * 9.80665 * tail * 0.3048
        baro.append(dict(depth_ft=depth, tail_ft=tail, beta_0=b0, beta_1=b1, injury=float(B(p1 / p2, b0, b1))))
    with open(OUT, "w") as fh:
        json.dump(dict(reference="Stryke.stryke.simulation.{Kaplan,Propeller,Francis,Pump}, Stryke.barotrauma.baro_injury_prob (unmodified)",
                       params=P15, strike=strike, baro=baro), fh)
    print(OUT, os.path.getsize(OUT) // 1024, "KiB", len(strike), "strike sets", len(baro), "baro cases")


if __name__ == "__main__":
    main()
