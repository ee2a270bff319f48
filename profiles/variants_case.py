"""Small end-to-end case that launches every passage kernel variant (general fp32 / fp64, generic throughput, NVRTC
specialised) on three facilities plus the bootstrap kernel and requires identical tallies: a quick consistency run to
put under a profiler or a debugger.   python profiles/variants_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fixtures  # noqa: E402
import gpu_util as G  # noqa: E402
from stryke_b200 import engine  # noqa: E402

for name, prj in (("c2", fixtures.c2_facility(n_fish=3000, iterations=3, days=2, baro=True)),
                  ("cabot", fixtures.cabot_station(iterations=2, days=3, seed=3, median_Q=7000.0, fish=2500.0)),
                  ("c3", fixtures.c3_population(n_species=4, iterations=6, days=5))):
    sim, plan = G.make_plan(prj, "san_" + name)
    args = (plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id)
    ref = None
    for variant in (1, 2, 3):
        r = engine.run_cells(*args, precision=32, seed=5, max_daily_fish=10 ** 7, kernel_variant=variant)
        if ref is None:
            ref = r
        assert np.array_equal(r.daily, ref.daily) and np.array_equal(r.state_daily, ref.state_daily), (name, variant)
    r64 = engine.run_cells(*args, precision=64, seed=5, max_daily_fish=10 ** 7)
    print(name, "cells", plan.n_cells, "fish", int(ref.cell_n.sum()), "ok")
m = engine.bootstrap_means(np.arange(1000, dtype=float), 64, seed=1)
print("bootstrap", float(m.mean()))
