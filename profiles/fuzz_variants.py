"""One-off fuzzing beyond the test-suite: many random facilities, every kernel variant must return the same tallies and
the fp64 injected replay must equal the oracle.   python profiles/fuzz_variants.py <first seed> <last seed>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fixtures, stryke_oracle as O  # noqa: E402
import gpu_util as G  # noqa: E402
from stryke_b200 import engine  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    prj = fixtures.random_facility(seed, n_fish=int(3000 + 997 * (seed % 7)), iterations=2, days=5, baro=bool(seed % 3 == 0))
    sim, plan = G.make_plan(prj, f"fz{seed}")
    args = (plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id)
    ref = None
    for variant in (1, 2, 3):
        r = engine.run_cells(*args, precision=32, seed=seed, max_daily_fish=10 ** 7, kernel_variant=variant)
        if ref is None:
            ref = r
        elif not (np.array_equal(r.daily, ref.daily) and np.array_equal(r.state_daily, ref.state_daily) and np.array_equal(r.cell_n, ref.cell_n)):
            print("VARIANT MISMATCH seed", seed, "variant", variant, flush=True)
            bad += 1
    # fp64 injected vs oracle
    model = O.Model.from_project(prj)
    ref_idx = G.reference_cell_index(plan)
    rng = np.random.default_rng(seed)
    prow = prj["population"][0]
    sp = O.species_record(model, prow, None)
    M, n, days = len(model.moves), int(prow["Fish"]), len(prj["hydrograph"])
    draws, results = [], []
    for r_ in range(plan.n_cells):
        z = rng.standard_normal(n)
        L = O.lengths_from_normals(z, prow["length shape"], prow["length location"], prow["length scale"])
        d = O.Draws.random(rng, M, n, L, ghost=True)
        draws.append((d, z))
        results.append(O.simulate_cell(model, sp, prj["hydrograph"][r_ % days]["Discharge"], prow["Scenario"], d))
    inj, off = G.injected_from_cells(plan, ref_idx, draws, [0.0] * plan.n_cells, [None] * plan.n_cells, M)
    res = G.run_injected(plan, inj, precision=64)
    for c, r_ in enumerate(ref_idx):
        if not (np.array_equal(G.full_daily(res.cell_n[c], res.daily[c]), results[r_].daily)
                and G.state_daily_dict(plan, res.state_daily[c]) == results[r_].state_daily):
            print("ORACLE MISMATCH seed", seed, "cell", c, flush=True)
            bad += 1
print("fuzzed seeds", lo, "..", hi - 1, "mismatches:", bad)
sys.exit(1 if bad else 0)
