"""High-power statistical comparison of the native-RNG throughput kernel with the NumPy oracle: tens of millions of fish
on both sides (the oracle fanned out over the host cores), z-scores of every tallied proportion.  Detects biases of the
order of 1e-4 in survival / routing / cause probabilities (e.g. from the 16-bit r/R, cause and Box-Muller inputs).
    python profiles/high_power_check.py [fixture] [fish per cell] [cells per process]"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def project(fixture, n_fish, iterations):
    from oracle import fixtures
    kw = dict(n_fish=n_fish, iterations=iterations, days=1)
    if fixture == "c2b":
        return fixtures.c2_facility(baro=True, **kw)
    if fixture == "c5":
        return fixtures.c5_cascade(**kw)
    if fixture == "c4":
        return fixtures.c4_barotrauma(**kw)
    return fixtures.c2_facility(**kw)


def worker(args):
    fixture, n_fish, iterations, seed = args
    from oracle import stryke_oracle as O
    out = O.run_project(project(fixture, n_fish, iterations), seed=seed)
    daily = np.sum([r["daily"] for r in out], axis=0)
    sd = {}
    for r in out:
        for k, (s, c) in r["state_daily"].items():
            a = sd.setdefault(k, [0, 0])
            a[0] += int(s); a[1] += int(c)
    return daily, sd


def main():
    fixture = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_fish = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
    per_proc = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    procs = os.cpu_count() or 1
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(worker, [(fixture, n_fish, per_proc, 1000 + i) for i in range(procs)])
    o_daily = np.sum([p[0] for p in parts], axis=0)
    o_sd = {}
    for _, sd in parts:
        for k, (s, c) in sd.items():
            a = o_sd.setdefault(k, [0, 0])
            a[0] += s; a[1] += c
    import gpu_util as G
    from stryke_b200 import engine
    prj = project(fixture, n_fish, per_proc * procs)
    sim, plan = G.make_plan(prj, "hp")
    res = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id,
                           precision=32, seed=20241, max_daily_fish=2_000_000_000)
    g_daily = G.full_daily(int(res.cell_n.astype(np.int64).sum()), res.daily.astype(np.int64).sum(axis=0))
    g_sd = {}
    for c in range(plan.n_cells):
        for k, (s, cnt) in G.state_daily_dict(plan, res.state_daily[c]).items():
            a = g_sd.setdefault(k, [0, 0])
            a[0] += int(s); a[1] += int(cnt)
    n_tot = int(o_daily[0])
    assert n_tot == int(g_daily[0]), (n_tot, g_daily[0])
    zs = []

    def z2(s1, n1, s2, n2):
        p = (s1 + s2) / (n1 + n2)
        if p <= 0 or p >= 1:
            return 0.0
        return (s1 / n1 - s2 / n2) / np.sqrt(p * (1 - p) * (1 / n1 + 1 / n2))
    for j, nm in ((1, "entrained"), (2, "survived"), (4, "impingement"), (5, "blade_strike"), (6, "barotrauma")):
        zs.append((f"daily {nm}", z2(g_daily[j], n_tot, o_daily[j], n_tot), g_daily[j] / n_tot, o_daily[j] / n_tot))
    for k in sorted(g_sd):
        (s1, c1), (s2, c2) = g_sd[k], o_sd[k]
        zs.append((f"arrive {k}", z2(c1, n_tot, c2, n_tot), c1 / n_tot, c2 / n_tot))
        if c1 and c2:
            zs.append((f"survive {k}", z2(s1, c1, s2, c2), s1 / c1, s2 / c2))
    zs.sort(key=lambda t: -abs(t[1]))
    print(f"{fixture}: {n_tot} fish on each side ({procs} oracle processes); {len(zs)} proportions; largest |z|:")
    for name, z, a, b in zs[:8]:
        print(f"  {name:34s} z = {z:+.2f}   device {a:.6f}   oracle {b:.6f}   diff {a - b:+.2e}")


if __name__ == "__main__":
    main()
