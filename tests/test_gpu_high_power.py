"""High-power statistical check at the sizes the product is used at: 1e10 fish per side.

Side A: the throughput kernel (fp32, compact word budget: 16-bit centred uniforms for r/R, cause and the Box-Muller
inputs, 12/13-bit ones at barotrauma levels, ONE cause uniform).  Side B: the fp64 general kernel on its *wide* stream
(``rng_mode = 1``): every draw a 53-bit uniform of its own, Box-Muller from two of them, the reference's sequential cause
rolls (stryke.py:1545-1560) and the reference's operation order.  Different seeds, so the two sides are independent
samples; every tallied proportion — arrivals and survival per (move, node), the Daily columns — must agree within
|z| < 4 (alpha = 6e-5 per proportion, < 1 % for the ~100 proportions of the largest case).  A bias of 1e-5 in any
probability of order 0.05 would show as |z| > 4.6 here."""
import numpy as np
import pytest

from oracle import fixtures
import gpu_util as G
from stryke_b200 import engine

pytestmark = pytest.mark.gpu

CASES = {
    "c2": ("c2_facility", dict(n_fish=1_000_000, iterations=1000, days=10)),
    "c2_baro": ("c2_facility", dict(n_fish=1_000_000, iterations=1000, days=10, baro=True)),
    "c4": ("c4_barotrauma", dict(n_fish=1_000_000, iterations=1000, days=10)),
    "c4_pump_friction": ("c4_barotrauma", dict(n_fish=1_000_000, iterations=1000, days=10, pump=True, friction=True)),
    "c5": ("c5_cascade", dict(n_fish=1_000_000, iterations=1000, days=10)),
}


def totals(res, n_days, iterations):
    """Sums over the iterations, per day (engine order is day-major)."""
    n = res.cell_n.astype(np.int64).reshape(n_days, iterations).sum(axis=1)
    d = res.daily.astype(np.int64).reshape(n_days, iterations, -1).sum(axis=1)
    s = res.state_daily.astype(np.int64).reshape(n_days, iterations, res.state_daily.shape[1], 2).sum(axis=1)
    return n, d, s


def z_two(k1, n1, k2, n2):
    p = (k1 + k2) / np.maximum(n1 + n2, 1)
    se = np.sqrt(np.maximum(p * (1 - p) * (1.0 / np.maximum(n1, 1) + 1.0 / np.maximum(n2, 1)), 1e-300))
    z = (k1 / np.maximum(n1, 1) - k2 / np.maximum(n2, 1)) / se
    return np.where((n1 > 0) & (n2 > 0) & (k1 + k2 > 50) & ((n1 + n2) - (k1 + k2) > 50), z, 0.0)


@pytest.mark.parametrize("case", sorted(CASES))
def test_throughput_kernel_against_the_wide_fp64_stream_at_1e10_fish(case):
    fixture, kw = CASES[case]
    prj = getattr(fixtures, fixture)(**kw)
    sim, plan = G.make_plan(prj, "hp_" + case)
    cells = (plan.grid, None, None)
    fast = engine.run_cells(plan.fac, plan.days, plan.species_table, *cells, precision=32, seed=0xA11CE, max_daily_fish=10 ** 9)
    wide = engine.run_cells(plan.fac, plan.days, plan.species_table, *cells, precision=64, seed=0xB0B, max_daily_fish=10 ** 9,
                            rng_mode=1)
    D, I = kw["days"], kw["iterations"]
    n1, d1, s1 = totals(fast, D, I)
    n2, d2, s2 = totals(wide, D, I)
    assert n1.sum() == n2.sum() == 10 ** 10
    zs = []
    # per day (the day's discharge sets routing and strike): arrivals per (move, node) among all fish, survival among arrivals
    zs.append(z_two(s1[..., 1], n1[:, None], s2[..., 1], n2[:, None]))
    zs.append(z_two(s1[..., 0], s1[..., 1], s2[..., 0], s2[..., 1]))
    # Daily: entrained of all, survived of entrained, the three cause counters per fish
    zs.append(z_two(d1[:, 0], n1, d2[:, 0], n2))
    zs.append(z_two(d1[:, 1], d1[:, 0], d2[:, 1], d2[:, 0]))
    for j in (2, 3, 4):
        zs.append(z_two(d1[:, j], n1, d2[:, j], n2))
    z = np.concatenate([np.ravel(x) for x in zs])
    n_tests = int((z != 0).sum())
    assert n_tests > 50
    # per-day proportions: hundreds of tests -> compare with the Gaussian extreme for that many, and pool over the days for power
    assert np.abs(z).max() < 4.5, (case, float(np.abs(z).max()), n_tests)
    zp = []
    zp.append(z_two(s1[..., 1].sum(0), n1.sum(), s2[..., 1].sum(0), n2.sum()))
    zp.append(z_two(s1[..., 0].sum(0), s1[..., 1].sum(0), s2[..., 0].sum(0), s2[..., 1].sum(0)))
    zp.append(np.atleast_1d(z_two(d1[:, 0].sum(), n1.sum(), d2[:, 0].sum(), n2.sum())))
    zp.append(np.atleast_1d(z_two(d1[:, 1].sum(), d1[:, 0].sum(), d2[:, 1].sum(), d2[:, 0].sum())))
    for j in (2, 3, 4):
        zp.append(np.atleast_1d(z_two(d1[:, j].sum(), n1.sum(), d2[:, j].sum(), n2.sum())))
    zp = np.concatenate([np.ravel(x) for x in zp])
    assert np.abs(zp).max() < 4.0, (case, float(np.abs(zp).max()), int((zp != 0).sum()))
    print(f"{case}: {n_tests} per-day proportions max|z| {np.abs(z).max():.2f}; {int((zp != 0).sum())} pooled proportions "
          f"max|z| {np.abs(zp).max():.2f}")
