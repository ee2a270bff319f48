"""Edge cases of the device path: empty inputs, cells without fish, one-fish cells, one huge cell, many tiny cells."""
import numpy as np
import pytest

from oracle import fixtures
import gpu_util as G
from stryke_b200 import _abi, engine

pytestmark = pytest.mark.gpu


def run(plan, variant=0, cells=slice(None), seed=9, cap=2_000_000_000):
    return engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day[cells], plan.cell_species[cells],
                            plan.cell_id[cells], precision=32, seed=seed, max_daily_fish=cap, kernel_variant=variant)


def test_no_cells_is_not_an_error():
    sim, plan = G.make_plan(fixtures.c2_facility(n_fish=10, iterations=1, days=1), "empty")
    res = run(plan, cells=slice(0, 0))
    assert res.cell_n.shape == (0,) and res.daily.shape == (0, 5) and res.state_daily.shape[0] == 0


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_species_never_present_gives_zero_rows(variant):
    prj = fixtures.c3_population(n_species=3, iterations=20, days=5)
    for row in prj["population"]:
        row["occur_prob"] = 0.0
    sim, plan = G.make_plan(prj, "absent")
    res = run(plan, variant)
    assert (res.cell_n == 0).all() and not res.daily.any() and not res.state_daily.any()


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_one_fish_cells(variant):
    sim, plan = G.make_plan(fixtures.c2_facility(n_fish=1, iterations=300, days=3), "one")
    res = run(plan, variant)
    assert (res.cell_n == 1).all()
    sd = res.state_daily.astype(np.int64)
    assert (sd[:, 0, 1] == 1).all() and (sd[..., 1].sum(axis=1) <= plan.fac.n_moves).all() and (sd[..., 0] <= sd[..., 1]).all()
    assert np.array_equal(res.daily, run(plan, 1).daily)


def test_one_cell_of_three_hundred_million_fish():
    """A single cell far beyond any tile / chunk / 16-bit boundary: counters, fish indices and the tile walk in 64 bits."""
    n = 300_000_000
    sim, plan = G.make_plan(fixtures.c2_facility(n_fish=n, iterations=1, days=1), "huge")
    res = run(plan, 3)
    sd = res.state_daily.astype(np.int64)[0]
    fac = plan.fac
    assert res.cell_n[0] == n and sd[0, 1] == n
    for k in range(1, fac.n_moves):
        now, prev = slice(fac.level_off[k], fac.level_off[k + 1]), slice(fac.level_off[k - 1], fac.level_off[k])
        assert sd[now, 1].sum() == sd[prev, 0].sum()
    again = run(plan, 2)
    assert np.array_equal(again.state_daily, res.state_daily) and np.array_equal(again.daily, res.daily)


def test_many_tiny_cells_population_mode_all_variants_agree():
    prj = fixtures.c3_population(n_species=6, iterations=400, days=12)
    sim, plan = G.make_plan(prj, "tiny")
    a, b, c = (run(plan, v, cap=10 ** 7) for v in (1, 2, 3))
    assert a.cell_n.sum() > 0 and (a.cell_n > 0).mean() < 0.9
    for other in (b, c):
        assert np.array_equal(a.cell_n, other.cell_n) and np.array_equal(a.daily, other.daily)
        assert np.array_equal(a.state_daily, other.state_daily)


def _same(a, b):
    return (np.array_equal(a.cell_n, b.cell_n) and np.array_equal(a.daily, b.daily)
            and np.array_equal(a.state_daily, b.state_daily))


@pytest.mark.parametrize("order", ["engine", "shuffled", "ragged"])
def test_pooled_remainders_equal_one_tile_per_remainder(order, monkeypatch):
    """The specialised kernel packs the n mod 32 left-over fish of the cells of a 32-cell block into shared tiles
    (pool_block, passage_fast.cuh).  Same Philox counters (fish index, cell id) -> the tallies must not change, whether
    the blocks are uniform (engine order), mixed in species and day (shuffled cells: one tile per remainder) or the cell
    list is not a multiple of 32 long."""
    prj = fixtures.c3_population(n_species=5, iterations=90, days=7)
    sim, plan = G.make_plan(prj, "pool")
    cells = np.arange(plan.n_cells)
    if order == "shuffled":
        cells = np.random.default_rng(5).permutation(plan.n_cells)
    elif order == "ragged":
        cells = cells[3:-18]
    assert order != "ragged" or len(cells) % 32 != 0
    ref = run(plan, 1, cells=cells, cap=10 ** 7)
    monkeypatch.setenv("STRYKE_B200_POOL", "1")
    pooled = run(plan, 3, cells=cells, cap=10 ** 7)
    monkeypatch.setenv("STRYKE_B200_POOL", "0")
    plain = run(plan, 3, cells=cells, cap=10 ** 7)
    assert ref.cell_n.sum() > 0 and ((ref.cell_n & 31) != 0).any() and (ref.cell_n >= 32).any()
    assert _same(ref, pooled) and _same(ref, plain)


@pytest.mark.parametrize("n_fish", [1, 31, 32, 33, 63, 64, 95, 1000, 4097])
def test_pooled_remainders_fixed_counts_around_the_tile_size(n_fish, monkeypatch):
    """Fixed counts at and around multiples of the tile: no remainder at all, only remainders, both."""
    sim, plan = G.make_plan(fixtures.c2_facility(n_fish=n_fish, iterations=45, days=3), "poolfix")
    ref = run(plan, 1)
    monkeypatch.setenv("STRYKE_B200_POOL", "1")
    pooled = run(plan, 3)
    assert (ref.cell_n == n_fish).all() and _same(ref, pooled)


def test_pooled_remainders_next_to_a_huge_cell(monkeypatch):
    """One block holds a cell of 70 million fish (its full tiles span many chunks) between cells of a few fish."""
    prj = fixtures.c3_population(n_species=2, iterations=40, days=4)
    sim, plan = G.make_plan(prj, "poolhuge")
    table = np.array(plan.species_table, dtype=np.float64).reshape(-1, plan.species_table.shape[-1])
    big = table[0].copy()
    big[12], big[17], big[18] = 0.0, 1.0, 70_000_013.0      # STRYKE_ENT_FIXED, always present, n
    table = np.vstack([table, big[None, :]])
    species = plan.cell_species.copy()
    species[37] = len(table) - 1
    args = (plan.fac, plan.days, table, plan.cell_day, species, plan.cell_id)
    monkeypatch.setenv("STRYKE_B200_POOL", "1")
    a = engine.run_cells(*args, precision=32, seed=4, max_daily_fish=10 ** 8, kernel_variant=3)
    monkeypatch.setenv("STRYKE_B200_POOL", "0")
    b = engine.run_cells(*args, precision=32, seed=4, max_daily_fish=10 ** 8, kernel_variant=3)
    assert a.cell_n[37] == 70_000_013 and _same(a, b)
    # the population cap holds for fixed counts too (stryke.py:3052), as for drawn ones (:2610)
    with pytest.raises(_abi.StrykeError, match="exceeds STRYKE_MAX_DAILY_FISH"):
        engine.run_cells(*args, precision=32, seed=4, max_daily_fish=10 ** 7, kernel_variant=3)
    assert int(a.state_daily[37, 0, 1]) == 70_000_013


def test_kernel_timing_events_cover_every_launch():
    """stryke_b200_plan_time_kernel / plan_kernel_ms: the measurement hook bench.py's roofline uses."""
    import torch
    sim, plan_h = G.make_plan(fixtures.c2_facility(n_fish=200_000, iterations=4, days=2), "ktime")
    plan = engine.Plan(plan_h.fac, plan_h.days, plan_h.species_table, plan_h.cell_day, plan_h.cell_species, plan_h.cell_id,
                       device=0, precision=32, seed=3, max_daily_fish=10 ** 7)
    n, d, sd = plan.allocate()
    plan.launch(n, d, sd)                      # not timed
    assert plan.kernel_ms() == (0.0, 0)
    plan.time_kernel(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        plan.launch(n, d, sd)
    e1.record()
    ms, launches = plan.kernel_ms()
    torch.cuda.synchronize()
    assert launches == 3 and 0.0 < ms <= e0.elapsed_time(e1)
    assert plan.kernel_ms() == (0.0, 0)        # forgotten after the read
    plan.time_kernel(False)
    plan.launch(n, d, sd)
    assert plan.kernel_ms()[1] == 0
    assert plan.status() == 200_000 * 8


@pytest.mark.parametrize("fixture,kw", [("c5_cascade", dict(n_fish=1037, iterations=40, days=2)),            # 41 pairs: two pair rows per lane
                                        ("c5_cascade", dict(n_fish=333, iterations=50, days=2, dams=7)),
                                        ("c4_barotrauma", dict(n_fish=777, iterations=50, days=2)),          # barotrauma words
                                        ("cabot_station", dict(iterations=40, days=3, seed=2, fish=555.0)),  # 13 nodes in one level
                                        ("c1_single_unit", dict(iterations=60, days=20, fish=45.0))])
def test_pooled_remainders_on_other_facilities(fixture, kw, monkeypatch):
    """Pooling with more than 32 (move, node) pairs (several counters per lane in the accumulator flush), multi-word
    levels and barotrauma levels: identical to the general kernel."""
    sim, plan = G.make_plan(getattr(fixtures, fixture)(**kw), "poolfac")
    ref = run(plan, 1)
    monkeypatch.setenv("STRYKE_B200_POOL", "1")
    monkeypatch.setenv("STRYKE_B200_JIT", "1")
    pooled = run(plan, 3)
    assert ref.cell_n.sum() > 0 and _same(ref, pooled)
    if fixture == "c5_cascade" and "dams" not in kw:
        p = engine.Plan(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id, device=0,
                        precision=32, seed=9, max_daily_fish=2_000_000_000, kernel_variant=3)
        assert "pooled" in p.kernel() and plan.fac.n_pairs > 32
