"""Compact tally rows (``stryke_b200_run_compact`` / ``plan_launch_compact``) and implicit cell lists (``StrykeCellGroup``):
the same tallies as the dense layout, bit for bit, for every passage kernel, for narrow (16-bit) and wide (32-bit) cells,
for explicit and implicit cell lists, for one launch and for ranges / slabs of cells."""
import numpy as np
import pytest

from oracle import fixtures
import gpu_util as G
from stryke_b200 import _abi, engine

pytestmark = pytest.mark.gpu


def dense(plan, seed=5, variant=0, cap=10 ** 9, cells=None):
    args = (plan.grid, None, None) if cells is None else cells
    return engine.run_cells(plan.fac, plan.days, plan.species_table, *args, precision=32, seed=seed, max_daily_fish=cap,
                            kernel_variant=variant)


def compact(plan, seed=5, variant=0, cap=10 ** 9, cells=None, **kw):
    args = (plan.grid, None, None) if cells is None else cells
    return engine.run_compact(plan.fac, plan.days, plan.species_table, *args, precision=32, seed=seed, max_daily_fish=cap,
                              kernel_variant=variant, **kw)


def same(a, b):
    assert np.array_equal(a.cell_n, b.cell_n)
    assert np.array_equal(a.daily, b.daily)
    assert np.array_equal(a.state_daily, b.state_daily)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_population_mode_compact_equals_dense(variant):
    prj = fixtures.c3_population(n_species=4, iterations=300, days=6, extreme=True)
    sim, plan = G.make_plan(prj, "cmp")
    d = dense(plan, variant=variant)
    c = compact(plan, variant=variant)
    assert c.n_slots == int((d.cell_n > 0).sum()) + int((d.cell_n > c.narrow_max).sum())
    assert c.n_fish == int(d.cell_n.astype(np.int64).sum())
    same(c.dense(), d)
    # rows hold nothing for absent cells, and the block tables say where every present cell's row is
    cells, slot, wide = c.slots()
    assert np.array_equal(cells, np.flatnonzero(d.cell_n > 0))
    assert len(np.unique(slot)) == len(slot)


def test_implicit_cell_list_equals_explicit_arrays():
    prj = fixtures.c3_population(n_species=3, iterations=130, days=5, scenarios=["A", "B"], facility_scenario_column=False)
    sim, plan = G.make_plan(prj, "grid")
    a = dense(plan)
    b = dense(plan, cells=(plan.cell_day, plan.cell_species, plan.cell_id))
    same(a, b)
    same(compact(plan).dense(), compact(plan, cells=(plan.cell_day, plan.cell_species, plan.cell_id)).dense())
    # general kernel too (fp64, native stream)
    g1 = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.grid, None, None, precision=64, seed=3, max_daily_fish=10 ** 9)
    g2 = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id, precision=64,
                          seed=3, max_daily_fish=10 ** 9)
    same(g1, g2)


@pytest.mark.parametrize("n_fish", [31, 1000, 65534, 65535, 70000, 300000])
def test_narrow_and_wide_cells(n_fish):
    """One turbine level on C2: cells up to 65 534 fish keep 16-bit counters, larger ones take two slots of 32-bit ones."""
    prj = fixtures.c2_facility(n_fish=n_fish, iterations=40, days=2)
    sim, plan = G.make_plan(prj, "wide")
    for variant in (1, 3):
        d = dense(plan, variant=variant)
        c = compact(plan, variant=variant)
        assert c.narrow_max == 65534
        assert c.n_slots == plan.n_cells * (2 if n_fish > 65534 else 1)
        same(c.dense(), d)


def test_cascade_rows_and_mixed_widths():
    """41 pairs (two pair rows per lane), five turbine levels (narrow_max 13 106), cells on both sides of it."""
    prj = fixtures.c5_cascade(n_fish=1, iterations=64, days=2)
    sim, plan = G.make_plan(prj, "casc")
    sp = plan.species_table.copy()
    sp = np.concatenate([sp, sp, sp])
    sp[0, 18], sp[1, 18], sp[2, 18] = 900, 13106, 13107
    spc = (np.arange(plan.n_cells) % 3).astype(np.int32)
    cells = (plan.cell_day, spc, plan.cell_id)
    for variant in (1, 2, 3):
        d = engine.run_cells(plan.fac, plan.days, sp, *cells, precision=32, seed=9, max_daily_fish=10 ** 9, kernel_variant=variant)
        c = engine.run_compact(plan.fac, plan.days, sp, *cells, precision=32, seed=9, max_daily_fish=10 ** 9, kernel_variant=variant)
        assert c.narrow_max == 13106
        assert c.n_slots == plan.n_cells + int((spc == 2).sum())
        same(c.dense(), d)


@pytest.mark.parametrize("n_slabs", [1, 3, 7])
def test_slabs_give_the_same_rows(n_slabs):
    prj = fixtures.c3_population(n_species=4, iterations=700, days=5)
    sim, plan = G.make_plan(prj, "slab")
    ref = compact(plan, n_slabs=1)
    got = compact(plan, n_slabs=n_slabs)
    assert got.n_slots == ref.n_slots and got.n_fish == ref.n_fish
    assert np.array_equal(got.blk_row, ref.blk_row) and np.array_equal(got.blk_mask, ref.blk_mask)
    assert np.array_equal(got.rows, ref.rows)


def test_rows_buffer_too_small_is_an_error():
    prj = fixtures.c3_population(n_species=4, iterations=100, days=4)
    sim, plan = G.make_plan(prj, "cap")
    ok = compact(plan)
    with pytest.raises(_abi.StrykeError, match="rows buffer too small"):
        compact(plan, rows_cap=ok.n_slots - 1)
    again = compact(plan, rows_cap=ok.n_slots)
    assert np.array_equal(again.rows, ok.rows)


def test_plan_ranges_on_device_buffers():
    """plan_launch_compact over ranges of cells (what the multi-GPU path and bench.py do): slots continue from one range to
    the next, the totals and the rows equal the one-shot call's."""
    import torch
    prj = fixtures.c3_population(n_species=4, iterations=500, days=5)
    sim, plan = G.make_plan(prj, "rng")
    ref = compact(plan, n_slabs=1, seed=21)
    p = engine.Plan(plan.fac, plan.days, plan.species_table, plan.grid, None, None, device=0, precision=32, seed=21,
                    max_daily_fish=10 ** 9)
    C = plan.n_cells
    cell_n = torch.zeros(C, dtype=torch.int32, device="cuda")
    rows = torch.full((2 * C, plan.fac.n_pairs + 3), -1, dtype=torch.int32, device="cuda")
    after = torch.zeros(3, dtype=torch.int64, device="cuda")
    cuts = [0, 1024 * 1, 1024 * 5, C]
    for k in range(3):
        p.launch_compact(cell_n, rows, cuts[k], cuts[k + 1], first_slot=0 if k == 0 else -1, slots_after=after[k:k + 1])
    ns, nf = p.compact_totals()
    p.status()
    assert (ns, nf) == (ref.n_slots, ref.n_fish)
    assert int(after[2]) == ns and int(after[0]) <= int(after[1]) <= ns
    br, bm, bw, row_w, narrow_max = p.compact_tables()
    NB = (C + 31) // 32
    assert np.array_equal(br.cpu().numpy().view(np.uint32)[:NB], ref.blk_row)
    assert np.array_equal(bm.cpu().numpy().view(np.uint32)[:NB], ref.blk_mask)
    assert np.array_equal(rows[:ns].cpu().numpy().view(np.uint32), ref.rows)
    assert (rows[ns:] == -1).all()            # nothing is written past the used slots
    # a second pass with another seed reuses everything
    p.set_seed(22)
    p.launch_compact(cell_n, rows, 0, -1, first_slot=0)
    ns2, nf2 = p.compact_totals()
    ref2 = compact(plan, n_slabs=1, seed=22)
    assert (ns2, nf2) == (ref2.n_slots, ref2.n_fish)
    assert np.array_equal(rows[:ns2].cpu().numpy().view(np.uint32), ref2.rows)


def test_cell_tables_out_of_range_are_reported():
    prj = fixtures.c2_facility(n_fish=100, iterations=4, days=2)
    sim, plan = G.make_plan(prj, "bad")
    day = plan.cell_day.copy()
    day[3] = 99
    with pytest.raises(_abi.StrykeError, match="out of range"):
        engine.run_cells(plan.fac, plan.days, plan.species_table, day, plan.cell_species, plan.cell_id, precision=32, seed=1)


@pytest.mark.parametrize("big_full,own,chunk", [("0", "1", ("8", "2")), ("1", "1", ("3", "1")), ("6", "1", None), ("255", "1", ("64", "8")),
                                                ("6", "0", ("16", "4"))])
def test_owned_rows_equal_the_dense_tallies(monkeypatch, big_full, own, chunk):
    """Pooled plans write the rows of their 'owned' cells (at most big_full full tiles) once, from the warp that simulates the
    block's atomic part, and add into zeroed rows only for big cells: whatever the threshold, the chunk sizes (which move the
    ownership boundaries) and the slabs, the rows must hold the dense tallies."""
    monkeypatch.setenv("STRYKE_B200_BIG_FULL", big_full)
    monkeypatch.setenv("STRYKE_B200_OWN", own)
    if chunk:
        monkeypatch.setenv("STRYKE_B200_CHUNK_TILES", chunk[0])
        monkeypatch.setenv("STRYKE_B200_CHUNK_MIN", chunk[1])
    prj = fixtures.c3_population(n_species=4, iterations=300, days=6, extreme=True)
    sim, plan = G.make_plan(prj, "own")
    want = dense(plan, variant=2)                     # generic tables, no pooling
    got = compact(plan, variant=3)
    same(got.dense(), want)
    same(dense(plan, variant=3), want)                # the same kernel writing the dense layout
    same(compact(plan, variant=3, n_slabs=3).dense(), want)
    # narrow cap: cells above it take the 32-bit layout and count as big
    same(compact(plan, variant=3, seed=11).dense(), dense(plan, variant=2, seed=11))
