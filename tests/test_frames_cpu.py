"""Host logic between the device tallies and the reference's result frames, on the CPU: the compact-row decoder
(include/stryke_b200.h ``StrykeCompact``) against an encoder written from the header's description, the implicit cell grid
and its iteration slices, and ``runner.assemble_frames`` fed with the reference's own tallies (tests/golden) — dense,
compact, and split over several ranks — which must give the reference's Daily / State_Daily rows in the reference's order."""
import contextlib
import io

import numpy as np
import pytest

from golden_util import Golden
import gpu_util as G
from stryke_b200 import _abi, engine, runner
from stryke_b200.engine import CellResult, CompactResult

DAILY_COLS = ["pop_size", "num_entrained", "num_survived", "num_mortality", "mortality_impingement",
              "mortality_blade_strike", "mortality_barotrauma"]


def encode_compact(res: CellResult, narrow_max, first_slot=0, gap_every=None):
    """Dense tallies -> compact rows exactly as the header describes them."""
    C, P = res.cell_n.shape[0], res.state_daily.shape[1]
    W = P + 3
    NB = (C + 31) // 32
    blk = np.zeros((3, NB), np.uint32)
    rows = [np.full(W, 0xDEADBEEF, np.uint32) for _ in range(first_slot)]      # the buffer is indexed by absolute slot
    slot = first_slot
    for b in range(NB):
        if gap_every and b and b % gap_every == 0:
            for _ in range(5):                       # a new range of cells may start at any slot
                rows.append(np.full(W, 0xDEADBEEF, np.uint32)); slot += 1
        blk[0, b] = slot
        for j in range(32):
            c = 32 * b + j
            if c >= C or res.cell_n[c] <= 0:
                continue
            n, d, s = int(res.cell_n[c]), res.daily[c].astype(np.int64), res.state_daily[c].astype(np.int64)
            blk[1, b] |= np.uint32(1 << j)
            if n > narrow_max:
                blk[2, b] |= np.uint32(1 << j)
                w = np.zeros(2 * W, np.uint32)
                w[0], w[1:6] = n, d
                w[6:6 + 2 * P] = s.reshape(-1)
                rows += [w[:W], w[W:]]
                slot += 2
            else:
                w = np.zeros(W, np.uint32)
                w[0] = n | (d[4] << 16)
                w[1] = d[0] | (d[1] << 16)
                w[2] = d[2] | (d[3] << 16)
                w[3:3 + P] = s[:, 0] | (s[:, 1] << 16)
                rows.append(w)
                slot += 1
    rows = np.stack(rows) if rows else np.zeros((0, W), np.uint32)
    return CompactResult(C, P, W, narrow_max, len(rows), int(res.cell_n.sum()), blk[0], blk[1], blk[2], rows)


def golden_dense(g, plan):
    ref_idx = G.reference_cell_index(plan)
    P = plan.fac.n_pairs
    pair_of = {(int(plan.fac.pair_move[p]), plan.fac.node_names[int(plan.fac.pair_node[p])]): p for p in range(P)}
    n = np.zeros(plan.n_cells, np.int32)
    d = np.zeros((plan.n_cells, 5), np.uint32)
    s = np.zeros((plan.n_cells, P, 2), np.uint32)
    raw = {}
    for c, r in enumerate(ref_idx):
        cm = g.cells[r]
        n[c] = cm["n"]
        if cm["n"] == 0:
            continue
        d7 = g.ref(r, "daily")
        raw[c] = d7
        d[c] = [d7[1], d7[2], d7[4], d7[5], d7[6]]
        for (mv, nd), (su, cn) in g.ref_state_daily(r).items():
            s[c, pair_of[(mv, nd)]] = (su, cn)
    return CellResult(n, d, s, None), ref_idx


def check_frames(g, plan, sim, tallies):
    frames = runner.assemble_frames(sim, plan, tallies)
    daily, sd = frames["Daily"], frames["State_Daily"]
    assert len(daily) == len(g.cells)
    ref_daily = np.stack([g.ref(ci, "daily") for ci in range(len(g.cells))])
    assert np.array_equal(daily[DAILY_COLS].to_numpy(dtype=np.int64), ref_daily)
    assert [x.strip() for x in daily["species"]] == [cm["species"] for cm in g.cells]
    assert [x.strip() for x in daily["scenario"]] == [cm["scenario"] for cm in g.cells]
    assert list(daily["iteration"]) == [cm["iteration"] for cm in g.cells]
    assert [str(x) for x in daily["day"]] == [cm["day"] for cm in g.cells]
    want = []
    for ci, cm in enumerate(g.cells):
        if cm["n"] == 0:
            continue
        for m, s_, v in zip(g.ref(ci, "sd_move"), g.ref(ci, "sd_state"), g.ref(ci, "sd_vals")):
            want.append((cm["scenario"], cm["species"], cm["iteration"], cm["day"], int(m), g.model.nodes[int(s_)], float(v[0]), float(v[1])))
    got = [(r.scenario, r.species, int(r.iteration), str(r.day), int(r.move), r.state, float(r.successes), float(r["count"]))
           for _, r in sd.iterrows()] if sd is not None else []
    assert got == want


@pytest.mark.parametrize("case", ["c2_kaplan_francis", "c3_population", "c3_multi_scenario", "c5_cascade", "cabot_station"])
@pytest.mark.parametrize("layout", ["dense", "compact", "compact_wide", "two_ranks", "three_ranks"])
def test_frames_from_reference_tallies(case, layout):
    g = Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    res, ref_idx = golden_dense(g, plan)
    if layout == "dense":
        pieces = [(plan.grid, res)]
    elif layout == "compact":
        c = encode_compact(res, narrow_max=65534, first_slot=7, gap_every=3)
        back = c.dense()
        assert np.array_equal(back.cell_n, res.cell_n) and np.array_equal(back.daily, res.daily)
        assert np.array_equal(back.state_daily, res.state_daily)
        pieces = [(plan.grid, c)]
    elif layout == "compact_wide":
        pieces = [(plan.grid, encode_compact(res, narrow_max=3))]          # nearly every cell takes the two-slot 32-bit layout
        back = pieces[0][1].dense()
        assert np.array_equal(back.state_daily, res.state_daily) and np.array_equal(back.daily, res.daily)
    else:
        world = 2 if layout == "two_ranks" else 3
        day, spc, ids = plan.grid.explicit()
        pieces = []
        for r in range(world):
            sub = plan.grid.iterations_slice(r / world, (r + 1) / world)
            d2, s2, i2 = sub.explicit()
            pos = {int(i): k for k, i in enumerate(ids)}         # Philox ids identify cells across partitions
            sel = np.array([pos[int(i)] for i in i2], dtype=np.int64)
            assert np.array_equal(day[sel], d2) and np.array_equal(spc[sel], s2)
            part = CellResult(res.cell_n[sel], res.daily[sel], res.state_daily[sel], None)
            pieces.append((sub, encode_compact(part, narrow_max=65534) if r % 2 == 0 else part))
        assert sum(p[0].n_cells for p in pieces) == plan.n_cells
    tallies = runner.group_tallies(plan, pieces)
    check_frames(g, plan, sim, tallies)
    back = runner.dense_from_tallies(plan, tallies)
    assert np.array_equal(back.cell_n, res.cell_n) and np.array_equal(back.state_daily, res.state_daily)


def test_iteration_slices_partition_the_grid():
    from stryke_b200 import workloads
    prj = workloads.c3_population(n_species=5, iterations=11, days=7, scenarios=["A", "B"], facility_scenario_column=False)
    sim, plan = workloads.make_plan(prj, "slices")
    full = set(int(x) for x in plan.grid.explicit()[2])
    assert len(full) == plan.n_cells == 5 * 2 * 11 * 7
    for world in (1, 2, 3, 8, 16):
        seen = []
        for r in range(world):
            sub = plan.grid.iterations_slice(r / world, (r + 1) / world)
            seen += [int(x) for x in sub.explicit()[2]]
            assert sub.n_cells == len(sub.explicit()[0])
        assert sorted(seen) == sorted(full)


@pytest.mark.parametrize("case", ["c3_multi_scenario", "cabot_station"])
def test_result_file_in_the_reference_layout(case, tmp_path):
    """The frames run() stores also reach ``<name>.h5`` in the pandas table layout (``/<key>/table`` with ``index`` and
    ``values_block_N`` members, stryke.py:3344 / :3379): written by the store on close, read back here with the tests' own
    HDF5 reader (and by ``pandas.HDFStore`` where PyTables exists: tests/test_h5lite_cpu.py)."""
    import h5parse
    from stryke_b200 import store
    g = Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    res, _ = golden_dense(g, plan)
    frames = runner.assemble_frames(sim, plan, runner.group_tallies(plan, [(plan.grid, res)]))
    path = str(tmp_path / "result.h5")
    with store.FrameStore(path, mode="w") as st:
        for key, df in frames.items():
            if df is not None and len(df):
                st.append(key, df)
    store.wait_for_exports(path)                      # (the file is written on a worker thread after close)
    f = h5parse.File(path)
    tables = {k: o for k, o in f.walk() if o["kind"] == "dataset"}
    assert {"/Daily/table", "/State_Daily/table"} <= set(tables)
    for key in ("Daily", "State_Daily"):
        df = frames[key]
        grp, tab = f.get(key), tables[f"/{key}/table"]
        assert grp["attrs"]["pandas_type"] == "frame_table" and grp["attrs"]["non_index_axes"] == [(1, [str(c) for c in df.columns])]
        rec = f.read(tab)
        assert len(rec) == len(df) and rec.dtype.names[0] == "index"
        for name in grp["attrs"]["values_cols"]:
            for j, c in enumerate(tab["attrs"][f"{name}_kind"]):
                col, kind = df[c], tab["attrs"][f"{name}_dtype"]
                v = rec[name][:, j]
                if kind.startswith("bytes"):
                    assert [x.decode("utf-8") for x in v] == [str(x) for x in col]
                elif kind == "datetime64[ns]":
                    assert np.array_equal(v.view("datetime64[ns]"), col.to_numpy().astype("datetime64[ns]"))
                else:
                    assert np.array_equal(v, col.to_numpy())
    sd = tables["/State_Daily/table"]["dtype"]
    assert max(sd[n].subdtype[0].itemsize for n in sd.names if sd[n].subdtype and sd[n].subdtype[0].kind == "S") >= 256      # min_itemsize of the reference
