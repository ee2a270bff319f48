"""CPU tests: the C-ABI library loads and exports every symbol include/stryke_b200.h declares (no compute without a
GPU — and a loud failure, not a fallback, when compute is asked for), the result store, summary() arithmetic (ported
from the reference's tests/test_summary_state_daily.py and tests/test_driver_diagnostics.py), the .xlsx reader, and
the world_size-2 sharding + all-reduce logic over gloo."""
import contextlib
import io
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from stryke_b200 import _abi
    header = open(os.path.join(ROOT, "include", "stryke_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(stryke_b200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no prototypes found in the header"
    lib = _abi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_abi.EXPORTS) == declared
    assert lib.stryke_b200_abi_version() == _abi.ABI_VERSION == 2


def test_abi_struct_layout_matches_header():
    """sizeof() of the ctypes mirrors vs the C compiler's."""
    from stryke_b200 import _abi
    import ctypes as C
    src = '#include <stdio.h>\n#include "stryke_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(StrykeFacility),sizeof(StrykeDayTables),sizeof(StrykeSpeciesTable),sizeof(StrykeCells)," \
          "sizeof(StrykeInjected),sizeof(StrykeTrace),sizeof(StrykeOpts),sizeof(StrykeTallies),sizeof(StrykeCellGroup)," \
          "sizeof(StrykeCompact));return 0;}"
    exe = "/tmp/stryke_b200_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src.encode(), check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(t) for t in (_abi.Facility, _abi.DayTables, _abi.SpeciesTable, _abi.Cells, _abi.Injected, _abi.Trace,
                                  _abi.Opts, _abi.Tallies, _abi.CellGroup, _abi.Compact)]
    assert got == want


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stryke_b200 import _abi, engine
    with pytest.raises(_abi.StrykeError, match="no CPU fallback") as exc:
        engine.strike_survival(1, np.ones(15), np.array([1.0]), np.array([0.5]))
    assert exc.value.code == _abi.E_NOGPU
    from oracle import fixtures
    from stryke_b200 import simulation
    sim = simulation.__new__(simulation)
    fixtures.apply_to(sim, fixtures.c2_facility(n_fish=10), proj_dir="/tmp/stryke_b200_nogpu")
    with contextlib.redirect_stdout(io.StringIO()), pytest.raises(_abi.StrykeError, match="no CPU fallback"):
        sim.run()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stryke_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "stryke_oracle" not in text, f


def test_frame_store_roundtrip(tmp_path):
    from stryke_b200.store import FrameStore
    p = str(tmp_path / "proj.h5")
    with FrameStore(p, "w") as st:
        st["Nodes"] = pd.DataFrame({"Location": ["a", "b"]})
        st.append("Daily", pd.DataFrame({"x": [1, 2]}))
        st.append("Daily", pd.DataFrame({"x": [3]}))
        assert list(st["Daily"]["x"]) == [1, 2, 3]
    with FrameStore(p, "a") as st:
        st.append("Daily", pd.DataFrame({"x": [4]}))
        st.put("Yearly_Summary", pd.DataFrame({"y": [1.5]}), format="table", data_columns=True)
    with FrameStore(p, "r") as st:
        assert set(st.keys()) == {"/Nodes", "/Daily", "/Yearly_Summary"}
        assert list(st["Daily"]["x"]) == [1, 2, 3, 4] and "Daily" in st and "/Nodes" in st and "Nope" not in st
        with pytest.raises(KeyError):
            st["Nope"]


def _summary_project(tmp_path, state_rows, daily_rows, unit_hours=None, route_flows=None):
    from stryke_b200 import simulation
    from stryke_b200.store import open_store
    sim = simulation.__new__(simulation)
    sim.proj_dir, sim.output_name, sim.wks_dir = str(tmp_path), "summ", None
    sim.moves = np.arange(0, 3)
    with open_store(os.path.join(sim.proj_dir, "summ.h5"), "w") as st:
        st["Population"] = pd.DataFrame([{"Species": "Alosa"}])
        st["Flow Scenarios"] = pd.DataFrame([{"Scenario": "Spring"}])
        st["Unit_Parameters"] = pd.DataFrame([{"Qopt": 100.0, "Qcap": 120.0, "H": 10.0}], index=pd.Index(["U1"], name="Unit"))
        st["Daily"] = pd.DataFrame(daily_rows)
        st["State_Daily"] = pd.DataFrame(state_rows)
        if unit_hours is not None:
            st["Unit_Hours"] = pd.DataFrame(unit_hours)
        if route_flows is not None:
            st["Route_Flows"] = pd.DataFrame(route_flows)
    return sim


def _daily(it, day, pop, ent, surv):
    return {"species": "{:50}".format("Alosa"), "scenario": "{:50}".format("Spring"), "season": "{:50}".format("Spring"),
            "iteration": it, "day": pd.Timestamp(day), "flow": 100.0, "pop_size": pop, "num_entrained": ent, "num_survived": surv,
            "num_mortality": ent - surv, "mortality_impingement": 0, "mortality_blade_strike": ent - surv, "mortality_barotrauma": 0}


def _state(it, day, move, state, succ, count):
    return {"scenario": "Spring", "species": "Alosa", "iteration": it, "day": pd.Timestamp(day), "move": move, "state": state,
            "successes": float(succ), "count": float(count)}


def test_summary_from_state_daily_and_driver_diagnostics(tmp_path):
    """Arithmetic pinned by the reference's tests/test_summary_state_daily.py and tests/test_driver_diagnostics.py:80-99
    (20 h, 20 % off Qopt, 100 % of hours outside the band)."""
    days = ["2023-03-01", "2023-03-02"]
    state, daily = [], []
    for it in range(2):
        for d in days:
            daily.append(_daily(it, d, 100, 60, 54))
            state += [_state(it, d, 0, "river_node_0", 100, 100), _state(it, d, 1, "U1", 54, 60), _state(it, d, 1, "spillway", 38, 40),
                      _state(it, d, 2, "river_node_1", 92, 92)]
    uh = [{"scenario": "Spring", "iteration": 0, "day": pd.Timestamp(d), "route": "U1", "hours": 10.0, "discharge_cfs": 120.0} for d in days]
    rf = [{"scenario": "Spring", "iteration": 0, "day": pd.Timestamp(d), "route": r, "discharge_cfs": q}
          for d in days for r, q in (("U1", 120.0), ("spillway", 80.0))]
    sim = _summary_project(tmp_path, state, daily, uh, rf)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.summary()
    b = sim.beta_df.set_index("state")
    assert b.at["whole", "survival rate"] == pytest.approx(0.92) and b.at["U1", "survival rate"] == pytest.approx(0.9)
    assert b.at["spillway", "survival rate"] == pytest.approx(0.95) and b.at["whole", "variance"] == 0.0
    assert list(sim.beta_df_units_only["Passage Route"]) == ["U1", "spillway"]
    y = sim.cum_sum.iloc[0]
    assert y["prob_entrainment"] == pytest.approx(0.6) and y["mean_yearly_entrainment"] == 120.0
    assert y["mean_yearly_mortality"] == 12.0 and y["1_in_10_day_entrainment"] == 60.0
    d = sim.driver_diagnostics.set_index("route").loc["U1"]
    assert d["total_hours"] == 20.0 and d["mean_abs_pct_off_qopt"] == pytest.approx(20.0)
    assert d["pct_hours_outside_qopt_band"] == pytest.approx(100.0) and d["flow_share"] == 1.0
    assert d["mortality_weight"] == pytest.approx(0.1) and d["mean_discharge_cfs"] == 120.0
    from stryke_b200.store import open_store
    with open_store(os.path.join(sim.proj_dir, "summ.h5"), "r") as st:
        assert {"/Daily_Summary", "/Beta_Distributions", "/Beta_Distributions_Units", "/Yearly_Summary", "/Driver_Diagnostics"} <= set(st.keys())


def test_bootstrap_matches_the_reference_stream():
    """One vectorised randint call consumes MT19937 exactly like the reference's loop of np.random.choice calls."""
    from stryke_b200.summary import bootstrap_mean_ci
    data = np.random.default_rng(0).random(37)
    np.random.seed(42)
    means = [np.mean(np.random.choice(data, size=len(data), replace=True)) for _ in range(300)]
    want = (np.mean(data), np.percentile(means, 2.5), np.percentile(means, 97.5))
    np.random.seed(42)
    got = bootstrap_mean_ci(data, n_bootstrap=300, ci=95, chunk_elems=37 * 7)
    assert got == pytest.approx(want, rel=0, abs=0)


def test_xlsx_reader_and_appender(tmp_path):
    """Round trip through a workbook written by this package's own appender (sheet geometry of SURVEY.md Appendix F:
    header row = skiprows + 1, usecols by letter)."""
    import zipfile
    from stryke_b200 import xlsx
    path = str(tmp_path / "book.xlsx")
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("[Content_Types].xml", '<?xml version="1.0"?><Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="xml" ContentType="application/xml"/><Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/></Types>')
        z.writestr("_rels/.rels", '<?xml version="1.0"?><Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument" Target="xl/workbook.xml"/></Relationships>')
        z.writestr("xl/workbook.xml", '<?xml version="1.0"?><workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets></sheets></workbook>')
        z.writestr("xl/_rels/workbook.xml.rels", '<?xml version="1.0"?><Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships"></Relationships>')
    df = pd.DataFrame({"Location": ["river_node_0", "U1"], "Surv_Fun": ["a priori", "Kaplan"], "Survival": [1.0, 0.5], "n": [3, 4]})
    xlsx.append_sheets(path, {"Nodes": df.set_index("Location")})
    wb = xlsx.Workbook(path)
    got = wb.frame("Nodes", "A:D", skiprows=0)
    assert list(got.columns) == ["Location", "Surv_Fun", "Survival", "n"]
    assert list(got["Location"]) == ["river_node_0", "U1"] and got["Survival"].tolist() == [1.0, 0.5]
    assert got["n"].dtype == np.int64 and got["Survival"].dtype == np.float64
    assert list(wb.frame("Nodes", "B:C", skiprows=0).columns) == ["Surv_Fun", "Survival"]
    with pytest.raises(ValueError, match="already exists"):
        xlsx.append_sheets(path, {"Nodes": df})
    assert xlsx.col_index("B") == 1 and xlsx.col_index("AA") == 26 and xlsx.col_letters(27) == "AB"


def test_shard_bounds_cover_everything():
    from stryke_b200.distributed import shard_bounds, shard_by_weight
    for n in (0, 1, 7, 1000, 36_500_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    cuts = shard_by_weight(np.array([1, 1, 1, 1, 100, 1, 1, 1.0]), 2)
    assert cuts[0] == 0 and cuts[-1] == 8 and 4 <= cuts[1] <= 5


GLOO_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    from golden_util import Golden
    import gpu_util as G
    from test_frames_cpu import encode_compact, golden_dense, check_frames
    from stryke_b200 import runner
    from stryke_b200.engine import CellResult
    from stryke_b200.distributed import ShardLayout, pack_flat, unpack_flat, gather_flat, agree_on_status
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the reference's own tallies of a multi-scenario population-mode run (tests/golden), split by iteration over the ranks:
    # every rank packs ITS iterations as compact rows into its flat buffer, ONE all-gather, every rank assembles the frames
    g = Golden("c3_multi_scenario")
    sim, plan = G.make_plan(g.prj, "gloo")
    res, _ = golden_dense(g, plan)
    day, spc, ids = plan.grid.explicit()
    pos = {int(i): k for k, i in enumerate(ids)}
    grids = [plan.grid.iterations_slice(r / world, (r + 1) / world) for r in range(world)]
    mine = grids[rank]
    sel = np.array([pos[int(i)] for i in mine.explicit()[2]], dtype=np.int64)
    part = encode_compact(CellResult(res.cell_n[sel], res.daily[sel], res.state_daily[sel], None), narrow_max=65534)
    n_max = max(x.n_cells for x in grids)
    L = ShardLayout(n_max, plan.fac.n_pairs, slab_cells=1024 * ((n_max + 1023) // 1024), slab_cap=2 * n_max + 8)
    flat = torch.from_numpy(pack_flat(L, np.stack([part.blk_row, part.blk_mask, part.blk_wide]), part.rows, 0, mine.n_cells, part.n_slots))
    gathered = gather_flat(flat, world).numpy().reshape(world, L.total)
    pieces = []
    for r in range(world):
        n_cells, status, c = unpack_flat(L, gathered[r], plan.fac.n_pairs)
        assert n_cells == grids[r].n_cells and status == 0
        pieces.append((grids[r], c))
    check_frames(g, plan, sim, runner.group_tallies(plan, pieces))
    # an error on one rank reaches all of them before any further collective
    code = agree_on_status(-4 if rank == 1 else 0)
    ok = code == -4 and agree_on_status(0) == 0
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 3)
""")


def test_world_size_2_sharding_and_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29653", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    assert [p.wait(timeout=240) for p in procs] == [0, 0]


C_CLIENT = r'''
#include <stdio.h>
#include <string.h>
#include "stryke_b200.h"
/* what a C (or cgo / JNI) caller of the boundary does: link the shared library, call through the header */
int main(void) {
  if (stryke_b200_abi_version() != STRYKE_B200_ABI_VERSION) return 2;
  int n_dev = stryke_b200_device_count();
  StrykeFacility fac; StrykeDayTables days; StrykeSpeciesTable sp; StrykeCells cells; StrykeOpts opts; StrykeTallies out;
  memset(&fac, 0, sizeof fac); memset(&days, 0, sizeof days); memset(&sp, 0, sizeof sp); memset(&cells, 0, sizeof cells);
  memset(&opts, 0, sizeof opts); memset(&out, 0, sizeof out);
  int rc = stryke_b200_run(&fac, &days, &sp, &cells, &opts, &out);     /* empty tables: must be refused, not crash */
  printf("%d %d %s\n", n_dev, rc, stryke_b200_last_error());
  return rc == 0 ? 3 : 0;
}
'''


def test_a_plain_c_program_links_against_the_boundary(tmp_path):
    """The boundary is a C ABI: the header compiles as C (gcc -std=c99 -Wall -Werror), a C program links against
    libstryke_b200.so and gets an error code + message (never a crash) for invalid input."""
    from stryke_b200 import build as B
    lib = B.build()
    exe = str(tmp_path / "c_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    "-L", os.path.dirname(lib), "-lstryke_b200", "-Wl,-rpath," + os.path.dirname(lib)],
                   input=C_CLIENT.encode(), check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    n_dev, rc, msg = r.stdout.strip().split(" ", 2)
    assert int(rc) < 0 and msg


def test_fallback_store_exports_to_parquet_and_csv(tmp_path):
    """The directory store that stands in for the .h5 where PyTables is missing can be handed on as parquet / csv
    (and as a real HDF5 file where PyTables exists): same keys, same frames."""
    import pandas as pd
    from stryke_b200 import store
    path = str(tmp_path / "proj.h5")
    daily = pd.DataFrame({"species": ["a", "b"], "iteration": [0, 1], "day": pd.to_datetime(["2023-03-01", "2023-03-02"]),
                          "pop_size": [3, 4]})
    with store.FrameStore(path, mode="w") as st:
        st["Population"] = pd.DataFrame({"Species": ["a"], "Fish": [3.0]})
        st.append("Daily", daily)
        st.append("simulations/Spring/a", daily)
    out = store.export(path, str(tmp_path / "pq"), "parquet")
    assert sorted(os.path.basename(f) for f in out) == ["Daily.parquet", "Population.parquet", "simulations__Spring__a.parquet"]
    assert pd.read_parquet(os.path.join(str(tmp_path / "pq"), "Daily.parquet")).equals(daily)
    out = store.export(path, str(tmp_path / "csv"), "csv")
    assert len(out) == 3 and pd.read_csv(out[0]).shape[0] == 2
    # closing the store also wrote <name>.h5 itself, and export(..., "x.h5") writes one too: with PyTables through
    # pandas.HDFStore, without it through stryke_b200.h5lite — read here with the tests' own HDF5 reader
    h5 = store.export(path, str(tmp_path / "x.h5"))
    store.wait_for_exports(path)
    assert h5 == [str(tmp_path / "x.h5")] and os.path.isfile(path)
    if not store.pytables_available():
        import h5parse
        for file in (path, h5[0]):
            f = h5parse.File(file)
            assert sorted(k for k, o in f.walk() if o["kind"] == "dataset") == ["/Daily/table", "/Population/table",
                                                                               "/simulations/Spring/a/table"]
            rec = f.read(f.get("Daily/table"))
            t = f.get("Daily/table")["attrs"]
            assert t["values_block_1_kind"] == ["iteration", "pop_size"] and np.array_equal(rec["values_block_1"], [[0, 3], [1, 4]])
            assert t["values_block_0_dtype"] == "datetime64[ns]"
            assert np.array_equal(rec["values_block_0"][:, 0].view("datetime64[ns]"), daily["day"].to_numpy())
            assert [x.decode() for x in rec["values_block_2"][:, 0]] == ["a", "b"]


def test_slab_plan_and_layout():
    """Slab bounds are multiples of the range alignment, cover the cells, decrease in size; every slab's region of the rows
    buffer holds the rows the slab is expected to write (checked against an actual draw of the presence rolls)."""
    from stryke_b200 import workloads as W
    from stryke_b200.distributed import ALIGN, SLAB_SHARES, ShardLayout, slab_plan
    prj = W.c3_population(n_species=6, iterations=1000, days=60, scenarios=["A", "B"], facility_scenario_column=False)
    sim, plan = W.make_plan(prj, "slabplan")
    world = 2
    grids = [plan.grid.iterations_slice(r / world, (r + 1) / world) for r in range(world)]
    n_max = max(g.n_cells for g in grids)
    assert n_max >= (1 << 18)
    cell_lo, slot_lo = slab_plan(grids, plan.species_table, n_max)
    assert cell_lo[0] == 0 and cell_lo[-1] == n_max and len(cell_lo) == len(SLAB_SHARES) + 1
    assert all(c % ALIGN == 0 for c in cell_lo[:-1])
    sizes = np.diff(cell_lo)
    assert all(sizes[i] >= sizes[i + 1] for i in range(len(sizes) - 1))
    L = ShardLayout(n_max, plan.fac.n_pairs, cell_lo=cell_lo, slot_lo=slot_lo)
    assert L.n_slabs == len(SLAB_SHARES) and L.slots == slot_lo[-1] and L.total % 4 == 0
    assert [L.slab_words(k)[1] - L.slab_words(k)[0] for k in range(L.n_slabs)] == [int(x) * L.W for x in np.diff(slot_lo)]
    rng = np.random.default_rng(3)
    for g in grids:
        day, spc, ids = g.explicit()
        occur = plan.species_table[spc, 17]
        present = rng.random(len(spc)) <= np.where(np.isfinite(occur), occur, 1.0)
        for k in range(L.n_slabs):
            assert int(present[cell_lo[k]:cell_lo[k + 1]].sum()) <= slot_lo[k + 1] - slot_lo[k]
    # small lists: one slab; the uniform constructor still works
    one = slab_plan(grids, plan.species_table, 5000)
    assert one[0] == [0, 5000] and len(one[1]) == 2
    U = ShardLayout(5000, plan.fac.n_pairs, slab_cells=2048, slab_cap=3000)
    assert U.cell_lo == [0, 2048, 4096, 5000] and U.slot_lo == [0, 3000, 6000, 9000]
