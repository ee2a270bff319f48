"""Property-based checks (hypothesis) of the host-side building blocks that sit either side of the device path: the
spreadsheet reader/writer, the routing cdf, the cell sharding and the frame store."""
import os

import numpy as np
import pandas as pd
from hypothesis import given, settings, strategies as st

from stryke_b200 import routing, store, xlsx
from stryke_b200.distributed import shard_bounds, shard_by_weight

finite = st.floats(min_value=-1e9, max_value=1e9, allow_nan=False, allow_infinity=False, width=64)
names = st.text(alphabet=st.characters(whitelist_categories=("Lu", "Ll", "Nd"), whitelist_characters=" _-"), min_size=1, max_size=12)


@settings(max_examples=25, deadline=None)
@given(rows=st.lists(st.tuples(names, finite, st.integers(-10**6, 10**6), st.one_of(st.none(), finite)), min_size=1, max_size=12),
       startrow=st.integers(0, 6), startcol=st.integers(0, 3))
def test_workbook_roundtrip_keeps_values_and_geometry(tmp_path_factory, rows, startrow, startcol):
    path = str(tmp_path_factory.mktemp("wb") / "book.xlsx")
    df = pd.DataFrame(rows, columns=["name", "x", "k", "maybe"])
    df["name"] = df["name"].str.strip().replace("", "n")
    df["maybe"] = df["maybe"].astype(float)
    xlsx.new_workbook(path)
    xlsx.append_sheets(path, {"Sheet A": df}, startrow=startrow, startcol=startcol, index=False)
    back = xlsx.Workbook(path).frame("Sheet A", f"{xlsx.col_letters(startcol)}:{xlsx.col_letters(startcol + 3)}", skiprows=startrow)
    assert list(back.columns) == ["name", "x", "k", "maybe"] and len(back) == len(df)
    assert [str(v) for v in back["name"]] == [str(v) for v in df["name"]]
    assert np.allclose(back["x"].astype(float), df["x"], rtol=0, atol=0)
    assert list(back["k"].astype(int)) == list(df["k"])
    assert np.allclose(back["maybe"].astype(float), df["maybe"], equal_nan=True, rtol=0, atol=0)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.floats(min_value=0.0, max_value=1e6, allow_nan=False), min_size=1, max_size=16))
def test_routing_cdf_is_what_numpy_choice_searches(p):
    p = np.array(p)
    if p.sum() <= 0:
        return
    p = p / p.sum()
    cdf = routing.normalised_cdf(p)
    want = np.cumsum(p)
    want = want / want[-1]
    assert np.array_equal(cdf, want) and cdf[-1] == 1.0 and (np.diff(cdf) >= 0).all()
    u = np.random.default_rng(len(p)).random(64)
    assert np.array_equal(np.searchsorted(cdf, u, side="right"), [int((cdf <= x).sum()) for x in u])


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10**7), st.integers(1, 64))
def test_shards_partition_the_cells(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=50, deadline=None)
@given(st.lists(st.integers(0, 10**5), min_size=1, max_size=200), st.integers(1, 8))
def test_weighted_shards_are_contiguous_and_complete(weights, world):
    bounds = shard_by_weight(np.array(weights, dtype=np.int64), world)
    assert bounds[0] == 0 and bounds[-1] == len(weights) and len(bounds) == world + 1
    assert all(a <= b for a, b in zip(bounds, bounds[1:]))


@settings(max_examples=20, deadline=None)
@given(rows=st.lists(st.tuples(names, finite, st.integers(0, 1000)), min_size=1, max_size=20), parts=st.integers(1, 3))
def test_frame_store_append_concatenates(tmp_path_factory, rows, parts):
    path = str(tmp_path_factory.mktemp("st") / "proj.h5")
    df = pd.DataFrame(rows, columns=["species", "flow", "n"])
    with store.FrameStore(path, mode="w") as s:
        for _ in range(parts):
            s.append("Daily", df)
    with store.FrameStore(path, mode="r") as s:
        back = s["Daily"]
    assert len(back) == parts * len(df) and back["n"].sum() == parts * df["n"].sum()
    assert os.path.isdir(path + ".store")
