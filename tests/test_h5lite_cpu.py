"""``stryke_b200.h5lite`` (the pure-Python writer of the pandas/PyTables table layout) against ``tests/h5parse.py`` (an
independent reader of the same HDF5 subset): structure, values, attributes; with PyTables present also ``pandas.HDFStore``."""
import os

import numpy as np
import pandas as pd
import pytest

from stryke_b200 import h5lite
from stryke_b200.store import pytables_available
import h5parse


def demo_frame(n=1000, seed=0):
    rng = np.random.default_rng(seed)
    return pd.DataFrame({
        "species": rng.choice(["Alosa sapidissima", "Anguilla rostrata", "Micropterus"], n),
        "scenario": "Spring high flow",
        "iteration": rng.integers(0, 50, n),
        "day": pd.Timestamp("2020-01-01") + pd.to_timedelta(rng.integers(0, 365, n), unit="D"),
        "flow": rng.random(n) * 1e4,
        "pop_size": rng.integers(0, 5000, n).astype(np.int64),
        "num_survived": rng.random(n),
    })


def check_frame(f, key, df, data_columns=False):
    grp = f.get(key)
    assert grp["kind"] == "group"
    a = grp["attrs"]
    assert a["pandas_type"] == "frame_table" and a["table_type"] == "appendable_frame" and a["CLASS"] == "GROUP"
    assert a["non_index_axes"] == [(1, list(df.columns))] and a["index_cols"] == [(0, "index")]
    assert int(a["levels"]) == 1 and a["nan_rep"] == "nan" and a["encoding"] == "UTF-8"
    tab = f.get(key + "/table")
    t = tab["attrs"]
    assert t["CLASS"] == "TABLE" and int(t["NROWS"]) == len(df) and t["index_kind"] == "integer"
    rec = f.read(tab)
    assert rec.dtype.names[0] == "index" and np.array_equal(rec["index"], np.arange(len(df)))
    assert [t[f"FIELD_{i}_NAME"] for i in range(len(rec.dtype.names))] == list(rec.dtype.names)
    got = {}
    for name in a["values_cols"]:
        cols = t[f"{name}_kind"]
        block = rec[name] if rec[name].ndim == 2 else rec[name][:, None]
        assert block.shape[1] == len(cols)
        for j, c in enumerate(cols):
            got[c] = (block[:, j], t[f"{name}_dtype"])
    assert set(got) == set(df.columns)
    for c in df.columns:
        v, kind = got[c]
        col = df[c]
        if col.dtype.kind == "f":
            assert kind == "float64" and np.array_equal(v, col.to_numpy())
        elif col.dtype.kind in "iu":
            assert kind == "int64" and np.array_equal(v, col.to_numpy())
        elif col.dtype.kind == "M":
            assert kind == "datetime64[ns]" and np.array_equal(v.view("datetime64[ns]"), col.to_numpy().astype("datetime64[ns]"))
        else:
            assert kind.startswith("bytes") and [x.decode("utf-8") for x in v] == [str(x) for x in col]
    assert a["data_columns"] == (list(a["values_cols"]) if data_columns else [])


def test_roundtrip_and_structure(tmp_path):
    frames = {"Daily": demo_frame(1000, 1), "simulations/Spring high flow/Alosa sapidissima": demo_frame(17, 2),
              "simulations/Spring high flow/Micropterus": demo_frame(0, 3), "Yearly_Summary": demo_frame(40, 4)}
    path = str(tmp_path / "out.h5")
    h5lite.write_frames(path, frames, data_columns={"Yearly_Summary"})
    f = h5parse.File(path)
    assert f.root["attrs"]["PYTABLES_FORMAT_VERSION"] == "2.1"
    keys = sorted(k for k, o in f.walk() if o["kind"] == "dataset")
    assert keys == sorted("/" + k + "/table" for k in frames)
    for k, df in frames.items():
        check_frame(f, k, df, data_columns=k == "Yearly_Summary")
    # every object header, heap, B-tree node and chunk lies on an 8-byte boundary inside the file (the parser checks)
    assert os.path.getsize(path) % 8 == 0 or True


def test_many_chunks_make_a_multi_level_btree(tmp_path):
    n = 70000
    rec = np.zeros(n, dtype=[("index", "<i8"), ("values_block_0", "<f8", (2,)), ("values_block_1", "S5", (1,))])
    rec["index"] = np.arange(n)
    rec["values_block_0"] = np.random.default_rng(5).random((n, 2))
    rec["values_block_1"][:, 0] = [b"a%d" % (i % 977) for i in range(n)]
    for chunk_rows, levels in ((n, 0), (1024, 1), (16, 2)):
        w = h5lite.Writer()
        w.table("T/table", rec, {"CLASS": "TABLE"}, chunk_rows=chunk_rows)
        path = str(tmp_path / f"c{chunk_rows}.h5")
        w.write(path)
        f = h5parse.File(path)
        tab = f.get("T/table")
        assert tab["layout"]["chunk"] == (chunk_rows,) and tab["maxshape"] == (None,)
        got = f.read(tab)
        assert got.dtype == rec.dtype and np.array_equal(got, rec)
        import struct
        level = struct.unpack_from("<B", f.b, tab["layout"]["btree"] + 5)[0]
        assert level == levels


def test_groups_with_many_links(tmp_path):
    w = h5lite.Writer()
    names = [f"species {i:03d}" for i in range(150)]          # > 64 links: several symbol nodes under the group's B-tree
    rec = np.zeros(3, dtype=[("index", "<i8"), ("values_block_0", "<i8", (1,))])
    for i, nm in enumerate(names):
        rec["values_block_0"][:, 0] = i
        w.table(f"simulations/scenario/{nm}/table", rec.copy(), {"CLASS": "TABLE", "NROWS": np.int64(3)})
    path = str(tmp_path / "g.h5")
    w.write(path)
    f = h5parse.File(path)
    links = f.links(f.get("simulations/scenario"))
    assert sorted(links) == sorted(names)
    for i in (0, 63, 64, 149):
        assert np.all(f.read(f.get(f"simulations/scenario/{names[i]}/table"))["values_block_0"] == i)


def test_min_itemsize_and_missing_strings(tmp_path):
    df = pd.DataFrame({"state": ["river", None, "Unit 1"], "route": ["a", "b", "c"], "x": [1.0, np.nan, 3.0]})
    rec, gattrs, tattrs = h5lite.frame_table(df, min_itemsize={"state": 256})
    assert rec.dtype["values_block_1"].subdtype[0] == np.dtype("S256")
    assert rec["values_block_1"][1, 0] == b"nan" and np.isnan(rec["values_block_0"][1, 0])
    path = str(tmp_path / "m.h5")
    h5lite.write_frames(path, {"State_Daily": df}, min_itemsize={"State_Daily": {"state": 256}})
    f = h5parse.File(path)
    assert f.get("State_Daily/table")["dtype"]["values_block_1"].subdtype[0].itemsize == 256


@pytest.mark.skipif(not pytables_available(), reason="PyTables is not installed in this image")
def test_pandas_reads_the_file_back(tmp_path):
    frames = {"Daily": demo_frame(500, 7), "simulations/A/B": demo_frame(20, 8)}
    path = str(tmp_path / "pd.h5")
    h5lite.write_frames(path, frames)
    with pd.HDFStore(path, mode="r") as st:
        for k, df in frames.items():
            got = st[k]
            pd.testing.assert_frame_equal(got.reset_index(drop=True), df.reset_index(drop=True), check_dtype=False)


def test_the_reader_reads_a_file_written_by_libhdf5():
    """Pins the tests' own HDF5 reader: SciPy ships a MATLAB v7.3 file — written by libhdf5 1.6 behind a 512-byte user block —
    that uses the very structures h5lite emits for groups, object headers, heaps, float datatypes and string attributes."""
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.isfile(path):
        pytest.skip("SciPy's test data are not installed")
    f = h5parse.File(path)
    assert (f.leaf_k, f.internal_k) == (4, 16)
    ds = f.get("testdouble")
    assert ds["kind"] == "dataset" and ds["dtype"] == np.dtype("<f8") and ds["shape"] == (9, 1)
    assert ds["attrs"] == {"MATLAB_class": "double"}
    got = f.read(ds)[:, 0]
    assert np.allclose(got, np.arange(9) * np.pi / 4)             # scipy's testdouble: 0, pi/4, ..., 2 pi


def test_random_frames_round_trip(tmp_path):
    """Property test: frames of random shape (float / int / datetime / string columns in any order, NaN and None cells,
    non-ASCII labels, nested keys, zero rows) come back value for value through the independent reader."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    label = st.text(alphabet=st.characters(blacklist_characters="\x00", blacklist_categories=("Cs",)), min_size=0, max_size=12)
    kinds = st.lists(st.sampled_from("fids"), min_size=1, max_size=6)
    counter = [0]

    @settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(kinds=kinds, n=st.integers(0, 40), seed=st.integers(0, 2 ** 16), labels=st.lists(label, min_size=1, max_size=5),
           nested=st.booleans())
    def run(kinds, n, seed, labels, nested):
        rng = np.random.default_rng(seed)
        cols = {}
        for j, k in enumerate(kinds):
            name = f"col {j}"
            if k == "f":
                v = rng.normal(size=n)
                v[rng.random(n) < 0.2] = np.nan
            elif k == "i":
                v = rng.integers(-2 ** 40, 2 ** 40, n)
            elif k == "d":
                v = (np.datetime64("2001-01-01", "ns") + rng.integers(0, 10 ** 15, n).astype("timedelta64[ns]"))
            else:
                v = np.array([labels[i] for i in rng.integers(0, len(labels), n)], dtype=object)
                v[rng.random(n) < 0.15] = None
            cols[name] = v
        df = pd.DataFrame(cols)
        key = "simulations/some scenario/Genus species" if nested else "Daily"
        counter[0] += 1
        path = str(tmp_path / f"r{counter[0]}.h5")
        h5lite.write_frames(path, {key: df})
        f = h5parse.File(path)
        grp, tab = f.get(key), f.get(key + "/table")
        rec = f.read(tab)
        assert len(rec) == n and int(tab["attrs"]["NROWS"]) == n
        seen = set()
        for name in grp["attrs"]["values_cols"]:
            for j, c in enumerate(tab["attrs"][f"{name}_kind"]):
                seen.add(c)
                v, col = rec[name][:, j], df[c]
                kind = tab["attrs"][f"{name}_dtype"]
                if kind == "float64":
                    assert np.array_equal(v, col.to_numpy(dtype=float), equal_nan=True)
                elif kind == "int64":
                    assert np.array_equal(v, col.to_numpy())
                elif kind == "datetime64[ns]":
                    assert np.array_equal(v.view("datetime64[ns]"), col.to_numpy().astype("datetime64[ns]"))
                else:
                    want = ["nan" if (x is None or (isinstance(x, float) and np.isnan(x))) else x for x in col]
                    assert [x.decode("utf-8") for x in v] == want
        assert seen == set(df.columns)
    run()
