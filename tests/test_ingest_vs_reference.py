"""webapp_import against the reference's own webapp_import (stryke.py:489-796) on the same payload: every attribute the
run reads afterwards must hold the same frame / value.  Needs the reference tree (build container), so it is skipped on
the GPU box; it runs in the CPU suite."""
import contextlib
import io
import os

import numpy as np
import pandas as pd
import pytest

from oracle import fixtures, harness

pytestmark = pytest.mark.skipif(not harness.reference_available(), reason="needs /root/reference (build container only)")

FRAMES = ("nodes", "edges", "unit_params", "facility_params", "flow_scenarios_df", "operating_scenarios_df", "pop",
          "input_hydrograph_df")


def payload(prj, tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ingest_helpers", os.path.join(os.path.dirname(__file__), "test_gpu_ingest.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.webapp_project(prj, tmp_path)


def same_frame(a, b, name):
    assert list(a.columns) == list(b.columns), (name, list(a.columns), list(b.columns))
    assert [str(i) for i in a.index] == [str(i) for i in b.index], name
    for c in a.columns:
        x, y = a[c].to_numpy(), b[c].to_numpy()
        if np.issubdtype(np.asarray(x).dtype, np.number) and np.issubdtype(np.asarray(y).dtype, np.number):
            assert np.allclose(x.astype(float), y.astype(float), rtol=0, atol=0, equal_nan=True), (name, c)
        else:
            assert [str(v) for v in x] == [str(v) for v in y], (name, c)


@pytest.mark.parametrize("fixture,kw", [("c2_facility", dict(n_fish=50, iterations=2, days=3, baro=True)),
                                        ("c5_cascade", dict(n_fish=20, iterations=1, days=2, dams=3)),
                                        ("c3_population", dict(n_species=3, iterations=2, days=4))])
def test_webapp_import_builds_the_reference_attributes(fixture, kw, tmp_path):
    from stryke_b200 import simulation
    prj = getattr(fixtures, fixture)(**kw)
    (tmp_path / "ref").mkdir()
    (tmp_path / "ours").mkdir()
    S = harness.load_reference()
    data_ref, _ = payload(prj, tmp_path / "ref")
    ref = S.simulation.__new__(S.simulation)
    ref.proj_dir, ref.output_name = str(tmp_path / "ref"), "ref"
    with contextlib.redirect_stdout(io.StringIO()):
        ref.webapp_import(data_ref, "ref")
    data_our, _ = payload(prj, tmp_path / "ours")
    ours = simulation.__new__(simulation)
    ours.proj_dir, ours.output_name = str(tmp_path / "ours"), "ours"
    with contextlib.redirect_stdout(io.StringIO()):
        ours.webapp_import(data_our, "ours")
    for name in FRAMES:
        assert hasattr(ref, name) == hasattr(ours, name), name
        if hasattr(ref, name):
            same_frame(getattr(ours, name), getattr(ref, name), name)
    assert ours.output_units == ref.output_units
    assert list(ours.moves) == list(ref.moves)
    assert sorted(ours.graph.edges(data="weight")) == sorted(ref.graph.edges(data="weight"))
    assert list(ours.graph.nodes) == list(ref.graph.nodes)
    assert ours.surv_fun_dict == ref.surv_fun_dict
    assert list(ours.flow_scenarios) == list(ref.flow_scenarios)
    assert getattr(ours, "max_river_node", None) == getattr(ref, "max_river_node", None)
