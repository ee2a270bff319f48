"""summary() against the reference's own summary() (tests/golden/summary_*.json, oracle/make_summary_golden.py): the
frames the reference's run() wrote go into our store, np.random is seeded as it was for the reference, and every frame
our summary() writes must equal the reference's — means, standard deviations, bootstrap intervals (the NumPy path
consumes the MT19937 stream exactly like the reference's loop), yearly quantiles and the driver diagnostics."""
import contextlib
import io
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import fixtures
from stryke_b200 import simulation
from stryke_b200.store import open_store

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(f[:-5] for f in os.listdir(os.path.join(HERE, "golden")) if f.startswith("summary_") and f.endswith(".json"))


def frame(obj):
    df = pd.DataFrame(obj["data"], columns=obj["columns"], index=obj["index"])
    if "day" in df.columns:
        df["day"] = pd.to_datetime(df["day"])
    return df


@pytest.mark.parametrize("case", CASES)
def test_summary_frames_equal_the_reference_frames(case, tmp_path):
    with open(os.path.join(HERE, "golden", case + ".json")) as fh:
        g = json.load(fh)
    prj = getattr(fixtures, g["fixture"])(**g["kwargs"])
    sim = simulation.__new__(simulation)
    fixtures.apply_to(sim, prj, proj_dir=str(tmp_path), output_name=case)
    with open_store(sim.hdf_path, mode="w") as st:
        for key, df in (("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
                        ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges), ("Unit_Parameters", sim.unit_params),
                        ("Facilities", sim.facility_params), ("Hydrograph", sim.input_hydrograph_df)):
            st[key] = df
        for key, obj in g["run"].items():
            st[key] = frame(obj).reset_index(drop=True)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.create_route()          # run() leaves sim.moves / sim.graph behind; summary() reads them (stryke.py:3604)
    np.random.seed(g["summary_seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        sim.summary()
    with open_store(sim.hdf_path, mode="r") as st:
        got = {k: st[k] for k in g["summary"] if k in st}
    assert set(got) == set(g["summary"])
    for key, obj in g["summary"].items():
        want, have = frame(obj), got[key]
        assert list(have.columns) == list(want.columns), key
        assert len(have) == len(want), key
        assert [str(i) for i in have.index] == [str(i) for i in want.index], key
        for col in want.columns:
            w, h = want[col].to_numpy(), have[col].to_numpy()
            if np.issubdtype(np.asarray(w).dtype, np.number) and np.issubdtype(np.asarray(h).dtype, np.number):
                assert np.allclose(h.astype(float), w.astype(float), rtol=1e-12, atol=1e-12, equal_nan=True), (key, col, h[:5], w[:5])
            else:
                assert [str(x).strip() for x in h] == [str(x).strip() for x in w], (key, col)
