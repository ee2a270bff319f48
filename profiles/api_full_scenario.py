"""The class API at the reference's full single-scenario scale on one host + one GPU: 20 species x 365 days x 1000
iterations = 7.3e6 cells (C3 is five such scenarios).  simulation.run() and summary() from a webapp_import-style payload;
writes wall-clock times, frame sizes and peak RSS to gpurun_out/api_full_scenario.json.

    python profiles/api_full_scenario.py [iterations]
"""
import contextlib
import io
import json
import os
import resource
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from stryke_b200 import workloads as W
    from stryke_b200.store import open_store, wait_for_exports
    iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    os.environ.setdefault("STRYKE_MAX_DAILY_FISH", "2000000000")
    os.environ.setdefault("STRYKE_MAX_DAILY_STATE_ROWS", "2000000000000")
    prj = W.c3_population(n_species=20, iterations=iterations, days=365)
    tmp = tempfile.mkdtemp(prefix="api_full_")
    sim = W.make_sim(prj, "api", proj_dir=tmp)
    inputs = (("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
              ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges), ("Unit_Parameters", sim.unit_params),
              ("Facilities", sim.facility_params), ("Hydrograph", sim.input_hydrograph_df))
    with open_store(sim.hdf_path, mode="w") as st:
        for key, df in inputs:
            st[key] = df
    out = {"cells": 20 * 365 * iterations}
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        sim.run()
        t1 = time.perf_counter()
        sim.summary()
        t2 = time.perf_counter()
        wait_for_exports()
        t3 = time.perf_counter()
    fish = int(sum(int(t.n.sum()) for t in sim._last_tallies.values()))
    with open_store(sim.hdf_path, mode="r") as st:
        rows = {k: int(st._manifest["keys"][k]["rows"]) for k in st.keys()} if hasattr(st, "_manifest") else {}
    out.update({"fish_passages": fish, "run_s": t1 - t0, "summary_s": t2 - t1, "h5_export_wait_s": t3 - t2,
                "run_fish_passages_per_s": fish / (t1 - t0), "rows": rows,
                "h5_bytes": os.path.getsize(sim.hdf_path) if os.path.isfile(sim.hdf_path) else 0,
                "peak_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6,
                "yearly_summary_rows": int(len(sim.cum_sum)), "beta_rows": int(len(sim.beta_df))})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "api_full_scenario.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
