"""Wall-clock profile of simulation.run() and simulation.summary() on a population-mode project (C3 shape):
    python profiles/time_api.py <iterations>
Measurement tool (uses the oracle fixtures to describe the project)."""
import sys, os, time, io, contextlib, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ.setdefault("STRYKE_MAX_DAILY_FISH", "2000000000")
from oracle import fixtures
import gpu_util as G
import cProfile, pstats
it = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prj = fixtures.c3_population(n_species=20, iterations=it, days=365, curr_Q=9000.0)
os.makedirs("/tmp/stryke_b200_tests", exist_ok=True)
sim = G.make_sim(prj, "api")
from stryke_b200.store import open_store
with open_store(sim.hdf_path, mode="w") as st:      # what webapp_import / worksheet_import write before run()
    for key, df in (("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
                    ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges), ("Unit_Parameters", sim.unit_params),
                    ("Facilities", sim.facility_params), ("Hydrograph", sim.input_hydrograph_df)):
        st[key] = df
t = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    sim.run()
pr.disable()
t1 = time.perf_counter()
print("run() %.2f s" % (t1 - t))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    sim.summary()
pr.disable()
print("summary() %.2f s" % (time.perf_counter() - t1))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
import resource
print("peak RSS %.2f GB" % (resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0))
