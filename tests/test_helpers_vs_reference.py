"""Helper methods of the class API called side by side with the reference's own methods (live, under the harness) on
randomised facilities and flows: compute_unit_flow_allocations (stryke.py:1657), movement's recorded route flows
(:1823), daily_hours (:2309).  Needs the reference tree (build container); runs in the CPU suite."""
import contextlib
import io

import numpy as np
import pytest

from oracle import fixtures, harness

pytestmark = pytest.mark.skipif(not harness.reference_available(), reason="needs /root/reference (build container only)")


def both_sims(prj, tmp_path):
    from stryke_b200 import simulation
    from stryke_b200.tables import RunContext
    S = harness.load_reference()
    ref = S.simulation.__new__(S.simulation)
    fixtures.apply_to(ref, prj, proj_dir=str(tmp_path), output_name="ref")
    ours = simulation.__new__(simulation)
    fixtures.apply_to(ours, prj, proj_dir=str(tmp_path), output_name="ours")
    with contextlib.redirect_stdout(io.StringIO()):
        ref.create_route()
        ours.create_route()
    return ref, ours, RunContext(ours)


@pytest.mark.parametrize("seed", range(12))
def test_allocations_route_flows_and_hours_equal_the_reference(seed, tmp_path):
    prj = fixtures.random_facility(seed, n_fish=10, days=6, baro=bool(seed % 2))
    if seed % 3 == 0:       # shared penstocks on some facilities
        for u in prj["unit_params"]:
            u.update(Penstock_ID=f"{u['Facility']}-P{(u['Unit'] + 1) // 2}", Penstock_Qcap=float(u["Qcap"]) * 1.4)
    ref, ours, ctx = both_sims(prj, tmp_path)
    rng = np.random.default_rng(seed)
    flows = [h["Discharge"] for h in prj["hydrograph"]] + [float(rng.uniform(0, 40000)) for _ in range(4)]
    for q in flows:
        Q = ctx.q_dict(q, "Spring")
        args = (Q["curr_Q"], ctx.units, Q["min_Q"], Q["env_Q"], Q["bypass_Q"], Q["sta_cap"], ctx.unit_fac, ctx.penstock_id,
                ctx.penstock_cap, ctx.q_cap)
        a = ref.compute_unit_flow_allocations(*args)
        b = ours.compute_unit_flow_allocations(*args)
        assert {k: float(v) for k, v in a[0].items()} == {k: float(v) for k, v in b[0].items()}
        assert {k: float(v) for k, v in a[1].items()} == {k: float(v) for k, v in b[1].items()}
        assert float(a[2]) == float(b[2]) and float(a[3]) == float(b[3])
        with contextlib.redirect_stdout(io.StringIO()):
            ha, hb = ref.daily_hours(dict(Q), "Spring"), ours.daily_hours(dict(Q), "Spring")
        assert float(np.sum(ha[0])) == float(np.sum(hb[0])) and float(np.sum(ha[1])) == float(np.sum(hb[1]))
        assert {str(k): float(v) for k, v in ha[2].items()} == {str(k): float(v) for k, v in hb[2].items()}
        assert {str(k): float(v) for k, v in ha[3].items()} == {str(k): float(v) for k, v in hb[3].items()}
        for node in list(ours.graph.nodes):
            if ours.graph.out_degree(node) == 0:
                continue
            ra, rb = {}, {}
            np.random.seed(5)
            na = ref.movement(node, 1, 1.0, ref.graph, ctx.intake_vel, dict(Q), ctx.op_order, ctx.q_cap, ctx.unit_fac,
                              ctx.penstock_id, ctx.penstock_cap, route_flow_recorder=ra)
            np.random.seed(5)
            nb = ours.movement(node, 1, 1.0, ours.graph, ctx.intake_vel, dict(Q), ctx.op_order, ctx.q_cap, ctx.unit_fac,
                               ctx.penstock_id, ctx.penstock_cap, route_flow_recorder=rb)
            assert {k: float(v) for k, v in ra.items()} == {k: float(v) for k, v in rb.items()}, (q, node)
            assert str(na) == str(nb), (q, node)      # same legacy-stream uniform -> same successor
