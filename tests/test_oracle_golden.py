"""Pins oracle/stryke_oracle.py to the reference: replaying the slot-tagged draws recorded from the UNMODIFIED
reference (tests/golden/*.npz, oracle/make_goldens.py) must reproduce the reference's Daily and State_Daily rows and
every per-fish state, fp32 survival rate and survival flag bit-for-bit."""
import numpy as np
import pytest

from oracle import stryke_oracle as O
from golden_util import CASES, Golden


@pytest.mark.parametrize("case", CASES)
def test_oracle_replays_reference_bit_exact(case):
    g = Golden(case)
    n_sim = 0
    for ci, cm in enumerate(g.cells):
        ref_daily = g.ref(ci, "daily")
        prow = g.pops[(cm["species"], cm["scenario"])]
        occur = 1.0 if O._isnan(prow["occur_prob"]) else prow["occur_prob"]
        if cm["n"] == 0:
            assert not (occur >= cm["presence"])
            assert (ref_daily == 0).all()
            continue
        assert occur >= cm["presence"]
        if not O._isnan(prow["shape"]):   # population mode: entrainment-count restatement
            rate = O.entrainment_rate(prow["dist"], prow["shape"], prow["location"], prow["scale"],
                                      prow["max_ent_rate"], cm["count_draw"])
            assert O.daily_fish_count(rate, cm["curr_Q"]) == cm["n"]
        res = O.simulate_cell(g.model, g.species(cm), cm["curr_Q"], cm["scenario"], g.draws(ci))
        assert np.array_equal(res.daily, ref_daily), (ci, res.daily, ref_daily)
        assert res.state_daily == g.ref_state_daily(ci), ci
        assert np.array_equal(res.states, g.ref(ci, "states"))
        assert np.array_equal(res.rates, g.ref(ci, "rates"))          # fp32, bit-exact
        assert np.array_equal(res.survival, g.ref(ci, "survival"))
        n_sim += 1
    assert n_sim > 0


@pytest.mark.parametrize("case", CASES)
def test_oracle_routing_tables_match_reference_movement(case):
    """(locs, p) captured from the reference's own ``np.random.choice`` call sites (stryke.py:2058)."""
    g = Golden(case)
    for r in g.meta["routing"]:
        cm = g.cells[r["cell"]]
        rp = O.route_probs(g.model, r["location"], cm["curr_Q"], cm["scenario"])
        assert rp is not None
        locs, probs, _ = rp
        assert locs == r["locs"]
        assert np.array_equal(probs, np.array(r["probs"]))


@pytest.mark.parametrize("case", CASES)
def test_oracle_daily_hours_match_reference(case):
    g = Golden(case)
    for cm in g.cells:
        tot, hours = O.daily_hours_run_of_river(g.model, cm["curr_Q"], cm["scenario"])
        assert tot == cm["tot_hours"]
        assert hours == cm["hours"]
