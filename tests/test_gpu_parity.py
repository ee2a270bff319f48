"""GPU parity tests (run on the B200 box with ``-m gpu``); every call goes through the C ABI
(``stryke_b200_run`` / ``stryke_b200_strike_survival`` / ...), never through the oracle.

* injected-uniform replay of the UNMODIFIED reference's recordings (tests/golden/*.npz): pop sizes, Daily and
  State_Daily tallies, per-fish states, survival flags and fp32 rates must be bit-exact in fp64 mode;
* per-fish strike / barotrauma probabilities: <= 1e-6 relative (fp64), <= 1e-4 relative (fp32) — tolerances from
  BASELINE.json:north_star;
* oracle-generated draws at sizes the reference cannot reach (2e5 fish/cell): tallies bit-exact.
"""
import numpy as np
import pytest

from oracle import fixtures, stryke_oracle as O
from golden_util import CASES, Golden
import gpu_util as G

pytestmark = pytest.mark.gpu

KAPLAN = dict(H=40.0, RPM=120.0, D=12.0, Q=2500.0, ada=0.9, N=5.0, Qopt=2500.0, Qper=2500.0 / 3000.0, _lambda=0.2)
FRANCIS = dict(H=60.0, RPM=97.3, D=10.76, Q=2150.72, Qper=0.94, ada=0.94, N=13.0, iota=1.1, D1=11.22, D2=10.76, B=4.0,
               _lambda=0.2, Qopt=2150.72)
PUMP = dict(FRANCIS, Q_p=1800.0, gamma=0.153)
KEYS15 = ("H", "RPM", "D", "Q", "ada", "N", "Qopt", "Qper", "_lambda", "iota", "D1", "D2", "B", "Q_p", "gamma")


def p15(d):
    return np.array([float(d.get(k, np.nan)) for k in KEYS15])


@pytest.mark.parametrize("case", CASES)
def test_fp64_injected_replay_matches_reference_bit_exact(case):
    g = Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    inj, off, ref_idx = G.golden_injected(g, plan)
    res = G.run_injected(plan, inj, precision=64)
    tr = res.trace
    n_checked = 0
    for c, r in enumerate(ref_idx):
        cm = g.cells[r]
        assert int(res.cell_n[c]) == cm["n"], (case, c, r)
        assert np.array_equal(G.full_daily(res.cell_n[c], res.daily[c]), g.ref(r, "daily")), (case, c, r)
        if cm["n"] == 0:
            assert res.state_daily[c].sum() == 0
            continue
        assert G.state_daily_dict(plan, res.state_daily[c]) == g.ref_state_daily(r), (case, c, r)
        s = slice(off[c], off[c + 1])
        assert np.array_equal(tr.state[:, s], g.ref(r, "states")), (case, c)
        assert np.array_equal(tr.survival[:, s], g.ref(r, "survival")), (case, c)
        # CUDA's exp() and glibc's differ in the last ulp for some arguments; |loc + scale*exp(s z)| keeps that
        np.testing.assert_allclose(tr.length_ft64[s], g.draws(r).lengths_ft, rtol=1e-14, atol=0)
        assert np.array_equal(tr.length_ft[s], np.float32(g.draws(r).lengths_ft)) or \
            (tr.length_ft[s] != np.float32(g.draws(r).lengths_ft)).mean() < 1e-3
        ref_rates = g.ref(r, "rates")
        mism = tr.rate[:, s] != ref_rates
        # fp32 rounding of a probability that differs in the last fp64 ulp can flip one fp32 ulp; never more
        assert mism.mean() < 1e-3, (case, c, mism.sum())
        np.testing.assert_allclose(tr.rate[:, s], ref_rates, rtol=1.3e-7, atol=0)
        n_checked += 1
    assert n_checked > 0


@pytest.mark.parametrize("case", ["c2_kaplan_francis", "c2_barotrauma", "c4_propeller_baro", "c5_cascade", "c1_single_unit",
                                  "cabot_station_fish", "example_v250304"])
@pytest.mark.parametrize("precision,tol", [(64, 1e-6), (32, 1e-4)])
def test_per_fish_probabilities_within_tolerance(case, precision, tol):
    """Strike and barotrauma survival AND mortality probabilities of every evaluated (fish, move)."""
    g = Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    inj, off, ref_idx = G.golden_injected(g, plan)
    res = G.run_injected(plan, inj, precision=precision)
    tr = res.trace
    seen = 0
    for c, r in enumerate(ref_idx):
        cm = g.cells[r]
        if cm["n"] == 0:
            continue
        ora = O.simulate_cell(g.model, g.species(cm), cm["curr_Q"], cm["scenario"], g.draws(r))
        s = slice(off[c], off[c + 1])
        same_path = (tr.state[:, s] == ora.states) & (ora.survival == tr.survival[:, s])
        for dev, ref in ((tr.strike[:, s], ora.strike), (tr.baro[:, s], ora.baro)):
            m = ~np.isnan(ref) & same_path & ~np.isnan(dev)
            if precision == 64:
                assert np.array_equal(np.isnan(ref), np.isnan(dev))
            if m.any():
                np.testing.assert_allclose(dev[m], ref[m], rtol=tol, atol=0)
                mort = ref[m] < 1.0          # relative error of the mortality probability itself
                if mort.any():
                    np.testing.assert_allclose(1.0 - dev[m][mort], 1.0 - ref[m][mort], rtol=tol, atol=tol * 1e-3)
                seen += int(m.sum())
    assert seen > 0


@pytest.mark.parametrize("case", ["c2_kaplan_francis", "c2_barotrauma", "c5_cascade", "cabot_station_fish"])
def test_fp32_injected_tallies_statistically_identical(case):
    """fp32 arithmetic with the reference's uniforms: a survival flag can only flip when a dice lands within one
    fp32 ulp of the rate, so tallies agree up to a handful of fish."""
    g = Golden(case)
    sim, plan = G.make_plan(g.prj, case)
    inj, off, ref_idx = G.golden_injected(g, plan)
    res = G.run_injected(plan, inj, precision=32, trace=False)
    for c, r in enumerate(ref_idx):
        ref = g.ref(r, "daily")
        got = G.full_daily(res.cell_n[c], res.daily[c])
        assert got[0] == ref[0]
        assert np.abs(got[1:] - ref[1:]).max() <= 2, (case, c, got, ref)


@pytest.mark.parametrize("fixture,kw", [
    ("c2_facility", dict(n_fish=200000, iterations=2, days=2)),
    ("c2_facility", dict(n_fish=100000, iterations=1, days=2, baro=True)),
    ("c5_cascade", dict(n_fish=60000, iterations=1, days=2)),
    ("c4_barotrauma", dict(n_fish=50000, iterations=2, days=1, habitat="Benthic")),
    # BASELINE.json configs[3] as named: 3 Propeller + 3 Pump units (opt-in Pump runner, Q_p / gamma columns) and the
    # friction-loss pressure mode; the oracle composes Stryke/barotrauma.py's calc_* helpers as stryke_barotrauma.py:565-601
    ("c4_barotrauma", dict(n_fish=40000, iterations=1, days=3, pump=True, friction=True)),
    ("c4_barotrauma", dict(n_fish=40000, iterations=1, days=2, pump=True, habitat="Benthic")),
    ("cabot_station", dict(iterations=1, days=3, seed=5, median_Q=6000.0, fish=30000.0)),
    ("c5_cascade", dict(n_fish=8000, iterations=1, days=2, dams=9)),
] + [("random_facility", dict(seed=s, n_fish=15000, days=4, baro=bool(s % 2))) for s in range(10)])
def test_large_cells_oracle_draws_bit_exact(fixture, kw):
    """Sizes the reference cannot reach in test time: the oracle generates the slot-tagged draws, the GPU (fp64,
    injected) must return identical tallies, states and survival flags."""
    prj = getattr(fixtures, fixture)(**kw)
    model = O.Model.from_project(prj)
    sim, plan = G.make_plan(prj, fixture)
    ref_idx = G.reference_cell_index(plan)
    rng = np.random.default_rng(1234)
    M = len(model.moves)
    prow = prj["population"][0]
    sp = O.species_record(model, prow, {"Micropterus": "Centrarchidae"}.get(prow["Species"]))
    n = int(prow["Fish"])
    n_ref = plan.n_cells
    draws, results = [], []
    days = len(prj["hydrograph"])
    for r in range(n_ref):
        z = rng.standard_normal(n)
        L = O.lengths_from_normals(z, prow["length shape"], prow["length location"], prow["length scale"])
        d = O.Draws.random(rng, M, n, L, ghost=True)
        draws.append((d, z))
        q = prj["hydrograph"][r % days]["Discharge"]
        results.append(O.simulate_cell(model, sp, q, prow["Scenario"], d))
    inj, off = G.injected_from_cells(plan, ref_idx, draws, [0.0] * n_ref, [None] * n_ref, M)
    res = G.run_injected(plan, inj, precision=64)
    for c, r in enumerate(ref_idx):
        ora = results[r]
        assert np.array_equal(G.full_daily(res.cell_n[c], res.daily[c]), ora.daily), (c, r)
        assert G.state_daily_dict(plan, res.state_daily[c]) == ora.state_daily
        s = slice(off[c], off[c + 1])
        assert np.array_equal(res.trace.state[:, s], ora.states)
        assert np.array_equal(res.trace.survival[:, s], ora.survival)
        # per-fish probabilities where both sides evaluated them (strike: Kaplan / Propeller / Francis / Pump; barotrauma:
        # hydrostatic or friction-loss pressures): 1e-6 relative, the fp64 tolerance of BASELINE.json
        for dev, ref in ((res.trace.strike[:, s], ora.strike), (res.trace.baro[:, s], ora.baro)):
            both = np.isfinite(dev) & np.isfinite(ref)
            assert both.sum() == np.isfinite(ref).sum()
            np.testing.assert_allclose(dev[both], ref[both], rtol=1e-6, atol=1e-12)


def test_strike_known_answer_vectors():
    """SURVEY.md Appendix C (values computed by calling the reference's Kaplan/Propeller/Francis/Pump)."""
    from stryke_b200 import engine
    L = np.array([0.25, 0.5, 1.0, 2.0])
    kat = {
        ("Kaplan", 0.30): [0.9707441172827377, 0.9414882345654753, 0.8829764691309505, 0.7659529382619011],
        ("Kaplan", 0.65): [0.9771016930814648, 0.9542033861629295, 0.908406772325859, 0.816813544651718],
        ("Kaplan", 1.00): [0.9776439071793664, 0.9552878143587329, 0.9105756287174657, 0.8211512574349313],
        ("Propeller", None): [0.9774022968174205, 0.9548045936348412, 0.9096091872696823, 0.8192183745393646],
        ("Francis", None): [0.9336025523008413, 0.8672051046016827, 0.7344102092033654, 0.46882041840673083],
        ("Pump", None): [0.9184704362836714, 0.8369408725673428, 0.6738817451346856, 0.34776349026937114],
    }
    kinds = {"Kaplan": (1, KAPLAN), "Propeller": (2, KAPLAN), "Francis": (3, FRANCIS), "Pump": (4, PUMP)}
    for (name, rR), want in kat.items():
        k, params = kinds[name]
        for precision, tol in ((64, 1e-12), (32, 1e-6)):
            got = engine.strike_survival(k, p15(params), L, None if rR is None else np.full(4, rR), precision=precision)
            np.testing.assert_allclose(got, want, rtol=tol)
            np.testing.assert_allclose(1 - got, 1 - np.array(want), rtol=1e-6 if precision == 64 else 1e-4)


def test_barotrauma_known_answer_vectors():
    from stryke_b200 import engine
    for depth, tail, b0, b1, injury in ((9.9, 6.0, -5.301, 4.825, 0.007754677794343372),
                                        (30.0, 2.0, -4.657, 3.41, 0.0633677711333838),
                                        (0.3, 6.0, -3.308, 6.195, 0.013906127661367678)):
        got64 = engine.baro_survival([depth], tail, b0, b1, precision=64)[0]
        got32 = engine.baro_survival([depth], tail, b0, b1, precision=32)[0]
        assert abs((1 - got64) - injury) <= 1e-6 * injury
        assert abs((1 - got32) - injury) <= 1e-4 * injury


def _philox_np(ctr, key):
    c = [np.uint64(x) for x in ctr]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, W0, W1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    return [int(x) for x in c]


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10, plus the NumPy restatement on random counters."""
    from stryke_b200 import engine
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert tuple(_philox_np(ctr, key)) == want
        got = engine.philox(np.array([ctr], np.uint32), np.array([key], np.uint32))
        assert tuple(int(x) for x in got[0]) == want
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, (64, 2), dtype=np.uint64).astype(np.uint32)
    got = engine.philox(ctr, key)
    for i in range(64):
        assert [int(x) for x in got[i]] == _philox_np(ctr[i], key[i])


def _function_vectors():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "function_vectors.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("precision,tol", [(64, 1e-6), (32, 1e-4)])
def test_strike_functions_against_240_reference_parameter_sets(precision, tol):
    """tests/golden/function_vectors.json: the reference's Kaplan / Propeller / Francis / Pump called on 60 random
    parameter sets each (oracle/make_function_goldens.py).  Tolerances of BASELINE.json on the survival probability and
    on its complement (the strike probability), wherever the latter is not clipped."""
    from stryke_b200 import engine
    kinds = {"Kaplan": 1, "Propeller": 2, "Francis": 3, "Pump": 4}
    worst = 0.0
    for case in _function_vectors()["strike"]:
        want = np.array(case["survival"])
        got = engine.strike_survival(kinds[case["kind"]], np.array(case["params"]), np.array(case["length"]),
                                     np.array(case["rR"]), precision=precision)
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok)
        p_want, p_got = 1.0 - want[ok], 1.0 - got[ok]
        scale = np.maximum(np.abs(p_want), 1e-3)          # strike probability; survival may be far below 0 (reference does not clip)
        err = np.abs(p_got - p_want) / scale
        worst = max(worst, float(err.max()) if err.size else 0.0)
    assert worst <= tol, worst


@pytest.mark.parametrize("precision,tol", [(64, 1e-6), (32, 1e-4)])
def test_barotrauma_response_against_200_reference_cases(precision, tol):
    from stryke_b200 import engine
    worst = 0.0
    for c in _function_vectors()["baro"]:
        got = engine.baro_survival([c["depth_ft"]], c["tail_ft"], c["beta_0"], c["beta_1"], precision=precision)[0]
        worst = max(worst, abs((1.0 - got) - c["injury"]) / max(c["injury"], 1e-3))
    assert worst <= tol, worst
