"""stryke_b200.epri against the reference's own ``epri`` class (Stryke/stryke.py:4322-5084): tests/golden/epri_fits.json
holds, for twelve queries, what the unmodified class computed on its database (oracle/make_epri_goldens.py) and the
filtered entrainment rates it fitted."""
import importlib
import json
import os

import numpy as np
import pytest

E = importlib.import_module("stryke_b200.epri")     # the package attribute `epri` is the class, as in the reference

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "epri_fits.json")) as fh:
    GOLD = json.load(fh)
CASES = GOLD["cases"]
DATABASE = "/root/reference/Data/epri1997.csv"      # present in the build container only


@pytest.mark.parametrize("i", range(len(CASES)))
def test_fits_equal_the_reference_fits_on_the_same_rates(i):
    c = CASES[i]
    x = np.array(c["rates"])
    assert len(x) == c["sample_size"] and x.max() == c["max_ent_rate"]
    assert np.allclose(E.pareto_mle(x), c["pareto"], rtol=1e-12, atol=0)
    assert np.allclose(E.lognorm_mle(x), c["lognorm"], rtol=1e-12, atol=0)
    # scipy fits the Weibull with a Nelder-Mead simplex that stops at xtol = 1e-4: the exact root agrees to that
    # tolerance and its likelihood is at least as high
    mine, ref = E.weibull_min_mle(x), c["weibull"]
    assert np.allclose(mine, ref, rtol=5e-4, atol=0)
    loglik = lambda p: np.sum(np.log(p[0] / p[2]) + (p[0] - 1) * np.log(x / p[2]) - (x / p[2]) ** p[0])
    assert loglik(mine) >= loglik(ref) - 1e-9


def test_weibull_fit_solves_the_likelihood_equation():
    rng = np.random.default_rng(0)
    x = 0.02 * rng.weibull(0.6, 400)
    c, loc, scale = E.weibull_min_mle(x)
    lx = np.log(x)
    assert abs((x ** c * lx).sum() / (x ** c).sum() - 1 / c - lx.mean()) < 1e-10
    assert abs(scale - np.mean(x ** c) ** (1 / c)) < 1e-12 * scale and loc == 0.0


@pytest.mark.skipif(not os.path.isfile(DATABASE), reason="the EPRI database ships with the reference, not with this repository")
@pytest.mark.parametrize("i", range(len(CASES)))
def test_queries_select_the_reference_records(i):
    c = CASES[i]
    q = {k: (tuple(v) if k == "plant_cap" else v) for k, v in c["query"].items()}
    e = E.epri(database=DATABASE, quiet=True, **q)
    assert e.presence == c["presence"] and e.sample_size == c["sample_size"] and float(e.max_ent_rate) == c["max_ent_rate"]
    assert np.array_equal(np.asarray(e.epri.FishPerMft3.values, float), np.array(c["rates"]))
    assert e.cohort_counts().tolist() == c["cohorts"]
    e.ParetoFit(); e.LogNormalFit(); e.WeibullMinFit()
    assert np.allclose(e.dist_pareto, c["pareto"], rtol=1e-12) and np.allclose(e.dist_lognorm, c["lognorm"], rtol=1e-12)
    row = e.population_row("Pareto")
    assert row["dist"] == "Pareto" and row["occur_prob"] == c["presence"] and row["shape"] == e.dist_pareto[0]
    if c["extreme"] is not None and i < 3:
        assert np.allclose(e.ExtremeFit(), c["extreme"], rtol=1e-6)


@pytest.mark.skipif(not os.path.isfile(DATABASE), reason="the EPRI database ships with the reference, not with this repository")
def test_fit_table_batches_queries_and_length_summary_follows_the_cohorts():
    qs = [dict(label=f"q{i}", **{k: (tuple(v) if k == "plant_cap" else v) for k, v in c["query"].items()}) for i, c in enumerate(CASES)]
    qs.append(dict(label="empty", Genus="Nonexistentus"))
    tab = E.fit_table(qs, database=DATABASE)
    assert list(tab["label"]) == [q["label"] for q in qs]
    for i, c in enumerate(CASES):
        r = tab.iloc[i]
        assert r["sample_size"] == c["sample_size"] and r["occur_prob"] == c["presence"]
        assert np.isclose(r["pareto_shape"], c["pareto"][0], rtol=1e-12) and np.isclose(r["log_normal_scale"], c["lognorm"][2], rtol=1e-12)
        assert np.isclose(r["weibull_shape"], c["weibull"][0], rtol=5e-4)
    assert tab.iloc[-1]["sample_size"] == 0 and np.isnan(tab.iloc[-1]["pareto_shape"])
    e = E.epri(database=DATABASE, quiet=True, Genus="Micropterus", Month=[3, 4, 5], HUC02=[4])
    np.random.seed(3)
    s, loc, scale = e.LengthSummary()
    assert len(e.lengths) == sum(CASES[0]["cohorts"]) and e.lengths.min() >= 0 and e.lengths.max() <= 100
    assert s > 0 and scale > 0 and abs(e.length_mean_cm - e.lengths.mean()) < 1e-12
