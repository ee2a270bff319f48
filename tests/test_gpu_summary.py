"""summary() on the device: the bootstrap kernel (stryke_b200_bootstrap_means) against NumPy resampling of the same series
(reference: bootstrap_mean_ci, Stryke/stryke.py:143-162)."""
import numpy as np
import pytest

from stryke_b200 import engine, summary

pytestmark = pytest.mark.gpu


def test_resampled_means_of_degenerate_series_are_exact():
    assert np.array_equal(engine.bootstrap_means([0.375], 50, seed=1), np.full(50, 0.375))
    assert np.array_equal(engine.bootstrap_means(np.full(1000, 2.5), 64, seed=2), np.full(64, 2.5))
    ints = engine.bootstrap_means(np.arange(7, dtype=float), 500, seed=3) * 7       # sums of 7 integers in [0, 6]
    assert np.abs(ints - np.round(ints)).max() < 1e-12 and ints.min() >= 0 and ints.max() <= 42.0 + 1e-12


def test_same_seed_same_means_other_seed_other_means():
    x = np.random.default_rng(0).random(5000)
    a, b, c = (engine.bootstrap_means(x, 300, seed=s) for s in (11, 11, 12))
    assert np.array_equal(a, b) and not np.array_equal(a, c)


@pytest.mark.parametrize("n,n_boot", [(37, 4000), (1000, 4000), (36500, 2000)])
def test_resampled_means_have_the_bootstrap_distribution(n, n_boot):
    """Mean of the resampled means = sample mean, their variance = population variance / n (both within the Monte
    Carlo error of n_boot draws), every index equally likely, and the 95 % interval agrees with NumPy's."""
    rng = np.random.default_rng(n)
    x = np.where(rng.random(n) < 0.6, 0.0, rng.gamma(0.7, 30.0, n))      # zero-inflated daily counts, like Daily.num_entrained
    m = engine.bootstrap_means(x, n_boot, seed=99)
    se = x.std() / np.sqrt(n)
    assert abs(m.mean() - x.mean()) < 5 * se / np.sqrt(n_boot)
    assert abs(m.std(ddof=1) / se - 1) < 5 / np.sqrt(2 * n_boot)
    ref = x[rng.integers(0, n, size=(n_boot, n))].mean(axis=1)
    for q in (2.5, 97.5):       # percentile estimators of two independent bootstrap samples: sd ~ 2.1 se / sqrt(n_boot)
        assert abs(np.percentile(m, q) - np.percentile(ref, q)) < 6 * 2.1 * se / np.sqrt(n_boot)
    # index uniformity: resampling the identity series gives means whose average is (n - 1) / 2
    ident = engine.bootstrap_means(np.arange(n, dtype=float), n_boot, seed=5)
    assert abs(ident.mean() - (n - 1) / 2) < 5 * np.sqrt((n * n - 1) / 12 / n / n_boot)


def test_bootstrap_mean_ci_switches_to_the_device_for_large_series(monkeypatch):
    x = np.random.default_rng(4).poisson(3.0, 40000).astype(float)
    np.random.seed(7)
    mean_np, lo_np, hi_np = summary.bootstrap_mean_ci(x, n_bootstrap=1000)             # NumPy path (no device set)
    monkeypatch.setattr(summary, "_DEVICE", 0)
    before = engine.launch_count()
    np.random.seed(7)
    mean_d, lo_d, hi_d = summary.bootstrap_mean_ci(x, n_bootstrap=1000)
    assert engine.launch_count() == before + 1
    se = x.std() / np.sqrt(len(x))
    assert mean_d == mean_np and abs(lo_d - lo_np) < 0.5 * se and abs(hi_d - hi_np) < 0.5 * se
    assert lo_d < mean_d < hi_d
    skipped = summary.bootstrap_mean_ci(x, n_bootstrap=1000, result_used=False)
    assert engine.launch_count() == before + 1 and skipped[0] == mean_np
