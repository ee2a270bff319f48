"""EPRI (1997) entrainment database queries and distribution fits — the step before ``simulation.run()``: it supplies the
Population sheet's ``shape / location / scale / dist / max_ent_rate / occur_prob`` and length parameters.

Reference: ``class epri`` (Stryke/stryke.py:4322-5084).  Same constructor arguments, attributes and method names for
the query and the fits the simulation can consume (Pareto, Log Normal, Weibull, Extreme); plotting and the HTML/Excel
report helpers are out of scope.  The reference calls ``scipy.stats.<dist>.fit(values, floc=0)``; here the maximum-
likelihood estimates are written out (closed form for Pareto and Log Normal, a Newton solve of the profile-likelihood
equation for Weibull), which is what scipy >= 1.9 computes analytically too and what its generic optimiser converges to
in the pinned 1.8.1.  ``fit_table`` fits many (genus, months, region) queries in one pass — the table the web app keeps
as ``species_defaults`` (webapp/app.py:2870-4018).

The database itself (``Data/epri1997.csv`` of the reference, 17 600 rows) is not part of this repository: pass
``database=`` (path or DataFrame) or set ``$STRYKE_EPRI_DATABASE``.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

COHORTS = (("0_5", 0.0, 5.0), ("5_10", 5.0, 10.0), ("10_15", 10.0, 15.0), ("15_20", 15.0, 20.0), ("20_25", 20.0, 25.0),
           ("25_38", 25.0, 38.0), ("38_51", 38.0, 51.0), ("51_64", 51.0, 64.0), ("64_76", 64.0, 76.0), ("GT76", 76.0, 100.0))


# ---------------------------------------------------------------------------------------------------------------------
# maximum-likelihood fits with the location fixed at 0 (scipy parameter order: shape, loc, scale)
# ---------------------------------------------------------------------------------------------------------------------
def pareto_mle(x):
    """``scipy.stats.pareto.fit(x, floc=0)``: scale = min(x), b = n / sum(log(x / scale))."""
    x = np.asarray(x, dtype=np.float64)
    scale = x.min()
    b = len(x) / np.sum(np.log(x / scale))
    return float(b), 0.0, float(scale)


def lognorm_mle(x):
    """``scipy.stats.lognorm.fit(x, floc=0)``: s = std(log x) (ddof 0), scale = exp(mean(log x))."""
    lx = np.log(np.asarray(x, dtype=np.float64))
    mu = lx.mean()
    return float(np.sqrt(np.mean((lx - mu) ** 2))), 0.0, float(np.exp(mu))


def weibull_min_mle(x, tol=1e-13, max_iter=200):
    """``scipy.stats.weibull_min.fit(x, floc=0)``: the shape solves  sum(x^c log x) / sum(x^c) - 1/c = mean(log x)
    (monotone in c: Newton from the method-of-moments start), scale = mean(x^c)^(1/c)."""
    x = np.asarray(x, dtype=np.float64)
    lx = np.log(x)
    m = lx.mean()
    z = lx - lx.max()                     # x^c / max^c: no overflow
    sd = lx.std()
    c = 1.2 / sd if sd > 0 else 1.0
    for _ in range(max_iter):
        w = np.exp(c * z)
        s0, s1, s2 = w.sum(), (w * lx).sum(), (w * lx * lx).sum()
        g = s1 / s0 - 1.0 / c - m
        dg = s2 / s0 - (s1 / s0) ** 2 + 1.0 / (c * c)
        step = g / dg
        c_new = c - step
        if c_new <= 0:
            c_new = c / 2
        if abs(c_new - c) <= tol * c:
            c = c_new
            break
        c = c_new
    scale = np.exp(lx.max()) * np.mean(np.exp(c * z)) ** (1.0 / c)
    return float(c), 0.0, float(scale)


def genextreme_mle(x):
    """``scipy.stats.genextreme.fit(x, floc=0)`` (no closed form; scipy's own optimiser, as in the reference)."""
    from scipy.stats import genextreme
    c, loc, scale = genextreme.fit(np.asarray(x, dtype=np.float64), floc=0)
    return float(c), float(loc), float(scale)


FITS = {"Pareto": pareto_mle, "Log Normal": lognorm_mle, "Weibull": weibull_min_mle, "Extreme": genextreme_mle}


def load_database(database=None):
    if isinstance(database, pd.DataFrame):
        return database
    path = database or os.environ.get("STRYKE_EPRI_DATABASE")
    if not path:
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "epri1997.csv")
        path = here if os.path.isfile(here) else None
    if not path or not os.path.isfile(path):
        raise FileNotFoundError("EPRI entrainment database not found: pass database=<path to epri1997.csv> or set "
                                "$STRYKE_EPRI_DATABASE (the file ships with the reference under Data/).")
    return pd.read_csv(path, encoding="unicode_escape")


def _select(df, column, value, scalar_type=str):
    """The reference's filter idiom (stryke.py:4384-4445): scalar -> equality, otherwise membership."""
    if value is None:
        return df
    if isinstance(value, scalar_type):
        return df[df[column] == value]
    return df[df[column].isin(value)]


class epri:
    """Query the database and fit entrainment-rate distributions (stryke.py:4322-5084)."""

    def __init__(self, states=None, plant_cap=None, Family=None, Genus=None, Species=None, Month=None, HUC02=None,
                 HUC04=None, HUC06=None, HUC08=None, NIDID=None, River=None, database=None, quiet=False):
        df = load_database(database)
        if NIDID is not None:                               # excluded, not selected (stryke.py:4378-4382)
            df = df[df.NIDID != NIDID] if isinstance(NIDID, str) else df[~df["NIDID"].isin(NIDID)]
        df = _select(df, "State", states)
        if plant_cap is not None:
            df = df[df.Plant_cap_cfs > plant_cap[0]] if plant_cap[1] == ">" else df[df.Plant_cap_cfs <= plant_cap[0]]
        for column, value in (("Month", Month), ("Family", Family), ("Species", Species), ("Genus", Genus),
                              ("HUC02", HUC02), ("HUC04", HUC04), ("HUC06", HUC06)):
            df = _select(df, column, value)
        df = _select(df, "HUC08", HUC08, scalar_type=int)
        df = _select(df, "River", River)
        success, trials = df.Present.sum(), len(df)
        self.presence = round(float(success / trials), 4) if trials else float("nan")
        df = df.copy()
        numeric = df.select_dtypes("number").columns      # the reference's blanket fillna(0) (stryke.py:4453) breaks on pandas-3
        df[numeric] = df[numeric].fillna(0)               # string columns; only the numeric ones are ever used afterwards
        self.epri = df[df.Present == 1].copy()
        self.family, self.genus, self.species, self.month = Family, Genus, Species, Month
        self.HUC02, self.HUC04 = HUC02, HUC04
        self.max_ent_rate = self.epri.FishPerMft3.max()
        self.sample_size = len(self.epri)
        self.dist_pareto = self.dist_lognorm = self.dist_weibull = self.dist_extreme = None
        if not quiet:
            bar = "-" * 92
            print(bar)
            print("out of %s potential samples %s had this species present for %s probability of presence"
                  % (trials, success, self.presence))
            print(bar)
            print("There are %s records left to describe entrainment rates" % len(self.epri))
            print("The maximum entrainment rate for this fish is: %s" % self.max_ent_rate)
            print(bar)

    # ---- fits (stryke.py:4451-4634) ------------------------------------------------------------------------------------
    def _rates(self):
        x = np.asarray(self.epri.FishPerMft3.values, dtype=np.float64)
        if x.size == 0:
            raise ValueError("no entrainment records left to fit")
        return x

    def ParetoFit(self):
        self.dist_pareto = pareto_mle(self._rates())
        return self.dist_pareto

    def LogNormalFit(self):
        self.dist_lognorm = lognorm_mle(self._rates())
        return self.dist_lognorm

    def WeibullMinFit(self):
        self.dist_weibull = weibull_min_mle(self._rates())
        return self.dist_weibull

    def ExtremeFit(self):
        self.dist_extreme = genextreme_mle(self._rates())
        return self.dist_extreme

    # ---- lengths (stryke.py:4655-4714) ---------------------------------------------------------------------------------
    def cohort_counts(self):
        return np.array([np.int32(self.epri[c].sum()) if c in self.epri.columns else 0 for c, _, _ in COHORTS], dtype=np.int64)

    def LengthSummary(self):
        """Lengths drawn uniformly inside the database's size cohorts (same np.random call order as the reference), then
        a three-parameter lognormal fit (scipy, as the reference) -> ``len_dist``, ``length_mean_cm``, ``length_sd_cm``."""
        from scipy.stats import lognorm
        counts = self.cohort_counts()
        if counts.sum() == 0:
            raise ValueError("No fish length data available to fit length distribution.")
        self.lengths = np.concatenate([np.random.uniform(low=lo, high=hi, size=int(n)) for (_, lo, hi), n in zip(COHORTS, counts)])
        self.len_dist = lognorm.fit(self.lengths)
        self.length_mean_cm = float(np.mean(self.lengths))
        self.length_sd_cm = float(np.std(self.lengths, ddof=1)) if self.lengths.size > 1 else 0.0
        return self.len_dist

    def population_row(self, dist="Log Normal"):
        """The Population-sheet entries this query produces."""
        fit = {"Pareto": self.dist_pareto or self.ParetoFit(), "Log Normal": self.dist_lognorm or self.LogNormalFit(),
               "Weibull": self.dist_weibull or self.WeibullMinFit(), "Extreme": self.dist_extreme}[dist]
        if fit is None:
            fit = self.ExtremeFit()
        return dict(shape=fit[0], location=fit[1], scale=fit[2], dist=dist, max_ent_rate=float(self.max_ent_rate),
                    occur_prob=self.presence)


def fit_table(queries, database=None, dists=("Pareto", "Log Normal", "Weibull")):
    """Fit many queries against one load of the database.  ``queries``: iterable of dicts of ``epri`` keyword arguments
    (plus an optional ``label``).  Returns one row per query: presence, sample size, maximum rate and (shape, location,
    scale) of every requested distribution — queries without records give NaNs instead of raising."""
    db = load_database(database)
    rows = []
    for q in queries:
        q = dict(q)
        label = q.pop("label", None)
        e = epri(database=db, quiet=True, **q)
        row = dict(label=label, **{k: (v if not isinstance(v, (list, tuple)) else ",".join(map(str, v))) for k, v in q.items()},
                   occur_prob=e.presence, sample_size=e.sample_size, max_ent_rate=float(e.max_ent_rate) if e.sample_size else np.nan)
        for d in dists:
            key = d.lower().replace(" ", "_")
            if e.sample_size >= 2 and e._rates().min() > 0 and np.ptp(e._rates()) > 0:
                s, loc, sc = FITS[d](e._rates())
            else:
                s = loc = sc = np.nan
            row.update({f"{key}_shape": s, f"{key}_location": loc, f"{key}_scale": sc})
        rows.append(row)
    return pd.DataFrame(rows)
