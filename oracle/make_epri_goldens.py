"""Golden fits for stryke_b200.epri: run the UNMODIFIED reference class ``Stryke.stryke.epri`` (stryke.py:4322-5084) on its
own database for a set of queries and record what it computed, together with the filtered rates so the fit functions
can be checked without the database.

Run in the build container only:  python -m oracle.make_epri_goldens   ->  tests/golden/epri_fits.json
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np

from . import harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "epri_fits.json")

QUERIES = [
    dict(Genus="Micropterus", Month=[3, 4, 5], HUC02=[4]),
    dict(Genus="Micropterus", Month=[6, 7, 8]),
    dict(Genus="Ambloplites", Month=[12, 1, 2], HUC02=[4]),
    dict(Genus="Lepomis", Month=[6, 7, 8], HUC02=[5]),
    dict(Family="Catostomidae", Month=[3, 4, 5]),
    dict(Genus="Alosa"),
    dict(Species="Perca flavescens", Month=[9, 10, 11]),
    dict(Family="Ictaluridae", states=["WI", "MI"]),
    dict(Genus="Notropis", plant_cap=(1000, ">")),
    dict(Genus="Notropis", plant_cap=(1000, "<=")),
    dict(Genus="Sander", HUC04=[405, 707]),
    dict(Family="Salmonidae", Month=[1, 2, 3, 4, 5, 6]),
]


def main():
    import pandas as pd
    pd.set_option("future.infer_string", False)     # the reference's fillna(0) over string columns needs object dtype
    S = harness.load_reference()
    cases = []
    for q in QUERIES:
        kw = {k: (tuple(v) if k == "plant_cap" else v) for k, v in q.items()}
        with contextlib.redirect_stdout(io.StringIO()):
            e = S.epri(**kw)
            e.ParetoFit(); e.LogNormalFit(); e.WeibullMinFit()
            extreme = None
            try:
                e.ExtremeFit()
                extreme = [float(v) for v in e.dist_extreme]
            except Exception:
                pass
        cohorts = [int(np.int32(e.epri[c].sum())) for c in ("0_5", "5_10", "10_15", "15_20", "20_25", "25_38", "38_51", "51_64", "64_76", "GT76")]
        cases.append(dict(query=q, presence=e.presence, sample_size=int(e.sample_size), max_ent_rate=float(e.max_ent_rate),
                          rates=[float(v) for v in e.epri.FishPerMft3.values],
                          pareto=[float(v) for v in e.dist_pareto], lognorm=[float(v) for v in e.dist_lognorm],
                          weibull=[float(v) for v in e.dist_weibull], extreme=extreme, cohorts=cohorts))
        print(q, e.sample_size, e.presence, e.dist_pareto, e.dist_lognorm, e.dist_weibull)
    import scipy
    with open(OUT, "w") as fh:
        json.dump(dict(scipy=scipy.__version__, reference="Stryke.stryke.epri on Data/epri1997.csv (unmodified)", cases=cases), fh)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
