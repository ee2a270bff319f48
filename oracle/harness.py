"""Reference harness: runs the UNMODIFIED reference (``/root/reference/Stryke/stryke.py``) under import stubs
and an in-memory HDF store, and records every random draw tagged by (cell, move, fish, slot).

TEST INFRASTRUCTURE — only usable in the build container (``/root/reference`` does not exist on the GPU box).
It is used by ``oracle/make_goldens.py`` to generate the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent) to pin ``oracle/stryke_oracle.py``.
Recipe: SURVEY.md Appendix D.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

REFERENCE_ROOT = os.environ.get("STRYKE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Stryke", "stryke.py"))


class FakeStore:
    """Stands in for ``pd.HDFStore`` (PyTables is not installed here)."""
    _S: dict = {}

    def __init__(self, path, mode="a", **k):
        if mode == "w" or str(path) not in self._S:
            self._S[str(path)] = {}
        self.d = self._S[str(path)]

    def __setitem__(self, k, v):
        self.d["/" + k.lstrip("/")] = v

    def __getitem__(self, k):
        v = self.d["/" + k.lstrip("/")]
        v = pd.concat(v, ignore_index=True) if isinstance(v, list) else v
        if k.strip("/") == "Daily":  # pandas>=3 shim for stryke.py:3564
            v = v.copy()
            v[v.columns[6:]] = v[v.columns[6:]].astype(float)
        return v

    def __contains__(self, k):
        return "/" + k.lstrip("/") in self.d

    def get(self, k):
        return self[k]

    def keys(self):
        return list(self.d)

    def put(self, k, v, **kw):
        self[k] = v

    def flush(self, *a, **k):
        pass

    def close(self):  # also neutralises the close-inside-scenario-loop defect (stryke.py:3516, Appendix E1)
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_S = None


def load_reference():
    """Import ``Stryke.stryke`` from the reference tree with the five import-time stubs."""
    global _S
    if _S is not None:
        return _S
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for m in ["matplotlib", "matplotlib.pyplot", "xlrd", "hydrofunctions", "statsmodels", "statsmodels.api", "h5py"]:
        if m not in sys.modules:
            mod = MagicMock(name=m)
            mod.__path__ = []
            sys.modules[m] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    pd.HDFStore = FakeStore
    pd.DataFrame.to_hdf = lambda df, store, key=None, **kw: store.d.setdefault("/" + key.lstrip("/"), []).append(df.copy())
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        import Stryke.stryke as S
    _S = S
    return S


class RecGen(np.random.Generator):
    """PCG64 ``Generator`` that logs what the reference's module-level ``rng`` hands out (stryke.py:70)."""

    def __init__(self, seed):
        super().__init__(np.random.PCG64(seed))
        self.log = []

    def standard_normal(self, size=None, *a, **k):
        v = super().standard_normal(size, *a, **k)
        self.log.append(("standard_normal", np.array(v, dtype=float).ravel().copy()))
        return v

    def uniform(self, low=0.0, high=1.0, size=None):
        v = super().uniform(low, high, size)
        self.log.append(("uniform", float(np.ravel(low)[0]), float(np.ravel(high)[0]), np.array(v, dtype=float).ravel().copy()))
        return v

    def random(self, size=None, *a, **k):
        v = super().random(size, *a, **k)
        self.log.append(("random", np.array(v, dtype=float).ravel().copy()))
        return v


class Recorder:
    """Wraps the global ``np.random`` entry points the reference calls, and ``node_surv_rate`` / ``movement`` /
    ``daily_hours`` on one simulation object, so every draw lands in a (move, fish, slot) slot.

    Slots per (move k, evaluation e) where e = 0 is ``np.vectorize``'s probe call on fish 0 (the "ghost",
    SURVEY.md Appendix E3) and e = f + 1 is fish f:
      rR      raw uniform behind ``np.random.uniform(0.3, 1.0, 1)``            (stryke.py:900)
      depth   raw uniform behind ``np.random.uniform(depth_1, depth_2, 1)``    (stryke.py:1494)
      cause   up to 4 values of ``np.random.random()``                         (stryke.py:1545-1560)
      route   the uniform consumed by ``np.random.choice(locs, p=probs)``      (stryke.py:2058)
    plus ``dice[k, f]`` (stryke.py:3189), the presence uniform (stryke.py:3026) and the PCG64 draws.
    """

    def __init__(self, S, sim, seed):
        self.S, self.sim, self.seed = S, sim, seed
        self.cells = []
        self.cur = None
        self.routing = {}  # (day_index, location) -> (locs, probs)
        self._ctx = None

    # ---- cell bookkeeping -------------------------------------------------------------------------------------
    def _new_cell(self, Q_dict, scenario, hours):
        self.cur = dict(scenario=str(scenario), curr_Q=float(Q_dict["curr_Q"]), tot_hours=float(np.sum(hours[0])),
                        hours={str(k): float(v) for k, v in hours[2].items()}, presence=None, pcg=[], moves=[],
                        pcg_mark=len(self.S.rng.log))
        self.cells.append(self.cur)

    def _move(self):
        return self.cur["moves"][-1]

    # ---- wrappers ---------------------------------------------------------------------------------------------
    def install(self):
        S, sim, rec = self.S, self.sim, self
        self._orig = dict(uniform=np.random.uniform, random=np.random.random, choice=np.random.choice, normal=np.random.normal)
        o_normal = np.random.normal

        def normal(loc=0.0, scale=1.0, size=None):
            # fish lengths from Length_mean / Length_sd (stryke.py:3090): the legacy global stream, not the PCG64 one
            v = o_normal(loc, scale, size)
            if rec.cur is not None and size is not None and np.size(v) > 1:
                rec.cur["legacy_normal"] = (float(loc), float(scale), np.array(v, dtype=float).ravel().copy())
            return v
        o_uniform, o_random = np.random.uniform, np.random.random
        o_sample = np.random.random_sample

        def uniform(low=0.0, high=1.0, size=None):
            if rec.cur is not None and rec._ctx is not None and rec._ctx[0] == "surv" and size == 1:
                # legacy RandomState.uniform is ``low + (high - low) * random_sample()`` (verified bit-exact in
                # tests/test_oracle_vs_reference.py); record the raw uniform so it can be injected on the device
                u = o_sample(1)
                v = low + (high - low) * u
                e = rec._ctx[1]
                if float(low) == 0.3 and float(high) == 1.0:
                    rec._move()["rR"][e] = float(u[0])
                else:
                    rec._move()["depth"][e] = float(u[0])
                return v
            v = o_uniform(low, high, size)
            if rec.cur is None:
                return v
            if size is None:  # presence roll, stryke.py:3026
                rec.cur["presence"] = float(v)
            elif np.ndim(v) == 1 and rec._ctx is None and not (np.size(v) == 1 and rec._in_hours):
                # dice vector, stryke.py:3189 -> a new move begins
                rec.cur["moves"].append(dict(dice=np.array(v, dtype=float), n=int(np.size(v)), surv_calls=0,
                                             move_calls=0, rR={}, depth={}, cause={}, route={}))
            return v

        def random(size=None):
            v = o_random(size)
            if rec.cur is not None and rec._ctx is not None and rec._ctx[0] == "surv":
                rec._move()["cause"].setdefault(rec._ctx[1], []).append(float(v))
            return v

        def choice(a, size=None, replace=True, p=None):
            # legacy RandomState.choice with p: cdf = cumsum(p); cdf /= cdf[-1]; u = random_sample();
            # idx = searchsorted(cdf, u, side='right')   (numpy/random/mtrand.pyx)
            if p is None or size is not None:
                return rec._orig["choice"](a, size=size, replace=replace, p=p)
            p = np.asarray(p, dtype=float)
            if abs(p.sum() - 1.0) > np.sqrt(np.finfo(np.float64).eps):
                raise ValueError("probabilities do not sum to 1")
            u = float(o_sample())
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            idx = int(np.searchsorted(cdf, u, side="right"))
            if rec.cur is not None and rec._ctx is not None and rec._ctx[0] == "move":
                rec._move()["route"][rec._ctx[1]] = u
                rec.routing[(len(rec.cells) - 1, rec._ctx[2])] = (list(map(str, a)), p.copy())
            return np.asarray(a)[idx]

        np.random.uniform, np.random.random, np.random.choice, np.random.normal = uniform, random, choice, normal

        cls = type(sim)
        o_surv, o_move, o_hours = cls.node_surv_rate, cls.movement, cls.daily_hours
        rec._in_hours = False

        def node_surv_rate(*a, **k):
            m = rec._move()
            rec._ctx = ("surv", m["surv_calls"])
            m["surv_calls"] += 1
            try:
                return o_surv(sim, *a, **k)
            finally:
                rec._ctx = None

        def movement(location, *a, **k):
            m = rec._move()
            rec._ctx = ("move", m["move_calls"], str(location))
            m["move_calls"] += 1
            try:
                return o_move(sim, location, *a, **k)
            finally:
                rec._ctx = None

        def daily_hours(Q_dict, scenario, *a, **k):
            rec._in_hours = True
            rec.cur = rec.cur  # draws inside daily_hours belong to no move
            try:
                out = o_hours(sim, Q_dict, scenario, *a, **k)
            finally:
                rec._in_hours = False
            rec._new_cell(Q_dict, scenario, out)
            return out

        sim.node_surv_rate, sim.movement, sim.daily_hours = node_surv_rate, movement, daily_hours

    def uninstall(self):
        np.random.uniform, np.random.random, np.random.choice, np.random.normal = (
            self._orig[k] for k in ("uniform", "random", "choice", "normal"))
        for k in ("node_surv_rate", "movement", "daily_hours"):
            self.sim.__dict__.pop(k, None)

    def finish(self):
        log = self.S.rng.log
        for i, c in enumerate(self.cells):
            hi = self.cells[i + 1]["pcg_mark"] if i + 1 < len(self.cells) else len(log)
            c["pcg"] = log[c["pcg_mark"]:hi]


def run_reference(prj, seed=1, traces=True, summary=False, quiet=True):
    """Run the reference's ``run()`` on a project dict.  Returns (sim, store, recorder)."""
    import contextlib
    import io
    from . import fixtures
    S = load_reference()
    sim = S.simulation.__new__(S.simulation)
    fixtures.apply_to(sim, prj)
    FakeStore._S.pop(sim.hdf_path, None)
    S.STORE_AGENT_TRACES = bool(traces)
    S.STORE_SIMULATION_TABLE = bool(traces)
    np.random.seed(seed)
    S.rng = RecGen(seed)
    rec = Recorder(S, sim, seed)
    rec.install()
    try:
        with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
            sim.run()
    finally:
        rec.uninstall()
    rec.finish()
    store = FakeStore(sim.hdf_path)
    if summary:
        for key, df in [("Flow Scenarios", sim.flow_scenarios_df), ("Operating Scenarios", sim.operating_scenarios_df),
                        ("Population", sim.pop), ("Nodes", sim.nodes), ("Edges", sim.edges),
                        ("Unit_Parameters", sim.unit_params), ("Facilities", sim.facility_params),
                        ("Hydrograph", sim.input_hydrograph_df)]:
            store[key] = df
        with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
            sim.summary()
    return sim, store, rec
