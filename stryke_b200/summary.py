"""``simulation.summary()`` — reference: Stryke/stryke.py:3520-4101 (host-side, runs once per project).

Works from the ``Daily`` and ``State_Daily`` tables ``run()`` wrote (the ``simulations/{scen}/{spc}`` per-fish table
is optional in the reference and not produced here): whole-project and per-(move, state) survival probabilities
with empirical mean / std and bootstrap 95 % intervals, the yearly entrainment / mortality summary with 1-in-N day
quantiles, and the unit-level driver diagnostics.  The bootstrap is vectorised: the reference's
``n_bootstrap`` sequential ``np.random.choice(data, size=len(data))`` calls consume the global MT19937 stream exactly
like one ``np.random.randint(0, n, size=(n_bootstrap, n))`` call, so under the same ``np.random.seed`` the intervals
are identical (``bootstrap_mean_ci``, stryke.py:143-162).
"""
from __future__ import annotations

import logging
import os

import numpy as np
import pandas as pd

from .store import open_store

logger = logging.getLogger(__name__)

BETA_CAPTION = (
    "Means shown are empirical averages of per-iteration, per-day survival probabilities. "
    "95% intervals are bootstrap estimates. Direct beta distribution fits are sensitive to values exactly 0 or 1 "
    "and are not used for the reported mean.")


def summarize_ci(series):
    """stryke.py:136-140."""
    return pd.Series({"mean": series.mean(), "lower_95_CI": np.percentile(series, 2.5),
                      "upper_95_CI": np.percentile(series, 97.5)})


#: CUDA ordinal the large bootstraps run on (set by summarize(); None = NumPy only)
_DEVICE = None
#: resampled values (len(data) * n_bootstrap) from which the device kernel is used.  Below it NumPy is faster than an
#: upload + launch + download and reproduces the reference's MT19937 stream draw for draw.
DEVICE_BOOTSTRAP_MIN = 1 << 23


def _device_for(n, n_bootstrap):
    return _DEVICE if (_DEVICE is not None and n * n_bootstrap >= DEVICE_BOOTSTRAP_MIN) else None


def bootstrap_mean_ci(data, n_bootstrap=10000, ci=95, chunk_elems=1 << 24, result_used=True):
    """Mean and bootstrap interval of ``data`` (stryke.py:143-162).

    Small problems: NumPy, resamples drawn ``chunk`` rows at a time from the global MT19937 stream exactly as the
    reference's loop consumes it.  Large problems (>= DEVICE_BOOTSTRAP_MIN resampled values, a GPU present): the
    resampled means come from ``stryke_b200_bootstrap_means`` (Philox indices, seeded from ``np.random`` so that
    ``np.random.seed`` still makes a run reproducible); a call whose result the caller discards
    (``result_used=False``: the two draws at stryke.py:3832-3833) is then skipped altogether."""
    data = np.array(data)
    n = len(data)
    dev = _device_for(n, n_bootstrap)
    if dev is not None:
        if not result_used:
            return np.mean(data), np.nan, np.nan
        from . import engine
        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        means = engine.bootstrap_means(data.astype(np.float64), n_bootstrap, seed=seed, device=dev)
    else:
        means = np.empty(n_bootstrap)
        rows = max(1, int(chunk_elems // max(n, 1)))
        for lo in range(0, n_bootstrap, rows):
            hi = min(n_bootstrap, lo + rows)
            idx = np.random.randint(0, n, size=(hi - lo, n))
            means[lo:hi] = data[idx].mean(axis=1)
    lower = np.percentile(means, (100 - ci) / 2)
    upper = np.percentile(means, 100 - (100 - ci) / 2)
    return np.mean(data), lower, upper


def _mean_std_ci(prob_values, n_bootstrap=2000):
    s = pd.Series(prob_values)
    mean, std = s.mean(), s.std()
    if len(prob_values) < 2:
        return mean, std, mean, mean
    _, lcl, ucl = bootstrap_mean_ci(prob_values, n_bootstrap=n_bootstrap, ci=95)
    return mean, std, lcl, ucl


def _label_codes(col):
    """(codes, labels) of a label column: ``labels[codes]`` reproduces it.  Long frames are compared by these integer codes
    instead of string by string (3e6-row string compares per (species, scenario) were most of summary()'s time)."""
    if isinstance(col.dtype, pd.CategoricalDtype):
        return np.asarray(col.cat.codes), np.array([str(x) for x in col.cat.categories], dtype=object)
    codes, uniq = pd.factorize(col, sort=False)
    return np.asarray(codes), np.array([str(x) for x in uniq], dtype=object)


def _stripped(col):
    """``col.astype(str).str.strip()`` computed on the distinct labels only; also returns (codes, stripped labels)."""
    codes, labels = _label_codes(col)
    stripped = np.array([x.strip() for x in labels], dtype=object)
    return pd.Series(stripped[codes], index=col.index, dtype=object), codes, stripped


def _code_mask(codes, labels, value):
    """Rows whose (stripped) label equals ``value``: the labels are few, the rows are compared as integers."""
    hit = np.flatnonzero(labels == value)
    if hit.size == 0:
        return np.zeros(codes.shape[0], bool)
    return np.isin(codes, hit) if hit.size > 1 else codes == hit[0]


def _gpu_ordinal(sim):
    """The simulation's CUDA device when the library is built and a GPU is visible, else None."""
    try:
        from . import _abi
        if _abi.lib().stryke_b200_device_count() > 0:
            return int(getattr(sim, "device", 0))
    except Exception:      # summary() of an existing store also works on a machine without the library / a GPU
        pass
    return None


def summarize(sim):
    global _DEVICE
    _DEVICE = _gpu_ordinal(sim)
    try:
        return _summarize(sim)
    finally:
        _DEVICE = None


def _summarize(sim):
    if not hasattr(sim, "wks_dir"):
        sim.wks_dir = None
    hdf_path = os.path.join(sim.proj_dir, "%s.h5" % sim.output_name)
    with open_store(hdf_path, mode="r") as store:
        sim.beta_dict = {}
        species = store["Population"].Species.unique()
        scens = store["Flow Scenarios"].Scenario.unique()
        sim.daily_summary = store.get("Daily", categorical=True)
        num_cols = list(sim.daily_summary.columns[6:])
        sim.daily_summary[num_cols] = sim.daily_summary[num_cols].astype(float)      # stryke.py:3564
        sim.daily_summary["species_norm"], d_spc, d_spc_l = _stripped(sim.daily_summary["species"])
        sim.daily_summary["scenario_norm"], d_scn, d_scn_l = _stripped(sim.daily_summary["scenario"])
        daily_index = _DailyIndex(sim.daily_summary)      # (label codes taken while the columns are still categoricals)
        for c in ("species", "scenario"):                 # the frame the caller sees holds plain strings, as the reference's
            if isinstance(sim.daily_summary[c].dtype, pd.CategoricalDtype):
                sim.daily_summary[c] = sim.daily_summary[c].astype(object)
        print("\n" + "=" * 60, flush=True)
        print("SIMULATION SUMMARY STATISTICS", flush=True)
        print("=" * 60, flush=True)
        state_daily = store.get("State_Daily", categorical=True) if "State_Daily" in store else None
        if state_daily is not None:
            state_daily["day"] = pd.to_datetime(state_daily["day"])
            state_daily["move"] = pd.to_numeric(state_daily["move"], errors="coerce").fillna(0).astype(int)
            for c in ("successes", "count"):
                state_daily[c] = pd.to_numeric(state_daily[c], errors="coerce").fillna(0.0)
            state_daily["scenario"], s_scn, s_scn_l = _stripped(state_daily["scenario"])
            state_daily["species"], s_spc, s_spc_l = _stripped(state_daily["species"])
            st_codes, st_labels = _label_codes(state_daily["state"])
            state_daily["state"] = pd.Series(st_labels[st_codes], index=state_daily.index, dtype=object)
            # rows of one (scenario, species) made contiguous once (stable: the order inside a group is kept)
            s_group = s_scn.astype(np.int64) * max(len(s_spc_l), 1) + s_spc
            s_order = np.argsort(s_group, kind="stable")
            state_sorted = state_daily.iloc[s_order]
            s_group_sorted = s_group[s_order]
            st_codes_sorted = st_codes[s_order]
        moves = [int(m) for m in sim.moves]
        final_move = max(moves) if moves else 0
        group_sizes = sim.daily_summary.groupby(["species", "scenario"], sort=False).size() if len(species) > 1 else None
        for si, spc in enumerate(species):
            if si > 0:
                # the reference rebuilds the yearly summary inside its species loop (stryke.py:3770-3862 sit under
                # `for i in species`): the table of the last pass is the one that is kept, but every pass draws its
                # bootstraps from the global stream — replayed here so that a seeded run stays draw-for-draw identical
                yearly_summary(sim.daily_summary, draws_only=True, sizes=group_sizes, index=daily_index)
            for scen in scens:
                if state_daily is None:
                    logger.warning("Missing simulation data for key 'simulations/%s/%s'", scen, spc)
                    continue
                hs, hp = np.flatnonzero(s_scn_l == str(scen).strip()), np.flatnonzero(s_spc_l == str(spc).strip())
                if hs.size == 1 and hp.size == 1:
                    gkey = int(hs[0]) * max(len(s_spc_l), 1) + int(hp[0])
                    lo_, hi_ = np.searchsorted(s_group_sorted, [gkey, gkey + 1])
                    state_dat = state_sorted.iloc[lo_:hi_]
                    state_codes = (st_codes_sorted[lo_:hi_], st_labels)
                else:
                    state_codes = None
                    state_dat = state_daily[_code_mask(s_scn, s_scn_l, str(scen).strip()) & _code_mask(s_spc, s_spc_l, str(spc).strip())]
                if state_dat.empty:
                    logger.warning("Missing simulation aggregates for scenario '%s' species '%s'", scen, spc)
                    continue
                # whole-project survival per (iteration, day): successes at the final move / pop_size (:3616-3637)
                sub = sim.daily_summary[_code_mask(d_spc, d_spc_l, str(spc).strip())
                                        & _code_mask(d_scn, d_scn_l, str(scen).strip())][["iteration", "day", "pop_size"]].copy()
                sub["day"] = pd.to_datetime(sub["day"])
                sub = sub.groupby(["iteration", "day"], as_index=False)["pop_size"].max()
                fin = state_dat[state_dat["move"] == final_move]
                if fin.empty:
                    fin = pd.DataFrame(columns=["iteration", "day", "successes"])
                else:
                    fin = fin.groupby(["iteration", "day"], as_index=False)["successes"].sum()
                whole = sub.merge(fin, on=["iteration", "day"], how="left")
                whole["successes"] = whole["successes"].fillna(0.0)
                whole["count"] = whole["pop_size"].astype(float)
                whole["prob"] = np.where(whole["count"] > 0, whole["successes"] / whole["count"], 0.0)
                probs = whole["prob"].values
                if probs.size == 0:
                    raise ValueError(f"No survival probability data for scenario '{scen}' species '{spc}'.")
                sim.beta_caption = BETA_CAPTION
                sim.beta_dict["%s_%s_%s" % (scen, spc, "whole")] = [scen, spc, "whole", *_mean_std_ci(probs)]
                # per (move, state): one pass over the group's rows; keys in the reference's order — moves ascending, the
                # states of a move in order of first appearance (`for l in moves: for m in rs["state"].unique()`, :3688-3749)
                cnt_, suc_ = state_dat["count"].to_numpy(dtype=float), state_dat["successes"].to_numpy(dtype=float)
                prob_ = np.where(cnt_ > 0, suc_ / np.where(cnt_ > 0, cnt_, 1.0), 0.0)
                mv_ = state_dat["move"].to_numpy()
                stc_, stl_ = state_codes if state_codes is not None else _label_codes(state_dat["state"])
                key_ = mv_.astype(np.int64) * max(len(stl_), 1) + stc_
                uk, first, inv = np.unique(key_, return_index=True, return_inverse=True)
                order_ = np.argsort(np.asarray(inv), kind="stable")
                bounds = np.concatenate([[0], np.cumsum(np.bincount(np.asarray(inv), minlength=len(uk)))])
                by_move = {}
                for j in np.argsort(first, kind="stable"):          # (move, state) keys by first appearance
                    by_move.setdefault(int(uk[j] // max(len(stl_), 1)), []).append(j)
                for l in moves:
                    for j in by_move.get(int(l), []):
                        m = stl_[int(uk[j] % max(len(stl_), 1))]
                        pv = prob_[order_[bounds[j]:bounds[j + 1]]]
                        if pv.size == 0:
                            continue
                        sim.beta_dict["%s_%s_%s" % (scen, spc, m)] = [scen, spc, m, *_mean_std_ci(pv)]
        sim.beta_df = pd.DataFrame.from_dict(sim.beta_dict, orient="index",
                                             columns=["scenario number", "species", "state", "survival rate", "variance", "ll", "ul"])
        st = sim.beta_df["state"].astype(str)
        sim.beta_df_units_only = sim.beta_df[(st != "whole") & (~st.str.contains("river_node", case=False, na=False))
                                             & (~st.str.contains("^river$", case=False, na=False, regex=True))][
            ["state", "survival rate", "variance", "ll", "ul"]].copy()
        sim.beta_df_units_only.rename(columns={"state": "Passage Route", "survival rate": "Mean", "variance": "Variance",
                                               "ll": "Lower 95% CI", "ul": "Upper 95% CI"}, inplace=True)
        sim.cum_sum = yearly_summary(sim.daily_summary, index=daily_index)
        if sim.wks_dir and str(sim.wks_dir).endswith((".xlsx", ".xls")):
            try:
                from . import xlsx
                xlsx.append_sheets(sim.wks_dir, {"beta fit": sim.beta_df, "daily summary": sim.daily_summary,
                                                 "yearly summary": sim.cum_sum})
            except (PermissionError, FileNotFoundError, OSError, ValueError) as exc:
                raise RuntimeError(f"Failed to write Excel summaries to {sim.wks_dir}: {exc}") from exc
        else:
            logger.info("Web app mode detected - Excel export skipped (results in HDF5 only)")
        sim.driver_diagnostics = driver_diagnostics(
            store["Unit_Parameters"] if "Unit_Parameters" in store else None,
            store["Route_Flows"] if "Route_Flows" in store else None,
            store["Unit_Hours"] if "Unit_Hours" in store else None, sim.beta_df_units_only)
        print("=" * 60 + "\n", flush=True)
    with open_store(hdf_path, mode="a") as store:
        kw = dict(format="table", data_columns=True)
        store.put("Daily_Summary", sim.daily_summary, **kw)
        store.put("Beta_Distributions", sim.beta_df, **kw)
        store.put("Beta_Distributions_Units", sim.beta_df_units_only, **kw)
        store.put("Yearly_Summary", sim.cum_sum, **kw)
        if isinstance(getattr(sim, "driver_diagnostics", None), pd.DataFrame):
            store.put("Driver_Diagnostics", sim.driver_diagnostics, **kw)
    return None


class _DailyIndex:
    """The Daily frame's rows grouped by (species, scenario) once: label codes, a stable order that makes every group
    contiguous, and the numeric columns as arrays in that order.  ``yearly_summary`` runs once per species of the project
    (the reference rebuilds it inside its species loop) — factorising 1e6-row label columns and masking the frame every
    time was two thirds of summary()'s time."""

    def __init__(self, daily):
        self.spc, self.spc_l = _label_codes(daily["species"])
        self.scn, self.scn_l = _label_codes(daily["scenario"])
        self.nscn = max(len(self.scn_l), 1)
        key = self.spc.astype(np.int64) * self.nscn + self.scn
        self.order = np.argsort(key, kind="stable")
        self.key_sorted = key[self.order]
        self.cols = {c: daily[c].to_numpy()[self.order] for c in ("iteration", "num_entrained", "num_survived", "pop_size")}
        self.first_seen = {}                      # labels in order of first appearance (what .unique() returns)
        for name, codes, labels in (("species", self.spc, self.spc_l), ("scenario", self.scn, self.scn_l)):
            _, first = np.unique(codes, return_index=True)
            self.first_seen[name] = [labels[codes[i]] for i in np.sort(first)]

    def group(self, species, scenario):
        """Row range [lo, hi) (in the sorted order) of one (species, scenario); empty when a label does not occur."""
        hs, hc = np.flatnonzero(self.spc_l == str(species)), np.flatnonzero(self.scn_l == str(scenario))
        if hs.size != 1 or hc.size != 1:
            return 0, 0
        k = int(hs[0]) * self.nscn + int(hc[0])
        lo, hi = np.searchsorted(self.key_sorted, [k, k + 1])
        return int(lo), int(hi)


def yearly_summary(daily, draws_only=False, sizes=None, index=None):
    """Per (species, scenario): entrainment probability, yearly totals with 95 % percentiles over iterations, bootstrap
    means and 1-in-10/100/1000-day quantiles of daily entrainment and mortality (stryke.py:3770-3862).
    ``draws_only``: consume the random draws of one pass and return nothing (see summarize)."""
    cols = ["species", "scenario", "prob_entrainment", "mean_yearly_entrainment", "lcl_yearly_entrainment",
            "ucl_yearly_entrainment", "1_in_10_day_entrainment", "1_in_100_day_entrainment", "1_in_1000_day_entrainment",
            "mean_yearly_mortality", "lcl_yearly_mortality", "ucl_yearly_mortality", "1_in_10_day_mortality",
            "1_in_100_day_mortality", "1_in_1000_day_mortality"]
    out = {c: [] for c in cols}
    if draws_only and sizes is None:
        sizes = daily.groupby(["species", "scenario"], sort=False).size()
    if draws_only and all(_device_for(int(n), 10000) is not None for n in sizes.values):
        return None                           # every group on the device path: the discarded bootstraps draw nothing
    D = index if index is not None else _DailyIndex(daily)
    # the reference's group order: the keys of groupby(["species", "scenario", "iteration"]) (sorted), .unique() of each
    species_order, scenario_order = sorted(D.first_seen["species"]), sorted(D.first_seen["scenario"])
    for fishy in species_order:
        for scen in scenario_order:
            lo, hi = D.group(fishy, scen)
            if draws_only and _device_for(hi - lo, 10000) is not None:
                continue                      # large group on the device path: its discarded bootstraps draw nothing
            out["species"].append(fishy)
            out["scenario"].append(scen)
            if hi == lo:
                for c in cols[2:]:
                    out[c].append(0.0)
                continue
            ent = D.cols["num_entrained"][lo:hi].astype(float)
            mort = ent - D.cols["num_survived"][lo:hi].astype(float)
            if draws_only:
                bootstrap_mean_ci(ent, result_used=False)
                bootstrap_mean_ci(mort, result_used=False)
                continue
            it_codes, it_labels = pd.factorize(D.cols["iteration"][lo:hi], sort=True)
            n_it = len(it_labels)
            pop_it = np.bincount(it_codes, weights=D.cols["pop_size"][lo:hi].astype(float), minlength=n_it)
            ent_it = np.bincount(it_codes, weights=ent, minlength=n_it)
            mort_it = np.bincount(it_codes, weights=mort, minlength=n_it)
            probs = np.where(pop_it > 0, ent_it / np.where(pop_it > 0, pop_it, 1.0), 0.0)
            out["prob_entrainment"].append(float(np.mean(probs)) if len(probs) else 0.0)
            # drawn by the reference (:3832-3833) although nothing reads the results: kept for stream identity on the
            # NumPy path, skipped on the device path
            bootstrap_mean_ci(ent, result_used=False)
            bootstrap_mean_ci(mort, result_used=False)
            for what, per_it, per_day in (("entrainment", ent_it, ent), ("mortality", mort_it, mort)):
                ci = summarize_ci(pd.Series(per_it))
                out[f"mean_yearly_{what}"].append(ci["mean"])
                out[f"lcl_yearly_{what}"].append(ci["lower_95_CI"])
                out[f"ucl_yearly_{what}"].append(ci["upper_95_CI"])
                for T in (10, 100, 1000):
                    out[f"1_in_{T}_day_{what}"].append(float(np.quantile(per_day, 1 - 1 / T)))
    return None if draws_only else pd.DataFrame.from_dict(out, orient="columns")


def driver_diagnostics(unit_params, route_flows, unit_hours, beta_units):
    """Unit-level diagnostics table (stryke.py:3913-4084)."""
    if not isinstance(unit_params, pd.DataFrame) or unit_params.empty:
        logger.info("Driver diagnostics skipped: Unit_Parameters missing or empty.")
        return None
    diag = unit_params.copy()
    diag["route"] = diag.index.astype(str)
    pref = ["H", "RPM", "D", "N", "Qopt", "Qcap", "D1", "D2", "B", "ada", "intake_vel", "ps_D", "ps_length", "fb_depth",
            "submergence_depth", "elevation_head"]
    diag = diag[["route"] + [c for c in pref if c in diag.columns]].reset_index(drop=True)
    if isinstance(beta_units, pd.DataFrame) and not beta_units.empty:
        bu = beta_units.rename(columns={"Passage Route": "route", "state": "route", "Mean": "survival_mean",
                                        "survival rate": "survival_mean", "Variance": "survival_variance",
                                        "variance": "survival_variance", "Lower 95% CI": "survival_lcl", "ll": "survival_lcl",
                                        "Upper 95% CI": "survival_ucl", "ul": "survival_ucl"})
        keep = [c for c in ("route", "survival_mean", "survival_variance", "survival_lcl", "survival_ucl") if c in bu.columns]
        if "route" in keep and len(keep) > 1:
            bu = bu[keep].groupby("route", as_index=False)[keep[1:]].mean()
            diag = diag.merge(bu, on="route", how="left")
    have_rf = isinstance(route_flows, pd.DataFrame) and not route_flows.empty and {"route", "discharge_cfs"} <= set(route_flows.columns)
    if have_rf:
        rf = route_flows.copy()
        rf["route"] = rf["route"].astype(str)
        mean_q = rf.groupby("route")["discharge_cfs"].mean()
        diag = diag.merge(mean_q.rename("mean_discharge_cfs"), left_on="route", right_index=True, how="left")
        unit_mean = mean_q[mean_q.index.isin(set(diag["route"].dropna().astype(str)))]
        share = None
        if float(unit_mean.sum()) > 0:
            share = unit_mean / float(unit_mean.sum())
        elif float(mean_q.sum()) > 0:
            share = mean_q / float(mean_q.sum())
        if share is not None:
            diag = diag.merge(share.rename("flow_share"), left_on="route", right_index=True, how="left")
    if "flow_share" in diag.columns and "survival_mean" in diag.columns:
        diag["mortality_weight"] = diag["flow_share"] * (1 - diag["survival_mean"])
    dispatch = None
    if isinstance(unit_hours, pd.DataFrame) and not unit_hours.empty and "Qopt" in diag.columns:
        uh = unit_hours.copy()
        if {"route", "hours"} <= set(uh.columns):
            uh["route"] = uh["route"].astype(str)
            uh["hours"] = pd.to_numeric(uh["hours"], errors="coerce")
            if "discharge_cfs" in uh.columns:
                uh["discharge_cfs"] = pd.to_numeric(uh["discharge_cfs"], errors="coerce")
                dispatch = uh
            elif have_rf:
                dispatch = rf.merge(uh, on=["scenario", "iteration", "day", "route"], how="outer")
                dispatch["hours"] = pd.to_numeric(dispatch["hours"], errors="coerce")
                dispatch["discharge_cfs"] = pd.to_numeric(dispatch["discharge_cfs"], errors="coerce")
    if isinstance(dispatch, pd.DataFrame) and not dispatch.empty:
        dispatch = dispatch.merge(diag[["route", "Qopt"]], on="route", how="left")
        dispatch["Qopt"] = pd.to_numeric(dispatch["Qopt"], errors="coerce")
        dispatch = dispatch[(dispatch["hours"] > 0.0) & (dispatch["Qopt"] > 0.0) & dispatch["discharge_cfs"].notna()].copy()
        if not dispatch.empty:
            off = (dispatch["discharge_cfs"] - dispatch["Qopt"]).abs() / dispatch["Qopt"] * 100.0
            outside = (dispatch["discharge_cfs"] < dispatch["Qopt"] * 0.9) | (dispatch["discharge_cfs"] > dispatch["Qopt"] * 1.1)
            dispatch["hours_outside_band"] = dispatch["hours"] * outside.astype(float)
            dispatch["abs_pct_off_hours"] = off * dispatch["hours"]
            hs = dispatch.groupby("route").agg(total_hours=("hours", "sum"), hours_outside_band=("hours_outside_band", "sum"),
                                               abs_pct_off_hours=("abs_pct_off_hours", "sum"))
            hs["mean_abs_pct_off_qopt"] = hs["abs_pct_off_hours"] / hs["total_hours"]
            hs["pct_hours_outside_qopt_band"] = hs["hours_outside_band"] / hs["total_hours"] * 100.0
            diag = diag.merge(hs[["total_hours", "mean_abs_pct_off_qopt", "pct_hours_outside_qopt_band"]],
                              left_on="route", right_index=True, how="left")
    return diag
