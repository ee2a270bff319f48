"""Project result store with the reference's HDF keys.

The reference writes ``<proj_dir>/<output_name>.h5`` through ``pandas.HDFStore`` (PyTables).  When PyTables is
importable this module uses exactly that (same keys, table format, ``min_itemsize``).  PyTables / libhdf5 are absent
from this image, so otherwise a same-keys container is written next to where the ``.h5`` would be:
``<output_name>.h5.store/`` holding, per key, a directory of Parquet parts (``pyarrow``; one part per ``put`` / ``append``,
so appending never rewrites what is already there) plus ``manifest.json`` — a non-executable, documented format that
anything reading Arrow / Parquet opens, and a loud log line says that no ``.h5`` was produced.  Nothing here unpickles:
opening a project directory somebody else wrote cannot run code.  ``python -m stryke_b200.store <project>.h5 <target.h5>``
converts such a store to a real ``HDFStore`` file where PyTables exists (the web app's host).  The API below is the subset
of ``HDFStore`` that ``run()``, ``summary()`` and downstream readers use (``[]``, ``put``, ``append``, ``keys``, ``in``,
context manager)."""
from __future__ import annotations

import json
import logging
import os
import shutil

import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)
_WARNED = set()
_EXPORTS = {}      # h5 path -> thread writing it from the Parquet store (see FrameStore.close)


def wait_for_exports(path=None):
    """Block until the ``.h5`` export of ``path`` (default: of every store closed by this process) is on disk."""
    for p in ([str(path)] if path is not None else list(_EXPORTS)):
        t = _EXPORTS.pop(p, None)
        if t is not None:
            t.join()


def _export_job(h5_path):
    try:
        FrameStore(h5_path, mode="r").write_h5(h5_path)
    except Exception as exc:      # the Parquet store is the working copy; the .h5 is an export of it
        logger.warning("could not write %s: %s", h5_path, exc)


def pytables_available():
    try:
        import tables  # noqa: F401
        return True
    except Exception:
        return False


def _encode_mixed(df):
    """Object columns that Arrow cannot type (numbers and strings in one column, as in the reference's ``Flow`` column:
    'hydrograph' or a discharge) are stored cell by cell as JSON text; returns (frame to write, names of such columns)."""
    import pyarrow as pa
    mixed = []
    out = df
    for c in df.columns:
        col = df[c]
        if col.dtype != object:
            continue
        try:
            pa.array(col.to_numpy(), from_pandas=True)
        except (pa.ArrowInvalid, pa.ArrowTypeError):
            mixed.append(c)
    if mixed:
        out = df.copy()
        for c in mixed:
            out[c] = [json.dumps(_plain(v)) for v in df[c]]
    return out, mixed


def _plain(v):
    if v is None or (isinstance(v, float) and np.isnan(v)):
        return None
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (np.bool_,)):
        return bool(v)
    if isinstance(v, (str, int, float, bool)):
        return v
    return str(v)


class FrameStore:
    """Directory-backed stand-in for ``pd.HDFStore`` (keys are stored with a leading '/'): Parquet parts per key."""

    FORMAT = "stryke_b200.framestore/2 (parquet parts)"

    def __init__(self, path, mode="a"):
        self.h5_path = str(path)
        self.dir = self.h5_path + ".store"
        self.mode = mode
        if mode != "r":
            wait_for_exports(self.h5_path)      # an export still reading the parts this store is about to change
        if mode == "w" and os.path.isdir(self.dir):
            shutil.rmtree(self.dir)
        if mode == "r" and not os.path.isdir(self.dir):
            raise FileNotFoundError(f"no result store at {self.dir}")
        os.makedirs(self.dir, exist_ok=True)
        self._manifest_path = os.path.join(self.dir, "manifest.json")
        self._manifest = {"format": self.FORMAT, "keys": {}}
        if os.path.isfile(self._manifest_path):
            with open(self._manifest_path) as fh:
                m = json.load(fh)
            if m.get("format") != self.FORMAT:
                raise RuntimeError(f"{self.dir}: unknown store format {m.get('format')!r} (written by another version; "
                                   "stores are never unpickled — re-run the project)")
            self._manifest = m
        self._dirty = False
        if mode != "r" and self.h5_path not in _WARNED:
            _WARNED.add(self.h5_path)
            logger.warning("PyTables is not installed: results go to %s (Parquet parts per HDF key); %s is written from them by "
                           "stryke_b200.h5lite when the store is closed (pandas table layout, frames up to STRYKE_B200_H5_MAX_ROWS "
                           "rows; not verified against libhdf5 on this machine).  `python -m stryke_b200.store %s <target.h5>` "
                           "converts a store explicitly (through pandas.HDFStore where PyTables exists)",
                           self.dir, self.h5_path, self.h5_path)

    @staticmethod
    def _norm(key):
        return "/" + str(key).strip("/")

    def _dir(self, key):
        safe = self._norm(key).strip("/").replace("/", "__").replace(" ", "_")
        return os.path.join(self.dir, safe)

    def keys(self):
        return sorted(self._manifest["keys"])

    def __contains__(self, key):
        return self._norm(key) in self._manifest["keys"]

    def __getitem__(self, key):
        return self.get(key)

    def get(self, key, categorical=False):
        """The frame stored under ``key``.  ``categorical=True`` leaves dictionary-encoded label columns as pandas
        categoricals instead of turning them back into the object columns the writer had (summary() compares long frames by
        the integer codes and builds the object columns itself)."""
        k = self._norm(key)
        rec = self._manifest["keys"].get(k)
        if rec is None:
            raise KeyError(f"No object named {key} in the file")
        parts = [self._read_part(os.path.join(self._dir(k), f"part-{i:05d}.parquet"), rec, categorical) for i in range(rec["parts"])]
        df = parts[0] if len(parts) == 1 else pd.concat(parts, ignore_index=rec.get("range_index", True))
        return df

    @staticmethod
    def _read_part(path, rec, categorical=False):
        df = pd.read_parquet(path)
        for c in rec.get("json_columns", []):
            df[c] = pd.Series([json.loads(v) if v is not None else None for v in df[c]], index=df.index, dtype=object)
            df[c] = df[c].where(df[c].notna(), np.nan)
        for c in rec.get("object_columns", []):          # keep what the writer had as plain object columns
            if c in df.columns and df[c].dtype != object and not (categorical and isinstance(df[c].dtype, pd.CategoricalDtype)):
                df[c] = df[c].astype(object)
        return df

    def __setitem__(self, key, df):
        self.put(key, df)

    def _write_part(self, k, df, fresh):
        d = self._dir(k)
        if fresh and os.path.isdir(d):
            shutil.rmtree(d)
        os.makedirs(d, exist_ok=True)
        rec = self._manifest["keys"].get(k) if not fresh else None
        if rec is None:
            rec = {"parts": 0, "rows": 0, "columns": [str(c) for c in df.columns], "json_columns": [], "object_columns": [],
                   "range_index": isinstance(df.index, pd.RangeIndex)}
        out, mixed = _encode_mixed(df)
        rec["json_columns"] = sorted(set(rec["json_columns"]) | set(str(c) for c in mixed))
        rec["object_columns"] = sorted(set(rec["object_columns"]) | {str(c) for c in df.columns if df[c].dtype == object and c not in mixed})
        out.columns = [str(c) for c in out.columns]
        # label columns handed over as categoricals (runner._labels on long frames) are dictionary-encoded as they are and
        # come back as the plain object columns the reference writes
        rec["object_columns"] = sorted(set(rec["object_columns"]) | {str(c) for c in df.columns if isinstance(df[c].dtype, pd.CategoricalDtype)})
        out.to_parquet(os.path.join(d, f"part-{rec['parts']:05d}.parquet"), engine="pyarrow",
                       index=None if not isinstance(df.index, pd.RangeIndex) else False)
        rec["parts"] += 1
        rec["rows"] += int(len(df))
        self._manifest["keys"][k] = rec
        self._write_manifest()
        self._dirty = True

    def put(self, key, df, **kwargs):
        self._write_part(self._norm(key), df, fresh=True)

    def append(self, key, df, **kwargs):
        """A new Parquet part; what is already stored is not touched."""
        self._write_part(self._norm(key), df, fresh=False)

    def flush(self, *a, **k):
        pass

    def _write_manifest(self):
        tmp = self._manifest_path + ".tmp"
        with open(tmp, "w") as fh:
            json.dump(self._manifest, fh, indent=1)
        os.replace(tmp, self._manifest_path)

    def close(self):
        """Closing a store that was written to also (re)writes ``<name>.h5`` in the reference's HDF5 layout — on a worker
        thread (``$STRYKE_B200_H5_EXPORT``: ``async`` default, ``sync``, ``off``): run() and summary() return when the working
        store is complete; the next writer of the same store, ``wait_for_exports`` and interpreter exit wait for the file."""
        if self._dirty and self.mode != "r":
            self._dirty = False
            how = os.environ.get("STRYKE_B200_H5_EXPORT", "async").lower()
            if how == "sync":
                _export_job(self.h5_path)
            elif how != "off":
                import threading
                t = threading.Thread(target=_export_job, args=(self.h5_path,), name="stryke_b200-h5-export")
                _EXPORTS[self.h5_path] = t
                t.start()

    def write_h5(self, target, max_rows=None):
        """Every key of the store as ``/<key>/table`` of one HDF5 file (``stryke_b200.h5lite``).  Keys with more than
        ``max_rows`` rows (default ``$STRYKE_B200_H5_MAX_ROWS`` or 2 000 000; < 0: no limit, 0: write nothing) are left out
        with a log line: building a 1e8-row State_Daily table in memory is a deliberate step, not a side effect of close()."""
        from . import h5lite
        if max_rows is None:
            max_rows = int(os.environ.get("STRYKE_B200_H5_MAX_ROWS", "2000000"))
        if max_rows == 0:
            return None
        frames, skipped = {}, []
        for k in self.keys():
            if 0 < max_rows < int(self._manifest["keys"][k].get("rows", 0)):
                skipped.append(k)
                continue
            frames[k.strip("/")] = self.get(k, categorical=True)
        if skipped:
            logger.warning("%s: %s left out of the HDF5 export (more than %d rows; STRYKE_B200_H5_MAX_ROWS=-1 or "
                           "`python -m stryke_b200.store %s <target.h5>` writes them)", target, ", ".join(skipped), max_rows, self.h5_path)
        summary_keys = {"Daily_Summary", "Beta_Distributions", "Beta_Distributions_Units", "Yearly_Summary", "Driver_Diagnostics"}
        tmp = str(target) + ".tmp"
        h5lite.write_frames(tmp, frames, data_columns={k for k in frames if k in summary_keys},
                            min_itemsize={"State_Daily": {"scenario": 128, "species": 128, "state": 256},
                                          "Route_Flows": {"scenario": 128, "route": 256}, "Unit_Hours": {"scenario": 128, "route": 256}})
        os.replace(tmp, target)
        return target

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _HDF(object):
    """``pd.HDFStore`` with ``append`` defaulting to the table format the reference uses (stryke.py:3344, :3379)."""

    def __init__(self, path, mode):
        self._s = pd.HDFStore(path, mode=mode)

    def __getattr__(self, name):
        return getattr(self._s, name)

    def __getitem__(self, key):
        return self._s[key]

    def get(self, key, categorical=False):      # (same signature as FrameStore.get; HDF tables hold plain object columns)
        return self._s[key]

    def __setitem__(self, key, df):
        self._s[key] = df

    def __contains__(self, key):
        return "/" + str(key).strip("/") in self._s.keys()

    def append(self, key, df, **kwargs):
        df.to_hdf(self._s, key=key, mode="a", format="table", append=True, **kwargs)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._s.close()
        return False


def open_store(path, mode="a"):
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    if pytables_available():
        if mode == "w" and os.path.exists(path):
            os.remove(path)
        return _HDF(path, mode)
    return FrameStore(path, mode)


def export(path, target, fmt=None):
    """Copy a fallback store (``<name>.h5.store/``) into a container other tools read: ``fmt='hdf'`` writes a real
    ``pandas.HDFStore`` file with the reference's keys (needs PyTables — run it where the web app runs), ``'parquet'``
    one ``.parquet`` per key (pyarrow), ``'csv'`` one ``.csv`` per key.  ``path`` is the ``.h5`` path the project used."""
    fmt = fmt or ("hdf" if str(target).endswith((".h5", ".hdf", ".hdf5")) else "parquet")
    src = FrameStore(path, mode="r")
    keys = src.keys()
    if fmt == "hdf" and not pytables_available():      # no PyTables here: the built-in writer, every key, no row limit
        logger.warning("PyTables is not installed: %s is written by stryke_b200.h5lite (pandas table layout)", target)
        return [src.write_h5(target, max_rows=-1)]
    if fmt == "hdf":
        with pd.HDFStore(target, mode="w") as out:
            for k in keys:
                df = src[k]
                table = k.strip("/") in ("Daily", "State_Daily", "Route_Flows", "Unit_Hours", "Daily_Summary", "Beta_Distributions",
                                         "Beta_Distributions_Units", "Yearly_Summary", "Driver_Diagnostics") or "simulations" in k
                out.put(k, df, format="table" if table else "fixed", **({"data_columns": True} if table else {}))
        return [target]
    os.makedirs(target, exist_ok=True)
    written = []
    for k in keys:
        name = k.strip("/").replace("/", "__").replace(" ", "_")
        df = src[k]
        if fmt == "parquet":
            p = os.path.join(target, name + ".parquet")
            df.to_parquet(p)
        elif fmt == "csv":
            p = os.path.join(target, name + ".csv")
            df.to_csv(p)
        else:
            raise ValueError(f"unknown format {fmt!r}")
        written.append(p)
    return written


if __name__ == "__main__":      # python -m stryke_b200.store <project>.h5 <target> [hdf|parquet|csv]
    import sys
    if len(sys.argv) < 3:
        raise SystemExit("usage: python -m stryke_b200.store <project>.h5 <target.h5 | target_dir> [hdf|parquet|csv]")
    for f in export(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None):
        print(f)
