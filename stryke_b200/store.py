"""Project result store with the reference's HDF keys.

The reference writes ``<proj_dir>/<output_name>.h5`` through ``pandas.HDFStore`` (PyTables).  When PyTables is
importable this module uses exactly that (same keys, table format, ``min_itemsize``).  PyTables / libhdf5 are absent
from this image, so otherwise a same-keys container is written next to where the ``.h5`` would be:
``<output_name>.h5.store/`` holding one pickled DataFrame per key plus ``manifest.json``.  The API below is the
subset of ``HDFStore`` that ``run()``, ``summary()`` and downstream readers use (``[]``, ``put``, ``append``,
``keys``, ``in``, context manager)."""
from __future__ import annotations

import json
import os
import shutil

import pandas as pd


def pytables_available():
    try:
        import tables  # noqa: F401
        return True
    except Exception:
        return False


class FrameStore:
    """Directory-backed stand-in for ``pd.HDFStore`` (keys are stored with a leading '/')."""

    def __init__(self, path, mode="a"):
        self.h5_path = str(path)
        self.dir = self.h5_path + ".store"
        self.mode = mode
        if mode == "w" and os.path.isdir(self.dir):
            shutil.rmtree(self.dir)
        if mode == "r" and not os.path.isdir(self.dir):
            raise FileNotFoundError(f"no result store at {self.dir}")
        os.makedirs(self.dir, exist_ok=True)
        self._manifest_path = os.path.join(self.dir, "manifest.json")
        self._manifest = {}
        if os.path.isfile(self._manifest_path):
            with open(self._manifest_path) as fh:
                self._manifest = json.load(fh)
        self._pending = {}

    @staticmethod
    def _norm(key):
        return "/" + str(key).strip("/")

    def _file(self, key):
        safe = self._norm(key).strip("/").replace("/", "__").replace(" ", "_")
        return os.path.join(self.dir, safe + ".pkl")

    def keys(self):
        return sorted(set(self._manifest) | set(self._pending))

    def __contains__(self, key):
        return self._norm(key) in self._manifest or self._norm(key) in self._pending

    def __getitem__(self, key):
        k = self._norm(key)
        parts = []
        if k in self._manifest:
            parts.append(pd.read_pickle(self._file(k)))
        parts += self._pending.get(k, [])
        if not parts:
            raise KeyError(f"No object named {key} in the file")
        return parts[0] if len(parts) == 1 else pd.concat(parts, ignore_index=True)

    def get(self, key):
        return self[key]

    def __setitem__(self, key, df):
        self.put(key, df)

    def put(self, key, df, **kwargs):
        k = self._norm(key)
        self._pending.pop(k, None)
        df.to_pickle(self._file(k))
        self._manifest[k] = {"rows": int(len(df)), "columns": [str(c) for c in df.columns]}
        self._write_manifest()

    def append(self, key, df, **kwargs):
        self._pending.setdefault(self._norm(key), []).append(df)

    def flush(self, *a, **k):
        for k_, parts in list(self._pending.items()):
            base = [pd.read_pickle(self._file(k_))] if k_ in self._manifest else []
            df = pd.concat(base + parts, ignore_index=True) if (base or len(parts) > 1) else parts[0]
            df.to_pickle(self._file(k_))
            self._manifest[k_] = {"rows": int(len(df)), "columns": [str(c) for c in df.columns]}
        self._pending = {}
        self._write_manifest()

    def _write_manifest(self):
        with open(self._manifest_path, "w") as fh:
            json.dump(self._manifest, fh, indent=1)

    def close(self):
        if self.mode != "r":
            self.flush()

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
    if fmt == "hdf":
        if not pytables_available():
            raise RuntimeError("writing HDF5 needs PyTables (pip install tables); use fmt='parquet' or 'csv' here")
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
