"""ctypes binding of ``libstryke_b200.so`` (C ABI declared in include/stryke_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m stryke_b200.build``.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is visible when a compute entry point is called, the
call raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# STRYKE_B200_LIB: file name of another build of the library in this directory (e.g. the bounds-checked debug build made
# with `STRYKE_B200_BOUNDS_CHECK=1 python -m stryke_b200.build --out libstryke_b200_chk.so`)
LIB_PATH = os.path.join(_HERE, os.path.basename(os.environ.get("STRYKE_B200_LIB", "") or "libstryke_b200.so"))

KIND = {"a priori": 0, "Kaplan": 1, "Propeller": 2, "Francis": 3, "Pump": 4}
ENT = {"fixed": 0, "Pareto": 1, "Log Normal": 2, "Weibull": 3, "Extreme": 4}
UNIT_STATIC_W, UNIT_COEF_W, SPECIES_W, DAILY_W = 4, 8, 20, 5
MAX_NODES, MAX_UNITS, MAX_EDGES, MAX_MOVES, MAX_PAIRS = 255, 64, 1024, 64, 128
E_MAX_FISH, E_NOGPU = -4, -5
FLAG_NO_PROMOTE = 1
ABI_VERSION = 2

_p = C.c_void_p


class Facility(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_units", C.c_int32), ("n_moves", C.c_int32), ("n_edges", C.c_int32),
                ("n_pairs", C.c_int32), ("start_node", C.c_int32), ("barotrauma", C.c_int32),
                ("node_kind", _p), ("node_apriori", _p), ("node_unit", _p), ("node_has_U", _p), ("edge_off", _p),
                ("edge_dst", _p), ("move_lex_rank", _p), ("level_off", _p), ("pair_node", _p), ("unit_static", _p)]


class DayTables(C.Structure):
    _fields_ = [("n_days", C.c_int32), ("curr_Q", _p), ("edge_cdf", _p), ("node_stay", _p), ("unit_coef", _p)]


class SpeciesTable(C.Structure):
    _fields_ = [("n_species", C.c_int32), ("params", _p)]


class CellGroup(C.Structure):
    _fields_ = [("cell0", C.c_int64), ("iterations", C.c_int32), ("n_active_days", C.c_int32), ("day0", C.c_int32),
                ("species", C.c_int32), ("active_off", C.c_int32), ("reserved", C.c_int32), ("id_base", C.c_uint64)]


GROUP_DTYPE = np.dtype([("cell0", np.int64), ("iterations", np.int32), ("n_active_days", np.int32), ("day0", np.int32),
                        ("species", np.int32), ("active_off", np.int32), ("reserved", np.int32), ("id_base", np.uint64)])
assert GROUP_DTYPE.itemsize == C.sizeof(CellGroup) == 40


class Cells(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("cell_day", _p), ("cell_species", _p), ("cell_id", _p),
                ("n_groups", C.c_int32), ("n_active", C.c_int32), ("groups", _p), ("active_days", _p),
                ("id_days", C.c_uint64), ("id_add", C.c_uint64)]


class Compact(C.Structure):
    _fields_ = [("n_blocks", C.c_int64), ("blk_row", _p), ("blk_mask", _p), ("blk_wide", _p), ("rows", _p),
                ("rows_cap", C.c_int64), ("n_slots", C.c_int64), ("n_fish", C.c_int64), ("row_w", C.c_int32),
                ("narrow_max", C.c_int32), ("n_slabs", C.c_int32)]


class Injected(C.Structure):
    _fields_ = [("n_fish_total", C.c_int64), ("fish_off", _p), ("presence", _p), ("count_draw", _p),
                ("length_z", _p), ("dice", _p), ("rR", _p), ("depth", _p), ("cause", _p), ("route", _p),
                ("ghost_rR", _p), ("ghost_depth", _p), ("ghost_cause", _p)]


class Trace(C.Structure):
    _fields_ = [("length_ft", _p), ("length_ft64", _p), ("state", _p), ("rate", _p), ("survival", _p),
                ("strike", _p), ("baro", _p), ("n_fish", C.c_int64)]


class Peaking(C.Structure):
    _fields_ = [("n_units", C.c_int32), ("params", _p), ("hours", _p)]


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("seed", C.c_uint64), ("max_daily_fish", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("injected", C.POINTER(Injected)), ("trace", C.POINTER(Trace)),
                ("kernel_variant", C.c_int32), ("rng_mode", C.c_int32), ("flags", C.c_uint32), ("peaking", C.POINTER(Peaking))]


class Tallies(C.Structure):
    _fields_ = [("cell_n", _p), ("daily", _p), ("state_daily", _p)]


class StrykeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None


def lib():
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m stryke_b200.build` "
                          "(nvcc -gencode arch=compute_100a,code=sm_100a). stryke_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.stryke_b200_last_error.restype = C.c_char_p
    L.stryke_b200_abi_version.restype = C.c_int
    L.stryke_b200_device_count.restype = C.c_int
    L.stryke_b200_launch_count.restype = C.c_uint64
    P = C.POINTER
    L.stryke_b200_run.argtypes = [P(Facility), P(DayTables), P(SpeciesTable), P(Cells), P(Opts), P(Tallies)]
    L.stryke_b200_plan_create.argtypes = [P(Facility), P(DayTables), P(SpeciesTable), P(Cells), P(Opts), P(_p)]
    L.stryke_b200_plan_launch.argtypes = [_p, _p, _p, _p, _p]
    L.stryke_b200_run_compact.argtypes = [P(Facility), P(DayTables), P(SpeciesTable), P(Cells), P(Opts), P(Compact)]
    L.stryke_b200_plan_launch_compact.argtypes = [_p, C.c_int64, C.c_int64, _p, _p, C.c_int64, C.c_int64, _p, _p]
    L.stryke_b200_plan_compact_tables.argtypes = [_p, P(_p), P(_p), P(_p), P(C.c_int32), P(C.c_int32)]
    L.stryke_b200_plan_compact_totals.argtypes = [_p, _p, P(C.c_int64), P(C.c_int64)]
    L.stryke_b200_plan_set_compact_tables.argtypes = [_p, _p, _p, _p]
    L.stryke_b200_plan_set_compact_tables.restype = C.c_int
    L.stryke_b200_plan_set_seed.argtypes = [_p, C.c_uint64]
    L.stryke_b200_plan_kernel.argtypes = [_p]
    L.stryke_b200_plan_kernel.restype = C.c_char_p
    L.stryke_b200_set_source_dir.argtypes = [C.c_char_p]
    L.stryke_b200_set_source_dir.restype = C.c_int
    L.stryke_b200_jit_selftest.restype = C.c_int
    L.stryke_b200_plan_status.argtypes = [_p, _p, P(C.c_int64)]
    L.stryke_b200_plan_time_kernel.argtypes = [_p, C.c_int]
    L.stryke_b200_plan_kernel_ms.argtypes = [_p, P(C.c_double), P(C.c_int32)]
    L.stryke_b200_plan_fetch_trace.argtypes = [_p, P(Trace), _p]
    L.stryke_b200_plan_fetch_hours.argtypes = [_p, _p, _p]
    L.stryke_b200_plan_fetch_hours.restype = C.c_int
    L.stryke_b200_plan_destroy.argtypes = [_p]
    L.stryke_b200_strike_survival.argtypes = [C.c_int, _p, C.c_int64, _p, _p, C.c_int, C.c_int, _p]
    L.stryke_b200_baro_survival.argtypes = [C.c_int64, _p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _p]
    L.stryke_b200_philox.argtypes = [C.c_int64, _p, _p, C.c_int, _p]
    L.stryke_b200_calibrate_issue.argtypes = [C.c_int, _p, P(C.c_double), P(C.c_double)]
    for name in ("run_compact", "plan_launch_compact", "plan_compact_tables", "plan_compact_totals", "run", "plan_create", "plan_launch", "plan_set_seed", "plan_status", "plan_fetch_trace", "plan_destroy",
                 "plan_time_kernel", "plan_kernel_ms",
                 "strike_survival", "baro_survival", "philox", "calibrate_issue"):
        getattr(L, "stryke_b200_" + name).restype = C.c_int
    L.stryke_b200_set_source_dir(os.path.join(_HERE, "csrc").encode())
    _lib = L
    return L


EXPORTS = ["stryke_b200_plan_set_compact_tables", "stryke_b200_plan_fetch_hours", "stryke_b200_run_compact", "stryke_b200_plan_launch_compact", "stryke_b200_plan_compact_tables",
           "stryke_b200_plan_compact_totals", "stryke_b200_last_error", "stryke_b200_abi_version", "stryke_b200_device_count", "stryke_b200_launch_count",
           "stryke_b200_run", "stryke_b200_plan_create", "stryke_b200_plan_launch", "stryke_b200_plan_kernel", "stryke_b200_set_source_dir", "stryke_b200_jit_selftest", "stryke_b200_plan_set_seed", "stryke_b200_plan_status",
           "stryke_b200_plan_fetch_trace", "stryke_b200_plan_destroy", "stryke_b200_plan_time_kernel", "stryke_b200_plan_kernel_ms", "stryke_b200_strike_survival",
           "stryke_b200_baro_survival", "stryke_b200_philox", "stryke_b200_calibrate_issue", "stryke_b200_bootstrap_means"]


def check(rc):
    if rc != 0:
        raise StrykeError(rc, lib().stryke_b200_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(_p)


def arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
