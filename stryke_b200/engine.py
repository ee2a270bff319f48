"""Thin Python layer over the C ABI: one-shot host-buffer runs and device-resident plans.

``run_cells``  -> ``stryke_b200_run``  (host tables in, host tallies out; H2D/D2H inside the call)
``Plan``       -> ``stryke_b200_plan_*`` (tables resident in HBM, tallies in caller-owned torch tensors on the
                  caller's CUDA stream; used by bench.py and the multi-GPU path, where the tally tensors are
                  all-reduced with NCCL through torch.distributed)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .tables import DayTables, FacilityTables


@dataclass
class Injected:
    """Slot-tagged draws for bit-exact replay (layout: include/stryke_b200.h ``StrykeInjected``)."""
    fish_off: np.ndarray
    presence: np.ndarray
    count_draw: np.ndarray
    length_z: np.ndarray
    dice: np.ndarray
    rR: np.ndarray
    depth: np.ndarray
    cause: np.ndarray        # (M, 4, n_fish)
    route: np.ndarray
    ghost_rR: np.ndarray | None = None
    ghost_depth: np.ndarray | None = None
    ghost_cause: np.ndarray | None = None   # (M, 4, n_cells)

    def c_struct(self, keep):
        s = _abi.Injected()
        self.fish_off = _abi.arr(self.fish_off, np.int64)
        s.n_fish_total = int(self.fish_off[-1])
        s.fish_off = _abi.ptr(self.fish_off)
        for k in ("presence", "count_draw", "length_z", "dice", "rR", "depth", "cause", "route", "ghost_rR",
                  "ghost_depth", "ghost_cause"):
            v = getattr(self, k)
            if v is not None:
                v = _abi.arr(v, np.float64)
                setattr(self, k, v)
                keep.append(v)
            setattr(s, k, _abi.ptr(v))
        return s


@dataclass
class TraceBuffers:
    length_ft: np.ndarray
    length_ft64: np.ndarray
    state: np.ndarray
    rate: np.ndarray
    survival: np.ndarray
    strike: np.ndarray | None
    baro: np.ndarray | None

    @classmethod
    def allocate(cls, n_moves, n_fish, probs=True):
        return cls(np.zeros(n_fish, np.float32), np.zeros(n_fish, np.float64), np.zeros((n_moves, n_fish), np.uint8),
                   np.zeros((n_moves, n_fish), np.float32), np.zeros((n_moves, n_fish), np.uint8),
                   np.zeros((n_moves, n_fish), np.float64) if probs else None,
                   np.zeros((n_moves, n_fish), np.float64) if probs else None)

    def c_struct(self):
        t = _abi.Trace()
        for k in ("length_ft", "length_ft64", "state", "rate", "survival", "strike", "baro"):
            setattr(t, k, _abi.ptr(getattr(self, k)))
        t.n_fish = int(getattr(self, "n_fish", 0))      # > 0: capacity known from an earlier run (population mode)
        return t


def tally_tensors(n_cells, n_pairs, device, with_base=False):
    """(cell_n[C], daily[C,5], state_daily[C,P,2]) int32 views of one zeroed flat device buffer."""
    import torch
    C_, W, P = int(n_cells), _abi.DAILY_W, int(n_pairs)
    base = torch.zeros(C_ * (1 + W + 2 * P), dtype=torch.int32, device=torch.device("cuda", device))
    out = (base[:C_], base[C_:C_ * (1 + W)].view(C_, W), base[C_ * (1 + W):].view(C_, P, 2))
    return (*out, base) if with_base else out


@dataclass
class CellResult:
    cell_n: np.ndarray        # (C,) int32 pop_size
    daily: np.ndarray         # (C, 5) uint32: num_entrained, num_survived, mortality_impingement, _blade_strike, _barotrauma
    state_daily: np.ndarray   # (C, n_pairs, 2) uint32: successes, count
    trace: TraceBuffers | None = None


def _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, keep, kernel_variant=0):
    o = _abi.Opts()
    o.kernel_variant = int(kernel_variant)
    o.device, o.precision, o.seed = int(device), int(precision), int(seed) & 0xFFFFFFFFFFFFFFFF
    o.max_daily_fish, o.blocks_per_sm = int(max_daily_fish), int(blocks_per_sm)
    if injected is not None:
        inj = injected.c_struct(keep)
        keep.append(inj)
        o.injected = C.pointer(inj)
    if trace is not None:
        tr = trace.c_struct()
        keep.append(tr)
        o.trace = C.pointer(tr)
    return o


def _tables(fac: FacilityTables, days: DayTables, species: np.ndarray, cell_day, cell_species, cell_id, keep):
    species = _abi.arr(species, np.float64).reshape(-1, _abi.SPECIES_W)
    cell_day, cell_species = _abi.arr(cell_day, np.int32), _abi.arr(cell_species, np.int32)
    cell_id = _abi.arr(cell_id, np.uint64)
    keep += [species, cell_day, cell_species, cell_id]
    st = _abi.SpeciesTable()
    st.n_species, st.params = species.shape[0], _abi.ptr(species)
    cl = _abi.Cells()
    cl.n_cells = cell_day.shape[0]
    cl.cell_day, cl.cell_species, cl.cell_id = _abi.ptr(cell_day), _abi.ptr(cell_species), _abi.ptr(cell_id)
    return fac.c_struct(), days.c_struct(), st, cl


def run_cells(fac, days, species, cell_day, cell_species, cell_id, *, device=0, precision=32, seed=0,
              max_daily_fish=100000, blocks_per_sm=0, injected: Injected | None = None,
              trace: TraceBuffers | None = None, kernel_variant=0, out=None) -> CellResult:
    """One call of ``stryke_b200_run``: everything between the presence roll and the Daily / State_Daily tallies
    (stryke.py:3025-3366) for the listed cells, on the GPU, host buffers in and out."""
    keep = []
    f, d, st, cl = _tables(fac, days, species, cell_day, cell_species, cell_id, keep)
    o = _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, keep, kernel_variant)
    n = int(cl.n_cells)
    if out is not None:      # caller-owned (e.g. pinned) host buffers
        out_n, out_d, out_s = out
        for name, a, dt, shape in (("cell_n", out_n, np.int32, (n,)), ("daily", out_d, np.uint32, (n, _abi.DAILY_W)),
                                   ("state_daily", out_s, np.uint32, (n, fac.n_pairs, 2))):
            if a.dtype != dt or a.shape != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"out buffer {name}: need C-contiguous {np.dtype(dt).name}{list(shape)}, "
                                 f"got {a.dtype.name}{list(a.shape)}")
    else:
        out_n = np.zeros(n, np.int32)
        out_d = np.zeros((n, _abi.DAILY_W), np.uint32)
        out_s = np.zeros((n, fac.n_pairs, 2), np.uint32)
    t = _abi.Tallies()
    t.cell_n, t.daily, t.state_daily = _abi.ptr(out_n), _abi.ptr(out_d), _abi.ptr(out_s)
    _abi.check(_abi.lib().stryke_b200_run(C.byref(f), C.byref(d), C.byref(st), C.byref(cl), C.byref(o), C.byref(t)))
    return CellResult(out_n, out_d, out_s, trace)


class Plan:
    """Device-resident tables + asynchronous launches into caller-owned device buffers."""

    def __init__(self, fac, days, species, cell_day, cell_species, cell_id, *, device=0, precision=32, seed=0,
                 max_daily_fish=100000, blocks_per_sm=0, injected=None, trace=None, kernel_variant=0):
        self._keep = []
        f, d, st, cl = _tables(fac, days, species, cell_day, cell_species, cell_id, self._keep)
        o = _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, self._keep, kernel_variant)
        self.n_cells, self.n_pairs, self.device = int(cl.n_cells), fac.n_pairs, int(device)
        self._h = C.c_void_p()
        _abi.check(_abi.lib().stryke_b200_plan_create(C.byref(f), C.byref(d), C.byref(st), C.byref(cl), C.byref(o),
                                                      C.byref(self._h)))

    def allocate(self, with_base=False):
        """Tally tensors on the plan's device (torch is plumbing: device memory + streams + NCCL).  The three tensors
        are views of ONE flat int32 buffer (``with_base=True`` also returns it), so a multi-GPU step finishes with a
        single collective."""
        return tally_tensors(self.n_cells, self.n_pairs, self.device, with_base)

    def launch(self, cell_n, daily, state_daily, stream=None):
        """Enqueue count -> scan -> passage on ``stream`` (int handle; default: torch's current stream)."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _abi.check(_abi.lib().stryke_b200_plan_launch(self._h, C.c_void_p(cell_n.data_ptr()), C.c_void_p(daily.data_ptr()),
                                                      C.c_void_p(state_daily.data_ptr()), C.c_void_p(stream)))

    def kernel(self):
        """Name of the passage kernel this plan launches (general / generic throughput / NVRTC-specialised)."""
        return _abi.lib().stryke_b200_plan_kernel(self._h).decode()

    def time_kernel(self, enable=True):
        """Bracket the passage kernel of every later launch with CUDA events (measurement support)."""
        _abi.check(_abi.lib().stryke_b200_plan_time_kernel(self._h, C.c_int(1 if enable else 0)))

    def kernel_ms(self):
        """(summed passage-kernel milliseconds, launches) since the last call; waits for those launches."""
        ms, n = C.c_double(0.0), C.c_int32(0)
        _abi.check(_abi.lib().stryke_b200_plan_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def set_seed(self, seed):
        _abi.check(_abi.lib().stryke_b200_plan_set_seed(self._h, C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)))

    def status(self, stream=None):
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        nf = C.c_int64(0)
        _abi.check(_abi.lib().stryke_b200_plan_status(self._h, C.c_void_p(stream), C.byref(nf)))
        return int(nf.value)

    def close(self):
        if self._h:
            _abi.lib().stryke_b200_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def strike_survival(kind, params15, length_ft, rR=None, precision=64, device=0):
    """Kaplan / Propeller / Francis / Pump blade-strike survival on the device (stryke.py:846-1127)."""
    L = _abi.arr(np.atleast_1d(length_ft), np.float64)
    r = None if rR is None else _abi.arr(np.broadcast_to(np.atleast_1d(rR), L.shape), np.float64)
    p = _abi.arr(params15, np.float64)
    out = np.zeros_like(L)
    _abi.check(_abi.lib().stryke_b200_strike_survival(int(kind), _abi.ptr(p), L.shape[0], _abi.ptr(L), _abi.ptr(r),
                                                      int(precision), int(device), _abi.ptr(out)))
    return out


def baro_survival(fish_depth_ft, tail_depth_ft, beta_0, beta_1, precision=64, device=0):
    """1 - Pflugrath injury probability for fish acclimated at ``fish_depth_ft`` (stryke.py:1496-1510)."""
    d = _abi.arr(np.atleast_1d(fish_depth_ft), np.float64)
    out = np.zeros_like(d)
    _abi.check(_abi.lib().stryke_b200_baro_survival(d.shape[0], _abi.ptr(d), float(tail_depth_ft), float(beta_0),
                                                    float(beta_1), int(precision), int(device), _abi.ptr(out)))
    return out


def philox(ctr4, key2, device=0):
    c = _abi.arr(ctr4, np.uint32).reshape(-1, 4)
    k = _abi.arr(np.broadcast_to(np.asarray(key2, np.uint32).reshape(-1, 2), (c.shape[0], 2)), np.uint32)
    out = np.zeros_like(c)
    _abi.check(_abi.lib().stryke_b200_philox(c.shape[0], _abi.ptr(c), _abi.ptr(k), int(device), _abi.ptr(out)))
    return out


def calibrate_issue(device=0, stream=0):
    ops, ms = C.c_double(0), C.c_double(0)
    _abi.check(_abi.lib().stryke_b200_calibrate_issue(int(device), C.c_void_p(stream), C.byref(ops), C.byref(ms)))
    return ops.value, ms.value


def bootstrap_means(data, n_boot, seed=0, device=0):
    """Means of ``n_boot`` resamples (with replacement) of ``data``, computed on the device (bootstrap_mean_ci's loop,
    stryke.py:155-159)."""
    a = np.ascontiguousarray(data, dtype=np.float64)
    out = np.empty(int(n_boot), np.float64)
    lib = _abi.lib()
    lib.stryke_b200_bootstrap_means.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_int32, C.c_void_p]
    _abi.check(lib.stryke_b200_bootstrap_means(_abi.ptr(a), a.size, int(n_boot), C.c_uint64(int(seed) & (2 ** 64 - 1)), int(device),
                                               _abi.ptr(out)))
    return out


def launch_count():
    return int(_abi.lib().stryke_b200_launch_count())
