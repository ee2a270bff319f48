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


def _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, keep, kernel_variant=0, rng_mode=0, flags=0,
          peaking=None, n_cells=0):
    """``peaking``: array [n_species, n_units, 5] (include/stryke_b200.h ``StrykePeaking``); the hours drawn on the device come
    back in ``keep[-1]`` ... see ``run_cells`` / ``run_compact`` (attribute ``hours`` of their results)."""
    o = _abi.Opts()
    if peaking is not None:
        pk = _abi.arr(peaking, np.float64)
        hours = np.zeros((int(n_cells), pk.shape[1]), np.float32)
        st = _abi.Peaking()
        st.n_units, st.params, st.hours = int(pk.shape[1]), _abi.ptr(pk), _abi.ptr(hours)
        keep += [pk, hours, st]
        o.peaking = C.pointer(st)
        o._hours = hours
    o.kernel_variant, o.rng_mode, o.flags = int(kernel_variant), int(rng_mode), int(flags)
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


@dataclass
class CellGrid:
    """Implicit cell list (include/stryke_b200.h ``StrykeCellGroup``): the reference's loops over iterations and days
    (stryke.py:2897-2902) per (scenario, species) group, described by a few numbers per group instead of 16 bytes per cell.
    Cell ``cell0 + d * iterations + i`` of a group = iteration ``i`` on its ``d``-th day with operating hours."""
    groups: np.ndarray          # structured, _abi.GROUP_DTYPE, ordered by cell0
    active_days: np.ndarray     # int32: day offsets (relative to the group's day0) of the days with tot_hours > 0
    id_days: int                # Philox id = (id_base + iteration) * id_days + day offset + id_add
    id_add: int = 0

    @property
    def n_cells(self):
        g = self.groups
        return int((g["iterations"].astype(np.int64) * g["n_active_days"]).sum())

    def explicit(self):
        """The same list as three arrays (cell_day, cell_species, cell_id)."""
        day, spc, ids = [], [], []
        for g in self.groups:
            rel = self.active_days[g["active_off"]:g["active_off"] + g["n_active_days"]].astype(np.int64)
            I = int(g["iterations"])
            dd = np.repeat(rel, I)
            ii = np.tile(np.arange(I, dtype=np.uint64), len(rel))
            day.append((int(g["day0"]) + dd).astype(np.int32))
            spc.append(np.full(dd.shape[0], int(g["species"]), np.int32))
            ids.append((np.uint64(g["id_base"]) + ii) * np.uint64(self.id_days) + dd.astype(np.uint64) + np.uint64(self.id_add))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        return cat(day, np.int32), cat(spc, np.int32), cat(ids, np.uint64)

    def iterations_slice(self, lo_frac, hi_frac):
        """The sub-grid holding iterations [floor(I * lo_frac), floor(I * hi_frac)) of every group — how the cells are
        sharded over GPUs (iterations are independent: stryke.py:2897).  Groups left without iterations are dropped."""
        g = self.groups.copy()
        lo = np.floor(g["iterations"] * lo_frac + 1e-9).astype(np.int64)
        hi = np.floor(g["iterations"] * hi_frac + 1e-9).astype(np.int64)
        g["id_base"] = g["id_base"] + lo.astype(np.uint64)
        g["iterations"] = (hi - lo).astype(np.int32)
        its = (hi - lo)
        keep = its > 0
        g, its_lo = g[keep], lo[keep]
        n = g["iterations"].astype(np.int64) * g["n_active_days"]
        g["cell0"] = np.concatenate([[0], np.cumsum(n)[:-1]]) if len(g) else g["cell0"]
        out = CellGrid(g, self.active_days, self.id_days, self.id_add)
        out.iter_lo = its_lo           # first iteration of every kept group (frame assembly maps rows back with it)
        out.kept = np.nonzero(keep)[0]
        return out


def _tables(fac: FacilityTables, days: DayTables, species: np.ndarray, cell_day, cell_species, cell_id, keep):
    """C structs of the tables; ``cell_day`` may be a ``CellGrid`` (then the other two are None)."""
    species = _abi.arr(species, np.float64).reshape(-1, _abi.SPECIES_W)
    keep.append(species)
    st = _abi.SpeciesTable()
    st.n_species, st.params = species.shape[0], _abi.ptr(species)
    cl = _abi.Cells()
    if isinstance(cell_day, CellGrid):
        grid = cell_day
        groups = np.ascontiguousarray(grid.groups, dtype=_abi.GROUP_DTYPE)
        act = _abi.arr(grid.active_days, np.int32)
        keep += [groups, act]
        cl.n_cells = grid.n_cells
        cl.n_groups, cl.n_active = len(groups), act.shape[0]
        cl.groups, cl.active_days = groups.ctypes.data_as(_abi._p), _abi.ptr(act)
        cl.id_days, cl.id_add = int(grid.id_days), int(grid.id_add) & 0xFFFFFFFFFFFFFFFF
    else:
        cell_day, cell_species = _abi.arr(cell_day, np.int32), _abi.arr(cell_species, np.int32)
        cell_id = _abi.arr(cell_id, np.uint64)
        keep += [cell_day, cell_species, cell_id]
        cl.n_cells = cell_day.shape[0]
        cl.cell_day, cl.cell_species, cl.cell_id = _abi.ptr(cell_day), _abi.ptr(cell_species), _abi.ptr(cell_id)
    return fac.c_struct(), days.c_struct(), st, cl


@dataclass
class CompactResult:
    """Compact tallies of ``stryke_b200_run_compact`` (layout: include/stryke_b200.h ``StrykeCompact``)."""
    n_cells: int
    n_pairs: int
    row_w: int
    narrow_max: int
    n_slots: int
    n_fish: int
    blk_row: np.ndarray
    blk_mask: np.ndarray
    blk_wide: np.ndarray
    rows: np.ndarray            # (n_slots, row_w) uint32

    def present(self):
        """bool[n_cells]: the cell holds fish (its species was present that day)."""
        return np.unpackbits(np.ascontiguousarray(self.blk_mask).view(np.uint8), bitorder="little")[:self.n_cells].astype(bool)

    def wide(self):
        return np.unpackbits(np.ascontiguousarray(self.blk_wide).view(np.uint8), bitorder="little")[:self.n_cells].astype(bool)

    def slots(self):
        """(cells with fish, their slot, wide flag of those cells): the slot of cell 32 b + j is blk_row[b] + the slots of
        the block's present cells before it (launches over ranges of cells may start a range at any slot)."""
        NB = len(self.blk_row)
        bits = lambda a: np.unpackbits(np.ascontiguousarray(a).view(np.uint8), bitorder="little")
        pres, wide = bits(self.blk_mask).astype(bool), bits(self.blk_wide).astype(bool)
        per = pres.astype(np.int64) + wide
        cs = np.cumsum(per) - per
        slot = cs - np.repeat(cs[::32], 32) + np.repeat(self.blk_row.astype(np.int64), 32)
        cells = np.flatnonzero(pres[:self.n_cells])
        return cells, slot[cells], wide[cells]

    def columns(self):
        """Per present cell (in cell order): cells, pop_size, daily[., 5] (num_entrained, num_survived, mortality_impingement,
        _blade_strike, _barotrauma), state[., n_pairs, 2] (successes, count) as int64."""
        cells, slot, wide = self.slots()
        P, W = self.n_pairs, self.row_w
        flat = np.ascontiguousarray(self.rows).reshape(-1)
        h = flat.view(np.uint16).reshape(-1, 2 * W)          # little-endian halves: low, high
        contiguous = len(slot) > 0 and not wide.any() and slot[-1] - slot[0] == len(slot) - 1
        r = h[slot[0]:slot[-1] + 1] if contiguous else h[slot]     # (a run without wide cells and gaps: a view, no gather)
        n = r[:, 0].astype(np.int64)
        daily = np.stack([r[:, 2], r[:, 3], r[:, 4], r[:, 5], r[:, 1]], axis=1).astype(np.int64)
        state = r[:, 6:6 + 2 * P].reshape(-1, P, 2)              # uint16 while every cell is narrow
        wj = np.flatnonzero(wide)
        if len(wj):                                          # cells above narrow_max use 32-bit counters over two slots
            state = state.astype(np.uint32)
            idx = (slot[wj] * W)[:, None] + np.arange(2 * P + 6)[None, :]
            w = flat[idx].astype(np.int64)
            n[wj], daily[wj], state[wj] = w[:, 0], w[:, 1:6], w[:, 6:6 + 2 * P].reshape(-1, P, 2)
        return cells, n, daily, state

    def dense(self) -> "CellResult":
        """The same tallies in the dense layout of ``run_cells``."""
        cells, n, daily, state = self.columns()
        out_n = np.zeros(self.n_cells, np.int32)
        out_d = np.zeros((self.n_cells, _abi.DAILY_W), np.uint32)
        out_s = np.zeros((self.n_cells, self.n_pairs, 2), np.uint32)
        out_n[cells], out_d[cells], out_s[cells] = n, daily, state
        return CellResult(out_n, out_d, out_s, None)


def _pinned(shape, dtype):
    """Pinned host array when torch is around (copies overlap kernels), else a plain one."""
    import sys
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    if "torch" in sys.modules:
        import torch
        try:
            t = torch.empty(max(n, 1) * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
            a = t.numpy().view(dtype)[:n].reshape(shape)
            return a, t
        except Exception:
            pass
    return np.empty(shape, dtype), None


def run_compact(fac, days, species, cells, cell_species=None, cell_id=None, *, device=0, precision=32, seed=0,
                max_daily_fish=100000, blocks_per_sm=0, kernel_variant=0, rows_cap=None, n_slabs=0, buffers=None,
                peaking=None) -> CompactResult:
    """One call of ``stryke_b200_run_compact``: the hot path with one tally row per cell that holds fish.  ``cells`` is a
    ``CellGrid`` or the ``cell_day`` array.  ``rows_cap`` (slots) defaults to two per cell, which always suffices; pass an
    estimate for population-mode projects (the call fails loudly and names the count it needed when it was too small)."""
    keep = []
    f, d, st, cl = _tables(fac, days, species, cells, cell_species, cell_id, keep)
    o = _opts(device, precision, seed, max_daily_fish, blocks_per_sm, None, None, keep, kernel_variant, 0, 0, peaking, int(cl.n_cells))
    C_ = int(cl.n_cells)
    NB = (C_ + 31) // 32
    W = fac.n_pairs + 3
    if rows_cap is None:
        rows_cap = 2 * C_
    rows_cap = max(int(rows_cap), 1)
    if buffers is None:
        buffers = compact_buffers(C_, fac.n_pairs, rows_cap)
    blk, rows, _owners = buffers
    if blk.shape != (3, max(NB, 1)) or rows.shape[0] < rows_cap or rows.shape[1] != W:
        raise ValueError("compact buffers do not fit this call")
    c = _abi.Compact()
    c.n_blocks = NB
    c.blk_row, c.blk_mask, c.blk_wide = (blk[i].ctypes.data_as(_abi._p) for i in range(3))
    c.rows, c.rows_cap, c.n_slabs = rows.ctypes.data_as(_abi._p), rows_cap, int(n_slabs)
    _abi.check(_abi.lib().stryke_b200_run_compact(C.byref(f), C.byref(d), C.byref(st), C.byref(cl), C.byref(o), C.byref(c)))
    ns = int(c.n_slots)
    res = CompactResult(C_, fac.n_pairs, int(c.row_w), int(c.narrow_max), ns, int(c.n_fish), blk[0, :NB], blk[1, :NB], blk[2, :NB],
                        rows[:ns])
    res.hours = getattr(o, "_hours", None)
    return res


def compact_buffers(n_cells, n_pairs, rows_cap):
    """Host buffers for ``run_compact`` (pinned when torch is loaded): (block tables [3, n_blocks], rows [rows_cap, row_w])."""
    NB = max((int(n_cells) + 31) // 32, 1)
    blk, o1 = _pinned((3, NB), np.uint32)
    rows, o2 = _pinned((max(int(rows_cap), 1), int(n_pairs) + 3), np.uint32)
    return blk, rows, (o1, o2)


def run_cells(fac, days, species, cell_day, cell_species, cell_id, *, device=0, precision=32, seed=0,
              max_daily_fish=100000, blocks_per_sm=0, injected: Injected | None = None,
              trace: TraceBuffers | None = None, kernel_variant=0, out=None, rng_mode=0, flags=0, peaking=None) -> CellResult:
    """One call of ``stryke_b200_run``: everything between the presence roll and the Daily / State_Daily tallies
    (stryke.py:3025-3366) for the listed cells, on the GPU, host buffers in and out."""
    keep = []
    f, d, st, cl = _tables(fac, days, species, cell_day, cell_species, cell_id, keep)
    o = _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, keep, kernel_variant, rng_mode, flags,
              peaking, int(cl.n_cells))
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
    res = CellResult(out_n, out_d, out_s, trace)
    res.hours = getattr(o, "_hours", None)        # peaking operations: hours per (cell, unit) drawn on the device
    return res


class Plan:
    """Device-resident tables + asynchronous launches into caller-owned device buffers."""

    def __init__(self, fac, days, species, cell_day, cell_species, cell_id, *, device=0, precision=32, seed=0,
                 max_daily_fish=100000, blocks_per_sm=0, injected=None, trace=None, kernel_variant=0, rng_mode=0, flags=0):
        self._keep = []
        f, d, st, cl = _tables(fac, days, species, cell_day, cell_species, cell_id, self._keep)
        o = _opts(device, precision, seed, max_daily_fish, blocks_per_sm, injected, trace, self._keep, kernel_variant, rng_mode, flags)
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

    def launch_compact(self, cell_n, rows, cell_lo=0, cell_hi=-1, first_slot=0, slots_after=None, stream=None):
        """Enqueue the hot path for cells [cell_lo, cell_hi) with compact rows (``rows``: uint32/int32 tensor of
        [rows_cap, n_pairs + 3]; ``cell_n``: int32[n_cells]).  ``first_slot`` < 0 continues after the previous range;
        ``slots_after``: optional int64 device tensor (1 element) that receives the slot count after this range."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        sa = C.c_void_p(slots_after.data_ptr()) if slots_after is not None else None
        _abi.check(_abi.lib().stryke_b200_plan_launch_compact(self._h, int(cell_lo), int(cell_hi), C.c_void_p(cell_n.data_ptr()),
                                                              C.c_void_p(rows.data_ptr()), int(rows.shape[0]), int(first_slot), sa,
                                                              C.c_void_p(stream)))

    def set_compact_tables(self, blk_row, blk_mask, blk_wide):
        """The launches write the per-block tables into these int32 device tensors ([n_blocks] each) instead of the plan's own."""
        _abi.check(_abi.lib().stryke_b200_plan_set_compact_tables(self._h, C.c_void_p(blk_row.data_ptr()), C.c_void_p(blk_mask.data_ptr()),
                                                                  C.c_void_p(blk_wide.data_ptr())))

    def compact_tables(self):
        """(blk_row, blk_mask, blk_wide) as int32 torch tensors over the plan's device memory, row_w, narrow_max."""
        import torch
        ptrs = [C.c_void_p() for _ in range(3)]
        rw, nm = C.c_int32(0), C.c_int32(0)
        _abi.check(_abi.lib().stryke_b200_plan_compact_tables(self._h, C.byref(ptrs[0]), C.byref(ptrs[1]), C.byref(ptrs[2]),
                                                              C.byref(rw), C.byref(nm)))
        NB = max((self.n_cells + 31) // 32, 1)

        class _Dev:
            def __init__(self, p):
                self.__cuda_array_interface__ = {"shape": (NB,), "typestr": "<i4", "data": (int(p.value), False), "version": 2}
        ts = [torch.as_tensor(_Dev(p), device=torch.device("cuda", self.device)) for p in ptrs]
        return ts[0], ts[1], ts[2], int(rw.value), int(nm.value)

    def compact_totals(self, stream=None):
        """(slots, fish) counted by the compact launches so far (waits for the stream)."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        ns, nf = C.c_int64(0), C.c_int64(0)
        _abi.check(_abi.lib().stryke_b200_plan_compact_totals(self._h, C.c_void_p(stream), C.byref(ns), C.byref(nf)))
        return int(ns.value), int(nf.value)

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
