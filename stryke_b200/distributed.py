"""Multi-GPU layer: one process per GPU, cells sharded over ranks, the ranks' compact tally rows gathered at the end.

Scenario x species x iteration x day cells are independent (the reference's loops share nothing but RNG streams and the
HDF append, Stryke/stryke.py:2806-2902) and the Philox counter is keyed by the logical cell id, so any partition of the
cell list gives the same tallies (tests/test_gpu_native_rng.py).  A rank simulates the iterations [I r / N, I (r + 1) / N) of
every (scenario, species) group — the loop the reference runs at stryke.py:2897 — into compact rows (one row per cell
that holds fish, 16-bit counters; include/stryke_b200.h ``StrykeCompact``) inside ONE flat device buffer, and the ranks
exchange those buffers: every rank ends up with every rank's rows (what the north star calls the tally all-reduce; the
shards are disjoint, so the sum is a concatenation and nobody ships zeros).

Exchange.  ``push`` (NCCL process groups with CUDA peer access): the cell list of a rank is cut into slabs, every slab owns a
fixed region of the buffer, and as soon as a slab's kernels are done a copy stream pushes its region straight into every
peer's gathered buffer over NVLink (``cudaMemcpyPeerAsync`` through torch's symmetric memory — copy engines, no SMs), under
the kernels of the next slab; one tiny all-reduce at the end is the completion fence.  ``gather`` (anything else, e.g. gloo in
the CPU tests): one ``all_gather_into_tensor`` of the flat buffers after the last slab.
There is no exchange inside the simulation, hence no fused compute + collective kernel on this path.
"""
from __future__ import annotations

import os

import numpy as np

ALIGN = 1024      # ranges of cells start at multiples of this (include/stryke_b200.h STRYKE_RANGE_ALIGN)
HDR = 16          # int32 words at the head of a rank's flat buffer: [0:2) slots used (int64), [2] status code, [3] n_cells


def shard_bounds(n_cells, rank, world):
    """Contiguous block of cells of ``rank``: sizes differ by at most one."""
    base, rem = divmod(int(n_cells), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_weight(weights, world):
    """Contiguous blocks with (nearly) equal total weight — weights = expected fish per cell.  Returns world+1 bounds."""
    w = np.asarray(weights, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)


def expected_slots(grid, species, margin=1.08, extra=4096):
    """Upper estimate of the rows a grid needs: cells x occur_prob per group (+ margin); two slots per cell for groups
    whose fixed count may not fit 16-bit counters."""
    tot = 0.0
    for g in grid.groups:
        sp = species[int(g["species"])]
        cells = float(g["iterations"]) * float(g["n_active_days"])
        occur = sp[17] if np.isfinite(sp[17]) else 1.0
        fixed = int(sp[12]) == 0
        per = 2.0 if (fixed and sp[18] > 13000) else 1.0
        p = min(max(occur, 0.0), 1.0)
        tot += per * (cells * p * margin + 6.0 * np.sqrt(max(cells * p * (1 - p), 1.0))) + (0.02 * cells if not fixed else 0.0)
    return int(tot) + extra


N_SLABS = 4


def default_slab_cells(n_cells, rows_bytes):
    """Cells per slab (a multiple of ALIGN).  Large lists are cut into N_SLABS slabs: the export of one slab runs under the
    kernels of the next, and every launch costs ~0.1 ms of chunk starts and tail, so not more than that needs.  (Round 2
    first sized slabs for L2 — 16 on C3 — so that the rows count_kernel zeroed were still cached when the passage kernel
    added to them; pooled plans now write the rows of their owned cells once and nothing is zeroed for them.)"""
    import os
    n_cells = max(int(n_cells), 1)
    if n_cells < (1 << 18):
        return max(ALIGN, -(-n_cells // ALIGN) * ALIGN)
    n_slabs = N_SLABS
    if os.environ.get("STRYKE_B200_SLABS"):
        n_slabs = max(1, int(os.environ["STRYKE_B200_SLABS"]))
    return max(ALIGN, -(-n_cells // n_slabs // ALIGN) * ALIGN)


SLAB_SHARES = (0.40, 0.30, 0.20, 0.10)      # cells per slab when the rows are exported (peer push): the LAST export is not hidden


def group_density(grid, species, margin=1.08):
    """Per group of ``grid``: (first cell, cells, expected rows per cell with margin, variance of the rows per cell)."""
    out = []
    for g in grid.groups:
        sp = species[int(g["species"])]
        cells = float(g["iterations"]) * float(g["n_active_days"])
        occur = sp[17] if np.isfinite(sp[17]) else 1.0
        fixed = int(sp[12]) == 0
        per = 2.0 if (fixed and sp[18] > 13000) else 1.0
        p = min(max(occur, 0.0), 1.0)
        out.append((float(g["cell0"]), cells, per * p * margin + (0.0 if fixed else 0.02), per * per * p * (1 - p)))
    return out


def slab_plan(grids, species, n_cells_max, shares=SLAB_SHARES, scale=1.0, extra=1024):
    """(cell_lo, slot_lo): slab k of every rank covers its cells [cell_lo[k], cell_lo[k+1]) (multiples of ALIGN) and owns
    the slots [slot_lo[k], slot_lo[k+1]) of the rows buffer — sized for the rows the slab is expected to need (cells x
    occur_prob of the groups it crosses, + 6 sigma + margin), the largest over the ranks' grids, times ``scale``.  The shares
    decrease: the export of slab k runs under the kernels of slab k+1, only the last one is exposed."""
    n = max(int(n_cells_max), 1)
    if n < (1 << 18):
        shares = (1.0,)
    elif os.environ.get("STRYKE_B200_SLABS"):
        k = max(1, int(os.environ["STRYKE_B200_SLABS"]))
        shares = tuple([1.0 / k] * k)
    cum = np.cumsum((0.0,) + tuple(shares)) / float(sum(shares))
    cell_lo = [min(n, int(-(-(c * n) // ALIGN) * ALIGN)) for c in cum]
    cell_lo[0], cell_lo[-1] = 0, n
    cell_lo = sorted(set(cell_lo))
    need = np.zeros(len(cell_lo) - 1)
    for g in grids:
        dens = group_density(g, species)
        for k in range(len(cell_lo) - 1):
            lo, hi = float(cell_lo[k]), float(cell_lo[k + 1])
            e = v = 0.0
            for c0, cells, d, var in dens:
                ov = max(0.0, min(hi, c0 + cells) - max(lo, c0))
                e += ov * d
                v += ov * var
            need[k] = max(need[k], e + 6.0 * np.sqrt(max(v, 1.0)))
    caps = [int(scale * x) + extra for x in need]
    slot_lo = [0]
    for c in caps:
        slot_lo.append(slot_lo[-1] + c)
    return [int(c) for c in cell_lo], slot_lo


class ShardLayout:
    """Geometry of one rank's flat buffer: header, block tables, slab regions of rows (identical on every rank)."""

    def __init__(self, n_cells_max, n_pairs, slab_cells=None, slab_cap=None, cell_lo=None, slot_lo=None):
        """Either uniform slabs (``slab_cells`` cells and ``slab_cap`` slots each) or explicit bounds (``slab_plan``)."""
        self.W = int(n_pairs) + 3
        self.NB = (int(n_cells_max) + 31) // 32
        if cell_lo is None:
            n_slabs = max(1, -(-int(n_cells_max) // int(slab_cells)))
            cell_lo = [min(int(n_cells_max), k * int(slab_cells)) for k in range(n_slabs)] + [int(n_cells_max)]
            slot_lo = [k * int(slab_cap) for k in range(n_slabs + 1)]
        self.cell_lo = [int(c) for c in cell_lo]
        self.slot_lo = [int(x) for x in slot_lo]
        self.n_slabs = len(self.cell_lo) - 1
        self.slots = self.slot_lo[-1]
        self.blk_off = HDR
        self.rows_off = HDR + 3 * self.NB
        self.rows_off += (-self.rows_off) % 4                    # rows start 16-byte aligned
        self.total = self.rows_off + self.slots * self.W
        self.total += (-self.total) % 4

    def slab_words(self, k):
        return self.rows_off + self.slot_lo[k] * self.W, self.rows_off + self.slot_lo[k + 1] * self.W


def pack_flat(L, blk, rows, status=0, n_cells=0, slots_used=0):
    """A rank's flat int32 buffer (numpy) from block tables [3, <= NB] and rows [<= L.slots, W] — the layout the
    kernels fill in place on the device; used by the host-side tests of the exchange."""
    flat = np.zeros(L.total, np.int32)
    u = flat.view(np.uint32)
    u[0], u[1] = int(slots_used) & 0xFFFFFFFF, int(slots_used) >> 32
    flat[2], flat[3] = int(status), int(n_cells)
    b = np.zeros((3, L.NB), np.uint32)
    b[:, :blk.shape[1]] = blk
    u[L.blk_off:L.blk_off + 3 * L.NB] = b.reshape(-1)
    r = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1)
    u[L.rows_off:L.rows_off + r.shape[0]] = r
    return flat


def unpack_flat(L, flat, n_pairs):
    """(n_cells, status code, engine.CompactResult) of one rank's flat buffer (host int32 array of ``L.total`` words)."""
    from .engine import CompactResult
    u = np.ascontiguousarray(flat).view(np.uint32)
    n_cells, status = int(flat[3]), int(flat[2])
    NB = (n_cells + 31) // 32
    blk = u[L.blk_off:L.blk_off + 3 * L.NB].reshape(3, L.NB)
    rows = u[L.rows_off:L.rows_off + L.slots * L.W].reshape(-1, L.W)
    return n_cells, status, CompactResult(n_cells, int(n_pairs), L.W, 0, rows.shape[0], 0, blk[0, :NB], blk[1, :NB], blk[2, :NB], rows)


def gather_flat(local, world, group=None):
    """ONE ``all_gather_into_tensor`` of the ranks' flat buffers (NCCL on GPUs, gloo in the CPU tests): [world * len(local)]."""
    import torch
    import torch.distributed as dist
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if world == 1:
        out.copy_(local)
    else:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class ShardedRun:
    """One rank's share of a run, resident on its GPU: plan, flat result buffer, gathered buffer, streams."""

    def __init__(self, fac, days, species, grid, rank, world, device, *, precision=32, seed=0, max_daily_fish=100000,
                 kernel_variant=0, slab_cells=None, slab_cap=None, n_cells_max=None, group=None, exchange="auto", gather_to="root",
                 slabs=None):
        """``gather_to``: "root" — every rank's rows end up on rank 0, the one rank that writes the frames (each rank pushes
        1/(N-1) of what "all" costs it); "all" — on every rank (the all-reduce semantics: any rank can run ``summary()``)."""
        import torch
        import torch.distributed as dist
        from . import engine
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device, self.group = int(rank), int(world), int(device), group
        self.gather_to = gather_to
        self.targets = [p for p in range(int(world)) if p != int(rank)] if gather_to == "all" else ([0] if int(rank) != 0 else [])
        self.grid, self.fac = grid, fac
        self.n_cells = grid.n_cells
        self.dev = torch.device("cuda", device)
        self.plan = engine.Plan(fac, days, species, grid, None, None, device=device, precision=precision, seed=seed,
                                max_daily_fish=max_daily_fish, kernel_variant=kernel_variant)
        n_max = int(n_cells_max if n_cells_max is not None else self.n_cells)
        if slabs is not None:                                    # (cell_lo, slot_lo) of slab_plan: the same on every rank
            self.L = ShardLayout(n_max, fac.n_pairs, cell_lo=slabs[0], slot_lo=slabs[1])
        elif slab_cells is None and slab_cap is None:            # (ranks whose grids differ must be given a common plan)
            cell_lo, slot_lo = slab_plan([grid], np.asarray(species).reshape(-1, 20), n_max,
                                         shares=SLAB_SHARES if int(world) > 1 else (1.0,))      # one GPU: nothing to export between launches
            self.L = ShardLayout(n_max, fac.n_pairs, cell_lo=cell_lo, slot_lo=slot_lo)
        else:
            est = expected_slots(grid, np.asarray(species).reshape(-1, 20))
            if slab_cells is None:
                slab_cells = default_slab_cells(n_max, est * (fac.n_pairs + 3) * 4)
            if slab_cap is None:   # (expected_slots carries the margin; a slab that overflows is reported: slab_overflow -> the caller retries)
                slab_cap = max(1024, int(est / max(1, -(-n_max // slab_cells)) * 1.04) + 1024)
            self.L = ShardLayout(n_max, fac.n_pairs, slab_cells, slab_cap)
        L = self.L
        self.mode = "gather"
        self.symm = None
        if exchange in ("auto", "push") and world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.gathered = symm_mem.empty(world * L.total, dtype=torch.int32, device=self.dev)
                self.symm = symm_mem.rendezvous(self.gathered, group=group if group is not None else dist.group.WORLD)
                self.peers = [self.symm.get_buffer(p, (world * L.total,), torch.int32) for p in range(world)]
                self.mode = "push"
            except Exception as exc:      # no peer access / API not there: one collective at the end instead
                if exchange == "push":
                    raise
                self.push_error = repr(exc)
        if self.mode != "push":
            self.gathered = torch.zeros(world * L.total, dtype=torch.int32, device=self.dev)
        # this rank's buffer IS its region of the gathered buffer: kernels write the rows in place
        self.buf = self.gathered[self.rank * L.total:(self.rank + 1) * L.total]
        self.buf.zero_()
        self.cell_n = torch.zeros(max(self.n_cells, 1), dtype=torch.int32, device=self.dev)
        self.rows = self.buf[L.rows_off:L.rows_off + L.slots * L.W].view(L.slots, L.W)
        self.blk = self.buf[L.blk_off:L.blk_off + 3 * L.NB].view(3, L.NB)
        self.slots_after = torch.zeros(L.n_slabs + 1, dtype=torch.int64, device=self.dev)
        self.compute = torch.cuda.current_stream(self.dev)
        self.copy = torch.cuda.Stream(self.dev)
        # one stream per peer: the pushes of a slab to different GPUs run on different copy engines at the same time
        self.peer_streams = {p: torch.cuda.Stream(self.dev) for p in self.targets} if self.mode == "push" else {}
        self.fence = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.plan.set_compact_tables(self.blk[0], self.blk[1], self.blk[2])      # count_kernel writes them into the flat buffer
        self.slab_events = [torch.cuda.Event() for _ in range(L.n_slabs)]
        self.host = None

    # -- one pass over this rank's cells ------------------------------------------------------------------------
    def launch(self, seed=None, to_host=False):
        """Enqueue every slab (count + passage kernels) and its export (peer pushes and / or the copy to pinned host memory)."""
        torch, L = self.torch, self.L
        if seed is not None:
            self.plan.set_seed(seed)
        nC = self.n_cells
        cs = self.compute.cuda_stream
        if to_host and self.host is None:
            self.host = torch.empty(L.total, dtype=torch.int32).pin_memory()
        self.copy.wait_stream(self.compute)               # the previous pass's exports are ordered before this pass's kernels
        self.compute.wait_stream(self.copy)
        for ps in self.peer_streams.values():
            self.compute.wait_stream(ps)
        for k in range(L.n_slabs):
            lo, hi = min(nC, L.cell_lo[k]), min(nC, L.cell_lo[k + 1])
            if lo >= hi:
                break
            self.plan.launch_compact(self.cell_n, self.rows, lo, hi, first_slot=L.slot_lo[k],
                                     slots_after=self.slots_after[k:k + 1], stream=cs)
            self.slab_events[k].record(self.compute)
            w0, w1 = L.slab_words(k)
            if self.mode == "push":
                base = self.rank * L.total
                for p, ps in self.peer_streams.items():
                    with torch.cuda.stream(ps):
                        ps.wait_event(self.slab_events[k])
                        self.peers[p][base + w0:base + w1].copy_(self.buf[w0:w1], non_blocking=True)
            if to_host:
                with torch.cuda.stream(self.copy):
                    self.copy.wait_event(self.slab_events[k])
                    self.host[w0:w1].copy_(self.buf[w0:w1], non_blocking=True)

    def exchange(self, status_code=0, to_host=False):
        """Finish the pass: header + block tables to the peers, completion fence (push) or the one all-gather."""
        torch, dist, L = self.torch, self.dist, self.L
        hdr = torch.tensor([0, 0, int(status_code), self.n_cells], dtype=torch.int32)
        self.buf[2:4].copy_(hdr[2:4].to(self.dev, non_blocking=True), non_blocking=True)
        self.buf[0:2].copy_(self.slots_after[max(0, self.slabs_used() - 1):][:1].view(torch.int32)[:2], non_blocking=True)
        if self.world > 1:
            if self.mode == "push":
                base = self.rank * L.total
                for p, ps in self.peer_streams.items():
                    with torch.cuda.stream(ps):
                        ps.wait_stream(self.compute)
                        self.peers[p][base:base + L.rows_off].copy_(self.buf[:L.rows_off], non_blocking=True)
                self.copy.wait_stream(self.compute)
                with torch.cuda.stream(self.copy):
                    for ps in self.peer_streams.values():
                        self.copy.wait_stream(ps)
                    if to_host:
                        self.host[:L.rows_off].copy_(self.buf[:L.rows_off], non_blocking=True)
                    dist.all_reduce(self.fence, group=self.group)      # every rank's pushes are stream-ordered before its fence
                self.compute.wait_stream(self.copy)
            else:
                self.compute.wait_stream(self.copy)
                if to_host:
                    with torch.cuda.stream(self.copy):
                        self.copy.wait_stream(self.compute)
                        self.host[:L.rows_off].copy_(self.buf[:L.rows_off], non_blocking=True)
                mine = self.buf.clone() if dist.get_backend(self.group) != "nccl" else self.buf
                dist.all_gather_into_tensor(self.gathered, mine, group=self.group)
                self.compute.wait_stream(self.copy)
        elif to_host:
            with torch.cuda.stream(self.copy):
                self.copy.wait_stream(self.compute)
                self.host[:L.rows_off].copy_(self.buf[:L.rows_off], non_blocking=True)
            self.compute.wait_stream(self.copy)

    def step(self, seed=None, to_host=False):
        self.launch(seed, to_host)
        self.exchange(0, to_host)

    # -- results -----------------------------------------------------------------------------------------------------
    def results(self):
        """Per rank: (n_cells, status code, CompactResult) from the gathered buffer (device -> host copy of everything).  With
        ``gather_to="root"`` only rank 0 holds the other ranks' rows: the other ranks get None."""
        self.torch.cuda.synchronize(self.dev)
        if self.gather_to == "root" and self.rank != 0 and self.mode == "push":
            return None
        g = self.gathered.cpu().numpy().reshape(self.world, self.L.total)
        return [unpack_flat(self.L, g[r], self.fac.n_pairs) for r in range(self.world)]

    def slabs_used(self):
        """Slabs that hold cells of this rank (the last ones may be empty on a rank with fewer cells than ``n_cells_max``)."""
        return sum(1 for k in range(self.L.n_slabs) if self.L.cell_lo[k] < self.n_cells)

    def slots_used(self):
        """Rows this rank's last pass wrote (synchronises)."""
        self.torch.cuda.synchronize(self.dev)
        after = self.slots_after.cpu().numpy()
        return int(sum(int(after[k]) - self.L.slot_lo[k] for k in range(self.slabs_used())))

    def slab_overflow(self):
        """True when some slab used more slots than its region holds (the caller retries with larger regions)."""
        self.torch.cuda.synchronize(self.dev)
        after = self.slots_after.cpu().numpy()
        L = self.L
        return any(after[k] > L.slot_lo[k + 1] for k in range(self.slabs_used()))

    def close(self):
        self.plan.close()


def allreduce_tallies(tensors, group=None):
    """Sum dense tally tensors over ranks in place (small fixed-count projects; kept for callers of the dense layout)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def agree_on_status(code, device=None, group=None):
    """The largest-magnitude status code over the ranks (0 = every rank succeeded): an error on one rank must stop all of
    them BEFORE they enter a collective the failing rank will never reach."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(code)
    t = torch.tensor([-int(code)], dtype=torch.int32, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return -int(t.item())


def run_sharded(plan_host, rank, world, device, *, precision=32, seed=0, max_daily_fish=100000, group=None, gather_to="root"):
    """Simulate this rank's iterations of ``plan_host`` (a ``runner.RunPlan``) on ``device`` and exchange the rows.

    Returns a list with one (grid slice, CompactResult) per rank, in rank order — on rank 0, or on EVERY rank with
    ``gather_to="all"`` (the other ranks get None otherwise).  An engine error on any rank (population cap, non-finite count)
    is raised on all ranks with the failing rank's message; nobody hangs in a collective."""
    from . import _abi
    grids = [plan_host.grid.iterations_slice(r / world, (r + 1) / world) for r in range(world)]
    mine = grids[rank]
    n_max = max(g.n_cells for g in grids)
    species = plan_host.species_table
    cap_scale = 1.0
    while True:
        run = ShardedRun(plan_host.fac, plan_host.days, species, mine, rank, world, device, precision=precision, seed=seed,
                         max_daily_fish=max_daily_fish, n_cells_max=n_max, group=group, gather_to=gather_to,
                         slabs=slab_plan(grids, species, n_max, scale=cap_scale))
        code, msg = 0, ""
        try:
            run.launch(seed)
            run.plan.status()
        except _abi.StrykeError as exc:
            code, msg = exc.code, str(exc)
        over = 1 if (code == 0 and run.slab_overflow()) else 0
        worst = agree_on_status(code, run.dev if world > 1 and run.dist.get_backend(group) == "nccl" else None, group)
        if worst != 0:
            run.close()
            if code != 0:
                raise _abi.StrykeError(code, msg)
            raise _abi.StrykeError(worst, f"rank {rank}: another rank stopped with status {worst}")
        if agree_on_status(-over, run.dev if world > 1 and run.dist.get_backend(group) == "nccl" else None, group) != 0:
            run.close()
            cap_scale *= 2.0           # some slab region was too small somewhere: every rank retries with room to spare
            continue
        run.exchange(0)
        res = run.results()
        out = None if res is None else [(grids[r], x[2]) for r, x in enumerate(res)]
        run.close()
        return out
