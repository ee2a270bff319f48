"""Multi-GPU check of the public sharded calls (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py

Every rank simulates its iterations of every (scenario, species) group into compact rows; after the exchange every rank must
hold exactly what a single GPU computes for the whole project (Philox is keyed by the cell, not by the rank) — with the rows
pushed slab by slab into the peers' buffers over NVLink (``push``) and with the one all-gather (``gather``); an engine error
on one rank must surface on all of them; ``simulation.run()`` under torchrun must write the single-GPU frames."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from stryke_b200 import _abi, engine, runner, workloads as W
    from stryke_b200.distributed import ShardedRun, run_sharded, unpack_flat
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def all_ok(ok, what):
        flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
        dist.all_reduce(flag)
        if int(flag.item()):
            raise SystemExit(f"{what}: failed on {int(flag.item())} rank(s)")

    for name, prj in (("c3", W.c3_population(n_species=5, iterations=30, days=20, scenarios=["A", "B"], facility_scenario_column=False)),
                      ("c5", W.c5_cascade(n_fish=40000, iterations=7, days=3))):
        sim, plan = W.make_plan(prj, f"mg_{name}_{rank}")
        pieces = run_sharded(plan, rank, world, device=local, seed=123, max_daily_fish=10 ** 7, gather_to="all")
        got = runner.dense_from_tallies(plan, runner.group_tallies(plan, pieces))
        whole = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.grid, None, None, device=local, precision=32, seed=123,
                                 max_daily_fish=10 ** 7)
        ok = (np.array_equal(got.cell_n, whole.cell_n) and np.array_equal(got.daily, whole.daily)
              and np.array_equal(got.state_daily, whole.state_daily))
        print(f"rank {rank}/{world} {name}: cells {plan.n_cells} (mine {pieces[rank][0].n_cells}), fish {int(whole.cell_n.sum())}, "
              f"sharded == single GPU: {ok}", flush=True)
        all_ok(ok, name)

    # both exchanges give the same gathered buffers, slab by slab (several slabs of 1024 cells on a 6 000-cell shard)
    prj = W.c3_population(n_species=4, iterations=300, days=10)
    sim, plan = W.make_plan(prj, f"mg_x_{rank}")
    grids = [plan.grid.iterations_slice(r / world, (r + 1) / world) for r in range(world)]
    bufs = {}
    for mode in ("push", "gather"):
        run = ShardedRun(plan.fac, plan.days, plan.species_table, grids[rank], rank, world, local, seed=9, max_daily_fish=10 ** 7,
                         slab_cells=1024, slab_cap=2200, n_cells_max=max(g.n_cells for g in grids), exchange=mode, gather_to="all")
        assert run.mode == mode, (run.mode, getattr(run, "push_error", None))
        for seed in (9, 10):                 # twice: a second pass reuses the buffers
            run.step(seed)
        assert not run.slab_overflow()
        bufs[mode] = run.gathered.cpu().numpy().copy()
        res = run.results()
        run.close()
    ok = np.array_equal(bufs["push"], bufs["gather"])
    pieces = [(grids[r], res[r][2]) for r in range(world)]
    got = runner.dense_from_tallies(plan, runner.group_tallies(plan, pieces))
    whole = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.grid, None, None, device=local, precision=32, seed=10,
                             max_daily_fish=10 ** 7)
    ok = ok and np.array_equal(got.cell_n, whole.cell_n) and np.array_equal(got.state_daily, whole.state_daily)
    print(f"rank {rank}: push == gather == single GPU: {ok}", flush=True)
    all_ok(ok, "exchange modes")

    # gather to the root only (what simulation.run() and bench.py use): rank 0 holds everybody's rows, the others only their own
    root = run_sharded(plan, rank, world, device=local, seed=10, max_daily_fish=10 ** 7, gather_to="root")
    ok = (root is None) if rank != 0 else True
    if rank == 0:
        got = runner.dense_from_tallies(plan, runner.group_tallies(plan, root))
        ok = np.array_equal(got.cell_n, whole.cell_n) and np.array_equal(got.state_daily, whole.state_daily) and np.array_equal(got.daily, whole.daily)
    print(f"rank {rank}: gather to root == single GPU: {ok}", flush=True)
    all_ok(ok, "gather to root")

    # an error on ONE rank (population cap exceeded in a cell only that rank owns) is raised on every rank, nobody hangs
    prj = W.c3_population(n_species=2, iterations=8 * world, days=30)
    sim, plan = W.make_plan(prj, f"mg_e_{rank}")
    whole = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.grid, None, None, device=local, precision=32, seed=77,
                             max_daily_fish=10 ** 7)
    cap = int(whole.cell_n.max()) - 1                 # exactly the largest cell(s) trip the cap
    raised = False
    try:
        run_sharded(plan, rank, world, device=local, seed=77, max_daily_fish=cap)
    except _abi.StrykeError as exc:
        raised = exc.code == _abi.E_MAX_FISH
    print(f"rank {rank}: cap error raised here: {raised}", flush=True)
    all_ok(raised, "error agreement")

    # the class API: sim.run() under torchrun shards by itself; rank 0 writes the frames a single GPU would write
    import shutil
    import tempfile
    from stryke_b200.store import open_store
    prj = W.c3_population(n_species=4, iterations=12, days=15)
    tmp = tempfile.mkdtemp(prefix=f"mg_run_{rank}_")
    sim = W.make_sim(prj, "sharded", proj_dir=tmp)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
    ok = 1
    if rank == 0:
        single = W.make_sim(prj, "single", proj_dir=tmp)
        single.device = local
        with contextlib.redirect_stdout(io.StringIO()):
            plan1 = runner.RunPlan(single)
        res1 = engine.run_cells(plan1.fac, plan1.days, plan1.species_table, plan1.grid, None, None,
                                device=local, precision=single.precision, seed=single.seed, max_daily_fish=10 ** 7)
        want = runner.assemble_frames(single, plan1, runner.group_tallies(plan1, [(plan1.grid, res1)]))
        with open_store(sim.hdf_path, mode="r") as st:
            ok = int(all(st[k].reset_index(drop=True).equals(want[k].reset_index(drop=True)) for k in ("Daily", "State_Daily", "Unit_Hours")))
        print(f"rank 0: sim.run() under torchrun wrote the single-GPU frames: {bool(ok)} ({len(want['Daily'])} Daily rows)", flush=True)
    all_ok(bool(ok), "sim.run() under torchrun")
    from stryke_b200.store import wait_for_exports
    wait_for_exports()                        # (rank 0's .h5 export runs on a worker thread)
    shutil.rmtree(tmp, ignore_errors=True)
    dist.destroy_process_group()
    if rank == 0:
        print("multi_gpu_check: all good", flush=True)


if __name__ == "__main__":
    main()
