"""Multi-GPU check of the public sharded call (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py

Every rank simulates its block of cells into its slice of the full tally tensors; after the ONE all-reduce every rank
must hold exactly what a single GPU computes for the whole project (Philox is keyed by the cell, not by the rank)."""
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
    from oracle import fixtures
    import gpu_util as G
    from stryke_b200 import engine
    from stryke_b200.distributed import run_sharded, shard_bounds
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name, prj in (("c3", fixtures.c3_population(n_species=5, iterations=30, days=20)),
                      ("c5", fixtures.c5_cascade(n_fish=40000, iterations=7, days=3))):
        with contextlib.redirect_stdout(io.StringIO()):
            sim, plan = G.make_plan(prj, f"mg_{name}_{rank}")
        n, d, s = run_sharded(plan, rank, world, device=local, seed=123, max_daily_fish=10 ** 7)
        whole = engine.run_cells(plan.fac, plan.days, plan.species_table, plan.cell_day, plan.cell_species, plan.cell_id,
                                 device=local, precision=32, seed=123, max_daily_fish=10 ** 7)
        ok = (np.array_equal(n.cpu().numpy(), whole.cell_n) and np.array_equal(d.cpu().numpy().view(np.uint32), whole.daily)
              and np.array_equal(s.cpu().numpy().view(np.uint32), whole.state_daily))
        lo, hi = shard_bounds(plan.n_cells, rank, world)
        print(f"rank {rank}/{world} {name}: cells {plan.n_cells} (mine {lo}:{hi}), fish {int(whole.cell_n.sum())}, sharded == single GPU: {ok}", flush=True)
        flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
        dist.all_reduce(flag)
        if int(flag.item()):
            raise SystemExit(f"{name}: sharded result differs on {int(flag.item())} rank(s)")
    # the class API: sim.run() under torchrun shards by itself; rank 0 writes the frames a single GPU would write
    import shutil
    import tempfile
    import pandas as pd
    from stryke_b200 import runner
    from stryke_b200.store import open_store
    prj = fixtures.c3_population(n_species=4, iterations=12, days=15)
    tmp = tempfile.mkdtemp(prefix=f"mg_run_{rank}_")
    sim = G.make_sim(prj, "sharded", proj_dir=tmp)
    sim.device = local
    with contextlib.redirect_stdout(io.StringIO()):
        sim.run()
    ok = 1
    if rank == 0:
        single = G.make_sim(prj, "single", proj_dir=tmp)
        single.device = local
        plan1 = runner.RunPlan(single)
        res1 = engine.run_cells(plan1.fac, plan1.days, plan1.species_table, plan1.cell_day, plan1.cell_species, plan1.cell_id,
                                device=local, precision=single.precision, seed=single.seed, max_daily_fish=10 ** 7)
        want = runner.assemble_frames(single, plan1, res1)
        with open_store(sim.hdf_path, mode="r") as st:
            ok = int(all(st[k].reset_index(drop=True).equals(want[k].reset_index(drop=True)) for k in ("Daily", "State_Daily", "Unit_Hours")))
        print(f"rank 0: sim.run() under torchrun wrote the single-GPU frames: {bool(ok)} ({len(want['Daily'])} Daily rows)", flush=True)
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    shutil.rmtree(tmp, ignore_errors=True)
    if int(flag.item()):
        raise SystemExit("sim.run() under torchrun differs from the single-GPU frames")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
