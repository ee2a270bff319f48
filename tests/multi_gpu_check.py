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
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
