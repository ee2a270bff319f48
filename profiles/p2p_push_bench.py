"""Push bandwidth of the exchange (torch symmetric memory, peer copies on side streams), idle and under the persistent passage
kernel: torchrun --nproc-per-node N profiles/p2p_push_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
n = 32 << 20          # int32 words: 128 MB per peer region
g = symm_mem.empty(world * n, dtype=torch.int32, device=f"cuda:{local}")
h = symm_mem.rendezvous(g, group=dist.group.WORLD)
peers = [h.get_buffer(p, (world * n,), torch.int32) for p in range(world)]
src = g[rank * n:(rank + 1) * n]
src.fill_(rank)
streams = {p: torch.cuda.Stream() for p in range(world) if p != rank}
print(rank, "peer tensor device", peers[(rank + 1) % world].device, flush=True)

def push():
    for p, s in streams.items():
        with torch.cuda.stream(s):
            peers[p][rank * n:(rank + 1) * n].copy_(src, non_blocking=True)

def timed(label, reps=10, busy=None):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if busy is not None:
            busy()
        push()
    for s in streams.values():
        s.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gb = reps * (world - 1) * n * 4 / 1e9
    if rank == 0:
        print(f"{label}: {gb / dt:.1f} GB/s egress per rank ({gb:.2f} GB in {dt * 1e3:.1f} ms)", flush=True)

timed("idle")
# under load: the C3 passage kernel on the compute stream
from stryke_b200 import engine, workloads as W
prj = W.c3_population(n_species=4, iterations=1000, days=200)
sim, plan = W.make_plan(prj, f"p2p{rank}")
p = engine.Plan(plan.fac, plan.days, plan.species_table, plan.grid, None, None, device=local, precision=32, seed=1, max_daily_fish=10 ** 9)
cn = torch.zeros(plan.n_cells, dtype=torch.int32, device=f"cuda:{local}")
rows = torch.zeros((plan.n_cells, plan.fac.n_pairs + 3), dtype=torch.int32, device=f"cuda:{local}")
def busy():
    p.launch_compact(cn, rows, 0, -1, first_slot=0)
busy(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(10): busy()
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"kernel alone: {e0.elapsed_time(e1) / 10:.2f} ms per pass", flush=True)
timed("under the passage kernel", busy=busy)
e0.record()
for _ in range(10): busy(); 
e1.record(); torch.cuda.synchronize()
dist.barrier(); dist.destroy_process_group()
