"""CPU, world_size 2 over gloo: the N>1 path of bench.py shards independent ensemble members over ranks with no
data-path collective; only the timing reduction (MAX over ranks) and barriers use the process group."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import oracle
    prj = bench.make_workload((6, 7, 4), member=rank)          # each rank: its own member (Ks differs)
    sim = oracle.simulation(prj)
    steps = 0
    for _ in range(3):
        sim.step()
        steps += 1
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max-over-ranks timing reduction
    units = torch.tensor([float(sim.n * steps)], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ks = torch.tensor([prj.soil["TABLE"][0, 0, 0]], dtype=torch.float64)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, ks)
    if rank == 0:
        out.put((float(t.item()), float(units.item()), [float(g.item()) for g in gathered], sim.n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_members():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tmax, units, ks, n = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0 and units == 2 * 3 * n
    assert ks[0] != ks[1]                                      # different members on different ranks
