"""CPU: host logic of the row-block partition (strip ranges, gluing per-rank vectors into the global numbering), also across
two gloo ranks (the data path itself -- peer-memory halos inside the kernels -- needs GPUs: tests/test_gpu_partition.py)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_partition_rows_cover_and_balance():
    from pycathy_wrapper_b200.capi import partition_rows
    for nrow, world in [(20, 2), (1000, 8), (37, 5), (7, 1)]:
        rs = partition_rows(nrow, world)
        assert rs[0][0] == 0 and rs[-1][1] == nrow + 1
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1


def _infos(nrow, ncol, nstr, world, W=2):
    from pycathy_wrapper_b200.capi import partition_rows
    nc1 = ncol + 1
    out = []
    for a, b in partition_rows(nrow, world):
        lo, hi = max(0, a - W), min(nrow + 1, b + W)
        out.append({"win_row0": lo, "win_rows": hi - lo, "own_row0": a, "own_row1": b, "nnod_local": (hi - lo) * nc1,
                    "n_local": (hi - lo) * nc1 * (nstr + 1), "nnod_global": (nrow + 1) * nc1, "n_global": (nrow + 1) * nc1 * (nstr + 1)})
    return out


def test_assemble_global_from_windows():
    from pycathy_wrapper_b200.partition import assemble_global
    nrow, ncol, nstr, world = 11, 4, 3, 3
    nc1, gnnod = ncol + 1, (nrow + 1) * (ncol + 1)
    g = np.arange(gnnod * (nstr + 1), dtype=float)
    infos = _infos(nrow, ncol, nstr, world)
    pieces = []
    for inf in infos:
        loc = g.reshape(nstr + 1, gnnod)[:, inf["win_row0"] * nc1:(inf["win_row0"] + inf["win_rows"]) * nc1].copy()
        ghost = np.ones_like(loc, dtype=bool)
        ghost[:, (inf["own_row0"] - inf["win_row0"]) * nc1:(inf["own_row1"] - inf["win_row0"]) * nc1] = False
        loc[ghost] = -1.0                      # ghost rows must not leak into the global vector
        pieces.append(loc.ravel())
    assert np.array_equal(assemble_global(pieces, infos, nstr, nc1), g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pycathy_wrapper_b200.partition import assemble_global
    nrow, ncol, nstr = 9, 3, 2
    nc1, gnnod = ncol + 1, (nrow + 1) * (ncol + 1)
    infos = _infos(nrow, ncol, nstr, world)
    inf = infos[rank]
    g = np.arange(gnnod * (nstr + 1), dtype=float) * 0.5
    loc = g.reshape(nstr + 1, gnnod)[:, inf["win_row0"] * nc1:(inf["win_row0"] + inf["win_rows"]) * nc1].copy().ravel()
    objs = [None] * world
    dist.all_gather_object(objs, (loc, inf))           # what PartitionedSimulation.gather does
    out = assemble_global([o[0] for o in objs], [o[1] for o in objs], nstr, nc1)
    if rank == 0:
        q.put(bool(np.array_equal(out, g)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_across_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
