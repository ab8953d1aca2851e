"""Row-block partition of ONE large mesh over several GPUs (BASELINE config 5).

The mesh is cut into strips of DEM rows; every rank holds its strip plus two ghost node rows per interior side with the same
layer-major numbering, so all kernels keep their 15-point DIA stencil.  Halo rows and reduction scalars travel through peer
memory over NVLink inside the kernels (csrc/cathy_b200.cu, "Row-block partition"); this module only wires the ranks up:

  * ``PartitionedSimulation``  one rank per process (torchrun), CUDA IPC handles gathered with torch.distributed
  * ``LocalPartition``         all ranks in one process, one host thread per rank (several GPUs, or one GPU shared)
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from .capi import CathyLibraryError, Simulation, partition_rows


def assemble_global(pieces, infos, nstr: int, nc1: int) -> np.ndarray:
    """Glue per-rank N-vectors (local numbering, ghosts included) into the global layer-major vector from the OWNED rows."""
    gnnod = infos[0]["nnod_global"]
    out = np.empty(gnnod * (nstr + 1))
    g = out.reshape(nstr + 1, gnnod)
    for v, inf in zip(pieces, infos):
        loc = np.asarray(v).reshape(nstr + 1, inf["nnod_local"])
        a, b, w0 = inf["own_row0"], inf["own_row1"], inf["win_row0"]
        g[:, a * nc1:b * nc1] = loc[:, (a - w0) * nc1:(b - w0) * nc1]
    return out


class PartitionedSimulation:
    """One rank of a partitioned run; ``torch.distributed`` must be initialised (NCCL, one process per GPU)."""

    def __init__(self, lib, prj, device: int, group=None, **kw):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.ranges = partition_rows(prj.nrow, self.world)
        a, b = self.ranges[self.rank]
        self.prj = prj
        self.sim = Simulation(lib, prj, device=device, dd=(self.world, self.rank, a, b), **kw)
        h = torch.tensor(list(self.sim.dd_export()), dtype=torch.uint8, device=torch.device("cuda", device))
        allh = [torch.empty_like(h) for _ in range(self.world)]
        dist.all_gather(allh, h, group=group)
        handles = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
        dist.barrier(group=group)                    # every box exists and is zeroed before anyone maps it
        self.sim.dd_connect(handles)                 # collective: finishes the set-up (initial sums are combined)
        self.info = self.sim.dd_info()
        self.n_global = self.info["n_global"]

    def step(self):
        return self.sim.step()

    def gather(self, key: str = "psi"):
        """Global vector on every rank (host); for tests and output writing, not for the hot path."""
        st = self.sim.state()[key]
        objs = [None] * self.world
        self.dist.all_gather_object(objs, (st, self.info), group=self.group)
        return assemble_global([o[0] for o in objs], [o[1] for o in objs], self.prj.nstr, self.prj.ncol + 1)

    def close(self):
        self.sim.close()


class LocalPartition:
    """All ranks inside this process: ``devices[r]`` is the GPU of rank r (repeat an ordinal to share one GPU -- then set the
    environment variable CATHY_PCG_GRID so that the persistent solver kernels of all ranks are co-resident)."""

    def __init__(self, lib, prj, devices, **kw):
        self.world = len(devices)
        self.prj = prj
        self.ranges = partition_rows(prj.nrow, self.world)
        self.sims = [Simulation(lib, prj, device=d, dd=(self.world, r, *self.ranges[r]), **kw) for r, d in enumerate(devices)]
        arr = (C.c_void_p * self.world)(*[s.h for s in self.sims])
        for s in self.sims:
            rc = lib.f["dd_connect_local"](s.h, arr)
            if rc != 0:
                raise CathyLibraryError(f"dd_connect_local failed ({rc}): {lib.error()}")
        self._parallel(lambda s: s._ck(lib.f["dd_start"](s.h), "dd_start"))
        self.infos = [s.dd_info() for s in self.sims]

    def _parallel(self, fn):
        out, err = [None] * self.world, [None] * self.world

        def run(r):
            try:
                out[r] = fn(self.sims[r])
            except Exception as e:      # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def step(self):
        """One accepted time step on every rank (they rendezvous inside the kernels); returns the per-rank reports."""
        return self._parallel(lambda s: s.step())

    def gather(self, key: str = "psi"):
        return assemble_global([s.state()[key] for s in self.sims], self.infos, self.prj.nstr, self.prj.ncol + 1)

    def close(self):
        for s in self.sims:
            s.close()
