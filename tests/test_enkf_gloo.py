"""CPU, world_size 2 over gloo: the member-sharded analysis of pycathy_wrapper_b200.da.sharded_enkf_update (all_gather of
the predicted observations, all_reduce of row sums and of the partial cross covariance) reproduces the single-process
oracle.  The local stages are supplied by a numpy stand-in here (the CUDA stages need a GPU; they are checked against
the same oracle in tests/test_gpu_enkf.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


class NumpyOps:
    """numpy restatement of the four local stages of csrc/cathy_enkf.cu (test double)."""

    def gain(self, hx, y, R, sakov):
        m, ne = hx.shape
        S = hx - hx.sum(1, keepdims=True) * (1.0 / ne)
        D = (y if y.ndim == 2 else y[:, None]) - hx
        B = D / np.diag(R)[:, None] if sakov else np.linalg.solve(S @ S.T / (ne - 1) + R.T, D)
        return S, B

    def rowsum(self, X):
        return X.sum(dim=1)

    def crosscov(self, X, mean, S_local, ne_total):
        return (X - mean[:, None]) @ torch.from_numpy(S_local.T.copy()) / (ne_total - 1)

    def update(self, X, P, L, B_local, mean, bbar, inflate, n_infl, inflate2):
        PL = P.clone()
        if L is not None:
            PL[: L.shape[0]] *= L
        Xa = X + PL @ torch.from_numpy(B_local)
        ma = mean + PL @ torch.from_numpy(bbar)
        fac = torch.full((X.shape[0], 1), float(inflate2), dtype=torch.float64)
        fac[:n_infl] = inflate
        X.copy_(ma[:, None] + fac * (Xa - ma[:, None]))
        return X


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pycathy_wrapper_b200 import da
    g = np.load(os.path.join(GOLDEN, "enkf_golden.npz"))
    X = np.vstack([g["X"], g["theta"]])
    ne = X.shape[1]
    cut = 20                                  # ragged split: 20 + 12 members
    cols = slice(0, cut) if rank == 0 else slice(cut, ne)
    Xl = torch.from_numpy(np.ascontiguousarray(X[:, cols]))
    HXl = torch.from_numpy(np.ascontiguousarray(g["HX"][:, cols]))
    L = torch.from_numpy(g["L"])
    da.sharded_enkf_update(Xl, HXl, g["y"], g["R"], sakov=False, L=L, inflate=1.05, n_infl=g["X"].shape[0], inflate2=1.1, ops=NumpyOps())
    out = [torch.zeros((X.shape[0], cut)), torch.zeros((X.shape[0], cut))]
    pad = torch.zeros((X.shape[0], cut), dtype=torch.float64)
    pad[:, : Xl.shape[1]] = Xl
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank == 0:
        q.put(torch.cat([out[0], out[1][:, : ne - cut]], dim=1).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_matches_oracle():
    from oracle import enkf_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    Xa = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(GOLDEN, "enkf_golden.npz"))
    ns = g["X"].shape[0]
    assert np.allclose(Xa[:ns], g["loc_analysis"], rtol=1e-12, atol=1e-13)
    assert np.allclose(Xa[ns:], g["loc_param"], rtol=1e-12, atol=1e-13)
    r = o.enkf_analysis_localized_with_inflation(g["y"], g["R"], g["X"], g["theta"], g["HX"], L=g["L"], Sakov=False,
                                                 inflate_states=1.05, inflate_params=1.1)
    assert np.allclose(Xa[:ns], r[9], rtol=1e-12, atol=1e-13)
