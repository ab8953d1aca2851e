"""GPU: row-block partition of one mesh (BASELINE config 5 in small) against the unpartitioned run of the same project on the
same device: identical accepted steps, Picard and PCG iteration counts; heads equal to round-off (the partition only changes
the summation order of the dot products).  Two and three ranks share ONE GPU here (one host thread per rank, persistent solver
kernels sized to be co-resident); the multi-process NVLink path is exercised by bench.py --workload partitioned."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _project(nrow=24, ncol=10, nstr=6, tmax=400.0):
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    d = tempfile.mkdtemp(prefix="cathy_dd_")
    synthetic.make_project(d, nrow, ncol, nstr, ic=("wt", 0.8), ISIMGR=1, DELTAT=1.0, DTMAX=50.0, TMAX=tmax, TIMPRT=[tmax],
                           hspatm=0, atmbc=[(0.0, np.zeros((nrow + 1) * (ncol + 1))),
                                            (60.0, np.linspace(1.0e-5, 3.0e-5, (nrow + 1) * (ncol + 1))),
                                            (1.0e9, np.linspace(1.0e-5, 3.0e-5, (nrow + 1) * (ncol + 1)))])
    return load_project(d)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_run_equals_single_domain(gpu_lib, world):
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.partition import LocalPartition
    prj = _project()
    os.environ["CATHY_PCG_GRID"] = str(148 // (world + 1))      # world partitioned kernels + nothing else must be co-resident
    try:
        os.environ["CATHY_PCG_ALGO"] = "1"      # the unpartitioned reference runs the SAME kernel as the ranks (k_pcg_tma on the column-major
        os.environ["CATHY_PCG_CM"] = "1"        # permutation), so that PCG iteration counts are comparable one to one
        try:
            ref = Simulation(gpu_lib, prj)
            assert ref.solver_info()["kernel"] == 6
        finally:
            os.environ.pop("CATHY_PCG_ALGO", None)
            os.environ.pop("CATHY_PCG_CM", None)
        part = LocalPartition(gpu_lib, prj, [0] * world)
        assert all(s_.solver_info()["kernel"] == 6 for s_ in part.sims)
        assert sum(i["own_row1"] - i["own_row0"] for i in part.infos) == prj.nrow + 1
        k = 0
        while True:
            r = ref.step()
            reps = part.step()
            k += 1
            for q in reps:
                assert (q.nstep, q.iter, q.kbackt) == (r.nstep, r.iter, r.kbackt), (k, q.nstep, q.iter, r.iter)
                assert [q.it[i].niter for i in range(q.n_iter_rec)] == [r.it[i].niter for i in range(r.n_iter_rec)]
                assert abs(q.deltat - r.deltat) <= 1e-12 * r.deltat
                assert abs(q.store1 - r.store1) <= 1e-11 * abs(r.store1)
                assert q.it[q.n_iter_rec - 1].ikmax == r.it[r.n_iter_rec - 1].ikmax       # GLOBAL node number of the max-norm change
                assert abs(q.erras - r.erras) <= 1e-9 * max(abs(r.vin) + abs(r.vout), 1e-30) + 1e-14
            assert all(q.store1 == reps[0].store1 and q.it[0].pinf == reps[0].it[0].pinf for q in reps)   # ranks agree bit for bit
            if r.finished:
                assert all(q.finished for q in reps)
                break
        pg, pr = part.gather("psi"), ref.state()["psi"]
        assert np.max(np.abs(pg - pr)) <= 1e-9 * np.abs(pr).max()
        assert np.max(np.abs(part.gather("sw") - ref.state()["sw"])) <= 1e-9
        part.close()
        ref.close()
    finally:
        os.environ.pop("CATHY_PCG_GRID", None)
