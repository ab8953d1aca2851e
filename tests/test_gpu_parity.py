"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Tolerances: pressure head 1e-6 relative or 1e-8 m absolute (BASELINE.json north_star); integer
quantities (accepted step sequence, nonlinear iteration counts, back-steps) exact."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def psi_close(a, b, rtol=1e-6, atol=1e-8):
    d = np.abs(a - b)
    return bool(np.all(d <= np.maximum(rtol * np.abs(b), atol))), float(d.max())


def full_from_upper(topol, ja, coef, n):
    import scipy.sparse as sp
    rows = np.repeat(np.arange(n), np.diff(topol))
    U = sp.csr_matrix((coef, (rows, ja - 1)), shape=(n, n))
    return (U + sp.triu(U, 1).T).tocsr()


@pytest.fixture(scope="module")
def weill():
    from pycathy_wrapper_b200.project import load_project
    return load_project(os.path.join(GOLDEN, "weill_exemple"))


def test_mesh_and_initial_storage(gpu_lib, oracle_mod, weill):
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, weill), oracle_mod.simulation(weill)
    assert (g.nnod, g.n, g.nt, g.nterm, g.nnz) == (c.nnod, c.n, c.nt, c.nterm, c.nnz) == (441, 7056, 36000, 52111, 97166)
    for a, b in zip(g.mesh(), c.mesh()):
        assert np.array_equal(a, b)
    assert abs(g.initial_storage() - c.initial_storage()) <= 1e-12 * c.initial_storage()


@pytest.mark.parametrize("dt", [0.02, 5.0, 100.0])
def test_assembled_system_matches_oracle(gpu_lib, oracle_mod, weill, dt):
    """PICUNS+ASSPIC+RHSPIC+CFMATP+RHSGRV+BCPIC: same CSR pattern (bit exact), values to 1e-12 relative."""
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, weill), oracle_mod.simulation(weill)
    tg, jg, ag, bg = g.debug_assemble(dt)
    tc, jc, ac, bc = c.debug_assemble(dt)
    assert np.array_equal(tg, tc) and np.array_equal(jg, jc)
    scale = np.abs(ac).max()
    big = ac > 1e80
    assert np.array_equal(big, ag > 1e80)
    assert np.max(np.abs(ag[~big] - ac[~big])) <= 1e-12 * np.abs(ac[~big]).max()
    assert np.max(np.abs(bg - bc)) <= 1e-11 * max(np.abs(bc).max(), 1e-30), (np.abs(bg - bc).max(), np.abs(bc).max())
    assert scale > 0


def test_spmv_matches_oracle(gpu_lib, oracle_mod, weill):
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, weill), oracle_mod.simulation(weill)
    g.debug_assemble(1.0)
    c.debug_assemble(1.0)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(g.n)
    x[:441] = 0.0                      # Dirichlet rows carry the 1.7e91 penalty: keep them out of the comparison
    yg, _ = g.debug_spmv(x)
    yc, _ = c.debug_spmv(x)
    assert np.max(np.abs(yg[441:] - yc[441:])) <= 1e-12 * np.abs(yc[441:]).max()


@pytest.mark.parametrize("dt", [0.02, 50.0])
def test_pcg_solution_matches_oracle(gpu_lib, oracle_mod, weill, dt):
    """SYMSLV: the device PCG (different preconditioner) must reach the same solution of the same system."""
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, weill), oracle_mod.simulation(weill)
    g.debug_assemble(dt)
    c.debug_assemble(dt)
    xg, ng, eg, _ = g.debug_solve()
    xc, nc, ec, _ = c.debug_solve()
    assert eg <= 1e-10 and ec <= 1e-10
    assert ng < 20 * 500
    assert np.max(np.abs(xg - xc)) <= 1e-7 * max(np.abs(xc).max(), 1e-12), (np.abs(xg - xc).max(), np.abs(xc).max(), ng, nc)


@pytest.mark.parametrize("size", [(6, 7, 4), (7, 6, 5), (7, 7, 3), (6, 6, 4), (20, 20, 15), (31, 30, 15), (60, 50, 8), (120, 120, 20)])
def test_pcg_kernel_variants_agree(gpu_lib, tmp_path, monkeypatch, size):
    """k_pcg (streaming), k_pcg_res (resident, one row per thread) and k_pcg_res2 (resident, paired rows: aligned 16-byte loads +
    lane shuffles) run the same recurrence: same iteration count, solutions equal to rounding.  The sizes cover the three possible
    parity combinations of the stencil offsets NC1, NNOD-NC1-1, NNOD-1 (template variants 6, 5, 3 of k_pcg_res2), odd and even row
    counts, CTAs without rows, and (last size) more than one pass of 2048 rows per CTA."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    nrow, ncol, nstr = size
    d = str(tmp_path / "prj")
    synthetic.make_project(d, nrow, ncol, nstr, ic=("wt", 0.6), ISIMGR=1, TMAX=100.0, TIMPRT=[100.0], NODVP=[1])
    prj = load_project(d)
    sols = {}
    for algo in (1, 3, 4):
        monkeypatch.setenv("CATHY_PCG_ALGO", str(algo))
        sim = Simulation(gpu_lib, prj)
        assert sim.solver_info()["kernel"] == algo
        sim.debug_assemble(7.0)
        sols[algo] = sim.debug_solve()[:3]
        sim.close()
    # the streaming kernels on the column-major permutation of the same system (k' = s L + l; default for meshes beyond the resident
    # kernels' reach and for partitioned runs): other stencil offsets, other summation order, same recurrence.  5 = k_pcg (direct
    # loads), 6 = k_pcg_tma (tiles staged by TMA bulk copies: windows, lower-triangle slices, two-stage mbarrier pipeline)
    monkeypatch.setenv("CATHY_PCG_ALGO", "1")
    monkeypatch.setenv("CATHY_PCG_CM", "1")
    for kern, tma in ((5, "0"), (6, "1")):
        monkeypatch.setenv("CATHY_PCG_TMA", tma)
        sim = Simulation(gpu_lib, prj)
        assert sim.solver_info()["kernel"] == kern
        sim.debug_assemble(7.0)
        sols[kern] = sim.debug_solve()[:3]
        sim.close()
    monkeypatch.delenv("CATHY_PCG_CM")
    monkeypatch.delenv("CATHY_PCG_TMA")
    monkeypatch.delenv("CATHY_PCG_ALGO")
    algos = [3, 4, 5, 6]
    if prj.n <= 16 * 1024:
        # small meshes, the default: k_pcg_cl -- one thread-block cluster, matrix + vectors in shared memory, cluster-scope reductions;
        # k_pcg_res2 inside one cluster (cluster barrier instead of the global-memory grid barrier) is the opt-in middle step
        # (8: k_pcg_cl2, one cluster barrier per iteration, when the stencil window fits; 7: k_pcg_cl; 9 here = k_pcg_res2 inside a cluster)
        for kern, env in ((8, {}), (7, {"CATHY_PCG_CL2": "0"}), (9, {"CATHY_PCG_CLUSTER": "4"})):
            for kk, vv in env.items():
                monkeypatch.setenv(kk, vv)
            sim = Simulation(gpu_lib, prj)
            info = sim.solver_info()
            if kern == 9:
                assert info["kernel"] == 4 and info["grid"] == 4, info
            else:
                assert info["kernel"] in ((7, 8) if kern == 8 else (7,)), info
            sim.debug_assemble(7.0)
            sols[kern] = sim.debug_solve()[:3]
            sim.close()
            for kk in env:
                monkeypatch.delenv(kk)
        algos += [7, 8, 9]
    x1, n1, e1 = sols[1]
    for algo in algos:
        x, nit, err = sols[algo]
        assert abs(nit - n1) <= 1 and err <= 1e-10
        assert np.max(np.abs(x - x1)) <= 1e-10 * max(np.abs(x1).max(), 1e-300), (algo, np.abs(x - x1).max(), np.abs(x1).max())


def test_config2_full_size_properties(gpu_lib, monkeypatch):
    """BASELINE config 2 at its full size (200 x 200 DEM x 20 layers, 848,421 nodes: too large for the oracle to finish in
    seconds), checked through size-independent properties: the assembled operator is symmetric and linear, the PCG solution
    satisfies the exported system (true residual recomputed on the host from the CSR the handle exports), the streaming and the
    resident PCG kernels drive the run through the same accepted steps to the same heads, saturation stays inside its bounds and the
    mass-balance error of every step stays below 1e-6 of the storage."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from pycathy_wrapper_b200.capi import Simulation
    prj = bench.make_workload((200, 200, 20))
    sim = Simulation(gpu_lib, prj)
    assert (sim.n, sim.nt, sim.nterm, sim.nnz) == (848421, 4800000, 6592841, 12337261)      # closed forms of SURVEY.md section 8
    assert sim.solver_info()["kernel"] == 4
    topol, ja, coef, rhs = sim.debug_assemble(1.0)
    n = sim.n
    rows = np.repeat(np.arange(n), np.diff(topol))
    import scipy.sparse as sp
    U = sp.csr_matrix((coef, (rows, ja - 1)), shape=(n, n))
    A = (U + sp.triu(U, 1).T).tocsr()
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    ax, _ = sim.debug_spmv(x)
    ay, _ = sim.debug_spmv(y)
    axy, _ = sim.debug_spmv(2.0 * x - 3.0 * y)
    scale = np.abs(ax).max()
    assert np.abs(ax - A @ x).max() <= 1e-12 * scale                       # the DIA operator is the exported CSR
    assert np.abs(axy - (2.0 * ax - 3.0 * ay)).max() <= 1e-12 * scale      # linearity
    assert abs(y @ ax - x @ ay) <= 1e-11 * abs(y @ ax)                     # symmetry
    sol, nit, err, _ = sim.debug_solve()
    assert err <= 1e-10 and 1 < nit < 500
    assert np.linalg.norm(rhs - A @ sol) <= 2e-10 * np.linalg.norm(rhs)    # true residual of the resident PCG's answer
    sim.close()
    runs = {}
    for algo in (1, 4):
        monkeypatch.setenv("CATHY_PCG_ALGO", str(algo))
        g = Simulation(gpu_lib, prj)
        seq = []
        for _ in range(6):
            r = g.step()
            seq.append((r.nstep, r.iter, r.kbackt, round(r.deltat, 12)))
            assert abs(r.erras) <= 1e-6 * abs(r.store1), (r.erras, r.store1)      # set by the nonlinear tolerance TOLUNS, observed 2e-7
        st = g.state()
        assert st["sw"].max() <= 1.0 + 1e-12 and st["sw"].min() >= 0.15 / 0.55 - 1e-12
        runs[algo] = (seq, st["psi"].copy())
        g.close()
    assert runs[1][0] == runs[4][0]
    ok, dmax = psi_close(runs[4][1], runs[1][1], rtol=1e-8, atol=1e-10)
    assert ok, dmax


def _bench_module():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    return bench


def test_config2_full_size_first_steps_match_oracle(gpu_lib, oracle_mod):
    """BASELINE config 2 at the size the headline number is quoted on (200 x 200 DEM x 20 layers, 848,421 nodes; bench.py's own
    workload): the first 6 accepted steps of the device against the CPU oracle (about 3 s of CPU per step) -- same
    (nstep, Picard iterations, back-steps, dt) sequence, storage, and heads inside the 1e-6 relative / 1e-8 m band.
    Here every one of the 148 persistent CTAs of k_pcg_res2 owns ~5.7 k rows."""
    prj = _bench_module().make_workload((200, 200, 20))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj, nsteps=6)
    assert g.solver_info()["kernel"] == 4 and rg.nstep == 6
    sg, sc = g.state(), c.state()
    ok, dmax = psi_close(sg["psi"], sc["psi"])
    assert ok, dmax
    assert np.max(np.abs(sg["sw"] - sc["sw"])) < 1e-6
    assert np.array_equal(sg["ifatm"], sc["ifatm"])


def test_config3_full_size_first_steps_match_oracle(gpu_lib, oracle_mod):
    """BASELINE config 3 at full size (same mesh, Newton + BiCGSTAB, surface routing from the first step; bench.py's `coupled`
    workload): first 3 accepted steps against the oracle's ILU(0)-BiCGSTAB (about 20 s of CPU per step) -- same step sequence,
    Newton iteration counts, routing sub-steps, heads inside the band, outlet discharge to 1e-6."""
    prj = _bench_module().make_workload((200, 200, 20), iopt=2, routing=True)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj, nsteps=3, store_rtol=1e-8)
    assert rg.nsurf > 0
    sg, sc = g.state(), c.state()
    ok, dmax = psi_close(sg["psi"], sc["psi"])
    assert ok, dmax
    assert np.array_equal(sg["ifatm"], sc["ifatm"])
    assert abs(rg.q_outlet_1 - rc.q_outlet_1) <= 1e-6 * max(abs(rc.q_outlet_1), 1e-12)


def test_mid_size_82k_full_run_matches_oracle_and_reference_elf(gpu_lib, oracle_mod, tmp_path):
    """The largest mesh a shipped reference ELF holds (100 x 50 DEM x 15 layers, 82,416 nodes): whole run on the device against
    the oracle step by step, and the final heads against the ELF's own psi output (tests/golden/mid82k, produced by
    make_golden.py mid82k; the oracle is byte-identical to it: test_oracle_golden.py::test_oracle_82k_mesh_reproduces_reference_elf)."""
    import importlib.util
    from pycathy_wrapper_b200.project import load_project
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    prj = load_project(mg.mid82k_project(str(tmp_path / "p")))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    gold = np.loadtxt(os.path.join(GOLDEN, "mid82k", "golden", "mbeconv"), skiprows=3, usecols=range(4))
    assert rg.nstep == int(gold[-1, 0]) and g.n == 82416
    sg, sc = g.state(), c.state()
    ok, dmax = psi_close(sg["psi"], sc["psi"])
    assert ok, dmax
    psi_ref = np.load(os.path.join(GOLDEN, "mid82k", "golden", "psi.npz"))["values"][-1]
    ok, dmax = psi_close(sg["psi"], psi_ref, rtol=2e-6, atol=1e-7)        # the ELF's file carries 7 significant digits
    assert ok, dmax


@pytest.mark.parametrize("case", ["weill", (6, 7, 4), (7, 6, 5), (2, 2, 1), (3, 9, 2), "zones"])
def test_assembly_with_derived_tet_indices_equals_stored_lists(gpu_lib, weill, tmp_path, monkeypatch, case):
    """k_assemble_a (tet indices = base(k) + per-class offset, the tables verified entry by entry at cathy_create) sums the same
    contributions in the same order as k_assemble (stored index lists): matrix and right-hand side are bit-identical."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    if case == "weill":
        prj = weill
    elif case == "zones":
        from test_oracle_golden import _zoned_project
        prj = load_project(_zoned_project(str(tmp_path / "z"), 1))
    else:
        nrow, ncol, nstr = case
        prj = load_project(synthetic.make_project(str(tmp_path / "p"), nrow, ncol, nstr, ic=("wt", 0.6), ISIMGR=1, TMAX=100.0, TIMPRT=[100.0], NODVP=[1]))
    out = {}
    for stored in (0, 1):
        if stored:
            monkeypatch.setenv("CATHY_PLAN_STORED", "1")
        sim = Simulation(gpu_lib, prj)
        assert sim.plan_info()["analytic"] == (not stored)
        sim.step()
        out[stored] = sim.debug_assemble(3.0)
        sim.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_newton_assembly_with_derived_tet_indices_equals_stored_lists(gpu_lib, newton20, monkeypatch):
    """k_assemble_newton<true> (derived tet indices, the tables of k_assemble_a) against k_assemble_newton<false> (stored lists):
    Jacobian (stiffness + derivative terms, upper and lower) and right-hand side bit-identical, with active derivative terms."""
    from pycathy_wrapper_b200.capi import Simulation
    out = {}
    for stored in (0, 1):
        if stored:
            monkeypatch.setenv("CATHY_PLAN_STORED", "1")
        sim = Simulation(gpu_lib, newton20)
        assert sim.plan_info()["analytic"] == (not stored)
        for _ in range(3):
            sim.step()
        out[stored] = sim.debug_assemble(2.0)
        sim.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_state_async_equals_state(gpu_lib, weill):
    """cathy_get_state_async + cathy_state_wait (snapshot drained by a second stream while the next step computes) returns
    exactly what the blocking cathy_get_state returns for the same step."""
    from pycathy_wrapper_b200.capi import Simulation
    g = Simulation(gpu_lib, weill)
    bufs = [g.state_buffers(pinned=True) for _ in range(2)]
    want = []
    for i in range(4):
        g.step()
        want.append({k: v.copy() for k, v in g.state().items()})
        g.state_async(bufs[i & 1])
        if i >= 1:      # the previous read-back completes while this step was computed; check it after the wait
            g.state_wait()
            for k, v in want[i].items():
                assert np.array_equal(bufs[i & 1][k], v), k
    g.state_wait()
    g.close()


def _run_both(gpu_lib, oracle_mod, prj, nsteps=None, store_rtol=1e-9, strict_attempts=True):
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    k = 0
    while True:
        rg, rc = g.step(), c.step()
        k += 1
        assert (rg.nstep, rg.iter, rg.kbackt, rg.nsurf) == (rc.nstep, rc.iter, rc.kbackt, rc.nsurf), \
            f"step {k}: gpu (nstep,iter,back,nsurf)={(rg.nstep, rg.iter, rg.kbackt, rg.nsurf)} oracle={(rc.nstep, rc.iter, rc.kbackt, rc.nsurf)}"
        assert abs(rg.deltat - rc.deltat) <= 1e-12 * rc.deltat and abs(rg.time - rc.time) <= 1e-12 * rc.time
        assert abs(rg.store1 - rc.store1) <= store_rtol * abs(rc.store1)
        assert abs(rg.erras) <= max(2.0 * abs(rc.erras), 1e-9 * abs(rc.store1)), (rg.erras, rc.erras)
        assert rg.finished == rc.finished
        if rc.kbackt > 0:       # the failed attempts as output/iter lists them (cathy_attempt_log): same attempts, same nonlinear iterations
            ag, ac = g.attempt_log(), c.attempt_log()
            assert len(ag) == len(ac) == rc.kbackt
            for (dg, tg, recg), (dc, tc, recc) in zip(ag, ac):
                assert abs(dg - dc) <= 1e-12 * dc and abs(tg - tc) <= 1e-12 * tc
                if not strict_attempts:
                    continue            # same back-step, possibly another way to fail (ITUNS exhausted instead of a linear solve that
                                        # gave up) and the norms of a diverging attempt are noise
                assert len(recg) == len(recc)
                # the last iteration of a failed attempt may sit on a linear solve that did not converge (LSFAIL): its norm is
                # whatever the two solvers left behind, so only the iterations before it are compared
                for a, b in zip(recg[:-1], recc[:-1]):
                    assert abs(a.pinf - b.pinf) <= 1e-4 * abs(b.pinf) + 1e-9, (a.pinf, b.pinf)
        if rg.finished or (nsteps and k >= nsteps):
            break
    return g, c, rg, rc


def test_hillslope_full_run_same_steps_and_heads(gpu_lib, oracle_mod, weill):
    """BASELINE config 1 end to end: 235 accepted steps incl. 8 back-steps and 1,793 routing sub-steps."""
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, weill)
    assert rg.nstep == 235
    sg, sc = g.state(), c.state()
    ok, dmax = psi_close(sg["psi"], sc["psi"])
    assert ok, dmax
    assert np.array_equal(sg["ifatm"], sc["ifatm"])
    assert np.max(np.abs(sg["sw"] - sc["sw"])) < 1e-6
    assert abs(rg.q_outlet_1 - rc.q_outlet_1) <= 1e-6 * max(abs(rc.q_outlet_1), 1e-12)
    gold = np.load(os.path.join(GOLDEN, "weill_exemple", "golden", "psi.npz"))
    ok, dmax = psi_close(sg["psi"], gold["values"][-1], rtol=2e-6, atol=1e-7)   # file carries 7 significant digits
    assert ok, dmax


def test_storm_full_run(gpu_lib, oracle_mod):
    from pycathy_wrapper_b200.project import load_project
    prj = load_project(os.path.join(GOLDEN, "storm20"))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert rg.q_outlet_1 > 0


def test_initial_ponding_with_routing(gpu_lib, oracle_mod, tmp_path):
    """IPOND = 1: 5 mm of water ponded on a saturated hillslope, routed from the first step on (the small version of bench.py's
    `coupled` workload; the oracle is pinned byte-for-byte against the ELF on this case in tests/test_oracle_golden.py).
    Routing rasters: the reference pre-processor's output for this 20 x 20 DEM, committed with the storm20 fixture."""
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import _ponded_project
    d = _ponded_project(str(tmp_path / "p"))
    src = os.path.join(GOLDEN, "storm20", "prepro")
    for f in os.listdir(src):
        if f.startswith("dtm_") or f == "qoi_a":
            shutil.copy(os.path.join(src, f), os.path.join(d, "prepro", f))
    prj = load_project(d)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert rg.nsurf > 0 and rg.q_outlet_1 > 0


def test_infiltration_subsurface_only(gpu_lib, oracle_mod, tmp_path):
    """ISIMGR=1 (SWITCH_OLD path), unsaturated start, rain pulse, geometric layers."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    p = synthetic.make_project(str(tmp_path / "inf"), 12, 17, 8, ic=("uniform", -1.0), TMAX=1200.0, TIMPRT=[1200.0],
                               atmbc=[(0.0, 0.0), (10.0, 2.0e-5), (600.0, 2.0e-5), (610.0, 0.0), (1e9, 0.0)])
    prj = load_project(p)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax


def _bc_project(tmp_path, name, neubc_text, with_dir=True, ic=("wt", 1.0)):
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    nrow, ncol, nstr = 10, 12, 6
    nnod = (nrow + 1) * (ncol + 1)
    # prescribed heads on the bottom nodes of the last surface row (explicit 3-D list), changing at t = 150 s
    nodes = [nstr * nnod + nrow * (ncol + 1) + j + 1 for j in range(ncol + 1)]
    dirbc = ("0.0 TIME\n0 %d\n%s\n%s\n150.0 TIME\n0 %d\n%s\n%s\n1e9 TIME\n0 0\n"
             % (len(nodes), " ".join(map(str, nodes)), " ".join(["1.2"] * len(nodes)),
                len(nodes), " ".join(map(str, nodes)), " ".join(["0.8"] * len(nodes))))
    p = synthetic.make_project(str(tmp_path / name), nrow, ncol, nstr, ic=ic, TMAX=400.0, TIMPRT=[400.0],
                               atmbc=[(0.0, 1.0e-5), (1e9, 1.0e-5)], dirbc_text=dirbc if with_dir else None, neubc_text=neubc_text)
    return load_project(p)


def test_dirichlet_and_neumann_nodes(gpu_lib, oracle_mod, tmp_path):
    """nansfdirbc / nansfneubc on the device: prescribed heads (two records), three injection nodes."""
    prj = _bc_project(tmp_path, "bc1", "0.0 TIME\n0 3\n200 310 420\n2.0e-6 -1.0e-6 3.0e-6\n1e9 TIME\n0 0\n")
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert abs(rg.ndin + rg.ndout - rc.ndin - rc.ndout) <= 1e-6 * max(abs(rc.ndin + rc.ndout), 1e-12)
    assert abs(rg.nnin - rc.nnin) <= 1e-12 and rc.nnin > 0


def test_free_drainage_bottom(gpu_lib, oracle_mod, tmp_path):
    """NODIN2 < 0: unit-gradient drainage at every bottom node (NEUMANN, SRC/neumann.f)."""
    prj = _bc_project(tmp_path, "bc2", "0.0 TIME\n-1 0\n1e9 TIME\n-1 0\n", with_dir=False, ic=("uniform", -0.5))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert rc.nnout < 0 and abs(rg.nnout - rc.nnout) <= 1e-6 * abs(rc.nnout)


def test_run_processor_writes_reference_format(gpu_lib, tmp_path):
    """Through the plugin-level call: output files parse like the reference's and match the golden ones."""
    from pycathy_wrapper_b200.processor import run_processor
    dst = str(tmp_path / "prj")
    shutil.copytree(os.path.join(GOLDEN, "weill_exemple"), dst)
    res = run_processor(dst)
    assert res.nstep == 235 and res.finished_ok
    got = np.loadtxt(os.path.join(dst, "output", "mbeconv"), skiprows=3)
    ref = np.loadtxt(os.path.join(GOLDEN, "weill_exemple", "golden", "mbeconv"), skiprows=3)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, [0, 3]], ref[:, [0, 3]])            # NSTEP, NLIN (nonlinear its) identical
    assert np.allclose(got[:, 1:3], ref[:, 1:3], rtol=1e-6)          # DELTAT, TIME
    assert np.allclose(got[:, 5], ref[:, 5], rtol=1e-6)              # STORE1
    with open(os.path.join(dst, "output", "psi")) as fh:
        assert fh.readline().endswith("NSTEP   TIME\n")


def test_missing_library_fails_loudly(monkeypatch):
    from pycathy_wrapper_b200 import capi
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "library_path", lambda: "/nonexistent/libcathy_b200.so")
    with pytest.raises(capi.CathyLibraryError):
        capi.load_library()


# ---------------------------------------------------------------------------------------------
# Newton scheme (IOPT = 2): SRC/newton.f on the device against the oracle (itself byte-identical to the reference's Newton ELF)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def newton20():
    from pycathy_wrapper_b200.project import load_project
    return load_project(os.path.join(GOLDEN, "newton20"))


@pytest.mark.parametrize("dt", [0.5, 20.0])
def test_newton_jacobian_matches_oracle(gpu_lib, oracle_mod, newton20, dt):
    """NEWUNS+ASSNEW+RHSNEW+CFMATN+RHSGRV+BCNEW: same full-row CSR pattern (bit exact), Jacobian and RHS to 1e-10 relative.
    Three warm-up steps first so that psi differs from the previous time level and every derivative term is active."""
    import scipy.sparse as sp
    from pycathy_wrapper_b200.capi import Simulation
    g, c = Simulation(gpu_lib, newton20), oracle_mod.simulation(newton20)
    for _ in range(3):
        g.step(); c.step()
    tg, jg, ag, bg = g.debug_assemble(dt)
    tc, jc, ac, bc = c.debug_assemble(dt)
    assert np.array_equal(tg, tc) and np.array_equal(jg, jc)
    big = ac > 1e80
    assert np.array_equal(big, ag > 1e80)
    assert np.max(np.abs(ag[~big] - ac[~big])) <= 1e-10 * np.abs(ac[~big]).max()
    assert np.max(np.abs(bg - bc)) <= 1e-9 * max(np.abs(bc).max(), 1e-30)
    n = g.n
    J = sp.csr_matrix((ac, jc - 1, tc - 1), shape=(n, n))
    assert abs(J - J.T).max() > 0            # the derivative terms make it nonsymmetric
    xg, ng, eg, _ = g.debug_solve()
    xc, nc, ec, _ = c.debug_solve()
    assert eg <= 1e-10
    assert np.max(np.abs(xg - xc)) <= 1e-7 * max(np.abs(xc).max(), 1e-12), (np.abs(xg - xc).max(), np.abs(xc).max(), ng, nc)


def test_newton_illconditioned_saturated_solve(gpu_lib, oracle_mod):
    """Hydrostatic, fully saturated start of the storm: the Newton increment is ~5e4 x the RHS (ill-conditioned); the device
    BiCGSTAB must still land on the oracle's ILU(0)-BiCGSTAB solution."""
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    prj = load_project(os.path.join(GOLDEN, "storm20n"))
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    tg, jg, ag, bg = g.debug_assemble(1.0)
    tc, jc, ac, bc = c.debug_assemble(1.0)
    assert np.array_equal(jg, jc)
    big = ac > 1e80
    assert np.max(np.abs(ag[~big] - ac[~big])) <= 1e-10 * np.abs(ac[~big]).max()
    assert np.max(np.abs(bg - bc)) <= 1e-9 * np.abs(bc).max()
    xg, ng, eg, _ = g.debug_solve()
    xc, nc, ec, _ = c.debug_solve()
    assert np.abs(xc).max() > 0.1
    assert np.max(np.abs(xg - xc)) <= 1e-6 * np.abs(xc).max(), (np.abs(xg - xc).max(), np.abs(xc).max(), ng, eg)


def test_newton_full_run_same_steps_and_heads(gpu_lib, oracle_mod):
    """Infiltration pulse under Newton: all 130 accepted steps, nonlinear iteration counts and heads as the reference ELF."""
    from pycathy_wrapper_b200.project import load_project
    prj = load_project(os.path.join(GOLDEN, "newton20"))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    assert rg.nstep == 130
    sg, sc = g.state(), c.state()
    ok, dmax = psi_close(sg["psi"], sc["psi"])
    assert ok, dmax
    assert np.array_equal(sg["ifatm"], sc["ifatm"])
    gold = np.load(os.path.join(GOLDEN, "newton20", "golden", "psi.npz"))
    ok, dmax = psi_close(sg["psi"], gold["values"][-1], rtol=2e-6, atol=1e-7)   # the ELF's file carries 7 significant digits
    assert ok, dmax


def test_newton_boustrophedon_sweeps(gpu_lib, oracle_mod, monkeypatch):
    """k_bicgstab with CATHY_BICG_ZIGZAG=1 (the second product and the s-update sweep from the last row to the first: default only
    when the Jacobian is larger than the L2, i.e. never at test sizes): the inner products are summed in another order, the
    runs must still follow the reference -- all 130 steps of the infiltration pulse and the first 150 of the coupled storm."""
    from pycathy_wrapper_b200.project import load_project
    monkeypatch.setenv("CATHY_BICG_ZIGZAG", "1")
    prj = load_project(os.path.join(GOLDEN, "newton20"))
    from pycathy_wrapper_b200.capi import Simulation
    probe = Simulation(gpu_lib, prj)
    assert probe.solver_info()["kernel"] == 11          # resident-vector solver (bicg_res.cuh), zigzag forced on
    probe.close()
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    assert rg.nstep == 130
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    prj = load_project(os.path.join(GOLDEN, "storm20n"))
    # with the sums taken in the other order one of the nine failed attempts of step 1 fails differently (the third linear solve scrapes
    # through and the attempt runs out of iterations instead): same back-steps, same accepted steps; the default order fails exactly
    # like the reference (test_newton_coupled_storm compares the attempts strictly)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj, nsteps=150, store_rtol=1e-8, strict_attempts=False)
    assert rg.nstep == 150 and g.state()["ifatm"].tolist() == c.state()["ifatm"].tolist()
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax


def test_newton_coupled_storm(gpu_lib, oracle_mod):
    """Newton + surface routing (BASELINE config 3 in small, 556 accepted steps in the reference): the first step back-steps 9 times
    in the reference because its third linear solve fails -- the device must fail there too -- and every following accepted step
    (ponding, routing sub-steps, atmospheric switching at every iteration) must match exactly: (NSTEP, Newton iterations,
    back-steps, routing sub-steps, DELTAT) are asserted for ALL steps up to 387.

    Step 388 is a knife-edge of the reference's own algorithm, not of the device solver: the reference restatement run with
    TOLCG halved leaves its own trajectory at exactly that step (tests/test_oracle_golden.py::
    test_oracle_coupled_storm_is_sensitive_at_step_388: 2 instead of 4 Newton iterations, 543 instead of 556 steps, heads 4e-3 m
    apart at the end), because with switching at every iteration a rounding-level perturbation grows ~10x every 40 steps
    (oracle against itself: 3.6e-9 m at step 150, 5e-7 at 300, 2.6e-4 at 380).  The device (vertical-line BiCGSTAB, residual
    1e-13) follows the same curve: heads inside the 1e-6 / 1e-8 m band at step 150 (measured 6e-9 m), 1.7e-6 m at step 300,
    8e-4 m at step 380, fork at 388.  After the fork only what any two runs of the reference itself share is asserted: same end
    time, same stored volume, heads within the reference's own sensitivity (5e-3 m)."""
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    prj = load_project(os.path.join(GOLDEN, "storm20n"))
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    bounds = {150: None, 300: 1e-5, 380: 3e-3}           # None: the parity band itself
    rg = rc = None
    for k in range(1, 388):
        rg, rc = g.step(), c.step()
        assert (rg.nstep, rg.iter, rg.kbackt, rg.nsurf) == (rc.nstep, rc.iter, rc.kbackt, rc.nsurf), \
            f"step {k}: gpu (nstep,iter,back,nsurf)={(rg.nstep, rg.iter, rg.kbackt, rg.nsurf)} oracle={(rc.nstep, rc.iter, rc.kbackt, rc.nsurf)}"
        assert abs(rg.deltat - rc.deltat) <= 1e-12 * rc.deltat and abs(rg.time - rc.time) <= 1e-12 * rc.time
        if k <= 300:
            assert abs(rg.store1 - rc.store1) <= 1e-8 * abs(rc.store1)
        if k in bounds:
            pg, pc = g.state()["psi"], c.state()["psi"]
            if bounds[k] is None:
                ok, dmax = psi_close(pg, pc)
                assert ok, dmax
                assert g.state()["ifatm"].tolist() == c.state()["ifatm"].tolist()
            else:
                assert np.max(np.abs(pg - pc)) <= bounds[k], (k, np.max(np.abs(pg - pc)))
    assert rg.nstep == 387 and rg.kback_total == rc.kback_total and rg.klsfai_total == rc.klsfai_total
    while not rg.finished:
        rg = g.step()
    while not rc.finished:
        rc = c.step()
    assert abs(rg.time - rc.time) <= 1e-9 * rc.time and rg.noback == rc.noback == 0
    assert abs(rg.nstep - rc.nstep) <= 20 and rc.nstep == 556
    assert abs(rg.store1 - rc.store1) <= 1e-6 * rc.store1
    assert np.max(np.abs(g.state()["psi"] - c.state()["psi"])) <= 5e-3


def test_vtk_and_velocities_against_reference(gpu_lib, oracle_mod, tmp_path):
    """Row (f)1 of SURVEY 8: vtk/1NN.vtk written by the device path.  Structure lines (header, points, cells) identical to the
    reference ELF's files; pressure / saturation within the head tolerance; VEL3D element velocities and VNOD3D nodal
    velocities within 1e-6 of the largest velocity (the device recomputes the basis coefficients with fused multiply-adds)."""
    import gzip
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.processor import run_processor
    from pycathy_wrapper_b200.project import load_project
    dst = str(tmp_path / "vtk6")
    shutil.copytree(os.path.join(GOLDEN, "vtk6"), dst)
    res = run_processor(dst, lib=gpu_lib)
    assert res.finished_ok
    for f in ("100.vtk", "101.vtk", "102.vtk", "103.vtk"):
        with gzip.open(os.path.join(GOLDEN, "vtk6", "golden", f + ".gz"), "rt") as fh:
            gold = fh.read().split("\n")
        ours = open(os.path.join(dst, "vtk", f)).read().split("\n")
        assert len(ours) == len(gold)
        ipd = gold.index([ln for ln in gold if ln.startswith("POINT_DATA")][0])
        assert ours[:ipd + 3] == gold[:ipd + 3]                     # header, TIME, POINTS, CELLS, CELL_TYPES: character for character
        num = lambda L: np.array([[float(v) for v in ln.split()] for ln in L if ln and (ln[0] == " " or ln[0] == "-")])
        n = int(gold[ipd].split()[1])
        pg, po = num(gold[ipd + 3:ipd + 3 + n]).ravel(), num(ours[ipd + 3:ipd + 3 + n]).ravel()
        assert np.all(np.abs(po - pg) <= np.maximum(1e-6 * np.abs(pg), 1e-8))
        iv = gold.index("VECTORS velocity float")
        vg, vo = num(gold[iv + 1:]), num(ours[iv + 1:])
        assert vg.shape == vo.shape and vg.shape[1] == 3
        assert np.max(np.abs(vo - vg)) <= 1e-6 * np.abs(vg).max()
    prj = load_project(dst)
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    for _ in range(5):
        g.step(); c.step()
    vg, vc = g.velocity(), c.velocity()
    for k in ("uu", "vv", "ww", "unod", "vnod", "wnod"):
        assert np.max(np.abs(vg[k] - vc[k])) <= 1e-6 * max(np.abs(vc[k]).max(), 1e-30), k


def test_auxiliary_outputs_against_reference(gpu_lib, tmp_path):
    """Device run of the vtk6 fixture through the `cathy` process boundary: the auxiliary files pyCATHY reads (hgatmsf,
    dtcoupling, hgsfdet, wtdepth, recharge, fort.777, psisurf, satsurf, swsurf, velnod, velelt) have the reference's layout
    line for line; numbers within 1e-5 relative of each column's scale (files carry 4-7 significant digits)."""
    import gzip
    from pycathy_wrapper_b200.processor import run_processor
    dst = str(tmp_path / "vtk6")
    shutil.copytree(os.path.join(GOLDEN, "vtk6"), dst)
    res = run_processor(dst, lib=gpu_lib)
    assert res.finished_ok

    def tokens(txt):
        out = []
        for ln in txt.split("\n"):
            if ln.startswith("#") or not ln.strip():
                out.append(ln if not ln.startswith("#Total") else "#Total")
                continue
            row = []
            for t in ln.split():
                try:
                    row.append(float(t))
                except ValueError:
                    row.append(t)
            out.append(row)
        return out

    for f in ("hgatmsf", "hgsfdet", "wtdepth", "recharge", "fort.777", "psisurf", "satsurf", "swsurf", "velnod", "velelt", "dtcoupling", "hgflag"):
        with gzip.open(os.path.join(GOLDEN, "vtk6", "golden", f + ".gz"), "rt") as fh:
            gold = tokens(fh.read())
        ours = tokens(open(os.path.join(dst, f if f == "fort.777" else os.path.join("output", f))).read())
        assert len(ours) == len(gold), f
        nums_g, nums_o = [], []
        for a, b in zip(ours, gold):
            if isinstance(b, str):
                assert a == b, (f, a, b)
                continue
            assert len(a) == len(b), (f, a, b)
            ncol = len(b) - (2 if f == "dtcoupling" else 0)            # last two dtcoupling columns are CPU seconds
            for u, v in zip(a[:ncol], b[:ncol]):
                if isinstance(v, str):
                    assert u == v, (f, u, v)
                else:
                    nums_o.append(u); nums_g.append(v)
        g, o = np.array(nums_g), np.array(nums_o)
        if g.size:
            assert np.max(np.abs(o - g)) <= 1e-5 * max(np.abs(g).max(), 1e-30), (f, np.abs(o - g).max(), np.abs(g).max())


@pytest.mark.parametrize("ivghu,xvg_case", [(1, 0), (1, 1), (2, None), (3, None), (4, None)])
def test_huyakorn_and_brooks_corey_curves(gpu_lib, oracle_mod, tmp_path, ivghu, xvg_case):
    """IVGHU = 1 (extended van Genuchten: SRC/fxvmc.f, fxvkr.f, fxvdmc.f, PNOT of SRC/chparm.f:36-78), 2, 3 (Huyakorn) and 4
    (Brooks-Corey) moisture curves (SRC/chpic0.f:37-99, SRC/chvelo.f, fhu*.f, fbc*.f): same accepted steps and heads as the oracle,
    which is byte-identical to the reference ELF on these cases
    (tests/test_oracle_golden.py::test_oracle_other_moisture_curves_against_reference_elf_when_available)."""
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import _curve_project
    d, nstep = _curve_project(str(tmp_path / "p"), ivghu, xvg_case)
    prj = load_project(d)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    assert rg.nstep == nstep
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert np.max(np.abs(g.state()["sw"] - c.state()["sw"])) < 1e-6


@pytest.mark.parametrize("ivghu,xvg_case", [(1, 0), (1, 1), (2, None), (3, None), (4, None)])
def test_other_moisture_curves_under_newton(gpu_lib, oracle_mod, tmp_path, ivghu, xvg_case):
    """CHNEW0's IVGHU = 1..4 branches (SRC/chnew0.f:39-89) in `k_curves_newton_alt`: Jacobian and RHS of a system with active
    derivative terms to 1e-10 / 1e-9, then the whole run with the accepted steps of the reference's Newton ELF (the oracle is byte-identical
    to it on these projects: test_oracle_other_moisture_curves_newton_against_reference_elf_when_available) and the oracle's heads."""
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import _curve_project
    d, nstep = _curve_project(str(tmp_path / "p"), ivghu, xvg_case, newton=True)
    prj = load_project(d)
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    for _ in range(12):
        g.step(); c.step()
    tg, jg, ag, bg = g.debug_assemble(5.0)
    tc, jc, ac, bc = c.debug_assemble(5.0)
    assert np.array_equal(tg, tc) and np.array_equal(jg, jc)
    big = ac > 1e80
    assert np.array_equal(big, ag > 1e80)
    assert np.max(np.abs(ag[~big] - ac[~big])) <= 1e-10 * np.abs(ac[~big]).max()
    assert np.max(np.abs(bg - bc)) <= 1e-9 * max(np.abs(bc).max(), 1e-30)
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    assert rg.nstep == nstep
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert np.max(np.abs(g.state()["sw"] - c.state()["sw"])) < 1e-6


def test_extended_van_genuchten_input_check(gpu_lib, tmp_path):
    """IVGHU = 1 with a specific storage above the curve's steepest slope (here: positive VGPSAT): cathy_create fails with the
    reference's CHPARM message instead of running on NaNs (SRC/chparm.f:48-52)."""
    from pycathy_wrapper_b200.capi import CathyLibraryError, Simulation
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import _xvg_bad_project
    prj = load_project(_xvg_bad_project(str(tmp_path / "p")))
    with pytest.raises(CathyLibraryError, match="DMCMAX"):
        Simulation(gpu_lib, prj)


@pytest.mark.parametrize("ivert", [0, 1, 2])
def test_soil_zones_and_ivert(gpu_lib, oracle_mod, tmp_path, ivert):
    """NZONE = 3 soil zones with per-(layer, zone) soils and IVERT = 0, 1, 2 (SRC/gen3d.f, SRC/tpnodi.f): same accepted steps and
    heads as the oracle, which is byte-identical to the reference's MAXZON=4 ELF on these very projects
    (tests/test_oracle_golden.py::test_oracle_soil_zones_and_ivert_against_reference_elf_when_available)."""
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import _zoned_project
    prj = load_project(_zoned_project(str(tmp_path / "p"), ivert))
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    assert rg.nstep == 44
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert np.max(np.abs(g.state()["sw"] - c.state()["sw"])) < 1e-6


def test_constant_relaxation(gpu_lib, oracle_mod, tmp_path):
    """NLRELX = 1 (SRC/relax.f): the relaxed heads enter the convergence norms and the next iterate; 276 accepted steps as the ELF."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    d = synthetic.make_project(str(tmp_path / "p"), 8, 9, 5, ic=("wt", 0.8), ISIMGR=1, TMAX=600.0, TIMPRT=[300.0, 600.0], DELTAT=1.0, DTMIN=1e-4,
                               DTMAX=50.0, NODVP=[4], NLRELX=1, OMEGA=0.7, atmbc=[(0.0, 0.0), (60.0, 2.0e-5), (1.0e9, 2.0e-5)])
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, load_project(d))
    assert rg.nstep == 276
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax


@pytest.mark.parametrize("case", ["faces_picard", "faces_newton", "seep9", "seep9n"])
def test_seepage_faces_match_oracle(gpu_lib, oracle_mod, tmp_path, case):
    """Seepage faces on the device (seepage.cuh; SRC/extall.f, sfinit.f, the SFEX branches of bcpic.f / bcnew.f, bkpic.f / bknew.f,
    fluxmb.f): two ten-node faces on the 20x20x15 hillslope (SFINIT, exit points moving up and down, ISFCVG = 1) under Picard and
    Newton, and the two one-node projects on which the oracle is pinned against the reference ELFs (tests/golden/seep9, seep9n;
    here without the ELFs' SFINIT quirk).  Same accepted steps, nonlinear iterations and back-steps; seepage flux SFFLW and volume
    VSFFLW of every step; heads 1e-6 / 1e-8 m."""
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import seepage_hillslope
    if case.startswith("faces"):
        prj = load_project(seepage_hillslope(str(tmp_path / "p"), iopt=2 if case.endswith("newton") else 1))
    else:
        prj = load_project(os.path.join(GOLDEN, case))
    g, c = Simulation(gpu_lib, prj), oracle_mod.simulation(prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])       # after SFINIT
    assert ok, dmax
    k, vsf, active, idle = 0, 0.0, 0, 0
    while True:
        rg, rc = g.step(), c.step()
        k += 1
        assert (rg.nstep, rg.iter, rg.kbackt) == (rc.nstep, rc.iter, rc.kbackt), f"step {k}: gpu {(rg.nstep, rg.iter, rg.kbackt)} oracle {(rc.nstep, rc.iter, rc.kbackt)}"
        assert abs(rg.deltat - rc.deltat) <= 1e-12 * rc.deltat
        assert abs(rg.sfflw - rc.sfflw) <= 1e-7 * abs(rc.sfflw) + 1e-16, (k, rg.sfflw, rc.sfflw)
        assert abs(rg.vsfflw - rc.vsfflw) <= 1e-7 * abs(rc.vsfflw) + 1e-16
        assert abs(rg.vout - rc.vout) <= 1e-7 * abs(rc.vout) + 1e-14
        vsf += rg.vsfflw
        active += rc.sfflw < 0.0
        idle += rc.sfflw == 0.0
        if rg.finished:
            assert rc.finished
            break
    assert vsf < 0.0 and active > 10
    if not case.startswith("faces"):
        assert idle > 10            # the one-node face was switched off as well
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    g.close()


@pytest.mark.parametrize("size", [(60, 50, 8), (33, 47, 5)])
def test_spmv_tma_matches_plain_product(gpu_lib, tmp_path, monkeypatch, size):
    """k_spmv_tma (the product staged by TMA bulk copies in the column-major permutation, what the HBM-resident SpMV figure of the bench
    line is measured with) against the plain gather kernel in the same numbering and in the reference numbering."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.project import load_project
    nrow, ncol, nstr = size
    prj = load_project(synthetic.make_project(str(tmp_path / "p"), nrow, ncol, nstr, ic=("wt", 0.6), ISIMGR=1, TMAX=100.0, TIMPRT=[100.0], NODVP=[1]))
    x = np.random.default_rng(5).standard_normal(prj.n)
    out = {}
    for name, env in (("layer", {"CATHY_PCG_ALGO": "1", "CATHY_PCG_CM": "0"}),
                      ("cm_plain", {"CATHY_PCG_ALGO": "1", "CATHY_PCG_CM": "1", "CATHY_PCG_TMA": "1", "CATHY_SPMV_PLAIN": "1"}),
                      ("cm_tma", {"CATHY_PCG_ALGO": "1", "CATHY_PCG_CM": "1", "CATHY_PCG_TMA": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sim = Simulation(gpu_lib, prj)
        if name != "layer":
            assert sim.solver_info()["kernel"] == 6
        sim.debug_assemble(5.0)
        out[name], _ = sim.debug_spmv(x, reps=2)
        sim.close()
        for k in env:
            monkeypatch.delenv(k)
    scale = np.abs(out["layer"]).max()
    assert np.max(np.abs(out["cm_tma"] - out["cm_plain"])) <= 1e-13 * scale
    assert np.max(np.abs(out["cm_tma"] - out["layer"])) <= 1e-12 * scale


@pytest.mark.parametrize("parm", [dict(KSLOPE=1, TOLKSL=0.01), dict(KSLOPE=2, TOLKSL=0.01), dict(NLRELX=2), dict(KSLOPE=1, TOLKSL=0.002, NLRELX=2),
                                  dict(KSLOPE=3, TOLKSL=0.01, PSEL=-0.7, PSER=5.0), dict(KSLOPE=4, PSEL=-0.6, PSER=-0.2)],
                         ids=["kslope1", "kslope2", "nlrelx2", "kslope1+nlrelx2", "kslope3", "kslope4"])
def test_chord_slopes_and_variable_relaxation(gpu_lib, oracle_mod, tmp_path, parm):
    """KSLOPE = 1, 2 (k_curves_chord + the PTOLD copies) and NLRELX = 2 (k_relxom_*: OMEGA formed on the device from the signed
    maximum head change) against the oracle, which is byte-identical to the ELF on the same projects (test_oracle_golden.py)."""
    from pycathy_wrapper_b200.project import load_project
    from test_oracle_golden import option_project
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, load_project(option_project(str(tmp_path / "p"), **parm)))
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert np.max(np.abs(g.state()["sw"] - c.state()["sw"])) < 1e-6
