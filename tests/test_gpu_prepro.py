"""GPU: the device pre-processor (csrc/cathy_prepro.cu through cathy_prepro_run) against the golden files of the
reference's own ELF `pycppp`, against the oracle on DEMs no fixture holds, at the `./pycppp` process boundary, and --
at sizes the oracle cannot reach -- through the properties a drainage network has."""
import os
import shutil
import subprocess
import sys
import tarfile

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from test_prepro_oracle import CASES, golden_files, unpack

pytestmark = pytest.mark.gpu


def set_hap(text, **records):
    """Replace the value of the hap.in records whose description starts with the given keys."""
    lines = text.splitlines()
    for key, val in records.items():
        hit = [n for n, ln in enumerate(lines) if ln.startswith(key)]
        assert len(hit) == 1, key
        lines[hit[0]] = lines[hit[0]][:lines[hit[0]].index("=") + 1] + " " + val
    return "\n".join(lines) + "\n"


@pytest.fixture(scope="module")
def pp(gpu_lib):
    from pycathy_wrapper_b200 import preprocessor
    preprocessor.load_prepro_library()
    return preprocessor


@pytest.mark.parametrize("case", CASES)
def test_device_preprocessor_reproduces_pycppp_byte_for_byte(pp, case, tmp_path):
    d = unpack(case, tmp_path)
    out = str(tmp_path / "out")
    os.makedirs(out)
    shutil.copy(d + "/hap.in.orig", out + "/hap.in")
    shutil.copy(d + "/dtm_13.val", out + "/dtm_13.val")
    res = pp.run_preprocessor(out)
    assert res.info["n_launches"] >= 8 and res.info["n_waves"] > 1
    for f in golden_files(d):
        assert open(os.path.join(out, f)).read() == open(os.path.join(d, f)).read(), f


def test_device_depit_counts_the_reference_modifications(pp, tmp_path):
    """DEPIT's own counter: 102,325 raises in thousands of sweeps on the 16 x 14 fixture (pt = 1.3e-7), none on the plane."""
    d = unpack("deep_pits", tmp_path)
    res = pp.terrain_analysis(open(d + "/hap.in.orig").read(), open(d + "/dtm_13.val").read())
    assert res.info["n_modifications"] == 102325
    d = unpack("plane17", tmp_path)
    res = pp.terrain_analysis(open(d + "/hap.in.orig").read(), open(d + "/dtm_13.val").read())
    assert res.info["n_modifications"] == 0


def test_device_preprocessor_reproduces_the_committed_200x200_rasters(pp, tmp_path):
    """BASELINE config 3's routing inputs: tests/golden/route200_prepro.tar.xz was written by the reference ELF for the
    synthetic 200 x 200 DEM (tests/golden/make_route200.py); the device pre-processor must write the same 16 files."""
    from pycathy_wrapper_b200 import synthetic
    prj = str(tmp_path / "prj")
    synthetic.make_project(prj, 200, 200, 20, ic=("wt", 1.0), ISIMGR=2, TMAX=100.0, TIMPRT=[100.0], NODVP=[1])
    ref = str(tmp_path / "ref")
    os.makedirs(ref)
    with tarfile.open(os.path.join(GOLDEN, "route200_prepro.tar.xz")) as tf:
        tf.extractall(ref, filter="data")
    res = pp.run_preprocessor(os.path.join(prj, "prepro"))
    assert res.info["n_cells"] == 40000
    for f in sorted(os.listdir(ref)):
        assert open(os.path.join(prj, "prepro", f)).read() == open(os.path.join(ref, f)).read(), f


def test_device_preprocessor_matches_oracle_on_a_masked_rough_dem(pp, tmp_path):
    from oracle import prepro_oracle as po
    from pycathy_wrapper_b200 import synthetic
    rng = np.random.default_rng(5)
    nr, nc = 90, 70
    r, c = np.mgrid[0:nr, 0:nc]
    z = 8.0 - 0.03 * c - 0.021 * r + 0.015 * rng.standard_normal((nr, nc))
    z[(r - 45) ** 2 / 1.6 + (c - 35) ** 2 > 1150] = -9999.0
    z[40:44, 30:33] = -9999.0                                           # a hole inside the catchment
    for kw in ({}, {"Threshold on the contour curvature": "-0.100E+11", "Drainage directions method": "2",
                    "Upstream deviation memory": "0.100E+01", "Nondispersive channel flow": "1",
                    "Threshold on the support area": "0.500000000E+01"}):
        d = str(tmp_path / ("m%d" % len(kw)))
        os.makedirs(d)
        synthetic.write_hapin(d + "/hap.in", nr, nc, 0.5, 0.5)
        text = set_hap(open(d + "/hap.in").read(), **{"Depit threshold slope": "0.100E-02"}, **kw)
        open(d + "/hap.in", "w").write(text)
        np.savetxt(d + "/dtm_13.val", z, fmt="%.6f")
        o = po.Prepro(open(d + "/hap.in").read(), open(d + "/dtm_13.val").read()).run()
        g = pp.terrain_analysis(open(d + "/hap.in").read(), open(d + "/dtm_13.val").read())
        assert g.info["n_modifications"] == o.n_modifiche > 0
        assert list(g.order[:o.N_celle]) == o.qoi[1:]
        for a, b in (("p_outflow_1", "p1"), ("p_outflow_2", "p2"), ("hcID", "hcID"), ("dmID", "dmID")):
            assert np.array_equal(getattr(g, a), getattr(o, b)[1:]), a
        for a, b in (("quota", "quota"), ("A_inflow", "A_inflow"), ("w_1", "w_1"), ("w_2", "w_2"), ("local_slope_1", "ls_1"),
                     ("local_slope_2", "ls_2"), ("epl_1", "epl_1"), ("epl_2", "epl_2"), ("Ws1_sf_1", "Ws_1"), ("Ws1_sf_2", "Ws_2"),
                     ("kSs1_sf_1", "kSs_1"), ("kSs1_sf_2", "kSs_2"), ("b1_sf", "b1"), ("y1_sf", "y1"), ("nrc", "nrc")):
            x, y = getattr(g, a), np.asarray(getattr(o, b)[1:])
            pres = g.present
            assert np.array_equal(x[pres], y[pres]), a                    # bit-identical, doubles and singles alike
        assert abs(g.info["mean_s_max"] - o.mean_s_max) == 0.0


def test_pycppp_launcher_at_the_process_boundary(pp, tmp_path):
    """pyCATHY's own call: subprocess.run(["./pycppp"], input="2\\n0\\n1\\n") with cwd = <project>/prepro."""
    d = unpack("mask", tmp_path)
    run = str(tmp_path / "prepro")
    os.makedirs(run)
    shutil.copy(d + "/hap.in.orig", run + "/hap.in")
    shutil.copy(d + "/dtm_13.val", run + "/dtm_13.val")
    shutil.copy(os.path.join(ROOT, "pycathy_wrapper_b200", "pycppp"), run + "/pycppp")
    env = dict(os.environ, CATHY_B200_HOME=ROOT)
    env["PATH"] = os.path.dirname(sys.executable) + os.pathsep + env.get("PATH", "")
    p = subprocess.run(["./pycppp"], cwd=run, env=env, input="2\n0\n1\n", text=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "mrbb completed" in p.stdout
    for f in golden_files(d):
        assert open(os.path.join(run, f)).read() == open(os.path.join(d, f)).read(), f
    # the two fatal messages pyCATHY greps for in stdout (PY/cathy_tools.py:393-398)
    flat = run + "_flat"
    os.makedirs(flat)
    shutil.copy(d + "/hap.in.orig", flat + "/hap.in")
    shutil.copy(run + "/pycppp", flat + "/pycppp")
    np.savetxt(flat + "/dtm_13.val", np.full((36, 30), 2.0), fmt="%.3f")
    p = subprocess.run(["./pycppp"], cwd=flat, env=env, input="2\n0\n1\n", text=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode != 0 and "catchment with more than one outlet cell!" in p.stdout
    text = set_hap(open(flat + "/hap.in").read(), **{"Rivulet spacing": "0.300"})
    open(flat + "/hap.in", "w").write(text)
    p = subprocess.run(["./pycppp"], cwd=flat, env=env, input="2\n0\n1\n", text=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode != 0 and "DEM resolution is not a multiple of the rivulet spacing!" in p.stdout


def test_device_preprocessor_refuses_what_it_does_not_build(pp, tmp_path):
    d = unpack("bcc", tmp_path)
    with pytest.raises(pp.PreproError, match="boundary channel"):          # the reference's own check (wbb_sr.f90:144-158)
        pp.terrain_analysis(set_hap(open(d + "/hap.in.orig").read(), **{"Coefficient for boundary channel": "1.00"}), open(d + "/dtm_13.val").read())
    t = set_hap(open(d + "/hap.in.orig").read(), **{"Boundary channel constraction": "0"})
    t3 = set_hap(t, **{"Channel initiation method": "3"})
    with pytest.raises(pp.PreproError, match="nchc = 3"):
        pp.terrain_analysis(t3, open(d + "/dtm_13.val").read())
    from pycathy_wrapper_b200 import synthetic
    synthetic.write_hapin(str(tmp_path / "strip.in"), 1, 12, 0.5, 0.5)
    with pytest.raises(pp.PreproError, match="one cell wide"):
        pp.terrain_analysis(open(str(tmp_path / "strip.in")).read(), " ".join("%.3f" % (2.0 - 0.01 * k) for k in range(12)) + "\n")
    z = np.loadtxt(d + "/dtm_13.val")
    z[3, 3] = 0.0
    with pytest.raises(pp.PreproError, match="non-positive elevation"):
        pp.terrain_analysis(t, "\n".join(" ".join("%.6f" % v for v in row) for row in z) + "\n")


def test_drainage_network_properties_at_one_million_cells(pp, tmp_path):
    """1000 x 1000 (BASELINE config 5's DEM), far beyond the oracle: size-independent properties of the result."""
    from pycathy_wrapper_b200 import synthetic
    nr = nc = 1000
    z = synthetic.synthetic_dem(nr, nc)
    d = str(tmp_path / "big")
    os.makedirs(d)
    synthetic.write_hapin(d + "/hap.in", nr, nc, 0.5, 0.5)
    hap = open(d + "/hap.in").read()
    dtm = "\n".join(" ".join(repr(float(v)) for v in row) for row in z) + "\n"
    g = pp.terrain_analysis(hap, dtm)
    n = nr * nc
    order = g.order[:n].astype(np.int64) - 1
    assert g.info["n_cells"] == n and np.array_equal(np.sort(order), np.arange(n))          # a permutation
    q = g.quota
    assert np.all(np.diff(q[order]) <= 0)                                                     # descending elevation
    M = nr
    di = np.array([0, -1, -1, -1, 0, 0, 0, 1, 1, 1])
    dj = np.array([0, -1, 0, 1, -1, 0, 1, -1, 0, 1])
    outlet = order[-1]
    cells = np.arange(n)
    inner = cells != outlet
    w = g.w_1.astype(np.float64) + g.w_2.astype(np.float64)
    assert np.all(np.abs(w[inner] - 1.0) < 2e-7)                                              # the two weights share the cell's outflow
    for p, wk in ((g.p_outflow_1, g.w_1), (g.p_outflow_2, g.w_2)):
        use = inner & (wk > 0)
        rcv = cells[use] + M * di[p[use]] + dj[p[use]]
        assert np.all(q[rcv] < q[cells[use]])                                                 # water only runs downhill
    # every cell's area arrives at the outlet: upstream area of the outlet + its own cell = the catchment
    A_cell = 0.25
    assert abs(g.A_inflow[outlet] + A_cell - n * A_cell) < 1e-6 * n * A_cell
    # recompute A_inflow from the directions and weights in float64 (any order): agrees to rounding
    A = np.zeros(n)
    Aout = np.zeros(n)
    for c in order:                                                                           # descending: donors first
        Aout[c] = A[c] + A_cell
        if c == outlet:
            break
        for p, wk in ((g.p_outflow_1, g.w_1), (g.p_outflow_2, g.w_2)):
            if wk[c] > 0:
                A[c + M * di[p[c]] + dj[p[c]]] += Aout[c] * float(wk[c])
    assert np.max(np.abs(A - g.A_inflow) / np.maximum(A, 1.0)) < 1e-12
    print("1000x1000: device %.1f ms, %d waves, %d launches, stages %s" % (g.info["device_ms"], g.info["n_waves"], g.info["n_launches"], g.info["stage_ms"]))


def test_coupled_run_on_a_dem_preprocessed_on_the_device(pp, gpu_lib, oracle_mod, tmp_path):
    """What the pre-processor is for: a coupled surface / subsurface run (ISIMGR = 2) on a DEM no raster set was
    shipped for.  30 x 24 rough DEM -> device pre-processor -> the processor reads its files (project.py, SRC/datin.f:325-372) ->
    ponded storm routed over the drainage network, device against oracle step by step."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    from test_gpu_parity import _run_both, psi_close
    rng = np.random.default_rng(17)
    nr, nc = 30, 24
    r, c = np.mgrid[0:nr, 0:nc]
    dem = 2.0 - 0.02 * r - 0.011 * c + 0.003 * rng.standard_normal((nr, nc))
    for pr_, pc_ in ((7, 9), (15, 5), (22, 17)):
        dem[pr_, pc_] -= 0.06                                               # three one-cell pits for DEPIT to fill
    # the storm of the storm20 fixture (thin surface layers, rain on a saturated hillslope): well conditioned, unlike a ponded start on
    # a coarse top layer, where device and oracle part ways through rounding-level differences at the first BC switches
    d = synthetic.make_project(str(tmp_path / "prj"), nr, nc, 15, dem=dem, ic=("hydrostatic",), ISIMGR=2, TMAX=900.0, TIMPRT=[900.0], DELTAT=1.0, DTMIN=1e-4,
                               NODVP=[1], atmbc=[(0.0, 0.0), (60.0, 1.0e-4), (600.0, 1.0e-4), (660.0, 0.0), (1.0e9, 0.0)],
                               zratio=[0.002, 0.004, 0.006, 0.008, 0.01, 0.01, 0.02, 0.02, 0.05, 0.05, 0.1, 0.1, 0.2, 0.2, 0.22])
    text = set_hap(open(os.path.join(d, "prepro", "hap.in")).read(), **{"Depit threshold slope": "0.500E-03"})
    open(os.path.join(d, "prepro", "hap.in"), "w").write(text)
    res = pp.run_preprocessor(os.path.join(d, "prepro"))
    assert res.info["n_modifications"] == 3                                  # the processor gets the DEPITTED dem
    prj = load_project(d)
    assert np.allclose(np.asarray(prj.dem).reshape(nr, nc), res.north_first("quota"), rtol=2e-12, atol=0)      # the file carries 12 digits
    g, c, rg, rc = _run_both(gpu_lib, oracle_mod, prj)
    ok, dmax = psi_close(g.state()["psi"], c.state()["psi"])
    assert ok, dmax
    assert rg.nstep > 150 and rg.q_outlet_1 > 0
