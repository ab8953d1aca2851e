"""Generates the committed fixtures under tests/golden/ (run HERE, where /root/reference is mounted).

  weill_exemple/   inputs + prepro rasters of the reference's bundled hillslope (BASELINE config 1,
                   /root/reference/examples/SSHydro/weill_exemple) and the outputs COMMITTED in the
                   reference next to them (mbeconv, cumflowvol, hgraph, vp verbatim; psi/sw as .npz).
  seep9, seep9n/   one-node seepage face switching on and off (Picard / Newton ELF), see seepage()
  storm20/         a synthetic 20x20x15 storm on a saturated hillslope (rain -> ponding, runoff routing, back-steps);
                   prepro rasters from the reference's pre-processor ELF, outputs from the reference's
                   processor ELF (oracle/_ref), both run by this script.

Usage: python tests/golden/make_golden.py
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from pycathy_wrapper_b200 import synthetic  # noqa: E402

KEEP_PREPRO = ["dem", "zone", "lakes_map", "qoi_a", "dtm_w_1", "dtm_w_2", "dtm_p_outflow_1", "dtm_p_outflow_2",
               "dtm_local_slope_1", "dtm_local_slope_2", "dtm_epl_1", "dtm_epl_2", "dtm_kSs1_sf_1", "dtm_kSs1_sf_2",
               "dtm_Ws1_sf_1", "dtm_Ws1_sf_2", "dtm_b1_sf", "dtm_y1_sf", "dtm_nrc", "hap.in"]
VERBATIM = ["mbeconv", "cumflowvol", "hgraph", "vp", "iter"]
AUX = ["hgatmsf", "hgnansf", "hgsfdet", "hgnansfdirdet", "hgnansfneudet", "wtdepth", "recharge", "fort.777", "psisurf", "satsurf", "swsurf",
       "velnod", "velelt", "hgflag", "dtcoupling"]


def read_blocks(path):
    steps, times, blocks, cur = [], [], [], []
    with open(path) as fh:
        for ln in fh:
            if "NSTEP" in ln:
                if cur:
                    blocks.append(np.array(cur))
                    cur = []
                a = ln.split()
                steps.append(int(a[0]))
                times.append(float(a[1]))
            elif "HSPSW" in ln:
                continue
            else:
                cur.extend(float(v) for v in ln.split())
    blocks.append(np.array(cur))
    return np.array(steps), np.array(times), np.array(blocks)


def stash(src_prj, out_dir, dst):
    for sub in ("input", "prepro", "golden"):
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
    shutil.copy(os.path.join(src_prj, "cathy.fnames"), os.path.join(dst, "cathy.fnames"))
    for f in os.listdir(os.path.join(src_prj, "input")):
        shutil.copy(os.path.join(src_prj, "input", f), os.path.join(dst, "input", f))
    for f in KEEP_PREPRO:
        if os.path.exists(os.path.join(src_prj, "prepro", f)):
            shutil.copy(os.path.join(src_prj, "prepro", f), os.path.join(dst, "prepro", f))
    for f in VERBATIM:
        shutil.copy(os.path.join(out_dir, f), os.path.join(dst, "golden", f))
    for f in ("psi", "sw"):
        s, t, b = read_blocks(os.path.join(out_dir, f))
        np.savez_compressed(os.path.join(dst, "golden", f + ".npz"), nstep=s, time=t, values=b)
    for r, _d, fs in os.walk(dst):
        os.chmod(r, 0o755)
        for f in fs:
            os.chmod(os.path.join(r, f), 0o644)


def main():
    ref = "/root/reference/examples/SSHydro/weill_exemple"
    stash(ref, os.path.join(ref, "output"), os.path.join(HERE, "weill_exemple"))
    tmp = "/tmp/golden_storm20"
    shutil.rmtree(tmp, ignore_errors=True)
    synthetic.make_project(tmp, 20, 20, 15, ic=("hydrostatic",), ISIMGR=2, TMAX=1800.0, TIMPRT=[900.0, 1800.0], DELTAT=1.0, DTMIN=1e-4,
                           NODVP=[441], atmbc=[(0.0, 0.0), (60.0, 1.0e-4), (600.0, 1.0e-4), (660.0, 0.0), (1.0e9, 0.0)],
                           zratio=[0.002, 0.004, 0.006, 0.008, 0.01, 0.01, 0.02, 0.02, 0.05, 0.05, 0.1, 0.1, 0.2, 0.2, 0.22])
    oracle.run_prepro(tmp)
    oracle.run_reference(tmp, tmp + "_ref", "20x20x15")
    stash(tmp, os.path.join(tmp + "_ref", "output"), os.path.join(HERE, "storm20"))
    newton()
    print("golden fixtures written")


def newton():
    """Newton scheme (IOPT=2, ISOLV=0: ILU(0)+BiCGSTAB) fixtures from the reference ELF built with Newton storage
    (examples/SSHydro/weil_exemple_outputs_plot/cathy): an infiltration pulse (newton20) and the coupled storm with
    surface routing (storm20n).  The ELF prints NaN in mbeconv's storage columns on this path (its STORMB reads SWNEW,
    which only the Picard routines set); the step-sequence columns and psi/sw/vp are the pins."""
    tmp = "/tmp/golden_newton20"
    shutil.rmtree(tmp, ignore_errors=True)
    synthetic.make_project(tmp, 20, 20, 15, ic=("wt", 1.0), ISIMGR=1, IOPT=2, ISOLV=0, TMAX=400.0, TIMPRT=[200.0, 400.0], DELTAT=1.0,
                           DTMIN=1e-4, DTMAX=100.0, NODVP=[5],
                           atmbc=[(0.0, 0.0), (60.0, 2.0e-5), (1800.0, 2.0e-5), (1860.0, 0.0), (1.0e9, 0.0)])
    shutil.rmtree(tmp + "_ref", ignore_errors=True)
    oracle.run_reference(tmp, tmp + "_ref", "20x20x15_newton")
    stash(tmp, os.path.join(tmp + "_ref", "output"), os.path.join(HERE, "newton20"))
    tmp = "/tmp/golden_storm20n"
    shutil.rmtree(tmp, ignore_errors=True)
    synthetic.make_project(tmp, 20, 20, 15, ic=("hydrostatic",), ISIMGR=2, IOPT=2, ISOLV=0, TMAX=700.0, TIMPRT=[350.0, 700.0], DELTAT=1.0,
                           DTMIN=1e-4, NODVP=[441], atmbc=[(0.0, 0.0), (60.0, 1.0e-4), (600.0, 1.0e-4), (660.0, 0.0), (1.0e9, 0.0)],
                           zratio=[0.002, 0.004, 0.006, 0.008, 0.01, 0.01, 0.02, 0.02, 0.05, 0.05, 0.1, 0.1, 0.2, 0.2, 0.22])
    oracle.run_prepro(tmp)
    shutil.rmtree(tmp + "_ref", ignore_errors=True)
    oracle.run_reference(tmp, tmp + "_ref", "20x20x15_newton")
    stash(tmp, os.path.join(tmp + "_ref", "output"), os.path.join(HERE, "storm20n"))


def vtk():
    """vtk/1NN.vtk of the reference ELF (VTKF=4: pressure, saturation, element conductivity, element Darcy velocity printed
    list-directed with 17 significant digits) on a small 6x5x3 mesh; stored gzipped."""
    import gzip
    tmp = "/tmp/golden_vtk6"
    shutil.rmtree(tmp, ignore_errors=True)
    synthetic.make_project(tmp, 6, 5, 3, ic=("wt", 0.6), ISIMGR=1, TMAX=120.0, TIMPRT=[60.0, 120.0], DELTAT=1.0, DTMIN=1e-4, DTMAX=50.0,
                           VTKF=4, IPRT=4, NODVP=[3], atmbc=[(0.0, 0.0), (30.0, 3.0e-5), (1.0e9, 3.0e-5)])
    shutil.rmtree(tmp + "_ref", ignore_errors=True)
    oracle.run_reference(tmp, tmp + "_ref", "20x20x15")
    dst = os.path.join(HERE, "vtk6")
    stash(tmp, os.path.join(tmp + "_ref", "output"), dst)
    for f in AUX:            # auxiliary outputs of the same run (SRC/detout.f, SRC/cathy_main.f:3628-3695)
        src = os.path.join(tmp + "_ref", f if f == "fort.777" else os.path.join("output", f))
        with open(src, "rb") as fi, gzip.GzipFile(os.path.join(dst, "golden", f + ".gz"), "wb", mtime=0) as fo:
            fo.write(fi.read())
    for f in sorted(os.listdir(os.path.join(tmp + "_ref", "vtk"))):
        with open(os.path.join(tmp + "_ref", "vtk", f), "rb") as fi, gzip.GzipFile(os.path.join(dst, "golden", f + ".gz"), "wb", mtime=0) as fo:
            fo.write(fi.read())


def mid82k_project(path):
    """100 x 50 DEM x 15 layers = 82,416 nodes, 450,000 tets: the largest mesh a shipped reference ELF holds
    (examplesTmp/SSHydro/ERA5_ETp_spatially_from_weill/cathy).  Infiltration pulse on the synthetic
    hillslope with a water table 1 m deep -- the bench workload of BASELINE config 2 at a tenth of its size."""
    return synthetic.make_project(path, 100, 50, 15, ic=("wt", 1.0), ISIMGR=1, TMAX=300.0, TIMPRT=[150.0, 300.0], DELTAT=1.0, DTMIN=1e-2,
                                  DTMAX=100.0, NODVP=[5], atmbc=[(0.0, 0.0), (60.0, 2.0e-5), (1.0e9, 2.0e-5)])


def mid82k():
    """Outputs of the reference ELF on the 82,416-node project (about 90 s of CPU): mbeconv, vp verbatim, psi/sw as .npz.
    The inputs are NOT stored: tests regenerate them with mid82k_project (deterministic generator)."""
    tmp = "/tmp/golden_mid82k"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.rmtree(tmp + "_ref", ignore_errors=True)
    mid82k_project(tmp)
    oracle.run_reference(tmp, tmp + "_ref", "100x50x15", timeout=3600)
    dst = os.path.join(HERE, "mid82k", "golden")
    os.makedirs(dst, exist_ok=True)
    out = os.path.join(tmp + "_ref", "output")
    for f in ("mbeconv", "vp"):
        shutil.copy(os.path.join(out, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    for f in ("psi", "sw"):
        s, t, b = read_blocks(os.path.join(out, f))
        np.savez_compressed(os.path.join(dst, f + ".npz"), nstep=s, time=t, values=b)


def seepage_project(path, newton=False):
    """8 x 9 x 5 hillslope with ONE potential seepage-face node 0.82 m deep at the foot of the slope (what the shipped ELFs, built with
    NSFMAX = NNSFMX = 1, can hold) just below the initial water table, and from t = 300 s a prescribed head of -0.6 m at the node
    under it: the seepage node is switched on by EXTALL, drains, is switched off when its back-calculated flux turns positive, and
    so on (95 transitions in 298 accepted steps under Picard, with back-steps at the start)."""
    nnod = 10 * 9
    node = 2 * nnod + 10 * 8 + 5
    kw = dict(IOPT=2, ISOLV=0) if newton else {}
    return synthetic.make_project(path, 8, 9, 5, ic=("wt", 2.2), ISIMGR=1, TMAX=1800.0, TIMPRT=[900.0, 1800.0], DELTAT=1.0, DTMIN=1e-4,
                                  DTMAX=50.0, NODVP=[4], atmbc=[(0.0, 0.0), (60.0, 0.0), (900.0, 0.0), (960.0, -3e-5), (1.0e9, -3e-5)],
                                  ISFCVG=1, seepage_faces=[[node]],
                                  dirbc_text="0.0 TIME\n0 0\n300.0 TIME\n0 1\n%d\n-0.6\n1e9 TIME\n0 1\n%d\n-0.6\n" % (node + nnod, node + nnod), **kw)


def seepage():
    """Seepage-face fixtures from the reference ELFs (Picard and Newton builds).  The ELFs run SFINIT over zero faces and with SFCHEK
    false because SFVONE's sentinel store overruns NSFNOD (see ORACLE_SF_ELF_QUIRK in oracle/cathy_oracle.c); everything after the
    initialisation -- BCPIC/SHLPIC, BKPIC, FLUXMB, EXTALL, MASBAL's VSFFLW, hgsfdet -- is the reference's own arithmetic."""
    global VERBATIM
    keep = VERBATIM
    VERBATIM = keep + ["hgsfdet", "hgatmsf"]
    for name, newton, which in (("seep9", False, "20x20x15"), ("seep9n", True, "20x20x15_newton")):
        tmp = "/tmp/golden_" + name
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.rmtree(tmp + "_ref", ignore_errors=True)
        seepage_project(tmp, newton)
        oracle.run_reference(tmp, tmp + "_ref", which, timeout=600)
        stash(tmp, os.path.join(tmp + "_ref", "output"), os.path.join(HERE, name))
    VERBATIM = keep


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "seepage":
        seepage()
    elif len(sys.argv) > 1 and sys.argv[1] == "newton":
        newton()
    elif len(sys.argv) > 1 and sys.argv[1] == "mid82k":
        mid82k()
    elif len(sys.argv) > 1 and sys.argv[1] == "vtk":
        vtk()
    else:
        main()
