"""Golden vectors of the pre-processor: small DEMs run through the reference's own ELF `pycppp`
(oracle/_ref/bin/pycppp, staged by oracle/build_ref.sh from /root/reference/examples/SSHydro/weill_exemple/prepro/pycppp)
with the answers pyCATHY gives it ("2 0 1": GRASS header, nodata 0, HAP pointers; PY/cathy_tools.py:379).
Output: tests/golden/prepro/<case>.tar.xz = hap.in.orig + dtm_13.val (inputs) and every file MRBB_SR / HG / WPARFILE wrote.
Usage: python tests/golden/make_golden_prepro.py"""
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pycathy_wrapper_b200 import synthetic  # noqa: E402

OUT = ["hap.in", "qoi_a", "dem", "lakes_map", "zone", "dtm_w_1", "dtm_w_2", "dtm_p_outflow_1", "dtm_p_outflow_2", "dtm_A_inflow",
       "dtm_local_slope_1", "dtm_local_slope_2", "dtm_epl_1", "dtm_epl_2", "dtm_kSs1_sf_1", "dtm_kSs1_sf_2", "dtm_Ws1_sf_1",
       "dtm_Ws1_sf_2", "dtm_b1_sf", "dtm_y1_sf", "dtm_hcID", "dtm_q_output", "dtm_nrc"]
KEYS = {"pt": "Depit threshold slope", "imethod": "Drainage directions method", "lambda": "Upstream deviation memory",
        "cc": "Threshold on the contour curvature", "ndcf": "Nondispersive channel flow", "nchc": "Channel initiation method",
        "A_thr": "Threshold on the support area", "ASk_thr": "Threshold on the AS**k", "kas": "Exponent k",
        "vo": "Drainage direction of the outlet", "bcc": "Boundary channel constraction", "cqm": "Coefficient for boundary",
        "cqg": "Coefficient for outlet"}


def hapin(path, nrow, ncol, dx, **kw):
    synthetic.write_hapin(path, nrow, ncol, dx, dx)
    t = open(path).read().splitlines()
    for k, v in kw.items():
        for n, ln in enumerate(t):
            if ln.startswith(KEYS[k]):
                t[n] = ln[:ln.index("=") + 1] + " " + str(v)
    open(path, "w").write("\n".join(t) + "\n")


def rough(rng, nr, nc, amp):
    r, c = np.mgrid[0:nr, 0:nc]
    return 5.0 - 0.01 * c - 0.025 * r + amp * rng.standard_normal((nr, nc))


def cases():
    rng = np.random.default_rng(7)
    yield "plane17", rough(rng, 30, 17, 0.0), {}
    yield "rough_lad", rough(rng, 24, 31, 0.01), {"pt": "0.100E-02"}
    yield "rough_ltd_pbm_d8", rough(rng, 33, 28, 0.02), {"pt": "0.100E-02", "cc": "-0.100E+11", "imethod": 2, "lambda": "0.100E+01"}
    yield "rough_pbm", rough(rng, 33, 28, 0.02), {"pt": "0.100E-02", "lambda": "0.100E+01"}
    yield "chan_ndcf", rough(rng, 40, 40, 0.02), {"pt": "0.100E-02", "A_thr": "0.200000000E+01", "ndcf": 1}
    yield "chan_ask", rough(rng, 40, 40, 0.02), {"pt": "0.100E-02", "nchc": 2, "ASk_thr": "0.01", "kas": "2.00"}
    z = rough(rng, 36, 30, 0.02)
    r, c = np.mgrid[0:36, 0:30]
    z[(r - 18) ** 2 / 1.3 + (c - 15) ** 2 > 190] = -9999.0
    yield "mask", z, {"pt": "0.100E-02"}
    yield "mask_d8", z, {"pt": "0.100E-02", "cc": "-0.100E+11", "A_thr": "0.100000000E+01"}
    yield "deep_pits", rough(rng, 16, 14, 0.02), {}                   # pt = 1.3e-7: thousands of DEPIT sweeps
    yield "bcc", rough(rng, 30, 30, 0.02), {"pt": "0.100E-02", "bcc": 1}
    # edge shapes (appended so that the random stream of the cases above is unchanged)
    yield "tiny_2x3", rough(rng, 2, 3, 0.0), {}
    # rasters one cell wide have no facet at all: the ELF then carries an UNINITIALISED channel flag from cell to cell
    # (PRE/dsf.f90:64,470; 32565 in dtm_hcID here) -- not a fixture; the draws stay so that `split` keeps its random stream
    rough(rng, 1, 12, 0.001)
    rough(rng, 9, 1, 0.02)
    z2 = rough(rng, 12, 10, 0.01)
    z2[:, 4] = -9999.0                                               # a null column splits the raster: two basins, one sort order
    z2[5, 4] = 4.7
    yield "split", z2, {"pt": "0.100E-02"}


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "pycppp")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "oracle", "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    os.makedirs(os.path.join(HERE, "prepro"), exist_ok=True)
    for tag, z, kw in cases():
        d = "/tmp/golden_prepro_" + tag
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        hapin(d + "/hap.in", z.shape[0], z.shape[1], 0.5, **kw)
        shutil.copy(d + "/hap.in", d + "/hap.in.orig")
        np.savetxt(d + "/dtm_13.val", z, fmt="%.6f", delimiter="\t")
        p = subprocess.run([exe], cwd=d, env=env, input="2\n0\n1\n", text=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        mods = [ln.split("=")[1].split("(")[0].strip() for ln in p.stdout.splitlines() if "(total)" in ln]
        if not os.path.exists(d + "/qoi_a"):
            raise SystemExit(tag + ": pycppp failed\n" + p.stdout[-1500:])
        out = os.path.join(HERE, "prepro", tag + ".tar.xz")
        subprocess.run(["tar", "--sort=name", "--mtime=2000-01-01", "--owner=0", "--group=0", "-cJf", out, "hap.in.orig", "dtm_13.val", *OUT],
                       cwd=d, check=True)
        print(tag, z.shape, "DEPIT modifications", mods, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
