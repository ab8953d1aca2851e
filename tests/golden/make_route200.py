"""Surface-routing rasters of the 200 x 200 synthetic DEM (BASELINE config 3), made by the reference's own pre-processor
(oracle/_ref/bin/pycppp, staged by oracle/build_ref.sh from /root/reference/examples/SSHydro/weill_exemple/prepro/pycppp)
on the DEM that pycathy_wrapper_b200.synthetic.make_project writes.  Output: tests/golden/route200_prepro.tar.xz, the 16 files
that an ISIMGR=2 run reads from prepro/ (SRC/datin.f; list in synthetic.FNAMES).  Usage: python tests/golden/make_route200.py"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402
from pycathy_wrapper_b200 import synthetic  # noqa: E402

FILES = ["qoi_a", "dtm_w_1", "dtm_w_2", "dtm_p_outflow_1", "dtm_p_outflow_2", "dtm_local_slope_1", "dtm_local_slope_2", "dtm_epl_1", "dtm_epl_2",
         "dtm_kSs1_sf_1", "dtm_kSs1_sf_2", "dtm_Ws1_sf_1", "dtm_Ws1_sf_2", "dtm_b1_sf", "dtm_y1_sf", "dtm_nrc"]


def main():
    tmp = "/tmp/route200_prj"
    shutil.rmtree(tmp, ignore_errors=True)
    synthetic.make_project(tmp, 200, 200, 20, ic=("wt", 1.0), ISIMGR=2, TMAX=100.0, TIMPRT=[100.0], NODVP=[1])
    oracle.run_prepro(tmp, timeout=1500)
    out = os.path.join(HERE, "route200_prepro.tar.xz")
    subprocess.run(["tar", "--sort=name", "--mtime=2000-01-01", "--owner=0", "--group=0", "-cJf", out, *FILES], cwd=os.path.join(tmp, "prepro"), check=True)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
