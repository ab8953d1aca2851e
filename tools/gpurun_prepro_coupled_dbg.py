"""Debug: the coupled run of tests/test_gpu_prepro.py::test_coupled_run_on_a_dem_preprocessed_on_the_device, device against oracle,
printing the head / ponding difference step by step."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import oracle  # noqa: E402
from pycathy_wrapper_b200 import preprocessor as pp, synthetic  # noqa: E402
from pycathy_wrapper_b200.capi import Simulation, load_library  # noqa: E402
from pycathy_wrapper_b200.project import load_project  # noqa: E402
from test_gpu_prepro import set_hap  # noqa: E402

amp = float(os.environ.get("AMP", "0.02"))
rng = np.random.default_rng(17)
nr, nc = 30, 24
r, c = np.mgrid[0:nr, 0:nc]
dem = 2.0 - 0.02 * r - 0.011 * c + 0.003 * rng.standard_normal((nr, nc))
for pr_, pc_ in ((7, 9), (15, 5), (22, 17)):
    dem[pr_, pc_] -= 0.06
d = synthetic.make_project(tempfile.mkdtemp() + "/prj", nr, nc, 15, dem=dem, ic=("hydrostatic",), ISIMGR=2, TMAX=900.0, TIMPRT=[900.0], DELTAT=1.0, DTMIN=1e-4,
                           NODVP=[1], atmbc=[(0.0, 0.0), (60.0, 1.0e-4), (600.0, 1.0e-4), (660.0, 0.0), (1.0e9, 0.0)],
                           zratio=[0.002, 0.004, 0.006, 0.008, 0.01, 0.01, 0.02, 0.02, 0.05, 0.05, 0.1, 0.1, 0.2, 0.2, 0.22])
text = set_hap(open(os.path.join(d, "prepro", "hap.in")).read(), **{"Depit threshold slope": "0.500E-03"})
open(os.path.join(d, "prepro", "hap.in"), "w").write(text)
res = pp.run_preprocessor(os.path.join(d, "prepro"))
print("prepro", {k: v for k, v in res.info.items() if k not in ("hap_text", "stage_ms")})
prj = load_project(d)
oracle.load()
G, C = Simulation(load_library(), prj), oracle.simulation(prj)
k = 0
while True:
    rg, rc = G.step(), C.step()
    k += 1
    sg, sc = G.state(), C.state()
    dpsi = float(np.max(np.abs(sg["psi"] - sc["psi"])))
    keys = [x for x in ("pond", "pondnod", "ponding") if x in sg]
    dp = float(np.max(np.abs(sg[keys[0]] - sc[keys[0]]))) if keys else -1.0
    same = (rg.nstep, rg.iter, rg.kbackt, rg.nsurf) == (rc.nstep, rc.iter, rc.kbackt, rc.nsurf)
    if k <= 0:
        dd = np.abs(sg["pond"] - sc["pond"])
        top = np.argsort(dd)[::-1][:6]
        print("step", k, "nsurf", rg.nsurf, rc.nsurf, "pond diff top nodes", [(int(t), "%.3e" % dd[t], "%.6e" % sg["pond"][t], "%.6e" % sc["pond"][t]) for t in top])
        for t in top[:2]:
            irow, icol = divmod(int(t), nc + 1)          # node row from the north, node column
            print("   node", int(t), "row", irow, "col", icol, "cells around: ")
            for cr in (irow - 1, irow):
                for cc in (icol - 1, icol):
                    if 0 <= cr < nr and 0 <= cc < nc:
                        f = lambda name: res.north_first(name)[cr, cc]
                        print("      cell", cr, cc, "q %.6f p1 %d p2 %d w1 %.6f w2 %.6f ls1 %.4e ls2 %.4e hc %d" % (f("quota"), f("p_outflow_1"), f("p_outflow_2"), f("w_1"), f("w_2"), f("local_slope_1"), f("local_slope_2"), f("hcID")))
    if k % 20 == 0 or not same:
        print(k, "gpu", (rg.nstep, rg.iter, rg.kbackt, rg.nsurf), "ora", (rc.nstep, rc.iter, rc.kbackt, rc.nsurf), "dt %.5g" % rc.deltat,
              "dpsi %.3e dpond %.3e" % (dpsi, dp), "q_out %.6e %.6e" % (rg.q_outlet_1, rc.q_outlet_1), "ifatm differ", int(np.sum(sg["ifatm"] != sc["ifatm"])))
    if not same or rg.finished or k > 2000:
        break
print(list(sg.keys()))
