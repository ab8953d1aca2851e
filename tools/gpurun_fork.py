"""Where does the device run of the coupled Newton storm (tests/golden/storm20n, 556 steps) leave the oracle's step sequence?
Prints the first step whose (nstep, iter, kbackt, nsurf) differs, the head difference at step 150 / at the fork / at the end."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from oracle import oracle  # noqa: E402
from pycathy_wrapper_b200.capi import Simulation, load_library  # noqa: E402
from pycathy_wrapper_b200.project import load_project  # noqa: E402

prj = load_project(os.path.join(ROOT, "tests", "golden", "storm20n"))
G, C = Simulation(load_library(), prj), oracle.simulation(prj)
fork = None
k = 0
while True:
    rg, rc = G.step(), C.step()
    k += 1
    same = (rg.nstep, rg.iter, rg.kbackt, rg.nsurf) == (rc.nstep, rc.iter, rc.kbackt, rc.nsurf) and abs(rg.deltat - rc.deltat) <= 1e-12 * rc.deltat
    if k in (150, 300, 380) or (not same and fork is None):
        d = np.abs(G.state()["psi"] - C.state()["psi"]).max()
        print("step", k, "same" if same else "FORK gpu %s oracle %s" % ((rg.nstep, rg.iter, rg.kbackt, rg.nsurf, rg.deltat), (rc.nstep, rc.iter, rc.kbackt, rc.nsurf, rc.deltat)), "max |dpsi| %.3e" % d, flush=True)
    if not same and fork is None:
        fork = k
        break
    if rg.finished:
        break
while not rg.finished:
    rg = G.step()
while not rc.finished:
    rc = C.step()
print("fork", fork, "end steps gpu", rg.nstep, "oracle", rc.nstep, "time", rg.time, rc.time, "store rel", abs(rg.store1 - rc.store1) / rc.store1,
      "max |dpsi| end %.3e" % np.abs(G.state()["psi"] - C.state()["psi"]).max())
