import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as g
g.build()
from pycathy_wrapper_b200.capi import Simulation, load_library
from pycathy_wrapper_b200.project import load_project
from oracle import oracle
prj = load_project("tests/golden/storm20n")
c = oracle.simulation(prj)
crec, cpsi150 = [], None
while True:
    r = c.step(); crec.append((r.nstep, r.iter, r.kbackt, r.nsurf, r.deltat))
    if r.nstep == 150: cpsi150 = c.state()["psi"].copy()
    if r.finished: break
cpsi = c.state()["psi"]
for scale in (1e-2, 1e-3, 1e-4):
    s = Simulation(load_library(), prj, tolcg_scale=scale)
    k, first, d150, lin = 0, None, None, 0
    t0 = time.time()
    while True:
        r = s.step(); k += 1; lin += r.pcg_iters
        if first is None and (k > len(crec) or (r.nstep, r.iter, r.kbackt, r.nsurf) != crec[k - 1][:4]): first = k
        if r.nstep == 150: d150 = np.abs(s.state()["psi"] - cpsi150).max()
        if r.finished: break
    print("scale", scale, "steps", k, "oracle steps", len(crec), "first divergence", first, "dmax@150", d150, "dmax end", np.abs(s.state()["psi"] - cpsi).max(), "lin its", lin, "wall", time.time() - t0)
