import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle
from pycathy_wrapper_b200.capi import Simulation, load_library
from pycathy_wrapper_b200.project import load_project
prj = load_project('tests/golden/storm20')
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g, c = Simulation(load_library(), prj, tolcg_scale=tol), oracle.simulation(prj)
for k in range(1, 8):
    rg, rc = g.step(), c.step()
    print('step', k, 'gpu', (rg.nstep, rg.iter, rg.kbackt, rg.nsurf, rg.deltat), 'cpu', (rc.nstep, rc.iter, rc.kbackt, rc.nsurf, rc.deltat))
    for i in range(max(rg.n_iter_rec, rc.n_iter_rec)):
        a = rg.it[i] if i < rg.n_iter_rec else None
        b = rc.it[i] if i < rc.n_iter_rec else None
        print('   it', i+1, 'gpu', None if a is None else (a.niter, '%.6e'%a.pinf, a.ikmax, '%.6e'%a.fl2), 'cpu', None if b is None else (b.niter, '%.6e'%b.pinf, b.ikmax, '%.6e'%b.fl2))
    sg, sc = g.state(), c.state()
    print('   psi maxdiff %.3e  ifatm diff %d  atmact maxdiff %.3e pond maxdiff %.3e' % (np.abs(sg['psi']-sc['psi']).max(), int((sg['ifatm']!=sc['ifatm']).sum()), np.abs(sg['atmact']-sc['atmact']).max(), np.abs(sg['pond']-sc['pond']).max()))
    if (rg.iter != rc.iter): break
