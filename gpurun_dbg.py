import os, sys, time, tempfile, shutil
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as g
g.build()
from pycathy_wrapper_b200 import da, synthetic
from pycathy_wrapper_b200.capi import Simulation, load_library
from pycathy_wrapper_b200.project import load_project
lib = load_library()
prjs = []
for k in range(2):
    d = tempfile.mkdtemp()
    ks = 1.88e-4 * (1 + 0.3 * k)
    row = (ks, ks, ks, 1.0e-5, 0.55, 1.46, 0.15, 0.03125)
    synthetic.make_project(d, 100, 100, 15, ic=("wt", 1.0 + 0.1 * k), ISIMGR=1, DELTAT=10.0, DTMIN=1e-2, DTMAX=300.0, TMAX=1800.0, TIMPRT=[1800.0],
                           NODVP=[1], soil_rows=[row] * 15, atmbc=[(0.0, 5.0e-6), (1.0e9, 5.0e-6)])
    prjs.append(load_project(d))
t0 = time.time()
ens = da.Ensemble(lib, prjs, device=0)
print("build s", time.time() - t0)
for cyc in range(3):
    for j, s in enumerate(ens.sims):
        k = 0; t0 = time.time(); ms = 0; its = 0; lin = 0
        while True:
            r = s.step(); k += 1; ms += r.gpu_ms; its += r.iter; lin += r.pcg_iters
            if k <= 3 or r.finished: print("  cyc", cyc, "member", j, "step", r.nstep, "dt", r.deltat, "t", r.time, "iter", r.iter, "back", r.kbackt, "lin", r.pcg_iters, "ms", round(r.gpu_ms, 3))
            if r.finished: break
        print(" cyc", cyc, "member", j, "steps", k, "wall", time.time() - t0, "gpu ms", ms, "nl its", its, "lin its", lin)
    nn = prjs[0].nnod
    info = ens.analysis(np.linspace(0, nn - 1, 64).astype(np.int64), 0.55, np.full(64, 0.4), np.diag(np.full(64, 4e-4)))
    ens.restart(1800.0, 10.0)
