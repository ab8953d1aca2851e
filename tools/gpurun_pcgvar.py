import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pycathy_wrapper_b200.capi import Simulation, load_library
lib = load_library()
variants = [("algo1", {"CATHY_PCG_ALGO": "1"}), ("res", {"CATHY_PCG_ALGO": "3"}), ("res2", {"CATHY_PCG_ALGO": "4"}),
            ("res2_xglobal", {"CATHY_PCG_ALGO": "4", "CATHY_PCG_RES_X": "0"})]
for size in [(20, 20, 15), (21, 19, 7), (100, 100, 15), (200, 200, 20), (250, 250, 20)]:
    prj = bench.make_workload(size)
    sols = {}
    for name, env in variants:
        for k in ("CATHY_PCG_ALGO", "CATHY_PCG_RES_X", "CATHY_PCG_RES_PREFETCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        sim = Simulation(lib, prj, tolcg_scale=1e-30, ITMXCG=10)      # 200 iterations, never converges: pure per-iteration cost
        sim.debug_assemble(10.0)
        sim.debug_solve()
        best = 1e9
        for r in range(3):
            x, nit, err, ms = sim.debug_solve()
            best = min(best, ms)
        n = sim.n
        sim2 = Simulation(lib, prj)
        sim2.debug_assemble(10.0)
        xs, nit2, err2, ms2 = sim2.debug_solve()
        sols[name] = xs
        print(f"size {size} n {n} {name}: {nit} its {best:.3f} ms -> {1e3*best/nit:.2f} us/iter = {168.0*n/(best/nit*1e-3)/1e9:.0f} GB/s (168 B/row) ; converged: {nit2} its err {err2:.2e} {ms2:.3f} ms", flush=True)
        sim.close(); sim2.close()
    for nm in ("res", "res2", "res2_xglobal"):
        print("   max |x_algo1 - x_%s| / max|x| =" % nm, np.abs(sols["algo1"] - sols[nm]).max() / np.abs(sols["algo1"]).max(), flush=True)
