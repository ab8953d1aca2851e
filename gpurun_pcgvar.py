import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from pycathy_wrapper_b200.capi import Simulation, load_library
lib = load_library()
for size in [(20, 20, 15), (100, 100, 15), (200, 200, 20), (400, 400, 20)]:
    prj = bench.make_workload(size)
    sols = {}
    for fs in (0, 1):
        os.environ['CATHY_PCG_FASTSYNC'] = str(fs)
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
        sols[fs] = xs
        print(f"size {size} fastsync {fs}: {nit} its {best:.3f} ms -> {1e3*best/nit:.2f} us/iter = {168.0*n/(best/nit*1e-3)/1e9:.0f} GB/s ; converged: {nit2} its err {err2:.2e} {ms2:.3f} ms", flush=True)
        sim.close(); sim2.close()
    print("   max |x0-x1| / max|x| =", np.abs(sols[0] - sols[1]).max() / np.abs(sols[0]).max(), flush=True)
