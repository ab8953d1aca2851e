import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from pycathy_wrapper_b200.capi import Simulation, load_library
lib = load_library()
for size in [(20, 20, 15), (100, 100, 15), (200, 200, 20), (400, 400, 20)]:
    prj = bench.make_workload(size)
    sols = {}
    for algo in (1, 2):
        os.environ['CATHY_PCG_ALGO'] = str(algo)
        sim = Simulation(lib, prj, tolcg_scale=1e-30, ITMXCG=10)      # 200 iterations, never converges: pure per-iteration cost
        sim.debug_assemble(10.0)
        sim.debug_solve()
        best = 1e9
        for r in range(3):
            sim.debug_assemble(10.0)
            x, nit, err, ms = sim.debug_solve()
            best = min(best, ms)
        n = sim.n
        sim2 = Simulation(lib, prj)
        sim2.debug_assemble(10.0)
        xs, nit2, err2, ms2 = sim2.debug_solve()
        sols[algo] = xs
        bpr = 168.0 if algo == 1 else 144.0
        print(f"size {size} algo {algo}: {nit} its {best:.3f} ms -> {1e3*best/nit:.2f} us/iter = {bpr*n/(best/nit*1e-3)/1e9:.0f} GB/s ; converged: {nit2} its err {err2:.2e} {ms2:.3f} ms", flush=True)
        sim.close(); sim2.close()
    print("   max |x1-x2| / max|x| =", np.abs(sols[1] - sols[2]).max() / np.abs(sols[1]).max(), flush=True)
