import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from pycathy_wrapper_b200.capi import Simulation, load_library
lib = load_library()
for size in [(20,20,15),(200,200,20)]:
    prj = bench.make_workload(size)
    for block in (256, 512, 1024):
        for custom in (0, 1):
            os.environ['CATHY_PCG_BLOCK']=str(block); os.environ['CATHY_PCG_CUSTOM_BARRIER']=str(custom)
            sim = Simulation(lib, prj, tolcg_scale=1e-30, ITMXCG=50)
            sim.debug_assemble(10.0)
            sim.debug_solve()
            best=1e9
            for r in range(3):
                x, nit, err, ms = sim.debug_solve()
                best=min(best, ms)
            sim2 = Simulation(lib, prj)
            sim2.debug_assemble(10.0)
            xs, nit2, err2, ms2 = sim2.debug_solve()
            print(f"size {size} block {block} custom {custom}: {nit} its {best:.3f} ms -> {1e3*best/nit:.2f} us/iter ; converged solve: {nit2} its err {err2:.2e} {ms2:.3f} ms", flush=True)
            sim.close(); sim2.close()
