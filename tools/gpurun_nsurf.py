"""nsurf (routing sub-steps) and level structure of the coupled bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import __graft_entry__ as g
g.build()
from pycathy_wrapper_b200.capi import Simulation, load_library
prj = bench.make_workload((200, 200, 20), iopt=2, routing=True)
sim = Simulation(load_library(), prj)
print("steps nsurf:", [(r.nstep, r.nsurf, r.iter, round(r.gpu_ms, 2)) for r in (sim.step() for _ in range(8))])
