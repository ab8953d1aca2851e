import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ["CATHY_PCG_GRID"] = "49"
import __graft_entry__ as g
g.build()
sys.path.insert(0, "tests")
from test_gpu_partition import _project
from pycathy_wrapper_b200.capi import Simulation, load_library
from pycathy_wrapper_b200.partition import LocalPartition
lib = load_library()
prj = _project()
t0 = time.time()
part = LocalPartition(lib, prj, [0, 0])
print("created+started", time.time() - t0, part.infos, flush=True)
try:
    reps = part.step()
    print("step ok", [(r.nstep, r.iter, r.pcg_iters) for r in reps], time.time() - t0, flush=True)
    for _ in range(5):
        reps = part.step()
    print("6 steps ok", [(r.nstep, r.iter, r.pcg_iters, r.time) for r in reps], flush=True)
except Exception as e:
    print("FAILED", e, time.time() - t0, flush=True)
