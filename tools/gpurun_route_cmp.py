"""Config 1 (bundled hillslope, 1,793 routing sub-steps in 235 steps): time of the whole run and of routing with k_route alone
(CATHY_ROUTE_WAVE=0) and with the wavefront kernel for multi-sub-step calls (default); results must be identical."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g
g.build()
from pycathy_wrapper_b200.capi import Simulation, load_library
from pycathy_wrapper_b200.project import load_project
prj = load_project(os.path.join(ROOT, "tests", "golden", "weill_exemple"))
out = {}
for mode in ("0", "1"):
    os.environ["CATHY_ROUTE_WAVE"] = mode
    sim = Simulation(load_library(), prj)
    t0 = time.perf_counter(); ms = 0.0; ns = 0; seq = []
    while True:
        r = sim.step(); ms += r.gpu_ms; ns += r.nsurf; seq.append((r.nstep, r.iter, r.kbackt, r.nsurf))
        if r.finished: break
    out[mode] = (sim.state()["psi"].copy(), seq, r.q_outlet_1)
    print("CATHY_ROUTE_WAVE=%s: %d steps, %d sub-steps, device %.1f ms, wall %.2f s" % (mode, r.nstep, ns, ms, time.perf_counter() - t0))
    sim.close()
print("identical steps:", out["0"][1] == out["1"][1], "max |dpsi|:", np.abs(out["0"][0] - out["1"][0]).max(), "outlet:", out["0"][2], out["1"][2])
