"""Where a SMALL member spends its time (BASELINE config 1, 7,056 nodes): wall vs device time per accepted step for one handle,
then the aggregate rate of M members advancing concurrently (one host thread each).  usage: python tools/gpurun_small.py [M ...]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pycathy_wrapper_b200.capi import Simulation, load_library  # noqa: E402
from pycathy_wrapper_b200.project import load_project  # noqa: E402

lib = load_library()
prj = load_project(os.path.join(ROOT, "tests", "golden", "weill_exemple"))


def run(sim, nmax=10 ** 9):
    k, dev, launches, its = 0, 0.0, 0, 0
    global PCG
    while True:
        r = sim.step()
        k += 1
        dev += r.gpu_ms
        launches += r.launches
        its += r.iter
        PCG[0] += r.pcg_ms; PCG[1] += r.pcg_iters; PCG[2] += r.pcg_solves; PCG[3] += r.nsurf
        if r.finished or k >= nmax:
            return k, dev, launches, its


PCG = [0.0, 0, 0, 0]
s = Simulation(lib, prj)
run(s, 20)
s.close()
s = Simulation(lib, prj)
PCG = [0.0, 0, 0, 0]
t0 = time.perf_counter()
k, dev, launches, its = run(s)
w = time.perf_counter() - t0
print("single member: %d steps, %d nonlinear its, wall %.3f s (%.3f ms/step), device %.3f ms/step, %.1f launches/step, %.1f us wall per launch"
      % (k, its, w, 1e3 * w / k, dev / k, launches / k, 1e6 * w / launches), flush=True)
print("  linear solves: %.3f ms/step in the PCG kernel (%d solves, %d iterations, %.2f us/iteration incl. launch), routing sub-steps %d" % (PCG[0] / k, PCG[2], PCG[1], 1e3 * PCG[0] / max(PCG[1], 1), PCG[3]), flush=True)
s.close()
base = k / w
if os.environ.get("ONLY_SINGLE"):
    sys.exit(0)
for m in [int(v) for v in sys.argv[1:]] or [4, 8, 16, 32]:
    os.environ["CATHY_PCG_GRID"] = str(max(1, 148 // m))
    sims = [Simulation(lib, prj) for _ in range(m)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=m) as ex:
        res = list(ex.map(run, sims))
    w = time.perf_counter() - t0
    tot = sum(r[0] for r in res)
    print("%d concurrent members: %d member-steps in %.3f s = %.0f member-steps/s = %.2f x one member" % (m, tot, w, tot / w, tot / w / base), flush=True)
    for s in sims:
        s.close()
