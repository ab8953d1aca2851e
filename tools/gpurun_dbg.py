"""Where does the end-to-end loop lose time against the device-resident loop?  (bench.py's e2e leg, split into its parts)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pycathy_wrapper_b200.capi import Simulation, load_library
lib = load_library()
prj = bench.make_workload((200, 200, 20))
forcing = torch.from_numpy(np.ascontiguousarray(prj.atm_values[1]).copy()).pin_memory().numpy()
for mode in ("step", "step+upload", "step+state_async", "all", "all+blocking_state"):
    sim = Simulation(lib, prj)
    bufs = [sim.state_buffers(pinned=True) for _ in range(2)]
    for i in range(3):
        sim.step()
        if "async" in mode or mode == "all":
            sim.state_async(bufs[i & 1])
    sim.state_wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tp = []
    ts = []
    for i in range(20):
        if "upload" in mode or mode.startswith("all"):
            sim.upload_atm_record(1, forcing)
        b = time.perf_counter()
        sim.step()
        a = time.perf_counter()
        if mode == "all+blocking_state":
            sim.state(bufs[i & 1])
        elif "state_async" in mode or mode == "all":
            sim.state_async(bufs[i & 1])
        tp.append(time.perf_counter() - a)
        ts.append(a - b)
    sim.state_wait()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-22s %.3f ms/step   state call: mean %.3f max %.3f ms; step call mean %.3f ms; first 5 state calls %s" % (mode, 1e3 * dt / 20, 1e3 * np.mean(tp), 1e3 * np.max(tp), 1e3 * np.mean(ts), ["%.2f" % (1e3 * v) for v in tp[:5]]), flush=True)
    sim.close()
