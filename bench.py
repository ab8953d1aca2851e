#!/usr/bin/env python
"""bench.py -- headline benchmark of the CATHY Richards hot path on B200.

Metric (BASELINE.json): node-timesteps/s = N x accepted time steps / seconds of the time loop.
Headline workload (BASELINE.json configs[1]): synthetic 200x200 DEM x 20 layers (848,421 nodes,
4.8 M tetrahedra), van Genuchten, Picard + PCG, infiltration pulse, fp64.
A "step" is one ACCEPTED time step of the hot path (all its Picard iterations, linear solves,
mass balance, boundary switching and any back-stepped attempts).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size NROWxNCOLxNSTR]
                  [--workload picard|newton|coupled|enkf|partitioned|prepro] [--sharded auto|off|enkf|partitioned|enkf,partitioned]

The ONE JSON line of the default run carries, next to the headline `value` / `e2e` / `roofline` / `cpu_baseline`:
  * `sharded` -- the two shardings BASELINE.json's north_star names, measured in the SAME invocation at the same N:
      `enkf`        BASELINE config 4: 256 EnKF members on a 100x100x15 catchment, members sharded over the ranks, the analysis
                    with NCCL all-gather / all-reduce (strong scaling: 256 members at every N);
      `partitioned` BASELINE config 5: ONE 1000x1000x30 mesh (31 M nodes) row-block partitioned over the ranks, halo rows and
                    reduction scalars through peer memory inside the PCG kernel (strong scaling).
    The headline keeps one independent forward run per rank (`scaling: weak`, no collective), so N=1 equals the plain bench.
  * at N=1: `roofline_hbm` (the same PCG and SpMV on a mesh of 3.4 M nodes, 3.4 x the L2, HBM-resident), `full_run`
    (the whole TMAX = 7200 s of the headline workload), `setup_s`, `io_s`, `prepro` (one short pass of --workload prepro: the
    pre-processor on a 1000x1000 DEM with the reference's own ELF timed beside it).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pycathy_wrapper_b200 import synthetic  # noqa: E402
from pycathy_wrapper_b200.project import load_project  # noqa: E402

PCG_BYTES_PER_ROW_ITER = 168.0     # the CG recurrence in the DIA layout, every operand touched once per phase:
                                   # phase A 104 B + phase B 64 B per row and iteration (fp64) -- the ALGORITHMIC bytes
PCG_BYTES_PER_ROW_SETUP = 152.0    # x0, residual and first preconditioner application
# k_pcg_res2 keeps r, p, B (and x) of a CTA's rows in shared memory: what it still moves through L2/HBM per row and iteration
PCG_RES_BYTES_PER_ROW_ITER = 88.0  # 8 diagonals + z read (phase A); reciprocal diagonal + z write (phase B); +16 when x is not resident
SPMV_BYTES_PER_ROW = 80.0          # 8 diagonals + x + y
T_START = time.perf_counter()


def _nccl_to_stderr():
    """stdout must carry the ONE JSON line only.  NCCL prints its banner and (at NCCL_DEBUG=INFO) its communicator lines to
    stdout: they are redirected to stderr, not suppressed, so that the driver still sees how many ranks the communicator has."""
    if "NCCL_DEBUG" in os.environ and "NCCL_DEBUG_FILE" not in os.environ:
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"


def ncu_traffic(kernel: str):
    """DRAM bytes (read + write) of ONE captured launch of `kernel` from the committed `ncu --set full` summary
    (profiles/ncu_traffic.json, written from the .ncu-rep of the same bench command); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        rec = json.load(fh).get(kernel)
    return rec


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


ROUTE200 = os.path.join(ROOT, "tests", "golden", "route200_prepro.tar.xz")


def make_workload(size, member: int = 0, iopt: int = 1, routing: bool = False):
    nrow, ncol, nstr = size
    d = tempfile.mkdtemp(prefix="cathy_bench_")
    ks = 1.88e-4 * (1.0 + 0.05 * member)                     # ensemble members differ in Ks (weak scaling replicas)
    row = (ks, ks, ks, 1.0e-5, 0.55, 1.46, 0.15, 0.03125)
    if routing:
        # BASELINE config 3: Newton + coupled surface routing.  A storm (1e-4 m/s for 10 min) on a saturated hillslope (water table
        # at the surface) that starts with 5 mm of ponded water, so that SURF_FLOWTRA routes runoff from the first step on.  The
        # routing rasters of the 200 x 200 DEM come from the reference's own pre-processor (tests/golden/make_route200.py).
        rain = [(0.0, 0.0), (60.0, 1.0e-4), (600.0, 1.0e-4), (660.0, 0.0), (1.0e9, 0.0)]
        synthetic.make_project(d, nrow, ncol, nstr, ic=("hydrostatic",), pond=0.005, ISIMGR=2, DELTAT=1.0, DTMIN=1e-4, DTMAX=100.0, TMAX=7200.0, TIMPRT=[7200.0],
                               NODVP=[1], soil_rows=[row] * nstr, IOPT=iopt, ISOLV=0 if iopt == 2 else 2, atmbc=rain)
        if (nrow, ncol) == (200, 200):
            subprocess.run(["tar", "-xJf", ROUTE200, "-C", os.path.join(d, "prepro")], check=True)       # the reference ELF's own files
        else:
            from pycathy_wrapper_b200 import preprocessor                                                  # any other DEM: the device pre-processor
            preprocessor.run_preprocessor(os.path.join(d, "prepro"))
        prj = load_project(d)
        shutil.rmtree(d, ignore_errors=True)
        return prj
    synthetic.make_project(d, nrow, ncol, nstr, ic=("wt", 1.0), ISIMGR=1, DELTAT=1.0, DTMIN=1e-2, DTMAX=100.0,
                           TMAX=7200.0, TIMPRT=[7200.0], NODVP=[1], soil_rows=[row] * nstr, hspatm=0, IOPT=iopt, ISOLV=0 if iopt == 2 else 2,
                           atmbc=[(0.0, np.zeros((nrow + 1) * (ncol + 1))), (60.0, np.full((nrow + 1) * (ncol + 1), 2.0e-5)),
                                  (3600.0, np.full((nrow + 1) * (ncol + 1), 2.0e-5)),
                                  (3660.0, np.zeros((nrow + 1) * (ncol + 1))), (1.0e9, np.zeros((nrow + 1) * (ncol + 1)))])
    prj = load_project(d)
    shutil.rmtree(d, ignore_errors=True)
    return prj


def seq_entry(rep):
    """One accepted step as both arms print it: (NSTEP, nonlinear iterations, back-steps, DELTAT) -- equal lists in the `ours` and
    the `reference` line are a parity check of the driver-run pair itself."""
    return [int(rep.nstep), int(rep.iter), int(rep.kbackt), float("%.12g" % rep.deltat)]


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md clocks line): ONE long-lived
    `nvidia-smi -lms` child per rank instead of a new process every 200 ms."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.maxmhz, self.proc = gpu_index, [], set(), None, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                out = line.strip().split(",")
                try:
                    self.samples.append(float(out[0]))
                    self.maxmhz = float(out[1])
                    for nm, v in zip(names, out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(nm)
                except (ValueError, IndexError):
                    pass
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.maxmhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


class Ctx:
    """torch.distributed plumbing of one rank (one process per GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            _nccl_to_stderr()
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        import __graft_entry__ as g
        if self.rank == 0:
            g.build()
        if self.world > 1:
            dist.barrier()
        from pycathy_wrapper_b200.capi import load_library
        self.lib = load_library()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def rmax(self, x):
        return self._red(x, self.dist.ReduceOp.MAX)

    def rmin(self, x):
        return self._red(x, self.dist.ReduceOp.MIN)

    def rsum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ======================================================================================================
# headline: BASELINE config 2 (or --workload newton / coupled), one forward run per rank
# ======================================================================================================
def run_headline(ctx: Ctx, args, size) -> dict:
    from pycathy_wrapper_b200.capi import Simulation
    torch = ctx.torch
    world, rank, local, lib = ctx.world, ctx.rank, ctx.local, ctx.lib
    newton = args.workload in ("newton", "coupled")
    coupled = args.workload == "coupled"
    t_io = time.perf_counter()
    prj = make_workload(size, member=rank, iopt=2 if newton else 1, routing=coupled)
    t_io = time.perf_counter() - t_io

    # ---------------- device-resident measurement ("value") ----------------
    t_setup = time.perf_counter()
    sim = Simulation(lib, prj, device=local)
    t_setup = time.perf_counter() - t_setup
    n = sim.n
    for _ in range(args.warmup):
        sim.step()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.barrier()
    t0 = time.perf_counter()
    gpu_ms = pcg_ms = 0.0
    pcg_iters = pcg_solves = launches = nl_its = 0
    seq = []
    for _ in range(args.steps):
        rep = sim.step()
        gpu_ms += rep.gpu_ms
        pcg_ms += rep.pcg_ms
        pcg_iters += rep.pcg_iters
        pcg_solves += rep.pcg_solves
        launches += rep.launches
        nl_its += rep.iter
        seq.append(seq_entry(rep))
        if rep.finished:
            raise SystemExit("bench.py: workload finished before K steps; lower --steps")
    ctx.barrier()
    wall = time.perf_counter() - t0
    sampler.stop()
    # device time of the timed region = sum of per-step CUDA-event times on the simulation's stream
    dev_s = ctx.rmax(gpu_ms / 1e3)
    wall_s = ctx.rmax(wall)
    value = world * n * args.steps / wall_s
    # SpMV sample on the last assembled system (L2-resident at this size: reported as such, not as an HBM fraction)
    x = np.random.default_rng(0).standard_normal(n)
    _, spmv_ms = sim.debug_spmv(x, reps=50)
    solver = sim.solver_info()
    limits = sim.solver_limits()
    sim_nnz, sim_nt = sim.nnz, sim.nt
    # text I/O of one detailed output (psi + sw blocks, what DETOUT prints at a TIMPRT), for `io_s`
    t_out = None
    if rank == 0:
        from pycathy_wrapper_b200 import outputs as O
        st = sim.state()
        t1 = time.perf_counter()
        with tempfile.TemporaryFile("w") as fh:
            O.write_block(fh, 1, 1.0, st["psi"])
            O.write_block(fh, 1, 1.0, st["sw"])
        t_out = time.perf_counter() - t1
    sim.close()

    # ---------------- end-to-end through the C ABI with host buffers ("e2e") ----------------
    sim = Simulation(lib, prj, device=local)
    forcing = np.ascontiguousarray(prj.atm_values[1])          # pinned by torch below
    pin = torch.from_numpy(forcing.copy()).pin_memory()
    forcing = pin.numpy()
    hostbufs = [sim.state_buffers(pinned=True) for _ in range(2)]   # the caller's (pinned) host buffers for the per-step read-back
    sim.upload_atm_record(1, forcing)
    for i in range(args.warmup):
        sim.step()
        sim.upload_atm_record(1, forcing)
        sim.state_async(hostbufs[i & 1])
    sim.state_wait()
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        sim.step()
        # H2D: the NEXT step's forcing record, enqueued before this step's read-back starts to drain: both copies would
        # otherwise meet on the copy engines and the upload (on the compute stream) would wait for the 28 MB drain.  K steps = K uploads.
        sim.upload_atm_record(1, forcing)
        # D2H: psi, sw, ckrw, ... (what DETOUT prints) of THIS step, snapshotted on the device and drained by a second stream
        # while the next step computes; two host buffer sets alternate, the last read-back is awaited inside the timed region
        st = sim.state_async(hostbufs[i & 1])
    sim.state_wait()
    ctx.barrier()
    e2e_s = ctx.rmax(time.perf_counter() - t0)
    d2h = sum(v.nbytes for v in st.values())
    h2d = forcing.nbytes
    sim.close()
    e2e_value = world * n * args.steps / e2e_s

    peak, peak_src = measured_peaks()
    resident = solver["kernel"] in (3, 4, 11)
    if newton:      # BiCGSTAB: two SpMVs per iteration at SURVEY section 8d's CSR figure (12 nnz + 20 N bytes each) + 10 vector passes
        alg_row_iter = (2.0 * (12.0 * sim_nnz + 20.0 * n) + 80.0 * n) / n
        alg_bytes = pcg_iters * alg_row_iter * n
        moved_row_iter = 2.0 * 15.0 * 8.0 + (64.0 if solver["kernel"] == 11 else 208.0)
        kname = ("k_bicgstab_res (persistent line-preconditioned BiCGSTAB, Krylov vectors resident in shared memory)" if solver["kernel"] == 11
                 else "k_bicgstab (persistent right-preconditioned BiCGSTAB on the 15-diagonal Jacobian)")
    else:
        alg_row_iter = PCG_BYTES_PER_ROW_ITER
        alg_bytes = (pcg_iters * PCG_BYTES_PER_ROW_ITER + pcg_solves * PCG_BYTES_PER_ROW_SETUP) * n
        moved_row_iter = (PCG_RES_BYTES_PER_ROW_ITER + (0.0 if solver["x_resident"] else 16.0)) if resident else PCG_BYTES_PER_ROW_ITER
        kname = ("%s (persistent PCG, CG vectors resident in shared memory: SpMV on z + fused vector ops)" % ("k_pcg_res2" if solver["kernel"] == 4 else "k_pcg_res")
                 if resident else "k_pcg (persistent PCG: SpMV + fused vector ops)")
    alg_gbs = alg_bytes / (pcg_ms / 1e3) / 1e9 if pcg_ms > 0 else None
    moved_bytes = pcg_iters * moved_row_iter * n + (0 if newton else pcg_solves * PCG_BYTES_PER_ROW_SETUP * n)
    moved_gbs = moved_bytes / (pcg_ms / 1e3) / 1e9 if pcg_ms > 0 else None
    # `traffic`: measured DRAM bytes of one captured launch (ncu --set full of the same command), next to the bytes of an average launch
    trec = ncu_traffic(("k_pcg_res2" if solver["kernel"] == 4 else "k_pcg_res") if resident else "k_pcg") if (size == (200, 200, 20) and not newton) else None
    traffic = None
    if trec:
        traffic = {"dram_bytes_per_launch": trec["dram_bytes_per_launch"], "pcg_iters_in_launch": trec.get("pcg_iters_in_launch"),
                   "algorithmic_bytes_per_launch": alg_bytes / max(pcg_solves, 1), "source": trec.get("source")}
    working_set_mb = n * ((240.0 if newton else 64.0) + 24.0) / 1e6
    in_l2 = working_set_mb < 110.0
    wl = ("synthetic %dx%d DEM x %d layers (%d nodes, %d tets), van Genuchten, %s, %s, first %d accepted steps after %d warm-up"
          % (size[1], size[0], size[2], n, sim_nt,
             ("Newton+BiCGSTAB, coupled surface routing (BASELINE config 3)" if coupled else "Newton+BiCGSTAB (BASELINE config 3 without the surface-routing coupling)") if newton else "Picard+PCG",
             "storm of 1e-4 m/s on a saturated hillslope (INDP=2) with 5 mm of initial ponding (IPOND=1), ISIMGR=2: SURF_FLOWTRA routing every step" if coupled
             else "infiltration pulse on a hillslope with a water table 2 m deep (INDP=3), ISIMGR=1", args.steps, args.warmup))
    out = {
        "metric": "node-timesteps/s", "value": value, "unit": "node-timesteps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl,
                   "parallelism": "1 independent forward run (ensemble member) per GPU, no collective; the sharded workloads are under `sharded`" if world > 1 else "single forward run",
                   "l2": "inputs larger than L2: every nonlinear iteration streams the %.2f GB gather plan and the nodal soil constants through the 126 MB L2 "
                         "between two linear solves; inside ONE solve (one persistent launch) the matrix (%.0f MB) is re-read every iteration, no flush there" % (1.27e3 * n / 1e9, n * (240 if newton else 64) / 1e6),
                   "nonlinear_its": nl_its, "pcg_iters": pcg_iters, "pcg_solves": pcg_solves,
                   "linear_solver_limits": limits},
        "nonlinear_its": nl_its,
        "step_sequence": seq,
        "device_ms_per_step": 1e3 * dev_s / args.steps,
        "e2e": {"value": e2e_value, "unit": "node-timesteps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        # The solver kernel at THIS size works out of L2 + shared memory (ncu: DRAM ~1 % of peak, lts hit rate 95 %), so dividing
        # the algorithmic bytes by the HBM peak would measure the cache (frac 1.3 in round 1).  `achieved` therefore counts the bytes
        # the kernel really moves through the SM boundary (global-memory loads + stores of its design); the algorithmic figure is
        # kept beside it, and the HBM-resident measurement of the same solver is `roofline_hbm`.
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": moved_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (moved_gbs / peak) if moved_gbs else None,
                     "bytes_convention": "global-memory bytes the kernel moves per row and iteration (%.0f B); SURVEY 8d's algorithmic figure is %.0f B" % (moved_row_iter, alg_row_iter),
                     "served_from": ("L2 (working set %.0f MB < 126 MB): `frac` relates L2-served traffic to the HBM copy peak and is NOT an HBM utilisation; see roofline_hbm" % working_set_mb)
                                    if in_l2 else "HBM (working set %.0f MB > L2)" % working_set_mb,
                     "algorithmic": {"bytes_per_row_iter": alg_row_iter, "achieved": alg_gbs, "frac_of_hbm_peak": (alg_gbs / peak) if alg_gbs else None},
                     "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_detail": traffic,
                     "share_of_step": pcg_ms / gpu_ms if gpu_ms else None,
                     "us_per_iter": 1e3 * pcg_ms / max(pcg_iters, 1), "iters_per_solve": pcg_iters / max(pcg_solves, 1),
                     "spmv_only": {"ms": spmv_ms, "achieved": SPMV_BYTES_PER_ROW * n / (spmv_ms / 1e3) / 1e9,
                                   "served_from": "L2 (50 back-to-back products on the same %.0f MB)" % (n * 80 / 1e6) if n * 80 < 110e6 else "HBM",
                                   "csr_equivalent_gbs": (12.0 * sim_nnz + 20.0 * n) / (spmv_ms / 1e3) / 1e9}},
        "setup_s": {"cathy_create": t_setup, "what": "mesh, sparsity, static gather plan, initial state on the device (host build, single thread)"},
        "io_s": {"project_text_write_and_parse": t_io, "one_detailed_output_psi_sw": t_out,
                 "what": "synthetic project written and parsed through the reference's text formats; psi + sw blocks of one TIMPRT"},
    }
    return out


def run_roofline_hbm(ctx: Ctx, args) -> dict:
    """The same PCG kernel family and the SpMV on a mesh that does NOT fit the caches: 400x400 DEM x 20 layers = 3,376,821 nodes; the
    8 diagonals are 216 MB and one PCG iteration touches 567 MB (4.5 x the 126 MB L2).  Between two solves the assembly streams
    4.3 GB through the L2 (a natural flush); inside a solve every iteration re-reads more than the L2 holds."""
    from pycathy_wrapper_b200.capi import Simulation
    size = (400, 400, 20)
    prj = make_workload(size)
    sim = Simulation(ctx.lib, prj, device=ctx.local)
    n = sim.n
    for _ in range(2):
        sim.step()
    pcg_ms = 0.0
    iters = solves = 0
    for _ in range(4):
        rep = sim.step()
        pcg_ms += rep.pcg_ms; iters += rep.pcg_iters; solves += rep.pcg_solves
    x = np.random.default_rng(0).standard_normal(n)
    _, spmv_ms = sim.debug_spmv(x, reps=20)
    solver = sim.solver_info()
    nnz = sim.nnz
    sim.close()
    peak, peak_src = measured_peaks()
    b = (iters * PCG_BYTES_PER_ROW_ITER + solves * PCG_BYTES_PER_ROW_SETUP) * n
    ach = b / (pcg_ms / 1e3) / 1e9
    sp = SPMV_BYTES_PER_ROW * n / (spmv_ms / 1e3) / 1e9
    return {"bound": "hbm", "mesh": "400x400 DEM x 20 layers, %d nodes: matrix %.0f MB, %.0f MB touched per PCG iteration (L2: 126 MB)" % (n, n * 64 / 1e6, n * 168 / 1e6),
            "kernel": {1: "k_pcg (streaming, layer-major)", 5: "k_pcg on the column-major permutation (streaming, direct loads)",
                       6: "k_pcg_tma (column-major permutation; tiles staged by cp.async.bulk + mbarrier, producer warp, round-robin tiles)"}.get(solver["kernel"], "kernel %d" % solver["kernel"]),
            "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
            "bytes_per_row_iter": PCG_BYTES_PER_ROW_ITER, "us_per_iter": 1e3 * pcg_ms / max(iters, 1), "pcg_iters": iters, "pcg_solves": solves,
            "flush": "inputs larger than L2; 4.3 GB of assembly traffic between two solves",
            "spmv_only": {"kernel": "k_spmv_tma (y = A x with the TMA-staged tiles of k_pcg_tma's stencil phase)" if solver["kernel"] == 6 else "k_spmv (direct loads)",
                          "ms": spmv_ms, "achieved": sp, "frac": sp / peak, "bytes_per_row": SPMV_BYTES_PER_ROW,
                          "csr_equivalent_gbs": (12.0 * nnz + 20.0 * n) / (spmv_ms / 1e3) / 1e9,
                          "flush": "20 back-to-back products, each streaming 270 MB (2.1 x the L2)"}}


def run_full(ctx: Ctx, args, size, wall_cap_s: float = 90.0) -> dict:
    """The whole TMAX = 7200 s of the headline workload (the 20-step window above sits at t = 5..40 s, where dt is small and
    the system mass-matrix dominated): node-timesteps/s over all accepted steps, linear iterations per solve as a histogram."""
    from pycathy_wrapper_b200.capi import Simulation
    prj = make_workload(size)
    sim = Simulation(ctx.lib, prj, device=ctx.local)
    n = sim.n
    t0 = time.perf_counter()
    k = nl = back = 0
    gpu_ms = pcg_ms = 0.0
    its = []
    while True:
        rep = sim.step()
        k += 1; nl += rep.iter; back += rep.kbackt; gpu_ms += rep.gpu_ms; pcg_ms += rep.pcg_ms
        its.extend(int(rep.it[q].niter) for q in range(rep.n_iter_rec))
        if rep.finished or time.perf_counter() - t0 > wall_cap_s:
            break
    wall = time.perf_counter() - t0
    t_end, done = rep.time, bool(rep.finished)
    sim.close()
    its = np.array(its)
    edges = [0, 10, 20, 40, 80, 160, 320, 640, 100000]
    hist = {("%d-%d" % (edges[i] + 1, edges[i + 1])) if edges[i + 1] < 100000 else ">%d" % edges[i]: int(((its > edges[i]) & (its <= edges[i + 1])).sum()) for i in range(len(edges) - 1)}
    return {"value": n * k / wall, "unit": "node-timesteps/s", "accepted_steps": k, "nonlinear_its": nl, "back_steps": back, "simulated_s": t_end,
            "reached_tmax": done, "wall_s": wall, "ms_per_step": 1e3 * wall / k, "device_ms_per_step": gpu_ms / k,
            "pcg_its_per_solve": {"mean": float(its.mean()), "median": float(np.median(its)), "max": int(its.max()), "histogram": hist},
            "pcg_share_of_device_time": pcg_ms / gpu_ms if gpu_ms else None}


# ======================================================================================================
# sharded workload 1: BASELINE config 4, EnKF with members sharded over the ranks
# ======================================================================================================
def run_enkf(ctx: Ctx, args, size, cycles: int, warm: int) -> dict:
    """EnKF data assimilation, `--members` members (default 256) on a 100x100x15 catchment, members round-robin over the ranks,
    SWC observations at 64 surface nodes, one forecast window + one analysis per cycle.
    Metric: ensemble member-steps/s (accepted time steps summed over members / seconds), analysis included."""
    from pycathy_wrapper_b200 import da
    torch, dist = ctx.torch, ctx.dist
    world, rank, local = ctx.world, ctx.rank, ctx.local
    nrow, ncol, nstr = size
    ne = args.members
    mine = [k for k in range(ne) if k % world == rank]
    rng = np.random.default_rng(1234)
    lnk = 0.5 * rng.standard_normal(ne)                      # log-normal Ks, sigma = 0.5
    dwt = 0.25 * rng.standard_normal(ne)                     # IC: water-table depth perturbation, sigma = 0.25 m
    window = 1800.0
    prjs = []
    t_io = time.perf_counter()
    for k in mine:
        d = tempfile.mkdtemp(prefix="cathy_enkf_")
        ks = 1.88e-4 * float(np.exp(lnk[k]))
        row = (ks, ks, ks, 1.0e-5, 0.55, 1.46, 0.15, 0.03125)
        synthetic.make_project(d, nrow, ncol, nstr, ic=("wt", 1.0 + float(dwt[k])), ISIMGR=1, DELTAT=10.0, DTMIN=1e-2, DTMAX=300.0,
                               TMAX=window, TIMPRT=[window], NODVP=[1], soil_rows=[row] * nstr,
                               atmbc=[(0.0, 5.0e-6), (1.0e9, 5.0e-6)])
        prjs.append(load_project(d))
        shutil.rmtree(d, ignore_errors=True)
    t_io = time.perf_counter() - t_io
    t_build = time.perf_counter()
    ens = da.Ensemble(ctx.lib, prjs, device=local, concurrent=args.concurrent)
    t_build = time.perf_counter() - t_build
    n, nnod = ens.n, prjs[0].nnod
    m = 64
    obs_nodes = (np.linspace(0, nnod - 1, m).astype(np.int64))          # 64 surface-layer nodes
    R = np.diag(np.full(m, 0.02 ** 2))
    noise = 0.02 * np.random.default_rng(4321).standard_normal(m)        # synthetic truth = ensemble-mean SWC + observation noise
    failed_total = [0]
    kern = {"crosscov_ms": 0.0, "update_ms": 0.0, "collective_ms": 0.0, "n": 0}

    def cycle():
        steps = ens.forecast()
        failed_total[0] += len(ens.failed)
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        ctx.barrier()                    # rank skew (members differ in their step counts) is WAIT, not analysis work
        wait_ms = 1e3 * (time.perf_counter() - t_w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        HX = ens.predict_obs(obs_nodes, 0.55)
        ybar = HX.sum(dim=1)
        if world > 1:
            dist.all_reduce(ybar)
        y = (ybar / ne).cpu().numpy() + noise
        info = ens.analysis(obs_nodes, 0.55, y, R, sakov=False, inflate=1.02, HX=HX)
        e1.record()
        ens.restart(window, 10.0)
        torch.cuda.synchronize()
        tm = info.get("timing_ms", {})
        for k_ in ("crosscov_ms", "update_ms", "collective_ms"):
            kern[k_] += tm.get(k_, 0.0)
        kern["n"] += 1
        return steps, e0.elapsed_time(e1), wait_ms

    for _ in range(max(warm, 1)):
        cycle()
    kern.update({"crosscov_ms": 0.0, "update_ms": 0.0, "collective_ms": 0.0, "n": 0})
    sampler = ClockSampler(local)
    sampler.start()
    ctx.barrier()
    t0 = time.perf_counter()
    steps = 0
    ana_ms = wait_ms = 0.0
    for _ in range(cycles):
        s_, a_, w_ = cycle()
        steps += s_
        ana_ms += a_
        wait_ms += w_
    ctx.barrier()
    wall = ctx.rmax(time.perf_counter() - t0)
    sampler.stop()
    steps_all = ctx.rsum(float(steps))
    peak, peak_src = measured_peaks()
    ne_loc = len(mine)
    # k_enkf_crosscov reads X once (8 N ne_local) and writes P (8 N m); k_enkf_update reads X, P and writes X (16 N ne_local + 8 N m)
    cc_b, up_b = 8.0 * n * (ne_loc + m), 8.0 * n * (2 * ne_loc + m)
    ncy = max(kern["n"], 1)
    cc_ms, up_ms = ctx.rmax(kern["crosscov_ms"] / ncy), ctx.rmax(kern["update_ms"] / ncy)
    roof = {"bound": "hbm", "kernel": "k_enkf_crosscov_r8 + k_enkf_update_r8 (fp64 DMMA, small operand resident in shared memory; N = %d, Ne_local = %d, m = %d)" % (n, ne_loc, m),
            "bytes": cc_b + up_b, "ms": cc_ms + up_ms, "crosscov_ms": cc_ms, "update_ms": up_ms,
            "achieved": (cc_b + up_b) / ((cc_ms + up_ms) / 1e3) / 1e9 if cc_ms + up_ms > 0 else None, "peak": peak, "peak_source": peak_src, "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / peak if roof["achieved"] else None
    # the pair is bound by the fp64 tensor pipe, not by HBM: 2 N Ne m flop each against the DMMA rate measured on this pool's B200
    # with tools/bench_dmma.cu (37.1 TFLOP/s for DMMA.8x8x4 with 32 warps per SM; the plain DFMA pipe does 34.1)
    fl = 4.0 * n * ne_loc * m
    roof["tensor"] = {"bound": "tensor", "unit": "TFLOP/s", "peak": 37.1, "peak_source": "measured (tools/bench_dmma.cu, profiles/micro/r2f_dmma_peak.log)",
                      "achieved": fl / ((cc_ms + up_ms) / 1e3) / 1e12 if cc_ms + up_ms > 0 else None}
    roof["tensor"]["frac"] = roof["tensor"]["achieved"] / 37.1 if roof["tensor"]["achieved"] else None
    out = {"metric": "ensemble member-steps/s", "value": steps_all / wall, "unit": "member-steps/s", "n_gpus": world, "steps": cycles,
           "warmup": max(warm, 1), "ms_per_step": 1e3 * wall / cycles, "higher_is_better": True, "scaling": "strong",
           "config": {"workload": "BASELINE config 4: EnKF DA, %d members on a %dx%d DEM x %d layers (%d nodes), %d SWC observations, window %.0f s; a step = "
                                  "one forecast window of every member + one analysis (NCCL all-gather / all-reduce when sharded)" % (ne, ncol, nrow, nstr, n, m, window),
                      "parallelism": "members round-robin over %d GPU(s), %d concurrent per GPU" % (world, args.concurrent), "member_steps_per_cycle": steps_all / cycles},
           "node_member_steps_per_s": steps_all * n / wall,
           "analysis_ms": {"wait": ctx.rmax(wait_ms / cycles), "compute": ctx.rmax(ana_ms / cycles), "collectives": ctx.rmax(kern["collective_ms"] / ncy),
                           "what": "per cycle, max over ranks; wait = barrier after the forecast (rank skew), compute = predict_obs + gain + kernels + NCCL collectives"},
           "roofline": roof,
           "setup_s_per_rank": ctx.rmax(t_build), "io_s_per_rank": ctx.rmax(t_io), "failed_member_windows": ctx.rsum(float(failed_total[0])), "clocks": sampler.summary()}
    ens.close()
    return out


# ======================================================================================================
# sharded workload 2: BASELINE config 5, one mesh row-block partitioned over the ranks
# ======================================================================================================
def run_partitioned(ctx: Ctx, args, size, steps: int, warm: int) -> dict:
    """ONE mesh, row-block partitioned over the ranks (strips of DEM rows), Picard + PCG with halo rows and reduction scalars
    exchanged through peer memory over NVLink inside the solver kernel.  Strong scaling: the mesh is fixed."""
    from pycathy_wrapper_b200.capi import Simulation
    from pycathy_wrapper_b200.partition import PartitionedSimulation
    world, local = ctx.world, ctx.local
    nrow, ncol, nstr = size
    d = tempfile.mkdtemp(prefix="cathy_part_")
    t_io = time.perf_counter()
    synthetic.make_project(d, nrow, ncol, nstr, ic=("wt", 1.0), ISIMGR=1, DELTAT=1.0, DTMIN=1e-2, DTMAX=100.0, TMAX=600.0, TIMPRT=[600.0],
                           NODVP=[1], atmbc=[(0.0, 0.0), (60.0, 2.0e-5), (1.0e9, 2.0e-5)])
    prj = load_project(d)
    shutil.rmtree(d, ignore_errors=True)
    t_io = time.perf_counter() - t_io
    t_build = time.perf_counter()
    sim = PartitionedSimulation(ctx.lib, prj, device=local) if world > 1 else Simulation(ctx.lib, prj, device=local)
    t_build = time.perf_counter() - t_build
    n_global = sim.n_global if world > 1 else sim.n
    n_local = sim.sim.n if world > 1 else sim.n
    solver = (sim.sim if world > 1 else sim).solver_info()
    for _ in range(warm):
        sim.step()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.barrier()
    t0 = time.perf_counter()
    gpu_ms = pcg_ms = 0.0
    pcg_iters = pcg_solves = launches = nl = 0
    seq = []
    for _ in range(steps):
        rep = sim.step()
        gpu_ms += rep.gpu_ms; pcg_ms += rep.pcg_ms; pcg_iters += rep.pcg_iters; pcg_solves += rep.pcg_solves; launches += rep.launches; nl += rep.iter
        seq.append(seq_entry(rep))
    ctx.barrier()
    wall = ctx.rmax(time.perf_counter() - t0)
    sampler.stop()
    peak, peak_src = measured_peaks()
    pcg_bytes = (pcg_iters * PCG_BYTES_PER_ROW_ITER + pcg_solves * PCG_BYTES_PER_ROW_SETUP) * n_local
    pcg_ms_max = ctx.rmax(pcg_ms)
    achieved = pcg_bytes / (pcg_ms_max / 1e3) / 1e9 if pcg_ms_max > 0 else None
    halo = 2 * 2 * (nstr + 1) * (ncol + 1) * 8 if world > 1 else 0          # bytes stored into the neighbours per PCG iteration (interior rank)
    out = {"metric": "node-timesteps/s", "value": n_global * steps / wall, "unit": "node-timesteps/s", "n_gpus": world, "steps": steps,
           "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "strong",
           "config": {"workload": "BASELINE config 5: ONE synthetic %dx%d DEM x %d layers mesh (%d nodes), Picard+PCG, infiltration pulse; row-block partitioned into %d strip(s) of DEM rows, "
                                  "halo exchange + all-reduce through peer memory inside the PCG kernel" % (ncol, nrow, nstr, n_global, world),
                      "parallelism": "row-block x%d" % world, "nodes_per_rank": n_local, "nonlinear_its": nl, "pcg_iters": pcg_iters, "pcg_solves": pcg_solves,
                      "halo_bytes_per_pcg_iteration": halo},
           "step_sequence": seq,
           "device_ms_per_step": 1e3 * ctx.rmax(gpu_ms / 1e3) / steps, "pcg_us_per_iteration": 1e3 * pcg_ms_max / max(pcg_iters, 1),
           "gpu_launches": launches, "setup_s_per_rank": ctx.rmax(t_build), "io_s_per_rank": ctx.rmax(t_io), "clocks": sampler.summary(),
           "roofline": {"bound": "hbm", "kernel": "solver kernel %d (persistent PCG, per rank; 1 = k_pcg layer-major, 5 = k_pcg column-major, 6 = k_pcg_tma column-major with TMA-staged tiles)" % solver["kernel"], "achieved": achieved, "peak": peak, "peak_source": peak_src,
                        "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "bytes_per_row_iter": PCG_BYTES_PER_ROW_ITER,
                        "served_from": "HBM (per-rank matrix %.0f MB)" % (n_local * 64 / 1e6), "traffic": None}}
    sim.close()
    return out


def cpu_baseline(size, budget_s: float = 25.0, max_steps: int = 1000, iopt: int = 1, routing: bool = False):
    """The CPU oracle (a C port of the reference's algorithm; the reference ELFs cannot hold this mesh)
    timed on a bounded sample of the same workload: the first accepted step(s), single thread."""
    from oracle import oracle
    prj = make_workload(size, iopt=iopt, routing=routing)
    sim = oracle.simulation(prj)
    t0 = time.perf_counter()
    k = 0
    while True:
        rep = sim.step()
        k += 1
        if time.perf_counter() - t0 > budget_s or rep.finished or k >= max_steps:
            break
    dt = time.perf_counter() - t0
    return {"value": sim.n * k / dt, "unit": "node-timesteps/s", "cores": 1, "kind": "port",
            "sample": "first %d accepted time step(s) of the same workload (%.1f s of CPU work, sequential %s as in the reference)" % (k, dt, "ILU(0)-BiCGSTAB" if iopt == 2 else "IC(0)-PCG")}


def _ref_worker(job):
    """One reference process = one forward run on one host core (the reference has no intra-run threading; ensembles are
    parallel host processes, pyCATHY/DA/cathy_DA.py:1382-1390)."""
    size, member, iopt, routing, warmup, steps, budget = job
    from oracle import oracle
    prj = make_workload(size, member=member, iopt=iopt, routing=routing)
    sim = oracle.simulation(prj)
    t_w = time.perf_counter()
    nw = 0
    for _ in range(warmup):
        if time.perf_counter() - t_w > 30.0:
            break
        sim.step()
        nw += 1
    t0 = time.perf_counter()
    k = nl = 0
    seq = []
    for _ in range(steps):
        rep = sim.step()
        k += 1
        nl += rep.iter
        seq.append(seq_entry(rep))
        if time.perf_counter() - t0 > budget:
            break
    return {"n": sim.n, "k": k, "nw": nw, "dt": time.perf_counter() - t0, "nl": nl, "seq": seq}


def run_reference(args, size):
    """Reference arm: the reference's CPU implementation of the path (its C restatement, oracle/: no Fortran compiler here and the
    shipped ELFs hold <= 82,416 nodes) on the host cores -- min(N, cores) parallel processes, one forward run each, like the
    reference runs ensemble members.  Rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "prepro":
        return run_prepro_reference(args, size)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    newton = args.workload in ("newton", "coupled")
    cores = os.cpu_count() or 1
    procs = max(1, min(world, cores, 8))
    budget = 150.0
    jobs = [(size, r, 2 if newton else 1, args.workload == "coupled", args.warmup, args.steps, budget) for r in range(procs)]
    if procs == 1:
        res = [_ref_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_worker, jobs)
    n = res[0]["n"]
    dt = max(r["dt"] for r in res)
    kmin = min(r["k"] for r in res)
    v = sum(r["n"] * r["k"] for r in res) / dt
    # N GPUs advance N runs; fewer host processes than ranks is stated, not scaled up
    print(json.dumps({
        "impl": "reference", "metric": "node-timesteps/s", "value": v, "unit": "node-timesteps/s", "n_gpus": world,
        "steps": kmin, "warmup": res[0]["nw"], "ms_per_step": 1e3 * dt / max(kmin, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %dx%d DEM x %d layers (%d nodes), same project files as the GPU arm; CPU restatement of the reference "
                               "(oracle port: no Fortran compiler here and the shipped ELFs are dimensioned for <= 82,416 nodes)" % (size[1], size[0], size[2], n),
                   "parallelism": "%d independent forward run(s) as parallel host processes, one core each (%d host cores)" % (procs, cores)},
        "nonlinear_its": res[0]["nl"], "step_sequence": res[0]["seq"],
        "cpu_baseline": {"value": v, "unit": "node-timesteps/s", "cores": procs, "kind": "port",
                         "sample": "%d accepted step(s) after %d warm-up per process, time-boxed to %.0f s; %d process(es) x 1 core (the reference has no intra-run threading)" % (kmin, res[0]["nw"], budget, procs)},
        "e2e": {"value": v, "unit": "node-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ======================================================================================================
# pre-processor: hap.in + dtm_13.val -> routing rasters, one pass = one "step"

def _prepro_project(size):
    """<project>/prepro with hap.in and dtm_13.val of the synthetic DEM (the DEM of BASELINE config 5 at 1000x1000)."""
    nrow, ncol = size[0], size[1]
    d = tempfile.mkdtemp(prefix="cathy_prepro_")
    synthetic.write_hapin(os.path.join(d, "hap.in"), nrow, ncol, 0.5, 0.5)
    z = synthetic.synthetic_dem(nrow, ncol)
    np.savetxt(os.path.join(d, "dtm_13.val"), z, fmt="%.9f", delimiter="\t")
    shutil.copy(os.path.join(d, "hap.in"), os.path.join(d, "hap.in.orig"))
    return d


def _pycppp_seconds(d):
    """The reference's own ELF on the same directory (oracle/_ref/bin/pycppp, answers 2 / 0 / 1 as pyCATHY gives them)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "pycppp")
    if not os.path.exists(exe):
        return None
    run = tempfile.mkdtemp(prefix="cathy_prepro_ref_")
    shutil.copy(os.path.join(d, "hap.in.orig"), os.path.join(run, "hap.in"))
    shutil.copy(os.path.join(d, "dtm_13.val"), os.path.join(run, "dtm_13.val"))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "oracle", "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    t0 = time.perf_counter()
    subprocess.run([exe], cwd=run, env=env, input="2\n0\n1\n", text=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1800)
    dt = time.perf_counter() - t0
    ok = os.path.exists(os.path.join(run, "qoi_a"))
    same = None
    if ok:
        same = all(open(os.path.join(run, f)).read() == open(os.path.join(d, f)).read()
                   for f in ("qoi_a", "dem", "dtm_w_1", "dtm_w_2", "dtm_p_outflow_1", "dtm_p_outflow_2", "dtm_local_slope_1", "dtm_epl_2",
                             "dtm_kSs1_sf_1", "dtm_Ws1_sf_2", "dtm_nrc") if os.path.exists(os.path.join(d, f)))
    shutil.rmtree(run, ignore_errors=True)
    return {"seconds": dt, "ok": ok, "files_identical_to_ours": same}


def run_prepro(ctx: Ctx, args, size) -> dict:
    from pycathy_wrapper_b200 import preprocessor as pp
    torch = ctx.torch
    d = _prepro_project(size)
    hap = open(os.path.join(d, "hap.in")).read()
    dtm = open(os.path.join(d, "dtm_13.val")).read()
    for _ in range(args.warmup):
        res = pp.terrain_analysis(hap, dtm, device=ctx.local)
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ctx.barrier()
    dev_ms, stages = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = pp.terrain_analysis(hap, dtm, device=ctx.local)          # text -> device -> cell records (H2D + D2H inside)
        dev_ms.append(res.info["device_ms"])
        stages.append(res.info["stage_ms"])
    ctx.barrier()
    t_api = (time.perf_counter() - t0) / args.steps
    # end to end = the call a pyCATHY user makes: ./pycppp in <project>/prepro, files in, files out
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        shutil.copy(os.path.join(d, "hap.in.orig"), os.path.join(d, "hap.in"))
        res = pp.run_preprocessor(d, device=ctx.local)
    t_e2e = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
    sampler.stop()
    n = res.info["n_cells"]
    ms = float(np.mean(dev_ms))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbps_burst", peaks.get("hbm_gbps", 6547.5))) if isinstance(peaks, dict) else 6547.5
    st = {k: float(np.mean([s_[k] for s_ in stages])) for k in stages[0]}
    top = max(st, key=st.get)
    bytes_cell = 9 * 8 + 19 * 4 + 2 * 8 + 60            # 3x3 window of doubles in, 19 cell fields out, donor gathers
    out = {
        "metric": "pre-processor cells/s (hap.in + dtm_13.val -> routing rasters)", "value": n / (ms * 1e-3), "unit": "cells/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CATHY pre-processor on the synthetic %dx%d DEM (%d cells; the DEM of BASELINE config 5 at 1000x1000): CSORT, DEPIT, CSORT, CCA, SMEAN, "
                               "DSF, HG on the device; value = cells / CUDA-event time of the device part; inputs larger than L2 are not the point here: the path is "
                               "bound by dependency chains (DEPIT's sequential sweeps, %d drainage wavefronts), not by bandwidth" % (size[0], size[1], n, res.info["n_waves"])},
        "e2e": {"value": n / t_e2e, "unit": "cells/s", "seconds": t_e2e, "what": "run_preprocessor(<project>/prepro): parse hap.in + dtm_13.val, device, write 21 rasters + qoi_a + hap.in",
                "api_only_cells_per_s": n / t_api, "h2d_bytes_per_step": 9 * n, "d2h_bytes_per_step": (2 * 8 + 13 * 4 + 5 * 4) * n},
        "gpu_launches": int(res.info["n_launches"]) * args.steps,
        "stage_ms": st, "depit": {"modifications": res.info["n_modifications"], "sweeps": res.info["depit_sweeps"]}, "drainage_wavefronts": res.info["n_waves"],
        "roofline": {"bound": "hbm", "achieved": n * bytes_cell / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": n * bytes_cell / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": top, "note": "algorithmic bytes (%d B/cell) over the whole device part; the dominant stage `%s` is a chain of dependent steps by definition of its result "
                                            "(DESIGN section 9), so the HBM roof is not what bounds it" % (bytes_cell, top)},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu:
        ref = _pycppp_seconds(d)
        if ref and ref["ok"]:
            out["cpu_baseline"] = {"value": n / ref["seconds"], "unit": "cells/s", "cores": 1, "kind": "reference",
                                   "sample": "the reference's own ELF pycppp on the same hap.in + dtm_13.val, whole DEM, %.1f s; files identical to ours: %s" % (ref["seconds"], ref["files_identical_to_ours"])}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "cells/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref/bin/pycppp not staged on this box"}
    shutil.rmtree(d, ignore_errors=True)
    return out


def run_prepro_reference(args, size):
    d = _prepro_project(size)
    times = []
    for _ in range(max(1, min(args.steps, 2))):
        ref = _pycppp_seconds(d)
        if not ref or not ref["ok"]:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/pycppp not staged or failed"}), flush=True)
            return
        times.append(ref["seconds"])
    n = size[0] * size[1]
    v = n / float(np.mean(times))
    print(json.dumps({"impl": "reference", "metric": "pre-processor cells/s (hap.in + dtm_13.val -> routing rasters)", "value": v, "unit": "cells/s", "n_gpus": 1,
                      "steps": len(times), "warmup": 0, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f64", "data": "synthetic", "config": {"workload": "the reference ELF pycppp on the synthetic %dx%d DEM, text in, text out" % (size[0], size[1])},
                      "cpu_baseline": {"value": v, "unit": "cells/s", "cores": 1, "kind": "reference", "sample": "whole DEM, %d run(s)" % len(times)},
                      "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    shutil.rmtree(d, ignore_errors=True)


def run_ours(args, size):
    ctx = Ctx()
    elapsed = lambda: time.perf_counter() - T_START      # noqa: E731
    if args.workload == "prepro":
        if ctx.world > 1:
            raise SystemExit("bench.py: the pre-processor is a one-off set-up step of one DEM: replicas only, run it with --gpus 1")
        out = run_prepro(ctx, args, size)
    elif args.workload == "enkf":
        out = run_enkf(ctx, args, size, cycles=args.steps, warm=args.warmup)
        out.update({"vs_baseline": None, "dtype": "f64", "data": "synthetic"})
    elif args.workload == "partitioned":
        out = run_partitioned(ctx, args, size, steps=args.steps, warm=args.warmup)
        out.update({"vs_baseline": None, "dtype": "f64", "data": "synthetic"})
    else:
        out = run_headline(ctx, args, size)
        legs = [] if args.sharded == "off" else (["enkf", "partitioned"] if args.sharded == "auto" else args.sharded.split(","))
        if args.workload != "picard":
            legs = []
        if legs:
            out["sharded"] = {"note": "measured in this invocation at n_gpus = %d; strong scaling (fixed total work), so the driver's 1/2/4/8 runs give the curve" % ctx.world}
        for leg in legs:
            # a leg whose expected cost (mostly host-side set-up, ~1/world) does not fit the remaining budget is skipped and says so
            est = (40.0 + 130.0 / ctx.world) if leg == "enkf" else (30.0 + 100.0 / ctx.world)
            go = ctx.rmin(1.0 if elapsed() + est < args.budget_s else 0.0) > 0.5
            if not go:
                out["sharded"][leg] = {"skipped": "time budget (--budget-s %.0f, %.0f s elapsed)" % (args.budget_s, elapsed())}
                continue
            t_leg = time.perf_counter()
            try:
                if leg == "enkf":
                    r = run_enkf(ctx, args, (100, 100, 15), cycles=2, warm=1)
                    out["sharded"]["enkf_member_steps_per_s"] = r["value"]
                    out["sharded"]["analysis_ms"] = r["analysis_ms"]
                else:
                    psize = tuple(int(v) for v in args.partition_size.lower().split("x"))
                    r = run_partitioned(ctx, args, psize, steps=3, warm=2)
                    out["sharded"]["partitioned_node_timesteps_per_s"] = r["value"]
                r["leg_wall_s"] = time.perf_counter() - t_leg
                out["sharded"][leg] = r
            except Exception as e:      # noqa: BLE001 -- a failed leg must not lose the headline line
                out["sharded"][leg] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
            ctx.torch.cuda.empty_cache()
        if ctx.world == 1 and args.workload == "picard" and not args.quick:
            try:
                out["roofline_hbm"] = run_roofline_hbm(ctx, args)
            except Exception as e:      # noqa: BLE001
                out["roofline_hbm"] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
            try:
                out["full_run"] = run_full(ctx, args, size)
            except Exception as e:      # noqa: BLE001
                out["full_run"] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if ctx.world == 1 and args.workload == "picard" and (not args.quick or args.prepro_leg):
            # SURVEY 8f-3 in the driver's own record: one short pass of `--workload prepro` (1000x1000 DEM, the reference ELF timed beside it)
            try:
                r = run_prepro(ctx, argparse.Namespace(steps=3, warmup=1, no_cpu=args.no_cpu), (1000, 1000, 1))
                out["prepro"] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "stage_ms", "depit", "drainage_wavefronts", "gpu_launches", "cpu_baseline") if k in r}
                out["prepro"]["workload"] = r["config"]["workload"]
            except Exception as e:      # noqa: BLE001
                out["prepro"] = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
            newton = args.workload in ("newton", "coupled")
            out["cpu_baseline"] = cpu_baseline(size, budget_s=args.cpu_budget, iopt=2 if newton else 1, routing=args.workload == "coupled")
    out["bench_wall_s"] = elapsed()
    if ctx.rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default=None)
    ap.add_argument("--workload", default="picard", choices=["picard", "newton", "coupled", "enkf", "partitioned", "prepro"], help="picard: BASELINE config 2 (headline); newton: config 3's linearisation on the same mesh; coupled: config 3 (Newton + surface routing); enkf: config 4; partitioned: config 5")
    ap.add_argument("--sharded", default="auto", help="picard workload: which sharded legs run in the same invocation (auto = enkf,partitioned; off)")
    ap.add_argument("--partition-size", default="1000x1000x30", help="mesh of the `partitioned` leg (BASELINE config 5)")
    ap.add_argument("--members", type=int, default=256)
    ap.add_argument("--concurrent", type=int, default=4, help="enkf workload: ensemble members advancing concurrently per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip roofline_hbm, full_run and the pre-processor leg")
    ap.add_argument("--prepro-leg", action="store_true", help="run the pre-processor leg even with --quick")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--budget-s", type=float, default=540.0, help="wall-clock budget of the whole invocation; sharded legs that do not fit are skipped")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" and args.workload not in ("enkf", "partitioned") else args.warmup
    if args.size is None:
        args.size = {"picard": "200x200x20", "newton": "200x200x20", "coupled": "200x200x20", "enkf": "100x100x15", "partitioned": "1000x1000x30", "prepro": "1000x1000x1"}[args.workload]
    size = tuple(int(v) for v in args.size.lower().split("x"))
    if args.impl == "reference":
        run_reference(args, size)
    else:
        run_ours(args, size)


if __name__ == "__main__":
    main()
