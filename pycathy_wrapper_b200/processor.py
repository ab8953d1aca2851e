"""The `./cathy` replacement: run one CATHY forward simulation in a project directory.

Mirrors PROGRAM CATHY_MAIN's observable behaviour (SRC/cathy_main.f:2495-3945) at the
process boundary pyCATHY uses (pyCATHY/cathy_tools.py:593-740 run_processor;
:76-98 subprocess_run_multi): read ./cathy.fnames + input/* + prepro/*, write output/*.
All numerical work happens on the GPU behind libcathy_b200.so (capi.load_library());
this module only does text I/O, the cumulative book-keeping of the mass-balance columns
and the DETOUT scheduling.

File-clobbering rules of the reference are kept: a normal run leaves an
existing output/grid3d and output/xyz untouched; IPRT1=3 writes only the mesh files.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import outputs as O
from .capi import CathyLib, Simulation, load_library
from .project import CathyProject, load_project


class RunResult:
    def __init__(self):
        self.reports = []          # light-weight dict per accepted step
        self.nstep = 0
        self.wall_loop = 0.0       # seconds spent in the time loop (steps only)
        self.wall_setup = 0.0
        self.wall_io = 0.0
        self.gpu_ms = 0.0
        self.launches = 0
        self.n = 0
        self.finished_ok = True
        self.final_state = None
        self.solver_limits = None  # effective (itmax, tol) of the device linear solver, also written to output/risul


def _out(prj: CathyProject, unit: str) -> str:
    path = prj.fnames[unit]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    return path


def run_processor(project_dir: str, lib: CathyLib | None = None, write_files: bool = True, verbose: bool = False,
                  max_steps: int | None = None, keep_reports: bool = True, prj: CathyProject | None = None,
                  **overrides) -> RunResult:
    """Run the simulation described by `project_dir`.  `overrides` are parm entries
    (e.g. TMAX=600, DELTAT=1) applied on top of input/parm, like pyCATHY's
    run_processor(**kwargs) -> update_parm."""
    t_start = time.perf_counter()
    if lib is None:
        lib = load_library()                       # raises if the CUDA library is missing: no CPU path
    if prj is None:
        prj = load_project(project_dir)
    sim_kw = {k: overrides.pop(k) for k in ("precond", "device", "tolcg_scale", "itmxcg_scale") if k in overrides}
    sim = Simulation(lib, prj, **sim_kw, **overrides)
    parm = sim.parm
    res = RunResult()
    res.n = sim.n
    nnod, n, nstr = sim.nnod, sim.n, prj.nstr

    if parm["IPRT1"] == 3:                          # mesh-only mode, SRC/gen3d.f:89-112
        x, y, z, tet = sim.mesh()
        if write_files:
            O.write_xyz(_out(prj, "IOUT3"), nnod, n, x, y, z)
            O.write_grid3d(os.path.join(os.path.dirname(_out(prj, "IOUT3")), "grid3d"), nnod, n, sim.nt, tet, x, y, z)
            with open(_out(prj, "IOUT2"), "w") as fh:
                fh.write("\n\n IPRT1=3: Program terminating after output of X, Y, Z coordinate values\n")
        sim.close()
        return res

    lim = sim.solver_limits()
    if write_files and lim is not None:
        # the stopping rule the device solver really applies (ITMXCG x itmxcg_scale, TOLCG x tolcg_scale; include/cathy_b200.h) goes on
        # record in the run log, next to the parm values the user wrote
        with open(_out(prj, "IOUT2"), "w") as fh_r:
            fh_r.write(" cathy-b200 linear solver: %s preconditioner, ITMXCG = %d (parm %d x %g), TOLCG = %.3E (parm %.3E x %g)\n"
                       % (lim["preconditioner"], lim["itmax"], parm["ITMXCG"], lim["itmxcg_scale"], lim["tol"], parm["TOLCG"], lim["tolcg_scale"]))
    res.solver_limits = lim

    if getattr(prj, "transport_skipped", False):
        note = (" TRAFLAG=1: cathy-b200 runs the FLOW problem of this project only; the solute-transport add-on (one-way coupled, "
                "SRC/cathy_main.f:3304-3607) is not simulated and no concentration output is written\n")
        print(note, end="")
        if write_files:
            with open(_out(prj, "IOUT2"), "a") as fh_r:
                fh_r.write(note)

    x = y = z = None
    fh = {}
    if write_files:
        x, y, z, _ = sim.mesh(with_tetra=False)
        for key, unit in (("psi", "IOUT11"), ("sw", "IOUT13"), ("vp", "IOUT6"), ("mbeconv", "IOUT5"),
                          ("cumflowvol", "IOUT36"), ("iter", "IOUT4"), ("hgraph", "IOUT41"), ("pondhead", "IOUT42")):
            fh[key] = open(_out(prj, unit), "w")
        aux = (("hgatmsf", "IOUT7"), ("hgnansf", "IOUT8"), ("hgflag", "IOUT9"), ("velnod", "IOUT12"), ("velelt", "IOUT15"), ("psisurf", "IOUT16"),
               ("satsurf", "IOUT17"), ("swsurf", "IOUT18"), ("hgsfdet", "IOUT30"), ("hgnansfdirdet", "IOUT31"), ("hgnansfneudet", "IOUT32"),
               ("dtcoupling", "IOUT43"), ("recharge", "IOUT44"), ("wtdepth", "IOUT57"))
        for key, unit in aux:
            if unit in prj.fnames:
                fh[key] = open(_out(prj, unit), "w")
        fh["fort777"] = open(os.path.join(os.path.dirname(os.path.dirname(_out(prj, "IOUT11"))), "fort.777"), "w")   # unit 777: unnamed, lands in the cwd
        for key, head in (("hgatmsf", O.HGATMSF_HEADER), ("hgnansf", O.HGNANSF_HEADER), ("hgsfdet", O.HGSFDET_HEADER),
                          ("hgnansfdirdet", O.HGNANSFDIR_HEADER), ("hgnansfneudet", O.HGNANSFNEU_HEADER), ("wtdepth", O.WTDEPTH_HEADER)):
            if key in fh:
                fh[key].write(head)
        if parm["IPRT"] >= 4:
            if "velelt" in fh:
                fh["velelt"].write("  0   HSPVEL\n")
            fh["sw"].write("  0   HSPSW\n")
        fh["iter"].write(O.iter_header(parm))
        fh["mbeconv"].write(O.MBECONV_HEADER % O.fe(sim.initial_storage(), 13, 5))
        fh["cumflowvol"].write(O.CUMFLOWVOL_HEADER)
        fh["hgraph"].write("#%s\n" % O.fi(parm["NUM_QOUT"] + 1, 8))
        if parm["ISIMGR"] == 2:
            fh["hgraph"].write("#          TIME %s\n" % "".join(O.fi(v, 16) for v in [int(prj.surf["qoi"][-1])] + parm["ID_QOUT"]))

    vtk_state = {"tet0": None, "ks": None}
    area_cache = {}

    def arenod():
        """Nodal surface areas (SRC/area2d.f): a third of the area of every adjacent triangle, summed in triangle order."""
        if "a" not in area_cache:
            nc1 = prj.ncol + 1
            a = np.zeros(nnod)
            are3 = abs(0.5 * prj.dx * prj.dy) * (1.0 / 3.0)
            for i in range(prj.nrow):
                for j in range(prj.ncol):
                    n00 = i * nc1 + j
                    for nd in (n00, n00 + nc1, n00 + nc1 + 1, n00, n00 + 1, n00 + nc1 + 1):
                        a[nd] += are3
            area_cache["a"] = a
        return area_cache["a"]

    def vtkout(unit, tim, st):
        """VTKRIS3D (SRC/vtkris3d.f), called with unit 100 at time 0 and 100+KPRT at the detailed outputs when IPRT >= 2 and
        VTKF > 0 (SRC/cathy_main.f:2833-2837, 3727-3729, 3860-3862); files land in ./vtk like the reference's."""
        vtkf = int(parm.get("VTKF", 0))
        if parm["IPRT"] < 2 or vtkf <= 0:
            return
        if vtk_state["tet0"] is None:
            _, _, _, tet = sim.mesh()
            t0 = tet[:, :4].astype(np.int64) - 1
            if int(parm["IOPT"]) == 1:
                t0 = np.sort(t0, axis=1)                  # element nodes are kept sorted under Picard (SRC/grdsys.f:63)
            vtk_state["tet0"] = t0
            ntri3 = 3 * 2 * prj.nrow * prj.ncol
            lay = np.arange(sim.nt) // ntri3
            vtk_state["ks"] = prj.soil["TABLE"][lay, tet[:, 4] - 1, 0]   # KS(J) = PERMX(layer, zone), SRC/cathy_main.f:2688-2691
        vel = None
        if vtkf >= 4:
            v = sim.velocity(nodal=False)
            vel = (v["uu"], v["vv"], v["ww"])
        vdir = os.path.join(os.path.dirname(os.path.dirname(_out(prj, "IOUT11"))), "vtk")
        os.makedirs(vdir, exist_ok=True)
        O.write_vtk(os.path.join(vdir, "%3d.vtk" % unit), tim, x, y, z, vtk_state["tet0"], st["psi"], st["sw"], vtk_state["ks"], vel, vtkf)

    def detout(nstep, tim, vtk_unit=None):
        if not write_files:
            return
        st = sim.state()
        if vtk_unit is not None:
            vtkout(vtk_unit, tim, st)
        if parm["IPRT"] >= 1:
            O.write_block(fh["psi"], nstep, tim, st["psi"])
            O.write_block(fh["pondhead"], nstep, tim, st["pond"])
            if parm["IPRT"] >= 4:
                O.write_block(fh["sw"], nstep, tim, st["sw"])
        if parm["NUMVP"] > 0:
            O.write_vp(fh["vp"], nstep, tim, parm["NODVP"], nnod, nstr, x, y, z, st["psi"], st["sw"], st["ckrw"],
                       st["qtranie"])
        # SRC/detout.f:34-43 velocities, :99-122 surface tables (psisurf, satsurf, swsurf, recharge, fort.777)
        vel = None
        if parm["IPRT"] >= 2 and ("velnod" in fh or "velelt" in fh):
            vel = sim.velocity(nodal=True)
            if "velnod" in fh and parm["IPRT"] >= 1:
                O.write_velnod(fh["velnod"], nstep, tim, vel["unod"], vel["vnod"], vel["wnod"])
            if "velelt" in fh and parm["IPRT"] >= 4:
                O.write_velelt(fh["velelt"], tim, vel["uu"], vel["vv"], vel["ww"])
        psi2 = st["psi"].reshape(nstr + 1, nnod)
        satsur = np.where(psi2[0] >= 0.0, np.where(np.any(psi2[1:] < 0.0, axis=0), 2, 3), 1)
        eta = st["qtranie"].reshape(nstr + 1, nnod)
        etasum = np.zeros(nnod)
        for k in range(nstr + 1):
            etasum = etasum + eta[k]
        rec = sim.recharge()[0] if parm["IPRT"] >= 2 else np.zeros(nnod)
        area = arenod()
        for key, title, vals, integer in (("psisurf", "PRESSURE HEAD", psi2[0], False), ("satsurf", "SATSUR", satsur, True),
                                          ("swsurf", "SW", st["sw"][:nnod], False), ("recharge", "REC. FLUX", rec / area, False),
                                          ("fort777", "ACT. ETRA", etasum / area, False)):
            if key in fh:
                O.write_surface_table(fh[key], nstep, tim, title, x, y, vals, integer)

    detout(0, 0.0, vtk_unit=100)
    res.wall_setup = time.perf_counter() - t_start

    cum = dict(VSFTOT=0.0, VNDTOT=0.0, VNNTOT=0.0, VNUDTOT=0.0, VTOT=0.0, CVIN=0.0, CVOUT=0.0, CDSTOR=0.0,
               CERRAS=0.0, CAERAS=0.0)
    kprt = 1
    tot = dict(recvol=0.0, dtc_head=False, cpusub=0.0, vapot=0.0, vaact=0.0, nsurf=0, nsurft=0)
    if write_files and parm["IPRT"] >= 2 and "hgatmsf" in fh:
        # recharge at the initial conditions already counts into RECVOL with the first DELTAT (SRC/cathy_main.f:2814, SRC/recharge.f:45)
        tot["recvol"] = sim.recharge()[1] * float(parm["DELTAT"])
    nprt, timprt = parm["NPRT"], parm["TIMPRT"]
    last = None
    while True:
        t0 = time.perf_counter()
        rep = sim.step()
        res.wall_loop += time.perf_counter() - t0
        res.gpu_ms += rep.gpu_ms
        res.launches += rep.launches
        last = rep
        t1 = time.perf_counter()
        cum["VSFTOT"] += rep.vsfflw
        cum["VNDTOT"] += rep.vndin + rep.vndout
        cum["VNNTOT"] += rep.vnnin + rep.vnnout
        cum["VTOT"] += rep.vin + rep.vout
        cum["CVIN"] += rep.vin
        cum["CVOUT"] += rep.vout
        cum["CDSTOR"] += rep.dstore
        cum["CERRAS"] += rep.erras
        cum["CAERAS"] += abs(rep.erras)
        if keep_reports:
            res.reports.append(dict(nstep=rep.nstep, deltat=rep.deltat, time=rep.time, iter=rep.iter,
                                    nitert=rep.nitert, kbackt=rep.kbackt, nsurf=rep.nsurf, store1=rep.store1,
                                    dstore=rep.dstore, vin=rep.vin, vout=rep.vout, erras=rep.erras, errel=rep.errel,
                                    pinf=[rep.it[k].pinf for k in range(rep.n_iter_rec)],
                                    niter=[rep.it[k].niter for k in range(rep.n_iter_rec)],
                                    q_outlet=rep.q_outlet_1 + rep.q_outlet_2, gpu_ms=rep.gpu_ms))
        if write_files:
            # the reference prints one (NSTEP..) header per attempt: the back-stepped ones first (cathy_attempt_log)
            if rep.kbackt > 0:
                for dt_a, time_a, recs in sim.attempt_log():
                    fh["iter"].write(O.iter_step_line(rep.nstep, dt_a, time_a))
                    for k, r in enumerate(recs):
                        fh["iter"].write(O.iter_line(k + 1, r))
            fh["iter"].write(O.iter_step_line(rep.nstep, rep.deltat, rep.time))
            for k in range(rep.n_iter_rec):
                fh["iter"].write(O.iter_line(k + 1, rep.it[k]))
            fh["mbeconv"].write(O.mbeconv_line(
                rep.nstep, rep.deltat, rep.time, rep.iter, float(rep.nitert) / float(rep.iter), rep.store1,
                rep.store2, rep.dstore, cum["CDSTOR"], rep.vin, cum["CVIN"], rep.vout, cum["CVOUT"],
                rep.vin + rep.vout, cum["VTOT"], rep.erras, rep.errel, cum["CERRAS"], cum["CAERAS"]))
            fh["cumflowvol"].write(O.cumflowvol_line(rep.nstep, rep.deltat, rep.time, cum["VSFTOT"], cum["VNDTOT"],
                                                     cum["VNNTOT"], cum["VNUDTOT"], cum["VTOT"]))
            if parm["ISIMGR"] == 2:
                fh["hgraph"].write("".join(O.fe(v, 16, 8) for v in (rep.time, rep.q_outlet_1, rep.q_outlet_2, 0.0, 0.0)) + "\n")
            # SRC/cathy_main.f:3628, 3658-3695: water-table depth, recharge, detailed hydrograph files, coupling diagnostics
            if parm["NUMVP"] > 0 and "wtdepth" in fh:
                fh["wtdepth"].write(O.wtdepth_line(rep.time, sim.wtdepth(parm["NODVP"])))
            recflow = sim.recharge()[1] if (parm["IPRT"] >= 2 and "hgatmsf" in fh) else 0.0
            tot["recvol"] += recflow * rep.deltat
            if "hgatmsf" in fh:
                fh["hgatmsf"].write(O.hgatmsf_line(rep, recflow, tot["recvol"]))
            if "hgnansf" in fh:
                fh["hgnansf"].write(O.hgnansf_line(rep))
            for key, vol in (("hgsfdet", rep.vsfflw), ("hgnansfdirdet", rep.vndin + rep.vndout), ("hgnansfneudet", rep.vnnin + rep.vnnout)):
                if key in fh:
                    fh[key].write(O.det_line(rep, vol))
            if "dtcoupling" in fh:
                if not tot["dtc_head"]:
                    fh["dtcoupling"].write(O.dtcoupling_header(parm["ISIMGR"] == 2, nnod, prj.nrow * prj.ncol, rep.areatot))
                    tot["dtc_head"] = True
                cpusub = time.perf_counter() - t0
                fh["dtcoupling"].write(O.dtcoupling_line(rep, parm["ITUNS"], cpusub, 0.0))
                tot["cpusub"] += cpusub
                tot["vapot"] += rep.apot * rep.deltat
                tot["vaact"] += 0.5 * (rep.aact + rep.aact_prev) * rep.deltat
                tot["nsurf"] += rep.nsurf
                tot["nsurft"] += rep.nsurft
        if verbose:
            print(" TIME STEP: %6d  DELTAT: %12.4E  TIME: %12.4E  NL its %2d  lin its %4d  back-steps %d"
                  % (rep.nstep, rep.deltat, rep.time, rep.iter, rep.nitert, rep.kbackt), flush=True)
        if nprt > 0 and kprt <= nprt and rep.time >= timprt[kprt - 1]:
            detout(rep.nstep, rep.time, vtk_unit=100 + kprt)
            kprt += 1
        res.wall_io += time.perf_counter() - t1
        if rep.finished or (max_steps is not None and rep.nstep >= max_steps):
            break
    res.nstep = last.nstep
    res.finished_ok = not bool(last.noback)
    t1 = time.perf_counter()
    if max_steps is None:
        detout(last.nstep, parm["TMAX"], vtk_unit=100 + kprt)   # label 300: final DETOUT always carries TIME=TMAX
    res.final_state = sim.state()
    if write_files and last is not None:
        if "dtcoupling" in fh and tot["dtc_head"]:
            fh["dtcoupling"].write(O.dtcoupling_footer(last.kback_total, last.itrtot, parm["ITUNS"], tot["nsurf"], tot["nsurft"], tot["vapot"],
                                                       tot["vaact"], last.areatot, tot["cpusub"], 0.0))
        if "hgflag" in fh:
            fh["hgflag"].write(O.hgflag_text(list(last.hgflag)))
    for f in fh.values():
        f.close()
    res.wall_io += time.perf_counter() - t1
    sim.close()
    return res


def main(argv=None) -> int:
    """Entry used by the `cathy` launcher script: no arguments, cwd = project directory."""
    argv = sys.argv[1:] if argv is None else argv
    prj_dir = argv[0] if argv else os.getcwd()
    res = run_processor(prj_dir, verbose=True)
    return 0 if res.finished_ok else 1


if __name__ == "__main__":
    sys.exit(main())
