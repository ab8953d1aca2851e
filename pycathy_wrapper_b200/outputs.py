"""Writers for the CATHY output files pyCATHY parses (the drop-in boundary, output side).

Formats follow the reference FORMAT statements character for character so that
pyCATHY's readers (pyCATHY/importers/cathy_outputs.py: read_psi :480, read_sw :418,
read_vp :255, read_mbeconv :607, read_cumflowvol :223, read_grid3d :15) parse them
unchanged:
  psi/sw      SRC/detout.f:126-128   (I7,1PE16.8,'     NSTEP   TIME') + 5(1PE15.6)
  vp          SRC/detout.f:54-67     FORMAT 1070/1080
  mbeconv     SRC/cathy_main.f:3269  FORMAT 1220 (header) / 1240
  cumflowvol  SRC/cathy_main.f:3677  FORMAT 1060 (mbinit.f) / 1199
  iter        SRC/cathy_main.f:4010-4030 FORMAT 1060/1065, SRC/conver.f FORMAT 1070
  hgraph      SRC/cathy_main.f:4135-4136 FORMAT 1530/1540, SRC/detoutq.f FORMAT 1530
  grid3d/xyz  SRC/gen3d.f:96-102,1010-1050
"""
from __future__ import annotations

import numpy as np


def fe(x: float, w: int, d: int) -> str:
    """Fortran 1PEw.d edit descriptor."""
    s = "%.*E" % (d, x)
    mant, ex = s.split("E")
    e = int(ex)
    if abs(e) >= 100:                        # 3-digit exponents drop the 'E'
        s = "%s%+04d" % (mant, e)
    out = s.rjust(w)
    return out if len(out) <= w else "*" * w


def fi(i: int, w: int) -> str:
    s = str(int(i)).rjust(w)
    return s if len(s) <= w else "*" * w


def write_block(fh, nstep: int, time: float, values: np.ndarray) -> None:
    """DETOUT block: header + values 5 per line in 1PE15.6."""
    fh.write("%s%s     NSTEP   TIME\n" % (fi(nstep, 7), fe(time, 16, 8)))
    v = np.asarray(values, dtype=np.float64)
    txt = np.char.mod("%15.6E", v)
    big = np.abs(v) < 1e-99
    if np.any(big & (v != 0)) or np.any(np.abs(v) >= 1e100):
        txt = np.array([fe(float(t), 15, 6) for t in v])
    n = len(txt)
    full = (n // 5) * 5
    if full:
        rows = txt[:full].reshape(-1, 5)
        fh.write("\n".join("".join(r) for r in rows))
        fh.write("\n")
    if n > full:
        fh.write("".join(txt[full:]) + "\n")


MBECONV_HEADER = (
    "#INITIAL VOLUME OF WATER IN THE SUBSURFACE = %s\n"
    "#NSTEP       DELTAT         TIME  NLIN   AVG.LIN       STORE1       STORE2       DSTORE"
    "   CUM.DSTORE       VIN  CUM.VIN       VOUT    CUM.VOUT   VIN+VOUT   CUM. VIN  M.BAL.ERR"
    "   REL. MBE   CUM. MBE      CUM.\n"
    "#                                 ITER.    ITER.                             (>0 x inc)"
    "                                            "
    "                          + VOUT  (Vi+Vo-Ds)        (%%)             ABS(MBE)\n")


def mbeconv_line(nstep, deltat, time, it, avglin, store1, store2, dstore, cdstor, vin, cvin, vout, cvout,
                 vtot_step, vtot, erras, errel, cerras, caeras) -> str:
    return (fi(nstep, 6) + fe(deltat, 13, 6) + fe(time, 13, 6) + fi(it, 6) + fe(avglin, 10, 3)
            + fe(store1, 13, 6) + fe(store2, 13, 6) + fe(dstore, 13, 5) + fe(cdstor, 13, 5)
            + fe(vin, 10, 3) + fe(cvin, 10, 3)
            + "".join(fe(v, 11, 3) for v in (vout, cvout, vtot_step, vtot, erras, errel, cerras))
            + fe(caeras, 10, 3) + "\n")


CUMFLOWVOL_HEADER = (" " * 21 + "#***** Cumulative flow volumes ***** \n"
                     "# Nstep    Deltat      Time  SeepageF  nansfDir  nansfNeu   Nudging  net (VTOT)\n")


def cumflowvol_line(nstep, deltat, time, vsftot, vndtot, vnntot, vnudtot, vtot) -> str:
    return (fi(nstep, 7) + "".join(fe(v, 10, 2) for v in (deltat, time, vsftot, vndtot, vnntot, vnudtot))
            + fe(vtot, 12, 4) + "\n")


def iter_header(parm: dict) -> str:
    return ("     IOPT   (1 PICARD, 2 NEWTON)             = %s\n"
            "     NLRELX (0 NORELX,1 CONS RELX,2 VAR RELX)= %s\n"
            "     KSLOPE (0 ANA, 1 KSL/ANA, 2 KSL/C-DIFF,\n"
            "             3 LOC KSL/ANA, 4 LOC TAN-SLOPE) = %s\n"
            "     TOLUNS (TOLERANCE FOR NONLINEAR ITER)   = %s\n"
            "     TOLSWI (TOLERANCE FOR BC SWITCHING)     = %s\n"
            "     L2NORM (0 INFINITY NORM, ELSE L2 NORM)  = %s\n"
            "\n"
            " nlinr  linr converg error norms  node   PNEW at   POLD at   resid error norms\n"
            "  iter  iter       PL2      PINF IKMAX     IKMAX     IKMAX       FL2      FINF\n"
            " =============================================================================\n"
            % (fi(parm["IOPT"], 6), fi(parm["NLRELX"], 6), fi(parm["KSLOPE"], 6), fe(parm["TOLUNS"], 15, 5),
               fe(parm["TOLSWI"], 15, 5), fi(parm["L2NORM"], 6)))


def iter_step_line(nstep, deltat, time) -> str:
    return " " * 23 + " (NSTEP: %s  DELTAT: %s  TIME: %s)\n" % (fi(nstep, 5), fe(deltat, 12, 4), fe(time, 12, 4))


def iter_line(k, rec) -> str:
    return (fi(k, 6) + fi(rec.niter, 6) + fe(rec.pl2, 10, 3) + fe(rec.pinf, 10, 3) + fi(rec.ikmax, 6)
            + fe(rec.pnew_ik, 10, 2) + fe(rec.pold_ik, 10, 2) + fe(rec.fl2, 10, 3) + fe(rec.finf, 10, 3) + "\n")


def write_vp(fh, nstep, time, nodvp, nnod, nstr, x, y, z, psi, sw, ckrw, qtranie) -> None:
    fh.write("%s%s     NSTEP   TIME\n" % (fi(nstep, 7), fe(time, 16, 8)))
    for inod in nodvp:
        if 1 <= inod <= nnod:
            fh.write(" SURFACE NODE = %s  X = %s  Y = %s\n" % (fi(inod, 5), fe(x[inod - 1], 12, 4), fe(y[inod - 1], 12, 4)))
            fh.write("              Z  PRESSURE HEAD             SW           CKRW        QTRANIE\n")
            for k in range(nstr + 1):
                kk = k * nnod + inod - 1
                fh.write("".join(fe(v, 15, 6) for v in (z[kk], psi[kk], sw[kk], ckrw[kk], qtranie[kk], 0.0)) + "\n")


def write_grid3d(path, nnod, n, nt, tetra, x, y, z) -> None:
    with open(path, "w") as fh:
        fh.write("%s%s%s\n" % (fi(nnod, 9), fi(n, 9), fi(nt, 9)))
        np.savetxt(fh, tetra, fmt="%7d", delimiter="")
        np.savetxt(fh, np.column_stack([x, y, z]), fmt="%15.6E", delimiter="")


def write_xyz(path, nnod, n, x, y, z) -> None:
    with open(path, "w") as fh:
        fh.write("%s%s  NNOD   N\n" % (fi(nnod, 7), fi(n, 7)))
        idx = np.arange(1, n + 1)
        for i, a, b, c in zip(idx, x, y, z):
            fh.write("%7d%15.6E%15.6E%15.6E\n" % (i, a, b, c))


# ---------------------------------------------------------------------------------------------
# vtk/1NN.vtk (SRC/vtkris3d.f): legacy ASCII unstructured grid with pressure, saturation (VTKF >= 2), element
# conductivity (VTKF >= 3) and element Darcy velocity (VTKF >= 4).  Header lines use FORMATs 78/77/79/80/81/82-85, the value
# lists are LIST-DIRECTED writes, reproduced here with gfortran's rules for REAL(8) / INTEGER(4).
# ---------------------------------------------------------------------------------------------
def ld_real(x: float) -> str:
    """gfortran list-directed REAL(8): 17 significant digits; F editing in a 21-column field + 5 blanks when
    0.1 <= |x| < 1e17 (after rounding), ES26.17E3 otherwise; zero prints as 0.0000000000000000."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return ("NaN" if x != x else ("Infinity" if x > 0 else "-Infinity")).rjust(26)
    if x == 0.0:
        return "  -0.0000000000000000     " if np.signbit(x) else "   0.0000000000000000     "
    mant, ex = ("%.16E" % x).split("E")
    k = int(ex) + 1                                   # 10**(k-1) <= |x| < 10**k after rounding to 17 digits
    if 0 <= k <= 17:
        return ("%.*f" % (17 - k, x)).rjust(21) + "     "
    return ("%sE%+04d" % (mant, int(ex))).rjust(26)


def _ld_reals(v: np.ndarray) -> np.ndarray:
    """Vectorised ld_real for an array (one 26-character token per value)."""
    v = np.asarray(v, dtype=np.float64).ravel()
    out = np.empty(v.shape, dtype="U26")
    fin = np.isfinite(v) & (v != 0.0)
    ex = np.zeros(v.shape, dtype=np.int64)
    es = np.char.mod("%.16E", v[fin])
    ex[fin] = np.array([int(t[-4:]) if t[-4] in "+-" else int(t.split("E")[1]) for t in es], dtype=np.int64) if es.size else ex[fin]
    k = ex + 1
    use_f = fin & (k >= 0) & (k <= 17)
    for kk in np.unique(k[use_f]):
        m = use_f & (k == kk)
        out[m] = np.char.add(np.char.rjust(np.char.mod("%%.%df" % (17 - int(kk)), v[m]), 21), "     ")
    m = fin & ~use_f
    if np.any(m):
        out[m] = [ld_real(t) for t in v[m]]
    m = ~fin
    if np.any(m):
        out[m] = [ld_real(t) for t in v[m]]
    return out


def write_vtk(path: str, time: float, x, y, z, tetra0: np.ndarray, psi, sw, ks=None, vel=None, vtkf: int = 1) -> None:
    """tetra0: [NT][4] 0-based node ids in the processor's stored order (sorted ascending under Picard)."""
    n, nt = len(x), len(tetra0)
    with open(path, "w") as fh:
        fh.write("# vtk DataFile Version 2.0\n3D Unstructured Grid of Linear Triangles\nASCII\n")
        fh.write("DATASET UNSTRUCTURED_GRID\nFIELD FieldData  1\nTIME 1 1 double\n%18.5f\n" % time)
        fh.write("POINTS %8d float\n" % n)
        pts = np.char.add(np.char.add(np.char.mod("%16.8E", x), np.char.mod("%16.8E", y)), np.char.mod("%16.8E", z))
        fh.write("\n".join(np.char.add("    ", pts)))
        fh.write("\nCELLS %8d %8d\n" % (nt, nt * 5))
        t = np.asarray(tetra0)
        cells = np.char.add("4", np.char.add(np.char.add(np.char.mod("   %8d", t[:, 0]), np.char.mod("   %8d", t[:, 1])),
                                              np.char.add(np.char.mod("   %8d", t[:, 2]), np.char.mod("   %8d", t[:, 3]))))
        fh.write("\n".join(cells))
        fh.write("\nCELL_TYPES%8d\n" % nt)
        fh.write(("          10\n") * nt)
        fh.write("POINT_DATA %8d\nSCALARS pressure float\nLOOKUP_TABLE default\n" % n)
        fh.write("\n".join(_ld_reals(psi)))
        fh.write("\n")
        if vtkf >= 2:
            fh.write("SCALARS saturation float\nLOOKUP_TABLE default\n")
            fh.write("\n".join(_ld_reals(sw)))
            fh.write("\n")
        if vtkf >= 3:
            fh.write("CELL_DATA %8d\nSCALARS permeability float\nLOOKUP_TABLE default\n" % nt)
            fh.write("\n".join(_ld_reals(ks)))
            fh.write("\n")
        if vtkf >= 4:
            fh.write("VECTORS velocity float\n")
            a, b, c = (_ld_reals(v) for v in vel)
            fh.write("\n".join(np.char.add(np.char.add(a, b), c)))
            fh.write("\n")


# ---------------------------------------------------------------------------------------------
# Auxiliary per-step and per-DETOUT outputs pyCATHY plots or maps (read_hgatmsf, read_dtcoupling, read_hgsfdet, read_wtdepth,
# read_recharge, read_fort777 in pyCATHY/importers/cathy_outputs.py).  FORMAT numbers refer to SRC/cathy_main.f unless noted.
# ---------------------------------------------------------------------------------------------
HGATMSF_HEADER = ("#NSTP            DELTAT         TIME    POT. FLUX    ACT. FLUX    OVL. FLUX    RET. FLUX    SEEP FLUX"
                  "    REC. FLUX     REC.VOL.\n")                                                     # SRC/mbinit.f FORMAT 1000
HGNANSF_HEADER = "#NSTEP     DELTAT       TIME   NET NATM,NSF DIR FLUX   NET NATM,NSF NEU FLUX\n"     # mbinit.f 1010
HGSFDET_HEADER = "# NSTEP            DELTAT              TIME  NET SEEPFACE VOL  NET SEEPFACE FLX\n"  # mbinit.f 1020
HGNANSFDIR_HEADER = "# NSTEP            DELTAT              TIME NET NANSF DIR VOL NET NANSF DIR FLX\n"
HGNANSFNEU_HEADER = "# NSTEP            DELTAT              TIME NET NANSF NEU VOL NET NANSF NEU FLX\n"
WTDEPTH_HEADER = "#TIME      <--- WTDEPTH(NODVP(I)), I=1,2,...,NUMVP --->\n"                          # FORMAT 1232
_DTC = [("Step      (1)", "Time step"), ("Deltat    (2)", "Time step size"), ("Time      (3)", "See parm input file for units"),
        ("Back      (4)", "# of back-stepping occurrences"),
        ("NL-l      (5)", "# of nonlinear iterations for the successful (last) time step"),
        ("NL-a      (6)", "# of nonlinear iterations for the successful time step and any back-steps (= NL-last + ITUNS*Back)"),
        ("Sdt-l     (7)", "# of time steps in the surface routing module for the successful (last) subsurface module time step"),
        ("Sdt-a     (8)", "# of time steps in the surface routing module for the successful subsurface module time step and any back-steps"),
        ("Atmpot-vf (9)", "Potential atmospheric forcing (rain +ve / evap -ve) as a volumetric flux [L^3/T]"),
        ("Atmpot-v (10)", "Potential atmospheric forcing volume [L^3] (See parm input file for units)"),
        ("Atmpot-r (11)", "Potential atmospheric forcing rate [L/T]"), ("Atmpot-d (12)", "Potential atmospheric forcing depth [L]"),
        ("Atmact-vf(13)", "Actual infiltration (+ve) or exfiltration (-ve) at atmospheric BC nodes as a volumetric flux [L^3/T]"),
        ("Atmact-v (14)", "Actual infiltration (+ve) or exfiltration (-ve) volume [L^3]"),
        ("Atmact-r (15)", "Actual infiltration (+ve) or exfiltration (-ve) rate [L/T]"),
        ("Atmact-d (16)", "Actual infiltration (+ve) or exfiltration (-ve) depth [L]"),
        ("Horton   (17)", "Fraction of the surface nodes that are saturated or ponded due to Horton infiltration excess (Note: based on PNEW and not on IFATM)"),
        ("Dunne    (18)", "Fraction of the surface nodes that are saturated or ponded due to Dunne saturation excess (see previous note)"),
        ("Ponded   (19)", "Fraction of the surface nodes that are ponded (PNEW > PONDH_MIN)"),
        ("Satur    (20)", "Fraction of the surface nodes that are saturated or ponded (PNEW > 0)"),
        ("CPU-sub  (21)", "CPU seconds for the subsurface flow module"), ("CPU-surf (22)", "CPU seconds for the surface routing module")]


def dtcoupling_header(surf: bool, nnod: int, ncell: int, areatot: float) -> str:
    """SRC/inital.f FORMAT 1100 (coupled runs only) + 1110."""
    s = ""
    if surf:
        s += "#" + " " * 17 + " ***** Surface vs subsurface diagnostics ***** \n"
        s += "#NNOD    (# of surface nodes)            = %s\n#NCELL   (# of DEM cells)                = %s\n" % (fi(nnod, 6), fi(ncell, 6))
        s += "#AREATOT (total catchment surface area)  = %s\n" % fe(areatot, 15, 5)
    s += "".join("#%s : %s\n" % kv for kv in _DTC)
    s += ("#   (1)       (2)       (3)   (4)   (5)   (6)   (7)   (8)        (9)       (10)       (11)       (12)       (13)"
          "       (14)       (15)       (16)   (17)   (18)   (19)   (20)       (21)       (22)\n"
          "#  Step    Deltat      Time  Back  NL-l  NL-a Sdt-l Sdt-a  Atmpot-vf   Atmpot-v   Atmpot-r   Atmpot-d  Atmact-vf"
          "   Atmact-v   Atmact-r   Atmact-d Horton  Dunne Ponded  Satur   CPU-sub  CPU-surf\n")
    return s


def dtcoupling_line(rep, ituns: int, cpusub: float, cpusurf: float) -> str:
    """FORMAT 1170: I7,2(1PE10.3),5(I6),8(1PE11.3),4(0PF7.3),2(1PE11.3)"""
    at = rep.areatot
    aav = 0.5 * (rep.aact + rep.aact_prev)
    vals = [rep.apot, rep.apot * rep.deltat, rep.apot / at, (rep.apot * rep.deltat) / at, aav, aav * rep.deltat, aav / at, (aav * rep.deltat) / at]
    return (fi(rep.nstep, 7) + fe(rep.deltat, 10, 3) + fe(rep.time, 10, 3)
            + "".join(fi(v, 6) for v in (rep.kbackt, rep.iter, rep.iter + ituns * rep.kbackt, rep.nsurf, rep.nsurft))
            + "".join(fe(v, 11, 3) for v in vals) + "".join("%7.3f" % v for v in (rep.fhort, rep.fdunn, rep.fpond, rep.fsat))
            + fe(cpusub, 11, 3) + fe(cpusurf, 11, 3) + "\n")


def dtcoupling_footer(kback, itrtot, ituns, nsurft_t, nsurft_tb, vapot_t, vaact_t, areatot, cpusub_t, cpusurf_t) -> str:
    """FORMAT 1175"""
    return ("#" + "=" * 192 + "\n#Total:" + " " * 20 + "".join(fi(v, 6) for v in (kback, itrtot, itrtot + ituns * kback, nsurft_t, nsurft_tb))
            + "".join(" " * 11 + fe(v, 11, 3) for v in (vapot_t, vapot_t / areatot, vaact_t, vaact_t / areatot)) + " " * 28
            + fe(cpusub_t, 11, 3) + fe(cpusurf_t, 11, 3) + "\n")


def hgatmsf_line(rep, recflow: float, recvol: float) -> str:
    """FORMAT 1190: I10,2(1PE13.5),7(1PE13.5)"""
    return fi(rep.nstep, 10) + "".join(fe(v, 13, 5) for v in (rep.deltat, rep.time, rep.apot, rep.aact, rep.ovflow, rep.reflow, rep.sfflw,
                                                               recflow, recvol / rep.areatot)) + "\n"


def hgnansf_line(rep) -> str:
    """FORMAT 1195: I6,2(1PE11.3),2(11X,1PE13.5)"""
    return fi(rep.nstep, 6) + fe(rep.deltat, 11, 3) + fe(rep.time, 11, 3) + " " * 11 + fe(rep.ndin + rep.ndout, 13, 5) + " " * 11 + \
        fe(rep.nnin + rep.nnout, 13, 5) + "\n"


def det_line(rep, vol: float) -> str:
    """FORMAT 1197: I7,4(4X,1PE17.9) -- hgsfdet, hgnansfdirdet, hgnansfneudet"""
    return fi(rep.nstep, 7) + "".join("    " + fe(v, 17, 9) for v in (rep.deltat, rep.time, vol, vol / rep.deltat)) + "\n"


def wtdepth_line(time: float, wt) -> str:
    """SRC/wtdepth.f FORMAT 1020: 70(1PE15.6)"""
    vals = [time] + [float(v) for v in wt]
    return "\n".join("".join(fe(v, 15, 6) for v in vals[i:i + 70]) for i in range(0, len(vals), 70)) + "\n"


def hgflag_text(hg) -> str:
    """FORMAT 1260"""
    return "\n HGFLAG:    (1)    (2)    (3)    (4)    (5)    (6)    (7)    (8)    (9)\n" + " " * 8 + " ".join(fi(v, 6) for v in hg) + "\n"


def write_surface_table(fh, nstep: int, time: float, title: str, x, y, values, integer: bool = False) -> None:
    """psisurf / satsurf / swsurf / recharge / fort.777 block of DETOUT (SRC/detout.f FORMATs 1000, 2000-2042, 2060/2080)."""
    fh.write("%s%s     NSTEP   TIME\n" % (fi(nstep, 7), fe(time, 16, 8)))
    fh.write(" SURFACE NODE              X              Y%s\n" % title.rjust(15))
    n = len(values)
    idx = np.char.mod("%6d", np.arange(1, n + 1))
    xs, ys = np.char.mod("%15.6E", x[:n]), np.char.mod("%15.6E", y[:n])
    vs = np.char.mod("%15d", np.asarray(values, dtype=np.int64)) if integer else np.char.mod("%15.6E", values)
    rows = np.char.add(np.char.add(np.char.add(np.char.add("       ", idx), xs), ys), vs)
    fh.write("\n".join(rows))
    fh.write("\n")


def write_velnod(fh, nstep: int, time: float, u, v, w) -> None:
    """SRC/detout.f:34-35, FORMAT 1000 + 1040 (3 values per line = one node per line)."""
    fh.write("%s%s     NSTEP   TIME\n" % (fi(nstep, 7), fe(time, 16, 8)))
    rows = np.char.add(np.char.add(np.char.mod("%15.6E", u), np.char.mod("%15.6E", v)), np.char.mod("%15.6E", w))
    fh.write("\n".join(rows))
    fh.write("\n")


def write_velelt(fh, time: float, uu, vv, ww) -> None:
    """SRC/detout.f:40-43, FORMAT 1010 + three 1020 lists."""
    fh.write("%s       TIME\n" % fe(time, 16, 8))
    for arr in (uu, vv, ww):
        txt = np.char.mod("%15.6E", arr)
        n = len(txt)
        full = (n // 5) * 5
        if full:
            fh.write("\n".join("".join(r) for r in txt[:full].reshape(-1, 5)))
            fh.write("\n")
        if n > full:
            fh.write("".join(txt[full:]) + "\n")
