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
