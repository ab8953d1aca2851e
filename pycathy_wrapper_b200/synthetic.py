"""Deterministic synthetic CATHY projects (SURVEY.md 8d "Synthetic inputs").

Writes a complete project directory in the reference's own file formats
(cathy.fnames + input/* + prepro/*) so the same inputs can be fed to the reference
processor, the CPU oracle and the GPU path.  No RNG: the DEM is a tilted plane like
the bundled hillslope (prepro/dtm_13.val: z = 1 - 0.01*col - 0.025*row) plus small
sine roughness so that drainage directions have no ties.
"""
from __future__ import annotations

import os

import numpy as np

FNAMES = [
    "input/parm", "input/grid", "input/root_map", "input/soil", "input/ic", "input/atmbc", "input/sfbc",
    "input/nansfdirbc", "input/nansfneubc", "prepro/dem", "input/dem_parameters", "input/retctab",
    "input/posizione_serb", "input/livelli_iniz_s", "prepro/lakes_map", "prepro/zone", "input/effraininp",
    "prepro/qoi_a", "prepro/dtm_w_1", "prepro/dtm_w_2", "prepro/dtm_p_outflow_1", "prepro/dtm_p_outflow_2",
    "prepro/dtm_local_slope_1", "prepro/dtm_local_slope_2", "prepro/dtm_epl_1", "prepro/dtm_epl_2",
    "prepro/dtm_kSs1_sf_1", "prepro/dtm_kSs1_sf_2", "prepro/dtm_Ws1_sf_1", "prepro/dtm_Ws1_sf_2",
    "prepro/dtm_b1_sf", "prepro/dtm_y1_sf", "prepro/dtm_nrc", "input/nudging", "input/mesh", "input/base_map",
    "input/transp", "input/transp_ic", "input/transp_atmbc", "input/transp_dirbc",
    "output/debug", "output/risul", "output/xyz", "output/iter", "output/mbeconv", "output/vp", "output/hgatmsf",
    "output/hgnansf", "output/hgflag", "output/sfflag", "output/psi", "output/velnod", "output/sw", "output/ckrw",
    "output/velelt", "output/psisurf", "output/satsurf", "output/swsurf", "output/nansfdir", "output/nansfneu",
    "output/hgsfdet", "output/hgnansfdirdet", "output/hgnansfneudet", "output/cumflowvol", "output/net.ris",
    "output/hgraph", "output/pondhead", "output/dtcoupling", "output/recharge", "output/hgnudging",
    "output/tsnudging", "output/wtdepth", "output/transp", "output/masszone", "output/peatdef",
]

DEFAULT_PARM = dict(
    IPRT1=2, NCOUT=0, TRAFLAG=0, ISIMGR=1, PONDH_MIN=0.0, VELREC=0, KSLOPE=0, TOLKSL=0.01,
    PKRL=-3.0, PKRR=-1.0, PSEL=-3.0, PSER=-1.0, PDSE1L=-3.0, PDSE1R=-2.5, PDSE2L=-1.5, PDSE2R=-1.0,
    ISFONE=0, ISFCVG=0, DUPUIT=0, TETAF=1.0, LUMP=1, IOPT=1, NLRELX=0, OMEGA=0.8,
    L2NORM=0, TOLUNS=1e-4, TOLSWI=1e30, ERNLMX=1e30, ITUNS=10, ITUNS1=5, ITUNS2=7,
    ISOLV=2, ITMXCG=500, TOLCG=1e-10, DELTAT=1.0, DTMIN=1e-2, DTMAX=100.0, TMAX=7200.0,
    DTMAGA=0.0, DTMAGM=1.1, DTREDS=0.0, DTREDM=0.5, IPRT=4, VTKF=0, NPRT=1, TIMPRT=[7200.0],
    NUMVP=1, NODVP=[1], NR=0, NUM_QOUT=0,
)

DEFAULT_SOIL_ROW = (1.88e-4, 1.88e-4, 1.88e-4, 1.0e-5, 0.55, 1.46, 0.15, 0.03125)


def synthetic_dem(nrow: int, ncol: int) -> np.ndarray:
    r = np.arange(nrow)[:, None] / max(nrow, 1)
    c = np.arange(ncol)[None, :] / max(ncol, 1)
    z = 1.5 - 0.5 * r - 0.2 * c                      # ~0.7 m of relief across the rectangle, all > 0
    z = z + 1.0e-3 * np.sin(7.0 * np.arange(nrow)[:, None] + 0.3) * np.sin(5.0 * np.arange(ncol)[None, :] + 0.7)
    return np.round(z, 9)


def geometric_zratio(nstr: int, ratio: float = 1.25) -> np.ndarray:
    w = ratio ** np.arange(nstr)
    z = np.round(w / w.sum(), 6)
    z[-1] = 0.0
    # exact closure so that the reference's sequential sum hits 1 within 1e-14
    s = 0.0
    for v in z[:-1]:
        s += float(v)
    z[-1] = round(1.0 - s, 12)
    return z


def _raster(path: str, arr: np.ndarray, fmt: str, south=0.0, west=0.0) -> None:
    nrow, ncol = arr.shape
    with open(path, "w") as fh:
        fh.write("north:     0\nsouth:     %.8f\neast:      0\nwest:      %.8f\nrows:     %d\ncols:     %d\n"
                 % (south, west, nrow, ncol))
        np.savetxt(fh, arr, fmt=fmt, delimiter=" ")


def write_parm(path: str, p: dict) -> None:
    L = []
    L.append("%d %d %d\tIPRT1 NCOUT TRAFLAG" % (p["IPRT1"], p["NCOUT"], p["TRAFLAG"]))
    L.append("%d %r %d\tISIMGR PONDH_MIN VELREC" % (p["ISIMGR"], float(p["PONDH_MIN"]), p["VELREC"]))
    L.append("%d %r\tKSLOPE TOLKSL" % (p["KSLOPE"], float(p["TOLKSL"])))
    L.append("%r %r %r %r\tPKRL PKRR PSEL PSER" % (p["PKRL"], p["PKRR"], p["PSEL"], p["PSER"]))
    L.append("%r %r %r %r\tPDSE1L PDSE1R PDSE2L PDSE2R" % (p["PDSE1L"], p["PDSE1R"], p["PDSE2L"], p["PDSE2R"]))
    L.append("%d %d %d\tISFONE ISFCVG DUPUIT" % (p["ISFONE"], p["ISFCVG"], p["DUPUIT"]))
    L.append("%r %d %d\tTETAF LUMP IOPT" % (float(p["TETAF"]), p["LUMP"], p["IOPT"]))
    L.append("%d %r\tNLRELX OMEGA" % (p["NLRELX"], float(p["OMEGA"])))
    L.append("%d %r %r %r\tL2NORM TOLUNS TOLSWI ERNLMX" % (p["L2NORM"], float(p["TOLUNS"]), float(p["TOLSWI"]), float(p["ERNLMX"])))
    L.append("%d %d %d\tITUNS ITUNS1 ITUNS2" % (p["ITUNS"], p["ITUNS1"], p["ITUNS2"]))
    L.append("%d %d %r\tISOLV ITMXCG TOLCG" % (p["ISOLV"], p["ITMXCG"], float(p["TOLCG"])))
    L.append("%r %r %r %r\tDELTAT DTMIN DTMAX TMAX" % (float(p["DELTAT"]), float(p["DTMIN"]), float(p["DTMAX"]), float(p["TMAX"])))
    L.append("%r %r %r %r\tDTMAGA DTMAGM DTREDS DTREDM" % (float(p["DTMAGA"]), float(p["DTMAGM"]), float(p["DTREDS"]), float(p["DTREDM"])))
    L.append("%d %d %d %s\tIPRT VTKF NPRT (TIMPRT(I),I=1,NPRT)" % (p["IPRT"], p["VTKF"], len(p["TIMPRT"]), " ".join(repr(float(t)) for t in p["TIMPRT"])))
    L.append("%d %s\tNUMVP (NODVP(I),I=1,NUMVP)" % (len(p["NODVP"]), " ".join(str(int(v)) for v in p["NODVP"])))
    L.append("0\tNR")
    L.append("0\tNUM_QOUT")
    with open(path, "w") as fh:
        fh.write("\n".join(L) + "\n")


def write_hapin(path: str, nrow: int, ncol: int, dx: float, dy: float) -> None:
    """prepro/hap.in for the reference pre-processor (`cppp`); values as in the bundled hillslope."""
    rows = [
        ("Grid spacing along the x-direction", "%.2f" % dx), ("Grid spacing along the y-direction", "%.2f" % dy),
        ("DEM rectangle size along the x-direction", "%d" % ncol), ("DEM rectangle size along the y-direction", "%d" % nrow),
        ("Number of cells within the catchment", "%d" % (nrow * ncol)),
        ("X low left corner coordinate", "0.00000000"), ("Y low left corner coordinate", "0.00000000"),
    ]
    terr = [
        ("Depit threshold slope", "0.130E-06"), ("Drainage directions method (LAD:1,LTD:2)", "1"),
        ("Upstream deviation memory factor (CBM:0,PBM:1)", "0.000E+00"),
        ("Threshold on the contour curvature (NDM:-1E10;DM:+1E10)", "0.100E+13"),
        ("Nondispersive channel flow (0:not-required;1:required)", "0"),
        ("Channel initiation method (A:1,AS**k:2,ND:3)", "1"), ("Threshold on the support area (A)", "0.200000000E+04"),
        ("Threshold on the AS**k function", "16000.00"), ("Exponent k of the AS**k function", "2.00"),
        ("Threshold on the normalized divergence (ND)", "-0.100E-01"), ("Path threshold slope", "0.500E-03"),
        ("Drainage direction of the outlet cell (if necessary...) ", "4"),
        ("Boundary channel constraction (No:0,Yes:1)", "0"),
        ("Coefficient for boundary channel elevation definition", "0.50"),
        ("Coefficient for outlet cell elevation definition", "0.50"),
    ]
    bar = "-" * 78
    with open(path, "w") as fh:
        fh.write(bar + "\nSTRUCTURAL PARAMETERS\n" + bar + "\n")
        for k, v in rows:
            fh.write(("%s =" % k).ljust(56) + v.rjust(14) + "\n")
        fh.write(bar + "\nTERRAIN ANALYSIS PARAMETERS\n" + bar + "\n")
        for k, v in terr:
            fh.write(("%s =" % k).ljust(58) + v.rjust(16) + "\n")
        fh.write(bar + "\nRIVULET NETWORK PARAMETERS (HYDRAULIC GEOMETRY OF THE SINGLE RIVULET)\n" + bar + "\n")
        fh.write("Rivulet spacing =                                    %.3f\n" % dx)
        fh.write("Reference drainage area (As_rf) =                    0.200000000000E+04\n")
        fh.write("Flow discharge (Qsf_rf,w_rf) =                       1.000               1.000\n")
        fh.write("Water-surface width (Wsf_rf,b1_rf,b2_rf) =           1.000     0.000     0.000\n")
        fh.write("Resistance coefficient (kSsf_rf,y1_rf,y2_rf) =      24.004     0.000     0.000\n")
        fh.write("Initial flow discharge (Qsi_rf) =                    0.000\n")
        fh.write(bar + "\nCHANNEL NETWORK PARAMETERS\n" + bar + "\n")
        fh.write("Reference drainage area (As_cf) =                    0.200000000000E+04\n")
        fh.write("Flow discharge (Qsf_cf,w_cf) =                       1.000               1.000\n")
        fh.write("Water-surface width (Wsf_cf,b1_cf,b2_cf) =           5.000     0.260     0.500\n")
        fh.write("Resistance coefficient (kSsf_cf,y1_cf,y2_cf) =      66.500     0.000     0.000\n")
        fh.write("Initial flow discharge (Qsi_cf) =                    1.000\n" + bar + "\n")


def make_project(path: str, nrow: int, ncol: int, nstr: int, dx: float = 0.5, dy: float = 0.5, base: float = 3.0,
                 zratio=None, dem=None, soil_rows=None, ic=("uniform", -1.0), atmbc=None, hspatm: int = 1, ieto: int = 0,
                 pmin: float = -5.0, dirbc_text: str | None = None, neubc_text: str | None = None, ivghu: int = 0,
                 hu=(0.02, 2, 2, 0, 0.333), hun=1, huab=(-5, 1), bc=(1.2, 0, -0.345), zone=None, ivert: int = 0, pond: float | None = None, seepage_faces=None, **parm) -> str:
    """Write a full project directory.  `ic` = ("uniform", psi) | ("hydrostatic",) | ("wt", position);
    `atmbc` = list of (time, rate) pairs (homogeneous) -- rate in m/s, +ve = rain."""
    for sub in ("input", "prepro", "output", "vtk"):
        os.makedirs(os.path.join(path, sub), exist_ok=True)
    p = dict(DEFAULT_PARM)
    p.update({k.upper(): v for k, v in parm.items()})
    p["NPRT"] = len(p["TIMPRT"])
    with open(os.path.join(path, "cathy.fnames"), "w") as fh:
        fh.write("'.'\n")
        for nm in FNAMES:
            fh.write(("'%s'" % nm).ljust(36) + "\n")
    write_parm(os.path.join(path, "input", "parm"), p)
    dem = synthetic_dem(nrow, ncol) if dem is None else np.asarray(dem, dtype=np.float64)
    zr = geometric_zratio(nstr) if zratio is None else np.asarray(zratio, dtype=np.float64)
    _raster(os.path.join(path, "prepro", "dem"), dem, "%.12E")
    np.savetxt(os.path.join(path, "prepro", "dtm_13.val"), dem, fmt="%.9f", delimiter="\t")
    zone = np.ones((nrow, ncol), dtype=int) if zone is None else np.asarray(zone, dtype=int)
    nzone = int(zone.max())
    _raster(os.path.join(path, "prepro", "zone"), zone, "%d")
    _raster(os.path.join(path, "prepro", "lakes_map"), np.zeros((nrow, ncol), dtype=int), "%d")
    _raster(os.path.join(path, "input", "root_map"), np.ones((nrow, ncol), dtype=int), "%d")
    write_hapin(os.path.join(path, "prepro", "hap.in"), nrow, ncol, dx, dy)
    with open(os.path.join(path, "input", "dem_parameters"), "w") as fh:
        fh.write("%r\n%r\n1.0\n1\n%d\t%d\t25\n%d\t1\t%r\n%s\n" % (dx, dy, nzone, nstr, ivert, base, "\t".join(repr(float(v)) for v in zr)))
        fh.write("delta_x\ndelta_y\nfactor\ndostep\nnzone\tnstr\tn1\nivert\tisp\tbase\nzratio(i),i=1,nstr\n")
    rows = soil_rows if soil_rows is not None else [DEFAULT_SOIL_ROW] * (nstr * nzone)     # layer-major: all zones of layer 1, then layer 2, ...
    with open(os.path.join(path, "input", "soil"), "w") as fh:
        fh.write("%r\tPMIN\n0 1.0\tIPEAT SCF\n0.4 0.225\tCBETA0,CANG\n" % pmin)
        fh.write("0.0 -4.0 -150.0 1.0 1.0 1.0\tPCANA,PCREF,PCWLT,ZROOT,PZ,OMGC\n%d\tIVGHU\n" % ivghu)
        fh.write("%r %r %r %r %r\tHUALFA,HUBETA,HUGAMA,HUPSIA,HUSWR\n%r\tHUN\n%r %r\tHUA,HUB\n%r %r %r\tBCBETA,BCRMC,BCPSAT\n"
                 % (*hu, hun, *huab, *bc))
        for r in rows:
            fh.write(" ".join("%.6E" % v for v in r) + "\n")
    with open(os.path.join(path, "input", "ic"), "w") as fh:
        ipond = 1 if pond is not None else 0          # IPOND = 1: uniform initial ponding head (SRC/datin.f:380-403)
        if ic[0] == "uniform":
            fh.write("0 %d\tINDP IPOND\n%r\n" % (ipond, float(ic[1])))
        elif ic[0] == "hydrostatic":
            fh.write("2 %d\tINDP IPOND\n0\tWTPOSITION\n" % ipond)
        elif ic[0] == "wt":
            fh.write("3 %d\tINDP IPOND\n%r\tWTPOSITION\n" % (ipond, float(ic[1])))
        else:
            raise ValueError(ic)
        if ipond:
            fh.write("%r\tPONDING HEAD\n" % float(pond))
    atmbc = atmbc if atmbc is not None else [(0.0, 0.0), (1.0e9, 0.0)]
    with open(os.path.join(path, "input", "atmbc"), "w") as fh:
        fh.write("%d %d\tHSPATM IETO\n" % (hspatm, ieto))
        for t, v in atmbc:
            fh.write("%r\tTIME\n" % float(t))
            if np.ndim(v) == 0:
                fh.write("%r\tATMINP\n" % float(v))
            else:
                fh.write(" ".join("%.9E" % x for x in np.ravel(v)) + "\n")
    for nm, txt in (("nansfdirbc", dirbc_text), ("nansfneubc", neubc_text)):
        with open(os.path.join(path, "input", nm), "w") as fh:
            fh.write(txt if txt is not None else "0.0\tTIME\n0 0\n1.0e9\tTIME\n0 0\n")
    with open(os.path.join(path, "input", "sfbc"), "w") as fh:
        if seepage_faces:           # list of faces, each a list of 1-based 3-D node ids with descending elevation (SRC/sfvone.f)
            fh.write("0.0\n%d\n" % len(seepage_faces))
            for f in seepage_faces:
                fh.write("%d\n%s\n" % (len(f), " ".join(str(int(v)) for v in f)))
            fh.write("1.0e30\n0\n")
        else:
            fh.write("0\n0\n1.0e9\n0\n")
    with open(os.path.join(path, "input", "posizione_serb"), "w") as fh:
        fh.write("0\n")
    for nm in ("grid", "retctab", "livelli_iniz_s", "effraininp", "mesh", "base_map", "transp", "transp_ic",
               "transp_atmbc", "transp_dirbc"):
        open(os.path.join(path, "input", nm), "a").close()
    with open(os.path.join(path, "input", "nudging"), "w") as fh:
        fh.write("0 0 0\tNUDN,NUDT,NUDFLAG\n")
    # the reference opens every input unit of cathy.fnames (SRC/openio.f), also those it never reads
    for nm in FNAMES:
        if not nm.startswith("output/") and not os.path.exists(os.path.join(path, nm)):
            open(os.path.join(path, nm), "a").close()
    return path
