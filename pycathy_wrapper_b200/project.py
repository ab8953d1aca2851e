"""Reader for a CATHY project directory (the drop-in boundary, input side).

The reference processor reads ``./cathy.fnames`` and then ~25 whitespace separated
text files with Fortran list-directed ``READ(unit,*)`` statements
(reference: SRC/openio.f:35, SRC/datin.f:80-122,195-204,266,325-372,380-403,421-458,510-514,
SRC/atmone.f, SRC/bcone.f, SRC/rdndbc.f, SRC/readbc.f, SRC/rast_input_*.f; SRC =
pyCATHY/tests/weil_exemple/my_cathy_prj/src).  This module re-implements that
grammar in Python and returns plain numpy arrays; nothing here touches the GPU.

Scope of round 1: DEM based projects (ISIMGR = 1 or 2) whose DEM
rectangle is fully inside the catchment (no zero cells), no lakes/reservoirs, no
seepage faces, DOSTEP = 1, IVGHU = 0..4 (no look-up tables).  Anything else raises
``CathyInputError`` -- loudly, never silently ignored.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

import numpy as np


class CathyInputError(ValueError):
    """Raised when a project uses a feature outside the implemented scope or is malformed."""


# unit order of cathy.fnames (reference: SRC/openio.f:35-; one quoted path per record)
FNAMES_UNITS = [
    "IIN1", "IIN2", "IIN3", "IIN4", "IIN5", "IIN6", "IIN7", "IIN8", "IIN9", "IIN10", "IIN11",
    "IIN16", "IIN17", "IIN18", "IIN20", "IIN21", "IIN22", "IIN23", "IIN25", "IIN26", "IIN27",
    "IIN28", "IIN29", "IIN30", "IIN31", "IIN32", "IIN33", "IIN34", "IIN35", "IIN36", "IIN37",
    "IIN38", "IIN39", "IIN50", "IIN51", "IIN60", "IIN61", "IIN62", "IIN63", "IIN64",
    "IOUT1", "IOUT2", "IOUT3", "IOUT4", "IOUT5", "IOUT6", "IOUT7", "IOUT8", "IOUT9", "IOUT10",
    "IOUT11", "IOUT12", "IOUT13", "IOUT14", "IOUT15", "IOUT16", "IOUT17", "IOUT18", "IOUT19",
    "IOUT20", "IOUT30", "IOUT31", "IOUT32", "IOUT36", "IOUT40", "IOUT41", "IOUT42", "IOUT43",
    "IOUT44", "IOUT50", "IOUT51", "IOUT57", "IOUT80", "IOUT81", "IOUTPT",
]

_REPEAT = re.compile(r"^(\d+)\*(.+)$")


class ListDirectedReader:
    """Fortran list-directed input: each ``read`` starts on a fresh record, may continue
    over several records, and discards whatever is left on the last record it touched."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "r", errors="replace") as fh:
            self.lines = fh.read().splitlines()
        self.pos = 0

    def eof(self) -> bool:
        p = self.pos
        while p < len(self.lines) and not self.lines[p].strip():
            p += 1
        return p >= len(self.lines)

    def skip_record(self) -> None:
        self.pos += 1

    def raw_line(self) -> str:
        if self.pos >= len(self.lines):
            raise EOFError(self.path)
        s = self.lines[self.pos]
        self.pos += 1
        return s

    def read(self, n: int, conv=float) -> list:
        out: list = []
        if n == 0:
            # a READ with an empty list still consumes one record
            self.pos += 1
            return out
        while len(out) < n:
            if self.pos >= len(self.lines):
                raise EOFError(f"{self.path}: end of file while reading {n} items")
            toks = self.lines[self.pos].replace(",", " ").split()
            self.pos += 1
            for t in toks:
                m = _REPEAT.match(t)
                reps, body = (int(m.group(1)), m.group(2)) if m else (1, t)
                try:
                    v = conv(_fnum(body)) if conv is not str else body
                except ValueError:
                    # non numeric trailer (a comment) ends this record, like a Fortran
                    # read that already has what it needs; if it does not, it is an error
                    if len(out) < n:
                        raise CathyInputError(
                            f"{self.path}:{self.pos}: expected {n} values, found text {t!r} after {len(out)}")
                    break
                out.extend([v] * reps)
                if len(out) >= n:
                    break
        return out[:n]

    def read_f(self, n: int) -> np.ndarray:
        return np.asarray(self.read(n, float), dtype=np.float64)

    def read_i(self, n: int) -> np.ndarray:
        return np.asarray(self.read(n, _fint), dtype=np.int64)


def _fnum(tok: str) -> str:
    t = tok.strip().strip("'\"")
    return t.replace("d", "e").replace("D", "e")


def _fint(tok: str) -> int:
    return int(float(tok))


def read_raster(path: str, dtype=np.float64) -> tuple[np.ndarray, dict]:
    """6 header lines (north/south/east/west/rows/cols) + nrow records, first record = north row
    (reference: SRC/rast_input_dem.f; header fields are taken from column 8/7 onwards)."""
    with open(path, "r", errors="replace") as fh:
        lines = fh.read().splitlines()
    hdr = {}
    keys = ["north", "south", "east", "west", "rows", "cols"]
    for k, ln in zip(keys, lines[:6]):
        cut = ln[7:] if k not in ("rows", "cols") else ln[6:]
        hdr[k] = float(_fnum(cut.split()[0]))
    nrow, ncol = int(hdr["rows"]), int(hdr["cols"])
    vals: list[float] = []
    for ln in lines[6:]:
        if len(vals) >= nrow * ncol:
            break
        vals.extend(float(_fnum(t)) for t in ln.replace(",", " ").split())
    if len(vals) < nrow * ncol:
        raise CathyInputError(f"{path}: raster holds {len(vals)} values, header says {nrow}x{ncol}")
    arr = np.asarray(vals[: nrow * ncol], dtype=np.float64).reshape(nrow, ncol)
    if dtype is not np.float64:
        arr = arr.astype(dtype)
    return arr, hdr


def nodal_mean_of_cells(cells: np.ndarray) -> np.ndarray:
    """Surface-node value = mean over the adjacent triangles of their cell value
    (reference: SRC/triangoli.f:91-110 two triangles per cell sharing the NW-SE diagonal,
    SRC/tpnodi2d.f).  Returns an (nrow+1, ncol+1) array."""
    nrow, ncol = cells.shape
    acc = np.zeros((nrow + 1, ncol + 1))
    cnt = np.zeros((nrow + 1, ncol + 1))
    for (di, dj, w) in ((0, 0, 2), (1, 0, 1), (1, 1, 2), (0, 1, 1)):
        acc[di:di + nrow, dj:dj + ncol] += w * cells
        cnt[di:di + nrow, dj:dj + ncol] += w
    return acc / cnt


def read_fnames(prj: str) -> dict[str, str]:
    """cathy.fnames: record 1 = base directory, then one quoted relative path per unit."""
    path = os.path.join(prj, "cathy.fnames")
    with open(path, "r") as fh:
        lines = [ln for ln in fh.read().splitlines() if ln.strip()]
    def first_quoted(s: str) -> str:
        m = re.search(r"'([^']*)'", s)
        return m.group(1) if m else s.split()[0]
    base = first_quoted(lines[0])
    names = [first_quoted(ln) for ln in lines[1:]]
    out = {}
    for unit, nm in zip(FNAMES_UNITS, names):
        out[unit] = os.path.normpath(os.path.join(prj, base, nm))
    return out


@dataclass
class BCTable:
    """Time-record table of a nansfdirbc / nansfneubc file (SRC/bcone.f, rdndbc.f, readbc.f).
    ``nodes`` are 1-based 3-D node ids after the 2-D -> all-layers expansion."""
    times: list = field(default_factory=list)
    nodes: list = field(default_factory=list)    # list of int64 arrays
    values: list = field(default_factory=list)   # list of float64 arrays
    n2d: list = field(default_factory=list)      # NODIN2 of each record (<0: free drainage)


def read_bc_table(path: str, nnod: int, nstr: int) -> BCTable:
    tab = BCTable()
    rd = ListDirectedReader(path)
    while not rd.eof():
        try:
            t = rd.read(1)[0]
        except (EOFError, CathyInputError):
            break
        try:
            n2, n3 = rd.read(2, _fint)
        except EOFError:
            n2, n3 = 0, 0
        if n2 < 0:
            nbc = nnod + n3
            nodes = list(range(nnod * nstr + 1, nnod * nstr + nnod + 1))
            if n3:
                nodes += rd.read(n3, _fint)
            vals = [0.0] * nnod
            if n3:
                vals += rd.read(n3)
        elif n2 > 0:
            nbc = n2 * (nstr + 1) + n3
            surf = rd.read(n2, _fint)
            nodes = [s + k * nnod for k in range(nstr + 1) for s in surf]
            if n3:
                nodes += rd.read(n3, _fint)
            v2 = rd.read(n2)
            vals = [v for _k in range(nstr + 1) for v in v2]
            if n3:
                vals += rd.read(n3)
        else:
            nbc = n3
            nodes = rd.read(n3, _fint) if n3 else []
            vals = rd.read(n3) if n3 else []
        assert len(nodes) == nbc and len(vals) == nbc
        tab.times.append(float(t))
        tab.nodes.append(np.asarray(nodes, dtype=np.int64))
        tab.values.append(np.asarray(vals, dtype=np.float64))
        tab.n2d.append(int(n2))
    return tab


@dataclass
class CathyProject:
    """Everything DATIN + the *ONE routines read, as numpy arrays / python scalars."""
    path: str
    fnames: dict
    parm: dict
    transport_skipped: bool = False      # parm says TRAFLAG=1 and the caller opted in to run the flow problem only
    # mesh description
    nrow: int = 0
    ncol: int = 0
    nstr: int = 0
    nzone: int = 1
    n1: int = 25
    dx: float = 0.0
    dy: float = 0.0
    west: float = 0.0
    south: float = 0.0
    factor: float = 1.0
    dostep: int = 1
    ivert: int = 0
    isp: int = 0
    base: float = 0.0
    zratio: np.ndarray = None
    dem: np.ndarray = None
    zone: np.ndarray = None
    root_map: np.ndarray = None
    lakes_map: np.ndarray = None
    # soil
    soil: dict = None
    # ic
    indp: int = 0
    ipond: int = 0
    wtposition: float = 0.0
    ic_psi: np.ndarray = None
    ic_pond: np.ndarray = None
    # atmbc
    hspatm: int = 0
    ieto: int = 0
    atm_times: np.ndarray = None
    atm_values: np.ndarray = None     # (ntimes, 1) homogeneous or (ntimes, nnod)
    atm_none: bool = False            # HSPATM == 9999 or empty file
    # other BCs
    dirbc: BCTable = None
    neubc: BCTable = None
    # surface routing rasters (ISIMGR == 2)
    surf: dict = None

    @property
    def nnod(self) -> int:
        return (self.nrow + 1) * (self.ncol + 1)

    @property
    def n(self) -> int:
        return self.nnod * (self.nstr + 1)

    @property
    def ntri(self) -> int:
        return 2 * self.nrow * self.ncol

    @property
    def nt(self) -> int:
        return 3 * self.ntri * self.nstr


PARM_LAYOUT = [
    (("IPRT1", int), ("NCOUT", int), ("TRAFLAG", int)),
    (("ISIMGR", int), ("PONDH_MIN", float), ("VELREC", int)),
    (("KSLOPE", int), ("TOLKSL", float)),
    (("PKRL", float), ("PKRR", float), ("PSEL", float), ("PSER", float)),
    (("PDSE1L", float), ("PDSE1R", float), ("PDSE2L", float), ("PDSE2R", float)),
    (("ISFONE", int), ("ISFCVG", int), ("DUPUIT", int)),
    (("TETAF", float), ("LUMP", int), ("IOPT", int)),
    (("NLRELX", int), ("OMEGA", float)),
    (("L2NORM", int), ("TOLUNS", float), ("TOLSWI", float), ("ERNLMX", float)),
    (("ITUNS", int), ("ITUNS1", int), ("ITUNS2", int)),
    (("ISOLV", int), ("ITMXCG", int), ("TOLCG", float)),
    (("DELTAT", float), ("DTMIN", float), ("DTMAX", float), ("TMAX", float)),
    (("DTMAGA", float), ("DTMAGM", float), ("DTREDS", float), ("DTREDM", float)),
]


def read_parm(path: str) -> dict:
    """input/parm (reference: SRC/datin.f:80-122)."""
    rd = ListDirectedReader(path)
    p: dict = {}
    for rec in PARM_LAYOUT:
        vals = rd.read(len(rec))
        for (name, typ), v in zip(rec, vals):
            p[name] = int(v) if typ is int else float(v)
    # IPRT,VTKF,NPRT,(TIMPRT(I),I=1,NPRT): NPRT is only known after the third item, and the
    # list may continue on following records -- peek by re-reading the record.
    save = rd.pos
    head = rd.read(3)
    nprt = int(head[2])
    rd.pos = save
    vals = rd.read(3 + nprt)
    p["IPRT"], p["VTKF"], p["NPRT"] = int(vals[0]), int(vals[1]), nprt
    p["TIMPRT"] = [float(v) for v in vals[3:]]
    save = rd.pos
    numvp = int(rd.read(1)[0])
    rd.pos = save
    vals = rd.read(1 + numvp)
    p["NUMVP"] = numvp
    p["NODVP"] = [int(v) for v in vals[1:]]
    p["NR"] = int(rd.read(1)[0])
    p["CONTR"] = [int(v) for v in rd.read(p["NR"])] if p["NR"] else []
    try:
        save = rd.pos
        nq = int(rd.read(1)[0])
        rd.pos = save
        vals = rd.read(1 + nq)
        p["NUM_QOUT"] = nq
        p["ID_QOUT"] = [int(v) for v in vals[1:]]
    except (EOFError, CathyInputError):
        p["NUM_QOUT"], p["ID_QOUT"] = 0, []
    if p["ISIMGR"] <= 1:
        p["PONDH_MIN"] = 1.0e10            # SRC/datin.f:122
    return p


def load_project(prj: str, skip_transport: bool | None = None) -> CathyProject:
    """Read a CATHY project directory like DATIN does.  ``skip_transport`` (default: the environment variable
    CATHY_B200_SKIP_TRANSPORT=1): accept a project whose parm says TRAFLAG=1 and run its FLOW problem only.  In the reference the
    solute transport is a one-way add-on after each flow step (SRC/cathy_main.f:3304-3607: it reads the flow's velocities, the flow
    never reads a concentration), so psi, sw, vp, mbeconv, cumflowvol ... are unaffected; the concentration outputs are not written."""
    prj = os.path.abspath(prj)
    fn = read_fnames(prj)
    parm = read_parm(fn["IIN1"])
    P = CathyProject(path=prj, fnames=fn, parm=parm)
    isim = parm["ISIMGR"]
    if isim not in (1, 2):
        raise CathyInputError(f"ISIMGR={isim}: only DEM based runs (1: subsurface, 2: coupled) are implemented")
    if skip_transport is None:
        skip_transport = os.environ.get("CATHY_B200_SKIP_TRANSPORT", "0") == "1"
    P.transport_skipped = False
    if parm["TRAFLAG"] != 0:
        if not skip_transport:
            raise CathyInputError("TRAFLAG=1 (solute transport) is outside the hot-path scope; set CATHY_B200_SKIP_TRANSPORT=1 "
                                  "(or load_project(..., skip_transport=True)) to run the flow problem of this project without its transport add-on")
        P.transport_skipped = True

    P.dem, hdr = read_raster(fn["IIN10"])
    P.nrow, P.ncol = P.dem.shape
    P.west, P.south = hdr["west"], hdr["south"]
    P.zone, _ = read_raster(fn["IIN21"], np.int32)
    rd = ListDirectedReader(fn["IIN11"])
    P.dx, P.dy = rd.read(2)
    P.factor = rd.read(1)[0]
    P.dostep = int(rd.read(1)[0])
    P.nzone, P.nstr, P.n1 = (int(v) for v in rd.read(3))
    iv, isp, base = rd.read(3)
    P.ivert, P.isp, P.base = int(iv), int(isp), float(base)
    P.lakes_map, _ = read_raster(fn["IIN20"], np.int32)
    P.root_map, _ = read_raster(fn["IIN3"])
    P.zratio = rd.read_f(P.nstr)
    if abs(P.zratio.sum() - 1.0) > 1.0e-14 and abs(float(np.add.reduce(P.zratio)) - 1.0) > 1e-14:
        # the reference sums sequentially (SRC/datin.f:268-276); redo it the same way
        s = 0.0
        for z in P.zratio:
            s += float(z)
        if abs(s - 1.0) > 1.0e-14:
            raise CathyInputError(f"ZRATIO does not sum to 1 (sum={s!r})")
    if P.dostep != 1:
        raise CathyInputError("DOSTEP != 1 (DEM coarsening) is not implemented")
    if P.ivert not in (0, 1, 2, 4):
        raise CathyInputError("IVERT=3 (base map) is not implemented")
    if np.any(P.dem == 0) or np.any(P.zone == 0):
        raise CathyInputError("DEM/zone rasters with null cells (irregular catchment outline) are not implemented")
    if np.any(P.lakes_map > 0):
        raise CathyInputError("lakes / reservoirs are outside the hot-path scope (NUMRES forced to 0 upstream)")
    if P.zone.max() > P.nzone or P.zone.min() < 1:
        raise CathyInputError("zone raster holds ids outside 1..NZONE")

    nnod, n, nstr = P.nnod, P.n, P.nstr

    # ---- ic (SRC/datin.f:380-403)
    rd = ListDirectedReader(fn["IIN5"])
    P.indp, P.ipond = (int(v) for v in rd.read(2))
    if isim != 2:
        P.ipond = 0
    if P.indp in (3, 4):
        P.wtposition = rd.read(1)[0]
    if P.indp == 0:
        P.ic_psi = np.full(n, rd.read(1)[0])
    elif P.indp == 1:
        P.ic_psi = rd.read_f(n)
    elif P.indp in (2, 3, 4):
        P.ic_psi = np.zeros(n)
    else:
        raise CathyInputError(f"INDP={P.indp} unknown")
    P.ic_pond = np.zeros(nnod)
    if P.ipond == 1:
        P.ic_pond[:] = rd.read(1)[0]
    elif P.ipond == 2:
        P.ic_pond = rd.read_f(nnod)

    # ---- soil (SRC/datin.f:421-458,510-514)
    rd = ListDirectedReader(fn["IIN4"])
    soil: dict = {}
    soil["PMIN"] = rd.read(1)[0]
    ipeat, scf = rd.read(2)
    soil["IPEAT"], soil["SCF"] = int(ipeat), float(scf)
    soil["CBETA0"], soil["CANG"] = rd.read(2)
    nveg = max(int(nodal_mean_of_cells(P.root_map * P.factor).astype(np.int64).max()), 1)
    veg = np.zeros((nveg, 6))
    for i in range(nveg):
        veg[i] = rd.read(6)
    soil["VEG"] = veg                            # PCANA PCREF PCWLT ZROOT PZ OMGC
    soil["IVGHU"] = int(rd.read(1)[0])
    soil["HU"] = rd.read(5)
    soil["HUN"] = rd.read(1)[0]
    soil["HUAB"] = rd.read(2)
    soil["BC"] = rd.read(3)
    tab = np.zeros((nstr, P.nzone, 8))
    for i in range(nstr):
        for j in range(P.nzone):
            tab[i, j] = rd.read(8)
    soil["TABLE"] = tab                          # PERMX PERMY PERMZ ELSTOR POROS VGN VGRMC VGPSAT
    P.soil = soil
    if soil["IPEAT"] != 0:
        raise CathyInputError("IPEAT=1 (peat deformation) is outside the hot-path scope")
    if soil["IVGHU"] not in (0, 1, 2, 3, 4):
        raise CathyInputError(f"IVGHU={soil['IVGHU']}: van Genuchten (0), extended van Genuchten (1), Huyakorn (2, 3) and Brooks-Corey (4) are implemented")

    # ---- atmbc (SRC/atmone.f, SRC/atmnxt.f)
    rd = ListDirectedReader(fn["IIN6"])
    times, vals = [], []
    if rd.eof():
        P.atm_none = True
    else:
        P.hspatm, P.ieto = (int(v) for v in rd.read(2))
        if P.hspatm == 9999:
            P.atm_none = True
        else:
            while not rd.eof():
                try:
                    t = rd.read(1)[0]
                except (EOFError, CathyInputError):
                    break
                if P.hspatm == 0:
                    v = rd.read_f(nnod)
                else:
                    v = np.asarray(rd.read(1))
                times.append(t)
                vals.append(v)
    P.atm_times = np.asarray(times, dtype=np.float64)
    P.atm_values = np.asarray(vals, dtype=np.float64).reshape(len(times), -1) if times else np.zeros((0, 1))

    # ---- non-atmospheric, non seepage-face BCs
    P.dirbc = read_bc_table(fn["IIN8"], nnod, nstr)
    P.neubc = read_bc_table(fn["IIN9"], nnod, nstr)

    # ---- seepage faces (SRC/sfvone.f:27-66): TIME, NSF, then per face its node count and node ids; the record after it gives the
    # time from which a new set would apply (SRC/sfvnxt.f) -- only a set that holds for the whole run is supported
    P.seepage_faces = []
    rd = ListDirectedReader(fn["IIN7"])
    if not rd.eof():
        t0 = float(rd.read(1)[0])                 # SFVTIM(1)
        nsf = int(rd.read(1)[0])
        if nsf > 0 and t0 <= 0.0:                 # sfvone returns before reading anything when SFVTIM(1) > TIME = 0
            for _ in range(nsf):
                cnt = int(rd.read(1)[0])
                P.seepage_faces.append(np.asarray(rd.read_i(cnt), dtype=np.int32))
            t1 = float(rd.read(1)[0]) if not rd.eof() else 1.0e30
            if t1 <= float(P.parm["TMAX"]):
                raise CathyInputError("input/sfbc: a second seepage-face record at TIME=%g <= TMAX; time-varying seepage-face sets "
                                      "(SRC/sfvnxt.f) are not implemented" % t1)
        elif nsf > 0:
            raise CathyInputError("input/sfbc: first record at TIME=%g > 0 (the reference would start without seepage faces and "
                                  "switch them on later, SRC/sfvnxt.f); not implemented" % t0)

    # ---- surface routing inputs (SRC/datin.f:325-372)
    if isim == 2:
        S: dict = {}
        rdq = ListDirectedReader(fn["IIN23"])
        ncell = P.nrow * P.ncol
        n_listed = int(rdq.read_i(1)[0])            # first record of qoi_a: the number of catchment cells (PRE/hg.f90:31)
        if n_listed != ncell:
            raise CathyInputError("prepro/qoi_a lists %d catchment cells for a %d x %d DEM: the pre-processor saw null cells (irregular "
                                  "catchment outline), which the processor does not implement" % (n_listed, P.nrow, P.ncol))
        S["qoi"] = rdq.read_i(ncell)
        names = ["w_1", "w_2", "p_outflow_1", "p_outflow_2", "local_slope_1", "local_slope_2", "epl_1",
                 "epl_2", "kSs1_sf_1", "kSs1_sf_2", "Ws1_sf_1", "Ws1_sf_2", "b1_sf", "y1_sf", "nrc"]
        units = ["IIN25", "IIN26", "IIN27", "IIN28", "IIN29", "IIN30", "IIN31", "IIN32", "IIN33",
                 "IIN34", "IIN35", "IIN36", "IIN37", "IIN38", "IIN39"]
        for nm, u in zip(names, units):
            arr, _ = read_raster(fn[u])
            if arr.shape != (P.nrow, P.ncol):
                raise CathyInputError(f"{fn[u]}: raster shape {arr.shape} != DEM shape")
            S[nm] = arr
        rdr = ListDirectedReader(fn["IIN17"])
        if int(rdr.read(1)[0]) != 0:
            raise CathyInputError("reservoirs (posizione_serb) are outside the hot-path scope")
        P.surf = S
    return P
