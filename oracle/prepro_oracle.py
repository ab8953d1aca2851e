"""TEST INFRASTRUCTURE -- CPU restatement of the CATHY pre-processor `cppp`.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
(pycathy_wrapper_b200/preprocessor.py + csrc/cathy_prepro.cu) never does.

PRE = /root/reference/examples/SSHydro/weill_exemple/prepro/src (HAP v11.7, the sources pyCATHY compiles
into `pycppp`, PY/cathy_tools.py:353).  The program (PRE/cppp.f90:20-86) runs
    wbb_sr   dtm_13.val -> cell records (PRE/wbb_sr.f90:24-170; optional boundary channel :95-160)
    wparfile rewrites hap.in (PRE/mpar.f90:400-541) -- later stages RE-READ the rounded values (cca.f90:31, mrbb_sr.f90:24)
    csort    cells in descending elevation, unstable quicksort (PRE/csort.f90:14-67, qsort.f90:9-125)
    depit    raise pits by pt*dx above their lowest neighbour, sweep by sweep (PRE/depit.f90:14-154)
    csort    again
    cca      contour curvature -> drainage method per cell (PRE/cca.f90:10-96)
    smean    mean of the steepest facet slopes (PRE/smean.f90:10-171, facet.f90:10-47)
    dsf      two drainage directions, weights, slopes, path lengths, upstream area (PRE/dsf.f90:20-667)
    hg       hydraulic geometry of rivulets / channels (PRE/hg.f90:10-113)
    mrbb_sr  ASCII rasters dem, lakes_map, zone, dtm_* (PRE/mrbb_sr.f90:10-540) and qoi_a (hg.f90:31-37)
Pinned: byte-identical files against the reference's own ELF `pycppp` (oracle/_ref/bin/pycppp) -- see
tests/test_prepro_oracle.py, tests/golden/prepro/.  Pure-Python loops: small DEMs (<= ~40k cells) only.
Not restated: bb2shp_sr (ESRI shape files of the river network, read by no part of pyCATHY or CATHY),
the Strahler / Horton orders that only those shape files carry, dtm_Kc.txt.
"""
from __future__ import annotations

import math
import os

import numpy as np

F32 = np.float32
EPS64 = 2.220446049250313e-16
EPS32 = float(np.finfo(np.float32).eps)

# --------------------------------------------------------------------------- hap.in
# (key, kind) in file order; kind: 'd' REP, 's' RSP, 'i' integer; tuples for several values per record (PRE/mpar.f90:28-68)
HAP_RECORDS = [
    (("delta_x",), "d"), (("delta_y",), "d"), (("N",), "i"), (("M",), "i"), (("N_celle",), "i"), (("xllcorner",), "d"),
    (("yllcorner",), "d"),
    (("pt",), "d"), (("imethod",), "i"), (("lambda_",), "d"), (("CC_threshold",), "s"), (("ndcf",), "i"), (("nchc",), "i"),
    (("A_threshold",), "d"), (("ASk_threshold",), "s"), (("kas",), "s"), (("DN_threshold",), "s"), (("local_slope_t",), "s"),
    (("p_outflow_vo",), "i"), (("bcc",), "i"), (("cqm",), "s"), (("cqg",), "s"),
    (("dr",), "d"), (("As_rf",), "d"), (("Qsf_rf", "w_rf"), "s"), (("Wsf_rf", "b1_rf", "b2_rf"), "s"),
    (("kSsf_rf", "y1_rf", "y2_rf"), "s"), (("Qsi_rf",), "s"),
    (("As_cf",), "d"), (("Qsf_cf", "w_cf"), "s"), (("Wsf_cf", "b1_cf", "b2_cf"), "s"), (("kSsf_cf", "y1_cf", "y2_cf"), "s"),
    (("Qsi_cf",), "s"),
]


def _fnum(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


def parse_hap(text: str) -> dict:
    """RPARFILE / RROW (PRE/mpar.f90:83-394): every record is the text after the first '=' of the next line
    that has one, read list-directed.  RSP parameters are rounded to single precision on input."""
    lines = [ln for ln in text.splitlines() if "=" in ln]
    if len(lines) < len(HAP_RECORDS):
        raise ValueError("error when reading the parameter file")
    h = {}
    for (keys, kind), ln in zip(HAP_RECORDS, lines):
        toks = ln[ln.index("=") + 1:].replace(",", " ").split()
        for k, tok in zip(keys, toks):
            if kind == "i":
                h[k] = int(float(tok))
            elif kind == "s":
                h[k] = F32(_fnum(tok))
            else:
                h[k] = _fnum(tok)
        if len(toks) < len(keys):
            raise ValueError("error when reading the parameter file, " + keys[0])
    if math.fmod(h["delta_x"], h["dr"]) > EPS64:                       # mpar.f90:281 (epsilon(delta_x/dr))
        raise ValueError("DEM resolution is not a multiple of the rivulet spacing!")
    return h


def fmt_e(x: float, w: int, d: int) -> str:
    """Fortran Ew.d (no scale factor): 0.ddddE+ee."""
    x = float(x)
    if x == 0.0:
        s = "0." + "0" * d + "E+00"
    else:
        m, e = ("%.*E" % (d - 1, abs(x))).split("E")
        e = int(e) + 1
        s = ("-" if x < 0 else "") + "0." + m.replace(".", "") + "E%+03d" % e
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def fmt_f(x: float, w: int, d: int) -> str:
    s = "%.*f" % (d, float(x))
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def fmt_i(i: int, w: int) -> str:
    s = str(int(i))
    return s.rjust(w) if len(s) <= w else "*" * w


def format_hap(h: dict) -> str:
    """WPARFILE (PRE/mpar.f90:400-541)."""
    bar = "-" * 78
    sp = " "
    L = [bar, "STRUCTURAL PARAMETERS", bar,
         "Grid spacing along the x-direction = " + sp * 20 + fmt_f(h["delta_x"], 10, 2),
         "Grid spacing along the y-direction = " + sp * 20 + fmt_f(h["delta_y"], 10, 2),
         "DEM rectangle size along the x-direction = " + sp * 14 + fmt_i(h["N"], 7),
         "DEM rectangle size along the y-direction = " + sp * 14 + fmt_i(h["M"], 7),
         "Number of cells within the catchment = " + sp * 15 + fmt_i(h["N_celle"], 10),
         "X low left corner coordinate = " + sp * 22 + fmt_f(h["xllcorner"], 20, 8),
         "Y low left corner coordinate = " + sp * 22 + fmt_f(h["yllcorner"], 20, 8),
         bar, "TERRAIN ANALYSIS PARAMETERS", bar,
         "Depit threshold slope = " + sp * 38 + fmt_e(h["pt"], 10, 3),
         "Drainage directions method (LAD:1,LTD:2) = " + sp * 17 + fmt_i(h["imethod"], 4),
         "Upstream deviation memory factor (CBM:0,PBM:1) = " + sp * 13 + fmt_e(h["lambda_"], 10, 3),
         "Threshold on the contour curvature (NDM:-1E10;DM:+1E10) = " + sp * 4 + fmt_e(h["CC_threshold"], 10, 3),
         "Nondispersive channel flow (0:not-required;1:required) = " + sp * 6 + fmt_i(h["ndcf"], 1),
         "Channel initiation method (A:1,AS**k:2,ND:3) = " + sp * 13 + fmt_i(h["nchc"], 4),
         "Threshold on the support area (A) = " + sp * 26 + fmt_e(h["A_threshold"], 16, 9),
         "Threshold on the AS**k function = " + sp * 23 + fmt_f(h["ASk_threshold"], 10, 2),
         "Exponent k of the AS**k function = " + sp * 22 + fmt_f(h["kas"], 10, 2),
         "Threshold on the normalized divergence (ND) = " + sp * 16 + fmt_e(h["DN_threshold"], 10, 3),
         "Path threshold slope = " + sp * 39 + fmt_e(h["local_slope_t"], 10, 3),
         "Drainage direction of the outlet cell (if necessary...)  = " + sp * 4 + fmt_i(h["p_outflow_vo"], 1),
         "Boundary channel constraction (No:0,Yes:1) =" + sp * 19 + fmt_i(h["bcc"], 1),
         "Coefficient for boundary channel elevation definition =" + sp * 7 + fmt_f(h["cqm"], 5, 2),
         "Coefficient for outlet cell elevation definition =" + sp * 12 + fmt_f(h["cqg"], 5, 2),
         bar, "RIVULET NETWORK PARAMETERS (HYDRAULIC GEOMETRY OF THE SINGLE RIVULET)", bar,
         "Rivulet spacing = " + sp * 30 + fmt_f(h["dr"], 10, 3),
         "Reference drainage area (As_rf) = " + sp * 18 + fmt_e(h["As_rf"], 19, 12),
         "Flow discharge (Qsf_rf,w_rf) = " + sp * 17 + fmt_f(h["Qsf_rf"], 10, 3) + sp * 10 + fmt_f(h["w_rf"], 10, 3),
         "Water-surface width (Wsf_rf,b1_rf,b2_rf) = " + sp * 5 + "".join(fmt_f(h[k], 10, 3) for k in ("Wsf_rf", "b1_rf", "b2_rf")),
         "Resistance coefficient (kSsf_rf,y1_rf,y2_rf) = " + sp * 1 + "".join(fmt_f(h[k], 10, 3) for k in ("kSsf_rf", "y1_rf", "y2_rf")),
         "Initial flow discharge (Qsi_rf) = " + sp * 14 + fmt_f(h["Qsi_rf"], 10, 3),
         bar, "CHANNEL NETWORK PARAMETERS", bar,
         "Reference drainage area (As_cf) = " + sp * 18 + fmt_e(h["As_cf"], 19, 12),
         "Flow discharge (Qsf_cf,w_cf) = " + sp * 17 + fmt_f(h["Qsf_cf"], 10, 3) + sp * 10 + fmt_f(h["w_cf"], 10, 3),
         "Water-surface width (Wsf_cf,b1_cf,b2_cf) = " + sp * 5 + "".join(fmt_f(h[k], 10, 3) for k in ("Wsf_cf", "b1_cf", "b2_cf")),
         "Resistance coefficient (kSsf_cf,y1_cf,y2_cf) = " + sp * 1 + "".join(fmt_f(h[k], 10, 3) for k in ("kSsf_cf", "y1_cf", "y2_cf")),
         "Initial flow discharge (Qsi_cf) = " + sp * 14 + fmt_f(h["Qsi_cf"], 10, 3),
         bar]
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------- DEM input
def read_dtm13(text: str, N: int, M: int) -> np.ndarray:
    """dtm_13.val as WBB_SR reads it (PRE/wbb_sr.f90:66-88): M list-directed records of N values, northmost row first.
    Returns rows[M][N] with rows[0] = row j = M."""
    lines = text.splitlines()
    pos = 0
    rows = []
    for _ in range(M):
        vals: list[float] = []
        while len(vals) < N:
            if pos >= len(lines):
                raise ValueError("insufficient data in the file dtm_13.val")
            vals.extend(_fnum(t) for t in lines[pos].replace(",", " ").split())
            pos += 1
        rows.append(vals[:N])
    return np.array(rows, dtype=np.float64)


# --------------------------------------------------------------------------- quicksort
def qsort(n: int, arr: list, brr: list) -> None:
    """QSORT (PRE/qsort.f90:9-125): ascending, in place, arr/brr are 1-based lists (index 0 unused).
    Unstable; the permutation of equal keys is part of the result (qoi_a is written from it)."""
    MM, NSTACK = 7, 50
    istack = [0] * (NSTACK + 1)
    jstack = 0
    l, ir = 1, n
    while True:
        if ir - l < MM:
            for j in range(l + 1, ir + 1):
                a, b = arr[j], brr[j]
                i = j - 1
                while i >= 1:
                    if arr[i] <= a:
                        break
                    arr[i + 1] = arr[i]
                    brr[i + 1] = brr[i]
                    i -= 1
                arr[i + 1] = a
                brr[i + 1] = b
            if jstack == 0:
                return
            ir = istack[jstack]
            l = istack[jstack - 1]
            jstack -= 2
        else:
            k = (l + ir) // 2
            arr[k], arr[l + 1] = arr[l + 1], arr[k]
            brr[k], brr[l + 1] = brr[l + 1], brr[k]
            if arr[l + 1] > arr[ir]:
                arr[l + 1], arr[ir] = arr[ir], arr[l + 1]
                brr[l + 1], brr[ir] = brr[ir], brr[l + 1]
            if arr[l] > arr[ir]:
                arr[l], arr[ir] = arr[ir], arr[l]
                brr[l], brr[ir] = brr[ir], brr[l]
            if arr[l + 1] > arr[l]:
                arr[l + 1], arr[l] = arr[l], arr[l + 1]
                brr[l + 1], brr[l] = brr[l], brr[l + 1]
            i, j = l + 1, ir
            a, b = arr[l], brr[l]
            while True:
                i += 1
                while arr[i] < a:
                    i += 1
                j -= 1
                while arr[j] > a:
                    j -= 1
                if j < i:
                    break
                arr[i], arr[j] = arr[j], arr[i]
                brr[i], brr[j] = brr[j], brr[i]
            arr[l] = arr[j]
            arr[j] = a
            brr[l] = brr[j]
            brr[j] = b
            jstack += 2
            if jstack > NSTACK:
                raise RuntimeError("NSTACK too small!")
            if ir - i + 1 >= j - l:
                istack[jstack] = ir
                istack[jstack - 1] = i
                ir = j - 1
            else:
                istack[jstack] = j - 1
                istack[jstack - 1] = l
                l = i


def _div32(a, b) -> np.float32:
    """a / b with IEEE semantics (x / 0 = inf or NaN, as the Fortran executable computes it), rounded to REAL(4)."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return np.float32(np.float64(a) / np.float64(b))


def facet(e0: float, e1: float, e2: float, dx: float, dy: float):
    """FACET (PRE/facet.f90:10-47): aspect r and slope of the steepest direction inside one triangular facet."""
    pi = 4.0 * math.atan(1.0)
    s1 = (e0 - e1) / dx
    s2 = (e1 - e2) / dx
    if abs(s1) < EPS64:
        r = pi / 2.0 if s2 >= 0.0 else -pi / 2.0
    else:
        r = math.atan(s2 / s1)
    sp = math.sqrt(s1 * s1 + s2 * s2)
    sd = (e0 - e2) / math.sqrt(dx * dx + dy * dy)
    if 0.0 <= r <= pi / 4.0 and s1 >= 0.0:
        return r, sp
    if s1 > sd:
        return 0.0, s1
    return pi / 4.0, sd


# facets in the order DSF / SMEAN visit them: (index of e1, index of e2, sigma); e(1..9), e(5) = the cell (dsf.f90:103-245)
FACETS = [(2, 1, +1.0), (2, 3, -1.0), (6, 3, +1.0), (6, 9, -1.0), (8, 9, +1.0), (8, 7, -1.0), (4, 7, +1.0), (4, 1, -1.0)]


class Prepro:
    """State of one pre-processor run; arrays are indexed by i_basin = (i-1)*M + j, 1-based (index 0 unused)."""

    def __init__(self, hap_text: str, dtm13_text: str):
        self.hap_text_in = hap_text
        self.h = parse_hap(hap_text)
        self.messages: list[str] = []
        h = self.h
        self.N, self.M = h["N"], h["M"]
        N, M = self.N, self.M
        rows = read_dtm13(dtm13_text, N, M)
        nb = N * M + 1
        self.present = np.zeros(nb, dtype=bool)
        self.quota = np.zeros(nb)
        nodata = -9999.0
        # WBB_SR: records in file order (wbb_sr.f90:66-88); quota_min over ALL values with nodata replaced (:176-199)
        n_rec = 0
        for r in range(M):
            j = M - r
            for i in range(1, N + 1):
                v = rows[r, i - 1]
                if v > nodata:
                    n_rec += 1
                    ib = (i - 1) * M + j
                    self.present[ib] = True
                    self.quota[ib] = v
        self.N_celle = n_rec
        h["N_celle"] = n_rec
        quota_min = float(F32(8844.43))                           # wbb_sr.f90:185: a single-precision literal
        for r in range(M):
            rr = np.where(rows[r] > nodata, rows[r], quota_min)
            quota_min = min(quota_min, float(rr.min()))
        if h["bcc"] != 0:
            self._boundary_channel(quota_min)
        z = lambda dt: np.zeros(nb, dtype=dt)  # noqa: E731
        self.p1, self.p2, self.dmID, self.hcID = z(np.int32), z(np.int32), z(np.int32), z(np.int32)
        self.A_inflow, self.sumdev_num = z(np.float64), z(np.float64)
        for k in ("w_1", "w_2", "ls_1", "ls_2", "Ws_1", "Ws_2", "b1", "kSs_1", "kSs_2", "y1", "ASk", "DN", "epl_1", "epl_2", "nrc"):
            setattr(self, k, z(np.float32))
        self.lakes_map, self.q_output = z(np.int32), z(np.int32)
        self.zone = np.ones(nb, dtype=np.int32)
        self.hap_text_out = format_hap(h)                           # cppp.f90:27

    # ---- helpers
    def q(self, ib: int) -> float:
        """dtm_quota (PRE/mbbio.f90:766-776): -1 outside the catchment."""
        return self.quota[ib] if self.present[ib] else -1.0

    def ij(self, ib: int):
        jr = ib % self.M
        if jr != 0:
            return (ib - jr) // self.M + 1, jr
        return ib // self.M, self.M

    def _boundary_channel(self, quota_min: float) -> None:
        """WBB_SR boundary channel (PRE/wbb_sr.f90:95-160)."""
        h, N, M = self.h, self.N, self.M
        quota_gronda = quota_min * float(h["cqm"])
        quota_chiusura = quota_gronda * float(h["cqg"])
        cnt = 0
        last = 0
        for j in range(M, 0, -1):
            for i in range(1, N + 1):
                ib = (i - 1) * M + j
                if self.q(ib) < 0.0:
                    continue
                flag = False
                for ii in range(i - 1, i + 2):
                    if flag:
                        break
                    for jj in range(j - 1, j + 2):
                        if ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1:
                            flag = True
                        elif self.q((ii - 1) * M + jj) < 0.0:
                            flag = True
                        if flag:
                            cnt += 1
                            last = ib
                            self.quota[ib] = quota_gronda
                            break
        self.quota[last] = quota_chiusura
        sq2 = float(F32(math.sqrt(F32(2.0))))                       # sqrt(2.0): single precision intrinsic
        if quota_gronda + (cnt * h["delta_x"] * sq2 * h["pt"]) >= quota_min:
            raise ValueError("boundary channel: a smaller coefficient for boundary channel elevation definition is needed")

    # ---- CSORT
    def csort(self) -> None:
        N, M = self.N, self.M
        ibs = [0]
        qo = [0.0]
        for i in range(1, N + 1):
            for j in range(1, M + 1):
                ib = (i - 1) * M + j
                if self.present[ib]:
                    ibs.append(ib)
                    qo.append(float(self.quota[ib]))
        qsort(self.N_celle, qo, ibs)
        self.qoi = [0] + ibs[:0:-1]                                  # descending, 1-based

    # ---- DEPIT
    def depit(self) -> int:
        h, N, M, nc = self.h, self.N, self.M, self.N_celle
        eps = h["pt"] * h["delta_x"]
        ib_l = self.qoi[nc]
        if nc > 1:
            ib_sl = self.qoi[nc - 1]
            if (self.q(ib_sl) - self.q(ib_l)) < EPS64 * h["delta_x"]:
                raise ValueError("catchment with more than one outlet cell!")
        pit1 = [0] * (nc + 1)
        for i_qo in range(1, nc + 1):
            ib = self.qoi[i_qo]
            if self.q(ib) < 0.0:
                raise ValueError("negative elevation inside the catchment")
            pit1[nc - i_qo + 1] = ib
        n_pits = nc
        total = 0
        rec = np.cumsum(self.present) * self.present               # record number of a cell (any injective map does)
        while True:
            nn_mod = 0
            flagged = set()
            pit2 = [0]
            for n_pit in range(1, n_pits + 1):
                ib = pit1[n_pit]
                qc = self.q(ib)
                if ib == ib_l:
                    continue
                i, j = self.ij(ib)
                qmin = math.inf
                lower = False
                for ii in range(i - 1, i + 2):
                    for jj in range(j - 1, j + 2):
                        if (ii == i and jj == j) or ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1:
                            continue
                        qcc = self.q((ii - 1) * M + jj)
                        if qcc < 0.0:
                            continue
                        if qcc < qc:
                            lower = True
                            break
                        if qcc < qmin:
                            qmin = qcc
                    if lower:
                        break
                if lower:
                    continue
                if qc <= qmin:
                    self.quota[ib] = qmin + eps
                    total += 1
                    nn_mod += 1
                    for ii in range(i - 1, i + 2):
                        for jj in range(j - 1, j + 2):
                            if (ii == i and jj == j) or ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1:
                                continue
                            iib = (ii - 1) * M + jj
                            if rec[iib] == 0:
                                continue
                            if iib not in flagged:
                                flagged.add(iib)
                                pit2.append(iib)
            if nn_mod == 0:
                break
            nn = len(pit2) - 1
            qp = [0.0] + [float(self.quota[b]) for b in pit2[1:]]
            qsort(nn, qp, pit2)
            pit1 = pit2
            n_pits = nn
        self.n_modifiche = total
        return total

    # ---- CCA
    def cca(self) -> None:
        self.h = parse_hap(self.hap_text_out)                        # cca.f90:31 re-reads the REWRITTEN hap.in
        self.h["N_celle"] = self.N_celle
        h, N, M = self.h, self.N, self.M
        dx = h["delta_x"]
        dx2 = dx * dx
        Kp = F32(0.0)                                                # uninitialised local in the reference; carried over border cells
        p_small = float(F32(1.0e-9))
        cct = float(h["CC_threshold"])
        for j in range(M, 0, -1):
            for i in range(1, N + 1):
                ib = (i - 1) * M + j
                if not self.present[ib]:
                    continue
                dm = 2
                mesh = {}
                ok = True
                for ii in range(i - 1, i + 2):
                    for jj in range(j - 1, j + 2):
                        if ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1 or not self.present[(ii - 1) * M + jj]:
                            ok = False
                            break
                        v = float(self.quota[(ii - 1) * M + jj])
                        if v == 0.0:
                            v = -9999.0
                        mesh[(ii - i + 2, jj - j + 2)] = v
                    if not ok:
                        break
                if ok:
                    zx = (mesh[2, 3] - mesh[2, 1]) / (2 * dx)
                    zy = (mesh[1, 2] - mesh[3, 2]) / (2 * dx)
                    zxx = (mesh[2, 3] - 2 * mesh[2, 2] + mesh[2, 1]) / dx2
                    zyy = (mesh[1, 2] - 2 * mesh[2, 2] + mesh[3, 2]) / dx2
                    zxy = (-mesh[1, 1] + mesh[1, 3] + mesh[3, 1] - mesh[3, 3]) / (4 * dx2)
                    if abs(zx) > EPS64 or abs(zy) > EPS64:
                        p = zx * zx + zy * zy
                        q = p + 1.0
                    else:
                        p = p_small
                        q = p + 1.0
                    Kc = (zxx * zy * zy - 2.0 * zxy * zx * zy + zyy * zx * zx) / math.pow(p, 1.5)
                    if abs(Kc) < EPS64:
                        Kc = 0.0
                    Kp = F32((zxx * zx * zx + 2.0 * zxy * zx * zy + zyy * zy * zy) / (p * math.pow(q, 1.5)))
                    dm = 1 if Kc < cct else 2
                self.dmID[ib] = dm
                self.DN[ib] = Kp

    # ---- window of a cell as DSF / SMEAN build it (0 = absent)
    def window(self, i: int, j: int):
        N, M = self.N, self.M
        e = [0.0] * 10
        l = 0
        for ii in range(i - 1, i + 2):
            for jj in range(j - 1, j + 2):
                l += 1
                if ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1:
                    continue
                qv = self.q((ii - 1) * M + jj)
                if qv < 0.0:
                    continue
                e[l] = qv
        return e

    # ---- SMEAN
    def smean(self) -> None:
        h, nc = self.h, self.N_celle
        ib_out = self.qoi[nc]
        n_s, ssum = 0.0, 0.0
        for n in range(1, nc + 1):
            ib = self.qoi[n]
            if self.p1[ib] != 0 or self.p2[ib] != 0 or ib == ib_out:
                continue
            i, j = self.ij(ib)
            e = self.window(i, j)
            s_max = 0.0
            for a, b, _ in FACETS:
                if e[a] * e[b] != 0.0:
                    _, s = facet(e[5], e[a], e[b], h["delta_x"], h["delta_y"])
                    if s > s_max:
                        s_max = s
            n_s += 1.0
            ssum += s_max
        self.mean_s_max = ssum / n_s if n_s else float("nan")

    def channel_initiation(self, A_outflow: float, ASk, DN) -> int:
        h = self.h
        if h["nchc"] == 1:
            return 0 if A_outflow <= h["A_threshold"] else 1
        if h["nchc"] == 2:
            return 0 if F32(ASk) <= h["ASk_threshold"] else 1
        if h["nchc"] == 3:
            return 0 if F32(DN) >= h["DN_threshold"] else 1
        raise ValueError("nchc out of range!")

    # ---- DSF
    def dsf(self) -> None:
        h, N, M, nc = self.h, self.N, self.M, self.N_celle
        dx = h["delta_x"]
        pi = 4.0 * math.atan(1.0)
        rad2 = math.sqrt(2.0)
        dxy = math.sqrt(2.0) * dx
        A_cell = dx * dx
        lam = h["lambda_"]
        kas = float(h["kas"])
        ib_out = self.qoi[nc]
        hcID = 0                                                     # local of the reference, NOT reset per cell (dsf.f90:64)
        i = j = 0
        ib = 0
        for n in range(1, nc + 1):
            ib = self.qoi[n]
            if self.p1[ib] != 0 or self.p2[ib] != 0:
                continue
            i, j = self.ij(ib)
            if ib == ib_out:
                continue
            e = self.window(i, j)
            e0 = e[5]
            s_max = 0.0
            for a, b, sg in FACETS:
                if e[a] * e[b] != 0.0:
                    r, s = facet(e0, e[a], e[b], dx, h["delta_y"])
                    if s > s_max:
                        e1f, e2f, r_max, s_max, po1, po2, sigma = e[a], e[b], r, s, a, b, sg
            A_in = float(self.A_inflow[ib])
            A_out = A_in + A_cell
            sdn = float(self.sumdev_num[ib])
            sumdev = 0.0 if A_in == 0.0 else sdn / A_in
            if s_max > 0.0:
                if h["imethod"] == 1:
                    dev_1 = r_max
                    dev_2 = pi / 4 - r_max
                elif h["imethod"] == 2:
                    dev_1 = dx * math.sin(r_max)
                    dev_2 = dx * rad2 * math.sin(pi / 4.0 - r_max)
                else:
                    raise ValueError("unespected imethod!")
                if sigma == 1.0:
                    dev_2 = -dev_2
                else:
                    dev_1 = -dev_1
                if abs(dev_1) <= EPS64 or abs(dev_2) <= EPS64:
                    sumdev = 0.0
                dm = int(self.dmID[ib])
                hcID = int(self.hcID[ib])
                Kp = self.DN[ib]
                epl_1 = F32(dx)
                epl_2 = F32(dxy)
                ls_1 = F32((e0 - e1f) / float(epl_1))
                ls_2 = F32((e0 - e2f) / float(epl_2))
                ASk = F32(A_out * math.pow(s_max, kas))
                sumdev_1 = lam * sumdev + dev_1
                sumdev_2 = lam * sumdev + dev_2
                DN = _div32(Kp, -self.mean_s_max)
                if hcID == 0:
                    hcID = self.channel_initiation(A_out, ASk, DN)
                if hcID == 1 and h["ndcf"] == 1:
                    dm = 2
                a1, a2 = abs(sumdev_1), abs(sumdev_2)
                if dm == 1:
                    if abs(dev_1) <= EPS64:
                        w_1, w_2 = F32(1.0), F32(0.0)
                    elif abs(dev_2) <= EPS64:
                        w_1, w_2 = F32(0.0), F32(1.0)
                    else:
                        w_1 = F32(a2 / (a1 + a2))
                        w_2 = F32(a1 / (a1 + a2))
                        if w_1 < F32(1.0e-6):
                            w_1, w_2 = F32(0.0), F32(1.0)
                        if w_2 < F32(1.0e-6):
                            w_1, w_2 = F32(1.0), F32(0.0)
                elif dm == 2:
                    if abs(a1 - a2) / dx < 10e-14 and (e0 - e1f) > 0.0:
                        w_1, w_2, epl_2, ls_2 = F32(1.0), F32(0.0), F32(0.0), F32(0.0)
                    elif a1 < a2 and (e0 - e1f) > 0.0:
                        w_1, w_2, epl_2, ls_2 = F32(1.0), F32(0.0), F32(0.0), F32(0.0)
                    elif a1 > a2 or (e0 - e2f) > 0.0:
                        w_1, w_2, epl_1, ls_1 = F32(0.0), F32(1.0), F32(0.0), F32(0.0)
                    else:
                        raise ValueError("s_max < 0, unexpected case!")
                else:
                    raise ValueError("unexpected case!")
                cv1 = ib + M * _DI[po1] + _DJ[po1]
                cv2 = ib + M * _DI[po2] + _DJ[po2]
                if not (self.present[cv1] and self.present[cv2]):
                    raise ValueError("dtm_A_inflow")                 # the reference stops here (mbbio.f90:814-824)
                A1 = float(self.A_inflow[cv1])
                S1 = float(self.sumdev_num[cv1])
                A2 = float(self.A_inflow[cv2])
                S2 = float(self.sumdev_num[cv2])
                A1 = A1 + (A_out * float(w_1))
                A2 = A2 + (A_out * float(w_2))
                S1 = S1 + (A_out * float(w_1) * sumdev_1)
                S2 = S2 + (A_out * float(w_2) * sumdev_2)
                self.w_1[ib], self.w_2[ib], self.p1[ib], self.p2[ib], self.dmID[ib] = w_1, w_2, po1, po2, dm
                self.epl_1[ib], self.epl_2[ib], self.ls_1[ib], self.ls_2[ib] = epl_1, epl_2, ls_1, ls_2
                self.ASk[ib], self.DN[ib], self.hcID[ib] = ASk, DN, hcID
                self.A_inflow[cv1] = A1
                self.A_inflow[cv2] = A2                              # cv1 == cv2 never happens (two different neighbours)
                self.sumdev_num[cv1] = S1
                self.sumdev_num[cv2] = S2
                if abs(float(w_1)) > EPS32 and self.hcID[cv1] == 0:
                    self.hcID[cv1] = hcID
                if abs(float(w_2)) > EPS32 and self.hcID[cv2] == 0:
                    self.hcID[cv2] = hcID
            elif abs(s_max) < EPS64:
                Kp = self.DN[ib]
                emin = math.inf
                pL = 0
                for l in range(1, 10):
                    if e[l] != 0.0 and l != 5 and e[l] < emin:
                        emin = e[l]
                        pL = l
                if pL == 0:
                    raise ValueError("s_max = 0, unexpected case!")
                if pL % 2 == 0:
                    epl_1 = F32(dx)
                    ls_1 = F32((e0 - emin) / float(epl_1))
                    ASk = F32(A_out * math.pow(float(ls_1), kas))
                    DN = _div32(Kp, -self.mean_s_max)
                    if hcID == 0:
                        hcID = self.channel_initiation(A_out, ASk, DN)
                    cv1 = ib + M * _DI[pL] + _DJ[pL]
                    self.A_inflow[cv1] = float(self.A_inflow[cv1]) + A_out * 1.0
                    self.w_1[ib], self.p1[ib], self.epl_1[ib], self.ls_1[ib] = F32(1.0), pL, epl_1, ls_1
                    if self.hcID[cv1] == 0:
                        self.hcID[cv1] = hcID
                else:
                    epl_2 = F32(dxy)
                    ls_2 = F32((e0 - emin) / float(epl_2))
                    ASk = F32(A_out * math.pow(float(ls_2), kas))
                    DN = _div32(Kp, -self.mean_s_max)
                    sumdev_2 = lam * sumdev
                    cv2 = ib + M * _DI[pL] + _DJ[pL]
                    self.A_inflow[cv2] = float(self.A_inflow[cv2]) + A_out * 1.0
                    self.sumdev_num[cv2] = float(self.sumdev_num[cv2]) + A_out * 1.0 * sumdev_2
                    self.w_2[ib], self.p2[ib], self.epl_2[ib], self.ls_2[ib] = F32(1.0), pL, epl_2, ls_2
                    if self.hcID[cv2] == 0:
                        self.hcID[cv2] = hcID
                self.ASk[ib], self.DN[ib], self.hcID[ib] = ASk, DN, hcID
            else:
                raise ValueError("s_max < 0")
        # phantom channel end of the outlet cell (dsf.f90:531-600); i, j, ib are those of the outlet cell
        nvo = 0
        A_max = 0.0
        p_out = None
        ls_out = F32(0.0)
        for ii in range(i - 1, i + 2):
            for jj in range(j - 1, j + 2):
                if ii == 0 or ii == N + 1 or jj == 0 or jj == M + 1:
                    continue
                iib = (ii - 1) * M + jj
                if not self.present[iib]:
                    continue
                p_in = 3 * (ii - i) + (jj - j) + 5
                for pk, wk, lsk in ((self.p1, self.w_1, self.ls_1), (self.p2, self.w_2, self.ls_2)):
                    if p_in + int(pk[iib]) == 10:
                        Ao = (float(self.A_inflow[iib]) + A_cell) * float(wk[iib])
                        if Ao > A_max:
                            A_max = Ao
                            p_out = int(pk[iib])
                            ls_out = lsk[iib]
                            ivo = i + _DI[p_out]
                            jvo = j + p_out - 5 - 3 * (ivo - i)
                            ivb = ib + (M * (ivo - i) + (jvo - j))
                            nvo = 1 if (ivo == 0 or ivo == N + 1 or jvo == 0 or jvo == M + 1 or not self.present[ivb]) else 0
        if nvo == 0:
            p_out = h["p_outflow_vo"]
        if p_out % 2 == 0:
            self.p1[ib], self.w_1[ib], self.epl_1[ib], self.ls_1[ib] = p_out, F32(1.0), F32(dx), ls_out
        else:
            self.p2[ib], self.w_2[ib], self.epl_2[ib], self.ls_2[ib] = p_out, F32(1.0), F32(dxy), ls_out

    # ---- HG
    def hg(self) -> None:
        h = self.h
        A_cell = h["delta_x"] * h["delta_y"]
        for n in range(1, self.N_celle + 1):
            ib = self.qoi[n]
            A_out = float(self.A_inflow[ib]) + A_cell
            w_1, w_2 = self.w_1[ib], self.w_2[ib]
            sfx = "_rf" if self.hcID[ib] == 0 else "_cf"
            As, Qsf, w_, Wsf, b1r, b2r = h["As" + sfx], h["Qsf" + sfx], h["w" + sfx], h["Wsf" + sfx], h["b1" + sfx], h["b2" + sfx]
            kS, y1r, y2r = h["kSsf" + sfx], h["y1" + sfx], h["y2" + sfx]
            RA = A_out / As
            if self.b1[ib] == 0.0:
                self.b1[ib] = b1r
            b1 = self.b1[ib]

            def law(c, ex1, ex2, w):
                # c * Q**(-ex1) in single precision, times (RA*w)**(w_*(ex2-ex1)) in double, stored single (hg.f90:84-106)
                lead = F32(c * F32(math.pow(float(Qsf), float(-ex1))))
                return F32(float(lead) * math.pow(RA * float(w), float(F32(w_ * F32(ex2 - ex1)))))
            if self.Ws_1[ib] == 0.0 and abs(float(w_1)) > EPS32:
                self.Ws_1[ib] = law(Wsf, b1, b2r, w_1)
            if self.Ws_2[ib] == 0.0 and abs(float(w_2)) > EPS32:
                self.Ws_2[ib] = law(Wsf, b1, b2r, w_2)
            if self.y1[ib] == 0.0:
                self.y1[ib] = y1r
            y1 = self.y1[ib]
            if self.kSs_1[ib] == 0.0 and abs(float(w_1)) > EPS32:
                self.kSs_1[ib] = law(kS, y1, y2r, w_1)
            if self.kSs_2[ib] == 0.0 and abs(float(w_2)) > EPS32:
                self.kSs_2[ib] = law(kS, y1, y2r, w_2)
            if self.nrc[ib] == 0.0:
                self.nrc[ib] = F32(h["delta_x"] / h["dr"]) if self.hcID[ib] == 0 else F32(1.0)

    def run(self) -> "Prepro":
        self.csort()
        self.depit()
        self.csort()
        self.cca()
        self.smean()
        self.dsf()
        self.hg()
        return self

    # ---- MRBB_SR / RBB
    def raster(self, name: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> str:
        h, N, M = self.h, self.N, self.M
        kind, arr = RASTERS[name]
        vals = getattr(self, arr)
        grid = vals[1:].reshape(N, M)                                # [i-1][j-1]
        pres = self.present[1:].reshape(N, M)
        if ht == 2:
            head = ("north: %s\nsouth: %s\neast:  %s\nwest:  %s\nrows:  %s\ncols:  %s\n"
                    % (fmt_i(0, 5), fmt_f(h["yllcorner"], 20, 8), fmt_i(0, 5), fmt_f(h["xllcorner"], 20, 8), fmt_i(M, 5), fmt_i(N, 5)))
        elif ht == 1:
            head = ("ncols" + " " * 8 + fmt_i(N, 5) + "\nnrow" + " " * 9 + fmt_i(M, 5) + "\nxllcorner " + fmt_f(h["xllcorner"], 20, 8)
                    + "\nyllcorner " + fmt_f(h["yllcorner"], 20, 8) + "\ncellsize" + " " * 8 + fmt_f(h["delta_x"], 6, 2)
                    + "\nNODATA_value" + " " * 4 + fmt_i(-9999, 5) + "\n")
        else:
            head = ""
        out = [head]
        if kind == "i":
            if ips == 2 and name.startswith("dtm_p_outflow"):
                jd = np.array([0, 8, 16, 32, 4, 0, 64, 2, 1, 128])
                vals_i = jd[grid]
            else:
                vals_i = grid.astype(np.int64)
            g = np.where(pres, vals_i, int(nodata))
            imax = max(int(vals_i[pres].max()), abs(int(nodata)))
            imin = min(int(vals_i[pres].min()), int(nodata))
            if imax > 0:
                w = int(math.log10(float(F32(imax)))) + (3 if imin < 0 else 2)
            else:
                w = 2
            for j in range(M, 0, -1):
                out.append("".join(fmt_i(v, w) for v in g[:, j - 1]) + "\n")
        else:
            g = np.where(pres, grid.astype(np.float64), float(F32(nodata)))
            rmin = min(float(g[pres].min()), float(F32(nodata)))
            if name == "dtm_A_inflow":
                f = (lambda v: fmt_f(v, 15, 2)) if rmin < 0.0 else (lambda v: fmt_f(v, 14, 2))
            else:
                f = (lambda v: fmt_e(v, 20, 12)) if rmin < 0.0 else (lambda v: fmt_e(v, 21, 12))
            for j in range(M, 0, -1):
                out.append("".join(f(v) for v in g[:, j - 1]) + "\n")
        return "".join(out)

    def qoi_a(self) -> str:
        """hg.f90:31-37: list-directed integers, one per record."""
        return "".join("%12d\n" % v for v in [self.N_celle] + self.qoi[1:])

    def write(self, directory: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> None:
        with open(os.path.join(directory, "hap.in"), "w") as fh:
            fh.write(self.hap_text_out)
        for name in RASTERS:
            with open(os.path.join(directory, name), "w") as fh:
                fh.write(self.raster(name, ht, nodata, ips))
        with open(os.path.join(directory, "qoi_a"), "w") as fh:
            fh.write(self.qoi_a())


# p_outflow = 3*di + dj + 5 (dsf.f90:368-375): di along i (x), dj along j (y)
_DI = [0, -1, -1, -1, 0, 0, 0, 1, 1, 1]
_DJ = [0, -1, 0, 1, -1, 0, 1, -1, 0, 1]

# files MRBB_SR writes (mrbb_sr.f90:76-230), in its order: name -> (integer / real, attribute)
RASTERS = {
    "dem": ("r", "quota"), "lakes_map": ("i", "lakes_map"), "zone": ("i", "zone"), "dtm_w_1": ("r", "w_1"), "dtm_w_2": ("r", "w_2"),
    "dtm_p_outflow_1": ("i", "p1"), "dtm_p_outflow_2": ("i", "p2"), "dtm_A_inflow": ("r", "A_inflow"),
    "dtm_local_slope_1": ("r", "ls_1"), "dtm_local_slope_2": ("r", "ls_2"), "dtm_epl_1": ("r", "epl_1"), "dtm_epl_2": ("r", "epl_2"),
    "dtm_kSs1_sf_1": ("r", "kSs_1"), "dtm_kSs1_sf_2": ("r", "kSs_2"), "dtm_Ws1_sf_1": ("r", "Ws_1"), "dtm_Ws1_sf_2": ("r", "Ws_2"),
    "dtm_b1_sf": ("r", "b1"), "dtm_y1_sf": ("r", "y1"), "dtm_hcID": ("i", "hcID"), "dtm_q_output": ("i", "q_output"), "dtm_nrc": ("r", "nrc"),
}


def run_directory(directory: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> Prepro:
    """What `./pycppp` does in <project>/prepro with the answers 2, 0, 1 on stdin (PY/cathy_tools.py:378-389)."""
    with open(os.path.join(directory, "hap.in")) as fh:
        hap = fh.read()
    with open(os.path.join(directory, "dtm_13.val")) as fh:
        dtm = fh.read()
    p = Prepro(hap, dtm).run()
    p.write(directory, ht, nodata, ips)
    return p
