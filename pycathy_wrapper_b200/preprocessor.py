"""Drop-in for the CATHY pre-processor `pycppp`.

pyCATHY runs `./pycppp` in <project>/prepro with the answers "2\\n0\\n1\\n" on stdin (header type GRASS, nodata 0,
HAP pointer system; PY/cathy_tools.py:378-389, `CATHY.run_preprocessor` :331).  It reads `hap.in` and `dtm_13.val` and
writes `dem`, `lakes_map`, `zone`, the `dtm_*` rasters and `qoi_a` -- the surface-routing inputs of the processor
(SRC/datin.f:325-372) -- and rewrites `hap.in`.  PRE = /root/reference/examples/SSHydro/weill_exemple/prepro/src.

This module is the host side: text in, text out (PRE/mpar.f90 parser + WPARFILE, PRE/wbb_sr.f90:66-88 reader,
PRE/mrbb_sr.f90 RBB writer, PRE/hg.f90:31-37 qoi_a).  All terrain analysis (CSORT, DEPIT, CCA, SMEAN, DSF, HG) runs on
the GPU through `cathy_prepro_run` (include/cathy_prepro.h, csrc/cathy_prepro.cu).  There is no CPU path: without the
library or a CUDA device the call raises.  Not written: the ESRI shape files of BB2SHP (river_net.shp ...; nothing in
pyCATHY or CATHY reads them), `dtm_Kc.txt`, the binary scratch files basin_b / basin_i / qoi.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import numpy as np

from . import capi

ABI_VERSION = 1
_D = C.POINTER(C.c_double)
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int32)


class PreproError(RuntimeError):
    pass


class CathyPreproParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("N", C.c_int32), ("M", C.c_int32), ("imethod", C.c_int32), ("ndcf", C.c_int32),
        ("nchc", C.c_int32), ("p_outflow_vo", C.c_int32), ("bcc", C.c_int32),
        ("delta_x0", C.c_double), ("pt0", C.c_double), ("cqm0", C.c_float), ("cqg0", C.c_float),
        ("delta_x", C.c_double), ("delta_y", C.c_double), ("lambda_", C.c_double), ("A_threshold", C.c_double),
        ("CC_threshold", C.c_float), ("ASk_threshold", C.c_float), ("kas", C.c_float), ("_pad0", C.c_float),
        ("dr", C.c_double), ("As_rf", C.c_double), ("As_cf", C.c_double),
        ("Qsf_rf", C.c_float), ("w_rf", C.c_float), ("Wsf_rf", C.c_float), ("b1_rf", C.c_float), ("b2_rf", C.c_float),
        ("kSsf_rf", C.c_float), ("y1_rf", C.c_float), ("y2_rf", C.c_float),
        ("Qsf_cf", C.c_float), ("w_cf", C.c_float), ("Wsf_cf", C.c_float), ("b1_cf", C.c_float), ("b2_cf", C.c_float),
        ("kSsf_cf", C.c_float), ("y1_cf", C.c_float), ("y2_cf", C.c_float),
    ]


_OUT_ARRAYS = [("quota", np.float64), ("A_inflow", np.float64), ("w_1", np.float32), ("w_2", np.float32),
               ("local_slope_1", np.float32), ("local_slope_2", np.float32), ("epl_1", np.float32), ("epl_2", np.float32),
               ("Ws1_sf_1", np.float32), ("Ws1_sf_2", np.float32), ("b1_sf", np.float32), ("kSs1_sf_1", np.float32),
               ("kSs1_sf_2", np.float32), ("y1_sf", np.float32), ("nrc", np.float32),
               ("p_outflow_1", np.int32), ("p_outflow_2", np.int32), ("hcID", np.int32), ("dmID", np.int32), ("order", np.int32)]
_CT = {np.float64: _D, np.float32: _F, np.int32: _I}


class CathyPreproOut(C.Structure):
    _fields_ = [(name, _CT[dt]) for name, dt in _OUT_ARRAYS] + [
        ("n_cells", C.c_int32), ("n_modifications", C.c_int32), ("n_waves", C.c_int32), ("n_launches", C.c_int32),
        ("mean_s_max", C.c_double), ("device_ms", C.c_double), ("stage_ms", C.c_double * 8)]


_LIB = None


def load_prepro_library():
    """cathy_prepro_run / cathy_prepro_last_error of libcathy_b200.so; raises when the library is missing."""
    global _LIB
    if _LIB is None:
        path = capi.library_path()
        if not os.path.exists(path):
            raise capi.CathyLibraryError(f"{path} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        try:
            lib = C.CDLL(path)
            run, err = lib.cathy_prepro_run, lib.cathy_prepro_last_error
        except (OSError, AttributeError) as e:
            raise capi.CathyLibraryError(f"cannot bind the pre-processor entry points of {path}: {e}") from e
        run.argtypes = [C.POINTER(CathyPreproParams), _D, C.POINTER(C.c_uint8), C.c_int32, C.POINTER(CathyPreproOut)]
        run.restype = C.c_int32
        err.restype = C.c_char_p
        lib.cathy_prepro_format_real.argtypes = [_D, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        lib.cathy_prepro_format_real.restype = C.c_int64
        lib.cathy_prepro_format_int.argtypes = [_I, C.c_int64, C.c_int64, C.c_int32, C.c_char_p, C.c_int32]
        lib.cathy_prepro_format_int.restype = C.c_int64
        _LIB = (lib, run, err)
    return _LIB


# ----------------------------------------------------------------------------------------------- hap.in
# record order of RPARFILE (PRE/mpar.f90:83-394): names per record and their Fortran kind (d = REP, s = RSP, i = integer)
_RECORDS = (
    ("d", "delta_x"), ("d", "delta_y"), ("i", "N"), ("i", "M"), ("i", "N_celle"), ("d", "xllcorner"), ("d", "yllcorner"),
    ("d", "pt"), ("i", "imethod"), ("d", "lambda_"), ("s", "CC_threshold"), ("i", "ndcf"), ("i", "nchc"), ("d", "A_threshold"),
    ("s", "ASk_threshold"), ("s", "kas"), ("s", "DN_threshold"), ("s", "local_slope_t"), ("i", "p_outflow_vo"), ("i", "bcc"),
    ("s", "cqm"), ("s", "cqg"), ("d", "dr"), ("d", "As_rf"), ("s", "Qsf_rf w_rf"), ("s", "Wsf_rf b1_rf b2_rf"),
    ("s", "kSsf_rf y1_rf y2_rf"), ("s", "Qsi_rf"), ("d", "As_cf"), ("s", "Qsf_cf w_cf"), ("s", "Wsf_cf b1_cf b2_cf"),
    ("s", "kSsf_cf y1_cf y2_cf"), ("s", "Qsi_cf"),
)


def _num(tok: str) -> float:
    return float(tok.upper().replace("D", "E"))


def read_hapin(text: str) -> dict:
    """RPARFILE: the n-th record is what follows the first '=' of the n-th line holding one (RROW, PRE/mpar.f90:380-394)."""
    recs = [ln.split("=", 1)[1] for ln in text.splitlines() if "=" in ln]
    if len(recs) < len(_RECORDS):
        raise PreproError("error when reading the parameter file")
    hap = {}
    for (kind, names), rec in zip(_RECORDS, recs):
        toks = rec.replace(",", " ").split()
        names = names.split()
        if len(toks) < len(names):
            raise PreproError("error when reading the parameter file, " + names[0])
        for name, tok in zip(names, toks):
            v = _num(tok)
            hap[name] = int(v) if kind == "i" else (float(np.float32(v)) if kind == "s" else v)
    if math.fmod(hap["delta_x"], hap["dr"]) > np.finfo(np.float64).eps:        # PRE/mpar.f90:281-285
        raise PreproError("DEM resolution is not a multiple of the rivulet spacing!")
    return hap


def _E(x: float, w: int, d: int) -> str:
    """Fortran Ew.d."""
    if x == 0.0:
        body = "0." + "0" * d + "E+00"
    else:
        mant, ex = ("%.*E" % (d - 1, abs(x))).split("E")
        body = ("-" if x < 0 else "") + "0." + mant[0] + mant[2:] + "E%+03d" % (int(ex) + 1)
    if len(body) > w:
        body = body.replace("0.", ".", 1)
    return body.rjust(w) if len(body) <= w else "*" * w


def _Ff(x: float, w: int, d: int) -> str:
    body = "%.*f" % (d, x)
    if len(body) > w and body.lstrip("-").startswith("0."):
        body = body.replace("0.", ".", 1)
    return body.rjust(w) if len(body) <= w else "*" * w


def _Ii(i: int, w: int) -> str:
    body = "%d" % i
    return body.rjust(w) if len(body) <= w else "*" * w


def write_hapin(hap: dict) -> str:
    """WPARFILE (PRE/mpar.f90:400-541): hap.in as the pre-processor rewrites it after WBB."""
    h = hap
    bar = "-" * 78
    pad = lambda n: " " * n  # noqa: E731
    three = lambda *k: "".join(_Ff(h[x], 10, 3) for x in k)  # noqa: E731
    out = [
        bar, "STRUCTURAL PARAMETERS", bar,
        "Grid spacing along the x-direction = " + pad(20) + _Ff(h["delta_x"], 10, 2),
        "Grid spacing along the y-direction = " + pad(20) + _Ff(h["delta_y"], 10, 2),
        "DEM rectangle size along the x-direction = " + pad(14) + _Ii(h["N"], 7),
        "DEM rectangle size along the y-direction = " + pad(14) + _Ii(h["M"], 7),
        "Number of cells within the catchment = " + pad(15) + _Ii(h["N_celle"], 10),
        "X low left corner coordinate = " + pad(22) + _Ff(h["xllcorner"], 20, 8),
        "Y low left corner coordinate = " + pad(22) + _Ff(h["yllcorner"], 20, 8),
        bar, "TERRAIN ANALYSIS PARAMETERS", bar,
        "Depit threshold slope = " + pad(38) + _E(h["pt"], 10, 3),
        "Drainage directions method (LAD:1,LTD:2) = " + pad(17) + _Ii(h["imethod"], 4),
        "Upstream deviation memory factor (CBM:0,PBM:1) = " + pad(13) + _E(h["lambda_"], 10, 3),
        "Threshold on the contour curvature (NDM:-1E10;DM:+1E10) = " + pad(4) + _E(h["CC_threshold"], 10, 3),
        "Nondispersive channel flow (0:not-required;1:required) = " + pad(6) + _Ii(h["ndcf"], 1),
        "Channel initiation method (A:1,AS**k:2,ND:3) = " + pad(13) + _Ii(h["nchc"], 4),
        "Threshold on the support area (A) = " + pad(26) + _E(h["A_threshold"], 16, 9),
        "Threshold on the AS**k function = " + pad(23) + _Ff(h["ASk_threshold"], 10, 2),
        "Exponent k of the AS**k function = " + pad(22) + _Ff(h["kas"], 10, 2),
        "Threshold on the normalized divergence (ND) = " + pad(16) + _E(h["DN_threshold"], 10, 3),
        "Path threshold slope = " + pad(39) + _E(h["local_slope_t"], 10, 3),
        "Drainage direction of the outlet cell (if necessary...)  = " + pad(4) + _Ii(h["p_outflow_vo"], 1),
        "Boundary channel constraction (No:0,Yes:1) =" + pad(19) + _Ii(h["bcc"], 1),
        "Coefficient for boundary channel elevation definition =" + pad(7) + _Ff(h["cqm"], 5, 2),
        "Coefficient for outlet cell elevation definition =" + pad(12) + _Ff(h["cqg"], 5, 2),
        bar, "RIVULET NETWORK PARAMETERS (HYDRAULIC GEOMETRY OF THE SINGLE RIVULET)", bar,
        "Rivulet spacing = " + pad(30) + _Ff(h["dr"], 10, 3),
        "Reference drainage area (As_rf) = " + pad(18) + _E(h["As_rf"], 19, 12),
        "Flow discharge (Qsf_rf,w_rf) = " + pad(17) + _Ff(h["Qsf_rf"], 10, 3) + pad(10) + _Ff(h["w_rf"], 10, 3),
        "Water-surface width (Wsf_rf,b1_rf,b2_rf) = " + pad(5) + three("Wsf_rf", "b1_rf", "b2_rf"),
        "Resistance coefficient (kSsf_rf,y1_rf,y2_rf) = " + pad(1) + three("kSsf_rf", "y1_rf", "y2_rf"),
        "Initial flow discharge (Qsi_rf) = " + pad(14) + _Ff(h["Qsi_rf"], 10, 3),
        bar, "CHANNEL NETWORK PARAMETERS", bar,
        "Reference drainage area (As_cf) = " + pad(18) + _E(h["As_cf"], 19, 12),
        "Flow discharge (Qsf_cf,w_cf) = " + pad(17) + _Ff(h["Qsf_cf"], 10, 3) + pad(10) + _Ff(h["w_cf"], 10, 3),
        "Water-surface width (Wsf_cf,b1_cf,b2_cf) = " + pad(5) + three("Wsf_cf", "b1_cf", "b2_cf"),
        "Resistance coefficient (kSsf_cf,y1_cf,y2_cf) = " + pad(1) + three("kSsf_cf", "y1_cf", "y2_cf"),
        "Initial flow discharge (Qsi_cf) = " + pad(14) + _Ff(h["Qsi_cf"], 10, 3),
        bar,
    ]
    return "\n".join(out) + "\n"


def read_dtm13(text: str, N: int, M: int) -> np.ndarray:
    """dtm_13.val (PRE/wbb_sr.f90:66-88): M list-directed records of N values, north row first -> array [M][N]."""
    if "D" not in text and "d" not in text and "," not in text:
        # the common case, every record on its own line with exactly N values: one strtod pass over the text
        try:
            flat = np.fromstring(text, dtype=np.float64, sep=" ")
        except (ValueError, DeprecationWarning):
            flat = np.empty(0)
        if flat.size == N * M and text.count("\n") in (M, M - 1):
            return flat.reshape(M, N)
    rows = np.empty((M, N))
    it = iter(text.splitlines())
    for r in range(M):
        got: list[float] = []
        while len(got) < N:
            try:
                ln = next(it)
            except StopIteration:
                raise PreproError("insufficient data in the file dtm_13.val") from None
            got.extend(_num(t) for t in ln.replace(",", " ").split())
        rows[r] = got[:N]                                # what is left of the record is skipped, as a list-directed READ does
    return rows


# ----------------------------------------------------------------------------------------------- raster text (RBB)
def _e_column(v: np.ndarray, w: int, d: int) -> np.ndarray:
    """Vectorised Fortran Ew.d of a float64 vector -> array of strings."""
    v = np.asarray(v, dtype=np.float64)
    a = np.abs(v)
    if np.any((a != 0) & ((a < 1e-98) | (a >= 1e98))) or not np.all(np.isfinite(v)):
        return np.array([_E(float(x), w, d) for x in v])
    s = np.char.mod("%%.%dE" % (d - 1), a).astype("S%d" % (d + 5))         # d.ddd..E+ee
    c = s.view(np.uint8).reshape(len(v), d + 5).astype(np.int32)
    ex = (c[:, d + 3] - 48) * 10 + (c[:, d + 4] - 48)
    ex = np.where(c[:, d + 2] == ord("-"), -ex, ex) + 1
    ex = np.where(a == 0, 0, ex)
    out = np.full((len(v), w), ord(" "), dtype=np.uint8)
    body = d + 6                                                            # 0.<d digits>E+ee
    o = w - body
    out[:, o] = ord("0")
    out[:, o + 1] = ord(".")
    out[:, o + 2] = c[:, 0]
    out[:, o + 3:o + 2 + d] = c[:, 2:d + 1]
    out[:, o + 2 + d] = ord("E")
    out[:, o + 3 + d] = np.where(ex < 0, ord("-"), ord("+"))
    out[:, o + 4 + d] = 48 + np.abs(ex) // 10
    out[:, o + 5 + d] = 48 + np.abs(ex) % 10
    neg = v < 0
    if np.any(neg):
        if o < 1:
            return np.array([_E(float(x), w, d) for x in v])
        out[neg, o - 1] = ord("-")
    return out.view("S%d" % w).reshape(len(v)).astype(str)


def _block_real(v2d: np.ndarray, w: int, d: int, kind: int) -> bytes:
    """[M][N] doubles -> M records of N fields Ew.d (kind 0) / Fw.d (kind 1): native row-parallel formatter of the library,
    numpy when a value needs a form it does not write."""
    lib = load_prepro_library()[0]
    v = np.ascontiguousarray(v2d, dtype=np.float64)
    M, N = v.shape
    buf = np.empty(M * (N * w + 1), dtype=np.uint8)
    if lib.cathy_prepro_format_real(v.ctypes.data_as(_D), M, N, w, d, kind, buf.ctypes.data_as(C.c_char_p), 0) == buf.size:
        return buf.tobytes()
    if kind == 1:
        txt = np.char.mod("%%%d.%df" % (w, d), v)
    else:
        txt = _e_column(v.ravel(), w, d).reshape(M, N)
    return ("\n".join("".join(row) for row in txt) + "\n").encode("ascii")


def _block_int(v2d: np.ndarray, w: int) -> bytes:
    lib = load_prepro_library()[0]
    v = np.ascontiguousarray(v2d, dtype=np.int32)
    M, N = v.shape
    buf = np.empty(M * (N * w + 1), dtype=np.uint8)
    lib.cathy_prepro_format_int(v.ctypes.data_as(_I), M, N, w, buf.ctypes.data_as(C.c_char_p), 0)
    return buf.tobytes()


def _header(hap: dict, ht: int) -> str:
    N, M = hap["N"], hap["M"]
    if ht == 2:                                                             # GRASS ascii header (mrbb_sr.f90:436-443)
        return ("north: " + _Ii(0, 5) + "\nsouth: " + _Ff(hap["yllcorner"], 20, 8) + "\neast:  " + _Ii(0, 5) + "\nwest:  "
                + _Ff(hap["xllcorner"], 20, 8) + "\nrows:  " + _Ii(M, 5) + "\ncols:  " + _Ii(N, 5) + "\n")
    if ht == 1:                                                             # ESRI ascii header (:426-433)
        return ("ncols" + " " * 8 + _Ii(N, 5) + "\nnrow" + " " * 9 + _Ii(M, 5) + "\nxllcorner " + _Ff(hap["xllcorner"], 20, 8)
                + "\nyllcorner " + _Ff(hap["yllcorner"], 20, 8) + "\ncellsize" + " " * 8 + _Ff(hap["delta_x"], 6, 2)
                + "\nNODATA_value" + " " * 4 + _Ii(-9999, 5) + "\n")
    return ""


# file -> (kind, result field) in MRBB_SR's order (PRE/mrbb_sr.f90:76-230)
RASTERS = (
    ("dem", "r", "quota"), ("lakes_map", "i", "lakes_map"), ("zone", "i", "zone"), ("dtm_w_1", "r", "w_1"), ("dtm_w_2", "r", "w_2"),
    ("dtm_p_outflow_1", "i", "p_outflow_1"), ("dtm_p_outflow_2", "i", "p_outflow_2"), ("dtm_A_inflow", "a", "A_inflow"),
    ("dtm_local_slope_1", "r", "local_slope_1"), ("dtm_local_slope_2", "r", "local_slope_2"), ("dtm_epl_1", "r", "epl_1"),
    ("dtm_epl_2", "r", "epl_2"), ("dtm_kSs1_sf_1", "r", "kSs1_sf_1"), ("dtm_kSs1_sf_2", "r", "kSs1_sf_2"),
    ("dtm_Ws1_sf_1", "r", "Ws1_sf_1"), ("dtm_Ws1_sf_2", "r", "Ws1_sf_2"), ("dtm_b1_sf", "r", "b1_sf"), ("dtm_y1_sf", "r", "y1_sf"),
    ("dtm_hcID", "i", "hcID"), ("dtm_q_output", "i", "q_output"), ("dtm_nrc", "r", "nrc"),
)
_ARCGIS = np.array([0, 8, 16, 32, 4, 0, 64, 2, 1, 128])                     # Jenson-Domingue codes (mrbb_sr.f90:268-279)


class PreproResult:
    """Cell records after HG, as arrays indexed [i_basin - 1], plus the cell order."""

    def __init__(self, hap: dict, present: np.ndarray, fields: dict, info: dict):
        self.hap, self.present, self.info = hap, present, info
        self.__dict__.update(fields)
        n = len(present)
        self.lakes_map = np.zeros(n, dtype=np.int32)                       # cella_iniziale (PRE/mbbio.f90:163-203)
        self.zone = np.ones(n, dtype=np.int32)
        self.q_output = np.zeros(n, dtype=np.int32)

    def north_first(self, field: str) -> np.ndarray:
        """[M][N] view of a field the way the files list it: north row first, west to east."""
        N, M = self.hap["N"], self.hap["M"]
        return getattr(self, field).reshape(N, M).T[::-1]

    def raster_text(self, name: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> str:
        return self.raster_bytes(name, ht, nodata, ips).decode("ascii")

    def raster_bytes(self, name: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> bytes:
        """RBB (PRE/mrbb_sr.f90:243-470)."""
        kind, field = next((k, f) for n, k, f in RASTERS if n == name)
        N, M = self.hap["N"], self.hap["M"]
        pres = self.present.reshape(N, M).T[::-1]
        g = self.north_first(field)
        nodata32 = float(np.float32(nodata))
        if kind == "i":
            vi = g.astype(np.int64)
            if ips == 2 and name.startswith("dtm_p_outflow"):
                vi = _ARCGIS[vi]
            elif ips not in (1, 2) and name.startswith("dtm_p_outflow"):
                raise PreproError("unexpected case")
            imax = max(int(vi[pres].max()), abs(int(nodata32)))
            imin = min(int(vi[pres].min()), int(nodata32))
            w = 2 if imax == 0 else int(math.log10(float(np.float32(imax)))) + (3 if imin < 0 else 2)
            return _header(self.hap, ht).encode("ascii") + _block_int(np.where(pres, vi, int(nodata32)), w)
        vr = np.where(pres, g.astype(np.float64), nodata32)
        neg = min(float(vr[pres].min()), nodata32) < 0.0
        if kind == "a":
            return _header(self.hap, ht).encode("ascii") + _block_real(vr, 15 if neg else 14, 2, 1)
        return _header(self.hap, ht).encode("ascii") + _block_real(vr, 20 if neg else 21, 12, 0)

    def qoi_a_text(self) -> str:
        """hg.f90:31-37: N_celle then the cells in descending elevation, list-directed INTEGER*4 (width 12)."""
        v = np.concatenate(([self.info["n_cells"]], self.order[:self.info["n_cells"]])).astype(np.int32)
        return _block_int(v.reshape(-1, 1), 12).decode("ascii")

    def write(self, directory: str, ht: int = 2, nodata: float = 0.0, ips: int = 1) -> None:
        with open(os.path.join(directory, "hap.in"), "w") as fh:
            fh.write(self.info["hap_text"])
        def one(name):
            with open(os.path.join(directory, name), "wb") as fh:
                fh.write(self.raster_bytes(name, ht, nodata, ips))
        # 21 independent files: numpy, the native formatter and the file writes all release the GIL
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(one, [name for name, _, _ in RASTERS]))
        with open(os.path.join(directory, "qoi_a"), "w") as fh:
            fh.write(self.qoi_a_text())


def terrain_analysis(hap_text: str, dtm13_text: str, device: int = 0) -> PreproResult:
    """CPPP without its file output: hap.in text + dtm_13.val text -> cell records (device computation)."""
    _, run, err = load_prepro_library()
    hap0 = read_hapin(hap_text)
    N, M = hap0["N"], hap0["M"]
    rows = read_dtm13(dtm13_text, N, M)                                    # [M][N], north first
    by_cell = rows[::-1].T.reshape(-1)                                     # [(i-1)*M + (j-1)]
    present = np.ascontiguousarray(by_cell > -9999.0)                      # wbb_sr.f90:45,76
    hap0["N_celle"] = int(present.sum())
    hap_text_out = write_hapin(hap0)
    hap = read_hapin(hap_text_out)                                         # CCA and everything after it see these (cca.f90:31)
    P = CathyPreproParams(abi_version=ABI_VERSION, N=N, M=M, imethod=hap["imethod"], ndcf=hap["ndcf"], nchc=hap["nchc"],
                          p_outflow_vo=hap["p_outflow_vo"], bcc=hap0["bcc"], delta_x0=hap0["delta_x"], pt0=hap0["pt"],
                          cqm0=hap0["cqm"], cqg0=hap0["cqg"], delta_x=hap["delta_x"], delta_y=hap["delta_y"], lambda_=hap["lambda_"],
                          A_threshold=hap["A_threshold"], CC_threshold=hap["CC_threshold"], ASk_threshold=hap["ASk_threshold"],
                          kas=hap["kas"], dr=hap["dr"], As_rf=hap["As_rf"], As_cf=hap["As_cf"])
    for sfx in ("_rf", "_cf"):
        for k in ("Qsf", "w", "Wsf", "b1", "b2", "kSsf", "y1", "y2"):
            setattr(P, k + sfx, hap[k + sfx])
    n = N * M
    fields = {name: np.zeros(n, dtype=dt) for name, dt in _OUT_ARRAYS}
    out = CathyPreproOut()
    for name, dt in _OUT_ARRAYS:
        setattr(out, name, fields[name].ctypes.data_as(_CT[dt]))
    q_in = np.ascontiguousarray(by_cell, dtype=np.float64)
    pres8 = present.astype(np.uint8)
    rc = run(C.byref(P), q_in.ctypes.data_as(_D), pres8.ctypes.data_as(C.POINTER(C.c_uint8)), device, C.byref(out))
    if rc != 0:
        raise PreproError(err().decode() or f"cathy_prepro_run failed ({rc})")
    info = {"n_cells": out.n_cells, "n_modifications": out.n_modifications, "n_waves": out.n_waves, "n_launches": out.n_launches,
            "mean_s_max": out.mean_s_max, "device_ms": out.device_ms, "hap_text": hap_text_out,
            "stage_ms": dict(zip(("csort", "pitcheck", "depit", "csort2", "local_smean", "dsf_sweep", "outlet_hg"), list(out.stage_ms)[:7])),
            "depit_sweeps": int(out.stage_ms[7])}
    return PreproResult(hap, present, fields, info)


def run_preprocessor(prepro_dir: str, ht: int = 2, nodata: float = 0.0, ips: int = 1, device: int = 0, log=None) -> PreproResult:
    """What `./pycppp` does in <project>/prepro."""
    log = log or (lambda s: None)
    with open(os.path.join(prepro_dir, "hap.in")) as fh:
        hap_text = fh.read()
    with open(os.path.join(prepro_dir, "dtm_13.val")) as fh:
        dtm_text = fh.read()
    log(" wbb... csort... depit... cca... smean... dsf... hg... (B200)\n")
    res = terrain_analysis(hap_text, dtm_text, device)
    log(" number of processed cells = %d\n dem modifications = %d (total)\n" % (res.info["n_cells"], res.info["n_modifications"]))
    res.write(prepro_dir, ht, nodata, ips)
    log(" ...mrbb completed\n")
    return res


def main(argv=None) -> int:
    """Entry of the `pycppp` launcher: cwd (or argv[0]) = <project>/prepro, the three answers of MRBB_SR on stdin
    (header type, nodata value, pointer system; PRE/mrbb_sr.f90:27-72); the reference's fatal messages go to stdout,
    where pyCATHY looks for them (PY/cathy_tools.py:393-398)."""
    argv = list(sys.argv[1:] if argv is None else argv)
    d = argv[0] if argv else os.getcwd()
    ans = sys.stdin.read().split() if not sys.stdin.isatty() else []
    try:
        ht = int(ans[0]) if len(ans) > 0 else 2
        nodata = float(ans[1]) if len(ans) > 1 else 0.0
        ips = int(ans[2]) if len(ans) > 2 else 1
        run_preprocessor(d, ht, nodata, ips, device=int(os.environ.get("CATHY_B200_DEVICE", "0")), log=sys.stdout.write)
    except (PreproError, capi.CathyLibraryError, OSError, ValueError) as e:
        sys.stdout.write(" %s\n" % e)
        return 1
    return 0
