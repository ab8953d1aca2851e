"""ctypes mirror of include/cathy_b200.h and the loader of libcathy_b200.so.

The product path has NO fallback: if the CUDA library is missing or fails to load,
``load_library()`` raises ``CathyLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .project import CathyProject

ABI_VERSION = 8
MAXIT = 64
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)


class CathyLibraryError(RuntimeError):
    pass


class CathyProblem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nrow", C.c_int32), ("ncol", C.c_int32), ("nstr", C.c_int32), ("nzone", C.c_int32), ("nveg", C.c_int32),
        ("ivert", C.c_int32), ("_pad0", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("west", C.c_double), ("south", C.c_double),
        ("factor", C.c_double), ("base", C.c_double),
        ("dem", _D), ("zone", _I), ("root_map", _D), ("zratio", _D),
        ("permx", _D), ("permy", _D), ("permz", _D), ("elstor", _D), ("poros", _D), ("vgn", _D),
        ("vgrmc", _D), ("vgpsat", _D),
        ("pcana", _D), ("pcref", _D), ("pcwlt", _D), ("zroot", _D), ("pz", _D), ("omgc", _D),
        ("pmin", C.c_double), ("scf", C.c_double),
        ("ivghu", C.c_int32),
        ("isimgr", C.c_int32), ("kslope", C.c_int32), ("lump", C.c_int32), ("iopt", C.c_int32),
        ("nlrelx", C.c_int32), ("l2norm", C.c_int32),
        ("ituns", C.c_int32), ("ituns1", C.c_int32), ("ituns2", C.c_int32), ("isolv", C.c_int32),
        ("itmxcg", C.c_int32),
        ("pondh_min", C.c_double), ("tolksl", C.c_double), ("tetaf", C.c_double), ("omega", C.c_double),
        ("toluns", C.c_double), ("tolswi", C.c_double), ("ernlmx", C.c_double), ("tolcg", C.c_double),
        ("deltat", C.c_double), ("dtmin", C.c_double), ("dtmax", C.c_double), ("tmax", C.c_double),
        ("dtmaga", C.c_double), ("dtmagm", C.c_double), ("dtreds", C.c_double), ("dtredm", C.c_double),
        ("indp", C.c_int32), ("ipond", C.c_int32),
        ("wtposition", C.c_double),
        ("ic_psi", _D), ("ic_pond", _D),
        ("atm_none", C.c_int32), ("hspatm", C.c_int32), ("ieto", C.c_int32), ("natm", C.c_int32),
        ("atm_time", _D), ("atm_val", _D),
        ("ndir_rec", C.c_int32), ("nneu_rec", C.c_int32),
        ("dir_time", _D), ("dir_ptr", _I), ("dir_node", _I), ("dir_val", _D),
        ("neu_time", _D), ("neu_ptr", _I), ("neu_node", _I), ("neu_val", _D), ("neu_n2d", _I),
        ("qoi", _I),
        ("dtm_w_1", _D), ("dtm_w_2", _D), ("dtm_p_outflow_1", _D), ("dtm_p_outflow_2", _D),
        ("dtm_local_slope_1", _D), ("dtm_local_slope_2", _D), ("dtm_epl_1", _D), ("dtm_epl_2", _D),
        ("dtm_kss1_sf_1", _D), ("dtm_kss1_sf_2", _D), ("dtm_ws1_sf_1", _D), ("dtm_ws1_sf_2", _D),
        ("dtm_b1_sf", _D), ("dtm_y1_sf", _D), ("dtm_nrc", _D),
        ("precond", C.c_int32), ("device", C.c_int32),
        ("tolcg_scale", C.c_double),
        ("dd_world", C.c_int32), ("dd_rank", C.c_int32), ("dd_row0", C.c_int32), ("dd_row1", C.c_int32),
        ("hualfa", C.c_double), ("hubeta", C.c_double), ("hugama", C.c_double), ("hupsia", C.c_double), ("huswr", C.c_double),
        ("hun", C.c_double), ("hua", C.c_double), ("hub", C.c_double), ("bcbeta", C.c_double), ("bcrmc", C.c_double), ("bcpsat", C.c_double),
        ("itmxcg_scale", C.c_double),
        ("nsf", C.c_int32), ("isfone", C.c_int32), ("isfcvg", C.c_int32), ("dupuit", C.c_int32),
        ("sf_ptr", _I), ("sf_node", _I),
        ("psel", C.c_double), ("pser", C.c_double),
    ]


class CathyIterRecord(C.Structure):
    _fields_ = [("niter", C.c_int32), ("ikmax", C.c_int32), ("pl2", C.c_double), ("pinf", C.c_double),
                ("pnew_ik", C.c_double), ("pold_ik", C.c_double), ("fl2", C.c_double), ("finf", C.c_double)]


class CathyStepReport(C.Structure):
    _fields_ = [
        ("nstep", C.c_int32), ("iter", C.c_int32), ("nitert", C.c_int32), ("kbackt", C.c_int32),
        ("nsurf", C.c_int32), ("nsurft", C.c_int32), ("noback", C.c_int32), ("finished", C.c_int32),
        ("n_iter_rec", C.c_int32), ("ponding", C.c_int32), ("klsfai_total", C.c_int32), ("kback_total", C.c_int32),
        ("deltat", C.c_double), ("time", C.c_double),
        ("store1", C.c_double), ("store2", C.c_double), ("dstore", C.c_double),
        ("vin", C.c_double), ("vout", C.c_double), ("erras", C.c_double), ("errel", C.c_double),
        ("adin", C.c_double), ("adout", C.c_double), ("ndin", C.c_double), ("ndout", C.c_double),
        ("anin", C.c_double), ("anout", C.c_double), ("nnin", C.c_double), ("nnout", C.c_double), ("sfflw", C.c_double),
        ("vsfflw", C.c_double), ("vndin", C.c_double), ("vndout", C.c_double), ("vnnin", C.c_double), ("vnnout", C.c_double),
        ("apot", C.c_double), ("aact", C.c_double), ("ovflow", C.c_double), ("reflow", C.c_double),
        ("fhort", C.c_double), ("fdunn", C.c_double), ("fpond", C.c_double), ("fsat", C.c_double),
        ("next_deltat", C.c_double), ("next_time", C.c_double),
        ("ak_max", C.c_double), ("q_outlet_1", C.c_double), ("q_outlet_2", C.c_double), ("gpu_ms", C.c_double),
        ("launches", C.c_int64),
        ("pcg_ms", C.c_double), ("pcg_iters", C.c_int64), ("pcg_solves", C.c_int64),
        ("aact_prev", C.c_double), ("areatot", C.c_double),
        ("itrtot", C.c_int32), ("hgflag", C.c_int32 * 9),
        ("it", CathyIterRecord * MAXIT),
    ]


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None and a.size else C.cast(None, _D)


def _ip(a):
    return a.ctypes.data_as(_I) if a is not None and a.size else C.cast(None, _I)


class ProblemHolder:
    """Owns the numpy buffers a CathyProblem points to (they must outlive the create call)."""

    def __init__(self, prj: CathyProject, precond: int = 0, device: int = 0, tolcg_scale: float = 0.0,
                 dd: tuple | None = None, itmxcg_scale: float = 0.0, **overrides):
        p = dict(prj.parm)
        p.update({k.upper(): v for k, v in overrides.items()})
        self.parm = p
        self.keep: list = []
        s = CathyProblem()
        s.abi_version = ABI_VERSION
        s.nrow, s.ncol, s.nstr, s.nzone = prj.nrow, prj.ncol, prj.nstr, prj.nzone
        veg = prj.soil["VEG"]
        s.nveg = veg.shape[0]
        s.ivert = prj.ivert
        s.dx, s.dy, s.west, s.south, s.factor, s.base = prj.dx, prj.dy, prj.west, prj.south, prj.factor, prj.base

        def fd(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self.keep.append(a)
            return _dp(a)

        def fi(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            self.keep.append(a)
            return _ip(a)

        s.dem, s.zone, s.root_map, s.zratio = fd(prj.dem), fi(prj.zone), fd(prj.root_map), fd(prj.zratio)
        tab = prj.soil["TABLE"]
        for k, name in enumerate(["permx", "permy", "permz", "elstor", "poros", "vgn", "vgrmc", "vgpsat"]):
            setattr(s, name, fd(tab[:, :, k]))
        for k, name in enumerate(["pcana", "pcref", "pcwlt", "zroot", "pz", "omgc"]):
            setattr(s, name, fd(veg[:, k]))
        s.pmin, s.scf, s.ivghu = prj.soil["PMIN"], prj.soil["SCF"], prj.soil["IVGHU"]
        for name in ["isimgr", "kslope", "lump", "iopt", "nlrelx", "l2norm", "ituns", "ituns1", "ituns2",
                     "isolv", "itmxcg"]:
            setattr(s, name, int(p[name.upper()]))
        for name in ["pondh_min", "tolksl", "tetaf", "omega", "toluns", "tolswi", "ernlmx", "tolcg", "deltat",
                     "dtmin", "dtmax", "tmax", "dtmaga", "dtmagm", "dtreds", "dtredm"]:
            setattr(s, name, float(p[name.upper()]))
        s.psel, s.pser = float(p.get("PSEL", 0.0)), float(p.get("PSER", 0.0))
        s.indp, s.ipond, s.wtposition = prj.indp, prj.ipond, prj.wtposition
        s.ic_psi, s.ic_pond = fd(prj.ic_psi), fd(prj.ic_pond)
        s.atm_none, s.hspatm, s.ieto = int(prj.atm_none), prj.hspatm, prj.ieto
        s.natm = len(prj.atm_times)
        s.atm_time, s.atm_val = fd(prj.atm_times), fd(prj.atm_values)

        def table(tab_):
            n = len(tab_.times)
            ptr = np.zeros(n + 1, dtype=np.int32)
            for i, nd in enumerate(tab_.nodes):
                ptr[i + 1] = ptr[i] + len(nd)
            nodes = np.concatenate(tab_.nodes) if n else np.zeros(0, dtype=np.int64)
            vals = np.concatenate(tab_.values) if n else np.zeros(0)
            return n, fd(np.asarray(tab_.times)), fi(ptr), fi(nodes), fd(vals), fi(np.asarray(tab_.n2d))

        s.ndir_rec, s.dir_time, s.dir_ptr, s.dir_node, s.dir_val, _ = table(prj.dirbc)
        s.nneu_rec, s.neu_time, s.neu_ptr, s.neu_node, s.neu_val, s.neu_n2d = table(prj.neubc)
        if prj.surf is not None and int(p["ISIMGR"]) == 2:
            S = prj.surf
            s.qoi = fi(S["qoi"])
            for cname, key in [("dtm_w_1", "w_1"), ("dtm_w_2", "w_2"), ("dtm_p_outflow_1", "p_outflow_1"),
                               ("dtm_p_outflow_2", "p_outflow_2"), ("dtm_local_slope_1", "local_slope_1"),
                               ("dtm_local_slope_2", "local_slope_2"), ("dtm_epl_1", "epl_1"), ("dtm_epl_2", "epl_2"),
                               ("dtm_kss1_sf_1", "kSs1_sf_1"), ("dtm_kss1_sf_2", "kSs1_sf_2"),
                               ("dtm_ws1_sf_1", "Ws1_sf_1"), ("dtm_ws1_sf_2", "Ws1_sf_2"), ("dtm_b1_sf", "b1_sf"),
                               ("dtm_y1_sf", "y1_sf"), ("dtm_nrc", "nrc")]:
                setattr(s, cname, fd(S[key]))
        elif int(p["ISIMGR"]) == 2:
            raise ValueError("ISIMGR=2 needs the prepro rasters")
        s.precond, s.device, s.tolcg_scale, s.itmxcg_scale = precond, device, tolcg_scale, itmxcg_scale
        # seepage faces (input/sfbc, SRC/sfvone.f): list of node-id arrays, one per face
        faces = getattr(prj, "seepage_faces", None) or []
        s.nsf = len(faces)
        s.isfone, s.isfcvg, s.dupuit = int(p.get("ISFONE", 0)), int(p.get("ISFCVG", 0)), int(p.get("DUPUIT", 0))
        if faces:
            ptr = np.zeros(len(faces) + 1, dtype=np.int32)
            ptr[1:] = np.cumsum([len(f) for f in faces])
            s.sf_ptr, s.sf_node = fi(ptr), fi(np.concatenate([np.asarray(f, dtype=np.int32) for f in faces]))
        if dd is not None:      # (world, rank, row0, row1): row-block partition, see partition_rows()
            s.dd_world, s.dd_rank, s.dd_row0, s.dd_row1 = (int(v) for v in dd)
        else:
            s.dd_world, s.dd_rank, s.dd_row0, s.dd_row1 = 1, 0, 0, prj.nrow + 1
        hu, huab, bc = prj.soil.get("HU", [0.0] * 5), prj.soil.get("HUAB", [0.0, 0.0]), prj.soil.get("BC", [0.0] * 3)
        s.hualfa, s.hubeta, s.hugama, s.hupsia, s.huswr = (float(v) for v in hu)
        s.hun, s.hua, s.hub = float(prj.soil.get("HUN", 0.0)), float(huab[0]), float(huab[1])
        s.bcbeta, s.bcrmc, s.bcpsat = (float(v) for v in bc)
        self.struct = s


class CathyLib:
    """Binds one shared library exporting the cathy_b200.h entry points under ``prefix``."""

    SYMBOLS = ["sizeof_problem", "sizeof_report", "last_error", "create", "destroy", "get_dims", "get_mesh",
               "initial_storage", "step", "attempt_log", "get_state", "get_velocity", "get_recharge", "get_wtdepth", "set_psi", "upload_atm_record", "debug_assemble", "debug_spmv", "debug_solve"]

    # entry points only the product library has (in-process ensemble support); bound when present
    PRODUCT_ONLY = ["pack_state", "unpack_psi", "restart", "set_soil", "set_atm_table", "dd_export", "dd_connect", "dd_connect_local", "dd_start", "dd_info", "solver_info", "solver_limits", "plan_info", "get_state_async", "state_wait"]

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise CathyLibraryError(f"{path} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        try:
            self.lib = C.CDLL(path)
        except OSError as e:  # missing libcudart etc.
            raise CathyLibraryError(f"cannot load {path}: {e}") from e
        self.path, self.prefix = path, prefix
        f = {}
        for name in self.SYMBOLS:
            try:
                f[name] = getattr(self.lib, prefix + name)
            except AttributeError as e:
                raise CathyLibraryError(f"{path} does not export {prefix}{name}") from e
        for name in self.PRODUCT_ONLY:
            if prefix == "cathy_":
                try:
                    f[name] = getattr(self.lib, prefix + name)
                except AttributeError as e:
                    raise CathyLibraryError(f"{path} does not export {prefix}{name}") from e
        self.f = f
        if "pack_state" in f:
            f["pack_state"].argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64]
            f["unpack_psi"].argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
            f["restart"].argtypes = [C.c_void_p, C.c_double, C.c_double]
            f["set_soil"].argtypes = [C.c_void_p] + [_D] * 8
            f["set_atm_table"].argtypes = [C.c_void_p, C.c_int32, _D, _D]
            f["dd_export"].argtypes = [C.c_void_p, C.c_void_p]
            f["dd_connect"].argtypes = [C.c_void_p, C.c_void_p]
            f["dd_info"].argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            f["solver_info"].argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            f["plan_info"].argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            f["solver_limits"].argtypes = [C.c_void_p, _D]
            f["dd_connect_local"].argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
            f["dd_start"].argtypes = [C.c_void_p]
            f["get_state_async"].argtypes = [C.c_void_p, _D, _D, _D, _D, _D, _D, _D, _D, _I]
            f["state_wait"].argtypes = [C.c_void_p]
        f["sizeof_problem"].restype = C.c_int64
        f["sizeof_report"].restype = C.c_int64
        f["last_error"].restype = C.c_char_p
        f["create"].argtypes = [C.POINTER(CathyProblem), C.POINTER(C.c_void_p)]
        f["destroy"].argtypes = [C.c_void_p]
        f["destroy"].restype = None
        f["get_dims"].argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        f["get_mesh"].argtypes = [C.c_void_p, _D, _D, _D, _I]
        f["initial_storage"].argtypes = [C.c_void_p]
        f["initial_storage"].restype = C.c_double
        f["step"].argtypes = [C.c_void_p, C.POINTER(CathyStepReport)]
        f["attempt_log"].argtypes = [C.c_void_p, C.c_int32, _I, _D, _D, C.POINTER(CathyIterRecord)]
        f["get_state"].argtypes = [C.c_void_p, _D, _D, _D, _D, _D, _D, _D, _D, _I]
        f["get_velocity"].argtypes = [C.c_void_p, _D, _D, _D, _D, _D, _D]
        f["get_recharge"].argtypes = [C.c_void_p, _D, _D]
        f["get_wtdepth"].argtypes = [C.c_void_p, _I, C.c_int32, _D]
        f["set_psi"].argtypes = [C.c_void_p, _D]
        f["upload_atm_record"].argtypes = [C.c_void_p, C.c_int32, _D]
        f["debug_assemble"].argtypes = [C.c_void_p, C.c_double, _I, _I, _D, _D]
        f["debug_spmv"].argtypes = [C.c_void_p, _D, _D, C.c_int32, _D]
        f["debug_solve"].argtypes = [C.c_void_p, _D, _I, _D, _D]
        if f["sizeof_problem"]() != C.sizeof(CathyProblem) or f["sizeof_report"]() != C.sizeof(CathyStepReport):
            raise CathyLibraryError(
                f"{path}: struct layout mismatch (problem {f['sizeof_problem']()} vs {C.sizeof(CathyProblem)}, "
                f"report {f['sizeof_report']()} vs {C.sizeof(CathyStepReport)})")

    def error(self) -> str:
        return (self.f["last_error"]() or b"").decode()


class Simulation:
    """One simulation handle (device resident for the product, host resident for the oracle)."""

    def __init__(self, lib: CathyLib, prj: CathyProject, **kw):
        self.lib, self.prj = lib, prj
        self.holder = ProblemHolder(prj, **kw)
        self.parm = self.holder.parm
        h = C.c_void_p()
        rc = lib.f["create"](C.byref(self.holder.struct), C.byref(h))
        if rc != 0:
            raise CathyLibraryError(f"{lib.prefix}create failed ({rc}): {lib.error()}")
        self.h = h
        dims = (C.c_int64 * 5)()
        lib.f["get_dims"](h, dims)
        self.nnod, self.n, self.nt, self.nterm, self.nnz = (int(v) for v in dims)

    def close(self):
        if self.h:
            self.lib.f["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initial_storage(self) -> float:
        return float(self.lib.f["initial_storage"](self.h))

    def mesh(self, with_tetra=True):
        x, y, z = np.empty(self.n), np.empty(self.n), np.empty(self.n)
        tet = np.empty((self.nt, 5), dtype=np.int32) if with_tetra else None
        self.lib.f["get_mesh"](self.h, _dp(x), _dp(y), _dp(z), _ip(tet) if with_tetra else C.cast(None, _I))
        return x, y, z, tet

    def step(self) -> CathyStepReport:
        rep = CathyStepReport()
        rc = self.lib.f["step"](self.h, C.byref(rep))
        if rc != 0:
            raise CathyLibraryError(f"{self.lib.prefix}step failed ({rc}): {self.lib.error()}")
        return rep

    def attempt_log(self) -> list:
        """Failed attempts of the last step, in the order they were made: [(deltat, time, [CathyIterRecord, ...]), ...] -- what the
        reference lists in output/iter before the accepted attempt (cathy_attempt_log)."""
        na = self.lib.f["attempt_log"](self.h, 0, None, None, None, None)
        if na <= 0:
            return []
        nrec = np.zeros(na, dtype=np.int32)
        dt, tm = np.zeros(na), np.zeros(na)
        rec = (CathyIterRecord * (MAXIT * na))()
        self.lib.f["attempt_log"](self.h, na, nrec.ctypes.data_as(_I), dt.ctypes.data_as(_D), tm.ctypes.data_as(_D), rec)
        return [(float(dt[a]), float(tm[a]), [rec[a * MAXIT + k] for k in range(int(nrec[a]))]) for a in range(na)]

    def state_buffers(self, pinned: bool = False) -> dict:
        """Host arrays for ``state(out=...)``; pinned (page-locked, via torch) buffers make the device-to-host copies DMA at
        full PCIe speed and are reusable from step to step."""
        n, nn = self.n, self.nnod
        if pinned:
            import torch

            # one page-locked block, carved in the order cathy_get_state_async stages the arrays: the read-back is a single copy
            block = torch.empty(4 * n + 4 * nn, dtype=torch.float64).pin_memory().numpy()
            out, off = {}, 0
            for k, m in (("psi", n), ("sw", n), ("ckrw", n), ("qtranie", n), ("pond", nn), ("atmact", nn), ("atmpot", nn), ("ovfl", nn)):
                out[k] = block[off:off + m]
                off += m
            out["ifatm"] = torch.empty(nn, dtype=torch.int32).pin_memory().numpy()
            return out
        out = {k: np.empty(n) for k in ("psi", "sw", "ckrw", "qtranie")}
        out.update({k: np.empty(nn) for k in ("pond", "atmact", "atmpot", "ovfl")})
        out["ifatm"] = np.empty(nn, dtype=np.int32)
        return out

    def state(self, out: dict | None = None) -> dict:
        if out is None:
            out = self.state_buffers()
        rc = self.lib.f["get_state"](self.h, _dp(out["psi"]), _dp(out["sw"]), _dp(out["ckrw"]), _dp(out["qtranie"]),
                                     _dp(out["pond"]), _dp(out["atmact"]), _dp(out["atmpot"]), _dp(out["ovfl"]),
                                     _ip(out["ifatm"]))
        if rc != 0:
            raise CathyLibraryError(f"get_state failed ({rc}): {self.lib.error()}")
        return out

    def state_async(self, out: dict) -> dict:
        """Pipelined ``state``: returns at once, ``out`` (page-locked buffers from ``state_buffers(pinned=True)``) is valid after
        ``state_wait()``; later ``step()`` calls overlap with the copies.  Alternate between two buffer sets."""
        cache = self.__dict__.setdefault("_async_ptrs", {})
        ptrs = cache.get(id(out))
        if ptrs is None or ptrs[0] is not out:      # the ctypes pointers of a buffer set are built once (this call sits between two steps)
            ptrs = (out, (_dp(out["psi"]), _dp(out["sw"]), _dp(out["ckrw"]), _dp(out["qtranie"]), _dp(out["pond"]), _dp(out["atmact"]),
                          _dp(out["atmpot"]), _dp(out["ovfl"]), _ip(out["ifatm"])))
            cache[id(out)] = ptrs
        self._ck(self.lib.f["get_state_async"](self.h, *ptrs[1]), "get_state_async")
        return out

    def state_wait(self) -> None:
        self._ck(self.lib.f["state_wait"](self.h), "state_wait")

    def velocity(self, nodal: bool = True) -> dict:
        """Darcy velocities at the current state: per element (VEL3D) and, optionally, per node (VNOD3D)."""
        out = {k: np.empty(self.nt) for k in ("uu", "vv", "ww")}
        if nodal:
            out.update({k: np.empty(self.n) for k in ("unod", "vnod", "wnod")})
        null = C.cast(None, _D)
        self._ck(self.lib.f["get_velocity"](self.h, _dp(out["uu"]), _dp(out["vv"]), _dp(out["ww"]),
                                            _dp(out["unod"]) if nodal else null, _dp(out["vnod"]) if nodal else null,
                                            _dp(out["wnod"]) if nodal else null), "get_velocity")
        return out

    def recharge(self):
        """(RECNOD[NNOD], RECFLOW) of SRC/recharge.f at the current state."""
        rec = np.empty(self.nnod)
        flow = C.c_double()
        self._ck(self.lib.f["get_recharge"](self.h, _dp(rec), C.cast(C.byref(flow), _D)), "get_recharge")
        return rec, flow.value

    def wtdepth(self, nodvp) -> np.ndarray:
        nd = np.ascontiguousarray(nodvp, dtype=np.int32)
        wt = np.empty(len(nd))
        self._ck(self.lib.f["get_wtdepth"](self.h, _ip(nd), len(nd), _dp(wt)), "get_wtdepth")
        return wt

    def set_psi(self, psi: np.ndarray):
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        assert psi.size == self.n
        rc = self.lib.f["set_psi"](self.h, _dp(psi))
        if rc != 0:
            raise CathyLibraryError(f"set_psi failed ({rc}): {self.lib.error()}")

    def upload_atm_record(self, rec: int, vals: np.ndarray):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        rc = self.lib.f["upload_atm_record"](self.h, rec, _dp(vals))
        if rc != 0:
            raise CathyLibraryError(f"upload_atm_record failed ({rc}): {self.lib.error()}")

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise CathyLibraryError(f"{what} failed ({rc}): {self.lib.error()}")

    # ---- in-process ensemble support (product library only) ----
    def pack_state(self, which: int, dptr: int, ld: int, col: int):
        """Copy psi (which=0) or sw (1) into column `col` of a device matrix [n][ld] at address `dptr`."""
        self._ck(self.lib.f["pack_state"](self.h, which, C.c_void_p(dptr), ld, col), "pack_state")

    def unpack_psi(self, dptr: int, ld: int, col: int):
        self._ck(self.lib.f["unpack_psi"](self.h, C.c_void_p(dptr), ld, col), "unpack_psi")

    def restart(self, tmax: float = 0.0, deltat: float = 0.0):
        self._ck(self.lib.f["restart"](self.h, tmax, deltat), "restart")

    def set_soil(self, table: np.ndarray):
        """table: [nstr][nzone][8] = PERMX PERMY PERMZ ELSTOR POROS VGN VGRMC VGPSAT (the rows of input/soil)."""
        cols = [np.ascontiguousarray(table[:, :, k], dtype=np.float64) for k in range(8)]
        self._ck(self.lib.f["set_soil"](self.h, *[_dp(c) for c in cols]), "set_soil")

    def set_atm_table(self, times: np.ndarray, vals: np.ndarray):
        times = np.ascontiguousarray(times, dtype=np.float64)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self._ck(self.lib.f["set_atm_table"](self.h, len(times), _dp(times), _dp(vals)), "set_atm_table")

    # ---- row-block partition ----
    def dd_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.f["dd_export"](self.h, buf), "dd_export")
        return buf.raw

    def dd_connect(self, handles: bytes):
        self._ck(self.lib.f["dd_connect"](self.h, C.c_char_p(handles)), "dd_connect")

    def dd_info(self) -> dict:
        v = (C.c_int64 * 8)()
        self._ck(self.lib.f["dd_info"](self.h, v), "dd_info")
        keys = ["win_row0", "win_rows", "own_row0", "own_row1", "nnod_local", "n_local", "nnod_global", "n_global"]
        return dict(zip(keys, (int(x) for x in v)))

    def solver_info(self) -> dict:
        """Linear-solver kernel of this handle: kernel id (1 k_pcg, 2 k_pcg2, 3 k_pcg_res, 4 k_pcg_res2, 10 k_bicgstab), rows per CTA, x resident, grid."""
        v = (C.c_int64 * 4)()
        self._ck(self.lib.f["solver_info"](self.h, v), "solver_info")
        return dict(zip(["kernel", "rows_per_cta", "x_resident", "grid"], (int(x) for x in v)))

    def solver_limits(self) -> dict | None:
        """Effective stopping rule of the linear solver: ITMXCG x itmxcg_scale iterations, TOLCG x tolcg_scale relative residual
        (None for the CPU oracle, which runs the reference's own rule)."""
        if "solver_limits" not in self.lib.f:
            return None
        v = (C.c_double * 5)()
        self._ck(self.lib.f["solver_limits"](self.h, v), "solver_limits")
        return {"itmax": int(v[0]), "tol": float(v[1]), "itmxcg_scale": float(v[2]), "tolcg_scale": float(v[3]),
                "preconditioner": {1: "diagonal", 2: "vertical line"}.get(int(v[4]), "?")}

    def plan_info(self) -> dict:
        """Assembly plan of this handle: analytic = tet indices derived from the mesh structure instead of stored."""
        v = (C.c_int64 * 2)()
        self._ck(self.lib.f["plan_info"](self.h, v), "plan_info")
        return {"analytic": bool(v[0]), "table_width": int(v[1])}

    def debug_assemble(self, deltat: float):
        # Picard: symmetric upper CSR (NTERM entries); Newton: the Jacobian in full CSR (nnz entries)
        nent = self.nnz if int(self.parm.get("IOPT", 1)) == 2 else self.nterm
        topol = np.empty(self.n + 1, dtype=np.int32)
        ja = np.empty(nent, dtype=np.int32)
        coef = np.empty(nent)
        rhs = np.empty(self.n)
        rc = self.lib.f["debug_assemble"](self.h, deltat, _ip(topol), _ip(ja), _dp(coef), _dp(rhs))
        if rc != 0:
            raise CathyLibraryError(f"debug_assemble failed ({rc}): {self.lib.error()}")
        return topol, ja, coef, rhs

    def debug_spmv(self, x: np.ndarray, reps: int = 1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n)
        ms = C.c_double()
        rc = self.lib.f["debug_spmv"](self.h, _dp(x), _dp(y), reps, C.byref(ms))
        if rc != 0:
            raise CathyLibraryError(f"debug_spmv failed ({rc}): {self.lib.error()}")
        return y, ms.value

    def debug_solve(self):
        sol = np.empty(self.n)
        nit, err, ms = C.c_int32(), C.c_double(), C.c_double()
        rc = self.lib.f["debug_solve"](self.h, _dp(sol), C.byref(nit), C.byref(err), C.byref(ms))
        if rc != 0:
            raise CathyLibraryError(f"debug_solve failed ({rc}): {self.lib.error()}")
        return sol, nit.value, err.value, ms.value


_LIB = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcathy_b200.so")


def load_library() -> CathyLib:
    """The product library.  Raises CathyLibraryError when it is missing -- there is no CPU path."""
    global _LIB
    if _LIB is None:
        _LIB = CathyLib(library_path(), "cathy_")
    return _LIB


def partition_rows(nrow: int, world: int) -> list[tuple[int, int]]:
    """Owned global node-row ranges [a, b) of the (nrow + 1) DEM node rows for `world` ranks: contiguous strips of DEM rows
    (SURVEY.md section 8e), as even as possible."""
    rows = nrow + 1
    base, rem = divmod(rows, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < rem else 0)
        out.append((a, b))
        a = b
    return out
