"""Ensemble data assimilation on B200: host-side mirror of pyCATHY's analysis functions and an in-process ensemble.

Reference interfaces mirrored here (same names, argument order and return slots):
  * ``enkf_analysis``                           pyCATHY/DA/enkf.py:16-224
  * ``enkf_analysis_localized_with_inflation``  pyCATHY/DA/enkf.py:225-342
  * ``particle_filter_analysis``                pyCATHY/DA/pf.py:3-195
  * ``run_analysis``                            pyCATHY/DA/cathy_DA.py:86-260 (dispatcher on DA_type)
  * ``Ensemble``                                the forecast/analysis cycle of DA.run_DA_sequential
                                                (pyCATHY/DA/cathy_DA.py:1373-1491 forecast, :2684 state read-out,
                                                :1863-1875 restart files) kept device resident

All arithmetic runs in libcathy_b200.so (csrc/cathy_enkf.cu: fp64 tensor-core kernels); there is no CPU path --
``load_enkf_library()`` raises when the library is missing.  Multi-GPU: members (columns) are sharded over ranks, the
row sums and the partial cross covariance are all-reduced with torch.distributed (NCCL on GPUs), the predicted
observations are all-gathered; torch is used for device buffers and the process group only.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from .capi import CathyLibraryError, library_path

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int32)


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None else C.cast(None, _D)


class EnkfLib:
    SYMBOLS = ["enkf_last_error", "enkf_analysis_host", "enkf_gain", "enkf_rowsum", "enkf_scale", "enkf_crosscov",
               "enkf_update", "enkf_localization", "pf_weights", "pf_systematic_resample", "pf_gather_members"]

    def __init__(self, path: str):
        import os
        if not os.path.exists(path):
            raise CathyLibraryError(f"{path} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`")
        try:
            self.lib = C.CDLL(path)
        except OSError as e:
            raise CathyLibraryError(f"cannot load {path}: {e}") from e
        f = {}
        for name in self.SYMBOLS:
            try:
                f[name] = getattr(self.lib, "cathy_" + name)
            except AttributeError as e:
                raise CathyLibraryError(f"{path} does not export cathy_{name}") from e
        V, I32, I64, U64, DBL = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        f["enkf_last_error"].restype = C.c_char_p
        f["enkf_analysis_host"].argtypes = [_D, I64, I32, _D, _D, I32, _D, I32, I32, _D, I64, DBL, I64, DBL, _D, _D, _D, I32, _D]
        f["enkf_gain"].argtypes = [_D, _D, I32, _D, I32, I32, I32, _D, _D]
        f["enkf_rowsum"].argtypes = [V, I64, I32, V, U64]
        f["enkf_scale"].argtypes = [V, I64, DBL, U64]
        f["enkf_crosscov"].argtypes = [V, V, V, I64, I32, I32, I32, V, U64]
        f["enkf_update"].argtypes = [V, V, V, I64, V, V, V, DBL, I64, DBL, I64, I32, I32, V, U64]
        f["enkf_localization"].argtypes = [V, I64, V, I32, DBL, V, U64]
        f["pf_weights"].argtypes = [_D, _D, _D, I32, I32, _D, _D]
        f["pf_systematic_resample"].argtypes = [_D, I32, DBL, _I]
        f["pf_gather_members"].argtypes = [V, I64, I32, V, V, U64]
        self.f = f

    def check(self, rc: int, what: str):
        if rc != 0:
            raise CathyLibraryError(f"{what} failed ({rc}): {(self.f['enkf_last_error']() or b'').decode()}")


_ENKF = None


def load_enkf_library() -> EnkfLib:
    global _ENKF
    if _ENKF is None:
        _ENKF = EnkfLib(library_path())
    return _ENKF


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _analysis_host(X, HX, y, R, sakov, L, n_loc, inflate, n_infl, inflate2, device=0, want_P=True):
    lib = load_enkf_library()
    X, HX, R = _f64(X), _f64(HX), _f64(R)
    y = _f64(y)
    n, ne = X.shape
    m = HX.shape[0]
    if HX.shape != (m, ne) or R.shape != (m, m):
        raise ValueError(f"shape mismatch: X {X.shape}, HX {HX.shape}, R {R.shape}")
    y_is_matrix = int(y.ndim == 2)
    if y_is_matrix and y.shape != (m, ne):
        raise ValueError(f"data matrix must be (m, ne) = {(m, ne)}, got {y.shape}")
    if not y_is_matrix and y.shape != (m,):
        raise ValueError(f"data must have {m} entries, got {y.shape}")
    Lc = _f64(L) if L is not None else None
    Xa = np.empty_like(X)
    B = np.empty((m, ne))
    P = np.empty((n, m)) if want_P else None
    ms = C.c_double()
    rc = lib.f["enkf_analysis_host"](_dp(X), n, ne, _dp(HX), _dp(y), y_is_matrix, _dp(R), m, int(bool(sakov)), _dp(Lc),
                                     n_loc if Lc is not None else 0, float(inflate), n_infl, float(inflate2), _dp(Xa), _dp(B),
                                     _dp(P), device, C.byref(ms))
    lib.check(rc, "cathy_enkf_analysis_host")
    return Xa, B, P, ms.value


def enkf_analysis(data, data_cov, param, ensemble, predict_obs, **kwargs):
    """Mirror of pyCATHY/DA/enkf.py:16-224.  Returns the same 11-slot list; slot 1 (the mean) is a column instead of
    the reference's tiled copy and slot 2 (perturbations) is formed lazily on the host only when asked via
    ``return_perturbations=True`` -- both are N x Ne debugging copies the DA loop never reads."""
    Sakov = kwargs.pop("Sakov", False)
    want_pert = kwargs.pop("return_perturbations", False)
    device = kwargs.pop("device", 0)
    if isinstance(predict_obs, list):          # enkf.py:113-117
        if len(predict_obs) == 1:
            predict_obs = np.array(predict_obs[0])
        else:
            raise ValueError("predict_obs should be numpy array")
    ensemble = _f64(ensemble)
    sim_size, ens_size = ensemble.shape
    augm_state = np.vstack([ensemble, _f64(param)]) if len(param) > 0 else ensemble
    data = _f64(data)
    predict_obs = _f64(predict_obs)
    data_cov = _f64(data_cov)
    # the device forms C = S S^T/(Ne-1) + R^T from the R it is given (enkf.py:166 uses data_cov.T)
    Xa, B, P, _ms = _analysis_host(augm_state, predict_obs, data, data_cov, Sakov, None, 0, 1.0, 0, 1.0, device=device)
    obs_avg = (1.0 / ens_size) * np.tile(predict_obs.sum(1), (ens_size, 1)).T
    obs_pert = predict_obs - obs_avg
    data_pert = (data.T - predict_obs.T).T
    COV = data_cov.T if Sakov else (1.0 / (ens_size - 1)) * (obs_pert @ obs_pert.T) + data_cov.T   # m x m, diagnostics only
    mean = augm_state.mean(axis=1, keepdims=True)
    pert = augm_state - mean if want_pert else None
    return [augm_state, mean, pert, data_pert, obs_avg, obs_pert, COV, B, P, Xa[:sim_size, :], Xa[sim_size:, :].T]


def enkf_analysis_localized_with_inflation(data, data_cov, ensemble, param, predict_obs, L=None, **kwargs):
    """Mirror of pyCATHY/DA/enkf.py:225-342 (note the reference's argument order: ensemble BEFORE param)."""
    Sakov = kwargs.pop("Sakov")
    inflate_states = kwargs.pop("inflate_states")
    inflate_params = kwargs.pop("inflate_params")
    jitter_params = kwargs.pop("jitter_params")
    device = kwargs.pop("device", 0)
    ensemble = _f64(ensemble)
    param = _f64(param)
    sim_size, ens_size = ensemble.shape
    augm_state = np.vstack([ensemble, param])
    data = _f64(data).reshape(-1)
    predict_obs = _f64(predict_obs)
    data_cov = _f64(data_cov)
    if L is not None and tuple(L.shape) != (sim_size, predict_obs.shape[0]):
        raise ValueError(f"Localization matrix shape {L.shape} does not match P_xo[:states, :] {(sim_size, predict_obs.shape[0])}")
    # this variant adds data_cov untransposed (enkf.py:300); the device transposes what it is given
    Xa, B, P, _ms = _analysis_host(augm_state, predict_obs, data, data_cov.T, Sakov, L, sim_size, inflate_states, sim_size,
                                   inflate_params, device=device)
    analysis, analysis_param = Xa[:sim_size, :], Xa[sim_size:, :]
    if jitter_params > 0.0:                     # enkf.py:326-328 (global numpy RNG, like the reference)
        analysis_param = analysis_param + np.random.normal(0, jitter_params, analysis_param.shape)
    obs_avg = predict_obs.mean(axis=1, keepdims=True)
    obs_pert = predict_obs - obs_avg
    data_pert = data.reshape(-1, 1) - predict_obs
    COV = data_cov if Sakov else (obs_pert @ obs_pert.T) / (ens_size - 1) + data_cov
    mean = augm_state.mean(axis=1, keepdims=True)
    return [augm_state, mean, None, data_pert, obs_avg, obs_pert, COV, B, ensemble - mean[:sim_size], analysis, analysis_param]


def particle_filter_analysis(data, data_cov, param, ensemble, observation, **kwargs):
    """Mirror of pyCATHY/DA/pf.py:3-195 (bootstrap filter with systematic resampling; the hybrid weighted-EnKF branch
    is not built).  ``u`` (optional) fixes the uniform draw of the resampler for reproducible runs."""
    import torch
    if kwargs.get("use_enkf_update", False):
        raise NotImplementedError("particle_filter_analysis(use_enkf_update=True) is not implemented on the device")
    resample_threshold = kwargs.get("resample_threshold", 0.5)
    jitter_std_param = kwargs.get("jitter_std_param", 0.01)
    jitter_std_state = kwargs.get("jitter_std_state", 0.005)
    use_log_K = kwargs.get("use_log_K", False)
    u = kwargs.get("u", None)
    lib = load_enkf_library()
    ensemble, param = _f64(ensemble), _f64(param)
    sim_size, ens_size = ensemble.shape
    par_size = param.shape[0]
    observation = _f64(observation)
    if observation.shape[0] == ens_size:          # pf.py:44-45
        observation = np.ascontiguousarray(observation.T)
    data = np.atleast_1d(_f64(data).flatten())
    meas_size = data.shape[0]
    data_cov = _f64(data_cov)
    obs_std = _f64(np.sqrt(np.diag(data_cov)) if data_cov.ndim == 2 else np.sqrt(data_cov))
    weights = np.empty(ens_size)
    n_eff = C.c_double()
    lib.check(lib.f["pf_weights"](_dp(observation), _dp(data), _dp(obs_std), meas_size, ens_size, _dp(weights), C.byref(n_eff)),
              "cathy_pf_weights")
    n_eff = n_eff.value
    resampled = False
    if n_eff < resample_threshold * ens_size:     # pf.py:97-112
        resampled = True
        if u is None:
            u = np.random.rand()
        idx = np.empty(ens_size, dtype=np.int32)
        lib.check(lib.f["pf_systematic_resample"](_dp(weights), ens_size, float(u), idx.ctypes.data_as(_I)), "cathy_pf_systematic_resample")
        if not torch.cuda.is_available():
            raise CathyLibraryError("particle_filter_analysis: no CUDA device (there is no CPU path)")
        aug = torch.from_numpy(np.vstack([ensemble, param, observation])).cuda()
        out = torch.empty_like(aug)
        d_idx = torch.from_numpy(idx).cuda()
        lib.check(lib.f["pf_gather_members"](aug.data_ptr(), aug.shape[0], ens_size, d_idx.data_ptr(), out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "cathy_pf_gather_members")
        out = out.cpu().numpy()
        ensemble, param, observation = out[:sim_size], out[sim_size:sim_size + par_size], out[sim_size + par_size:]
        weights = np.ones(ens_size) / ens_size
    if resampled or jitter_std_param > 0:         # pf.py:127-151
        if jitter_std_state > 0:
            ensemble = ensemble + np.random.randn(*ensemble.shape) * jitter_std_state
        if jitter_std_param > 0:
            param = np.array(param)
            if use_log_K:
                n_cells = par_size // 2
                param[:n_cells, :] += np.random.randn(n_cells, ens_size) * jitter_std_param
                param[n_cells:, :] *= np.exp(np.random.randn(n_cells, ens_size) * jitter_std_param)
            else:
                param += np.random.randn(*param.shape) * jitter_std_param
    return {"Analysis": ensemble, "Analysisparam": param, "weights": weights, "n_eff": n_eff, "resampled": resampled,
            "observation": observation}


def build_localization_matrix(obs_pos, grid_pos, L, device: int = 0, as_tensor: bool = False):
    """Mirror of pyCATHY/DA/localisation.py:163-188 (Gaspari-Cohn weights between grid and observation positions, 2-D):
    returns the (n_grid, n_obs) matrix -- a numpy array, or the device tensor when ``as_tensor`` (to feed
    ``sharded_enkf_update(..., L=...)`` without a host round trip)."""
    import torch
    if not torch.cuda.is_available():
        raise CathyLibraryError("build_localization_matrix: no CUDA device (there is no CPU path)")
    lib = load_enkf_library()
    dev = torch.device("cuda", device)
    g = torch.from_numpy(_f64(np.asarray(grid_pos)[:, :2])).to(dev)
    o = torch.from_numpy(_f64(np.asarray(obs_pos)[:, :2])).to(dev)
    out = torch.empty((g.shape[0], o.shape[0]), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        lib.check(lib.f["enkf_localization"](g.data_ptr(), g.shape[0], o.data_ptr(), o.shape[0], float(L), out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "cathy_enkf_localization")
        torch.cuda.current_stream().synchronize()
    return out if as_tensor else out.cpu().numpy()


def run_analysis(DA_type, data, data_cov, param, list_update_parm, ensembleX, prediction, default_state="psi", **kwargs):
    """Mirror of the dispatcher pyCATHY/DA/cathy_DA.py:86-260 for the analysis kinds built on the device."""
    id_state = 1 if default_state == "sw" else 0
    Sakov = kwargs.pop("Sakov", True)
    L = kwargs.pop("localisationMatrix", None)
    inflate_states = kwargs.pop("inflation", None) or 1.0
    inflate_params = kwargs.pop("inflate_params", None) or 1.0
    jitter_params = kwargs.pop("jitter_params", None) or 0.0
    if DA_type == "pf":
        return particle_filter_analysis(data, data_cov, param, ensembleX[id_state], prediction, use_enkf_update=False,
                                        resample_threshold=0.5, jitter_std_param=0.01, jitter_std_state=0.05)
    if DA_type == "enkf_Evensen2009":
        return enkf_analysis(data, data_cov, param, ensembleX[id_state], prediction)
    if DA_type == "enkf_Evensen2009_Sakov":
        return enkf_analysis(data, data_cov, param, ensembleX[id_state], prediction, Sakov=True)
    if DA_type == "enkf_analysis_localized_with_inflation":
        return enkf_analysis_localized_with_inflation(data, data_cov, ensembleX[id_state], param, prediction, L, Sakov=Sakov,
                                                      inflate_states=inflate_states, inflate_params=inflate_params,
                                                      jitter_params=jitter_params)
    raise NotImplementedError(f"DA_type {DA_type!r} is not implemented on the device")


# ------------------------------------------------------------------------------------------------
# member-sharded analysis on device-resident ensembles
# ------------------------------------------------------------------------------------------------
class GpuOps:
    """Local stages of the analysis on torch CUDA tensors through the C ABI (csrc/cathy_enkf.cu)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.lib = load_enkf_library()
        if not torch.cuda.is_available():
            raise CathyLibraryError("GpuOps: no CUDA device (there is no CPU path)")

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def gain(self, hx, y, R, sakov):
        hx, y, R = _f64(hx), _f64(y), _f64(R)
        m, ne = hx.shape
        S, B = np.empty((m, ne)), np.empty((m, ne))
        self.lib.check(self.lib.f["enkf_gain"](_dp(hx), _dp(y), int(y.ndim == 2), _dp(R), m, ne, int(bool(sakov)), _dp(S), _dp(B)),
                       "cathy_enkf_gain")
        return S, B

    def rowsum(self, X):
        out = self.torch.empty(X.shape[0], dtype=self.torch.float64, device=X.device)
        self.lib.check(self.lib.f["enkf_rowsum"](X.data_ptr(), X.shape[0], X.shape[1], out.data_ptr(), self._stream()), "cathy_enkf_rowsum")
        return out

    def crosscov(self, X, mean, S_local, ne_total):
        t = self.torch
        m = S_local.shape[0]
        dS = t.from_numpy(_f64(S_local)).to(X.device)
        P = t.empty((X.shape[0], m), dtype=t.float64, device=X.device)
        self.lib.check(self.lib.f["enkf_crosscov"](X.data_ptr(), mean.data_ptr(), dS.data_ptr(), X.shape[0], X.shape[1], m, ne_total,
                                                   P.data_ptr(), self._stream()), "cathy_enkf_crosscov")
        return P

    def update(self, X, P, L, B_local, mean, bbar, inflate, n_infl, inflate2):
        t = self.torch
        dB = B_local if isinstance(B_local, t.Tensor) else t.from_numpy(_f64(B_local)).to(X.device)
        dbb = bbar if isinstance(bbar, t.Tensor) else t.from_numpy(_f64(bbar)).to(X.device)
        n_loc = L.shape[0] if L is not None else 0
        self.lib.check(self.lib.f["enkf_update"](X.data_ptr(), P.data_ptr(), L.data_ptr() if L is not None else None, n_loc, dB.data_ptr(),
                                                 mean.data_ptr(), dbb.data_ptr(), float(inflate), n_infl, float(inflate2), X.shape[0],
                                                 X.shape[1], P.shape[1], X.data_ptr(), self._stream()), "cathy_enkf_update")
        t.cuda.current_stream().synchronize()   # dB / dbb are released on return
        return X


def sharded_enkf_update(X_local, HX_local, y, R, sakov=False, L=None, inflate=1.0, n_infl=0, inflate2=1.0, group=None, ops=None):
    """EnKF analysis of a member-sharded ensemble, IN PLACE on ``X_local`` [n][ne_local] (torch tensor on this rank's
    device).  HX_local [m][ne_local] are this rank's predicted observations; y [m] or [m][ne_total] and R [m][m] are
    replicated host arrays.  Collectives: all_gather(HX), all_reduce(row sums), all_reduce(partial cross covariance).
    Members are ordered rank-major in the gathered matrices.  Returns (X_local, info dict)."""
    import torch
    import torch.distributed as dist
    ops = ops or GpuOps()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    ne_local = X_local.shape[1]
    HX_local = HX_local if isinstance(HX_local, torch.Tensor) else torch.from_numpy(_f64(HX_local))
    HX_local = HX_local.to(device=X_local.device, dtype=torch.float64).contiguous()
    m = HX_local.shape[0]
    if multi:
        counts = torch.zeros(world, dtype=torch.int64, device=X_local.device)
        counts[rank] = ne_local
        dist.all_reduce(counts, group=group)
        counts = [int(c) for c in counts.tolist()]
        nmax = max(counts)
        pad = torch.zeros((m, nmax), dtype=torch.float64, device=X_local.device)
        pad[:, :ne_local] = HX_local
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        HX_all = torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1)
    else:
        counts = [ne_local]
        HX_all = HX_local
    ne_total = sum(counts)
    c0 = sum(counts[:rank])
    S_all, B_all = ops.gain(HX_all.cpu().numpy(), y, R, sakov)
    bbar = B_all.mean(axis=1)
    # CUDA events around the two dense kernels and the collectives (torch's current stream is the one they are launched on)
    cuda = X_local.is_cuda

    class _NoEvent:
        def record(self):
            pass

    ev = [torch.cuda.Event(enable_timing=True) if cuda else _NoEvent() for _ in range(6)]
    ev[0].record()
    rs = ops.rowsum(X_local)
    if multi:
        dist.all_reduce(rs, group=group)
    mean = rs * (1.0 / ne_total)
    ev[1].record()
    P = ops.crosscov(X_local, mean, S_all[:, c0:c0 + ne_local], ne_total)
    ev[2].record()
    if multi:
        dist.all_reduce(P, group=group)
    ev[3].record()
    B_loc, bb = B_all[:, c0:c0 + ne_local], bbar
    if cuda:      # staged before the timed kernel
        B_loc, bb = torch.from_numpy(_f64(B_loc)).to(X_local.device), torch.from_numpy(_f64(bbar)).to(X_local.device)
    ev[4].record()
    ops.update(X_local, P, L, B_loc, mean, bb, inflate, n_infl, inflate2)
    ev[5].record()
    timing = {}
    if cuda:
        torch.cuda.current_stream().synchronize()
        timing = {"rowsum_mean_ms": ev[0].elapsed_time(ev[1]), "crosscov_ms": ev[1].elapsed_time(ev[2]), "collective_ms": ev[2].elapsed_time(ev[3]),
                  "update_ms": ev[4].elapsed_time(ev[5])}
    return X_local, {"ne_total": ne_total, "col0": c0, "B": B_all, "S": S_all, "P": P, "mean": mean, "timing_ms": timing}


# ------------------------------------------------------------------------------------------------
# in-process ensemble: forecast members on this rank's GPU, analysis across ranks
# ------------------------------------------------------------------------------------------------
class Ensemble:
    """``members``: list of (CathyProject, soil_table) for THIS rank.  Each member is one simulation handle on ``device``;
    states stay in HBM between assimilation windows (the reference writes input/ic and re-reads output/psi as text)."""

    def __init__(self, lib, projects, device: int = 0, group=None, concurrent: int = 1):
        """``concurrent`` > 1: that many members advance at the same time (one host thread each); their persistent solver
        kernels then use #SMs // concurrent CTAs each so that all of them stay co-resident (CATHY_PCG_GRID)."""
        import os
        import torch
        from .capi import Simulation
        self.torch = torch
        self.device, self.group = device, group
        self.concurrent = max(1, int(concurrent))
        if self.concurrent > 1:
            sms = torch.cuda.get_device_properties(device).multi_processor_count
            old = os.environ.get("CATHY_PCG_GRID")
            os.environ["CATHY_PCG_GRID"] = str(max(1, sms // self.concurrent))
        try:
            # cathy_create is mostly single-threaded host work (mesh, static gather plan) and ctypes releases the GIL inside it:
            # the members of a rank are built by a few host threads side by side
            nthr = max(1, min(len(projects), (os.cpu_count() or 1), 8))
            if nthr > 1:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(max_workers=nthr) as ex:
                    self.sims = list(ex.map(lambda prj: Simulation(lib, prj, device=device), projects))
            else:
                self.sims = [Simulation(lib, prj, device=device) for prj in projects]
        finally:
            if self.concurrent > 1:
                if old is None:
                    os.environ.pop("CATHY_PCG_GRID", None)
                else:
                    os.environ["CATHY_PCG_GRID"] = old
        self.n = self.sims[0].n if self.sims else 0
        self.ne_local = len(self.sims)
        dev = torch.device("cuda", device)
        self.X = torch.empty((self.n, self.ne_local), dtype=torch.float64, device=dev)    # psi, member minor
        self.SW = torch.empty((self.n, self.ne_local), dtype=torch.float64, device=dev)
        self.steps = 0

    def forecast(self) -> int:
        """Advance every local member to the end of its current window (TMAX); returns accepted steps summed over members.
        Members whose run stopped before TMAX (no convergence at DTMIN) are listed in ``self.failed`` -- pyCATHY's
        ``rejected_ens`` (pyCATHY/DA/cathy_DA.py:1450-1491 detects them from a short mbeconv)."""
        self.failed = []

        def run(j):
            s, k = self.sims[j], 0
            while True:
                rep = s.step()
                k += 1
                if rep.finished:
                    return k, bool(rep.noback)

        if self.concurrent > 1 and len(self.sims) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=self.concurrent) as ex:      # ctypes releases the GIL inside cathy_step
                res = list(ex.map(run, range(len(self.sims))))
        else:
            res = [run(j) for j in range(len(self.sims))]
        steps = sum(r[0] for r in res)
        self.failed = [j for j, r in enumerate(res) if r[1]]
        self.steps += steps
        return steps

    def gather_states(self):
        for j, s in enumerate(self.sims):
            s.pack_state(0, self.X.data_ptr(), self.ne_local, j)
            s.pack_state(1, self.SW.data_ptr(), self.ne_local, j)
        return self.X, self.SW

    def predict_obs(self, obs_nodes, porosity):
        """Predicted soil-water-content observations of the local members, [m][ne_local] on the device."""
        t = self.torch
        self.gather_states()
        idx = t.as_tensor(np.asarray(obs_nodes, dtype=np.int64), device=self.X.device)
        return self.SW.index_select(0, idx) * float(porosity)

    def analysis(self, obs_nodes, porosity, y, R, sakov=False, L=None, inflate=1.0, HX=None):
        """Assimilate soil-water-content observations at 0-based ``obs_nodes`` (theta = Sw * porosity, the 'swc' mapping of
        pyCATHY/DA/mapper.py) into the pressure-head ensemble, in place."""
        if HX is None:
            HX = self.predict_obs(obs_nodes, porosity)
        _, info = sharded_enkf_update(self.X, HX, y, R, sakov=sakov, L=L, inflate=inflate, n_infl=self.n, group=self.group)
        info["HX_local"] = HX
        return info

    def restart(self, tmax: float, deltat: float = 0.0):
        """Load the analysed heads back into the members and start the next window at time 0."""
        for j, s in enumerate(self.sims):
            s.unpack_psi(self.X.data_ptr(), self.ne_local, j)
            s.restart(tmax, deltat)

    def close(self):
        for s in self.sims:
            s.close()
        self.sims = []
