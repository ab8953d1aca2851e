"""Golden vectors for the EnKF analysis, produced by the REFERENCE's own functions
(/root/reference/pyCATHY/DA/enkf.py: enkf_analysis :16-224, enkf_analysis_localized_with_inflation :225-342).
pyCATHY cannot be imported here (matplotlib, shapely... are missing), so the module file is loaded directly with
`matplotlib` stubbed out -- its arithmetic is plain numpy.  Run HERE (needs /root/reference):
    python tests/golden/make_golden_enkf.py
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_enkf():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_enkf", "/root/reference/pyCATHY/DA/enkf.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_enkf()
    rng = np.random.default_rng(1234)
    n, ne, m, npar = 600, 32, 12, 3
    X = -1.0 + 0.3 * rng.standard_normal((n, ne))
    theta = np.log(1.88e-4) + 0.5 * rng.standard_normal((npar, ne))
    H = rng.choice(n, m, replace=False)
    HX = X[H, :] + 0.01 * rng.standard_normal((m, ne))
    y = X[H, :].mean(axis=1) + 0.2 * rng.standard_normal(m)
    R = np.diag(np.full(m, 0.02 ** 2)) + 1e-5 * np.ones((m, m))
    L = np.exp(-np.abs(np.arange(n)[:, None] - H[None, :]) / 80.0)
    out = dict(X=X, theta=theta, HX=HX, y=y, R=R, L=L)
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref.enkf_analysis(y.copy(), R.copy(), theta.copy(), X.copy(), HX.copy())
        out["full_analysis"], out["full_param"] = r[9], r[10]
        r = ref.enkf_analysis(y.copy(), R.copy(), theta.copy(), X.copy(), HX.copy(), Sakov=True)
        out["sakov_analysis"], out["sakov_param"] = r[9], r[10]
        r = ref.enkf_analysis(y.copy(), R.copy(), [], X.copy(), HX.copy())
        out["noparam_analysis"] = r[9]
        r = ref.enkf_analysis_localized_with_inflation(y.copy(), R.copy(), X.copy(), theta.copy(), HX.copy(), L=L.copy(), Sakov=False,
                                                       inflate_states=1.05, inflate_params=1.1, jitter_params=0.0)
        out["loc_analysis"], out["loc_param"] = r[9], r[10]
        # perturbed observations: one data column per member (data.ndim == 2 branch, enkf.py:123-124)
        Y = y[:, None] + 0.02 * rng.standard_normal((m, ne))
        out["Ymat"] = Y
        r = ref.enkf_analysis(Y.copy(), R.copy(), theta.copy(), X.copy(), HX.copy())
        out["pert_analysis"], out["pert_param"], out["pert_B"], out["pert_P"] = r[9], r[10], r[7], r[8]
        # particle filter (pf.py imports numpy only): weights, n_eff, systematic resampling; jitter off
        spec = importlib.util.spec_from_file_location("ref_pf", "/root/reference/pyCATHY/DA/pf.py")
        pf = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pf)
        Rpf = np.diag(np.full(m, 0.15 ** 2))
        np.random.seed(7)
        out["pf_u"] = np.array(np.random.RandomState(7).rand())
        r = pf.particle_filter_analysis(y.copy(), Rpf, theta.copy(), X.copy(), HX.copy(), jitter_std_param=0.0, jitter_std_state=0.0)
        out["pf_R"] = Rpf
        out["pf_analysis"], out["pf_param"], out["pf_weights"], out["pf_neff"] = r["Analysis"], r["Analysisparam"], r["weights"], np.array(r["n_eff"])
        out["pf_resampled"] = np.array(r["resampled"])
        Rpf2 = np.diag(np.full(m, 2.0 ** 2))           # wide likelihood: no resampling, weights returned as they are
        r = pf.particle_filter_analysis(y.copy(), Rpf2, theta.copy(), X.copy(), HX.copy(), jitter_std_param=0.0, jitter_std_state=0.0)
        out["pf2_R"] = Rpf2
        out["pf2_weights"], out["pf2_neff"], out["pf2_resampled"] = r["weights"], np.array(r["n_eff"]), np.array(r["resampled"])
    # Gaspari-Cohn localisation (localisation.py imports the pyCATHY package at module level, which cannot be imported here:
    # its two pure-numpy functions are extracted from the source text and executed as they are)
    src = open("/root/reference/pyCATHY/DA/localisation.py").read()
    ns = {"np": np}
    for name in ("gaspari_cohn", "build_localization_matrix"):
        i = src.index("def %s(" % name)
        j = src.index("\ndef ", i + 1) if "\ndef " in src[i + 1:] else len(src)
        exec(src[i:j], ns)
    grid = np.column_stack([rng.uniform(0, 10, 200), rng.uniform(0, 10, 200)])
    obs = np.column_stack([rng.uniform(0, 10, 9), rng.uniform(0, 10, 9)])
    obs[0] = grid[0]
    out["gc_grid"], out["gc_obs"], out["gc_radius"] = grid, obs, np.array(1.7)
    out["gc_matrix"] = ns["build_localization_matrix"](obs, grid, 1.7)
    np.savez_compressed(os.path.join(HERE, "enkf_golden.npz"), **out)
    print("enkf golden written:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
