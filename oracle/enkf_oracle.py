"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the ensemble analysis updates of pyCATHY's data assimilation.

Plain-numpy restatement of the arithmetic in
  * /root/reference/pyCATHY/DA/enkf.py:16-224   ``enkf_analysis``
  * /root/reference/pyCATHY/DA/enkf.py:225-342  ``enkf_analysis_localized_with_inflation``
  * /root/reference/pyCATHY/DA/pf.py:3-110      ``particle_filter_analysis`` (weights, n_eff, resampling; no jitter)
  * /root/reference/pyCATHY/DA/pf.py:197-211    ``systematic_resample``
  * /root/reference/pyCATHY/DA/localisation.py:136-188  ``gaspari_cohn``, ``build_localization_matrix``
Parity is PINNED: tests/golden/enkf_golden.npz holds outputs of the reference's own functions (module loaded from
/root/reference with matplotlib stubbed, see tests/golden/make_golden_enkf.py) and tests/test_enkf_oracle.py checks
this file against them.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product (pycathy_wrapper_b200/da.py) never does.
"""
from __future__ import annotations

import numpy as np


def enkf_analysis(data, data_cov, param, ensemble, predict_obs, Sakov=False):
    """enkf.py:16-224.  Returns the reference's 11-slot list."""
    ens_size = ensemble.shape[1]                       # :67
    sim_size = ensemble.shape[0]
    meas_size = data.shape[0]
    ensemble_mean = np.tile(np.mean(ensemble, axis=1, keepdims=True), (1, ens_size))    # :82-83
    if len(param) > 0:                                 # :86-97
        param_mean = np.tile(np.mean(param, axis=1, keepdims=True), (1, ens_size))
        augm_state_mean = np.vstack([ensemble_mean, param_mean])
        augm_state = np.vstack([ensemble, param])
    else:
        augm_state_mean = ensemble_mean
        augm_state = ensemble
    augm_state_pert = augm_state - augm_state_mean     # :106
    if data.ndim > 0:                                  # :123-126
        data_pert = (data.T - predict_obs.T).T
    else:
        data_pert = (data - predict_obs.T).T
    obs_avg = (1.0 / ens_size) * np.tile(predict_obs.reshape(meas_size, ens_size).sum(1), (ens_size, 1)).T   # :143-145
    obs_pert = predict_obs - obs_avg                   # :146
    if Sakov:                                          # :160-167
        COV = data_cov.T
        inv_data_pert = data_pert / np.diag(COV)[:, None]
    else:
        COV = (1.0 / (ens_size - 1)) * (obs_pert @ obs_pert.T) + data_cov.T
        inv_data_pert = np.linalg.solve(COV, data_pert)
    ensemble_pert = (1.0 / (ens_size - 1)) * (augm_state_pert @ obs_pert.T)     # :180
    analysis = augm_state + (ensemble_pert @ inv_data_pert)                     # :197
    analysis_param = analysis[sim_size:, :].T          # :205-206
    analysis = analysis[0:sim_size, :]
    return [augm_state, augm_state_mean, augm_state_pert, data_pert, obs_avg, obs_pert, COV, inv_data_pert,
            ensemble_pert, analysis, analysis_param]


def enkf_analysis_localized_with_inflation(data, data_cov, ensemble, param, predict_obs, L=None, Sakov=False,
                                           inflate_states=1.0, inflate_params=1.0):
    """enkf.py:225-342 (jitter_params = 0: the reference draws from the global numpy RNG there)."""
    ens_size = ensemble.shape[1]
    sim_size = ensemble.shape[0]
    ensemble_mean = np.mean(ensemble, axis=1, keepdims=True)       # :283-288
    param_mean = np.mean(param, axis=1, keepdims=True)
    augm_mean = np.vstack([ensemble_mean, param_mean])
    augm_state = np.vstack([ensemble, param])
    augm_state_pert = augm_state - np.tile(augm_mean, (1, ens_size))
    ensemble_pert = ensemble - np.tile(ensemble_mean, (1, ens_size))
    obs_avg = np.mean(predict_obs, axis=1, keepdims=True)          # :291-293
    obs_pert = predict_obs - obs_avg
    data_pert = data.reshape(-1, 1) - predict_obs
    if Sakov:                                                      # :296-301
        COV = data_cov
        inv_data_pert = data_pert / np.diag(COV)[:, None]
    else:
        COV = (obs_pert @ obs_pert.T) / (ens_size - 1) + data_cov
        inv_data_pert = np.linalg.solve(COV, data_pert)
    P_xo = (augm_state_pert @ obs_pert.T) / (ens_size - 1)         # :304
    if L is not None:                                              # :307-310
        if L.shape != P_xo[:sim_size, :].shape:
            raise ValueError("localisation matrix shape mismatch")
        P_xo[:sim_size, :] = P_xo[:sim_size, :] * L
    analysis_augm = augm_state + P_xo @ inv_data_pert              # :313-315
    analysis = analysis_augm[:sim_size, :]
    analysis_param = analysis_augm[sim_size:, :]
    if inflate_states != 1.0:                                      # :318-324
        mean_s = np.mean(analysis, axis=1, keepdims=True)
        analysis = mean_s + inflate_states * (analysis - mean_s)
    if inflate_params != 1.0:
        mean_p = np.mean(analysis_param, axis=1, keepdims=True)
        analysis_param = mean_p + inflate_params * (analysis_param - mean_p)
    return [augm_state, augm_mean, augm_state_pert, data_pert, obs_avg, obs_pert, COV, inv_data_pert, ensemble_pert,
            analysis, analysis_param]


def systematic_resample(weights, n_particles, u):
    """pf.py:197-211 with the uniform draw `u` supplied by the caller instead of np.random.rand()."""
    positions = (np.arange(n_particles) + u) / n_particles
    return np.searchsorted(np.cumsum(weights), positions)


def particle_filter_analysis(data, data_cov, param, ensemble, observation, resample_threshold=0.5, u=0.5):
    """pf.py:3-110 without jitter / hybrid update: weights, n_eff, systematic resampling of columns."""
    ens_size = ensemble.shape[1]
    if observation.shape[0] == ens_size:               # :44-45
        observation = observation.T
    data = np.atleast_1d(data.flatten())
    obs_std = np.sqrt(np.diag(data_cov)) if data_cov.ndim == 2 else np.sqrt(data_cov)   # :52-55
    log_weights = np.zeros(ens_size)
    for i in range(ens_size):                          # :66-73
        innovation = data - observation[:, i]
        log_weights[i] = -0.5 * np.sum((innovation / obs_std) ** 2)
    log_weights -= log_weights.max()                   # :76-78
    weights = np.exp(log_weights)
    weights /= weights.sum()
    n_eff = 1.0 / np.sum(weights ** 2)                 # :81
    resampled = False
    if n_eff < resample_threshold * ens_size:          # :97-112
        resampled = True
        indices = systematic_resample(weights, ens_size, u)
        ensemble = ensemble[:, indices]
        param = param[:, indices]
        observation = observation[:, indices]
        weights_out = np.ones(ens_size) / ens_size
    else:
        indices = np.arange(ens_size)
        weights_out = weights
    return {"Analysis": ensemble, "Analysisparam": param, "weights": weights_out, "raw_weights": weights, "n_eff": n_eff,
            "resampled": resampled, "observation": observation, "indices": indices}


def gaspari_cohn(r, L):
    """/root/reference/pyCATHY/DA/localisation.py:136-152 (verbatim arithmetic)."""
    r = np.abs(r) / L
    w = np.zeros_like(r)
    mask1 = r <= 1
    mask2 = (r > 1) & (r <= 2)
    w[mask1] = (((-0.25 * r[mask1] + 0.5) * r[mask1] + 0.625) * r[mask1] - 5 / 3) * r[mask1] ** 2 + 1
    w[mask2] = ((((r[mask2] / 12 - 0.5) * r[mask2] + 0.625) * r[mask2] + 5 / 3) * r[mask2] - 5) * r[mask2] + 4 - 2 / (3 * r[mask2])
    w[r > 2] = 0
    return w


def build_localization_matrix(obs_pos, grid_pos, L):
    """/root/reference/pyCATHY/DA/localisation.py:155-188."""
    n_grid, n_obs = grid_pos.shape[0], obs_pos.shape[0]
    loc = np.zeros((n_grid, n_obs))
    for i in range(n_grid):
        dx = obs_pos[:, 0] - grid_pos[i, 0]
        dy = obs_pos[:, 1] - grid_pos[i, 1]
        loc[i, :] = gaspari_cohn(np.sqrt(dx ** 2 + dy ** 2), L)
    return loc
