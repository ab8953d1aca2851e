"""One EnKF analysis at BASELINE config 4 size (N = 163,216 states, Ne = 256 members, m = 64 observations) on synthetic data --
the launch ncu captures for k_enkf_crosscov / k_enkf_update (tools/gpurun_r2_profiles.sh), and a CUDA-event timing of both."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

g.build()
from pycathy_wrapper_b200 import da  # noqa: E402

n, ne, m = 163216, int(os.environ.get("NE", "256")), 64
rng = np.random.default_rng(5)
X = torch.from_numpy(rng.standard_normal((n, ne))).cuda()
obs = np.linspace(0, n - 1, m).astype(np.int64)
HX = X.index_select(0, torch.as_tensor(obs, device="cuda")).contiguous()
y = HX.mean(dim=1).cpu().numpy() + 0.02 * rng.standard_normal(m)
R = np.diag(np.full(m, 0.02 ** 2))
for it in range(3):
    _, info = da.sharded_enkf_update(X, HX, y, R, sakov=False, inflate=1.02, n_infl=n)
t = info["timing_ms"]
bytes_cc, bytes_up = 8.0 * n * (ne + m), 8.0 * n * (2 * ne + m)
flop = 2.0 * n * ne * m
print("N=%d Ne=%d m=%d  crosscov %.3f ms (%.0f GB/s, %.1f TFLOP/s fp64)  update %.3f ms (%.0f GB/s, %.1f TFLOP/s fp64)"
      % (n, ne, m, t["crosscov_ms"], bytes_cc / t["crosscov_ms"] / 1e6, flop / t["crosscov_ms"] / 1e9, t["update_ms"], bytes_up / t["update_ms"] / 1e6,
         flop / t["update_ms"] / 1e9))
