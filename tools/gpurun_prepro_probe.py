"""One pass of the device pre-processor on the synthetic SIZE x SIZE DEM (default 1000): the workload the ncu launch list /
full captures of the pre-processor kernels are taken on (tools/gpurun_r2g.sh)."""
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycathy_wrapper_b200 import preprocessor as pp, synthetic  # noqa: E402

n = int(os.environ.get("SIZE", "1000"))
d = "/tmp/prepro_probe"
os.makedirs(d, exist_ok=True)
synthetic.write_hapin(d + "/hap.in", n, n, 0.5, 0.5)
b = io.StringIO()
np.savetxt(b, synthetic.synthetic_dem(n, n), fmt="%.9f", delimiter="\t")
res = pp.terrain_analysis(open(d + "/hap.in").read(), b.getvalue())
print({k: v for k, v in res.info.items() if k != "hap_text"})
