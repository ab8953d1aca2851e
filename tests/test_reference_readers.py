"""Boundary, output side: the files written through pycathy_wrapper_b200/outputs.py are parsed by the REFERENCE's own readers
(pyCATHY/importers/cathy_outputs.py: read_psi :480, read_sw :418, read_vp :255, read_mbeconv :607, read_cumflowvol :223,
read_grid3d :15, read_xyz :177, read_hgraph :337, read_wtdepth :198), loaded by file path with its unrelated imports
(matplotlib, xarray, pyCATHY.cathy_utils) stubbed -- the trick tests/golden/make_golden_enkf.py uses for enkf.py.
The run itself is done by the CPU oracle through the same processor + writers (no GPU here); the GPU side of the boundary
(`./cathy` as a child process, mesh-only mode, no-clobber rule) is tests/test_gpu_boundary.py.  Skipped where /root/reference
is not mounted (the GPU box)."""
import importlib.util
import os
import shutil
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN

REF = "/root/reference/pyCATHY/importers/cathy_outputs.py"


@pytest.fixture(scope="module")
def ref_readers():
    if not os.path.exists(REF):
        pytest.skip("reference tree not mounted")
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "xarray", "pyCATHY", "pyCATHY.cathy_utils")}
    for name in saved:
        if saved[name] is None:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.modules["pyCATHY"].cathy_utils = sys.modules["pyCATHY.cathy_utils"]
    spec = importlib.util.spec_from_file_location("ref_cathy_outputs", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yield mod
    for name, old in saved.items():
        if old is None:
            sys.modules.pop(name, None)


def test_outputs_parse_with_the_references_own_readers(oracle_mod, ref_readers, tmp_path):
    from pycathy_wrapper_b200.processor import run_processor
    dst = str(tmp_path / "prj")
    shutil.copytree(os.path.join(GOLDEN, "vtk6"), dst)
    # mesh-only mode first (IPRT1 = 3, SRC/gen3d.f:89-112): writes grid3d + xyz and stops
    run_processor(dst, lib=oracle_mod.load(), IPRT1=3)
    g3 = os.path.join(dst, "output", "grid3d")
    assert os.path.exists(g3) and os.path.exists(os.path.join(dst, "output", "xyz")) and not os.path.exists(os.path.join(dst, "output", "psi"))
    grid = ref_readers.read_grid3d(g3)
    before = open(g3, "rb").read()
    res = run_processor(dst, lib=oracle_mod.load())
    assert res.finished_ok
    assert open(g3, "rb").read() == before                       # a normal run leaves an existing grid3d untouched
    st = res.final_state
    n = st["psi"].size
    assert (int(grid["nnod3"]), int(grid["nel"])) == (n, res.reports[0]["nstep"] * 0 + grid["mesh_tetra"].shape[0])
    assert grid["mesh3d_nodes"].shape == (n, 3) and grid["mesh_tetra"].shape[1] == 5
    psi = ref_readers.read_psi(os.path.join(dst, "output", "psi"))
    assert psi.shape[1] == n and psi.index[0] == 0.0
    assert np.allclose(psi.iloc[-1].values, st["psi"], rtol=2e-6, atol=1e-7)          # 7 significant digits in the file
    sw, sw_times = ref_readers.read_sw(os.path.join(dst, "output", "sw"))
    assert sw.shape[1] == n and list(sw_times) == list(psi.index)
    assert np.allclose(sw.iloc[-1].values, st["sw"], rtol=2e-6, atol=1e-7)
    mb = ref_readers.read_mbeconv(os.path.join(dst, "output", "mbeconv"))
    assert list(mb["NSTEP"].astype(int)) == [r["nstep"] for r in res.reports]
    assert np.allclose(mb["DELTAT"].values, [r["deltat"] for r in res.reports], rtol=1e-6)
    assert np.allclose(mb["NLIN"].values, [r["iter"] for r in res.reports])
    assert np.allclose(mb["STORE1"].values, [r["store1"] for r in res.reports], rtol=1e-6)
    cf = np.atleast_2d(ref_readers.read_cumflowvol(os.path.join(dst, "output", "cumflowvol")))       # the reader skips 8 lines: header + first steps
    assert cf.shape[1] == 8 and int(cf[-1, 0]) == res.reports[-1]["nstep"] and abs(cf[-1, 2] - res.reports[-1]["time"]) <= 1e-2 * res.reports[-1]["time"]
    vp = ref_readers.read_vp(os.path.join(dst, "output", "vp"))
    assert len(vp) > 0
    xyz = ref_readers.read_xyz(os.path.join(dst, "output", "xyz"))
    assert len(xyz) == n
