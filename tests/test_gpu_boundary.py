"""GPU: the process boundary pyCATHY uses -- `subprocess.run(["./cathy"], cwd=project)` (pyCATHY/cathy_tools.py:669) with the
launcher copied into the project like the reference's executable, the mesh-only mode IPRT1 = 3 (SRC/gen3d.f:89-112; pyCATHY
calls it through run_preprocessor / create_mesh_vtk, cathy_tools.py:724-729) and the no-clobber rule for output/grid3d
(SURVEY.md 8b).  The outputs are parsed with plain numpy here (the reference's own readers are exercised on the same writers
in tests/test_reference_readers.py, where the reference tree is mounted)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _launcher_into(prj):
    exe = os.path.join(prj, "cathy")
    shutil.copy(os.path.join(ROOT, "pycathy_wrapper_b200", "cathy"), exe)
    os.chmod(exe, 0o755)
    env = dict(os.environ)
    env["CATHY_B200_HOME"] = ROOT                    # the copied launcher no longer sits inside the package
    return env


def _set_iprt1(prj, value):
    path = os.path.join(prj, "input", "parm")
    lines = open(path).read().split("\n")
    tok = lines[0].split()
    tok[0] = str(value)
    lines[0] = " ".join(tok[:3]) + "\tIPRT1 NCOUT TRAFLAG"
    open(path, "w").write("\n".join(lines))


def test_cathy_launcher_as_child_process(gpu_lib, tmp_path):
    """./cathy in the project directory, no arguments: exit code 0, the reference's output files, same accepted steps as the ELF."""
    prj = str(tmp_path / "prj")
    shutil.copytree(os.path.join(GOLDEN, "weill_exemple"), prj)
    env = _launcher_into(prj)
    p = subprocess.run(["./cathy"], cwd=prj, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "TIME STEP:" in p.stdout
    got = np.loadtxt(os.path.join(prj, "output", "mbeconv"), skiprows=3)
    ref = np.loadtxt(os.path.join(GOLDEN, "weill_exemple", "golden", "mbeconv"), skiprows=3)
    assert got.shape == ref.shape and np.array_equal(got[:, [0, 3]], ref[:, [0, 3]])          # NSTEP, nonlinear iterations
    assert np.allclose(got[:, 1:3], ref[:, 1:3], rtol=1e-6) and np.allclose(got[:, 5], ref[:, 5], rtol=1e-6)
    for f in ("psi", "sw", "vp", "cumflowvol", "hgraph", "iter", "risul"):
        assert os.path.getsize(os.path.join(prj, "output", f)) > 0, f
    assert "cathy-b200 linear solver" in open(os.path.join(prj, "output", "risul")).read()      # the effective ITMXCG / TOLCG are on record


def test_mesh_only_mode_and_no_clobber(gpu_lib, tmp_path):
    """IPRT1 = 3 writes grid3d + xyz and terminates without a time loop; a normal run afterwards leaves them untouched."""
    prj = str(tmp_path / "prj")
    shutil.copytree(os.path.join(GOLDEN, "vtk6"), prj)
    env = _launcher_into(prj)
    _set_iprt1(prj, 3)
    p = subprocess.run(["./cathy"], cwd=prj, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    out = os.path.join(prj, "output")
    assert os.path.exists(os.path.join(out, "grid3d")) and os.path.exists(os.path.join(out, "xyz"))
    assert not os.path.exists(os.path.join(out, "psi")) and not os.path.exists(os.path.join(out, "mbeconv"))
    head = open(os.path.join(out, "grid3d")).readline().split()
    nnod, n, nt = (int(float(v)) for v in head[:3])
    assert (nnod, n, nt) == (7 * 6, 7 * 6 * 4, 6 * 5 * 6 * 3)
    tet = np.loadtxt(os.path.join(out, "grid3d"), skiprows=1, max_rows=nt)
    assert tet.shape == (nt, 5) and tet[:, :4].min() == 1 and tet[:, :4].max() == n
    marker = b"# untouched\n"
    with open(os.path.join(out, "grid3d"), "ab") as fh:
        fh.write(marker)
    before = open(os.path.join(out, "grid3d"), "rb").read()
    _set_iprt1(prj, 2)
    p = subprocess.run(["./cathy"], cwd=prj, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    assert open(os.path.join(out, "grid3d"), "rb").read() == before
    assert os.path.getsize(os.path.join(out, "psi")) > 0


def test_missing_library_makes_the_launcher_fail_loudly(tmp_path):
    """No CPU fallback at the process boundary either: with the CUDA library out of reach the child exits non-zero with a message."""
    prj = str(tmp_path / "prj")
    shutil.copytree(os.path.join(GOLDEN, "vtk6"), prj)
    env = _launcher_into(prj)
    code = ("import sys; sys.path.insert(0, %r); from pycathy_wrapper_b200 import capi; capi.library_path = lambda: '/nonexistent/lib.so'; "
            "from pycathy_wrapper_b200.processor import main; sys.exit(main([%r]))" % (ROOT, prj))
    p = subprocess.run([sys.executable, "-c", code], cwd=prj, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert p.returncode != 0 and "CathyLibraryError" in p.stdout
