"""oracle/oracle.py -- TEST INFRASTRUCTURE: python binding of the CPU restatement and
a runner for the reference's own prebuilt processor (oracle/_ref).  Imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pycathy_wrapper_b200.capi import CathyLib, Simulation  # noqa: E402

_LIB = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"], check=True)
    return os.path.join(HERE, "liboracle.so")


def load() -> CathyLib:
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "cathy_oracle.c")):
            build()
        _LIB = CathyLib(path, "oracle_")
    return _LIB


def simulation(prj, **kw) -> Simulation:
    return Simulation(load(), prj, **kw)


REF_BIN = {"20x20x15": "cathy_20x20x15", "20x20x15_newton": "cathy_20x20x15_newton", "100x50x15": "cathy_100x50x15",
           "20x20x15_zones": "cathy_20x20x15_zones"}


def ref_available(which: str = "20x20x15") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "bin", REF_BIN[which])) and \
        os.path.exists(os.path.join(HERE, "_ref", "lib", "libgfortran.so.5"))


def run_reference(project_dir: str, workdir: str, which: str = "20x20x15", timeout: float = 3600.0) -> float:
    """Copy input/prepro/cathy.fnames of `project_dir` into `workdir`, run the reference ELF there,
    return wall seconds.  Outputs land in workdir/output."""
    os.makedirs(workdir, exist_ok=True)
    for sub in ("input", "prepro"):
        dst = os.path.join(workdir, sub)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(project_dir, sub), dst,
                        ignore=shutil.ignore_patterns("src", "*.pdf", "basin_*", "cppp", "pycppp"))
        for r, _d, fs in os.walk(dst):
            os.chmod(r, 0o755)
            for f in fs:
                os.chmod(os.path.join(r, f), 0o644)
    shutil.copy(os.path.join(project_dir, "cathy.fnames"), os.path.join(workdir, "cathy.fnames"))
    os.chmod(os.path.join(workdir, "cathy.fnames"), 0o644)
    for sub in ("output", "vtk"):
        os.makedirs(os.path.join(workdir, sub), exist_ok=True)
    exe = os.path.join(workdir, "cathy_ref")
    shutil.copy(os.path.join(HERE, "_ref", "bin", REF_BIN[which]), exe)
    os.chmod(exe, 0o755)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(HERE, "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    t0 = time.time()
    with open(os.path.join(workdir, "stdout_ref.txt"), "w") as so:
        subprocess.run([exe], cwd=workdir, env=env, stdout=so, stderr=subprocess.STDOUT, timeout=timeout, check=False)
    return time.time() - t0


def run_prepro(project_dir: str, timeout: float = 600.0) -> None:
    """Run the reference's prebuilt pre-processor (oracle/_ref/bin/pycppp) in <project>/prepro so that
    the surface-routing rasters (dtm_*, qoi_a) of an ISIMGR=2 project are the reference's own."""
    exe = os.path.join(HERE, "_ref", "bin", "pycppp")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(HERE, "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    p = subprocess.run([exe], cwd=os.path.join(project_dir, "prepro"), env=env, input="2\n0\n1\n", text=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    if not os.path.exists(os.path.join(project_dir, "prepro", "qoi_a")):
        raise RuntimeError("pycppp failed:\n" + p.stdout[-2000:])
