"""CPU: host-side logic -- project reader, Fortran-format writers, synthetic generator, C-ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_reader_parses_bundled_project():
    from pycathy_wrapper_b200.project import load_project
    p = load_project(os.path.join(GOLDEN, "weill_exemple"))
    assert (p.nrow, p.ncol, p.nstr, p.nnod, p.n, p.nt) == (20, 20, 15, 441, 7056, 36000)
    assert p.parm["ISIMGR"] == 2 and p.parm["TIMPRT"] == [1800.0, 7200.0] and p.parm["NODVP"] == [441]
    assert p.indp == 2 and p.ipond == 0 and p.hspatm == 1
    assert np.allclose(p.atm_times, [0.0, 86400.0]) and p.surf["qoi"][0] == 20
    assert abs(p.zratio.sum() - 1.0) < 1e-14 and p.soil["TABLE"].shape == (15, 1, 8)


def test_reader_rejects_out_of_scope_features(tmp_path):
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import CathyInputError, load_project
    d = synthetic.make_project(str(tmp_path / "a"), 4, 5, 3, TRAFLAG=1)
    with pytest.raises(CathyInputError, match="CATHY_B200_SKIP_TRANSPORT"):
        load_project(d)
    # explicit opt-in: the flow problem of a TRAFLAG=1 project (pyCATHY's template project carries TRAFLAG=1) is the TRAFLAG=0 problem
    p1, p0 = load_project(d, skip_transport=True), load_project(synthetic.make_project(str(tmp_path / "a0"), 4, 5, 3))
    assert p1.transport_skipped and not p0.transport_skipped
    assert p1.n == p0.n and np.array_equal(p1.dem, p0.dem) and np.array_equal(p1.ic_psi, p0.ic_psi)
    d = synthetic.make_project(str(tmp_path / "b"), 4, 5, 3)
    dem = os.path.join(d, "prepro", "dem")
    txt = open(dem).read().splitlines()
    txt[6] = "0.0 " + " ".join(txt[6].split()[1:])
    open(dem, "w").write("\n".join(txt) + "\n")
    with pytest.raises(CathyInputError):
        load_project(d)


def test_bc_table_grammar(tmp_path):
    from pycathy_wrapper_b200.project import read_bc_table
    f = tmp_path / "bc"
    f.write_text("0.0 TIME\n2 1\n3 4\n17\n1.5 2.5\n-0.25\n100.0 TIME\n0 0\n200. TIME\n-1 0\n")
    t = read_bc_table(str(f), nnod=6, nstr=2)
    assert t.times == [0.0, 100.0, 200.0] and t.n2d == [2, 0, -1]
    assert list(t.nodes[0]) == [3, 4, 9, 10, 15, 16, 17] and list(t.values[0]) == [1.5, 2.5, 1.5, 2.5, 1.5, 2.5, -0.25]
    assert len(t.nodes[1]) == 0 and list(t.nodes[2]) == [13, 14, 15, 16, 17, 18]


def test_fortran_edit_descriptors():
    from pycathy_wrapper_b200 import outputs as O
    assert O.fe(1.953125e-2, 13, 6) == " 1.953125E-02"
    assert O.fe(-2.11168e-6, 13, 5) == " -2.11168E-06"
    assert O.fe(0.0, 15, 6) == "   0.000000E+00"
    assert O.fe(1.7e-100, 15, 6) == "   1.700000-100"
    assert O.fi(235, 7) == "    235"
    line = O.mbeconv_line(1, 1.953125e-2, 1.953125e-2, 10, 9.3, 165.0, 165.0, -2.11168e-6, -2.11168e-6, 0.0, 0.0,
                          -1.054e-6, -1.054e-6, -1.054e-6, -1.054e-6, 1.058e-6, -100.4, 1.058e-6, 1.058e-6)
    ref = open(os.path.join(GOLDEN, "weill_exemple", "golden", "mbeconv")).read().splitlines()[3]
    assert line.rstrip("\n")[:60] == ref[:60] and len(line.rstrip("\n")) == len(ref)


def test_synthetic_project_round_trips(tmp_path):
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import load_project
    d = synthetic.make_project(str(tmp_path / "s"), 7, 9, 5, ic=("wt", 1.0), TMAX=50.0, TIMPRT=[50.0])
    p = load_project(d)
    assert (p.nrow, p.ncol, p.nstr) == (7, 9, 5) and p.indp == 3 and p.wtposition == 1.0
    s = 0.0
    for zr in p.zratio:
        s += float(zr)
    assert abs(s - 1.0) <= 1e-14
    assert np.all(p.dem > 0)


def test_cabi_library_exports_every_declared_symbol():
    """include/cathy_b200.h <-> libcathy_b200.so (no compute calls: this runs without a GPU)."""
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "cathy_b200.h")).read()
    declared = set(re.findall(r"\b(cathy_[a-z_0-9]+)\s*\(", hdr))
    lib = os.path.join(ROOT, "pycathy_wrapper_b200", "csrc", "libcathy_b200.so")
    nm = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (cathy_[a-z_0-9]+)", nm))
    assert declared and declared <= exported, declared - exported
    from pycathy_wrapper_b200 import capi
    L = capi.load_library()
    assert L.f["sizeof_problem"]() > 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pycathy_wrapper_b200")
    for r, _d, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(r, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", ""), os.path.join(r, f)


def test_reader_parses_seepage_faces(tmp_path):
    """input/sfbc (SRC/sfvone.f:27-66): TIME, NSF, then per face its node count and node ids; a set that changes during the run
    (SRC/sfvnxt.f) is refused; elevations must descend along a face (checked where the mesh is built: the oracle / the device)."""
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.project import CathyInputError, load_project
    nnod = 5 * 6
    faces = [[2 * nnod + 30, 3 * nnod + 30], [1 * nnod + 27, 2 * nnod + 27, 3 * nnod + 27]]
    d = synthetic.make_project(str(tmp_path / "a"), 4, 5, 3, seepage_faces=faces, ISFCVG=1, TMAX=100.0)
    p = load_project(d)
    assert [list(f) for f in p.seepage_faces] == faces and p.parm["ISFCVG"] == 1
    assert load_project(synthetic.make_project(str(tmp_path / "b"), 4, 5, 3)).seepage_faces == []
    with open(os.path.join(d, "input", "sfbc"), "w") as fh:
        fh.write("0.0\n1\n2\n%d %d\n50.0\n1\n2\n%d %d\n1e30\n0\n" % (faces[0][0], faces[0][1], faces[0][0], faces[0][1]))
    with pytest.raises(CathyInputError, match="time-varying"):
        load_project(d)


def test_oracle_checks_seepage_face_order(tmp_path):
    from oracle import oracle
    from pycathy_wrapper_b200 import synthetic
    from pycathy_wrapper_b200.capi import CathyLibraryError
    from pycathy_wrapper_b200.project import load_project
    nnod = 5 * 6
    d = synthetic.make_project(str(tmp_path / "a"), 4, 5, 3, seepage_faces=[[3 * nnod + 30, 2 * nnod + 30]], TMAX=100.0)
    with pytest.raises(CathyLibraryError, match="descending"):
        oracle.simulation(load_project(d))


def test_bench_reference_arm_of_the_preprocessor_workload_prints_the_contract_line():
    """bench.py --impl reference --workload prepro runs the reference's own ELF (CPU only) and prints one JSON line with the
    keys of the bench contract; without oracle/_ref it must say `unavailable` and exit 0."""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "prepro", "--size", "40x40x1", "--steps", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bin", "pycppp")):
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                  "cpu_baseline", "e2e"):
            assert k in line, k
        assert line["unit"] == "cells/s" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    else:
        assert "unavailable" in line


def test_loader_names_null_cells_when_only_qoi_a_reveals_them(tmp_path):
    """A project whose dem / zone rasters were rewritten as full rectangles but whose pre-processor saw an irregular catchment
    (shipped example: examples/SSHydro/meshing_from_subcachment): qoi_a then lists fewer cells than NROW x NCOL -- a clear refusal,
    not an end-of-file error."""
    import shutil
    import pytest
    from pycathy_wrapper_b200.project import CathyInputError, load_project
    dst = str(tmp_path / "prj")
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "weill_exemple"), dst)
    q = os.path.join(dst, "prepro", "qoi_a")
    lines = open(q).read().splitlines()
    open(q, "w").write("\n".join(["%12d" % 390] + lines[1:391]) + "\n")
    with pytest.raises(CathyInputError, match="qoi_a lists 390 catchment cells for a 20 x 20 DEM"):
        load_project(dst)
