"""CPU: the numpy oracle of the analysis updates (oracle/enkf_oracle.py) against golden vectors produced by the
REFERENCE's own functions (tests/golden/make_golden_enkf.py imports /root/reference/pyCATHY/DA/enkf.py and pf.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "enkf_golden.npz"))


def test_enkf_full_sakov_noparam(g):
    from oracle import enkf_oracle as o
    r = o.enkf_analysis(g["y"], g["R"], g["theta"], g["X"], g["HX"])
    assert np.array_equal(r[9], g["full_analysis"]) and np.array_equal(r[10], g["full_param"])
    r = o.enkf_analysis(g["y"], g["R"], g["theta"], g["X"], g["HX"], Sakov=True)
    assert np.array_equal(r[9], g["sakov_analysis"]) and np.array_equal(r[10], g["sakov_param"])
    r = o.enkf_analysis(g["y"], g["R"], [], g["X"], g["HX"])
    assert np.array_equal(r[9], g["noparam_analysis"])


def test_enkf_perturbed_observations(g):
    from oracle import enkf_oracle as o
    r = o.enkf_analysis(g["Ymat"], g["R"], g["theta"], g["X"], g["HX"])
    assert np.array_equal(r[9], g["pert_analysis"]) and np.array_equal(r[7], g["pert_B"]) and np.array_equal(r[8], g["pert_P"])


def test_enkf_localized_inflation(g):
    from oracle import enkf_oracle as o
    r = o.enkf_analysis_localized_with_inflation(g["y"], g["R"], g["X"], g["theta"], g["HX"], L=g["L"], Sakov=False,
                                                 inflate_states=1.05, inflate_params=1.1)
    assert np.array_equal(r[9], g["loc_analysis"]) and np.array_equal(r[10], g["loc_param"])


def test_particle_filter(g):
    from oracle import enkf_oracle as o
    r = o.particle_filter_analysis(g["y"], g["pf_R"], g["theta"], g["X"], g["HX"], u=float(g["pf_u"]))
    assert r["resampled"] and bool(g["pf_resampled"])
    assert np.array_equal(r["Analysis"], g["pf_analysis"]) and np.array_equal(r["Analysisparam"], g["pf_param"])
    assert r["n_eff"] == float(g["pf_neff"])
    r = o.particle_filter_analysis(g["y"], g["pf2_R"], g["theta"], g["X"], g["HX"])
    assert not r["resampled"] and np.array_equal(r["weights"], g["pf2_weights"]) and r["n_eff"] == float(g["pf2_neff"])


def test_enkf_library_exports_declared_symbols():
    """Every cathy_enkf_* / cathy_pf_* entry point declared in include/cathy_b200.h is exported (no compute without a GPU)."""
    import re
    from conftest import ROOT
    import __graft_entry__ as ge
    ge.build()
    from pycathy_wrapper_b200 import da
    lib = da.load_enkf_library()
    hdr = open(os.path.join(ROOT, "include", "cathy_b200.h")).read()
    names = set(re.findall(r"\b(cathy_(?:enkf|pf)_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 10
    for nm in names:
        assert hasattr(lib.lib, nm), nm


def test_gaspari_cohn_localisation(g):
    from oracle import enkf_oracle as o
    L = o.build_localization_matrix(g["gc_obs"], g["gc_grid"], float(g["gc_radius"]))
    assert np.array_equal(L, g["gc_matrix"]) and L[0, 0] == 1.0 and L.min() == 0.0
