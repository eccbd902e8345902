"""Host logic and the C ABI boundary, without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from uclchem_b200 import _capi, model
from uclchem_b200.params import NPARAM, PARAM_INDEX, default_params, params_from_dict

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "uclgpu.h").read_text()
    declared = set(re.findall(r"\b(uclgpu_\w+)\s*\(", header))
    declared -= {"uclgpu_param", "uclgpu_phys"}
    assert declared == set(_capi.EXPORTED)
    lib = ctypes.CDLL(str(_capi.library_path("default")))
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_network_info_and_defaults_match_python_mirror():
    L = _capi.Library("default")
    assert (L.nspec, L.nreac, L.naug) == (335, 3203, 338)
    assert L.species[:3] == ["H", "H+", "H2"] and L.species[-3:] == ["E-", "BULK", "SURFACE"]
    np.testing.assert_array_equal(L.default_params(3), default_params(3))   # defaultparameters.f90 twice


def test_param_dict_semantics():
    p = params_from_dict({"initialDens": [1e3, 1e4, 1e5], "freefall": True, "ionModel": "H", "outputFile": "x.dat"})
    assert p.shape == (NPARAM, 3)
    assert list(p[PARAM_INDEX["initialdens"]]) == [1e3, 1e4, 1e5]
    assert (p[PARAM_INDEX["freefall"]] == 1.0).all() and (p[PARAM_INDEX["ionmodel"]] == 1.0).all()
    with pytest.raises(KeyError):          # wrap.f90:966-970 unknown key -> PARAMETER_READ_ERROR
        params_from_dict({"initialDensity": 1e4})
    with pytest.raises(ValueError):
        params_from_dict({"initialDens": [1, 2, 3], "zeta": [1, 2]})
    with pytest.raises(ValueError, match="multi-parcel"):   # refused, not silently treated as one point
        params_from_dict({"points": 3})
    # single-precision default literals are kept
    assert default_params(1)[PARAM_INDEX["fhe"], 0] == float(np.float32(0.1))


def test_pre_flight_checklist_mirrors_reference():
    with pytest.raises(RuntimeError, match="Offending keys"):
        model.pre_flight_checklist(True, False, False, None, {"outputfile": "a.dat"})
    with pytest.raises(AssertionError):
        model.pre_flight_checklist(False, False, True, None, {})
    with pytest.raises(AssertionError):
        model.pre_flight_checklist(False, False, False, np.zeros(335), {})


def test_no_cpu_fallback_without_device():
    """Without a usable sm_100 device the library refuses to compute (-100), it never falls back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _capi.Library("default")
    with pytest.raises(_capi.UclgpuError, match="-100"):
        L.init()
    with pytest.raises(_capi.UclgpuError):
        L.run_grid(0, default_params(1))


def test_missing_library_fails_loudly():
    with pytest.raises(_capi.UclgpuError, match="no CPU fallback"):
        _capi.Library("no_such_network")


def test_every_key_of_the_reference_parser_is_accounted_for():
    """dictionaryParser (wrap.f90:699-983) accepts these keys (names read from the reference, lower case).
    All of them are parameters of the GPU path or documented as not supported; anything else raises the
    reference's PARAMETER_READ_ERROR semantics (KeyError here)."""
    from uclchem_b200.params import _IGNORED
    parser_keys = [
        "alpha", "beta", "gamma", "initialtemp", "initialdens", "finaldens", "currenttime", "finaltime", "radfield",
        "zeta", "freezefactor", "rout", "rin", "baseav", "points", "bm0", "endatfinaldensity", "freefall",
        "freefallfactor", "desorb", "h2desorb", "crdesorb", "uvdesorb", "thermdesorb", "instantsublimation",
        "cosmicrayattenuation", "ionmodel", "improvedh2crpdissociation", "ion", "fhe", "fc", "fo", "fn", "fs", "fmg",
        "fsi", "fcl", "fp", "ff", "outspecies", "writestep", "ebmaxh2", "epsilon", "uvcreff", "ebmaxcr", "phi",
        "ebmaxuvcr", "uv_yield", "metallicity", "omega", "reltol", "abstol_factor", "abstol_min", "jacobian",
        "abundsavefile", "abundloadfile", "outputfile", "ratefile", "fluxfile", "columnfile", "fh", "ntime",
        "trajecfile"]
    # not parameter COLUMNS: the per-reaction alpha/beta/gamma dictionaries (taken out of the dictionary by
    # uclchem_b200.model and passed as uclgpu_opts.coeff_*), the user-Jacobian switch (the Jacobian is always
    # analytic) and the legacy postprocess file inputs (model.postprocess takes arrays, like the reference's)
    unsupported = {"alpha", "beta", "gamma", "jacobian", "ntime", "trajecfile"}
    from uclchem_b200.model import _coefficients
    pd_ = {"alpha": {3: 1e-9}, "gamma": {10: 5.0, 11: 6.0}, "zeta": 2.0}
    assert _coefficients(pd_) == [(0, 2, 1e-9), (2, 9, 5.0), (2, 10, 6.0)] and pd_ == {"zeta": 2.0}
    for k in parser_keys:
        if k in unsupported:
            with pytest.raises(KeyError):
                params_from_dict({k: 1.0})
        else:
            assert k in PARAM_INDEX or k in _IGNORED, k
            params_from_dict({k: 1 if k in ("ion", "points") else ("x.dat" if k.endswith("file") else 1.0)})
