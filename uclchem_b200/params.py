"""Model parameters: names, defaults and dict -> column conversion.

Mirrors the reference's parameter block (``src/fortran_src/defaultparameters.f90:18-122``)
and its dictionary parser (``wrap.f90:699-983``): keys are case-insensitive,
unknown keys are an error (PARAMETER_READ_ERROR, ``wrap.f90:966-970``),
logicals accept Python bools.  The column order is the ``uclgpu_param`` enum of
``include/uclgpu.h``; unlike the reference (``wrap.f90:596``: defaults are never
re-applied between calls) every call starts from the defaults.
"""
from __future__ import annotations

import numpy as np

# (key, default) in uclgpu_param order
_PARAMS = [
    ("initialtemp", 10.0), ("initialdens", 1.0e2), ("finaldens", 1.0e5), ("currenttime", 0.0),
    ("finaltime", 5.0e6), ("radfield", 1.0), ("zeta", 1.0), ("rout", 0.05), ("rin", 0.0),
    ("baseav", 2.0), ("points", 1.0), ("bm0", 1.0),
    ("freezefactor", 1.0), ("endatfinaldensity", 0.0), ("freefall", 0.0), ("freefallfactor", 1.0),
    ("desorb", 1.0), ("h2desorb", 1.0), ("crdesorb", 1.0), ("uvdesorb", 1.0), ("thermdesorb", 1.0),
    ("instantsublimation", 0.0), ("cosmicrayattenuation", 0.0), ("ionmodel", 0.0),
    ("improvedh2crpdissociation", 0.0), ("enforcechargeconservation", 0.0),
    ("metallicity", 1.0), ("ion", 2.0), ("fh", 0.5), ("fhe", 0.1), ("fc", 1.77e-04), ("fo", 3.34e-04),
    ("fn", 6.18e-05), ("fs", 3.51e-6), ("fmg", 2.256e-06), ("fsi", 1.78e-06), ("fcl", 3.39e-08),
    ("fp", 7.78e-08), ("ffe", 2.01e-7), ("ff", 3.6e-08), ("fd", 0.0), ("fli", 0.0), ("fna", 0.0),
    ("fpah", 0.0), ("f15n", 0.0), ("f13c", 0.0), ("f18o", 0.0),
    ("reltol", 1e-8), ("abstol_factor", 1.0e-14), ("abstol_min", 1.0e-25), ("mxstep", 10000.0),
    ("ebmaxh2", 1.21e3), ("ebmaxcr", 1.21e3), ("ebmaxuvcr", 1.0e4), ("epsilon", 0.01),
    ("uv_yield", 0.03), ("phi", 1.0e5), ("uvcreff", 1.0e-3), ("omega", 0.5),
    ("temp_indx", 1.0), ("max_temperature", 300.0),
    ("shock_vel", 0.0), ("timestep_factor", 0.01), ("minimum_temperature", 0.0),
    ("collapse_mode", 0.0),
]
PARAM_NAMES = [k for k, _ in _PARAMS]
PARAM_INDEX = {k: i for i, k in enumerate(PARAM_NAMES)}
NPARAM = len(_PARAMS)
assert NPARAM == 65

# keys the reference's parser accepts but that do not influence the hot path
# (file names, output cadence); they are tolerated and ignored by the grid API.
_IGNORED = {
    "outputfile", "columnfile", "ratefile", "fluxfile", "writestep", "abundsavefile", "abundloadfile",
    "outspecies",
}
# default-real literals in defaultparameters.f90 that are NOT exactly representable:
# REAL(dp) :: x = <single-precision literal>  ->  value is float32-rounded
_F32_DEFAULTS = {"rout": 0.05, "fhe": 0.1, "epsilon": 0.01, "uv_yield": 0.03}

MODEL_KINDS = {"cloud": 0, "hot_core": 1, "cshock": 2, "collapse": 3, "jshock": 4, "postprocess": 5}


def default_params(ncell: int = 1) -> np.ndarray:
    """[NPARAM, ncell] array holding the defaults of defaultparameters.f90."""
    col = np.array([v for _, v in _PARAMS], dtype=np.float64)
    for k, v in _F32_DEFAULTS.items():
        col[PARAM_INDEX[k]] = float(np.float32(v))
    return np.repeat(col[:, None], ncell, axis=1)


def _to_float(key: str, v) -> float:
    if isinstance(v, (bool, np.bool_)):
        return 1.0 if v else 0.0
    if key == "ionmodel" and isinstance(v, str):
        if v.upper() not in ("L", "H"):
            raise ValueError("ionModel must be 'L' or 'H'")
        return 0.0 if v.upper() == "L" else 1.0
    return float(v)


def params_from_dict(param_dict: dict | None, ncell: int | None = None) -> np.ndarray:
    """Build the [NPARAM, ncell] parameter table from a reference-style dict.

    Scalars broadcast; array-valued entries become per-cell columns (all arrays
    must share one length).  Raises KeyError for unknown keys, the grid-level
    analogue of the reference's PARAMETER_READ_ERROR.
    """
    param_dict = dict(param_dict or {})
    lengths = {np.size(v) for v in param_dict.values() if np.ndim(v) > 0}
    if ncell is None:
        ncell = max(lengths) if lengths else 1
    for n in lengths:
        if n not in (1, ncell):
            raise ValueError(f"per-cell parameter of length {n} does not match ncell={ncell}")
    out = default_params(ncell)
    for key, v in param_dict.items():
        k = key.lower()
        if k in _IGNORED:
            continue
        if k not in PARAM_INDEX:
            raise KeyError(f"unknown parameter {key!r}")
        if np.ndim(v) > 0:
            out[PARAM_INDEX[k], :] = [_to_float(k, x) for x in np.ravel(v)]
        else:
            out[PARAM_INDEX[k], :] = _to_float(k, v)
    if (out[PARAM_INDEX["points"]] != 1.0).any():
        # multi-parcel clouds couple their parcels through column densities (chemistry.f90:169-181):
        # not a set of independent cells, and not built -- refuse rather than ignore
        raise ValueError("points > 1 (multi-parcel clouds) is not supported by the GPU path: "
                         "every cell of a grid is a single-point model")
    return out
