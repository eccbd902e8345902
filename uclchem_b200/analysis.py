"""Analysis hook of the reference on the host: ``rates_to_dy_and_flux`` (``src/uclchem/analysis.py:737-890``).

Given the arrays a model returns with ``return_rates=True`` (physics, abundances, rate coefficients per
output time) it re-evaluates GETYDOT: the flux of every reaction and ``ydot`` of every species.  The
reference builds an incidence matrix from its species / reaction objects; here the same quantities
come from the tables the MakeRates CUDA back-end already holds (flux factors of the extended state,
signed gather lists, three-phase transfer -- the formulas of ``table_emulator.TableEngine.rhs``, which
the CPU tests pin on the reference-generated ``odes.f90``), vectorised over the rows.

This is analysis, not the hot path: whole grids of states go through ``uclgpu_get_odes`` on the device.
"""
from __future__ import annotations

import numpy as np

from . import symbolic
from .network import EXT_BLR, EXT_INV_SM, EXT_ONE, EXT_SWAP_SM, N_EXT, Network
from .table_emulator import COV0, GAS_DUST_DENSITY_RATIO, NUM_SITES_PER_GRAIN

_SYM_CACHE: dict = {}


def _sym(net: Network):
    key = id(net)
    if key not in _SYM_CACHE:
        _SYM_CACHE[key] = symbolic.build(net)
    return _SYM_CACHE[key]


def rates_to_dy_and_flux(physics, abundances, rates, network: Network):
    """physics [T, 8] (column 1 = density), abundances [T, nspec], rates [T, nreac] (arrays or DataFrames)
    -> (dy [T, nspec], flux_by_reaction [T, nreac]); DataFrames in, DataFrames out."""
    frames = hasattr(abundances, "columns")
    cols = list(abundances.columns) if frames else None
    phys = np.asarray(physics, float)
    ab = np.asarray(abundances, float)
    rt = np.asarray(rates, float)
    sym, net = _sym(network), network
    T, neq = ab.shape[0], sym.neq
    y = np.empty((T, neq + N_EXT))
    y[:, : net.nspec] = ab
    y[:, net.nspec] = phys[:, 1]
    sm = np.maximum(1e-30, ab[:, sym.iS])
    sb = np.maximum(1e-30, ab[:, sym.iB])
    blr = np.minimum(1.0, NUM_SITES_PER_GRAIN / (GAS_DUST_DENSITY_RATIO * sb))
    swap = (rt[:, sym.swap_reacs] * ab[:, net.re[sym.swap_reacs, 0]]).sum(axis=1) * blr
    y[:, neq + EXT_ONE] = 1.0
    y[:, neq + EXT_BLR] = blr
    y[:, neq + EXT_INV_SM] = 1.0 / sm
    y[:, neq + EXT_SWAP_SM] = swap / sm
    flux = rt * np.prod(y[:, sym.flux_f], axis=2)
    rows = np.repeat(np.arange(net.nspec), np.diff(sym.g_ptr))
    dy = np.zeros((T, net.nspec))
    np.add.at(dy.T, rows, (flux[:, sym.g_reac] * sym.g_sign).T)
    # three-phase transfer (odes.f90 tail): uncorrected surface growth moved between surface and bulk
    surf, bulk = net.surface_list, net.bulk_list
    S = dy[:, surf].sum(axis=1)
    mb = dy[:, bulk].sum(axis=1)
    shrink = S < 0
    q = np.where(shrink, np.minimum(1.0, sb / sm) / sb, COV0)
    c = (S * q)[:, None] * np.where(shrink[:, None], ab[:, bulk], ab[:, surf])
    dy[:, surf] -= c
    dy[:, bulk] += c
    dy[:, sym.iB] = mb + c.sum(axis=1)
    dy[:, sym.iS] = S - c.sum(axis=1)
    if frames:
        import pandas as pd
        return pd.DataFrame(dy, columns=cols), pd.DataFrame(flux)
    return dy, flux
