"""uclchem_b200.analysis.rates_to_dy_and_flux against the reference-generated odes.f90 (golden RHS cases)
and against the oracle's F along a trajectory."""
import numpy as np
from conftest import GOLDEN

from uclchem_b200.analysis import rates_to_dy_and_flux


def test_dy_matches_oracle_getydot(oracle, net):
    """Same states as the golden RHS cases, with the scalars GETYDOT receives derived from the state the
    way chemistry.f90:195-200 derives them (the oracle's GETYDOT is pinned bit-exact on the
    reference-generated odes.f90 in tests/test_oracle_golden.py)."""
    from uclchem_b200.table_emulator import COV0, GAS_DUST_DENSITY_RATIO, NUM_SITES_PER_GRAIN
    g = np.load(GOLDEN / "getydot_cases.npz")
    for i in range(6):
        y, rate = g[f"y_{i}"], g[f"rate_{i}"]
        sm = max(1e-30, y[net.species_idx["nsurface"]])
        sb = max(1e-30, y[net.species_idx["nbulk"]])
        blr = min(1.0, NUM_SITES_PER_GRAIN / (GAS_DUST_DENSITY_RATIO * sb))
        ref, _ = oracle.getydot(rate, y, blr, COV0, sm, sb, y[-1])
        phys = np.zeros((1, 8))
        phys[0, 1] = y[-1]
        dy, flux = rates_to_dy_and_flux(phys, y[None, :335], rate[None, :], net)
        assert dy.shape == (1, 335) and flux.shape == (1, net.nreac)
        assert np.abs(dy[0] - ref[:335]).max() <= 1e-13 * np.abs(ref[:335]).max()


def test_dy_and_flux_along_an_oracle_trajectory(oracle, net):
    from uclchem_b200.params import params_from_dict
    p = params_from_dict({"initialDens": 1e4, "initialTemp": 10.0, "finalTime": 1e4})[:, 0]
    r = oracle.run_model(0, p, rates=True)
    assert r["flag"] == 0
    phys, ab, rt = r["physics"][1:], r["abund"][1:], r["rates"][1:]
    dy, flux = rates_to_dy_and_flux(phys, ab, rt, net)
    # every reaction's flux is its rate coefficient times its reactant abundances (and density factors folded
    # into the rate by the reference): non-negative, and zero where the rate is zero
    assert (flux >= 0).all() and ((rt == 0) <= (flux == 0)).all()
    # elements conserved by every reaction: carbon budget of dy vanishes
    import re
    w = np.array([sum(int(n or 1) for s, n in re.findall(r"(CL|MG|SI|HE|[A-Z])(\d*)", nm.lstrip("#@").rstrip("+-")) if s == "C")
                  for nm in net.names[:333]], float)
    scale = np.abs(dy[:, :333] * w).sum(axis=1)
    assert (np.abs(dy[:, :333] @ w) <= 1e-12 * scale + 1e-40).all()
    # DataFrame in -> DataFrame out, same numbers
    import pandas as pd
    dfy, dff = rates_to_dy_and_flux(pd.DataFrame(phys), pd.DataFrame(ab, columns=net.names), pd.DataFrame(rt), net)
    assert list(dfy.columns) == net.names and np.array_equal(dfy.to_numpy(), dy) and np.array_equal(dff.to_numpy(), flux)
