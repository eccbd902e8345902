"""Second network (config 5, `libuclgpu_crp_photo.so`) on the GPU against the oracle.

The library for this network (compact shared-memory layout, uclchem_b200/build.py) was finished after
the round's GPU budget was spent: it has been compiled and load-checked, never run.  The test is
therefore opt-in (UCLGPU_TEST_SECOND_NETWORK=1) and marked xfail(strict=False): it records the first
hardware outcome without gating the suite -- XPASS means parity is green, xfail means the library still
needs work; remove both markers once seen green."""
import numpy as np
import pytest
from conftest import ROOT, max_dex

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(__import__("os").environ.get("UCLGPU_TEST_SECOND_NETWORK") != "1",
                    reason="never run on hardware yet: opt in with UCLGPU_TEST_SECOND_NETWORK=1 under an outer "
                           "`timeout` (tools/next_round_first_call.sh does) so that a misbehaving first run "
                           "cannot take the rest of the GPU suite with it")
@pytest.mark.xfail(strict=False, reason="first run on hardware: compile-checked only so far")
def test_static_clouds_on_the_crp_photo_network_match_the_oracle():
    from oracle.oracle import Oracle
    from uclchem_b200._capi import Library
    from uclchem_b200.network import Network
    from uclchem_b200.params import params_from_dict
    net2 = Network.from_json(ROOT / "uclchem_b200" / "networks" / "crp_photo.json")
    L = Library("crp_photo")
    L.init([0])
    try:
        # tolerances of the reference's own test on this network (tests/test_photo_on_grain.py:112-113)
        p = params_from_dict({"initialDens": [1e4, 1e5], "initialTemp": [10.0, 20.0], "finalTime": 1e4,
                              "reltol": 1e-5, "abstol_min": 1e-15})
        out = L.run_grid(0, p, step_budget=200000)   # a cell that crawls is abandoned after seconds
        assert (out["flag"] == 0).all(), out["flag"]
        ref, _, flag, _ = Oracle(net2).run_grid(0, p, nthreads=2)
        assert (flag == 0).all()
        for c in range(2):
            # reltol 1e-5 on both sides: trajectories may differ at that level, far inside 0.01 dex
            assert max_dex(out["y_final"][c, : net2.nspec], ref[c, : net2.nspec]) <= 0.01
    finally:
        L.shutdown()
