import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that stalls (a model that crawls without a step budget) must fail, not hang the suite: ctypes
    releases the GIL during the library call, so pytest-timeout's watchdog thread can end the run."""
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900))


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The suite needs the in-tree libraries (`__graft_entry__.build()` makes them).  On a fresh
    checkout with nvcc present they are compiled here once (cross-compilation needs no GPU); building
    is not a fallback of any kind -- a missing library is still an error in the product path."""
    import shutil
    from uclchem_b200 import build
    from uclchem_b200._capi import library_path
    if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
        for tag in ("default", "crp_photo", "gar"):
            if not library_path(tag).exists():
                build.compile(tag)
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def net():
    from uclchem_b200.network import load_default
    return load_default()


@pytest.fixture(scope="session")
def oracle(net):
    """The CPU restatement of the reference algorithm (test infrastructure)."""
    from oracle.oracle import Oracle
    return Oracle(net)


@pytest.fixture(scope="session")
def lib():
    """The product: CUDA library behind the C ABI.  Fails loudly if it was not built."""
    from uclchem_b200._capi import get_library
    L = get_library("default")
    L.init()
    return L


def max_dex(a, b, floor=1e-15):
    """max |log10(a/b)| over entries of b above `floor` (the north-star parity metric)."""
    a, b = np.asarray(a), np.asarray(b)
    m = b > floor
    return float(np.abs(np.log10(a[m] / b[m])).max())


STATIC = {"endAtFinalDensity": False, "freefall": False, "initialDens": 1e4, "initialTemp": 10.0,
          "finalDens": 1e5, "finalTime": 1.0e6}
