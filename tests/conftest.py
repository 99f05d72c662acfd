import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    # the product library (nvcc cross-compiles without a GPU) and the oracle
    from ndt_2d_b200 import build as nb
    nb.build()
    from oracle import binding
    binding.build()


_ensure_built()


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled in place (oracle/_ref); None if absent."""
    from oracle import binding
    return binding.load_ref()


@pytest.fixture(scope="session", params=["orc", "ref"])
def either(request, oracle, ref):
    """Runs a test once against our restatement and once against the compiled reference."""
    if request.param == "orc":
        return oracle
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree not available here)")
    return ref


def have_gpu() -> bool:
    from ndt_2d_b200 import lib
    return lib.ndt2d_device_count() > 0


@pytest.fixture(scope="session")
def gpu():
    if not have_gpu():
        pytest.fail("this test is marked gpu but no CUDA device is visible "
                    "(ndt_2d_b200 has no CPU fallback)")
    return True
