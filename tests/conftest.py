import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the oracle, the corpus tools and (when nvcc is present) the product library."""
    import oracle_lib
    oracle_lib.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])
    lib = os.path.join(ROOT, "parallelparsing_b200", "lib", "libppb200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "parallelparsing_b200", "csrc")])
    yield


@pytest.fixture(scope="session")
def device():
    import parallelparsing_b200 as pp
    return pp.Device.default(0)
