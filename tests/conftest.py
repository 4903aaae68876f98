"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run on CPU only: they pin the oracle (oracle/) against the committed golden
vectors of the unmodified reference and against analytic known answers, exercise the host logic,
and check that the C-ABI library loads and exports every symbol include/pas_b200.h declares.
`-m gpu` tests are the parity tests proper: they call the CUDA path through the C ABI and compare it
with the golden fixtures and with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    from tests import parity
    return parity.load_golden()


@pytest.fixture(scope="session")
def pas():
    import precomputed_atmospheric_scattering_b200 as module
    return module


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as module
    module.build()
    return module
