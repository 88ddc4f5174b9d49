import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc4096():
    from oracle import binding as ob
    return ob.Oracle.default(4096, 20)


@pytest.fixture(scope="session")
def client4096():
    """HarnessClient for the server_test.cpp fixture shape: 10 items, ELEM_SIZE 7680, N=4096, t 20-bit."""
    from oracle import client as oc
    p = oc.create_pir_parameters(10, 7680, 1, 4096, 20)
    return oc.HarnessClient(p, seed=42)
