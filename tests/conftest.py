import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle.pyoracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.pyoracle import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libsnpref.so not built (reference tree absent)")
    return Oracle("reference")
