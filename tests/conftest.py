import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the reference's own object code, built where /root/reference exists)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle.oracle_api import Oracle
    Oracle.lib()
    return Oracle
