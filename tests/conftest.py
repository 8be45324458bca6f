import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import pyoracle
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def ctx():
    """one mcg_ctx on cuda:0 -- fails loudly when the extension or the GPU is missing"""
    from metacache_mpi_b200 import api
    c = api.Context(0)
    yield c
    c.close()
