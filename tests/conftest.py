import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference algorithm (oracle/sqp_oracle.cpp), built on demand with g++."""
    from oracle import bindings

    if not bindings.Oracle.available():
        bindings.build_oracle()
    return bindings.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled from /root/reference (only where oracle/_ref/libcorbo_ref.so exists)."""
    from oracle import bindings

    if not bindings.Reference.available():
        pytest.skip("oracle/_ref/libcorbo_ref.so not built (needs /root/reference)")
    return bindings.Reference()
