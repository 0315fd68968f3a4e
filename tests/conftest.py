import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own kernels (oracle/_ref/libqie_ref.so), GPU only."""
    from oracle.oracle import REF_SO, Ref
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libqie_ref.so not built (needs /root/reference at build time)")
    return Ref()
