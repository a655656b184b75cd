import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The reference-built oracle library (oracle/_ref/libgwat_ref.so); tests that need it skip when it is absent."""
    from oracle import gwat_ref
    if not gwat_ref.available():
        pytest.skip("oracle/_ref/libgwat_ref.so not built (needs /root/reference: make -C oracle)")
    gwat_ref.lib()
    return gwat_ref


@pytest.fixture(scope="session")
def ctx():
    """A gwat_b200 context on cuda:0.  Fails (does not skip) when the extension or the GPU is missing."""
    from gw_analysis_tools_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()
