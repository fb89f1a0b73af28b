import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libx264_cuda.so")) or \
       not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        ge.build()
    return ge.load_pkg()


@pytest.fixture(scope="session")
def port(pkg):
    import xo_api
    return xo_api.port()


@pytest.fixture(scope="session")
def ref(pkg):
    import xo_api
    if not xo_api.have_ref():
        pytest.skip("oracle/_ref not built (reference sources only exist in the build container)")
    return xo_api.ref()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)  # raises (loudly) when there is no GPU: no CPU fallback exists
    yield c
    c.close()
