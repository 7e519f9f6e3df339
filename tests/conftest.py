import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build every native artefact once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from tests import refapi
    return refapi.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    from tests import refapi
    r = refapi.Ref.try_load()
    if r is None:
        pytest.skip("oracle/_ref/libhydra_ref.so not present (it is built where /root/reference exists and travels as a binary)")
    return r


@pytest.fixture(scope="session")
def layer(built):
    import hydracore_b200 as hc
    lay = hc.CudaLayer()
    lay.SetShadowTrees(0)       # the parity tests compare with the CPU integrators, whose shadow rays see the first BVH tree only (IntegratorCommon::shadowTrace)
    yield lay
    lay.close()
