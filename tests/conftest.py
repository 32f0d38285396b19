import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail loudly, not skip: there is no CPU fallback to fall back to.
    pass


@pytest.fixture(scope="session")
def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gs_oracle
    gs_oracle.build()
    return gs_oracle


@pytest.fixture(scope="session")
def native():
    from genestrip_b200 import build_native, capi
    build_native()
    return capi


@pytest.fixture(scope="session")
def gpu_ctx(native):
    assert _has_cuda(), "GPU tests need a CUDA device; there is no CPU fallback"
    ctx = native.Context()
    yield ctx
    ctx.close()
