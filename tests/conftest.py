import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def _have_gpu() -> bool:
    try:
        import cudavideostream_b200 as cvs
        return cvs.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly rather than silently pass on nothing
    if config.getoption("-m") == "gpu" and not _have_gpu():
        raise pytest.UsageError("-m gpu requested but no sm_100 device / libcvs_b200.so is available")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def cvs():
    import cudavideostream_b200 as m
    m.load_library()
    return m
