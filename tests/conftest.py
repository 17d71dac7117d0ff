import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from ka9q_sdr_b200 import _lib
        return _lib.lib().ka9q_device_count() > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    """The verbatim reference library (oracle/_ref), with the reference mode table loaded."""
    from oracle import refbind as R
    if not R.available():
        pytest.skip("oracle/_ref/libka9q_ref.so not built")
    from ka9q_sdr_b200 import modes
    R.lib()
    R.set_fft_backend("standin")
    R.load_modes(modes.MODES.values())
    return R


GOLDEN = os.path.join(ROOT, "tests", "golden")
