import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


@pytest.fixture(autouse=True)
def _guarded_allocations(request):
    """With SDORB_GUARD=1 in the environment (the debug mode that stands in for compute-sanitizer, which is closed on this pool)
    every device buffer of libsdorb carries guard bands: after each GPU test all bands -- of the live buffers and of those freed
    during the test -- must be intact."""
    yield
    if os.environ.get("SDORB_GUARD", "0") not in ("", "0") and request.node.get_closest_marker("gpu"):
        from sdslam_b200 import api
        ex = api.ORBextractor(1, 1.2, 1, 20, max_width=64, max_height=64, max_batch=1)
        try:
            assert ex.guard_check() == 0
        finally:
            ex.close()
