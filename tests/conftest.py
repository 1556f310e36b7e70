import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")
    config.addinivalue_line("markers", "unverified(reason): a GPU test of code written after the round's GPU budget was spent -- it "
                            "has never run on a device.  Reported as xfail/xpass (non-strict) so that it can neither hide behind nor "
                            "break the verified suite; the marker is removed once the test has passed on a B200 (DESIGN.md section 8).")


def pytest_collection_modifyitems(config, items):
    for item in items:
        m = item.get_closest_marker("unverified")
        if m is not None:
            item.add_marker(pytest.mark.xfail(strict=False, reason="never run on a GPU yet: " + (m.args[0] if m.args else "")))
    # verified tests first, unverified ones last (a sticky CUDA error in new code must not take verified tests down with it)
    items.sort(key=lambda it: it.get_closest_marker("unverified") is not None)
