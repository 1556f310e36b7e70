import os
import sys

import pytest  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


if os.environ.get("GB_TEST_MOCK_LIB"):
    # tests/test_next_on_cpu_mock.py: run the GPU tests against the CPU mock of the library (tests/mock/)
    import grid_b200 as _gb
    _gb.LIB_PATH = os.environ["GB_TEST_MOCK_LIB"]
    _gb._LIB = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")
