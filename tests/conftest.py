import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_CHILD = os.environ.get("GB_UNVERIFIED_CHILD") == "1"
_CHILD_TIMEOUT_S = 240          # one unverified test (they are all 4^4 .. 8^4 lattices)
_CHILD_BUDGET_S = 1500          # all unverified tests of a session together; the rest are reported, not run
_child_spent = [0.0]


if os.environ.get("GB_TEST_MOCK_LIB"):
    # tests/test_next_on_cpu_mock.py: run the GPU tests of the SURVEY 8(f) rows against the CPU mock of the library (tests/mock/)
    import grid_b200 as _gb
    _gb.LIB_PATH = os.environ["GB_TEST_MOCK_LIB"]
    _gb._LIB = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")
    config.addinivalue_line("markers", "unverified(reason): a GPU test of code written after the round's GPU budget was spent -- it "
                            "has never run on a device.  Reported as xfail/xpass (non-strict) so that it can neither hide behind nor "
                            "break the verified suite, run last, and run in a child process with a time limit (a crash, a sticky CUDA error or a hang "
                            "in new code cannot take the session down); the marker is removed once the test has passed on a B200 "
                            "(DESIGN.md section 8).")


def pytest_collection_modifyitems(config, items):
    for item in items:
        m = item.get_closest_marker("unverified")
        if m is not None and not _CHILD:
            item.add_marker(pytest.mark.xfail(strict=False, reason="never run on a GPU yet: " + (m.args[0] if m.args else "")))
    # verified tests first, unverified ones last
    items.sort(key=lambda it: it.get_closest_marker("unverified") is not None)


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    """Unverified tests run in a child pytest process (GB_UNVERIFIED_CHILD=1 there: plain test, real exit code, 15 min limit)."""
    if _CHILD or pyfuncitem.get_closest_marker("unverified") is None:
        return None
    import time
    if _child_spent[0] > _CHILD_BUDGET_S:
        pytest.fail("not run: the session's %d s budget for unverified tests is spent" % _CHILD_BUDGET_S, pytrace=False)
    env = dict(os.environ, GB_UNVERIFIED_CHILD="1")
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", pyfuncitem.nodeid], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=_CHILD_TIMEOUT_S)
    except subprocess.TimeoutExpired:
        _child_spent[0] += time.time() - t0
        pytest.fail("unverified test exceeded %d s in its child process" % _CHILD_TIMEOUT_S, pytrace=False)
    _child_spent[0] += time.time() - t0
    if p.returncode != 0:
        pytest.fail("child pytest exit %d\n%s" % (p.returncode, (p.stdout + p.stderr)[-3000:]), pytrace=False)
    return True
