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
                            "break the verified suite, run last, and run in child processes with time limits (a crash, a sticky CUDA error or a hang "
                            "in new code cannot take the session down); the marker is removed once the test has passed on a B200 "
                            "(DESIGN.md section 8).")


def pytest_collection_modifyitems(config, items):
    for item in items:
        m = item.get_closest_marker("unverified")
        if m is not None and not _CHILD:
            item.add_marker(pytest.mark.xfail(strict=False, reason="never run on a GPU yet: " + (m.args[0] if m.args else "")))
    # verified tests first, unverified ones last
    items.sort(key=lambda it: it.get_closest_marker("unverified") is not None)


_file_outcomes = {}   # test file -> {nodeid: "passed" | "failed" | "skipped"} from ONE child run of all its unverified tests


def _run_child(args, timeout):
    import time
    env = dict(os.environ, GB_UNVERIFIED_CHILD="1")
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
        return p.returncode, p.stdout + p.stderr
    except subprocess.TimeoutExpired as e:
        return None, ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or ""))
    finally:
        _child_spent[0] += time.time() - t0


def _file_run(path):
    """All unverified tests of one file in ONE child process (a python + CUDA start-up per test would cost minutes of GPU time);
    -rA prints one 'PASSED nodeid' / 'FAILED nodeid' line per test.  Tests that have no line (the child crashed or hung before
    reaching them) are run one by one afterwards."""
    if path not in _file_outcomes:
        rc, out = _run_child(["-rA", "-m", "gpu", path], 4 * _CHILD_TIMEOUT_S)
        got = {}
        for line in out.splitlines():
            w = line.split()
            if len(w) >= 2 and w[0] in ("PASSED", "FAILED", "ERROR", "SKIPPED", "XFAIL", "XPASS") and "::" in w[1]:
                got[w[1]] = w[0].lower()
        _file_outcomes[path] = (got, out)
    return _file_outcomes[path]


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    """Unverified tests run in child pytest processes (GB_UNVERIFIED_CHILD=1 there: plain tests, real exit codes, time limits)."""
    if _CHILD or pyfuncitem.get_closest_marker("unverified") is None:
        return None
    if _child_spent[0] > _CHILD_BUDGET_S:
        pytest.fail("not run: the session's %d s budget for unverified tests is spent" % _CHILD_BUDGET_S, pytrace=False)
    got, out = _file_run(pyfuncitem.nodeid.split("::")[0])
    outcome = got.get(pyfuncitem.nodeid)
    if outcome == "passed":
        return True
    if outcome == "skipped":
        pytest.skip("skipped in the child process")
    if outcome is not None:
        pytest.fail("failed in the file's child process\n" + out[-3000:], pytrace=False)
    # no verdict from the file run (crash / hang before this test): run it alone
    rc, out1 = _run_child(["-x", pyfuncitem.nodeid], _CHILD_TIMEOUT_S)
    if rc is None:
        pytest.fail("unverified test exceeded %d s in its child process" % _CHILD_TIMEOUT_S, pytrace=False)
    if rc != 0:
        pytest.fail("child pytest exit %d\n%s" % (rc, out1[-3000:]), pytrace=False)
    return True
