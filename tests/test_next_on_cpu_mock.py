"""The GPU tests of the SURVEY 8(f) rows, run on the CPU against a MOCK of the library (tests/mock/).

What the mock is: the product's own source files for those rows -- grid_b200/csrc/solver.cu (ConjugateGradient, mixed / reliable-
update / multishift solvers and their fused update kernels), schur.cu (SchurRedBlack*Solve, physical 4D <-> 5D maps), force.cu
(DhopDir, DhopDeriv, MDeriv, Meo/MoeDeriv, MpcDeriv) and nersc.cu -- compiled for the host through a stand-in cuda_runtime.h and
a launch rewriter (kernels whose threads do not communicate run thread by thread), linked with a backend that implements what
those files CALL: field containers and BLAS on the same blocked layout, and the operator entry points served by the oracle.
So the orchestration, the Python mirror and the tests themselves are exercised here; the operator kernels, the leg mask in the
hopping kernel and the launch plumbing remain for the GPU (the same tests, there marked `unverified`).
The product library itself has no CPU path: this mock lives under tests/ and links oracle/, which the product never does."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["tests/test_next_schur_solve.py", "tests/test_next_force.py", "tests/test_next_multishift.py", "tests/test_next_relupcg.py",
         "tests/test_next_nersc_io.py"]


@pytest.fixture(scope="module")
def mock_lib(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tests", "mock"))
    try:
        import build_mock
        return build_mock.build(str(tmp_path_factory.mktemp("gridb200_mock")))
    finally:
        sys.path.pop(0)


def test_gpu_tests_of_the_next_rows_pass_on_the_cpu_mock(mock_lib):
    env = dict(os.environ, GB_UNVERIFIED_CHILD="1", GB_TEST_MOCK_LIB=mock_lib)
    # the C++ drivers are linked against the real library; everything else of these files runs
    p = subprocess.run([sys.executable, "-m", "pytest", *FILES, "-m", "gpu", "-k", "not driver", "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and "failed" not in p.stdout and "xfailed" not in p.stdout, tail
    npassed = int(p.stdout.strip().splitlines()[-1].split(" passed")[0].split()[-1])
    assert npassed >= 28, tail


def test_the_mock_is_not_reachable_from_the_product():
    """the product package never mentions the mock or its environment switch"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "GB_TEST_MOCK_LIB" not in src and "gb_mock" not in src, f
