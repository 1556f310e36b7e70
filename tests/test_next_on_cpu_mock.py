"""GPU tests run on the CPU against a MOCK build of the library (tests/mock/).

What the mock is: product source files whose kernels' threads never communicate -- grid_b200/csrc/fermop.cu (operator
compositions), dhop.cu (generic hopping kernel, double store, leg mask, serial-comms orchestration), cayley.cu (M5D, MooeeInv),
stag.cu (improved staggered operator incl. the three-deep halo path in its self-exchange form), solver.cu (CG, mixed / reliable-
update / multishift solvers and their fused update kernels), schur.cu, force.cu, nersc.cu -- compiled AS THEY ARE for the host
through a stand-in cuda_runtime.h and a launch rewriter (each launch becomes a loop over blocks and threads), linked with a
backend that supplies the rest: field containers, import / export, BLAS-1 and reductions as plain loops over the same blocked
layout, and stubs that switch the tuned paths off (dhop_fast / dhop_col, smat, peer-to-peer halos, NCCL).  No oracle inside:
the tests compare its results with the oracle, the golden fixtures and the compiled reference, exactly as they do on a GPU.

 * every `unverified` GPU test of the SURVEY 8(f) rows and of the N-rank staggered path passes on it (41 tests; all but the C++
   drivers, which link the real library);
 * so do the measured suite's golden-vector GPU tests (tests/test_golden.py), which is what says the mock itself can be trusted;
 * with host threads as ranks and mailboxes as the network, the N-rank checks of scripts/mgpu_check.py pass too: the decomposed
   Wilson / DWF hops (measured green on 2, 4, 8 B200) and the improved staggered operator with three-deep halos (not yet run on GPUs).
It cannot see: the tuned fp32 kernels, the dense s-space kernel (so Ls = 8 / 12 / 16 operators), reductions, real streams, NCCL
and peer-to-peer halos, launch configuration -- the device is still needed for those.  The product has no CPU path: this lives
under tests/ and is selected only by tests/conftest.py (GB_TEST_MOCK_LIB)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEXT = ["tests/test_next_schur_solve.py", "tests/test_next_force.py", "tests/test_next_multishift.py", "tests/test_next_relupcg.py",
        "tests/test_next_nersc_io.py", "tests/test_next_stag_halo_gpu.py"]


@pytest.fixture(scope="module")
def mock_lib(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tests", "mock"))
    try:
        import build_mock
        return build_mock.build(str(tmp_path_factory.mktemp("gridb200_mock")))
    finally:
        sys.path.pop(0)


def run_gpu_tests_on_mock(mock_lib, files, extra=()):
    env = dict(os.environ, GB_UNVERIFIED_CHILD="1", GB_TEST_MOCK_LIB=mock_lib)
    p = subprocess.run([sys.executable, "-m", "pytest", *files, "-m", "gpu", *extra, "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=2400)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0, tail
    last = p.stdout.strip().splitlines()[-1]
    assert " passed" in last and "failed" not in last and "xfailed" not in last and "error" not in last, tail
    return int(last.split(" passed")[0].split()[-1])


def test_unverified_gpu_tests_pass_on_the_cpu_mock(mock_lib):
    # the C++ drivers are linked against the real library; everything else of these files runs
    assert run_gpu_tests_on_mock(mock_lib, NEXT, ("-k", "not driver")) >= 41


def test_measured_golden_vector_gpu_tests_pass_on_the_cpu_mock(mock_lib):
    """the mock reproduces the reference's outputs through the product's generic path: 4^4 x Ls 4 Wilson / DWF / Moebius / staggered
    operators, CG and mixed CG (the same tests are green on the B200 with the tuned kernels)"""
    assert run_gpu_tests_on_mock(mock_lib, ["tests/test_golden.py"]) >= 10


def test_measured_parity_gpu_tests_pass_on_the_cpu_mock(mock_lib):
    """tests/test_gpu_parity.py (every operator entry, BLAS, reductions, CG on Wilson 8^4, DWF Ls 8, Moebius Ls 12; green on the B200)
    through the mock's generic path -- all but the host-pipelined Dhop and the device RNG, which the mock does not provide"""
    assert run_gpu_tests_on_mock(mock_lib, ["tests/test_gpu_parity.py"], ("-k", "not dhop_host and not device_random")) >= 300


DRIVERS = [("Test_dwf_cg_schur", ["--grid", "4.4.4.4", "--Ls", "4"], "PASS"), ("Test_dwf_multishift", ["--grid", "4.4.4.4", "--Ls", "4"], "PASS"),
           ("Test_dwf_force", ["--grid", "4.4.4.4", "--Ls", "4"], "PASS"), ("Test_dwf_mixedcg_prec", ["--grid", "4.4.4.4", "--Ls", "4"], "done"),
           ("Benchmark_staggered", ["--grid", "4.4.4.4", "--ncall", "2"], "done"), ("Benchmark_dwf_fp32", ["--grid", "4.4.4.4", "--Ls", "4", "--ncall", "2"], "done")]


@pytest.mark.parametrize("name,args,word", DRIVERS)
def test_cpp_drivers_run_on_the_cpu_mock(mock_lib, name, args, word):
    """the reference-shaped C++ programs (drivers/*.cc over include/gridb200.hpp), linked against the mock instead of the CUDA library:
    their asserts are the reference's (residuals, |x_mixed - x_double|, Deo + Doe = D, the force identity)"""
    d = os.path.dirname(mock_lib)
    exe = os.path.join(d, name)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "drivers", name + ".cc"), "-L" + d, "-lgridb200_mock", "-Wl,-rpath," + d])
    p = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and word in p.stdout, (p.stdout + p.stderr)[-2000:]


def test_n_rank_parity_on_the_cpu_mock(mock_lib):
    """tests/mock/mgpu_on_mock.py: ranks are host threads, halo messages go through the mock's mailboxes.  Decomposed Wilson / DWF /
    Moebius hops (overlapped and serial orchestration, gauge-face exchange, DhopDir legs across the boundary) and the improved
    staggered operator with three-deep halos on 2 and 4 ranks (1.1.1.2, 2.1.1.1, 1.2.1.1, 1.1.2.2, 1.1.1.4), reductions, CG and the
    Schur solve against the oracle on the global lattice -- what scripts/mgpu_check.py checks on N GPUs."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mock", "mgpu_on_mock.py"), mock_lib], cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0 and "MGPU_ON_MOCK PASS" in p.stdout, (p.stdout + p.stderr)[-3000:]


def test_the_mock_is_not_reachable_from_the_product():
    """the product package never mentions the mock or its environment switch"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "GB_TEST_MOCK_LIB" not in src and "gb_mock" not in src, f
