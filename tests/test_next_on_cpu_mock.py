"""GPU tests run on the CPU against a MOCK build of the library (tests/mock/).

What the mock is: the product's source files -- grid_b200/csrc/fermop.cu (operator compositions), dhop.cu (generic hopping kernel,
double store, leg mask, the multi-rank orchestration), dhop_fast.cu with dhop_fast.cuh / dhop_col.cuh (the tuned fp32 kernels: TMA
bulk copies of the links, cp.async rings, mbarriers, packed f32x2 math, the semi-fused multi-rank hop), halo_p2p.cu (pack + peer
stores + epoch flags), smat.cu (dense s-space kernel, persistent CTAs with a two-stage TMA pipeline), cayley.cu, stag.cu (improved
staggered operator incl. three-deep halos), solver.cu (CG, mixed / reliable-update / multishift solvers and their fused update
kernels), schur.cu, force.cu, nersc.cu -- compiled AS THEY ARE for the host through a stand-in cuda_runtime.h and a source rewriter:
a launch of a kernel whose threads do not talk to each other becomes a loop over blocks and threads; kernels that use shared memory,
barriers, mbarriers or asynchronous copies run one block at a time with a fibre per CUDA thread (tests/mock/simt.cpp), the one-line
PTX wrappers of the product (mbarrier, cp.async, cp.async.bulk, f32x2) are replaced by emulations and the system-scope flag
store / load of the peer-to-peer halos by host atomics.  A backend supplies the rest: field containers, import / export, BLAS-1 and
reductions as plain loops over the same blocked layout, and NCCL as mailboxes between threads.  No oracle inside: the tests compare
its results with the oracle, the golden fixtures and the compiled reference, exactly as they do on a GPU.

 * every GPU test of the SURVEY 8(f) rows, of the N-rank staggered path and of the tuned kernels' edge shapes passes on
   it (56 tests);
 * so do the measured suite's golden-vector and parity GPU tests -- now THROUGH the tuned kernels (the launch counters prove it), in
   both shapes: column-sweep kernel by default, micro-block kernel with GB_NO_COL=1, persistent s-space kernel looping over tiles;
 * with host threads as ranks the N-rank checks of scripts/mgpu_check.py pass too: peer-to-peer halos (a "peer mapping" is a plain
   pointer between threads), interior + exterior and semi-fused hops, improved staggered operator with three-deep halos.
It cannot see: memory ordering inside a block (fibres never run concurrently), bank conflicts, register pressure, launch limits,
real streams / events, the reductions of fields.cu -- the device is still needed for those and for every performance number.
The product has no CPU path: this lives under tests/ and is selected only by tests/conftest.py (GB_TEST_MOCK_LIB)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NEXT = ["tests/test_next_schur_solve.py", "tests/test_next_force.py", "tests/test_next_multishift.py", "tests/test_next_relupcg.py",
        "tests/test_next_nersc_io.py", "tests/test_next_stag_halo_gpu.py", "tests/test_next_tuned_shapes.py"]


@pytest.fixture(scope="module")
def mock_lib(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tests", "mock"))
    try:
        import build_mock
        return build_mock.build(str(tmp_path_factory.mktemp("gridb200_mock")))
    finally:
        sys.path.pop(0)


# every child run of this module, started side by side when the first test asks for one (they are independent processes; run one
# after the other they take four minutes): name -> (pytest arguments | script, counters to report, extra environment)
PARITY = ["tests/test_gpu_parity.py", "-k", "not device_random"]
JOBS = {
    "next_rows": (NEXT + ["-k", "not driver"], None, {}),
    "golden": (["tests/test_golden.py"], None, {}),
    "parity": (PARITY, ["dhop_col2_kernel", "dhop_col_kernel", "dhop_fast_kernel", "smat_kernel"], {}),
    "micro_block": (["tests/test_gpu_parity.py", "-k", "fast_and_generic or tiling or schur_operator or cg_matches"],
                    ["dhop_col2_kernel", "dhop_col_kernel", "dhop_fast_kernel", "smat_kernel"], {"GB_NO_COL": "1", "GB_MOCK_SM_COUNT": "3"}),
    "two_t_slices": (["tests/test_gpu_parity.py", "tests/test_next_tuned_shapes.py", "-k", "(fast_and_generic and dwf_col) or (edge_shapes and Ls8_n)"],
                     ["dhop_col_kernel<LS, 0, 0, 2>", "dhop_col_kernel<LS, 1, 0, 2>"], {"GB_COL_NT": "2", "GB_COL2": "0"}),
    "host_dhop": (["tests/test_gpu_self_halo.py", "-k", "host_dhop"], None, {}),
    "optional_forms": (["tests/test_gpu_recon12.py", "tests/test_gpu_self_halo.py", "-k", "recon12 or (compressed_halos and zt and nccl)"], None, {}),
    "n_rank": ("mgpu_on_mock.py", None, {}),
}


class Children:
    def __init__(self, mock_lib, outdir):
        self.procs = {}
        for name, (what, count, extra) in JOBS.items():
            # the children run side by side: keep the oracle's and numpy's thread pools small or they fight for the cores
            env = dict(os.environ, GB_TEST_MOCK_LIB=mock_lib, GB_MOCK_COUNT=";".join(count or []), OMP_NUM_THREADS="2",
                       OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", **extra)
            if isinstance(what, str):
                cmd = [sys.executable, os.path.join(ROOT, "tests", "mock", what), mock_lib]
            else:
                cmd = [sys.executable, os.path.join(ROOT, "tests", "mock", "run_counted.py"), *what, "-m", "gpu", "-q", "-p", "no:cacheprovider"]
            log = open(os.path.join(outdir, name + ".log"), "w")
            self.procs[name] = (subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=log, stderr=subprocess.STDOUT), log)

    def result(self, name):
        """-> (exit code, output) of the child, waiting for it if it is still running"""
        p, log = self.procs[name]
        try:
            p.wait(timeout=2400)
        except subprocess.TimeoutExpired:
            p.kill()
            raise
        log.close()
        return p.returncode, open(log.name).read()

    def passed_and_counts(self, name):
        rc, out = self.result(name)
        assert rc == 0, out[-3000:]
        summary = [l for l in out.splitlines() if " passed" in l][-1]
        assert "failed" not in summary and "xfailed" not in summary and "error" not in summary, out[-3000:]
        counts = {l.split()[1]: int(l.split()[2]) for l in out.splitlines() if l.startswith("COOP ")}   # (names come back without blanks)
        return int(summary.split(" passed")[0].split()[-1]), counts


@pytest.fixture(scope="module")
def children(mock_lib):
    c = Children(mock_lib, os.path.dirname(mock_lib))
    yield c
    for p, log in c.procs.values():
        if p.poll() is None:
            p.kill()


def test_next_row_gpu_tests_pass_on_the_cpu_mock(children):
    # the C++ drivers are linked against the real library; everything else of these files runs
    assert children.passed_and_counts("next_rows")[0] >= 56


def test_measured_golden_vector_gpu_tests_pass_on_the_cpu_mock(children):
    """the mock reproduces the reference's outputs: 4^4 x Ls 4 Wilson / DWF / Moebius / staggered operators, CG and mixed CG (the same
    tests are green on the B200)"""
    assert children.passed_and_counts("golden")[0] >= 10


def test_measured_parity_gpu_tests_pass_on_the_cpu_mock(children):
    """tests/test_gpu_parity.py (every operator entry, BLAS, reductions, CG on Wilson 8^4, DWF Ls 8, Moebius Ls 12; green on the B200)
    -- all but the device RNG, which the mock does not provide; the host-pipelined Dhop runs with copies as memmove and no stream
    concurrency.  The fp32 hops go through the
    column-sweep kernel, the s-space operators through the dense kernel (launch counters)"""
    n, c = children.passed_and_counts("parity")
    assert n >= 300 and c["dhop_col2_kernel"] > 50 and c["smat_kernel"] > 1000, (n, c)


def test_micro_block_kernel_and_persistent_s_space_kernel_on_the_cpu_mock(children):
    """the other tuned shapes: GB_NO_COL=1 sends every fp32 hop through the micro-block kernel (the interior pass on decomposed
    lattices), and with 3 "SMs" the s-space kernel's persistent CTAs loop over many tiles through their two-stage TMA pipeline"""
    n, c = children.passed_and_counts("micro_block")
    assert n >= 40 and c["dhop_col_kernel"] == 0 and c["dhop_col2_kernel"] == 0 and c["dhop_fast_kernel"] > 30 and c["smat_kernel"] > 100, (n, c)


def test_host_pipelined_dhop_on_decomposed_lattices_on_the_cpu_mock(children):
    """gb_op_dhop_host (dhop_host.cu) with a self halo in z, t, z+t (faces first, one exchange, slab hops reading the receive buffers)
    and x (import + hop + export), peer-to-peer and NCCL-path code, both precisions: the index arithmetic of the strided z-face import
    and of the slab hop's halo legs.  Stream / event ordering is the device's to prove (the same tests are green on the B200)."""
    assert children.passed_and_counts("host_dhop")[0] >= 32


def test_optional_forms_on_the_cpu_mock(children):
    """two-row (12-real) link storage (tests/test_gpu_recon12.py: every operator entry, boundary phases, decomposed lattice, the refusal of
    links that are not special unitary) and compressed halos (bf16 / fp32 on the wire; the mixed CG with the compressed inner operator runs here as the halfcomms driver)"""
    assert children.passed_and_counts("optional_forms")[0] >= 25


def test_two_t_slices_per_cta_column_kernel_on_the_cpu_mock(children):
    """GB_COL_NT=2 (opt-in variant, 512 threads: two adjacent t-slices share a CTA and read each other's ring slots for the t legs)"""
    n, c = children.passed_and_counts("two_t_slices")
    assert n >= 8 and c["dhop_col_kernel<LS,0,0,2>"] > 5 and c["dhop_col_kernel<LS,1,0,2>"] > 5, (n, c)


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


def test_halfcomms_driver_runs_on_the_cpu_mock(mock_lib):
    """drivers/Test_dwf_mixedcg_prec_halfcomms.cc (the reference's compiled-out program of that name, enabled): DomainWallFermionFH as the
    inner operator of the mixed and the reliable-update CG, z / t halos through the halo path (GB_SELF_HALO=12) so that they are compressed"""
    d = os.path.dirname(mock_lib)
    exe = os.path.join(d, "Test_dwf_mixedcg_prec_halfcomms")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "drivers", "Test_dwf_mixedcg_prec_halfcomms.cc"), "-L" + d, "-lgridb200_mock", "-Wl,-rpath," + d])
    p = subprocess.run([exe, "--grid", "4.4.4.4", "--Ls", "4"], capture_output=True, text=True, timeout=900, env=dict(os.environ, GB_SELF_HALO="12"))
    assert p.returncode == 0 and "done" in p.stdout, (p.stdout + p.stderr)[-2000:]


def test_n_rank_parity_on_the_cpu_mock(children):
    """tests/mock/mgpu_on_mock.py: ranks are host threads.  Decomposed Wilson / DWF / Moebius hops with the product's peer-to-peer
    halos (pack_send_kernel stores into the neighbour thread's receive buffer and publishes the epoch flag; the hop acquires it),
    overlapped and serial orchestration, gauge-face exchange, DhopDir legs across the boundary; the tuned fp32 path at Ls = 8 --
    the semi-fused launch on z / t splits (1.1.1.2, 1.1.2.2, 1.1.1.4), interior + exterior on a y split; the improved staggered operator
    with three-deep halos on 2 and 4 ranks; reductions, CG and the Schur solve against the oracle on the global lattice -- what
    scripts/mgpu_check.py checks on N GPUs.  The script fails if the semi-fused kernel or pack_send never ran."""
    rc, out = children.result("n_rank")
    assert rc == 0 and "MGPU_ON_MOCK PASS" in out, out[-3000:]


def test_the_mock_is_not_reachable_from_the_product():
    """the product package never mentions the mock or its environment switch"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "grid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "GB_TEST_MOCK_LIB" not in src and "gb_mock" not in src, f


@pytest.mark.parametrize("exe,args", [("Benchmark_dwf_fp32", ["--grid", "4.4.4.4", "-Ls", "8"]), ("Test_dwf_mixedcg_prec", ["--grid", "4.4.4.4", "--seconds", "1"])])
def test_unmodified_reference_programs_through_the_bridge_on_the_cpu_mock(mock_lib, exe, args):
    """bridge/_build/* are the reference's own Benchmark_dwf_fp32.cc / Test_dwf_mixedcg_prec.cc compiled unmodified against the C ABI
    (bridge/GridB200Bridge.h); preloading the mock library in front of libgridb200.so runs them here.  Their own asserts decide
    (Dhop against the Cshift implementation, Deo + Doe = D, |x_mixed - x_double| < 1e-4); tests/test_gpu_bridge.py is the GPU run."""
    path = os.path.join(ROOT, "bridge", "_build", exe)
    if not os.path.exists(path):
        pytest.skip("bridge/_build is made by __graft_entry__.build() where /root/reference is present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import test_gpu_bridge as tb
    finally:
        sys.path.pop(0)
    p = subprocess.run([path, *args], capture_output=True, text=True, timeout=1200, env=dict(os.environ, LD_PRELOAD=mock_lib, OMP_NUM_THREADS="4"))
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-2000:])
    (tb._check_benchmark if exe.startswith("Benchmark") else tb._check_mixedcg)(p.stdout)
