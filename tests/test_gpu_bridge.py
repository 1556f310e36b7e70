"""The reference's own acceptance programs for this path, UNMODIFIED, driving libgridb200.so.

bridge/GridB200Bridge.h subclasses the reference's DomainWallFermion<Impl> / MobiusFermion<Impl> and forwards the FermionOperator
virtuals (M, Mdag, Meooe*, Mooee*, MooeeInv*, Dhop, DhopOE, DhopEO, ImportGauge; ref: Grid/qcd/action/fermion/FermionOperator.h:63-77,144)
to the C ABI; bridge/Makefile compiles /root/reference/tests/Test_dwf_mixedcg_prec.cc and /root/reference/benchmarks/Benchmark_dwf_fp32.cc
as they are (the header is force-included and maps the operator names onto the bridge classes) and links them against the unmodified
reference (oracle/_ref/libgridref.so) and the CUDA library.  The binaries are built in this container (build(); the GPU box has no
/root/reference) and travel with the snapshot.  What is checked is the reference's own asserts -- the programs abort on failure:
  Benchmark_dwf_fp32.cc:321,379  Dhop and Dhop^dag against the reference's Cshift implementation (norm diff < 1e-4)
  Benchmark_dwf_fp32.cc:438      Deo + Doe == Dhop
  Test_dwf_mixedcg_prec.cc:212-215, 308-321  mixed-precision and double CG converge, |x_mixed - x_double|^2 < 1e-4
plus tighter bounds on the printed numbers here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "bridge", "_build")


def _run(exe, args, env=None, timeout=900):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: __graft_entry__.build() makes it where /root/reference is present")
    p = subprocess.run([path, *args], capture_output=True, text=True, timeout=timeout, env=dict(os.environ, OMP_NUM_THREADS="8", **(env or {})))
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-2000:])
    return p.stdout


def _check_benchmark(out):
    diffs = [float(x) for x in re.findall(r"norm (?:dag )?diff(?: even| odd)?\s+([0-9][0-9.eE+-]*)", out)]
    assert len(diffs) >= 3, out[-2000:]
    assert max(diffs) < 1e-9, diffs           # |Dhop - Cshift reference|^2 at unit-norm sources (fp32: ~1e-13), Deo + Doe - D exactly 0
    assert "Grid Finalize" in out


def _check_mixedcg(out):
    m = re.search(r"Diff between mixed and regular CG: ([0-9.eE+-]+)", out)
    assert m and float(m.group(1)) < 1e-10, out[-2000:]
    assert out.count("FlightRecorder is OK!") >= 2 and "Grid Finalize" in out


@pytest.mark.gpu
def test_reference_benchmark_dwf_fp32_drives_the_library():
    _check_benchmark(_run("Benchmark_dwf_fp32", ["--grid", "8.8.8.8", "-Ls", "16"]))


@pytest.mark.gpu
def test_reference_test_dwf_mixedcg_prec_drives_the_library():
    _check_mixedcg(_run("Test_dwf_mixedcg_prec", ["--grid", "8.8.8.8", "--seconds", "1"]))
