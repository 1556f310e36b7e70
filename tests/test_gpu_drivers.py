"""The C++ drivers shaped like the reference's own programs run to completion (their asserts are the reference's)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(name, *args, env=None):
    exe = os.path.join(ROOT, "drivers", name)
    assert os.path.exists(exe), f"{exe} missing: run make -C grid_b200"
    p = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def test_benchmark_dwf_fp32_driver():
    """ref: benchmarks/Benchmark_dwf_fp32.cc -- Deo + Doe == Dunprec (assert n2e < 1e-4, :438)"""
    out = run("Benchmark_dwf_fp32", "--grid", "8.8.8.8", "--Ls", "8", "--ncall", "10")
    assert "norm diff" in out and "done" in out


def test_dwf_mixedcg_prec_driver():
    """ref: tests/Test_dwf_mixedcg_prec.cc -- |x_mixed - x_double|^2 < 1e-4 (:212-215), Ls=12 as in the reference"""
    out = run("Test_dwf_mixedcg_prec", "--grid", "8.8.8.8", "--Ls", "12")
    assert "Diff between mixed and regular CG" in out and "done" in out


def test_dwf_mixedcg_prec_halfcomms_driver():
    """ref: tests/Test_dwf_mixedcg_prec_halfcomms.cc:71-114 (compiled out in the reference, :33-34) -- mixed-precision and
    reliable-update CG whose inner operator is DomainWallFermionFH (compressed halos) against the double CG; one GPU, so the z and t
    halos are routed through the halo path with GB_SELF_HALO=12"""
    out = run("Test_dwf_mixedcg_prec_halfcomms", "--grid", "8.8.8.8", "--Ls", "24", env={"GB_SELF_HALO": "12"})
    assert "Diff between mixed and regular CG" in out and "Diff between reliable update and regular CG" in out and "done" in out


def test_benchmark_staggered_driver():
    """ref: benchmarks/Benchmark_staggered.cc (+ the even-odd / anti-Hermiticity checks of tests/core/Test_staggered.cc)"""
    out = run("Benchmark_staggered", "--grid", "8.8.8.8", "--ncall", "10")
    assert "norm diff" in out and "CG iterations" in out and "done" in out
