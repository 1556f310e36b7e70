"""CPU check of the small kernels behind SURVEY 8(f) rows f1 / f2 (chiral_wall_kernel of schur.cu, insert_force_kernel of force.cu).

Their per-element bodies live in grid_b200/csrc/next_kernels.cuh as __host__ __device__ functions; tests/host/next_kernels_emul.cu
compiles them for the HOST and runs them for every global thread index on fields held in the device layout, against expectations
written here from the oracle (ImportUnphysicalFermion / ExportPhysicalFermion{Solution,Source}; the spin-traced outer product summed
over s of the force terms, ref: WilsonImpl.h:193-238).  What remains for the GPU is launch plumbing, not index arithmetic."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from grid_b200 import synthetic as syn
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("next_kernels") / "next_kernels_emul")
    subprocess.check_call([NVCC, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "host", "next_kernels_emul.cu")])
    return exe


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("dims,Ls", [((4, 4, 4, 4), 4), ((4, 6, 2, 8), 6), ((8, 4, 4, 2), 16), ((2, 2, 2, 2), 3)])
def test_kernel_bodies_on_the_cpu(emul, tmp_path, dims, Ls, prec):
    V4 = int(np.prod(dims))
    src4, src5, a5 = syn.random_fermion(dims, 1, seed=1), syn.random_fermion(dims, Ls, seed=2), syn.random_fermion(dims, Ls, seed=3)
    o = po.OracleOp(1, dims, Ls, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=1)
    files = {"src4": src4, "src5": src5, "a5": a5,
             "want_unphys": o.physical(po.IMPORT_UNPHYSICAL, src4), "want_sol": o.physical(po.EXPORT_PHYSICAL_SOLUTION, src5),
             "want_src": o.physical(po.EXPORT_PHYSICAL_SOURCE, src5)}
    b = src5.reshape(V4, Ls, 4, 3); a = a5.reshape(V4, Ls, 4, 3)
    force = np.einsum("xsac,xsad->xcd", b, np.conj(a))                 # sum_s sum_spin b[spin][c1] conj(a[spin][c2])
    files["want_force"] = np.repeat(force[:, None], 4, axis=1)
    for k, v in files.items():
        np.ascontiguousarray(v, dtype=np.complex128).tofile(tmp_path / f"{k}.bin")
    p = subprocess.run([emul, prec, *map(str, dims), str(Ls), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("max |diff|") == 5
