"""CPU check of the index logic of the multi-GPU improved staggered operator (three-deep Naik halos).

grid_b200/csrc/stag_halo.cuh holds every piece of index arithmetic the pack, double-store and hopping kernels use on a
decomposed lattice, as __host__ __device__ functions.  tests/host/stag_halo_emul.cu compiles them for the HOST (nvcc, no
kernel launch, no device) and runs them for every rank of an emulated processor grid, with the neighbour table of the
library's own gb_geometry_query: pack -> exchange -> 16-point neighbour lookup must return the global periodic neighbour of
every site (both parities), and the gauge-face exchange must hand the double store U_mu(x + d mu) for d = -3..+2
(ref: Grid/qcd/action/fermion/StaggeredImpl.h:105-162 ; Grid/stencil/Stencil.h:709 displacement < local extent).
The arithmetic of the kernels themselves is covered on the GPU by scripts/mgpu_check.py (N ranks) against the oracle.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("stag_halo") / "stag_halo_emul")
    libdir = os.path.join(ROOT, "grid_b200")
    subprocess.check_call([NVCC, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "host", "stag_halo_emul.cu"),
                           "-L" + libdir, "-lgridb200", "-Xlinker", "-rpath," + libdir])
    return exe


@pytest.mark.parametrize("gdims,mpi", [
    ((8, 8, 8, 8), (1, 1, 1, 2)),       # t split in two: forward and backward neighbour are the same rank
    ((8, 8, 8, 8), (2, 1, 1, 1)),       # x split: the parity constraint halves y on the face
    ((8, 12, 8, 8), (1, 2, 1, 1)),
    ((8, 8, 12, 16), (1, 1, 2, 4)),     # BASELINE config 4/5 processor grid, local extents 6 and 4 in the split dimensions
    ((8, 8, 8, 8), (2, 2, 2, 2)),       # every dimension split, local extent 4 = the minimum for a three-deep halo
    ((4, 4, 6, 8), (1, 1, 1, 1)),       # one rank: pure periodic wrap
])
def test_three_deep_halo_index_logic(emul, gdims, mpi):
    p = subprocess.run([emul, *map(str, gdims), *map(str, mpi)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 wrong" in p.stdout
