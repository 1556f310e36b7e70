#!/usr/bin/env python
"""Builds the TEST-ONLY CPU mock of libgridb200 (tests/mock/README.md): the product's fermop.cu, dhop.cu, cayley.cu, stag.cu, solver.cu, schur.cu, force.cu and nersc.cu,
rewritten for a host compiler by transform.py against shim/, linked with mock_backend.cpp.
usage: build_mock.py <output directory>  -> <output directory>/libgridb200_mock.so"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from transform import transform  # noqa: E402

PRODUCT_SOURCES = ["fermop.cu", "dhop.cu", "dhop_host.cu", "dhop_fast.cu", "halo_p2p.cu", "smat.cu", "cayley.cu", "stag.cu", "solver.cu", "schur.cu", "force.cu", "nersc.cu"]
# headers that hold kernels with inline PTX or shared memory: rewritten too, and found first on the include path
PRODUCT_HEADERS = ["dhop_fast.cuh", "dhop_col.cuh", "dhop_col2.cuh"]


def build(outdir, sanitize=False):
    """sanitize=True: AddressSanitizer + UBSan build (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so))"""
    os.makedirs(outdir, exist_ok=True)
    csrc = os.path.join(ROOT, "grid_b200", "csrc")
    cpps = []
    for f in PRODUCT_HEADERS:
        open(os.path.join(outdir, f), "w").write(f"// generated from grid_b200/csrc/{f} by tests/mock/transform.py\n" + transform(open(os.path.join(csrc, f)).read()))
    for f in PRODUCT_SOURCES:
        out = os.path.join(outdir, f.replace(".cu", "_mock.cpp"))
        open(out, "w").write(f"// generated from grid_b200/csrc/{f} by tests/mock/transform.py\n" + transform(open(os.path.join(csrc, f)).read()))
        cpps.append(out)
    lib = os.path.join(outdir, "libgridb200_mock.so")
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if sanitize else []
    flags = ["-std=c++17", "-O1", "-fPIC", "-pthread", *san, "-I", os.path.join(HERE, "shim"), "-I", outdir, "-I", csrc]
    srcs = cpps + [os.path.join(HERE, "mock_backend.cpp"), os.path.join(HERE, "simt.cpp")]
    objs = [os.path.join(outdir, os.path.basename(c)[:-4] + ".o") for c in srcs]
    procs = [subprocess.Popen(["g++", *flags, "-c", c, "-o", o]) for c, o in zip(srcs, objs)]   # one compiler per file, side by side
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("mock build failed")
    subprocess.check_call(["g++", "-shared", "-pthread", *san, "-o", lib, *objs, "-Wl,--no-undefined"])
    return lib


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_build"), sanitize="--sanitize" in sys.argv))
