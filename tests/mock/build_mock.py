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

PRODUCT_SOURCES = ["fermop.cu", "dhop.cu", "cayley.cu", "stag.cu", "solver.cu", "schur.cu", "force.cu", "nersc.cu"]


def build(outdir, sanitize=False):
    """sanitize=True: AddressSanitizer + UBSan build (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so))"""
    os.makedirs(outdir, exist_ok=True)
    csrc = os.path.join(ROOT, "grid_b200", "csrc")
    cpps = []
    for f in PRODUCT_SOURCES:
        out = os.path.join(outdir, f.replace(".cu", "_mock.cpp"))
        open(out, "w").write(f"// generated from grid_b200/csrc/{f} by tests/mock/transform.py\n" + transform(open(os.path.join(csrc, f)).read()))
        cpps.append(out)
    lib = os.path.join(outdir, "libgridb200_mock.so")
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if sanitize else []
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", *san, "-I", os.path.join(HERE, "shim"), "-I", csrc, "-o", lib, *cpps,
           os.path.join(HERE, "mock_backend.cpp"), "-Wl,--no-undefined"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_build"), sanitize="--sanitize" in sys.argv))
