#!/bin/bash
# ncu --set full of the round's other new kernels for the record: the hop-epilogue instantiations (four consecutive column-kernel launches of
# a CG iteration), the streaming CG kernels.
set -u
out=gpurun_out/r3l; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dhop_col2_kernel -s 20 -c 4 -o $out/col2_cg4 python scripts/cg_bench.py 32 16 single 40 > $out/n1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stri_kernel -s 6 -c 2 -o $out/stri_cg python scripts/cg_bench.py 32 16 single 40 > $out/n3.log 2>&1
ls -la $out | head
