#!/bin/bash
# GPU call C: full ncu capture of the second-generation column-sweep hop (one launch), plus the GPU test suite.
set -u
out=gpurun_out/r2c; mkdir -p $out
N=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dhop_col2 -s 3 -c 1 -o $out/col2_full python scripts/prof_dhop.py > $out/ncu.log 2>&1
tail -3 $out/ncu.log
( time python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
