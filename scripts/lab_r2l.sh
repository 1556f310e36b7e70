#!/bin/bash
# GPU call L: which component makes the 32^4 CG irreproducible?  Bisect by kernel form; memcheck + initcheck one short solve.
set -u
out=gpurun_out/r2l; mkdir -p $out
run() { env "$@" timeout 300 python scripts/cg_repro.py 32 16 6 2>&1 | tail -1 | tee -a $out/repro.jsonl | cut -c1-330; }
run GB_CG_UNFUSED=1
run GB_CG_UNFUSED=1 GB_COL2=0
run GB_CG_UNFUSED=1 GB_NO_COL=1
run GB_CG_UNFUSED=1 LAB_GENERIC=1
run GB_CG_UNFUSED=1 CUDA_LAUNCH_BLOCKING=1
run GB_COL2=0
run LAB_X=1
GB_CG_UNFUSED=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/cg_repro.py 32 16 1 1e-1 > $out/memcheck.log 2>&1; tail -4 $out/memcheck.log | cut -c1-300
GB_CG_UNFUSED=1 timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python scripts/cg_repro.py 32 16 1 1e-1 > $out/initcheck.log 2>&1; tail -4 $out/initcheck.log | cut -c1-300
