#!/bin/bash
# GPU call B: does dhop_col2_kernel run clean now?  compute-sanitizer on a small lattice, then the lab at full size.
set -u
out=gpurun_out/r2b; mkdir -p $out
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/lab_dhop.py 16 8 8 6 16 2 memcheck > $out/memcheck.log 2>&1
grep -E "ERROR SUMMARY|Unknown|Invalid|error" $out/memcheck.log | head -8; tail -2 $out/memcheck.log
GB_COL2_SYNC=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/lab_dhop.py 16 8 8 6 16 2 memcheck_sync > $out/memcheck_sync.log 2>&1
grep -E "ERROR SUMMARY|Unknown|Invalid|error" $out/memcheck_sync.log | head -8; tail -2 $out/memcheck_sync.log
GB_SELF_HALO=12 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/lab_dhop.py 16 8 8 6 16 2 memcheck_halo > $out/memcheck_halo.log 2>&1
grep -E "ERROR SUMMARY|Unknown|Invalid|error" $out/memcheck_halo.log | head -8; tail -2 $out/memcheck_halo.log
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab GB_COL2=1
lab GB_COL2_SYNC=1
lab GB_COL2=0
