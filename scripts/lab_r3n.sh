#!/bin/bash
set -u
out=gpurun_out/r3n; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_self_halo.py tests/test_gpu_stress.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl | cut -c1-250; }
for DIMS in "32 32 32 32" "64 64 32 16"; do
  lab LAB_X=1
  lab GB_SELF_HALO=8
  lab GB_SELF_HALO=8 GB_T_LL=0
  lab GB_SELF_HALO=12
  lab GB_SELF_HALO=12 GB_T_LL=0
done
for e in "GB_SELF_HALO=12" "GB_SELF_HALO=12 GB_T_LL=0" "LAB_X=1"; do env $e timeout 300 python scripts/cg_bench.py 32 16 single 300 | tail -1 | sed "s/^/$e /" | cut -c1-200; done
