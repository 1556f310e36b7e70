#!/bin/bash
set -u
out=gpurun_out/r4a; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_self_halo.py tests/test_gpu_parity.py -k "host" -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -3 $out/pytest.log
for e in "GB_SELF_HALO=0" "GB_SELF_HALO=8" "GB_SELF_HALO=8 GB_HOST_PIPE_DECOMP=0" "GB_SELF_HALO=12" "GB_SELF_HALO=12 GB_HOST_PIPE_DECOMP=0"; do
  env $e timeout 200 python scripts/e2e_decomp_lab.py 32 32 32 32 16 8 2>&1 | tail -1 | tee -a $out/e2e.jsonl
done
env GB_SELF_HALO=12 timeout 200 python scripts/e2e_decomp_lab.py 64 64 32 16 16 4 2>&1 | tail -1 | tee -a $out/e2e.jsonl
