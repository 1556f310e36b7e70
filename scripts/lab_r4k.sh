#!/bin/bash
set -u
out=gpurun_out/r4k; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_self_halo.py -k "host" -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
for e in "GB_SELF_HALO=0 GB_HOST_PIPE_ZCHUNKS=1" "GB_SELF_HALO=0" "GB_SELF_HALO=0 GB_HOST_PIPE_ZCHUNKS=8" "GB_SELF_HALO=12 GB_HOST_PIPE_ZCHUNKS=1" "GB_SELF_HALO=12"; do
  env $e timeout 100 python scripts/e2e_decomp_lab.py 32 32 32 32 16 10 2>&1 | tail -1 | sed "s/^/$e /" | tee -a $out/e2e.jsonl | cut -c1-330
done
