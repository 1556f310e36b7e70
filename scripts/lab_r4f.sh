#!/bin/bash
set -u
out=gpurun_out/r4f; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_recon12.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -3 $out/pytest.log
timeout 200 python scripts/recon12_lab.py 32 2>&1 | tee $out/recon12_lab.jsonl | tail -8
