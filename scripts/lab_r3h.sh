#!/bin/bash
set -u
out=gpurun_out/r3h; mkdir -p $out
timeout 900 python -m pytest tests/test_next_relupcg.py tests/test_next_multishift.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "batched or relup or multishift or checkerboards or cg" > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -3 $out/pytest.log
