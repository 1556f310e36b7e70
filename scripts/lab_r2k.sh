#!/bin/bash
# GPU call K: reproducibility probe (same call twice -> same bits?) at 16^4 and 32^4, bridge tests again.
set -u
out=gpurun_out/r2k; mkdir -p $out
timeout 600 python scripts/determinism_check.py 16 16 2>&1 | tee $out/det16.log | cut -c1-400
timeout 900 python scripts/determinism_check.py 32 16 2>&1 | tee $out/det32.log | cut -c1-400
timeout 900 python scripts/determinism_check.py 32 16 2>&1 | tee $out/det32b.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_bridge.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
