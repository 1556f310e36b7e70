#!/bin/bash
# GPU call: the whole -m gpu suite and smoke with the final code of the round.
set -u
out=gpurun_out/r3e; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $out/pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -5 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
