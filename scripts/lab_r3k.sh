#!/bin/bash
# Final GPU call of the round: the whole -m gpu suite, smoke, the default bench invocation and the reference arm.
set -u
out=gpurun_out/r3k; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $out/pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py ) > $out/bench_default.json 2> $out/bench_default.err
echo "bench default rc $?"; tail -3 $out/bench_default.err | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/bench_ref.json 2> $out/bench_ref.err
echo "ref rc $?"; tail -c 400 $out/bench_ref.json
