#!/bin/bash
set -u
out=gpurun_out/r3c; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "dhop_host" > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cg --no-config4 --no-config5 --no-cpu --e2e-steps 10 > $out/bench_n1.json 2> $out/bench.err; python - <<PY
import json
l=json.loads([x for x in open("$out/bench_n1.json").read().splitlines() if x.startswith("{")][-1])
print("ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "e2e", l["e2e"])
PY
