#!/bin/bash
# GPU call: host-pipeline reorder (e2e), stress tests incl. the decomposed forms, the bench line.
set -u
out=gpurun_out/r3b; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "stress or reproducible or dhop_host or conjugate" > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -3 $out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench.err; python - <<PY
import json
l=json.loads([x for x in open("$out/bench_n1.json").read().splitlines() if x.startswith("{")][-1])
print("ms", l["ms_per_step"], "frac", l["roofline"]["frac"], "e2e", l["e2e"]["ms_per_step"], l["e2e"]["value"], "cpu", l["cpu_baseline"]["value"])
print("cg", l["cg"]["ms_per_iteration"], l["cg"]["time_to_solution_s"], "e2e_cg", l["e2e_cg"]["time_to_solution_s"], "c4", l["config4"]["ms_per_step"], l["config4"]["cg"]["ms_per_iteration"], "c5", l["config5"]["ms_per_step"])
PY
