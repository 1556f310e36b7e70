#!/bin/bash
set -u
out=gpurun_out/r3i; mkdir -p $out
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu ) > $out/bench_n1.json 2> $out/bench.err
echo "rc $?"; python - <<PY
import json
l=json.loads([x for x in open("$out/bench_n1.json").read().splitlines() if x.startswith("{")][-1])
print("ms", l["ms_per_step"], l["roofline"]["frac"], "c4", l["config4"]["ms_per_step"], "strong", l["config4"].get("strong"), "c5", l["config5"])
PY
tail -3 $out/bench.err; nvidia-smi --query-gpu=memory.used --format=csv
