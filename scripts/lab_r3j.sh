#!/bin/bash
set -u
out=gpurun_out/r3j; mkdir -p $out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu ) > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench n2 rc $?"; python - <<PY
import json
l=json.loads([x for x in open("$out/bench_n2.json").read().splitlines() if x.startswith("{")][-1])
print(l["ms_per_step"], l["config4"]["ms_per_step"], l["config4"].get("strong"), l["parity_check"]["ok"])
PY
tail -2 $out/bench_n2.err
