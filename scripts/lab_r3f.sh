#!/bin/bash
# GPU call (8 GPUs): the bench line as the driver runs it, with the final code of the round.
set -u
out=gpurun_out/r3f; mkdir -p $out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 ) > $out/bench_n8.json 2> $out/bench_n8.err
echo "bench n8 rc $?"; tail -3 $out/bench_n8.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | tee $out/hop_lab.jsonl
