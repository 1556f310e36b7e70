#!/bin/bash
# GPU call W (8 GPUs): the bench line as the driver runs it (parity against the oracle, headline, cg, config4 = global 64^4 x 16, config5), N-rank parity.
set -u
out=gpurun_out/r2w; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 ) > $out/bench_n8.json 2> $out/bench_n8.err
echo "bench n8 rc $?"; tail -c 1500 $out/bench_n8.json; tail -3 $out/bench_n8.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/mgpu_check.py ) > $out/mgpu_check.log 2>&1
echo "mgpu_check rc $?"; tail -3 $out/mgpu_check.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | tee $out/hop_lab.jsonl
