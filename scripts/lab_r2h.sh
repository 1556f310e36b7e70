#!/bin/bash
# GPU call H (2 GPUs): the bench line under torchrun (parity against the oracle, headline, cg, config4 incl. halo bandwidth, config5), mgpu_check.
set -u
out=gpurun_out/${LAB_OUT:-r2h}; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 ) > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench n2 rc $?"; tail -c 4000 $out/bench_n2.json; tail -5 $out/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_check.py ) > $out/mgpu_check.log 2>&1
echo "mgpu_check rc $?"; tail -5 $out/mgpu_check.log
