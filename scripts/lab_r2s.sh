#!/bin/bash
# GPU call S (2 GPUs): hop-sent t faces on real NVLink: parity (mgpu_check), timings of both forms.
set -u
out=gpurun_out/r2s; mkdir -p $out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/mgpu_check.py ) > $out/mgpu_check.log 2>&1
echo "mgpu_check rc $?"; tail -2 $out/mgpu_check.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | tee $out/hop_lab.jsonl
