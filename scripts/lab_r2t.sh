#!/bin/bash
# GPU call T (2 GPUs): interleave ratio of the sender CTAs (hop-sent t faces).
set -u
out=gpurun_out/r2t; mkdir -p $out
for il in 0 1 2 3; do
  ( GB_SEND_INTERLEAVE=$il timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$il scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | sed "s/^{/{\"interleave_log2\": $il, /" | tee -a $out/hop_lab.jsonl | grep -v HOP_SENDS
done
