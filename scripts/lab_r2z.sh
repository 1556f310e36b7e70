#!/bin/bash
# GPU call Z (4 GPUs, 1.1.2.2): z+t split on real NVLink: N-rank parity, hop timings (default / box-launch z planes / pack-sent t), bench line.
set -u
out=gpurun_out/r2z; mkdir -p $out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_check.py ) > $out/mgpu_check.log 2>&1
echo "mgpu_check rc $?"; grep -c "^ok" $out/mgpu_check.log; grep -E "FAIL|PASS" $out/mgpu_check.log | tail -3
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | tee $out/hop_lab.jsonl
( GB_COL2_ZPLANES=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/mgpu_hop_lab.py ) 2>&1 | grep "^{" | sed 's/^{/{"GB_COL2_ZPLANES": 0, /' | tee -a $out/hop_lab.jsonl
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 ) > $out/bench_n4.json 2> $out/bench_n4.err
echo "bench n4 rc $?"; tail -c 600 $out/bench_n4.json
