#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash scripts/run_r2_verify.sh'
# Outputs under gpurun_out/r2_verify/.  Nothing here is a bench value of record.
set -u
out=gpurun_out/r2_verify; mkdir -p $out
nvidia-smi -L > $out/gpus.txt 2>&1
# 1. the whole GPU suite; unverified tests run last and report as xfail / xpass (-rxX lists them by name)
python -m pytest tests -m gpu -q -rxX -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -40 $out/pytest_gpu.log
# 2. N-rank parity (Wilson / DWF / Moebius as before, plus the improved staggered operator with three-deep halos)
ngpu=$(nvidia-smi -L | wc -l)
if [ "$ngpu" -ge 2 ]; then
  n=2; [ "$ngpu" -ge 8 ] && n=8
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/mgpu_check.py > $out/mgpu_check_n2.log 2>&1
  grep -E "FAIL|MGPU_CHECK|staggered" $out/mgpu_check_n2.log | tail -40
  # 3. staggered Dhop on N GPUs: strong (global 48^4) and weak (48^4 per GPU)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 scripts/stag_bench_mgpu.py 48 100 > $out/stag_bench_n${n}_strong.jsonl 2>$out/stag_bench_n${n}_strong.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29545 scripts/stag_bench_mgpu.py 48 100 weak > $out/stag_bench_n${n}_weak.jsonl 2>$out/stag_bench_n${n}_weak.err
  cat $out/stag_bench_n${n}_strong.jsonl $out/stag_bench_n${n}_weak.jsonl
  # 4. halo microbenchmark: peer-to-peer stores, then the NCCL send/recv path
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 scripts/halo_bench.py 32 200 > $out/halo_bench_n${n}.jsonl 2>$out/halo_bench_n${n}.err
  GB_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29556 scripts/halo_bench.py 32 200 >> $out/halo_bench_n${n}.jsonl 2>>$out/halo_bench_n${n}.err
  cat $out/halo_bench_n${n}.jsonl
fi
python scripts/stag_bench.py 48 100 > $out/stag_bench_n1.jsonl 2>&1; cat $out/stag_bench_n1.jsonl
# 5. BASELINE configs[3] weak-scaling series at ITS local volume (64.64.32.16 x Ls16 per GPU): 1 GPU and N GPUs
python bench.py --gpus 1 --local 64 64 32 16 --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > $out/bench_c4_n1.json 2>$out/bench_c4_n1.err
if [ "$ngpu" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus $n --local 64 64 32 16 --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 > $out/bench_c4_n${n}.json 2>$out/bench_c4_n${n}.err
fi
grep -o '"ms_per_step": [0-9.]*' $out/bench_c4_n*.json

# 6. Benchmark_usqcd-shaped table (Wilson / DWF4 / staggered DhopEO fp32 at local L^4, stream triad) next to BASELINE.md's Booster rows
python scripts/benchmark_usqcd.py > $out/usqcd_n1.jsonl 2>$out/usqcd_n1.err; cat $out/usqcd_n1.jsonl
