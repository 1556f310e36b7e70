#!/bin/bash
# GPU call P: hop epilogues of the Schur CG -- parity, reproducibility, ms per iteration against the streaming-pass form.
set -u
out=gpurun_out/r2p; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cg_fused.py tests/test_gpu_stress.py tests/test_gpu_full_size.py -m gpu -x -q -p no:cacheprovider -k "fused or conjugate or mixed or cg" > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -5 $out/pytest.log
for m in single mixed; do
  timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
  GB_NO_HOP_EPI=1 timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
done
env LAB_X=1 timeout 300 python scripts/cg_repro.py 32 16 5 2>&1 | tail -1 | tee -a $out/repro.jsonl | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 60 --csv --log-file $out/ncu_cg_launches.csv python scripts/cg_bench.py 32 16 single 60 > /dev/null 2>&1
echo "ncu rc $?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-config5 --no-cpu > $out/bench_n1.json 2> $out/bench.err; tail -c 1800 $out/bench_n1.json
