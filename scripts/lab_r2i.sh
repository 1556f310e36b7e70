#!/bin/bash
# GPU call I: fused Schur CG -- parity, then ms per iteration fused vs unfused at 32^4 x 16, launch list of one fused iteration.
set -u
out=gpurun_out/r2i; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cg_fused.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_full_size.py tests/test_next_multishift.py tests/test_next_relupcg.py tests/test_next_schur_solve.py -m gpu -x -q -p no:cacheprovider -k "cg or CG or solve or mixed or multishift or relup" > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -3 $out/pytest.log
for m in single mixed; do
  timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
  GB_CG_UNFUSED=1 timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 60 --csv --log-file $out/ncu_cg_launches.csv python scripts/cg_bench.py 32 16 single 60 > /dev/null 2>&1
echo "ncu rc $?"
