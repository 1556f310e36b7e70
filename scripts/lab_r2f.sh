#!/bin/bash
# GPU call F: third-generation column kernel (dhop_col3): parity on the tuned-shape / parity tests, then timing against col2.
set -u
out=gpurun_out/r2f; mkdir -p $out
for v in 1 2; do
  GB_COL3=$v timeout 600 python -m pytest tests/test_next_tuned_shapes.py tests/test_gpu_parity.py tests/test_gpu_self_halo.py -m gpu -x -q -p no:cacheprovider \
    -k "edge_shapes or fast_and_generic or tiling or schur_operator or dhop_full or dhop_oe_eo or self_halo" > $out/pytest_col3_$v.log 2>&1
  echo "col3=$v pytest rc $?" | tee -a $out/summary.txt; tail -3 $out/pytest_col3_$v.log
done
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab GB_COL3=0
lab GB_COL3=1
lab GB_COL3=2
lab GB_COL3=1 GB_COL_L2PF=1
lab GB_COL3=2 GB_COL_L2PF=1
lab GB_COL3=1 GB_COL_N=32
lab GB_COL3=2 GB_COL_N=32
lab GB_COL3=2 GB_COL_N=32 GB_COL_L2PF=1
lab GB_COL3=2 GB_COL_N=8
lab GB_COL3=2 GB_COL_RASTER=0
lab GB_COL3=2 GB_SELF_HALO=8
lab GB_COL3=2 GB_SELF_HALO=12
DIMS="64 64 32 16"
lab GB_COL3=0
lab GB_COL3=2
lab GB_COL3=2 GB_COL_L2PF=1
# one ncu pass of the better variant: DRAM bytes, stalls
for v in 1 2; do
GB_COL3=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
  --clock-control none -k regex:dhop_col3 -s 3 -c 1 --csv --log-file $out/ncu_col3_$v.csv python scripts/lab_dhop.py 32 32 32 32 16 5 ncu > /dev/null 2>&1
done
GB_COL3=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dhop_col3 -s 3 -c 1 -o $out/col3_full python scripts/lab_dhop.py 32 32 32 32 16 5 ncu > $out/ncu_full.log 2>&1
