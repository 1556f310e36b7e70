#!/bin/bash
# GPU call D: lab of the col2 variants (raster, deep prefetch), multi-rank forms on one GPU (self halo), DRAM bytes.
set -u
out=gpurun_out/r2d; mkdir -p $out
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab GB_COL2=0
lab GB_COL2_DEEP=0 GB_COL_RASTER=0
lab GB_COL2_DEEP=0 GB_COL_RASTER=1
lab GB_COL2_DEEP=1 GB_COL_RASTER=1
lab GB_COL2_DEEP=1 GB_COL_RASTER=0
lab GB_COL2_DEEP=1 GB_COL_RASTER=1 GB_COL_N=16
lab GB_COL2_DEEP=0 GB_COL_RASTER=1 GB_COL_N=16
lab GB_SELF_HALO=8
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=8
DIMS="64 64 32 16"
lab GB_COL2=0
lab GB_COL2_DEEP=1
lab GB_COL2_DEEP=0
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=8
for v in "GB_COL2_DEEP=0 GB_COL_RASTER=1" "GB_COL2_DEEP=1 GB_COL_RASTER=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
    --clock-control none -k regex:dhop_col2 -s 3 -c 1 --csv --log-file $out/ncu_$tag.csv python scripts/prof_dhop.py > /dev/null 2>&1
  grep -E "dram__|duration|hit_rate|wavefronts|long_score" $out/ncu_$tag.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
