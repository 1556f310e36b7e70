#!/bin/bash
# Round-2 GPU call A (one GPU): the whole -m gpu suite (markers gone, new self-halo and full-size reference tests), then the
# kernel lab of the second-generation column-sweep hop.  Outputs under gpurun_out/r2a/.  Nothing here is a bench value of record.
set -u
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
free -g > $out/host_mem.txt; nproc >> $out/host_mem.txt
( time python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=15 ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -30 $out/pytest_gpu.log
lab() { env "$@" python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>/dev/null | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab GB_COL2=0                       # round-1 kernel (cp.async ring, N = 16)
lab GB_COL2=1                       # TMA ring, N = Lz
lab GB_COL_N=16
lab GB_COL_RASTER=1
lab GB_COL_N=16 GB_COL_RASTER=1
lab GB_NO_COL=1                     # micro-block kernel
# the multi-rank forms on one GPU (halos to self): kernel-structure cost of a decomposed hop, no NVLink in it
lab GB_SELF_HALO=8
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=8
lab GB_SELF_HALO=12 LAB_OVERLAP=2
DIMS="64 64 32 16"                  # per-GPU volume of BASELINE configs[3]
lab GB_COL2=0
lab GB_COL2=1
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=8
# DRAM traffic of the old and the new kernel (ncu: one launch each; never a bench value)
for v in 0 1; do
  GB_COL2=$v ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:dhop_col -s 3 -c 1 --csv --log-file $out/ncu_col2_$v.csv python scripts/prof_dhop.py > /dev/null 2>&1
  tail -8 $out/ncu_col2_$v.csv | cut -c1-300
done
