#!/bin/bash
# Round-2 kernel lab, first pass (one GPU): knobs that exist in the library but whose B200 numbers are not on record.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/run_r2_lab.sh'
# Hypothesis behind it: the column-sweep hop moves 4.88 GB of DRAM traffic per full 32^4 x 16 hop against
# 3.825 GB algorithmic.  Accounted sources: (a) every column CTA reads N + 2 z-planes for N outputs (N = 16: +12.5 % input
# reads; N = Lz = 32: the two extra planes are the column's own first / last plane), (b) the y-rows at the edge of a wave of
# 296 CTAs (32 t x 4 x-blocks x 2.3 y-blocks) are re-read by the next wave (+22 %).  GB_COL_N sweeps (a).
set -u
out=gpurun_out/r2_lab; mkdir -p $out
q() { python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --e2e-steps 1 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'sm_mhz', d['clocks']['sm_mhz'])"; }
for n in 8 16 32; do GB_COL_N=$n q "GB_COL_N=$n"; done | tee $out/col_n.txt
GB_NO_COL=1 q "micro-block kernel (GB_NO_COL=1)" | tee -a $out/col_n.txt
GB_COL_NT=2 q "two t-slices per CTA (GB_COL_NT=2)" | tee -a $out/col_n.txt
# DRAM traffic of the best and the default setting (ncu: one launch each; never a bench value)
for n in 16 32; do
  GB_COL_N=$n ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dhop_col -s 3 -c 1 --csv \
    --log-file $out/ncu_col_n$n.csv python scripts/prof_dhop.py > /dev/null 2>&1
  tail -3 $out/ncu_col_n$n.csv
done
