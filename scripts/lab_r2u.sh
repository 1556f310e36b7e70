#!/bin/bash
# GPU call U: why is the t-decomposed launch slower than the single-rank one?  ncu of both on one GPU (self halo), 64.64.32.16 x 16.
set -u
out=gpurun_out/r2u; mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size
for env in "LAB_X=1" "GB_SELF_HALO=8" "GB_SELF_HALO=8 GB_HOP_SENDS_T=0" "GB_SELF_HALO=12"; do
  tag=$(echo $env | tr ' =' '__')
  env $env timeout 300 ncu --metrics $M --clock-control none -s 4 -c 6 --csv --log-file $out/ncu_$tag.csv python scripts/lab_dhop.py 64 64 32 16 16 3 ncu > /dev/null 2>&1
  echo "== $env"; python - <<PY
import csv
rows=list(csv.reader(open("$out/ncu_$tag.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ix={k:i for i,k in enumerate(h)}
cur=None
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    key=(r[ix['ID']], r[ix['Kernel Name']][:48])
    if key!=cur: print(); print(key[0], key[1], end=' | '); cur=key
    print(r[ix['Metric Name']].split('.')[0][-22:], r[ix['Metric Value']], r[ix['Metric Unit']], end=' ; ')
print()
PY
done
