#!/bin/bash
# GPU call: flag-in-data t halos (GB_T_LL): self-halo parity + stress, timings against the flag form, ncu traffic.
set -u
out=gpurun_out/r3m; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_self_halo.py tests/test_gpu_stress.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
GB_SELF_HALO=12 timeout 300 python scripts/hop_stress.py 32 16 2000 DhopEO 2>&1 | tail -1 | cut -c1-200
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl | cut -c1-250; }
for DIMS in "32 32 32 32" "64 64 32 16"; do
  lab LAB_X=1
  lab GB_SELF_HALO=8
  lab GB_SELF_HALO=8 GB_T_LL=0
  lab GB_SELF_HALO=12
  lab GB_SELF_HALO=12 GB_T_LL=0
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for env in "GB_SELF_HALO=8" "GB_SELF_HALO=8 GB_T_LL=0"; do
  tag=$(echo $env | tr ' =' '__')
  env $env timeout 300 ncu --metrics $M --clock-control none -s 5 -c 2 --csv --log-file $out/ncu_$tag.csv python scripts/lab_dhop.py 64 64 32 16 16 3 ncu > /dev/null 2>&1
  echo "== $env"; grep -E "dhop|pack" $out/ncu_$tag.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tr -d '"' | paste - - - - | cut -c1-330
done
