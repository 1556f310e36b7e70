#!/bin/bash
# GPU call: cache hints of the column kernel (GB_COL_HINTS bit mask: 1 streaming result stores, 2 evict-first links, 4 evict-last ring planes).
set -u
out=gpurun_out/r3d; mkdir -p $out
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl | cut -c1-230; }
DIMS="32 32 32 32"
for h in 0 1 2 3 4 5 6 7; do lab GB_COL_HINTS=$h; done
lab GB_COL_HINTS=0
DIMS="64 64 32 16"
for h in 0 1 3 7; do lab GB_COL_HINTS=$h; done
GB_COL_HINTS=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fast_and_generic or tiling or dhop_full" 2>&1 | tail -2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for h in 0 3 7; do
GB_COL_HINTS=$h timeout 300 ncu --metrics $M --clock-control none -s 3 -c 1 --csv --log-file $out/ncu_h$h.csv python scripts/lab_dhop.py 32 32 32 32 16 3 ncu > /dev/null 2>&1
echo "hints $h: $(grep dhop_col2 $out/ncu_h$h.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | paste - - - -)"
done
