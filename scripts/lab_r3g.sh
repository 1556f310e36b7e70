#!/bin/bash
# GPU call: tail split of the column kernel (GB_COL_TAIL): parity, stress, timings, CG.
set -u
out=gpurun_out/r3g; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_next_tuned_shapes.py tests/test_gpu_full_size.py tests/test_gpu_stress.py tests/test_gpu_cg_fused.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl | cut -c1-250; }
DIMS="32 32 32 32"
lab GB_COL_TAIL=1
lab GB_COL_TAIL=0
lab GB_COL_TAIL=1
lab GB_COL_TAIL=0
DIMS="64 64 32 16"
lab GB_COL_TAIL=1
lab GB_COL_TAIL=0
DIMS="48 48 48 48"
lab GB_COL_TAIL=1
lab GB_COL_TAIL=0
for tl in 1 0; do GB_COL_TAIL=$tl timeout 300 python scripts/cg_bench.py 32 16 mixed 300 | tail -1 | tee -a $out/cg.jsonl; done
for tl in 1 0; do GB_COL_TAIL=$tl timeout 300 python bench.py --steps 20 --warmup 5 --no-config4 --no-config5 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]); print('tail $tl', l['ms_per_step'], l['roofline']['frac'], l['cg']['ms_per_iteration'], l['cg']['time_to_solution_s'])"; done
