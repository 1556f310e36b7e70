#!/bin/bash
# GPU call Y: z-surface planes inside the column kernel: self-halo parity + stress, timings (one GPU), then sanitizers.
set -u
out=gpurun_out/r2y; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_self_halo.py tests/test_gpu_cg_fused.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -2 $out/pytest.log
GB_SELF_HALO=12 timeout 300 python scripts/hop_stress.py 32 16 2000 DhopEO 2>&1 | tail -1 | cut -c1-200
GB_SELF_HALO=4 timeout 300 python scripts/hop_stress.py 32 16 1000 Dhop 2>&1 | tail -1 | cut -c1-200
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl | cut -c1-260; }
for DIMS in "32 32 32 32" "64 64 32 16"; do
  lab LAB_X=1
  lab GB_SELF_HALO=4
  lab GB_SELF_HALO=12
  lab GB_SELF_HALO=12 GB_COL2_ZPLANES=0
done
bash scripts/lab_r2x.sh
