#!/bin/bash
# GPU call R: t faces sent by the hop kernel itself -- self-halo parity, reproducibility under the halo path, timings against the pack-kernel form.
set -u
out=gpurun_out/r2r; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_self_halo.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -4 $out/pytest.log
run() { env "$@" timeout 300 python scripts/hop_stress.py 32 16 $N $OP 2>&1 | tail -1 | tee -a $out/stress.jsonl | cut -c1-300; }
N=3000; OP=DhopEO
run GB_SELF_HALO=8
run GB_SELF_HALO=12
N=1000; OP=Dhop
run GB_SELF_HALO=12
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
for DIMS in "32 32 32 32" "64 64 32 16"; do
  lab LAB_X=1
  lab GB_SELF_HALO=8
  lab GB_SELF_HALO=8 GB_HOP_SENDS_T=0
  lab GB_SELF_HALO=12
  lab GB_SELF_HALO=12 GB_HOP_SENDS_T=0
done
