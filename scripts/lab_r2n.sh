#!/bin/bash
# GPU call N: the col2 fix (arrive depends on the z- loads): stress, reproducibility of the CG, speed.
set -u
out=gpurun_out/r2n; mkdir -p $out
run() { env "$@" timeout 300 python scripts/hop_stress.py 32 16 $N $OP 2>&1 | tail -1 | tee -a $out/stress.jsonl | cut -c1-300; }
N=10000; OP=DhopEO
run LAB_X=1
run GB_COL_N=32
N=3000; OP=Dhop
run LAB_X=1
N=1500; OP=HermOp
run LAB_X=1
env LAB_X=1 timeout 300 python scripts/cg_repro.py 32 16 8 2>&1 | tail -1 | tee -a $out/repro.jsonl | cut -c1-400
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab LAB_X=1
lab GB_COL2=0
lab GB_COL2_SYNC=1
lab GB_SELF_HALO=12
DIMS="64 64 32 16"
lab LAB_X=1
lab GB_SELF_HALO=12
timeout 300 python scripts/cg_bench.py 32 16 single 300 | tail -1 | tee -a $out/cg.jsonl
timeout 300 python scripts/cg_bench.py 32 16 mixed 300 | tail -1 | tee -a $out/cg.jsonl
