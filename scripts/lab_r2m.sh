#!/bin/bash
# GPU call M: hop stress (same input, 4000 calls, every output compared with the first) for the col2 kernel and its switches.
set -u
out=gpurun_out/r2m; mkdir -p $out
run() { env "$@" timeout 300 python scripts/hop_stress.py 32 16 $N $OP 2>&1 | tail -1 | tee -a $out/stress.jsonl | cut -c1-300; }
N=4000; OP=DhopEO
run LAB_X=1
run GB_COL2_SYNC=1
run GB_COL_N=32
run GB_COL_N=8
run GB_COL_RASTER=0
run GB_COL2=0
N=2000; OP=Dhop
run LAB_X=1
N=1000; OP=HermOp
run LAB_X=1
run GB_COL2=0
N=4000; OP=smat
run LAB_X=1
