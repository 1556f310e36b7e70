#!/bin/bash
# GPU call E: rasterisation / column-height / cluster / L2-prefetch sweep of dhop_col2, then the multi-rank forms on one GPU.
set -u
out=gpurun_out/r2e; mkdir -p $out
lab() { env "$@" timeout 120 python scripts/lab_dhop.py $DIMS 16 100 "$*" 2>&1 | tail -1 | tee -a $out/lab.jsonl; }
DIMS="32 32 32 32"
lab GB_COL2=0
lab GB_COL_N=16
lab GB_COL_N=8
lab GB_COL_N=4
lab GB_COL_N=8 GB_COL_RASTER=2
lab GB_COL_N=4 GB_COL_RASTER=2
lab GB_COL_N=16 GB_COL_RASTER=2
lab GB_COL_N=16 GB_COL_TB=8
lab GB_COL_N=16 GB_COL_TB=16
lab GB_COL_N=8 GB_COL_TB=8 GB_COL_RASTER=2
lab GB_COL_N=16 GB_COL_CLUSTER=2
lab GB_COL_N=16 GB_COL_CLUSTER=4
lab GB_COL_N=16 GB_COL_L2PF=1
lab GB_COL_N=32 GB_COL_L2PF=1
lab GB_COL_N=16 GB_COL_RASTER=0
lab GB_SELF_HALO=8
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=8
DIMS="64 64 32 16"
lab GB_COL2=0
lab GB_COL_N=16
lab GB_COL_N=8 GB_COL_RASTER=2
lab GB_SELF_HALO=12
lab GB_SELF_HALO=12 GB_COL2_DECOMP=0
lab GB_SELF_HALO=12 GB_PACK_STREAM=1 GB_PACK_CTAS=16
