#!/bin/bash
# Kernel-development loop WITHOUT a GPU: builds the CPU mock of the library (tests/mock/README.md) from the current sources and runs
# the tuned-kernel checks on it -- edge shapes, fast-vs-generic agreement, tiling invariance, the Schur operator, then (with "all")
# the whole parity file and the N-rank run with peer-to-peer halos and the semi-fused hop.  A variant of dhop_fast.cuh / dhop_col.cuh /
# smat.cu that breaks index arithmetic, shared-memory layout or the mbarrier protocol fails (or is reported as a deadlock) here,
# before it costs GPU minutes.  Says nothing about performance or intra-block memory ordering.
# usage: scripts/mock_check.sh [all] [extra environment, e.g. GB_COL_N=8 GB_NO_COL=1 GB_COL_NT=2 GB_MOCK_SM_COUNT=3 are read from the environment]
set -e
cd "$(dirname "$0")/.."
OUT=${GB_MOCK_DIR:-/tmp/gridb200_mock_check}
LIB=$(python tests/mock/build_mock.py "$OUT")
export GB_TEST_MOCK_LIB="$LIB" GB_MOCK_COUNT="dhop_col2_kernel;dhop_col_kernel;dhop_fast_kernel;smat_kernel;pack_send_kernel"
python tests/mock/run_counted.py tests/test_next_tuned_shapes.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x \
  -k "edge_shapes or fast_and_generic or tiling or schur_operator or dhop_full or dhop_oe_eo"
if [ "$1" = "all" ]; then
  python tests/mock/run_counted.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -p no:cacheprovider -x -k "not device_random"
  python tests/mock/run_counted.py tests/test_gpu_recon12.py tests/test_gpu_self_halo.py -m gpu -q -p no:cacheprovider -x -k "recon12 or host_dhop or (compressed_halos and zt)"
  python tests/mock/mgpu_on_mock.py "$LIB"
fi
