#!/bin/bash
# GPU call J: streaming tridiagonal CG passes (parity, ms per iteration), the reference's unmodified programs through the bridge, PCIe rates.
set -u
out=gpurun_out/r2j; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cg_fused.py tests/test_gpu_bridge.py tests/test_gpu_full_size.py -m gpu -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc $?"; tail -5 $out/pytest.log
for m in single mixed; do
  timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
  GB_NO_STRI=1 timeout 300 python scripts/cg_bench.py 32 16 $m 300 | tail -1 | tee -a $out/cg.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 400 -c 60 --csv --log-file $out/ncu_cg_launches.csv python scripts/cg_bench.py 32 16 single 60 > /dev/null 2>&1
echo "ncu rc $?"
timeout 300 python scripts/pcie_bw.py | tee $out/pcie_bw.json
( cd bridge/_build && timeout 600 ./Benchmark_dwf_fp32 --grid 16.16.16.16 -Ls 16 > ../../$out/bridge_benchmark_dwf_fp32_16.log 2>&1; echo "bridge bench rc $?"; grep -E "mflop/s =|norm (dag )?diff" ../../$out/bridge_benchmark_dwf_fp32_16.log )
