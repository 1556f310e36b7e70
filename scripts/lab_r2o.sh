#!/bin/bash
# GPU call O: the whole -m gpu suite, smoke, the bench line, the ncu launch list of the bench command, ncu --set full of the fixed col2 kernel
# (full hop and checkerboard hop) and of the fp64 generic kernel.
set -u
out=gpurun_out/r2o; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $out/pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py ) > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench rc $?"; tail -c 600 $out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-config4 --no-config5 > $out/bench_under_ncu.log 2>&1
echo "ncu launches rc $?"
N=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dhop_col2 -s 1 -c 1 -o $out/col2_fullhop python scripts/prof_dhop.py > $out/ncu_full1.log 2>&1
N=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dhop_col2 -s 4 -c 1 -o $out/col2_cbhop python scripts/prof_dhop.py > $out/ncu_full2.log 2>&1
PREC=f64 N=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dhop_kernel -s 3 -c 1 -o $out/generic_fp64_cbhop python scripts/prof_dhop.py > $out/ncu_full3.log 2>&1
ls -la $out
