#!/bin/bash
# Final GPU call of the round: the whole -m gpu suite, smoke, the default bench invocation, and an ncu pass over the generic kernel
# with the full and the two-row link store (4D Wilson fp32, 32^4).
set -u
out=gpurun_out/r4g; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $out/pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 600 python bench.py ) > $out/bench_default.json 2> $out/bench_default.err
echo "bench default rc $?"; tail -3 $out/bench_default.err | cut -c1-200
cat > /tmp/wl.py <<'P'
import sys; sys.path.insert(0, ".")
import grid_b200 as gb
ctx = gb.Context(0); grid = gb.GridCartesian(ctx, (32, 32, 32, 32))
D = gb.WilsonFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, 0.1)
src = gb.LatticeFermion(grid, 1, gb.F32).random(2); out = gb.LatticeFermion(grid, 1, gb.F32)
for nreal in (18, 12):
    D.set_link_reconstruct(nreal)
    for _ in range(4): D.Dhop(src, out, 0)
ctx.synchronize()
P
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:dhop_kernel --csv --log-file $out/ncu_wilson_recon12.csv python /tmp/wl.py > $out/ncu_wilson.log 2>&1
echo "ncu rc $?"; tail -9 $out/ncu_wilson_recon12.csv | cut -c1-220
