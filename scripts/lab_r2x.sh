#!/bin/bash
# GPU call X: compute-sanitizer (memcheck, synccheck, racecheck) on the current tuned kernels at a small lattice (16.8.8.8 x 16): plain hop,
# self-halo hop with hop-sent t faces, fused CG with hop epilogues.
set -u
out=gpurun_out/r2x; mkdir -p $out
cat > /tmp/san_workload.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import grid_b200 as gb
ctx = gb.Context(0)
dims, Ls = (16, 8, 8, 8), 16
grid = gb.GridCartesian(ctx, dims)
D = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, Ls, 0.1, 1.8, 1.5, 0.5)
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2); out = gb.LatticeFermion(grid, Ls, gb.F32)
so = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF); gb.pickCheckerboard(gb.Odd, so, src)
for dag in (0, 1):
    D.Dhop(src, out, dag)
x = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF).zero()
cg = gb.ConjugateGradient(1e-3, 40, err_on_no_conv=False)
cg(gb.SchurDiagMooeeOperator(D), so, x)
ctx.synchronize()
print("workload done", cg.IterationsToComplete, cg.TrueResidual, ctx.launch_count())
PY
for tool in memcheck synccheck racecheck; do
  for env in "LAB_X=1" "GB_SELF_HALO=12"; do
    tag=${tool}_$(echo $env | tr ' =' '__')
    env $env timeout 900 compute-sanitizer --tool $tool --print-limit 10 python /tmp/san_workload.py > $out/$tag.log 2>&1
    echo "$tag: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/$tag.log | tail -1) $(grep 'workload done' $out/$tag.log)"
  done
done
