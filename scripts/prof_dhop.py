"""Workload for ncu captures: a handful of DWF Dhop / DhopEO calls at BASELINE config 1 shape (32^4 x 16 fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb
L = int(os.environ.get("L", 32)); Ls = int(os.environ.get("LS", 16)); n = int(os.environ.get("N", 4))
prec = gb.F32 if os.environ.get("PREC", "f32") == "f32" else gb.F64
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L, L, L, L))
U = gb.LatticeGaugeField(grid, prec).random(1)
D = gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8)
src = gb.LatticeFermion(grid, Ls, prec).random(2)
out = gb.LatticeFermion(grid, Ls, prec)
se, ro = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
gb.pickCheckerboard(gb.Odd, se, src)
for _ in range(n):
    D.Dhop(src, out, 0)
for _ in range(n):
    D.DhopEO(se, ro, 0)
ctx.synchronize()
