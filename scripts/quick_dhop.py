"""Development timing of the hopping kernel (not the contract bench): python scripts/quick_dhop.py [L] [Ls] [prec]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 16
prec = gb.F32 if (len(sys.argv) <= 3 or sys.argv[3] == "f32") else gb.F64
tilings = [(0, 8, 0), (0, 0, 0), (0, 4, 0), (0, 16, 0), (8, 8, 0), (0, 8, 8), (0, 2, 0), (4, 4, 0), (0, 4, 4)]
ctx = gb.Context(0)
dims = (L, L, L, L)
grid = gb.GridCartesian(ctx, dims)
U = gb.LatticeGaugeField(grid, prec).random(1)
D = gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8) if Ls > 1 else gb.WilsonFermion(U, grid, 0.1)
src = gb.LatticeFermion(grid, Ls, prec).random(2)
out = gb.LatticeFermion(grid, Ls, prec)
se, ro = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
gb.pickCheckerboard(gb.Odd, se, src)
vol5 = L ** 4 * Ls
w = 4 if prec == gb.F32 else 8
bytes_site = 2 * 24 * w + 8 * 18 * w / Ls
for t in tilings:
    D.set_tiling(*t)
    for name, fn, vol in (("Dhop", lambda: D.Dhop(src, out, 0), vol5), ("DhopEO", lambda: D.DhopEO(se, ro, 0), vol5 // 2)):
        for _ in range(3):
            fn()
        n = 20
        ctx.timer_start()
        for _ in range(n):
            fn()
        ms = ctx.timer_stop() / n
        print(json.dumps(dict(op=name, tiling=t, ms=round(ms, 4), gflops=round(1320 * vol / ms / 1e6, 1),
                              alg_GBs=round(bytes_site * vol / ms / 1e6, 1))), flush=True)
