"""Development timing of the fp32 DWF hopping term under the library's environment switches (not the contract bench).
usage: [GB_*=...] python scripts/lab_dhop.py Lx Ly Lz Lt [Ls] [ncall] [tag]      prints one JSON line
GB_SELF_HALO=<mask> times the multi-rank hop forms on one GPU (halos to self), LAB_OVERLAP=0/1/2 picks the form."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb

dims = tuple(int(x) for x in sys.argv[1:5])
Ls = int(sys.argv[5]) if len(sys.argv) > 5 else 16
ncall = int(sys.argv[6]) if len(sys.argv) > 6 else 100
tag = sys.argv[7] if len(sys.argv) > 7 else ""
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, dims)
U = gb.LatticeGaugeField(grid, gb.F32).random(1)
D = gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8)
if "LAB_OVERLAP" in os.environ:
    D.set_overlap(int(os.environ["LAB_OVERLAP"]))
src, out = gb.LatticeFermion(grid, Ls, gb.F32).random(2), gb.LatticeFermion(grid, Ls, gb.F32)
se, ro = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
gb.pickCheckerboard(gb.Odd, se, src)
vol5 = grid.lsites * Ls
res = {"tag": tag, "dims": dims, "Ls": Ls, "env": {k: v for k, v in os.environ.items() if k.startswith("GB_") or k.startswith("LAB_")}}
for name, fn, vol in (("Dhop", lambda: D.Dhop(src, out, 0), vol5), ("DhopEO", lambda: D.DhopEO(se, ro, 0), vol5 // 2)):
    for _ in range(5):
        fn()
    ctx.synchronize()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(ncall):
        fn()
    ms = ctx.timer_stop() / ncall
    res[name] = {"ms": round(ms, 4), "frac_of_6546": round(228.0 * vol / ms / 1e6 / 6546.6, 4), "launches_per_call": (ctx.launch_count() - l0) / ncall}
print(json.dumps(res), flush=True)
