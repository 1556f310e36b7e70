"""Two-row (12-real) link storage against the full store, both through the generic hopping kernel: 4D Wilson (links dominate the
traffic) and DWF Ls 8 / 16, fp32 and fp64, one GPU.  usage: python scripts/recon12_lab.py [L]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L, L, L, L))
for prec, pname in ((gb.F32, "fp32"), (gb.F64, "fp64")):
    U = gb.LatticeGaugeField(grid, prec).random(1)
    for kind, Ls in (("wilson", 1), ("dwf", 8), ("dwf", 16)):
        if prec == gb.F64 and Ls == 16 and L > 24:
            continue
        D = gb.WilsonFermion(U, grid, 0.1) if kind == "wilson" else gb.DomainWallFermion(U, grid, Ls, 0.1, 1.8)
        D.set_fast_kernel(0)                      # generic kernel in both cases: the comparison is the link storage alone
        src = gb.LatticeFermion(grid, Ls, prec).random(2); out = gb.LatticeFermion(grid, Ls, prec)
        res = {}
        for nreal in (18, 12):
            D.set_link_reconstruct(nreal)
            for _ in range(5):
                D.Dhop(src, out, 0)
            ctx.synchronize(); ctx.timer_start()
            n = 20
            for _ in range(n):
                D.Dhop(src, out, 0)
            res[nreal] = ctx.timer_stop() / n
        print(json.dumps({"op": kind, "L": L, "Ls": Ls, "prec": pname, "ms_18": round(res[18], 4), "ms_12": round(res[12], 4), "speedup": round(res[18] / res[12], 3)}), flush=True)
        del D
