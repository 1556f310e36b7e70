import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb
L, Ls = 32, 16
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L,) * 4)
Uf = gb.LatticeGaugeField(grid, gb.F32).random(1)
Df = gb.MobiusFermion(Uf, grid, Ls, 0.1, 1.8, 1.5, 0.5)
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2)
so = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
gb.pickCheckerboard(gb.Odd, so, src)
Lf = gb.SchurDiagMooeeOperator(Df)
for rep in range(3):
    x = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF).zero()
    s = gb.ConjugateGradient(1e-5, 1000, err_on_no_conv=False)
    ctx.synchronize(); t0 = time.perf_counter()
    s(Lf, so, x)
    ctx.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps(dict(rep=rep, iters=s.IterationsToComplete, seconds=round(dt, 4), ms_per_it=round(1e3 * dt / s.IterationsToComplete, 3))), flush=True)
