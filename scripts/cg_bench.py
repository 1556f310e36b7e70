"""CG time-to-solution at BASELINE config 3 shape: even-odd Schur Moebius DWF, mixed-precision CG to 1e-8 on 32^4 x Ls16.
usage: python scripts/cg_bench.py [L] [Ls] [mixed|double|single] [maxit]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mode = sys.argv[3] if len(sys.argv) > 3 else "mixed"
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L,) * 4)
Ud = gb.LatticeGaugeField(grid, gb.F64).random(1)
Uf = gb.LatticeGaugeField(grid, gb.F32).import_lex(Ud.export_lex()) if L <= 32 else None
mass, M5, b, c = 0.1, 1.8, 1.5, 0.5
Dd = gb.MobiusFermion(Ud, grid, Ls, mass, M5, b, c)
Df = gb.MobiusFermion(Uf, grid, Ls, mass, M5, b, c)
src = gb.LatticeFermion(grid, Ls, gb.F64).random(2)
so = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
gb.pickCheckerboard(gb.Odd, so, src)
sol = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
sol.set_checkerboard(gb.Odd)
Ld, Lf = gb.SchurDiagMooeeOperator(Dd), gb.SchurDiagMooeeOperator(Df)
vol_cb = L ** 4 * Ls // 2
flops_it = (1452.0 * 4 + (8 + 4 + 8 + 4 + 4) * 12) * vol_cb   # ref: Test_dwf_mixedcg_prec.cc:138-142,171-172
ctx.synchronize()
l0 = ctx.launch_count()
t0 = time.perf_counter()
if mode == "mixed":
    s = gb.MixedPrecisionConjugateGradient(1e-8, maxit, 50, Lf, Ld)
    s(so, sol)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    its = s.TotalInnerIterations + s.TotalFinalStepIterations
    out = dict(mode=mode, inner=s.TotalInnerIterations, outer=s.TotalOuterIterations, final=s.TotalFinalStepIterations, true_resid=s.TrueResidual)
elif mode == "double":
    s = gb.ConjugateGradient(1e-8, maxit, err_on_no_conv=False)
    s(Ld, so, sol)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    its = s.IterationsToComplete
    out = dict(mode=mode, iterations=its, true_resid=s.TrueResidual)
else:
    sf, xf = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF).zero()
    gb.precisionChange(sf, so)
    s = gb.ConjugateGradient(1e-5, maxit, err_on_no_conv=False)
    ctx.synchronize(); t0 = time.perf_counter()
    s(Lf, sf, xf)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    its = s.IterationsToComplete
    out = dict(mode=mode, iterations=its, true_resid=s.TrueResidual)
out.update(L=L, Ls=Ls, seconds=round(dt, 4), ms_per_iteration=round(1e3 * dt / max(its, 1), 4), gflops=round(flops_it * its / dt / 1e9, 1),
           launches=ctx.launch_count() - l0)
print(json.dumps(out), flush=True)
