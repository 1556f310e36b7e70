"""Reproducibility of the fp32 Schur CG at L^4 x Ls: n solves from a zero guess, (iterations, true residual, sha1 of the solution) each.
usage: python scripts/cg_repro.py L Ls n [tol]   (kernel / solver forms through the GB_* environment switches)"""
import os, sys, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb
L, Ls, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-5
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L,) * 4)
Df = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, Ls, 0.1, 1.8, 1.5, 0.5)
if os.environ.get("LAB_GENERIC"):
    Df.set_fast_kernel(0)
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2)
so = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF); gb.pickCheckerboard(gb.Odd, so, src)
Lf = gb.SchurDiagMooeeOperator(Df)
rows = []
for i in range(n):
    x = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF).zero()
    cg = gb.ConjugateGradient(tol, 10000, err_on_no_conv=False)
    cg(Lf, so, x)
    rows.append((cg.IterationsToComplete, cg.TrueResidual, hashlib.sha1(x.export_lex().tobytes()).hexdigest()[:10]))
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith(("GB_", "LAB_")))
print(json.dumps({"tag": tag, "L": L, "distinct_solutions": len(set(r[2] for r in rows)), "true_resid_min_max": [min(r[1] for r in rows), max(r[1] for r in rows)], "rows": rows}), flush=True)
