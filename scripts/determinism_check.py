"""Run-to-run reproducibility on one GPU: the same call twice must give the same bits (fixed-order reductions, no atomics on data).
usage: python scripts/determinism_check.py [L] [Ls]"""
import os, sys, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import grid_b200 as gb
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Ls = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L,) * 4)
Ud = gb.LatticeGaugeField(grid, gb.F64).random(1)
Uf = gb.LatticeGaugeField(grid, gb.F32).random(1)
Dd = gb.MobiusFermion(Ud, grid, Ls, 0.1, 1.8, 1.5, 0.5)
Df = gb.MobiusFermion(Uf, grid, Ls, 0.1, 1.8, 1.5, 0.5)
def digest(f):
    return hashlib.sha1(f.export_lex().tobytes()).hexdigest()[:12]
out = {}
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2)
res = gb.LatticeFermion(grid, Ls, gb.F32)
h = []
for i in range(4):
    Df.Dhop(src, res, 0); h.append(digest(res))
out["Dhop"] = h
so = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF); gb.pickCheckerboard(gb.Odd, so, src)
re_ = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
h = []
for i in range(4):
    Df.DhopEO(so, re_, 0); h.append(digest(re_))
out["DhopEO"] = h
Lf, Ld = gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd)
h = []
for i in range(3):
    Lf.HermOp(so, re_) if hasattr(Lf, "HermOp") else None
    h.append(digest(re_))
out["HermOp"] = h
for mode, env in (("stri", {}), ("dense", {"GB_NO_STRI": "1"}), ("unfused", {"GB_CG_UNFUSED": "1"})):
    os.environ.update(env)
    rows = []
    for i in range(3):
        x = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF).zero()
        cg = gb.ConjugateGradient(1e-5, 10000, err_on_no_conv=False)
        cg(Lf, so, x)
        rows.append([cg.IterationsToComplete, cg.TrueResidual, digest(x)])
    out["cg_" + mode] = rows
    sd = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF); gb.precisionChange(sd, so)
    rows = []
    for i in range(3):
        xd = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
        m = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, Lf, Ld)
        m(sd, xd)
        rows.append([m.TotalInnerIterations, m.TotalOuterIterations, m.TotalFinalStepIterations, m.TrueResidual, digest(xd)])
    out["mixed_" + mode] = rows
    for k in env:
        os.environ.pop(k)
for k, v in out.items():
    print(k, json.dumps(v), "REPRODUCIBLE" if all(json.dumps(a) == json.dumps(v[0]) for a in v) else "DIFFERS", flush=True)
