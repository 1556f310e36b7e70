"""Stress test: the same hop n times; every result is compared on the device with the first one (|out_i - out_0|^2 must be exactly 0).
usage: python scripts/hop_stress.py L Ls n [op=DhopEO|Dhop|HermOp|smat]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_b200 as gb
L, Ls, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
opname = sys.argv[4] if len(sys.argv) > 4 else "DhopEO"
ctx = gb.Context(0)
grid = gb.GridCartesian(ctx, (L,) * 4)
Df = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).random(1), grid, Ls, 0.1, 1.8, 1.5, 0.5)
if os.environ.get("LAB_GENERIC"):
    Df.set_fast_kernel(0)
src = gb.LatticeFermion(grid, Ls, gb.F32).random(2)
if opname == "Dhop":
    a, r0, r1 = src, gb.LatticeFermion(grid, Ls, gb.F32), gb.LatticeFermion(grid, Ls, gb.F32)
    f = lambda o: Df.Dhop(a, o, 0)
else:
    a = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF); gb.pickCheckerboard(gb.Odd, a, src)
    r0, r1 = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
    Lf = gb.SchurDiagMooeeOperator(Df)
    f = {"DhopEO": lambda o: Df.DhopEO(a, o, 0), "HermOp": lambda o: Lf.HermOp(a, o), "smat": lambda o: Df.MooeeInv(a, o)}[opname]
f(r0)
bad = []
for i in range(n):
    f(r1)
    gb.axpy(r1, -1.0, r0, r1)
    d = gb.norm2(r1)
    if d != 0.0:
        bad.append((i, d))
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith(("GB_", "LAB_")))
print(json.dumps({"tag": tag, "op": opname, "L": L, "calls": n, "mismatches": len(bad), "first": bad[:5], "ref_norm2": gb.norm2(r0)}), flush=True)
