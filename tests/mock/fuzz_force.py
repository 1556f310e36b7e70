#!/usr/bin/env python
"""Random shapes / Ls / actions through the force terms on the CPU mock (tests/mock/README.md) against the oracle (which is pinned
bit for bit on the compiled reference for these entry points): DhopDeriv, MDeriv (+-dag), MpcDeriv / MpcDagDeriv, and the eight DhopDir
legs.  fp64, per-entry error relative to the largest entry < 1e-12.  Not part of the test suite (open-ended).
usage: fuzz_force.py <libgridb200_mock.so> <seed> <seconds>   (last recorded run: 501 cases, 0 disagreements)"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grid_b200 as gb                      # noqa: E402
from grid_b200 import synthetic as syn      # noqa: E402
from oracle import pyoracle as po           # noqa: E402

gb.LIB_PATH = sys.argv[1]
random.seed(int(sys.argv[2]))
t_end = time.time() + float(sys.argv[3])
ctx = gb.Context(0)
bad, ncase = [], 0


def rel(a, b):
    return float(np.max(np.abs(a.astype(np.complex128) - b)) / np.max(np.abs(b)))


while time.time() < t_end:
    dims = (random.choice([2, 4, 6, 8]), random.choice([2, 4, 6]), random.choice([2, 4, 6]), random.choice([2, 4, 8]))
    kind = random.choice(["wilson", "dwf", "mobius"])
    Ls = 1 if kind == "wilson" else random.choice([2, 4, 6, 8, 12])
    if np.prod(dims) * Ls > 5000:
        continue
    b, c = (1.5, 0.5) if kind == "mobius" else (1.0, 0.0)
    U = syn.hot_gauge(dims, seed=ncase + 1)
    o = po.OracleOp(0 if kind == "wilson" else 1, dims, Ls, mass=0.1, M5=1.8, b=b, c=c, prec=1)
    o.import_gauge(U)
    grid = gb.GridCartesian(ctx, dims)
    Umu = gb.LatticeGaugeField(grid, gb.F64).import_lex(U)
    D = gb.WilsonFermion(Umu, grid, 0.1) if kind == "wilson" else gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8) if kind == "dwf" else \
        gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, b, c)
    hA, hB = syn.random_fermion(dims, Ls, seed=700 + ncase), syn.random_fermion(dims, Ls, seed=800 + ncase)
    A, B = gb.LatticeFermion(grid, Ls, gb.F64).import_lex(hA), gb.LatticeFermion(grid, Ls, gb.F64).import_lex(hB)
    tag = f"{kind} dims {dims} Ls {Ls}"
    mat = gb.LatticeGaugeField(grid, gb.F64)
    for which, meth in ((0, D.DhopDeriv), (1, D.MDeriv)):
        for dag in (0, 1):
            meth(mat, A, B, dag)
            e = rel(mat.export_lex(dtype=np.complex128), o.deriv(which, hA, hB, dag))
            if not e < 1e-12:
                bad.append((tag, f"deriv {which} dag {dag}", e))
    uo, vo = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
    gb.pickCheckerboard(gb.Odd, uo, A); gb.pickCheckerboard(gb.Odd, vo, B)
    S = gb.SchurDifferentiableOperator(D)
    hu, hv = po.pick_checkerboard(dims, Ls, 1, hA), po.pick_checkerboard(dims, Ls, 1, hB)
    for which, meth in ((2, S.MpcDeriv), (3, S.MpcDagDeriv)):
        meth(mat, uo, vo)
        e = rel(mat.export_lex(dtype=np.complex128), o.deriv_eo(which, hu, hv))
        if not e < 1e-12:
            bad.append((tag, f"mpc deriv {which}", e))
    out = gb.LatticeFermion(grid, Ls, gb.F64)
    for d in range(4):
        for sgn in (1, -1):
            D.DhopDir(A, out, d, sgn)
            e = rel(out.export_lex(), o.dhop_dir(hA, d, sgn))
            if not e < 1e-12:
                bad.append((tag, f"DhopDir {d} {sgn}", e))
    ncase += 1
print("cases", ncase, "bad", len(bad), bad[:8], flush=True)
sys.exit(1 if bad else 0)
