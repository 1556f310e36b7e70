#!/usr/bin/env python
"""Random shapes / Ls / actions / precisions through the solver rows on the CPU mock (tests/mock/README.md), checked by identities that
need no oracle: the full propagator solve (SchurRedBlackDiagMooeeSolve) must satisfy |M x - src| / |src| < 20 tol, and every pole of
ConjugateGradientMultiShift must satisfy |(MpcDagMpc + pole) x - src| / |src| < 20 tol; every third case also runs the mixed-precision
solvers (reliable-update CG, MixedPrecisionConjugateGradient, ConjugateGradientMultiShiftMixedPrec: fp64 vectors, fp32 inner operator)
against the fp64 operator's residual < 2e-7.  Not part of the test suite (open-ended).
usage: fuzz_solvers.py <libgridb200_mock.so> <seed> <seconds>   (last recorded runs: 346 cases without and 257 with the mixed-precision solvers, then 3 seeds x 100 s with the
improved staggered solve; 0 violations)"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grid_b200 as gb                      # noqa: E402
from grid_b200 import synthetic as syn      # noqa: E402

gb.LIB_PATH = sys.argv[1]
random.seed(int(sys.argv[2]))
t_end = time.time() + float(sys.argv[3])
ctx = gb.Context(0)
bad, ncase = [], 0
while time.time() < t_end:
    dims = (random.choice([2, 4, 8]), random.choice([2, 4]), random.choice([2, 4, 6]), random.choice([2, 4, 8]))
    kind = random.choice(["wilson", "dwf", "mobius"])
    Ls = 1 if kind == "wilson" else random.choice([2, 4, 6, 8, 12, 16])
    if np.prod(dims) * Ls > 6000 or np.prod(dims) < 32:
        continue
    prec = random.choice([gb.F32, gb.F64])
    tol = 1e-5 if prec == gb.F32 else 1e-9
    mass = random.choice([0.05, 0.1, 0.3])
    grid = gb.GridCartesian(ctx, dims)
    if random.random() < 0.2:
        # improved staggered: SchurRedBlackStaggeredSolve, then |M x - src| / |src|
        sd_ = tuple(max(d, 4) for d in dims)
        g = gb.GridCartesian(ctx, sd_)
        Us = gb.LatticeGaugeField(g, prec).import_lex(syn.hot_gauge(sd_, seed=ncase + 1))
        Ds = gb.ImprovedStaggeredFermion(Us, Us, g, mass)
        rng = np.random.default_rng(ncase)
        V = int(np.prod(sd_))
        fs = gb.LatticeStaggeredFermion(g, 1, prec).import_lex((rng.random((V, 3)) + 1j * rng.random((V, 3))).astype(gb._cdtype(prec)))
        xs_, Ms = gb.LatticeStaggeredFermion(g, 1, prec).zero(), gb.LatticeStaggeredFermion(g, 1, prec)
        gb.SchurRedBlackStaggeredSolve(gb.ConjugateGradient(tol, 20000, err_on_no_conv=False))(Ds, fs, xs_)
        Ds.M(xs_, Ms)
        gb.axpy(Ms, -1.0, fs, Ms)
        r = np.sqrt(gb.norm2(Ms) / gb.norm2(fs))
        # CG stops on the residual of the preconditioned (m^2 - Deo Doe) system; the unpreconditioned one is larger by ~ 1 / m
        if not r < 3 * tol / mass:
            bad.append((f"staggered dims {sd_} prec {prec} mass {mass}", "staggered schur solve", r))
        ncase += 1
        continue
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(syn.hot_gauge(dims, seed=ncase + 1))
    D = gb.WilsonFermion(Umu, grid, mass + 0.3) if kind == "wilson" else gb.DomainWallFermion(Umu, grid, Ls, mass, 1.8) if kind == "dwf" else \
        gb.MobiusFermion(Umu, grid, Ls, mass, 1.8, 1.5, 0.5)
    tag = f"{kind} dims {dims} Ls {Ls} prec {prec} mass {mass}"
    src = gb.LatticeFermion(grid, Ls, prec).import_lex(syn.random_fermion(dims, Ls, seed=300 + ncase, dtype=gb._cdtype(prec)))
    # ---- M x = src through the red-black solve
    x, Mx = gb.LatticeFermion(grid, Ls, prec).zero(), gb.LatticeFermion(grid, Ls, prec)
    gb.SchurRedBlackDiagMooeeSolve(gb.ConjugateGradient(tol, 20000, err_on_no_conv=False))(D, src, x)
    D.M(x, Mx)
    gb.axpy(Mx, -1.0, src, Mx)
    r = np.sqrt(gb.norm2(Mx) / gb.norm2(src))
    if not r < 20 * tol:
        bad.append((tag, "schur solve", r))
    # ---- multishift on the odd checkerboard
    so = gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    gb.pickCheckerboard(gb.Odd, so, src)
    poles = sorted(random.sample([0.0, 0.01, 0.05, 0.2, 1.0, 4.0], random.choice([1, 2, 4])))
    res = [gb.LatticeFermion(grid, Ls, prec, gb.HALF) for _ in poles]
    lin = gb.SchurDiagMooeeOperator(D)
    gb.ConjugateGradientMultiShift(20000, gb.MultiShiftFunction(poles, tol))(lin, so, res)
    t = gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    for pole, xk in zip(poles, res):
        lin.HermOp(xk, t)
        gb.axpy(t, pole, xk, t)
        gb.axpy(t, -1.0, so, t)
        r = np.sqrt(gb.norm2(t) / gb.norm2(so))
        if not r < 20 * tol:
            bad.append((tag, f"multishift pole {pole} of {poles}", r))
    # ---- mixed-precision solvers (fp64 vectors, fp32 inner operator) on the odd checkerboard, every third case
    if ncase % 3 == 0 and kind != "wilson":
        U64 = gb.LatticeGaugeField(grid, gb.F64).import_lex(syn.hot_gauge(dims, seed=ncase + 1))
        U32 = gb.LatticeGaugeField(grid, gb.F32).import_lex(syn.hot_gauge(dims, seed=ncase + 1))
        mk = (lambda Um: gb.DomainWallFermion(Um, grid, Ls, mass, 1.8)) if kind == "dwf" else (lambda Um: gb.MobiusFermion(Um, grid, Ls, mass, 1.8, 1.5, 0.5))
        Ld, Lf = gb.SchurDiagMooeeOperator(mk(U64)), gb.SchurDiagMooeeOperator(mk(U32))
        sd = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
        gb.pickCheckerboard(gb.Odd, sd, gb.LatticeFermion(grid, Ls, gb.F64).import_lex(syn.random_fermion(dims, Ls, seed=300 + ncase)))
        td = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)

        def resid(xk, pole=0.0):
            Ld.HermOp(xk, td)
            gb.axpy(td, pole, xk, td)
            gb.axpy(td, -1.0, sd, td)
            return np.sqrt(gb.norm2(td) / gb.norm2(sd))
        xs = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
        gb.ConjugateGradientReliableUpdate(1e-8, 20000, random.choice([0.1, 0.5]), Lf, Ld, err_on_no_conv=False)(sd, xs)
        if not resid(xs) < 2e-7:
            bad.append((tag, "reliable-update CG", resid(xs)))
        xs = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
        gb.MixedPrecisionConjugateGradient(1e-8, 20000, 50, Lf, Ld)(sd, xs)
        if not resid(xs) < 2e-7:
            bad.append((tag, "mixed-precision CG", resid(xs)))
        resm = [gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF) for _ in poles]
        gb.ConjugateGradientMultiShiftMixedPrec(20000, gb.MultiShiftFunction(poles, 1e-8), Lf, random.choice([10, 50]))(Ld, sd, resm)
        for pole, xk in zip(poles, resm):
            if not resid(xk, pole) < 2e-7:
                bad.append((tag, f"mixed multishift pole {pole} of {poles}", resid(xk, pole)))
    ncase += 1
print("cases", ncase, "bad", len(bad), bad[:8], flush=True)
sys.exit(1 if bad else 0)
