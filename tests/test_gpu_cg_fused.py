"""The Schur conjugate gradient with its linear algebra folded into the s-space passes (csrc/smat.cu epilogues, csrc/fermop.cu
cg_fused_*; ref: ConjugateGradient.h:151-231 on SchurDiagMooeeOperator, LinearOperator.h:291-348) against (a) the same solver in its
unfused form (GB_CG_UNFUSED=1: separate inner product / axpy_norm / update kernels, the form round 1 measured) and (b) the CPU oracle.
Tolerances: iteration counts within max(1, 2 %); solutions 1e-4 (fp32) / 1e-9 (fp64) relative (the two forms differ in rounding
only: d = |Mpc p|^2 instead of Re<p, MpcDag Mpc p>, and Meooe5D(r + b p) is formed as Meooe5D r + b Meooe5D p)."""
import os

import numpy as np
import pytest

import grid_b200 as gb
from grid_b200 import synthetic as syn
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = gb.Context(0)
    yield c
    c.synchronize()


CASES = [((4, 4, 4, 8), 8, "dwf", 1.0, 0.0), ((8, 4, 4, 4), 12, "mobius", 1.5, 0.5), ((8, 8, 4, 4), 16, "mobius", 1.5, 0.5), ((4, 4, 4, 4), 16, "dwf", 1.0, 0.0)]


@pytest.mark.parametrize("dims,Ls,kind,b,c", CASES)
@pytest.mark.parametrize("prec,tol", [(gb.F64, 1e-9), (gb.F32, 1e-5)])
def test_fused_cg_matches_unfused_and_oracle(ctx, dims, Ls, kind, b, c, prec, tol):
    grid = gb.GridCartesian(ctx, dims)
    U = syn.hot_gauge(dims, seed=5)
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(U)
    D = gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, b, c) if kind == "mobius" else gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8)
    h = po.pick_checkerboard(dims, Ls, gb.Odd, syn.random_fermion(dims, Ls, seed=6, dtype=gb._cdtype(prec)))
    src = gb.LatticeFermion(grid, Ls, prec, gb.HALF).import_lex(h)
    src.set_checkerboard(gb.Odd)
    res = {}
    for form in ("fused", "unfused", "fused_no_hop_epilogue"):
        # fused = the default: at fp32 / Ls 16 the w = Mooee p - Dhop(..) and the residual-update passes ride on the hops as epilogues of
        # the column kernel (dhop_col2.cuh EPI 1 / 2); GB_NO_HOP_EPI=1 keeps them as separate streaming passes
        if form == "unfused":
            os.environ["GB_CG_UNFUSED"] = "1"
        if form == "fused_no_hop_epilogue":
            os.environ["GB_NO_HOP_EPI"] = "1"
        try:
            sol = gb.LatticeFermion(grid, Ls, prec, gb.HALF).zero()
            cg = gb.ConjugateGradient(tol, 10000)
            l0 = ctx.launch_count()
            cg(gb.SchurDiagMooeeOperator(D), src, sol)
            res[form] = (cg.IterationsToComplete, cg.TrueResidual, sol.export_lex(), (ctx.launch_count() - l0) / cg.IterationsToComplete)
        finally:
            os.environ.pop("GB_CG_UNFUSED", None)
            os.environ.pop("GB_NO_HOP_EPI", None)
    (it_f, tr_f, x_f, lf), (it_u, tr_u, x_u, lu) = res["fused"], res["unfused"]
    it_s, tr_s, x_s, ls_ = res["fused_no_hop_epilogue"]
    assert abs(it_s - it_u) <= max(1, 0.02 * it_u) and tr_s < 1.5 * tol
    assert np.linalg.norm((x_s - x_u).ravel()) / np.linalg.norm(x_u.ravel()) < (1e-4 if prec == gb.F32 else 1e-9)
    if prec == gb.F32 and Ls == 16 and dims[0] % 8 == 0 and not os.environ.get("GB_TEST_MOCK_LIB"):
        assert lf < ls_, (lf, ls_)                     # the hop epilogues were taken: two launches fewer per iteration
    assert abs(it_f - it_u) <= max(1, 0.02 * it_u), (it_f, it_u)
    assert tr_f < 1.5 * tol and tr_u < 1.5 * tol
    rel = np.linalg.norm((x_f - x_u).ravel()) / np.linalg.norm(x_u.ravel())
    assert rel < (1e-4 if prec == gb.F32 else 1e-9), rel
    if not os.environ.get("GB_TEST_MOCK_LIB"):          # (the CPU mock's BLAS stand-ins do not count launches)
        assert lf < lu - 2, (lf, lu)                  # 11 instead of 14 kernels per iteration
    orc = po.OracleOp(1, dims, Ls, mass=0.1, M5=1.8, b=b, c=c, prec=prec)
    orc.import_gauge(U)
    x_ref, info = orc.cg(gb.Odd, h, tol, 10000)
    assert abs(it_f - info["iterations"]) <= max(1, 0.02 * info["iterations"]), (it_f, info["iterations"])
    assert np.linalg.norm((x_f - x_ref).ravel()) / np.linalg.norm(x_ref.ravel()) < (1e-4 if prec == gb.F32 else 1e-7)


def test_fused_cg_nonzero_guess_and_restart(ctx):
    """a non-zero initial guess (the outer loop of the mixed-precision solver restarts from one) takes the same fused path"""
    dims, Ls = (4, 4, 4, 8), 8
    grid = gb.GridCartesian(ctx, dims)
    U = syn.hot_gauge(dims, seed=7)
    D = gb.DomainWallFermion(gb.LatticeGaugeField(grid, gb.F64).import_lex(U), grid, Ls, 0.1, 1.8)
    h = po.pick_checkerboard(dims, Ls, gb.Odd, syn.random_fermion(dims, Ls, seed=8))
    src = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).import_lex(h)
    src.set_checkerboard(gb.Odd)
    sol = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
    gb.ConjugateGradient(1e-4, 10000)(gb.SchurDiagMooeeOperator(D), src, sol)       # coarse solve
    cg = gb.ConjugateGradient(1e-10, 10000)
    cg(gb.SchurDiagMooeeOperator(D), src, sol)                                       # continue from it
    assert cg.TrueResidual < 1.5e-10
    ref = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
    cg2 = gb.ConjugateGradient(1e-10, 10000)
    cg2(gb.SchurDiagMooeeOperator(D), src, ref)
    assert np.linalg.norm((sol.export_lex() - ref.export_lex()).ravel()) / np.linalg.norm(ref.export_lex().ravel()) < 1e-8
