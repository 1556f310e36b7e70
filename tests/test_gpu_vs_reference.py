"""GPU parity against the REFERENCE ITSELF, run side by side on the same box: the CUDA path through the C ABI vs
oracle/_ref/libgridref.so (unmodified paboyle/Grid CPU code, built by oracle/Makefile.ref; it travels with the snapshot).
Sizes the reference finishes in seconds on the host cores: 8^4 (BASELINE configs[0]) and 8^3 x 16 x Ls 8/16.
north_star bars: per-site relative error <= 1e-6 fp32 / <= 1e-13 fp64, CG iteration count +-2 %, same true residual."""
import numpy as np
import pytest

import grid_b200 as gb
from grid_b200 import synthetic as syn
from oracle import pyref as pr

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built")]


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)))


@pytest.fixture(scope="module")
def ctx():
    return gb.Context(0)


def test_config0_wilson_dhop_fp64_8x8x8x8(ctx):
    """BASELINE configs[0]: Wilson Dhop fp64 on 8^4 random SU(3) (Benchmark_wilson shape)."""
    dims = (8, 8, 8, 8)
    U, src = syn.hot_gauge(dims, seed=1), syn.random_fermion(dims, 1, seed=2)
    ref = pr.RefOp(0, dims, 1, 0.1, prec=1); ref.import_gauge(U)
    grid = gb.GridCartesian(ctx, dims)
    D = gb.WilsonFermion(gb.LatticeGaugeField(grid, gb.F64).import_lex(U), grid, 0.1)
    fin, out = gb.LatticeFermion(grid, 1, gb.F64).import_lex(src), gb.LatticeFermion(grid, 1, gb.F64)
    for dag in (0, 1):
        D.Dhop(fin, out, dag)
        assert site_err(out.export_lex(), ref.apply(pr.OP_DHOP, src, dag=dag)) < 1e-13
    D.M(fin, out)
    assert site_err(out.export_lex(), ref.apply(pr.OP_M, src)) < 2e-13


@pytest.mark.parametrize("prec,tol", [(gb.F32, 1e-6), (gb.F64, 1e-13)])
@pytest.mark.parametrize("Ls", [8, 16])
def test_dwf_dhop_and_mobius_M(ctx, prec, tol, Ls):
    dims = (8, 8, 8, 16)
    U = syn.hot_gauge(dims, seed=3)
    src = syn.random_fermion(dims, Ls, seed=4, dtype=gb._cdtype(prec))
    ref = pr.RefOp(1, dims, Ls, 0.1, 1.8, 1.5, 0.5, prec=prec); ref.import_gauge(U)
    grid = gb.GridCartesian(ctx, dims)
    D = gb.MobiusFermion(gb.LatticeGaugeField(grid, prec).import_lex(U), grid, Ls, 0.1, 1.8, 1.5, 0.5)
    fin, out = gb.LatticeFermion(grid, Ls, prec).import_lex(src), gb.LatticeFermion(grid, Ls, prec)
    for dag in (0, 1):
        D.Dhop(fin, out, dag)
        assert site_err(out.export_lex(), ref.apply(pr.OP_DHOP, src, dag=dag)) < tol
    D.M(fin, out)
    assert site_err(out.export_lex(), ref.apply(pr.OP_M, src)) < 4 * tol
    half, hout = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    gb.pickCheckerboard(gb.Odd, half, fin)
    hsrc = ref.pick_checkerboard(1, src)
    assert np.array_equal(half.export_lex(), hsrc)
    D.DhopEO(half, hout, 0)
    assert site_err(hout.export_lex(), ref.apply(pr.OP_DHOP_EO, hsrc)) < tol
    gb.SchurDiagMooeeOperator(D).HermOp(half, hout)
    assert site_err(hout.export_lex(), ref.apply(pr.OP_HERMOP, hsrc, cb_in=1)) < 12 * tol


def test_schur_cg_and_mixed_cg_vs_reference(ctx):
    """ConjugateGradient and MixedPrecisionConjugateGradient on SchurDiagMooeeOperator(MobiusFermion), tol 1e-8."""
    dims, Ls = (8, 8, 8, 8), 8
    U = syn.hot_gauge(dims, seed=5)
    src = syn.random_fermion(dims, Ls, seed=6)
    rd = pr.RefOp(1, dims, Ls, 0.1, 1.8, 1.5, 0.5, prec=1); rd.import_gauge(U)
    rf = pr.RefOp(1, dims, Ls, 0.1, 1.8, 1.5, 0.5, prec=0); rf.import_gauge(U)
    hsrc = rd.pick_checkerboard(1, src)
    x_ref, info = rd.cg(1, hsrc, 1e-8, 10000)
    xm_ref, minfo = pr.mixed_cg(rd, rf, 1, hsrc, 1e-8, 10000, 50)
    grid = gb.GridCartesian(ctx, dims)
    Dd = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F64).import_lex(U), grid, Ls, 0.1, 1.8, 1.5, 0.5)
    Df = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).import_lex(U), grid, Ls, 0.1, 1.8, 1.5, 0.5)
    s = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).import_lex(hsrc); s.set_checkerboard(gb.Odd)
    sol = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
    cg = gb.ConjugateGradient(1e-8, 10000)
    cg(gb.SchurDiagMooeeOperator(Dd), s, sol)
    assert abs(cg.IterationsToComplete - info["iterations"]) <= max(1, 0.02 * info["iterations"]), (cg.IterationsToComplete, info)
    assert abs(cg.TrueResidual - info["true_residual"]) < 0.05 * info["true_residual"]
    assert site_err(sol.export_lex(), x_ref) < 1e-7
    solm = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
    mcg = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd))
    mcg(s, solm)
    assert mcg.TotalOuterIterations == minfo["outer"], (mcg.TotalOuterIterations, minfo)
    # fp32 inner solves stop where rounding decides: the reference's own count moves by +-4 in ~100 between runs (threaded
    # reductions), so two fp32 implementations are held to 5 % here; the fp64 solve above carries the +-2 % bar
    assert abs(mcg.TotalInnerIterations - minfo["inner"]) <= max(3, 0.05 * minfo["inner"]), (mcg.TotalInnerIterations, minfo)
    assert mcg.TrueResidual < 1e-8 * 1.05 and minfo["true_residual"] < 1e-8 * 1.05
    assert site_err(solm.export_lex(), xm_ref) < 1e-6
