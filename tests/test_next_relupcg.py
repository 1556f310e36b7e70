"""SURVEY 8(f) row 3 -- ConjugateGradientReliableUpdate (ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:36-270;
driver tests/solver/Test_dwf_relupcg_prec.cc:88-104): fp32 CG with fp64 reliable updates, then an fp64 clean-up CG.

 * CPU: the oracle restatement reproduces the compiled reference's iteration / update counts and solution (fixture
   tests/golden/next_golden.npz; live comparison where oracle/_ref exists) and agrees with plain fp64 CG.
 * GPU: gb_relup_cg_schur reproduces the same fixture.
"""
import os

import numpy as np
import pytest

from grid_b200 import synthetic as syn
from oracle import pyoracle as po
from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "dirac_golden.npz"))
N = np.load(os.path.join(HERE, "golden", "next_golden.npz"))
DIMS, LS = tuple(int(x) for x in G["dims"]), int(G["Ls"])
DELTA = 0.1


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    nb = np.linalg.norm(b, axis=1)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-3 * np.sqrt(np.mean(nb ** 2)) + 1e-300)))


def check(info, x):
    ref_it, ref_up = int(N["mobius/relup_cg/iterations"]), int(N["mobius/relup_cg/reliable_updates"])
    assert abs(info["iterations"] - ref_it) <= max(1, 0.02 * ref_it), (info, ref_it)
    assert abs(info["reliable_updates"] - ref_up) <= 1, (info, ref_up)
    assert abs(info["cleanup_iterations"] - int(N["mobius/relup_cg/cleanup_iterations"])) <= 2
    assert info["true_residual"] < 2e-8
    assert site_err(x, N["mobius/relup_cg/solution"]) < 1e-6


def test_oracle_relup_cg_matches_reference_outputs():
    od = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=1); od.import_gauge(G["U"])
    of = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=0); of.import_gauge(G["U"])
    src = po.pick_checkerboard(DIMS, LS, 1, G["src5"])
    x, info = po.relup_cg(od, of, 1, src, 1e-8, 5000, DELTA)
    assert info["converged"] == 1 and info["reliable_updates"] >= 2
    check(info, x)
    # same solution as plain double-precision CG, in a comparable number of iterations (the reference's own comparison, :106-120)
    xd, cg = od.cg(1, src, 1e-8, 5000)
    assert site_err(x, xd) < 1e-6 and info["iterations"] <= 1.3 * cg["iterations"]


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("delta", [0.1, 0.5])
def test_oracle_vs_reference_relup_cg(delta):
    dims, Ls = (4, 6, 8, 4), 6
    U = syn.hot_gauge(dims, seed=21)
    mk = lambda mod, prec: mod(1, dims, Ls, mass=0.1, M5=1.8, b=1.0, c=0.0, prec=prec)
    od, of, rd, rf = mk(po.OracleOp, 1), mk(po.OracleOp, 0), mk(pr.RefOp, 1), mk(pr.RefOp, 0)
    for o in (od, of, rd, rf):
        o.import_gauge(U)
    src = po.pick_checkerboard(dims, Ls, 1, syn.random_fermion(dims, Ls, seed=3))
    a, ia = po.relup_cg(od, of, 1, src, 1e-8, 5000, delta)
    b, ib = pr.relup_cg(rd, rf, 1, src, 1e-8, 5000, delta)
    assert abs(ia["iterations"] - ib["iterations"]) <= max(1, 0.02 * ib["iterations"]), (ia, ib)
    assert abs(ia["reliable_updates"] - ib["reliable_updates"]) <= 1
    assert site_err(a, b) < 1e-6


@pytest.mark.gpu
def test_cuda_relup_cg_matches_reference():
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Dd = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F64).import_lex(G["U"]), grid, LS, 0.1, 1.8, 1.5, 0.5)
    Df = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).import_lex(G["U"]), grid, LS, 0.1, 1.8, 1.5, 0.5)
    full = gb.LatticeFermion(grid, LS, gb.F64).import_lex(G["src5"])
    src, sol = gb.LatticeFermion(grid, LS, gb.F64, gb.HALF), gb.LatticeFermion(grid, LS, gb.F64, gb.HALF).zero()
    gb.pickCheckerboard(gb.Odd, src, full)
    mCG = gb.ConjugateGradientReliableUpdate(1e-8, 5000, DELTA, gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd))
    mCG(src, sol)
    check(dict(iterations=mCG.IterationsToComplete, reliable_updates=mCG.ReliableUpdatesPerformed, cleanup_iterations=mCG.IterationsToCleanup,
               true_residual=mCG.TrueResidual), sol.export_lex())
    assert sol.Checkerboard() == gb.Odd


@pytest.mark.gpu
def test_cuda_mixed_cg_batched():
    """MixedPrecisionConjugateGradientBatched through the C ABI (gb_mixed_cg_batched_schur; ref: ConjugateGradientMixedPrecBatched.h:36-213):
    three right-hand sides of different norms, one restart schedule.  Against the oracle's restatement (itself pinned against the
    compiled reference in tests/test_oracle_vs_reference.py) and, where it travelled with the snapshot, the compiled reference: same
    number of restarts, inner iterations per right-hand side within 5 % (fp32 solves), patch-up iterations +-2, same solutions; and
    each solution satisfies HermOp x = b to the tolerance."""
    import grid_b200 as gb
    from oracle import pyref as pr
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Dd = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F64).import_lex(G["U"]), grid, LS, 0.1, 1.8, 1.5, 0.5)
    Df = gb.MobiusFermion(gb.LatticeGaugeField(grid, gb.F32).import_lex(G["U"]), grid, LS, 0.1, 1.8, 1.5, 0.5)
    Ld, Lf = gb.SchurDiagMooeeOperator(Dd), gb.SchurDiagMooeeOperator(Df)
    hosts = [f * po.pick_checkerboard(DIMS, LS, 1, syn.random_fermion(DIMS, LS, seed=sd)) for sd, f in ((31, 1.0), (32, 3.0), (33, 0.2))]
    srcs, sols = [], []
    for h in hosts:
        s = gb.LatticeFermion(grid, LS, gb.F64, gb.HALF).import_lex(h)
        s.set_checkerboard(gb.Odd)
        srcs.append(s); sols.append(gb.LatticeFermion(grid, LS, gb.F64, gb.HALF).zero())
    B = gb.MixedPrecisionConjugateGradientBatched(1e-8, 10000, 50, 10000, Lf, Ld)
    B(srcs, sols)
    od = po.OracleOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=1); od.import_gauge(G["U"])
    of = po.OracleOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=0); of.import_gauge(G["U"])
    xo, io = po.mixed_cg_batched(od, of, 1, np.stack(hosts), 1e-8, 10000, 50, 10000)
    refs = [("oracle", xo, io)]
    if pr.available():
        rd = pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=1); rd.import_gauge(G["U"])
        rf = pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=0); rf.import_gauge(G["U"].astype(np.complex64))
        xr, ir = pr.mixed_cg_batched(rd, rf, 1, np.stack(hosts), 1e-8, 10000, 50, 10000)
        refs.append(("reference", xr, ir))
    for name, xref, iref in refs:
        assert B.TotalOuterIterations == iref["outer"], (name, B.TotalOuterIterations, iref)
        for a, b in zip(B.TotalInnerIterations, iref["inner"]):
            assert abs(a - b) <= max(3, 0.05 * b), (name, B.TotalInnerIterations, iref)
        for a, b in zip(B.TotalFinalStepIterations, iref["final"]):
            assert abs(a - b) <= 2, (name, B.TotalFinalStepIterations, iref)   # 1-3 clean-up iterations, decided by fp32 rounding
        for i in range(3):
            assert site_err(sols[i].export_lex(), xref[i]) < 1e-6, (name, i)
    assert max(B.TrueResidual) < 1.5e-8
    for s, x in zip(srcs, sols):
        r = s.like()
        Ld.HermOp(x, r)
        gb.axpy(r, -1.0, s, r)
        assert (gb.norm2(r) / gb.norm2(s)) ** 0.5 < 2e-8
        assert x.Checkerboard() == gb.Odd
