"""GPU check of the multi-rank code path of the improved staggered operator on ONE device.

GB_STAG_SELF_HALO=<bitmask> makes gb_op_create_staggered treat the chosen undecomposed dimensions as decomposed: the gauge faces
for the double store and the three-deep colour-vector halos are packed, "exchanged" with the rank itself and consumed through
the halo lookups of stag_dhop_kernel<..., COMM=1>.  The neighbour is this rank, so the result must be the periodic one:
 * equal to the oracle (fp64 <= 1e-13, fp32 <= 1e-6 relative to the rms site norm), and
 * BIT-IDENTICAL to the plain single-rank kernel (same operands, same order of operations).
The index logic is also checked on the CPU for real processor grids (tests/test_stag_halo_host.py); N-rank runs against the
oracle are scripts/mgpu_check.py.
"""
import os

import numpy as np
import pytest

from grid_b200 import synthetic as syn
from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu]
DIMS = (8, 6, 4, 8)


def stag_err(got, ref):
    d = np.linalg.norm(got.astype(np.complex128) - ref, axis=1); nb = np.linalg.norm(ref, axis=1)
    return float(np.max(d / np.maximum(nb, np.sqrt(np.mean(nb ** 2)))))


def make_op(gb, grid, U, prec, mask):
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(U)
    if mask:
        os.environ["GB_STAG_SELF_HALO"] = str(mask)
    try:
        return gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
    finally:
        os.environ.pop("GB_STAG_SELF_HALO", None)


@pytest.mark.parametrize("prec_name", ["f64", "f32"])
@pytest.mark.parametrize("mask", [8, 4, 2, 1, 12, 15])
def test_self_halo_path_matches_oracle_and_plain_kernel(mask, prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol = 1e-13 if prec == gb.F64 else 1e-6
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    U = syn.hot_gauge(DIMS, seed=31)
    rng = np.random.default_rng(32)
    V = int(np.prod(DIMS))
    src = (rng.random((V, 3)) + 1j * rng.random((V, 3))).astype(gb._cdtype(prec))
    orc = po.StagOracleOp(DIMS, 0.1, prec=1)
    orc.import_gauge(U)
    Dh, Dp = make_op(gb, grid, U, prec, mask), make_op(gb, grid, U, prec, 0)
    fin = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(src)
    oh, op_ = gb.LatticeStaggeredFermion(grid, 1, prec), gb.LatticeStaggeredFermion(grid, 1, prec)
    for overlap in (True, False):          # interior || exchange, then exterior  vs  exchange, then one kernel
        Dh.set_overlap(overlap)
        for dag in (0, 1):
            Dh.Dhop(fin, oh, dag); Dp.Dhop(fin, op_, dag)
            got = oh.export_lex()
            assert stag_err(got, orc.apply(po.OP_DHOP, src.astype(np.complex128), dag=dag)) < tol, (mask, dag, overlap)
            assert np.array_equal(got, op_.export_lex()), (mask, dag, overlap)
    Dh.set_overlap(True)
    Dh.M(fin, oh); Dp.M(fin, op_)
    assert np.array_equal(oh.export_lex(), op_.export_lex())
    assert stag_err(oh.export_lex(), orc.apply(po.OP_M, src.astype(np.complex128))) < tol
    he, ho = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
    r1, r2 = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
    gb.pickCheckerboard(gb.Odd, ho, fin); gb.pickCheckerboard(gb.Even, he, fin)
    for hin, meth in ((ho, "DhopEO"), (he, "DhopOE")):
        for dag in (0, 1):
            getattr(Dh, meth)(hin, r1, dag); getattr(Dp, meth)(hin, r2, dag)
            assert np.array_equal(r1.export_lex(), r2.export_lex()), (mask, meth, dag)
    ref = orc.apply(po.OP_DHOP_EO, po.pick_checkerboard_sites(DIMS, 1, src.astype(np.complex128)))
    Dh.DhopEO(ho, r1, 0)
    assert stag_err(r1.export_lex(), ref) < tol


def test_self_halo_cg_matches_plain():
    """SchurStaggeredOperator CG through the halo path: same iteration count and solution as the plain kernel."""
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    U = syn.hot_gauge(DIMS, seed=31)
    Dh, Dp = make_op(gb, grid, U, gb.F64, 12), make_op(gb, grid, U, gb.F64, 0)
    src = gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF).random(5)
    src.set_checkerboard(gb.Odd)
    out = []
    for D in (Dh, Dp):
        sol = gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF).zero()
        cg = gb.ConjugateGradient(1e-8, 5000)
        cg(gb.SchurStaggeredOperator(D), src, sol)
        out.append((cg.IterationsToComplete, sol.export_lex()))
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1], out[1][1])
