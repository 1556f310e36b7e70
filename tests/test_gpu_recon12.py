"""12-real SU(3) link reconstruction (BASELINE north star: "optional 12-real SU(3) link reconstruction"; the reference always keeps
18 reals per link, WilsonImpl.h:127-171 DoubleStore): gb_op_set_link_reconstruct(op, 12) stores two rows of the bare link, rebuilds
the third as conj(row0 x row1) in registers and applies the folded-in -1/2 and boundary phase to the product.  Same operator, so the
same bar as the full store: per-site <= 1e-6 (fp32) / 1e-13 (fp64) against the fp64 oracle."""
import os

import numpy as np
import pytest

import grid_b200 as gb
from grid_b200 import synthetic as syn
from oracle import pyoracle as po
from test_gpu_parity import site_rel_err, TOL_HOP, TOL_COMPOSITE

pytestmark = pytest.mark.gpu

SHAPES = {
    "wilson": dict(dims=(8, 4, 4, 6), Ls=1, kind="wilson"),
    "dwf8": dict(dims=(8, 8, 4, 4), Ls=8, kind="dwf"),
    "mobius6": dict(dims=(4, 4, 6, 4), Ls=6, kind="mobius", b=1.5, c=0.5),
}
PHASES = {"periodic": None, "antiperiodic-t": [1, 1, 1, -1], "complex": [np.exp(0.3j), 1, np.exp(-1.1j), -1]}


@pytest.fixture(scope="module")
def ctx():
    c = gb.Context(0)
    yield c
    c.synchronize()


def build(ctx, sh, prec, phases, U=None, self_halo=0):
    dims, Ls, kind = sh["dims"], sh["Ls"], sh["kind"]
    grid = gb.GridCartesian(ctx, dims)
    U = syn.hot_gauge(dims, seed=21) if U is None else U
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(U)
    if self_halo:
        os.environ["GB_SELF_HALO"] = str(self_halo)
    try:
        if kind == "wilson":
            D = gb.WilsonFermion(Umu, grid, 0.1, phases)
        elif kind == "dwf":
            D = gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8, phases)
        else:
            D = gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, sh["b"], sh["c"], phases)
        f = gb.LatticeFermion(grid, Ls, prec).zero()
        D.Dhop(f, gb.LatticeFermion(grid, Ls, prec), 0)     # the halo state is created by the first hop
    finally:
        os.environ.pop("GB_SELF_HALO", None)
    orc = po.OracleOp(0 if kind == "wilson" else 1, dims, Ls, mass=0.1, M5=1.8, b=sh.get("b", 1.0), c=sh.get("c", 0.0), prec=1)
    orc.import_gauge(U, phases)
    return grid, Umu, D, orc


def check_entries(grid, D, orc, sh, prec, tag):
    dims, Ls = sh["dims"], sh["Ls"]
    src = syn.random_fermion(dims, Ls, seed=22, dtype=gb._cdtype(prec))
    src64 = src.astype(np.complex128)
    fin, fout = gb.LatticeFermion(grid, Ls, prec).import_lex(src), gb.LatticeFermion(grid, Ls, prec)
    he, ho, r = (gb.LatticeFermion(grid, Ls, prec, gb.HALF) for _ in range(3))
    gb.pickCheckerboard(gb.Even, he, fin); gb.pickCheckerboard(gb.Odd, ho, fin)
    for dag in (0, 1):
        D.Dhop(fin, fout, dag)
        assert site_rel_err(fout.export_lex(), orc.apply(po.OP_DHOP, src64, dag=dag)) < TOL_HOP[prec], (tag, "Dhop", dag)
        D.DhopEO(ho, r, dag)
        assert site_rel_err(r.export_lex(), orc.apply(po.OP_DHOP_EO, po.pick_checkerboard(dims, Ls, 1, src64), dag=dag)) < TOL_HOP[prec], (tag, "DhopEO", dag)
        D.DhopOE(he, r, dag)
        assert site_rel_err(r.export_lex(), orc.apply(po.OP_DHOP_OE, po.pick_checkerboard(dims, Ls, 0, src64), dag=dag)) < TOL_HOP[prec], (tag, "DhopOE", dag)
    D.M(fin, fout)
    assert site_rel_err(fout.export_lex(), orc.apply(po.OP_M, src64)) < TOL_COMPOSITE[prec], (tag, "M")
    D.Mdag(fin, fout)
    assert site_rel_err(fout.export_lex(), orc.apply(po.OP_MDAG, src64)) < TOL_COMPOSITE[prec], (tag, "Mdag")


@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32", "f64"])
@pytest.mark.parametrize("ph", list(PHASES))
@pytest.mark.parametrize("shape", list(SHAPES))
def test_recon12_matches_the_oracle(ctx, shape, ph, prec):
    sh = SHAPES[shape]
    grid, Umu, D, orc = build(ctx, sh, prec, PHASES[ph])
    D.set_link_reconstruct(12)
    check_entries(grid, D, orc, sh, prec, (shape, ph, prec, 12))
    D.set_link_reconstruct(18)
    check_entries(grid, D, orc, sh, prec, (shape, ph, prec, 18))


@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32", "f64"])
def test_recon12_on_a_decomposed_lattice_and_with_the_tuned_shape(ctx, prec):
    """z+t halos through the halo path (one GPU: self halo), Ls 16 on a shape the tuned kernels would take: with two-row links the
    generic kernel serves every form (overlapped = interior + exterior slabs, serial), boundary phases on the global boundary"""
    sh = dict(dims=(8, 8, 8, 8), Ls=16, kind="dwf")
    grid, Umu, D, orc = build(ctx, sh, prec, [1, 1, 1, -1], self_halo=12)
    D.set_link_reconstruct(12)
    for overlap in (1, 2, 0):
        D.set_overlap(overlap)
        check_entries(grid, D, orc, sh, prec, ("self-halo zt", overlap))


def test_recon12_schur_cg_iterations(ctx):
    """the CG on the two-row operator is the same solve: same iteration count as on the full store, solution within fp64 roundoff"""
    sh = SHAPES["dwf8"]
    dims, Ls = sh["dims"], sh["Ls"]
    grid, Umu, D, orc = build(ctx, sh, gb.F64, None)
    src = syn.random_fermion(dims, Ls, seed=23)
    so = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
    gb.pickCheckerboard(gb.Odd, so, gb.LatticeFermion(grid, Ls, gb.F64).import_lex(src))
    sols, iters = [], []
    for nreal in (18, 12):
        D.set_link_reconstruct(nreal)
        x = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
        cg = gb.ConjugateGradient(1e-8, 10000)
        cg(gb.SchurDiagMooeeOperator(D), so, x)
        sols.append(x.export_lex()); iters.append(cg.IterationsToComplete)
        assert cg.TrueResidual < 1.5e-8
    assert abs(iters[0] - iters[1]) <= 1, iters
    assert np.linalg.norm((sols[0] - sols[1]).ravel()) / np.linalg.norm(sols[0].ravel()) < 1e-9


def test_recon12_refuses_links_that_are_not_special_unitary(ctx):
    """smeared / rescaled links are not SU(3): the third row cannot be rebuilt, the call says so and the operator keeps the full store"""
    sh = SHAPES["wilson"]
    U = syn.hot_gauge(sh["dims"], seed=24)
    bad = U.copy(); bad[5, 2] *= 1.001
    grid, Umu, D, orc = build(ctx, sh, gb.F64, None, U=bad)
    with pytest.raises(gb.GridB200Error):
        D.set_link_reconstruct(12)
    check_entries(grid, D, orc, sh, gb.F64, "full store after the refusal")
    # a good field imported afterwards can be reconstructed; importing the bad one again is refused at ImportGauge
    good = gb.LatticeGaugeField(grid, gb.F64).import_lex(U)
    D.ImportGauge(good)
    D.set_link_reconstruct(12)
    orc.import_gauge(U, None)
    check_entries(grid, D, orc, sh, gb.F64, "two rows after ImportGauge")
    with pytest.raises(gb.GridB200Error):
        D.ImportGauge(Umu)
