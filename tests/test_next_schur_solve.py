"""SURVEY 8(f) row 1 -- the full propagator solve around the red-black CG:
SchurRedBlackDiagMooeeSolve / SchurRedBlackStaggeredSolve (ref: Grid/algorithms/iterative/SchurRedBlack.h:238-290,294-349,385-430),
CayleyFermion5D::Dminus[Dag] and the physical 4D <-> 5D maps (ref: CayleyFermion5DImplementation.h:58-153).

 * CPU: the oracle reproduces tests/golden/next_golden.npz (outputs of the compiled reference, generator
   tests/golden/make_golden_next.py), agrees with the compiled reference on a second lattice where oracle/_ref exists, and
   satisfies the identities the reference's own tests rely on (M * solve(src) == src, tests/solver/Test_dwf_cg_schur.cc:46-61).
 * GPU: the CUDA path through the C ABI reproduces the same fixtures (fp64 <= 2e-13 composite, fp32 <= 4e-6), same CG iteration
   count +-2 %, same residuals.
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
OPS = {"wilson": dict(kind=0, Ls=1, b=1.0, c=0.0, src="src4"), "dwf": dict(kind=1, Ls=LS, b=1.0, c=0.0, src="src5"),
       "mobius": dict(kind=1, Ls=LS, b=1.5, c=0.5, src="src5")}


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    nb = np.linalg.norm(b, axis=1)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-3 * np.sqrt(np.mean(nb ** 2)) + 1e-300)))


def oracle_op(name, prec=1):
    if name == "stag":
        o = po.StagOracleOp(DIMS, 0.1, prec=prec)
        o.import_gauge(G["U"])
        return o, G["src_stag"]
    cfg = OPS[name]
    o = po.OracleOp(cfg["kind"], DIMS, cfg["Ls"], mass=0.1, M5=1.8, b=cfg["b"], c=cfg["c"], prec=prec)
    o.import_gauge(G["U"])
    return o, G[cfg["src"]]


# ---------------------------------------------------------------------------------------------- CPU: oracle vs the reference's outputs
def test_oracle_dminus_and_physical_maps_match_reference_outputs():
    o, src5 = oracle_op("mobius")
    assert site_err(o.apply(po.OP_DMINUS, src5), N["mobius/DMINUS"]) < 1e-13
    assert site_err(o.apply(po.OP_DMINUS_DAG, src5), N["mobius/DMINUS_DAG"]) < 1e-13
    assert site_err(o.physical(po.IMPORT_PHYSICAL_SOURCE, G["src4"]), N["mobius/physical/0"]) < 1e-13
    # the wall maps only move components: bit-exact, including the zeros in the bulk
    assert np.array_equal(o.physical(po.IMPORT_UNPHYSICAL, G["src4"]), N["mobius/physical/1"])
    assert np.array_equal(o.physical(po.EXPORT_PHYSICAL_SOLUTION, src5), N["mobius/physical/2"])
    assert np.array_equal(o.physical(po.EXPORT_PHYSICAL_SOURCE, src5), N["mobius/physical/3"])


@pytest.mark.parametrize("name", ["mobius", "stag"])
def test_oracle_redblack_source_and_solution_match_reference_outputs(name):
    o, src = oracle_op(name)
    e, od = o.redblack_source(src)
    assert np.array_equal(e, N[f"{name}/rb_source/e"])
    assert site_err(od, N[f"{name}/rb_source/o"]) < 4e-13
    pick = po.pick_checkerboard(DIMS, LS, 1, src) if name != "stag" else po.pick_checkerboard_sites(DIMS, 1, src)
    assert site_err(o.redblack_solution(pick, e), N[f"{name}/rb_solution"]) < 4e-13


@pytest.mark.parametrize("name", ["wilson", "dwf", "mobius", "stag"])
def test_oracle_schur_solve_matches_reference_outputs(name):
    o, src = oracle_op(name)
    x, info = o.schur_solve(src, 1e-8, 5000)
    ref_it = int(N[f"{name}/schur_solve/iterations"])
    assert abs(info["iterations"] - ref_it) <= max(1, 0.02 * ref_it), (info, ref_it)
    for k in ("true_residual", "unprec_residual"):
        r = float(N[f"{name}/schur_solve/{k}"])
        assert 0.6 < info[k] / r < 1.6, (k, info[k], r)
    assert site_err(x, N[f"{name}/schur_solve/solution"]) < 1e-6
    # the defining property (tests/solver/Test_dwf_cg_schur.cc:46-61): M x = src
    Mx = o.apply(po.OP_M, x)          # bound: the reference's own unpreconditioned residual for this solve (staggered: 1e-7, borderline)
    assert np.linalg.norm(Mx - src) / np.linalg.norm(src) < 3 * float(N[f"{name}/schur_solve/unprec_residual"])


def test_oracle_physical_map_identities():
    """Export(Import) projections and the 5D -> 4D reduction used by every propagator code path."""
    o, src5 = oracle_op("mobius")
    s4 = G["src4"]
    imp = o.physical(po.IMPORT_UNPHYSICAL, s4)
    # ExportPhysicalFermionSource inverts ImportUnphysicalFermion (P+ at s=0 and P- at s=Ls-1 are put back together)
    assert np.array_equal(o.physical(po.EXPORT_PHYSICAL_SOURCE, imp), s4)
    # ExportPhysicalFermionSolution picks the opposite chiralities: nothing of the walls' content survives
    assert np.count_nonzero(o.physical(po.EXPORT_PHYSICAL_SOLUTION, imp)) == 0
    # Dminus = 1 - cs DW: with c = 0 (Shamir) it is the identity, so ImportPhysicalFermionSource == ImportUnphysicalFermion
    d, _ = oracle_op("dwf")
    assert np.array_equal(d.physical(po.IMPORT_PHYSICAL_SOURCE, s4), d.physical(po.IMPORT_UNPHYSICAL, s4))
    assert np.array_equal(d.apply(po.OP_DMINUS, src5), src5)
    # adjointness: <a, Dminus b> == <DminusDag a, b>
    a = syn.random_fermion(DIMS, LS, seed=9)
    lhs = po.inner_product(a, o.apply(po.OP_DMINUS, src5)); rhs = po.inner_product(o.apply(po.OP_DMINUS_DAG, a), src5)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


# ---------------------------------------------------------------------------------------------- CPU: oracle vs the compiled reference, second lattice
DIMS2, LS2 = (4, 6, 8, 4), 6


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("bc", [(1.0, 0.0), (1.5, 0.5)])
def test_oracle_vs_reference_propagator_solve(prec, bc):
    tol = 1e-13 if prec else 1e-6
    U = syn.hot_gauge(DIMS2, seed=21)
    o = po.OracleOp(1, DIMS2, LS2, mass=0.1, M5=1.8, b=bc[0], c=bc[1], prec=prec); o.import_gauge(U)
    r = pr.RefOp(1, DIMS2, LS2, mass=0.1, M5=1.8, b=bc[0], c=bc[1], prec=prec); r.import_gauge(U)
    x5 = syn.random_fermion(DIMS2, LS2, seed=3).astype(po._cdtype(prec))
    x4 = syn.random_fermion(DIMS2, 1, seed=4).astype(po._cdtype(prec))
    for which in (po.OP_DMINUS, po.OP_DMINUS_DAG):
        assert site_err(o.apply(which, x5), r.apply(which, x5)) < 2 * tol
    for w in range(4):
        assert site_err(o.physical(w, x4 if w < 2 else x5), r.physical(w, x4 if w < 2 else x5)) < 2 * tol, w
    eo, oo = o.redblack_source(x5); er, orr = r.redblack_source(x5)
    assert np.array_equal(eo, er) and site_err(oo, orr) < 8 * tol
    so = syn.random_fermion(DIMS2, LS2, seed=7)[: x5.shape[0] // 2].astype(po._cdtype(prec))
    assert site_err(o.redblack_solution(so, eo), r.redblack_solution(so, eo)) < 8 * tol
    cgtol = 1e-8 if prec else 1e-5
    s1, i1 = o.schur_solve(x5, cgtol, 2000); s2, i2 = r.schur_solve(x5, cgtol, 2000)
    assert abs(i1["iterations"] - i2["iterations"]) <= max(1, 0.02 * i2["iterations"])
    assert 0.6 < i1["unprec_residual"] / i2["unprec_residual"] < 1.6
    assert site_err(s1, s2) < (1e-6 if prec else 1e-3)


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
def test_oracle_vs_reference_staggered_solve():
    U = syn.hot_gauge(DIMS2, seed=21)
    o = po.StagOracleOp(DIMS2, 0.1, prec=1); o.import_gauge(U)
    r = pr.RefOp(2, DIMS2, 1, 0.1, 9.0 / 8.0, -1.0 / 24.0, 1.0, prec=1); r.import_gauge(U)
    rng = np.random.default_rng(11)
    V = int(np.prod(DIMS2))
    x = rng.random((V, 3)) + 1j * rng.random((V, 3))
    eo, oo = o.redblack_source(x); er, orr = r.redblack_source(x)
    assert np.array_equal(eo, er) and site_err(oo, orr) < 1e-12
    s1, i1 = o.schur_solve(x, 1e-8, 5000); s2, i2 = r.schur_solve(x, 1e-8, 5000)
    assert abs(i1["iterations"] - i2["iterations"]) <= max(1, 0.02 * i2["iterations"])
    assert site_err(s1, s2) < 1e-6


# ---------------------------------------------------------------------------------------------- GPU: the CUDA path vs the reference's outputs
def _device_op(gb, grid, name, prec):
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(G["U"])
    if name == "stag":
        return gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
    cfg = OPS[name]
    if cfg["kind"] == 0:
        return gb.WilsonFermion(Umu, grid, 0.1)
    if cfg["b"] == 1.0:
        return gb.DomainWallFermion(Umu, grid, cfg["Ls"], 0.1, 1.8)
    return gb.MobiusFermion(Umu, grid, cfg["Ls"], 0.1, 1.8, cfg["b"], cfg["c"])


def _field(gb, grid, name, prec, kind=None, Ls=None):
    kind = gb.FULL if kind is None else kind
    if name == "stag":
        return gb.LatticeStaggeredFermion(grid, 1, prec, kind)
    return gb.LatticeFermion(grid, OPS[name]["Ls"] if Ls is None else Ls, prec, kind)


@pytest.mark.gpu
@pytest.mark.parametrize("prec_name", ["f64", "f32"])
def test_cuda_dminus_and_physical_maps(prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol = 2e-13 if prec == gb.F64 else 4e-6
    dt = gb._cdtype(prec)
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, "mobius", prec)
    src5 = gb.LatticeFermion(grid, LS, prec).import_lex(G["src5"].astype(dt))
    src4 = gb.LatticeFermion(grid, 1, prec).import_lex(G["src4"].astype(dt))
    out5, out4 = gb.LatticeFermion(grid, LS, prec), gb.LatticeFermion(grid, 1, prec)
    D.Dminus(src5, out5)
    assert site_err(out5.export_lex(), N["mobius/DMINUS"]) < tol
    D.DminusDag(src5, out5)
    assert site_err(out5.export_lex(), N["mobius/DMINUS_DAG"]) < tol
    D.ImportPhysicalFermionSource(src4, out5)
    assert site_err(out5.export_lex(), N["mobius/physical/0"]) < tol
    D.ImportUnphysicalFermion(src4, out5)
    assert np.array_equal(out5.export_lex(), N["mobius/physical/1"].astype(dt))
    D.ExportPhysicalFermionSolution(src5, out4)
    assert np.array_equal(out4.export_lex(), N["mobius/physical/2"].astype(dt))
    D.ExportPhysicalFermionSource(src5, out4)
    assert np.array_equal(out4.export_lex(), N["mobius/physical/3"].astype(dt))
    # 4D operators: every map is a copy (ref: FermionOperator.h:172-191)
    W = _device_op(gb, grid, "wilson", prec)
    W.ImportPhysicalFermionSource(src4, out4)
    assert np.array_equal(out4.export_lex(), G["src4"].astype(dt))
    W.Dminus(src4, out4)
    assert np.array_equal(out4.export_lex(), G["src4"].astype(dt))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mobius", "stag"])
def test_cuda_redblack_source_and_solution(name):
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, name, gb.F64)
    src_host = G["src_stag"] if name == "stag" else G[OPS[name]["src"]]
    src = _field(gb, grid, name, gb.F64).import_lex(src_host)
    e, o, so = (_field(gb, grid, name, gb.F64, gb.HALF) for _ in range(3))
    S = (gb.SchurRedBlackStaggeredSolve if name == "stag" else gb.SchurRedBlackDiagMooeeSolve)(gb.ConjugateGradient(1e-8, 5000))
    S.RedBlackSource(D, src, e, o)
    assert e.Checkerboard() == gb.Even and o.Checkerboard() == gb.Odd
    assert np.array_equal(e.export_lex(), N[f"{name}/rb_source/e"])
    assert site_err(o.export_lex(), N[f"{name}/rb_source/o"]) < 4e-13
    gb.pickCheckerboard(gb.Odd, so, src)
    sol = _field(gb, grid, name, gb.F64)
    S.RedBlackSolution(D, so, e, sol)
    assert site_err(sol.export_lex(), N[f"{name}/rb_solution"]) < 4e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["wilson", "dwf", "mobius", "stag"])
def test_cuda_schur_solve_matches_reference(name):
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, name, gb.F64)
    src_host = G["src_stag"] if name == "stag" else G[OPS[name]["src"]]
    src, sol = _field(gb, grid, name, gb.F64).import_lex(src_host), _field(gb, grid, name, gb.F64)
    CG = gb.ConjugateGradient(1e-8, 5000)
    S = (gb.SchurRedBlackStaggeredSolve if name == "stag" else gb.SchurRedBlackDiagMooeeSolve)(CG)
    S(D, src, sol)
    ref_it = int(N[f"{name}/schur_solve/iterations"])
    assert abs(CG.IterationsToComplete - ref_it) <= max(1, 0.02 * ref_it), (CG.IterationsToComplete, ref_it)
    assert 0.6 < CG.TrueResidual / float(N[f"{name}/schur_solve/true_residual"]) < 1.6
    assert 0.6 < S.TrueUnprecResidual / float(N[f"{name}/schur_solve/unprec_residual"]) < 1.6
    x = sol.export_lex()
    assert site_err(x, N[f"{name}/schur_solve/solution"]) < 1e-6
    # the same solve through the generic path (any OperatorFunction as the red-black solver; here subtractGuess exercises it)
    sol2 = _field(gb, grid, name, gb.F64)
    CG2 = gb.ConjugateGradient(1e-8, 5000)
    S2 = type(S)(CG2, initSubGuess=True)     # zero guess: subtracting it changes nothing
    S2(D, src, sol2)
    assert CG2.IterationsToComplete == CG.IterationsToComplete
    assert site_err(sol2.export_lex(), x) < 1e-10
    # M sol = src, checked by the oracle on the host
    o, _ = oracle_op(name)
    assert np.linalg.norm(o.apply(po.OP_M, x) - src_host) / np.linalg.norm(src_host) < 3 * float(N[f"{name}/schur_solve/unprec_residual"])


@pytest.mark.gpu
def test_cuda_schur_solve_mixed_precision():
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Dd, Df = _device_op(gb, grid, "mobius", gb.F64), _device_op(gb, grid, "mobius", gb.F32)
    src, sol = gb.LatticeFermion(grid, LS, gb.F64).import_lex(G["src5"]), gb.LatticeFermion(grid, LS, gb.F64)
    info = gb.schur_solve_mixed(Df, Dd, src, sol, 1e-8, 10000, 50)
    assert info["outer"] >= 1 and info["true_residual"] < 1e-7
    assert 0.3 < info["unprec_residual"] / float(N["mobius/schur_solve/unprec_residual"]) < 3.0
    assert site_err(sol.export_lex(), N["mobius/schur_solve/solution"]) < 1e-6


@pytest.mark.gpu
def test_cuda_propagator_column_4d_to_4d():
    """One column of a propagator the way physics callers drive it: 4D source -> ImportPhysicalFermionSource -> Schur solve ->
    ExportPhysicalFermionSolution, against the oracle doing the same chain."""
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, "mobius", gb.F64)
    src4 = gb.LatticeFermion(grid, 1, gb.F64).import_lex(G["src4"])
    src5, sol5, sol4 = gb.LatticeFermion(grid, LS, gb.F64), gb.LatticeFermion(grid, LS, gb.F64), gb.LatticeFermion(grid, 1, gb.F64)
    D.ImportPhysicalFermionSource(src4, src5)
    gb.SchurRedBlackDiagMooeeSolve(gb.ConjugateGradient(1e-9, 5000))(D, src5, sol5)
    D.ExportPhysicalFermionSolution(sol5, sol4)
    o, _ = oracle_op("mobius")
    x5, _ = o.schur_solve(o.physical(po.IMPORT_PHYSICAL_SOURCE, G["src4"]), 1e-9, 5000)
    assert site_err(sol4.export_lex(), o.physical(po.EXPORT_PHYSICAL_SOLUTION, x5)) < 1e-6


@pytest.mark.gpu
def test_dwf_cg_schur_driver():
    """ref: tests/solver/Test_dwf_cg_schur.cc -- SchurRedBlackDiagMooeeSolve(CG)(Ddwf, src, result), unpreconditioned residual,
    and one 4D -> 5D -> 4D propagator column, through the C++ mirror (include/gridb200.hpp)"""
    import subprocess
    exe = os.path.join(os.path.dirname(HERE), "drivers", "Test_dwf_cg_schur")
    assert os.path.exists(exe), f"{exe} missing: run make -C grid_b200"
    p = subprocess.run([exe, "--grid", "8.8.8.8", "--Ls", "8"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "true unprec resid" in p.stdout and "PASS" in p.stdout


@pytest.mark.gpu
def test_cuda_unpreconditioned_cg_on_mdagm():
    """MdagMLinearOperator (ref: LinearOperator.h:74-105) through ConjugateGradient's generic path: Mdag M x = b on the full grid,
    checked with the oracle's M / Mdag on the host, and against the Schur solve of M y = b (x = M^-1 Mdag^-1 b => M x' = ... )."""
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, "wilson", gb.F64)
    b = gb.LatticeFermion(grid, 1, gb.F64).import_lex(G["src4"])
    x = gb.LatticeFermion(grid, 1, gb.F64).zero()
    CG = gb.ConjugateGradient(1e-9, 5000)
    CG(gb.MdagMLinearOperator(D), b, x)
    o, _ = oracle_op("wilson")
    xs = x.export_lex()
    r = o.apply(po.OP_MDAG, o.apply(po.OP_M, xs)) - G["src4"]
    assert np.linalg.norm(r) / np.linalg.norm(G["src4"]) < 1e-8
    assert CG.IterationsToComplete > 5 and CG.TrueResidual < 1e-8
