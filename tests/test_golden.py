"""Golden-vector tests.  tests/golden/dirac_golden.npz holds inputs and OUTPUTS OF THE REFERENCE ITSELF (unmodified
paboyle/Grid compiled from /root/reference; generator: tests/golden/make_golden.py).  It travels with the repository, so
these run where neither /root/reference nor oracle/_ref exists.

 * CPU (not gpu): the oracle reproduces every stored reference output (fp64: <= 1e-13 per site; CG: same iteration count).
 * GPU (-m gpu) : the CUDA path through the C ABI reproduces them too (fp64 <= 1e-13 hop / 2e-13 composite, fp32 <= 1e-6 /
   4e-6 as in tests/test_gpu_parity.py), CG iteration count within +-2 %, same true residual.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dirac_golden.npz"))
DIMS, LS = tuple(int(x) for x in G["dims"]), int(G["Ls"])
OPS = {"wilson": dict(kind=0, Ls=1, b=1.0, c=0.0, phases=None, src="src4"),
       "wilson_apbc": dict(kind=0, Ls=1, b=1.0, c=0.0, phases=[1.0, 1.0, 1.0, -1.0], src="src4"),
       "dwf": dict(kind=1, Ls=LS, b=1.0, c=0.0, phases=None, src="src5"),
       "mobius": dict(kind=1, Ls=LS, b=1.5, c=0.5, phases=None, src="src5")}
ENTRY = dict(DHOP=po.OP_DHOP, M=po.OP_M, MDAG=po.OP_MDAG, DW=po.OP_DW, DHOP_OE=po.OP_DHOP_OE, DHOP_EO=po.OP_DHOP_EO, MEOOE=po.OP_MEOOE,
             MEOOE_DAG=po.OP_MEOOE_DAG, MOOEE=po.OP_MOOEE, MOOEE_DAG=po.OP_MOOEE_DAG, MOOEE_INV=po.OP_MOOEE_INV,
             MOOEE_INV_DAG=po.OP_MOOEE_INV_DAG, MPC=po.OP_MPC, MPC_DAG=po.OP_MPC_DAG, HERMOP=po.OP_HERMOP)
METHOD = dict(DHOP="Dhop", M="M", MDAG="Mdag", DW="DW", DHOP_OE="DhopOE", DHOP_EO="DhopEO", MEOOE="Meooe", MEOOE_DAG="MeooeDag",
              MOOEE="Mooee", MOOEE_DAG="MooeeDag", MOOEE_INV="MooeeInv", MOOEE_INV_DAG="MooeeInvDag")


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)))


def cases(name):
    """(key, entry, dag, cb_in or None) for every stored operator output of `name`"""
    out = []
    for key in G.files:
        parts = key.split("/")
        if parts[0] != name or parts[1] not in ENTRY or key.endswith("_f32"):
            continue
        tag = parts[2]
        dag = int(tag[3:]) if tag.startswith("dag") else 0
        cb = int(tag[2:]) if tag.startswith("cb") else (0 if parts[1] == "DHOP_OE" else 1 if parts[1] == "DHOP_EO" else None)
        out.append((key, parts[1], dag, cb))
    return out


def test_fixture_is_complete():
    assert DIMS == (4, 4, 4, 4) and LS == 4
    assert [len(cases(n)) for n in OPS] == [19, 19, 21, 13]
    U = G["U"]
    assert np.allclose(np.einsum("smij,smkj->smik", U, np.conj(U)), np.eye(3)[None, None], atol=1e-13)   # SU(3) links


# ---------------------------------------------------------------------------------------------- oracle vs the reference's outputs
@pytest.mark.parametrize("name", list(OPS))
def test_oracle_reproduces_reference_outputs(name):
    cfg = OPS[name]
    o = po.OracleOp(cfg["kind"], DIMS, cfg["Ls"], mass=float(G["mass"]), M5=float(G["M5"]), b=cfg["b"], c=cfg["c"], prec=1)
    o.import_gauge(G["U"], cfg["phases"])
    src = G[cfg["src"]]
    for cb in (0, 1):
        if f"{name}/pick/cb{cb}" in G.files:
            assert np.array_equal(po.pick_checkerboard(DIMS, cfg["Ls"], cb, src), G[f"{name}/pick/cb{cb}"])
    for key, entry, dag, cb in cases(name):
        x = src if cb is None else po.pick_checkerboard(DIMS, cfg["Ls"], cb, src)
        e = site_err(o.apply(ENTRY[entry], x, dag=dag, cb_in=cb or 0), G[key])
        assert e < 1e-13, (key, e)
    x, info = o.cg(1, po.pick_checkerboard(DIMS, cfg["Ls"], 1, src), 1e-8, 5000)
    assert abs(info["iterations"] - int(G[f"{name}/cg/iterations"])) <= 1      # threaded reductions: order dependent in the last bit
    assert abs(info["true_residual"] - float(G[f"{name}/cg/true_residual"])) < 1e-3 * float(G[f"{name}/cg/true_residual"])
    assert site_err(x, G[f"{name}/cg/solution"]) < 1e-9


def test_oracle_fp32_hop_and_mixed_cg_match_reference():
    of = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, prec=0); of.import_gauge(G["U"])
    od = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, prec=1); od.import_gauge(G["U"])
    assert site_err(of.apply(po.OP_DHOP, G["src5"].astype(np.complex64)), G["dwf/DHOP/dag0_f32"]) < 1e-6
    x, info = po.mixed_cg(od, of, 1, po.pick_checkerboard(DIMS, LS, 1, G["src5"]), 1e-8, 10000, 50)
    assert info["outer"] == int(G["dwf/mixed_cg/outer"])
    assert abs(info["inner"] - int(G["dwf/mixed_cg/inner"])) <= max(3, 0.05 * int(G["dwf/mixed_cg/inner"]))   # see test_oracle_vs_reference
    assert site_err(x, G["dwf/mixed_cg/solution"]) < 1e-6


# ---------------------------------------------------------------------------------------------- CUDA path vs the reference's outputs
def _device_op(gb, ctx, name, prec, grid=None):
    cfg = OPS[name]
    grid = grid or gb.GridCartesian(ctx, DIMS)
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(G["U"])
    mass, M5 = float(G["mass"]), float(G["M5"])
    if cfg["kind"] == 0:
        D = gb.WilsonFermion(Umu, grid, mass, cfg["phases"])
    elif cfg["b"] == 1.0:
        D = gb.DomainWallFermion(Umu, grid, cfg["Ls"], mass, M5, cfg["phases"])
    else:
        D = gb.MobiusFermion(Umu, grid, cfg["Ls"], mass, M5, cfg["b"], cfg["c"], cfg["phases"])
    return grid, D


@pytest.mark.gpu
@pytest.mark.parametrize("prec_name", ["f64", "f32"])
@pytest.mark.parametrize("name", list(OPS))
def test_cuda_path_reproduces_reference_outputs(name, prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol_hop, tol_comp = (1e-13, 2e-13) if prec == gb.F64 else (1e-6, 4e-6)
    ctx = gb.Context(0)
    cfg = OPS[name]
    grid, D = _device_op(gb, ctx, name, prec)
    lin = gb.SchurDiagMooeeOperator(D)
    src = G[cfg["src"]].astype(gb._cdtype(prec))
    full = gb.LatticeFermion(grid, cfg["Ls"], prec).import_lex(src)
    for key, entry, dag, cb in cases(name):
        kind = gb.FULL if cb is None else gb.HALF
        fin, out = gb.LatticeFermion(grid, cfg["Ls"], prec, kind), gb.LatticeFermion(grid, cfg["Ls"], prec, kind)
        if cb is None:
            fin.import_lex(src)
        else:
            gb.pickCheckerboard(cb, fin, full)
        if entry in ("DHOP", "DHOP_OE", "DHOP_EO", "DW"):
            getattr(D, METHOD[entry])(fin, out, dag)
            tol = tol_hop
        elif entry in ("MPC", "MPC_DAG", "HERMOP"):
            {"MPC": lin.Mpc, "MPC_DAG": lin.MpcDag, "HERMOP": lin.HermOp}[entry](fin, out)
            tol = 3 * tol_comp
        else:
            getattr(D, METHOD[entry])(fin, out)
            tol = tol_comp
        e = site_err(out.export_lex(), G[key])
        assert e < tol, (key, e)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(OPS))
def test_cuda_cg_matches_reference_iterations_and_residual(name):
    import grid_b200 as gb
    ctx = gb.Context(0)
    cfg = OPS[name]
    grid, D = _device_op(gb, ctx, name, gb.F64)
    full = gb.LatticeFermion(grid, cfg["Ls"], gb.F64).import_lex(G[cfg["src"]])
    src, sol = gb.LatticeFermion(grid, cfg["Ls"], gb.F64, gb.HALF), gb.LatticeFermion(grid, cfg["Ls"], gb.F64, gb.HALF).zero()
    gb.pickCheckerboard(gb.Odd, src, full)
    cg = gb.ConjugateGradient(1e-8, 5000)
    cg(gb.SchurDiagMooeeOperator(D), src, sol)
    ref_it, ref_tr = int(G[f"{name}/cg/iterations"]), float(G[f"{name}/cg/true_residual"])
    assert abs(cg.IterationsToComplete - ref_it) <= max(1, 0.02 * ref_it), (cg.IterationsToComplete, ref_it)
    assert abs(cg.TrueResidual - ref_tr) < 0.05 * ref_tr
    assert site_err(sol.export_lex(), G[f"{name}/cg/solution"]) < 1e-7


@pytest.mark.gpu
def test_cuda_mixed_cg_matches_reference():
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid, Dd = _device_op(gb, ctx, "dwf", gb.F64)
    _, Df = _device_op(gb, ctx, "dwf", gb.F32, grid)
    full = gb.LatticeFermion(grid, LS, gb.F64).import_lex(G["src5"])
    src, sol = gb.LatticeFermion(grid, LS, gb.F64, gb.HALF), gb.LatticeFermion(grid, LS, gb.F64, gb.HALF).zero()
    gb.pickCheckerboard(gb.Odd, src, full)
    mcg = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd))
    mcg(src, sol)
    assert mcg.TotalOuterIterations == int(G["dwf/mixed_cg/outer"])
    assert abs(mcg.TotalInnerIterations - int(G["dwf/mixed_cg/inner"])) <= max(3, 0.05 * int(G["dwf/mixed_cg/inner"]))
    assert mcg.TrueResidual < 1e-8 * 10
    assert site_err(sol.export_lex(), G["dwf/mixed_cg/solution"]) < 1e-6


# ---------------------------------------------------------------------------------------------- improved staggered
def stag_cases():
    out = []
    for key in G.files:
        parts = key.split("/")
        if parts[0] != "stag" or parts[1] not in ENTRY:
            continue
        tag = parts[2]
        dag = int(tag[3:]) if tag.startswith("dag") else 0
        cb = int(tag[2:]) if tag.startswith("cb") else (0 if parts[1] == "DHOP_OE" else 1 if parts[1] == "DHOP_EO" else None)
        out.append((key, parts[1], dag, cb))
    return out


def test_stag_oracle_reproduces_reference_outputs():
    assert len(stag_cases()) == 20
    o = po.StagOracleOp(DIMS, 0.1, prec=1)
    o.import_gauge(G["U"])
    src = G["src_stag"]
    for cb in (0, 1):
        assert np.array_equal(po.pick_checkerboard_sites(DIMS, cb, src), G[f"stag/pick/cb{cb}"])
    for key, entry, dag, cb in stag_cases():
        x = src if cb is None else po.pick_checkerboard_sites(DIMS, cb, src)
        e = site_err(o.apply(ENTRY[entry], x, dag=dag, cb_in=cb or 0), G[key])
        assert e < 1e-13, (key, e)
    x, info = o.cg(1, po.pick_checkerboard_sites(DIMS, 1, src), 1e-8, 5000)
    # ~265 iterations on an ill-conditioned operator: the oracle's threaded reductions are order dependent, +-1 happens
    assert abs(info["iterations"] - int(G["stag/cg/iterations"])) <= max(1, 0.02 * int(G["stag/cg/iterations"]))
    assert site_err(x, G["stag/cg/solution"]) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("prec_name", ["f64", "f32"])
def test_cuda_staggered_reproduces_reference_outputs(prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol = 1e-13 if prec == gb.F64 else 1e-6
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(G["U"])
    D = gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
    lin = gb.SchurStaggeredOperator(D)
    src = G["src_stag"].astype(gb._cdtype(prec))
    full = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(src)
    for key, entry, dag, cb in stag_cases():
        kind = gb.FULL if cb is None else gb.HALF
        fin, out = gb.LatticeStaggeredFermion(grid, 1, prec, kind), gb.LatticeStaggeredFermion(grid, 1, prec, kind)
        if cb is None:
            fin.import_lex(src)
        else:
            gb.pickCheckerboard(cb, fin, full)
            assert np.array_equal(fin.export_lex(), G[f"stag/pick/cb{cb}"].astype(gb._cdtype(prec)))
        if entry in ("DHOP", "DHOP_OE", "DHOP_EO"):
            getattr(D, METHOD[entry])(fin, out, dag)
        elif entry in ("MPC", "HERMOP"):
            {"MPC": lin.Mpc, "HERMOP": lin.HermOp}[entry](fin, out)
        else:
            getattr(D, METHOD[entry])(fin, out)
        e = site_err(out.export_lex(), G[key])
        assert e < (4 * tol if entry in ("MPC", "HERMOP") else tol), (key, e)
    if prec == gb.F64:
        s, sol = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF).zero()
        gb.pickCheckerboard(gb.Odd, s, full)
        cg = gb.ConjugateGradient(1e-8, 5000)
        cg(lin, s, sol)
        ref_it = int(G["stag/cg/iterations"])
        assert abs(cg.IterationsToComplete - ref_it) <= max(1, 0.02 * ref_it), (cg.IterationsToComplete, ref_it)
        # ~265 iterations, residual falls ~10 % per iteration near the end: stopping one iteration apart (allowed: +-2 %)
        # moves the true residual by that much, so the bar is "converged to the same tolerance", not 5 %
        ref_tr = float(G["stag/cg/true_residual"])
        assert cg.TrueResidual < 1e-8 and ref_tr < 1e-8 and 0.6 < cg.TrueResidual / ref_tr < 1.6
        assert site_err(sol.export_lex(), G["stag/cg/solution"]) < 1e-6
