"""SURVEY 8(f) row 2 -- single hop legs and the fermion force terms (full-grid): DhopDir, DhopDeriv, MDeriv
(ref: WilsonFermion5DImplementation.h:183-275, WilsonImpl.h:173-238, CayleyFermion5DImplementation.h:347-360).

 * CPU: the oracle reproduces the compiled reference's outputs (tests/golden/next_golden.npz; live on a second lattice where
   oracle/_ref exists), the eight DhopDir legs sum to Dhop, and MDeriv predicts the change of S = |M phi|^2 under a small
   change of the links -- the reference's own force test (tests/forces/Test_dwf_force.cc:60-141) restated.
 * GPU: the CUDA path reproduces the fixtures (fp64 <= 2e-13, fp32 <= 4e-6) and the sum-of-legs identity.
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
OPS = {"wilson": dict(kind=0, Ls=1, b=1.0, c=0.0, src="src4"), "mobius": dict(kind=1, Ls=LS, b=1.5, c=0.5, src="src5")}


def rel_err(a, b):
    return float(np.max(np.abs(a.astype(np.complex128) - b)) / np.max(np.abs(b)))


def oracle_op(name, prec=1, U=None, dims=DIMS):
    cfg = OPS[name]
    o = po.OracleOp(cfg["kind"], dims, cfg["Ls"], mass=0.1, M5=1.8, b=cfg["b"], c=cfg["c"], prec=prec)
    o.import_gauge(G["U"] if U is None else U)
    return o


@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_force_terms_match_reference_outputs(name):
    o = oracle_op(name)
    A, B = G[OPS[name]["src"]], N[f"{name}/src_b"]
    for which in (0, 1):
        for dag in (0, 1):
            assert rel_err(o.deriv(which, A, B, dag), N[f"{name}/deriv/{which}_{dag}"]) < 1e-13, (which, dag)
    if name == "mobius":
        assert rel_err(o.dhop_dir(A, 1, 1), N["mobius/dhop_dir/1_1"]) < 1e-14
        assert rel_err(o.dhop_dir(A, 3, -1), N["mobius/dhop_dir/3_-1"]) < 1e-14


@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_eight_legs_sum_to_dhop(name):
    o = oracle_op(name)
    x = G[OPS[name]["src"]]
    total = sum(o.dhop_dir(x, d, s) for d in range(4) for s in (1, -1))
    assert rel_err(total, o.apply(po.OP_DHOP, x)) < 1e-14
    # a leg only reads the neighbour in its own direction: it vanishes when that neighbour slice is zeroed
    y = x.reshape(DIMS[3], DIMS[2], DIMS[1], DIMS[0], -1).copy()
    y[:, :, :, 1] = 0                                     # kill x = 1
    leg = o.dhop_dir(y.reshape(x.shape), 0, 1).reshape(y.shape)
    assert np.count_nonzero(leg[:, :, :, 0]) == 0          # sites at x = 0 read x + 1 = 1


def _ta(M):
    """Ta: traceless anti-Hermitian part (ref: Grid/tensors/Tensor_Ta.h)"""
    A = 0.5 * (M - np.conj(np.swapaxes(M, -1, -2)))
    tr = np.trace(A, axis1=-2, axis2=-1) / 3.0
    return A - tr[..., None, None] * np.eye(3)


@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_mderiv_predicts_the_action_change(name):
    """tests/forces/Test_dwf_force.cc:60-141: S = |M phi|^2; U' = exp(dt P) U; S' - S == dt sum tr(P 2 Ta(UdSdU)) + O(dt^2)."""
    from scipy.linalg import expm
    rng = np.random.default_rng(7)
    U = G["U"]
    o = oracle_op(name)
    phi = G[OPS[name]["src"]]
    Mphi = o.apply(po.OP_M, phi)
    S = np.vdot(Mphi, Mphi).real
    UdSdU = o.deriv(1, Mphi, phi, 0) + o.deriv(1, phi, Mphi, 1)
    P = _ta(rng.normal(size=U.shape) + 1j * rng.normal(size=U.shape))     # momenta: traceless anti-Hermitian
    dt = 1e-5
    Up = np.einsum("smij,smjk->smik", np.array([[expm(dt * P[s, m]) for m in range(4)] for s in range(U.shape[0])]), U)
    o2 = oracle_op(name, U=Up)
    Mp = o2.apply(po.OP_M, phi)
    dS = np.vdot(Mp, Mp).real - S
    pred = dt * np.einsum("smij,smji->", P, 2.0 * _ta(UdSdU)).real
    assert abs(dS - pred) < 2e-3 * abs(pred), (dS, pred)


@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_even_odd_force_matches_reference_outputs_and_predicts_the_action_change(name):
    """SchurDifferentiableOperator::MpcDeriv / MpcDagDeriv (ref: EvenOddSchurDifferentiable.h:52-137) against the stored reference
    outputs, then the even-odd analogue of the force test: S = |Mpc phi_o|^2, dS == dt sum tr(P 2 Ta(MpcDeriv + MpcDagDeriv))."""
    from scipy.linalg import expm
    o = oracle_op(name)
    Ls = OPS[name]["Ls"]
    uo, vo = po.pick_checkerboard(DIMS, Ls, 1, G[OPS[name]["src"]]), po.pick_checkerboard(DIMS, Ls, 1, N[f"{name}/src_b"])
    assert rel_err(o.deriv_eo(2, uo, vo), N[f"{name}/mpc_deriv/0"]) < 1e-13
    assert rel_err(o.deriv_eo(3, uo, vo), N[f"{name}/mpc_deriv/1"]) < 1e-13
    # MeoDeriv / MoeDeriv only write the sites of U's parity
    ue = po.pick_checkerboard(DIMS, Ls, 0, G[OPS[name]["src"]])
    m = o.deriv_eo(0, ue, vo)
    par = np.indices(DIMS[::-1]).sum(axis=0).reshape(-1) & 1          # lexicographic site parity
    assert np.count_nonzero(m[par == 1]) == 0 and np.count_nonzero(m[par == 0]) > 0
    phi = uo
    Mphi = o.apply(po.OP_MPC, phi, cb_in=1)
    S = np.vdot(Mphi, Mphi).real
    UdSdU = o.deriv_eo(2, Mphi, phi) + o.deriv_eo(3, phi, Mphi)
    rng = np.random.default_rng(8)
    U = G["U"]
    P = _ta(rng.normal(size=U.shape) + 1j * rng.normal(size=U.shape))
    dt = 1e-5
    Up = np.einsum("smij,smjk->smik", np.array([[expm(dt * P[s, m]) for m in range(4)] for s in range(U.shape[0])]), U)
    Mp = oracle_op(name, U=Up).apply(po.OP_MPC, phi, cb_in=1)
    dS = np.vdot(Mp, Mp).real - S
    pred = dt * np.einsum("smij,smji->", P, 2.0 * _ta(UdSdU)).real
    assert abs(dS - pred) < 2e-3 * abs(pred), (dS, pred)


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_vs_reference_force_terms(name, prec):
    dims = (4, 6, 8, 4)
    cfg = OPS[name]
    U = syn.hot_gauge(dims, seed=21)
    o = oracle_op(name, prec, U, dims)
    r = pr.RefOp(cfg["kind"], dims, cfg["Ls"], mass=0.1, M5=1.8, b=cfg["b"], c=cfg["c"], prec=prec); r.import_gauge(U)
    x = syn.random_fermion(dims, cfg["Ls"], seed=3).astype(po._cdtype(prec)); y = syn.random_fermion(dims, cfg["Ls"], seed=5).astype(po._cdtype(prec))
    tol = 1e-13 if prec else 1e-6
    for d in range(4):
        for s in (1, -1):
            assert rel_err(o.dhop_dir(x, d, s), r.dhop_dir(x, d, s).astype(np.complex128)) < tol
    for which in (0, 1):
        for dag in (0, 1):
            assert rel_err(o.deriv(which, x, y, dag), r.deriv(which, x, y, dag).astype(np.complex128)) < 4 * tol
    xo, yo = po.pick_checkerboard(dims, cfg["Ls"], 1, x), po.pick_checkerboard(dims, cfg["Ls"], 1, y)
    for which in (2, 3):
        assert rel_err(o.deriv_eo(which, xo, yo), r.deriv_eo(which, xo, yo).astype(np.complex128)) < 8 * tol


# ---------------------------------------------------------------------------------------------- GPU
def _device_op(gb, grid, name, prec):
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(G["U"])
    if name == "wilson":
        return gb.WilsonFermion(Umu, grid, 0.1)
    return gb.MobiusFermion(Umu, grid, LS, 0.1, 1.8, 1.5, 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("prec_name", ["f64", "f32"])
@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_cuda_dhop_dir_and_force_terms(name, prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol = 2e-13 if prec == gb.F64 else 4e-6
    dt = gb._cdtype(prec)
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, name, prec)
    Ls = OPS[name]["Ls"]
    A = gb.LatticeFermion(grid, Ls, prec).import_lex(G[OPS[name]["src"]].astype(dt))
    B = gb.LatticeFermion(grid, Ls, prec).import_lex(N[f"{name}/src_b"].astype(dt))
    out, total = gb.LatticeFermion(grid, Ls, prec), gb.LatticeFermion(grid, Ls, prec).zero()
    o = oracle_op(name)
    for d in range(4):
        for s in (1, -1):
            D.DhopDir(A, out, d, s)
            assert rel_err(out.export_lex(), o.dhop_dir(G[OPS[name]["src"]], d, s)) < tol, (d, s)
            gb.axpy(total, 1.0, out, total)
    D.Dhop(A, out, 0)
    assert rel_err(total.export_lex(), out.export_lex().astype(np.complex128)) < 4 * tol          # the eight legs sum to Dhop
    if name == "mobius":
        D.DhopDir(A, out, 1, 1)
        assert rel_err(out.export_lex(), N["mobius/dhop_dir/1_1"]) < tol
    mat = gb.LatticeGaugeField(grid, prec)
    for which, meth in ((0, D.DhopDeriv), (1, D.MDeriv)):
        for dag in (0, 1):
            meth(mat, A, B, dag)
            assert rel_err(mat.export_lex(dtype=dt), N[f"{name}/deriv/{which}_{dag}"]) < 4 * tol, (which, dag)
    # the tuned kernels are back in charge afterwards: a plain hop still matches the oracle
    D.Dhop(A, out, 1)
    assert rel_err(out.export_lex(), o.apply(po.OP_DHOP, G[OPS[name]["src"]], dag=1)) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("prec_name", ["f64", "f32"])
@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_cuda_even_odd_force_terms(name, prec_name):
    import grid_b200 as gb
    prec = gb.F64 if prec_name == "f64" else gb.F32
    tol = 4e-13 if prec == gb.F64 else 8e-6
    dt = gb._cdtype(prec)
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, name, prec)
    Ls = OPS[name]["Ls"]
    A = gb.LatticeFermion(grid, Ls, prec).import_lex(G[OPS[name]["src"]].astype(dt))
    B = gb.LatticeFermion(grid, Ls, prec).import_lex(N[f"{name}/src_b"].astype(dt))
    uo, vo, ue = (gb.LatticeFermion(grid, Ls, prec, gb.HALF) for _ in range(3))
    gb.pickCheckerboard(gb.Odd, uo, A); gb.pickCheckerboard(gb.Odd, vo, B); gb.pickCheckerboard(gb.Even, ue, A)
    S = gb.SchurDifferentiableOperator(D)
    F = gb.LatticeGaugeField(grid, prec)
    S.MpcDeriv(F, uo, vo)
    assert rel_err(F.export_lex(dtype=dt), N[f"{name}/mpc_deriv/0"]) < tol
    S.MpcDagDeriv(F, uo, vo)
    assert rel_err(F.export_lex(dtype=dt), N[f"{name}/mpc_deriv/1"]) < tol
    # MeoDeriv against the oracle; it must leave the Odd sites of the force field alone
    o = oracle_op(name)
    before = F.export_lex(dtype=dt)
    D.MeoDeriv(F, ue, vo, 0)
    got = F.export_lex(dtype=dt)
    want = o.deriv_eo(0, po.pick_checkerboard(DIMS, Ls, 0, G[OPS[name]["src"]]), po.pick_checkerboard(DIMS, Ls, 1, N[f"{name}/src_b"]))
    par = np.indices(DIMS[::-1]).sum(axis=0).reshape(-1) & 1
    assert rel_err(got[par == 0], want[par == 0]) < tol
    assert np.array_equal(got[par == 1], before[par == 1])


@pytest.mark.gpu
def test_cuda_mdir_all():
    """Mdir / MdirAll (the directional pieces multigrid coarsening reads, ref: CayleyFermion5DImplementation.h:331-344): out[p] is
    leg p of Dhop applied to Meo5D psi; the eight of them sum to Meooe-style Dhop(Meooe5D psi)."""
    import grid_b200 as gb
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, "mobius", gb.F64)
    psi = gb.LatticeFermion(grid, LS, gb.F64).import_lex(G["src5"])
    outs = [gb.LatticeFermion(grid, LS, gb.F64) for _ in range(8)]
    D.MdirAll(psi, outs)
    o = oracle_op("mobius")
    Din = o.apply(po.OP_MEOOE5D, G["src5"])
    for p, f in enumerate(outs):
        assert rel_err(f.export_lex(), o.dhop_dir(Din, p & 3, 1 if p < 4 else -1)) < 2e-13, p
    one = gb.LatticeFermion(grid, LS, gb.F64)
    D.Mdir(psi, one, 2, -1)
    assert np.array_equal(one.export_lex(), outs[6].export_lex())


@pytest.mark.gpu
def test_dwf_force_driver():
    """ref: tests/forces/Test_dwf_force.cc -- dS predicted from MDeriv against the measured change of |M phi|^2, through the C++ mirror"""
    import subprocess
    exe = os.path.join(os.path.dirname(HERE), "drivers", "Test_dwf_force")
    assert os.path.exists(exe), f"{exe} missing: run make -C grid_b200"
    p = subprocess.run([exe, "--grid", "8.8.8.8", "--Ls", "8"], capture_output=True, text=True, timeout=200)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "predict dS" in p.stdout and "PASS" in p.stdout


def _two_flavour_eo_action_and_force(apply_mpc, cg_solve, mpc_deriv, phi_o):
    """TwoFlavourEvenOddPseudoFermionAction (ref: Grid/qcd/action/pseudofermion/TwoFlavourEvenOdd.h:112-182):
    S = phi^dag (Mpc^dag Mpc)^-1 phi ;  deriv: X = (Mpc^dag Mpc)^-1 phi, Y = Mpc X, dSdU = MpcDeriv(Y, X) + MpcDagDeriv(X, Y)."""
    X = cg_solve(phi_o)
    Y = apply_mpc(X)
    S = np.vdot(phi_o, X).real
    return S, mpc_deriv(2, Y, X) + mpc_deriv(3, X, Y)


@pytest.mark.parametrize("name", ["wilson", "mobius"])
def test_oracle_two_flavour_even_odd_pseudofermion_force(name):
    """The caller SURVEY 8b names (HMC pseudofermion actions): with the reference's conventions dS = -dt sum tr(P 2 Ta(dSdU))."""
    from scipy.linalg import expm
    Ls = OPS[name]["Ls"]
    phi_o = po.pick_checkerboard(DIMS, Ls, 1, G[OPS[name]["src"]])

    def run(U):
        o = oracle_op(name, U=U)
        return _two_flavour_eo_action_and_force(lambda x: o.apply(po.OP_MPC, x, cb_in=1), lambda b: o.cg(1, b, 1e-12, 5000)[0], lambda w, a, b: o.deriv_eo(w, a, b), phi_o)
    U = G["U"]
    S, dSdU = run(U)
    rng = np.random.default_rng(9)
    P = _ta(rng.normal(size=U.shape) + 1j * rng.normal(size=U.shape))
    dt = 1e-5
    Up = np.einsum("smij,smjk->smik", np.array([[expm(dt * P[s, m]) for m in range(4)] for s in range(U.shape[0])]), U)
    Sp, _ = run(Up)
    pred = -dt * np.einsum("smij,smji->", P, 2.0 * _ta(dSdU)).real
    assert abs((Sp - S) - pred) < 5e-3 * abs(pred), (Sp - S, pred)


@pytest.mark.gpu
def test_cuda_two_flavour_even_odd_pseudofermion_force():
    """The same action and force through the CUDA path (ConjugateGradient on SchurDifferentiableOperator + MpcDeriv / MpcDagDeriv),
    against the oracle's force for the same field."""
    import grid_b200 as gb
    name = "mobius"
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    D = _device_op(gb, grid, name, gb.F64)
    full = gb.LatticeFermion(grid, LS, gb.F64).import_lex(G["src5"])
    phi, X, Y = (gb.LatticeFermion(grid, LS, gb.F64, gb.HALF) for _ in range(3))
    gb.pickCheckerboard(gb.Odd, phi, full)
    Mpc = gb.SchurDifferentiableOperator(D)
    X.zero()
    gb.ConjugateGradient(1e-12, 5000)(Mpc, phi, X)          # DerivativeSolver(Mpc, PhiOdd, X)
    Mpc.Mpc(X, Y)
    F1, F2 = gb.LatticeGaugeField(grid, gb.F64), gb.LatticeGaugeField(grid, gb.F64)
    Mpc.MpcDeriv(F1, Y, X)
    Mpc.MpcDagDeriv(F2, X, Y)
    got = F1.export_lex() + F2.export_lex()
    o = oracle_op(name)
    phi_o = po.pick_checkerboard(DIMS, LS, 1, G["src5"])
    S, want = _two_flavour_eo_action_and_force(lambda x: o.apply(po.OP_MPC, x, cb_in=1), lambda b: o.cg(1, b, 1e-12, 5000)[0], lambda w, a, b: o.deriv_eo(w, a, b), phi_o)
    assert rel_err(got, want) < 1e-8
    assert abs(gb.innerProduct(phi, X).real - S) < 1e-9 * abs(S)
