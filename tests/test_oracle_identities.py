"""Pins the CPU oracle with the reference's own executable identities.

The reference stores no golden vectors for this path; what its tests assert are identities between
independent implementations.  Each test names the reference check it restates.  All CPU, fp64 unless noted.
"""
import numpy as np
import pytest

from grid_b200 import synthetic as syn
from oracle import pyoracle as po

DIMS = (4, 4, 4, 6)   # small but with unequal extents so index-order bugs show
LS = 6


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


@pytest.fixture(scope="module")
def gauge():
    return syn.hot_gauge(DIMS, seed=11)


@pytest.fixture(scope="module")
def wilson(gauge):
    op = po.OracleOp(0, DIMS, 1, mass=0.1, prec=1)
    op.import_gauge(gauge)
    return op


def make_cayley(gauge, b=1.0, c=0.0, prec=1, Ls=LS, mass=0.1, M5=1.8, phases=None):
    op = po.OracleOp(1, DIMS, Ls, mass=mass, M5=M5, b=b, c=c, prec=prec)
    op.import_gauge(gauge, phases)
    return op


@pytest.fixture(scope="module")
def dwf(gauge):
    return make_cayley(gauge)


@pytest.fixture(scope="module")
def mobius(gauge):
    return make_cayley(gauge, b=1.5, c=0.5)


# ---- ref: benchmarks/Benchmark_wilson.cc:122-145,220-257 ; Benchmark_dwf_fp32.cc:214-245,324-379
@pytest.mark.parametrize("dag", [0, 1])
def test_dhop_equals_naive_cshift_form_4d(gauge, wilson, dag):
    src = syn.random_fermion(DIMS, 1, seed=3)
    assert rel(wilson.apply(po.OP_DHOP, src, dag=dag), po.dhop_naive(DIMS, 1, gauge, src, dag=dag)) < 1e-14


@pytest.mark.parametrize("dag", [0, 1])
def test_dhop_equals_naive_cshift_form_5d(gauge, dwf, dag):
    src = syn.random_fermion(DIMS, LS, seed=4, normalise=True)
    assert rel(dwf.apply(po.OP_DHOP, src, dag=dag), po.dhop_naive(DIMS, LS, gauge, src, dag=dag)) < 1e-14


def test_dhop_fp32_vs_fp64(gauge):
    src = syn.random_fermion(DIMS, LS, seed=4, normalise=True)
    op32 = make_cayley(gauge, prec=0)
    r32 = op32.apply(po.OP_DHOP, src.astype(np.complex64))
    r64 = po.dhop_naive(DIMS, LS, gauge, src)
    # the reference asserts norm2(err) < 1e-4 on a unit-norm source (Benchmark_dwf_fp32.cc:308-321)
    assert np.linalg.norm((r32 - r64).ravel()) ** 2 < 1e-10
    assert rel(r32, r64) < 1e-6


# ---- ref: Benchmark_dwf_fp32.cc:424-446 (Deo + Doe == Dunprec)
@pytest.mark.parametrize("dag", [0, 1])
def test_deo_plus_doe_is_dunprec(dwf, dag):
    src = syn.random_fermion(DIMS, LS, seed=5)
    se, so = po.pick_checkerboard(DIMS, LS, 0, src), po.pick_checkerboard(DIMS, LS, 1, src)
    r_e = dwf.apply(po.OP_DHOP_EO, so, dag=dag)
    r_o = dwf.apply(po.OP_DHOP_OE, se, dag=dag)
    full = np.zeros_like(src)
    po.set_checkerboard(DIMS, LS, 0, full, r_e)
    po.set_checkerboard(DIMS, LS, 1, full, r_o)
    assert rel(full, dwf.apply(po.OP_DHOP, src, dag=dag)) < 1e-14


def test_pick_set_checkerboard_roundtrip():
    src = syn.random_fermion(DIMS, LS, seed=6)
    full = np.zeros_like(src)
    for cb in (0, 1):
        po.set_checkerboard(DIMS, LS, cb, full, po.pick_checkerboard(DIMS, LS, cb, src))
    assert np.array_equal(full, src)
    # parity of a picked site: coordinate-encoded field (ref: tests/Test_stencil.cc:70-80)
    v4 = int(np.prod(DIMS))
    idx = np.arange(v4)
    x = idx % DIMS[0]; y = (idx // DIMS[0]) % DIMS[1]; z = (idx // (DIMS[0] * DIMS[1])) % DIMS[2]; t = idx // (DIMS[0] * DIMS[1] * DIMS[2])
    enc = np.zeros((v4, 4, 3), dtype=np.complex128)
    enc[:, 0, 0] = x * 1000 + y * 100 + z * 10 + t
    enc[:, 0, 1] = (x + y + z + t) & 1
    for cb in (0, 1):
        h = po.pick_checkerboard(DIMS, 1, cb, enc)
        assert np.all(h[:, 0, 1].real == cb)
        # cb-lex order: x/2 fastest
        code = h[:, 0, 0].real.astype(int)
        hx, hy, hz, ht = code // 1000, (code // 100) % 10, (code // 10) % 10, code % 10
        icb = (hx // 2) + (DIMS[0] // 2) * (hy + DIMS[1] * (hz + DIMS[2] * ht))
        assert np.array_equal(icb, np.arange(v4 // 2))


# ---- ref: tests/core/Test_wilson_even_odd.cc:120-158 (adjointness of Meooe)
@pytest.mark.parametrize("opname", ["wilson", "dwf", "mobius"])
def test_meooe_adjoint(opname, request):
    op = request.getfixturevalue(opname)
    Ls = op.Ls
    phi = syn.random_fermion(DIMS, Ls, seed=7, gaussian=True)
    chi = syn.random_fermion(DIMS, Ls, seed=8, gaussian=True)
    phi_e, phi_o = (po.pick_checkerboard(DIMS, Ls, cb, phi) for cb in (0, 1))
    chi_e, chi_o = (po.pick_checkerboard(DIMS, Ls, cb, chi) for cb in (0, 1))
    dchi_o = op.apply(po.OP_MEOOE, chi_e, cb_in=0)       # even -> odd
    dchi_e = op.apply(po.OP_MEOOE, chi_o, cb_in=1)
    dphi_o = op.apply(po.OP_MEOOE_DAG, phi_e, cb_in=0)
    dphi_e = op.apply(po.OP_MEOOE_DAG, phi_o, cb_in=1)
    pDce = po.inner_product(phi_e, dchi_e); pDco = po.inner_product(phi_o, dchi_o)
    cDpe = po.inner_product(chi_e, dphi_e); cDpo = po.inner_product(chi_o, dphi_o)
    assert abs(pDce - np.conj(cDpo)) < 1e-10 * abs(pDce)
    assert abs(pDco - np.conj(cDpe)) < 1e-10 * abs(pDco)


# ---- ref: tests/core/Test_wilson_even_odd.cc:159-196 ; tests/debug/Test_cayley_even_odd.cc:47-113
@pytest.mark.parametrize("opname", ["wilson", "dwf", "mobius"])
def test_mooeeinv_mooee_is_identity(opname, request):
    op = request.getfixturevalue(opname)
    Ls = op.Ls
    phi = syn.random_fermion(DIMS, Ls, seed=9)
    for cb in (0, 1):
        h = po.pick_checkerboard(DIMS, Ls, cb, phi)
        assert rel(op.apply(po.OP_MOOEE_INV, op.apply(po.OP_MOOEE, h)), h) < 1e-13
        assert rel(op.apply(po.OP_MOOEE_INV_DAG, op.apply(po.OP_MOOEE_DAG, h)), h) < 1e-13
    # full-lattice fields too (Mooee is site-local in 4D)
    assert rel(op.apply(po.OP_MOOEE_INV, op.apply(po.OP_MOOEE, phi)), phi) < 1e-13


@pytest.mark.parametrize("opname", ["dwf", "mobius"])
def test_mooee_dag_is_adjoint(opname, request):
    op = request.getfixturevalue(opname)
    a = po.pick_checkerboard(DIMS, LS, 1, syn.random_fermion(DIMS, LS, seed=21, gaussian=True))
    b = po.pick_checkerboard(DIMS, LS, 1, syn.random_fermion(DIMS, LS, seed=22, gaussian=True))
    for fwd, adj in ((po.OP_MOOEE, po.OP_MOOEE_DAG), (po.OP_MOOEE_INV, po.OP_MOOEE_INV_DAG), (po.OP_MEOOE5D, po.OP_MEOOEDAG5D)):
        lhs = po.inner_product(a, op.apply(fwd, b))
        rhs = po.inner_product(op.apply(adj, a), b)
        assert abs(lhs - rhs) < 1e-11 * abs(lhs)


# ---- ref: tests/core/Test_wilson_even_odd.cc:197-224 (MpcDagMpc hermitian) and Test_dwf_even_odd.cc
@pytest.mark.parametrize("opname", ["wilson", "dwf", "mobius"])
def test_mpcdagmpc_hermitian_positive(opname, request):
    op = request.getfixturevalue(opname)
    Ls = op.Ls
    a = po.pick_checkerboard(DIMS, Ls, 1, syn.random_fermion(DIMS, Ls, seed=31, gaussian=True))
    b = po.pick_checkerboard(DIMS, Ls, 1, syn.random_fermion(DIMS, Ls, seed=32, gaussian=True))
    Aa, Ab = op.apply(po.OP_HERMOP, a, cb_in=1), op.apply(po.OP_HERMOP, b, cb_in=1)
    ab, ba = po.inner_product(a, Ab), po.inner_product(b, Aa)
    assert abs(ab - np.conj(ba)) < 1e-11 * abs(ab)
    aa = po.inner_product(a, Aa)
    assert aa.real > 0 and abs(aa.imag) < 1e-11 * aa.real
    # HermOp == MpcDag(Mpc())
    assert rel(op.apply(po.OP_MPC_DAG, op.apply(po.OP_MPC, a, cb_in=1), cb_in=1), Aa) < 1e-14


# ---- ref: tests/core/Test_dwf_even_odd.cc / Test_mobius_even_odd.cc: M == (Mee Meo ; Moe Moo) assembled
@pytest.mark.parametrize("opname", ["wilson", "dwf", "mobius"])
def test_unprec_m_equals_even_odd_assembly(opname, request):
    op = request.getfixturevalue(opname)
    Ls = op.Ls
    psi = syn.random_fermion(DIMS, Ls, seed=41)
    pe, po_ = po.pick_checkerboard(DIMS, Ls, 0, psi), po.pick_checkerboard(DIMS, Ls, 1, psi)
    for (m, meooe, mooee) in ((po.OP_M, po.OP_MEOOE, po.OP_MOOEE), (po.OP_MDAG, po.OP_MEOOE_DAG, po.OP_MOOEE_DAG)):
        r_e = op.apply(mooee, pe) + op.apply(meooe, po_, cb_in=1)
        r_o = op.apply(mooee, po_) + op.apply(meooe, pe, cb_in=0)
        full = np.zeros_like(psi)
        po.set_checkerboard(DIMS, Ls, 0, full, r_e)
        po.set_checkerboard(DIMS, Ls, 1, full, r_o)
        assert rel(full, op.apply(m, psi)) < 1e-13


@pytest.mark.parametrize("opname", ["wilson", "dwf", "mobius"])
def test_mdag_is_adjoint_of_m(opname, request):
    op = request.getfixturevalue(opname)
    a = syn.random_fermion(DIMS, op.Ls, seed=51, gaussian=True)
    b = syn.random_fermion(DIMS, op.Ls, seed=52, gaussian=True)
    lhs = po.inner_product(a, op.apply(po.OP_M, b))
    rhs = po.inner_product(op.apply(po.OP_MDAG, a), b)
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)


# ---- explicit 5D structure: M psi = (b Dw + 1) psi_s + (c Dw - 1)(P- psi_{s+1} + P+ psi_{s-1}), -m on the wrap
#      ref: CayleyFermion5DImplementation.h:274-286 ; DWFSlow.h:120-194 (reference's own naive DWF)
@pytest.mark.parametrize("b,c", [(1.0, 0.0), (1.5, 0.5)])
def test_cayley_m_against_explicit_formula(gauge, b, c):
    mass, M5 = 0.1, 1.8
    op = make_cayley(gauge, b=b, c=c, mass=mass, M5=M5)
    psi = syn.random_fermion(DIMS, LS, seed=61)
    v4 = int(np.prod(DIMS))
    p5 = psi.reshape(v4, LS, 4, 3)

    def dw(f):  # Dw = Dhop + (4 - M5), via the independent naive form
        return po.dhop_naive(DIMS, LS, gauge, f.reshape(-1, 4, 3)).reshape(v4, LS, 4, 3) + (4.0 - M5) * f

    Pm = np.zeros((4, 1)); Pm[2:] = 1
    Pp = np.zeros((4, 1)); Pp[:2] = 1
    hop = np.zeros_like(p5)
    for s in range(LS):
        up = p5[:, (s + 1) % LS] * Pm * (-mass if s == LS - 1 else 1.0)
        dn = p5[:, (s - 1) % LS] * Pp * (-mass if s == 0 else 1.0)
        hop[:, s] = up + dn
    expect = b * dw(p5) + p5 + c * dw(hop) - hop
    assert rel(op.apply(po.OP_M, psi).reshape(v4, LS, 4, 3), expect) < 1e-13


# ---- free field: plane waves diagonalise Dhop on unit gauge. ref: tests/core/Test_fft.cc:176-280 (momentum-space DWF)
def gamma_matrices():
    """Grid's chiral basis, read off Grid/qcd/spin/Gamma.h:552-558 (X), :420-426 (Y), :288-294 (Z), :156-162 (T)."""
    i = 1j
    gx = np.array([[0, 0, 0, i], [0, 0, i, 0], [0, -i, 0, 0], [-i, 0, 0, 0]])
    gy = np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=complex)
    gz = np.array([[0, 0, i, 0], [0, 0, 0, -i], [-i, 0, 0, 0], [0, i, 0, 0]])
    gt = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=complex)
    return [gx, gy, gz, gt]


def test_gamma_algebra():
    g = gamma_matrices()
    for mu in range(4):
        for nu in range(4):
            acomm = g[mu] @ g[nu] + g[nu] @ g[mu]
            assert np.allclose(acomm, 2 * np.eye(4) * (mu == nu))
    g5 = g[0] @ g[1] @ g[2] @ g[3]
    assert np.allclose(g5, np.diag([1, 1, -1, -1]))   # ref: Gamma.h:90-96


@pytest.mark.parametrize("dag", [0, 1])
def test_free_field_plane_wave(dag):
    U = syn.unit_gauge(DIMS)
    op = po.OracleOp(0, DIMS, 1, mass=0.1, prec=1)
    op.import_gauge(U)
    v4 = int(np.prod(DIMS))
    idx = np.arange(v4)
    coords = [idx % DIMS[0], (idx // DIMS[0]) % DIMS[1], (idx // (DIMS[0] * DIMS[1])) % DIMS[2], idx // (DIMS[0] * DIMS[1] * DIMS[2])]
    n = (1, 3, 2, 5)
    p = [2 * np.pi * n[mu] / DIMS[mu] for mu in range(4)]
    phase = np.exp(1j * sum(p[mu] * coords[mu] for mu in range(4)))
    rng = np.random.default_rng(5)
    u = rng.standard_normal((4, 3)) + 1j * rng.standard_normal((4, 3))
    psi = phase[:, None, None] * u[None]
    g = gamma_matrices()
    sgn = -1.0 if dag else 1.0
    # Dhop = -1/2 sum_mu [(1 -+ g)e^{ip} + (1 +- g)e^{-ip}] = -sum_mu [cos p_mu -+ i g_mu sin p_mu]
    K = -sum(np.cos(p[mu]) * np.eye(4) - sgn * 1j * np.sin(p[mu]) * g[mu] for mu in range(4))
    expect = phase[:, None, None] * (K @ u)[None]
    assert rel(op.apply(po.OP_DHOP, psi, dag=dag), expect) < 1e-13


# ---- boundary phases (anti-periodic in t): DoubleStore ref WilsonImpl.h:143-170
def test_antiperiodic_time_boundary(gauge):
    phases = [1, 1, 1, -1]
    op = po.OracleOp(0, DIMS, 1, mass=0.1, prec=1)
    op.import_gauge(gauge, phases)
    # equivalent: multiply the t-links on the last time slice by -1 and use periodic BCs
    v4 = int(np.prod(DIMS))
    t = np.arange(v4) // (DIMS[0] * DIMS[1] * DIMS[2])
    U2 = gauge.copy()
    U2[t == DIMS[3] - 1, 3] *= -1
    src = syn.random_fermion(DIMS, 1, seed=71)
    assert rel(op.apply(po.OP_DHOP, src), po.dhop_naive(DIMS, 1, U2, src)) < 1e-14


def test_doubled_links_layout(gauge, wilson):
    uds = wilson.doubled()
    assert np.allclose(uds[:, :4], -0.5 * gauge)
    v4 = int(np.prod(DIMS))
    idx = np.arange(v4)
    x = idx % DIMS[0]
    xm = (x - 1) % DIMS[0]
    back = idx - x + xm
    assert np.allclose(uds[:, 4], -0.5 * np.conj(np.swapaxes(gauge[back, 0], 1, 2)))


# ---- solvers. ref: tests/Test_dwf_mixedcg_prec.cc:136-215 ; ConjugateGradient.h:68-257
def test_cg_solves_and_reports_true_residual(dwf):
    src = po.pick_checkerboard(DIMS, LS, 1, syn.random_fermion(DIMS, LS, seed=81))
    sol, info = dwf.cg(1, src, 1e-8, 5000)
    assert info["converged"] == 1 and 0 < info["iterations"] < 5000
    r = dwf.apply(po.OP_HERMOP, sol, cb_in=1) - src
    tr = np.linalg.norm(r.ravel()) / np.linalg.norm(src.ravel())
    assert abs(tr - info["true_residual"]) < 1e-3 * tr and tr < 1e-7
    # restart from the converged solution: "guess is converged already" path (ConjugateGradient.h:129-135)
    _, info2 = dwf.cg(1, src, 1e-6, 5000, guess=sol)
    assert info2["iterations"] == 0


def test_mixed_cg_matches_double_cg(gauge):
    """ref: tests/Test_dwf_mixedcg_prec.cc:212-215 asserts |x_mixed - x_double|^2 < 1e-4"""
    op_d = make_cayley(gauge, prec=1)
    op_f = make_cayley(gauge, prec=0)
    src = po.pick_checkerboard(DIMS, LS, 1, syn.random_fermion(DIMS, LS, seed=91))
    x_d, info_d = op_d.cg(1, src, 1e-8, 10000)
    x_m, info_m = po.mixed_cg(op_d, op_f, 1, src, 1e-8, 10000, 50)
    assert info_m["converged"] == 1 and info_m["inner"] > 0 and info_m["outer"] >= 1
    assert np.linalg.norm((x_m - x_d).ravel()) ** 2 < 1e-4
    assert rel(x_m, x_d) < 1e-6
    assert info_m["true_residual"] < 1e-7


def test_cayley_coefficients_closed_forms():
    """SURVEY appendix A.4 closed forms for s-independent b,c (ref: CayleyFermion5DImplementation.h:486-530)."""
    Ls, m, M5, b, c = 8, 0.01, 1.8, 1.5, 0.5
    k = po.OracleOp(1, DIMS, Ls, mass=m, M5=M5, b=b, c=c).coeffs()
    bee, cee = b * (4 - M5) + 1, 1 - c * (4 - M5)
    assert np.allclose(k["bee"], bee) and np.allclose(k["cee"], cee)
    q = cee / bee
    for s in range(Ls - 1):
        assert np.isclose(k["lee"][s], -q) and np.isclose(k["uee"][s], -q)
        assert np.isclose(k["leem"][s], m * q * q ** s) and np.isclose(k["ueem"][s], m * q ** (s + 1))
    assert np.isclose(k["dee"][Ls - 1], bee + m * cee * q ** (Ls - 1))
