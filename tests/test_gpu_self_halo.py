"""GPU check of the MULTI-RANK code path of the Wilson / domain-wall hopping term on ONE device.

GB_SELF_HALO=<bitmask of dimensions>, read when an operator is created, makes the library treat the chosen undecomposed
dimensions as decomposed: the gauge faces of the double store, `pack_send_kernel` (boundary slices projected with the receiving
leg's projector; ref WilsonCompressor.h:461-511), the stores into the receive buffers (this rank's own: the "peer" of a periodic
dimension of extent 1 in the processor grid is the rank itself, ref Communicator_none.cc), the epoch flags, and then
  * the semi-fused launch `dhop_fast_kernel<LS,DAG,2>` (fp32, z / t splits): local legs, flag acquire, halo legs in one kernel,
  * the interior pass + exterior slabs (`set_overlap(2)`; always for fp64 and x / y splits;
    ref WilsonKernelsImplementation.h:167-285 DhopSiteInt / DhopSiteExt over the surface list),
  * the serial-comms form (`set_overlap(0)`: one all-legs kernel reading the halo buffers; ref WilsonFermion5DImplementation.h:386-411).
The neighbour is this rank, so the result must be the periodic one: every form is compared per site with the fp64 oracle on the
same lattice (<= 1e-6 fp32, <= 1e-13 fp64) -- NOT with the library's own single-rank kernel.
GB_NO_P2P=1 additionally routes the same hops through the NCCL-path code (pack_face_kernel + exchange -> device-to-device copies).
"""
import os

import numpy as np
import pytest

import grid_b200 as gb
from grid_b200 import synthetic as syn
from oracle import pyoracle as po
from test_gpu_parity import site_rel_err, TOL_HOP, TOL_COMPOSITE

pytestmark = pytest.mark.gpu

MASKS = {"z": 4, "t": 8, "zt": 12, "x": 1, "y": 2, "xyzt": 15}
SHAPES = {
    "dwf16": dict(dims=(8, 8, 8, 8), Ls=16, kind="dwf"),
    "mobius8": dict(dims=(16, 4, 6, 4), Ls=8, kind="mobius", b=1.5, c=0.5),
    "wilson": dict(dims=(8, 4, 4, 6), Ls=1, kind="wilson"),
}


def make(ctx, shape, prec, mask, no_p2p=False, phases=None, grid=None):
    dims, Ls, kind = shape["dims"], shape["Ls"], shape["kind"]
    grid = grid or gb.GridCartesian(ctx, dims)
    U = syn.hot_gauge(dims, seed=11)
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(U)
    os.environ["GB_SELF_HALO"] = str(mask)
    if no_p2p:
        os.environ["GB_NO_P2P"] = "1"
    try:
        if kind == "wilson":
            D = gb.WilsonFermion(Umu, grid, 0.1, phases)
        elif kind == "dwf":
            D = gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8, phases)
        else:
            D = gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, shape["b"], shape["c"], phases)
        # the peer-to-peer state is created by the first hop: run one while the environment still says what to do
        f = gb.LatticeFermion(grid, Ls, prec).zero()
        D.Dhop(f, gb.LatticeFermion(grid, Ls, prec), 0)
    finally:
        os.environ.pop("GB_SELF_HALO", None)
        os.environ.pop("GB_NO_P2P", None)
    orc = po.OracleOp(0 if kind == "wilson" else 1, dims, Ls, mass=0.1, M5=1.8, b=shape.get("b", 1.0), c=shape.get("c", 0.0), prec=prec)
    orc.import_gauge(U, phases)
    return grid, D, orc


@pytest.fixture(scope="module")
def ctx():
    c = gb.Context(0)
    yield c
    c.synchronize()


def check_all_entries(grid, D, orc, shape, prec, tag):
    dims, Ls = shape["dims"], shape["Ls"]
    src = syn.random_fermion(dims, Ls, seed=12, dtype=gb._cdtype(prec))
    src64 = src.astype(np.complex128)
    fin, fout = gb.LatticeFermion(grid, Ls, prec).import_lex(src), gb.LatticeFermion(grid, Ls, prec)
    he, ho = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    r = gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    gb.pickCheckerboard(gb.Even, he, fin); gb.pickCheckerboard(gb.Odd, ho, fin)
    for dag in (0, 1):
        D.Dhop(fin, fout, dag)
        assert site_rel_err(fout.export_lex(), orc.apply(po.OP_DHOP, src64, dag=dag)) < TOL_HOP[prec], (tag, "Dhop", dag)
        D.DhopEO(ho, r, dag)
        assert r.Checkerboard() == gb.Even
        assert site_rel_err(r.export_lex(), orc.apply(po.OP_DHOP_EO, po.pick_checkerboard(dims, Ls, 1, src64), dag=dag)) < TOL_HOP[prec], (tag, "DhopEO", dag)
        D.DhopOE(he, r, dag)
        assert site_rel_err(r.export_lex(), orc.apply(po.OP_DHOP_OE, po.pick_checkerboard(dims, Ls, 0, src64), dag=dag)) < TOL_HOP[prec], (tag, "DhopOE", dag)
    D.M(fin, fout)
    assert site_rel_err(fout.export_lex(), orc.apply(po.OP_M, src64)) < TOL_COMPOSITE[prec], (tag, "M")


@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32", "f64"])
@pytest.mark.parametrize("mask", list(MASKS), ids=list(MASKS))
@pytest.mark.parametrize("shape", list(SHAPES), ids=list(SHAPES))
def test_self_halo_hops_match_the_oracle(ctx, shape, mask, prec):
    sh = SHAPES[shape]
    grid, D, orc = make(ctx, sh, prec, MASKS[mask])
    # 1 = default overlapped form (semi-fused launch where it exists; on t splits the column kernel projects and sends the t faces
    # itself, GB_HOP_SENDS_T=0 leaves them to the pack kernel), 2 = interior + exterior slabs, 0 = serial comms
    for overlap, env in ((1, {}), (1, {"GB_HOP_SENDS_T": "0"}), (2, {}), (0, {})):
        D.set_overlap(overlap)
        os.environ.update(env)
        try:
            check_all_entries(grid, D, orc, sh, prec, (shape, mask, overlap, env))
        finally:
            for k in env:
                os.environ.pop(k, None)


@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32", "f64"])
def test_self_halo_nccl_path_code(ctx, prec):
    """GB_NO_P2P: pack_face_kernel + exchange (device-to-device copies for a self neighbour) + interior / exterior or serial."""
    sh = SHAPES["dwf16"]
    grid, D, orc = make(ctx, sh, prec, MASKS["zt"], no_p2p=True)
    for overlap in (1, 0):
        D.set_overlap(overlap)
        check_all_entries(grid, D, orc, sh, prec, ("nccl-path", overlap))


def test_self_halo_antiperiodic_phases(ctx):
    """boundary phases sit on the global boundary links, which with a self halo come through the gauge-face exchange"""
    ph = [1, 1, 1, -1]
    sh = SHAPES["dwf16"]
    grid, D, orc = make(ctx, sh, gb.F32, MASKS["zt"], phases=ph)
    check_all_entries(grid, D, orc, sh, gb.F32, "phases")


def test_self_halo_semi_fused_launch_is_taken(ctx):
    """fp32, t split, Ls 16: the default form is ONE launch per Dhop (the column kernel sends its own t faces: no pack kernel), two
    with the t faces left to the pack kernel (GB_HOP_SENDS_T=0); the interior + exterior form needs more (pack, interior, two slabs
    per parity pair)."""
    sh = SHAPES["dwf16"]
    grid, D, orc = make(ctx, sh, gb.F32, MASKS["t"])
    fin, fout = gb.LatticeFermion(grid, 16, gb.F32).zero(), gb.LatticeFermion(grid, 16, gb.F32)
    D.set_overlap(1)
    n0 = ctx.launch_count(); D.Dhop(fin, fout, 0); n1 = ctx.launch_count()
    os.environ["GB_HOP_SENDS_T"] = "0"
    try:
        D.Dhop(fin, fout, 0); n1b = ctx.launch_count()
    finally:
        os.environ.pop("GB_HOP_SENDS_T", None)
    D.set_overlap(2)
    D.Dhop(fin, fout, 0); n2 = ctx.launch_count()
    assert n1 - n0 == 1, n1 - n0
    assert n1b - n1 == 2, n1b - n1
    assert n2 - n1b > 2, n2 - n1b


def test_self_halo_cg_and_mixed_cg_match_the_oracle(ctx):
    """Schur CG (fp64: interior + exterior hops) and mixed-precision CG (fp32 semi-fused hops inside) through the halo path:
    same iteration count as the oracle's solver on the same fields (+-2 %), same true residual."""
    sh = dict(dims=(4, 4, 8, 8), Ls=8, kind="mobius", b=1.5, c=0.5)
    grid, Dd, orc_d = make(ctx, sh, gb.F64, MASKS["zt"])
    _, Df, orc_f = make(ctx, sh, gb.F32, MASKS["zt"], grid=grid)
    src = syn.random_fermion(sh["dims"], sh["Ls"], seed=21, dtype=np.complex128)
    h = po.pick_checkerboard(sh["dims"], sh["Ls"], gb.Odd, src)
    fsrc = gb.LatticeFermion(grid, sh["Ls"], gb.F64, gb.HALF).import_lex(h)
    fsrc.set_checkerboard(gb.Odd)
    sol = gb.LatticeFermion(grid, sh["Ls"], gb.F64, gb.HALF).zero()
    cg = gb.ConjugateGradient(1e-8, 10000)
    cg(gb.SchurDiagMooeeOperator(Dd), fsrc, sol)
    x_ref, info = orc_d.cg(gb.Odd, h, 1e-8, 10000)
    assert abs(cg.IterationsToComplete - info["iterations"]) <= max(1, 0.02 * info["iterations"])
    assert abs(cg.TrueResidual - info["true_residual"]) < 0.05 * info["true_residual"] + 1e-12
    x = sol.export_lex()
    assert np.linalg.norm((x - x_ref).ravel()) / np.linalg.norm(x_ref.ravel()) < 1e-7
    sol_m = gb.LatticeFermion(grid, sh["Ls"], gb.F64, gb.HALF).zero()
    mcg = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, gb.SchurDiagMooeeOperator(Df), gb.SchurDiagMooeeOperator(Dd))
    mcg(fsrc, sol_m)
    assert mcg.TrueResidual < 1e-7
    xm = sol_m.export_lex()
    assert np.linalg.norm((xm - x_ref).ravel()) / np.linalg.norm(x_ref.ravel()) < 1e-6
    _, minfo = po.mixed_cg(orc_d, orc_f, gb.Odd, h, 1e-8, 10000, 50)
    assert mcg.TotalOuterIterations == minfo["outer"]
    assert abs(mcg.TotalInnerIterations - minfo["inner"]) <= max(3, 0.08 * minfo["inner"])   # see tests/test_gpu_parity.py on the 8 %


@pytest.mark.parametrize("no_p2p", [False, True], ids=["p2p", "nccl-path"])
@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32", "f64"])
@pytest.mark.parametrize("mask", ["z", "t", "zt", "x"])
@pytest.mark.parametrize("shape", ["dwf16", "mobius8", "wilson"])
def test_self_halo_host_dhop_pipelined_on_decomposed_lattices(ctx, shape, mask, prec, no_p2p):
    """gb_op_dhop_host on a z / t / z+t decomposed lattice: the faces the neighbours need (t-slices 0 and Lt-1, z planes 0 and Lz-1
    of every slice, one strided H2D copy per face) go in first, ONE halo exchange, then the slices stream through H2D / hop / D2H
    with every slab hop reading the receive buffers -- against the fp64 oracle per site, both daggers, host precision equal to and
    different from the operator's.  An x split keeps the import + hop + export form (and the same result).
    ref: FermionOperator::Dhop (FermionOperator.h:75-77) on a decomposed Grid; WilsonFermion5DImplementation.h:386-411"""
    sh = SHAPES[shape]
    dims, Ls = sh["dims"], sh["Ls"]
    grid, D, orc = make(ctx, sh, prec, MASKS[mask], no_p2p=no_p2p)
    h = syn.random_fermion(dims, Ls, seed=31, dtype=gb._cdtype(prec))
    for dag in (0, 1):
        ref = orc.apply(po.OP_DHOP, h.astype(np.complex128), dag=dag)
        n0 = ctx.launch_count()
        got = D.Dhop_host(h, np.empty_like(h), dag)
        n1 = ctx.launch_count()
        assert site_rel_err(got, ref) < TOL_HOP[prec], (shape, mask, prec, dag)
        if mask != "x":   # pipelined: one hop launch per t-slice (plus layout kernels and the exchange), not one for the lattice
            assert n1 - n0 >= 3 * dims[3], (n1 - n0, dims)
        # a second call reuses receive buffers of the other epoch parity, a third the first again
        got = D.Dhop_host(h, np.empty_like(h), dag)
        assert site_rel_err(got, ref) < TOL_HOP[prec], (shape, mask, prec, dag, "second call")
    other = np.complex128 if prec == gb.F32 else np.complex64
    got = D.Dhop_host(h.astype(other), np.empty(h.shape, other), 0)
    assert site_rel_err(got, orc.apply(po.OP_DHOP, h.astype(np.complex128), dag=0)) < TOL_HOP[gb.F32], (shape, mask, "host precision differs")
    # the device-resident decomposed hop still agrees after the host calls moved the epochs on
    fin, fout = gb.LatticeFermion(grid, Ls, prec).import_lex(h), gb.LatticeFermion(grid, Ls, prec)
    D.Dhop(fin, fout, 0)
    assert site_rel_err(fout.export_lex(), orc.apply(po.OP_DHOP, h.astype(np.complex128), dag=0)) < TOL_HOP[prec]


def _surface_mask(dims, Ls, mask):
    """sites (x fastest, s innermost of the exported [V4*Ls] order) with at least one leg that leaves the rank in a masked dimension"""
    idx = np.arange(int(np.prod(dims)))
    surf = np.zeros(idx.shape, bool)
    for d in range(4):
        c = idx % dims[d]; idx = idx // dims[d]
        if (mask >> d) & 1:
            surf |= (c == 0) | (c == dims[d] - 1)
    return np.repeat(surf, Ls)


def _site_errs(got, ref):
    a = got.reshape(got.shape[0], -1).astype(np.complex128); b = ref.reshape(ref.shape[0], -1).astype(np.complex128)
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.mark.parametrize("no_p2p", [False, True], ids=["p2p", "nccl-path"])
@pytest.mark.parametrize("prec", [gb.F32, gb.F64], ids=["f32-bf16", "f64-f32"])
@pytest.mark.parametrize("mask", ["t", "zt", "xyzt"])
def test_self_halo_compressed_halos(ctx, mask, prec, no_p2p):
    """gb_op_set_halo_compression: the projected half spinors of every face travel one precision down (fp32 operator: bf16, fp64
    operator: fp32) and are widened by the consuming leg.  Sites without an off-rank leg stay within the operator's own tolerance of
    the fp64 oracle; surface sites within the comms precision (and measurably off the uncompressed result: the compression is on);
    every multi-rank form (overlapped, interior + exterior, serial), the checkerboard hops and the host-pipelined entry point.
    ref: FermionOperatorImpl.h:96-137 (CoeffRealHalfComms), WilsonCompressor.h:244-306, DomainWallVec5dImpl.h:204-206 (...ImplFH / DF)"""
    sh = SHAPES["dwf16"]
    dims, Ls = sh["dims"], sh["Ls"]
    grid, D, orc = make(ctx, sh, prec, MASKS[mask], no_p2p=no_p2p)
    surf = _surface_mask(dims, Ls, MASKS[mask])
    tol_comm = 8e-3 if prec == gb.F32 else 2e-6          # bf16: 2^-9 per component; fp32: 6e-8 per component
    src = syn.random_fermion(dims, Ls, seed=41, dtype=gb._cdtype(prec))
    src64 = src.astype(np.complex128)
    fin, fout = gb.LatticeFermion(grid, Ls, prec).import_lex(src), gb.LatticeFermion(grid, Ls, prec)
    ho, r = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
    gb.pickCheckerboard(gb.Odd, ho, fin)
    D.set_halo_compression(True)
    for overlap in (1, 2, 0):
        D.set_overlap(overlap)
        for dag in (0, 1):
            D.Dhop(fin, fout, dag)
            e = _site_errs(fout.export_lex(), orc.apply(po.OP_DHOP, src64, dag=dag))
            assert e[~surf].max() < TOL_HOP[prec], (mask, prec, overlap, dag, "interior sites must not see the compression")
            assert 20 * TOL_HOP[prec] < e[surf].max() < tol_comm, (mask, prec, overlap, dag, e[surf].max())
        D.DhopEO(ho, r, 0)
        full = gb.LatticeFermion(grid, Ls, prec).zero(); gb.setCheckerboard(full, r)
        eo = orc.apply(po.OP_DHOP_EO, po.pick_checkerboard(dims, Ls, 1, src64))
        ref = np.zeros(src64.shape, eo.dtype); po.set_checkerboard(dims, Ls, 0, ref, eo)
        even = np.linalg.norm(ref.reshape(ref.shape[0], -1), axis=1) > 0
        e = _site_errs(full.export_lex()[even], ref[even])
        assert e[~surf[even]].max() < TOL_HOP[prec] and e[surf[even]].max() < tol_comm, (mask, prec, overlap, "DhopEO")
    D.set_overlap(1)
    e = _site_errs(D.Dhop_host(src, np.empty_like(src), 0), orc.apply(po.OP_DHOP, src64, dag=0))
    assert e[~surf].max() < TOL_HOP[prec] and 20 * TOL_HOP[prec] < e[surf].max() < tol_comm, (mask, prec, "Dhop_host")
    D.set_halo_compression(False)
    D.Dhop(fin, fout, 0)
    assert _site_errs(fout.export_lex(), orc.apply(po.OP_DHOP, src64, dag=0)).max() < TOL_HOP[prec], "compression off again"


def test_self_halo_mixed_cg_with_compressed_halo_inner_operator(ctx):
    """the reference's Test_dwf_mixedcg_prec_halfcomms.cc:71-96 (compiled out there at :33-34): MixedPrecisionConjugateGradient whose
    inner fp32 operator exchanges half-precision halos still reaches the fp64 tolerance, because the outer defect correction is done
    with the uncompressed fp64 operator; the solution agrees with the plain fp64 CG"""
    sh = SHAPES["mobius8"]
    dims, Ls = sh["dims"], sh["Ls"]
    grid, Dd, orc = make(ctx, sh, gb.F64, MASKS["zt"])
    _, Df, _ = make(ctx, sh, gb.F32, MASKS["zt"], grid=grid)
    Df.set_halo_compression(True)
    src = syn.random_fermion(dims, Ls, seed=43)
    so = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
    gb.pickCheckerboard(gb.Odd, so, gb.LatticeFermion(grid, Ls, gb.F64).import_lex(src))
    lin_d, lin_f = gb.SchurDiagMooeeOperator(Dd), gb.SchurDiagMooeeOperator(Df)
    sol_m, sol_d = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero(), gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
    mcg = gb.MixedPrecisionConjugateGradient(1e-8, 10000, 50, lin_f, lin_d)
    mcg.InnerTolerance = 3.0e-5                                     # as the reference's program sets it (:88)
    mcg(so, sol_m)
    cg = gb.ConjugateGradient(1e-8, 10000)
    cg(lin_d, so, sol_d)
    xm, xd = sol_m.export_lex(), sol_d.export_lex()
    assert mcg.TrueResidual < 1e-7
    assert np.linalg.norm((xm - xd).ravel()) / np.linalg.norm(xd.ravel()) < 1e-6


def _proj_upper(psi, mu, sign):
    """upper two spin components of (1 + sign gamma_mu) psi in the reference's chiral basis (ref: Grid/qcd/spin/TwoSpinor.h:75-133), computed
    in psi's own precision exactly as the pack kernel does (one add per component)"""
    f0, f1, f2, f3 = (psi[:, k, :] for k in range(4))
    i = psi.dtype.type(1j)
    if mu == 0:
        return (f0 + i * f3, f1 + i * f2) if sign > 0 else (f0 - i * f3, f1 - i * f2)
    if mu == 1:
        return (f0 - f3, f1 + f2) if sign > 0 else (f0 + f3, f1 - f2)
    if mu == 2:
        return (f0 + i * f2, f1 - i * f3) if sign > 0 else (f0 - i * f2, f1 + i * f3)
    return (f0 + f2, f1 + f3) if sign > 0 else (f0 - f2, f1 - f3)


def _round_bf16(c64):
    """round-to-nearest-even of both parts of complex64 values to bf16 (kept in complex64)"""
    u = np.ascontiguousarray(c64).view(np.uint32)
    r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)) << np.uint32(16)
    return r.view(np.complex64).reshape(c64.shape)


@pytest.mark.parametrize("mask", ["t", "zt", "xyzt"])
def test_self_halo_compressed_halos_match_a_model_of_the_compressor(ctx, mask):
    """Exact statement of what bf16 halos do: a leg that leaves the rank sees the neighbour's PROJECTED half spinor rounded to bf16 (ref:
    WilsonCompressorTemplate::Compress, WilsonCompressor.h:271-277 -- project, then store in the comms precision), nothing else changes.
    Model built from the oracle's single legs (DhopDir): since the projection only reads the rounded upper components when the lower ones
    are zero, the compressed leg equals the oracle's leg applied to psi' = (round(h0), round(h1), 0, 0).  The library must agree with
    Dhop(psi) + sum over off-rank legs [leg(psi') - leg(psi)] per site to the fp32 tolerance -- surface sites included."""
    sh = SHAPES["dwf16"]
    dims, Ls = sh["dims"], sh["Ls"]
    grid, D, orc = make(ctx, sh, gb.F32, MASKS[mask])
    src = syn.random_fermion(dims, Ls, seed=47, dtype=np.complex64)
    src64 = src.astype(np.complex128)
    model = orc.apply(po.OP_DHOP, src64, dag=0).astype(np.complex128)
    idx = np.arange(int(np.prod(dims)))
    coord = []
    for d in range(4):
        coord.append(np.repeat(idx % dims[d], Ls)); idx = idx // dims[d]
    for mu in range(4):
        if not (MASKS[mask] >> mu) & 1:
            continue
        for disp in (+1, -1):
            sign = -1 if disp > 0 else +1                       # non-dag: forward legs carry (1 - gamma), backward legs (1 + gamma)
            h0, h1 = _proj_upper(src, mu, sign)
            psi_c = np.zeros_like(src)
            psi_c[:, 0, :] = _round_bf16(h0); psi_c[:, 1, :] = _round_bf16(h1)
            off_rank = coord[mu] == (dims[mu] - 1 if disp > 0 else 0)     # output sites whose (mu, disp) leg leaves the rank
            delta = orc.dhop_dir(psi_c.astype(np.complex128), mu, disp) - orc.dhop_dir(src64, mu, disp)
            model[off_rank] += delta[off_rank]
    fin, fout = gb.LatticeFermion(grid, Ls, gb.F32).import_lex(src), gb.LatticeFermion(grid, Ls, gb.F32)
    D.set_halo_compression(True)
    for overlap in (1, 0):
        D.set_overlap(overlap)
        D.Dhop(fin, fout, 0)
        assert site_rel_err(fout.export_lex(), model) < TOL_HOP[gb.F32], (mask, overlap)
    D.set_halo_compression(False)
