"""SURVEY 8(f) row 3 -- ConjugateGradientMultiShift (ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:84-343) on the
even-odd Schur operator, the solver RHMC drives with the same hopping kernels.

 * CPU: the oracle restatement (oracle/solvers_oracle.hpp) reproduces the compiled reference's per-shift iteration counts,
   residuals and solutions stored in tests/golden/next_golden.npz, agrees with the compiled reference on a second lattice where
   oracle/_ref exists, and every solution satisfies (HermOp + pole) x = src.
 * GPU: the CUDA path (fused multi-field update kernels, gb_cg_multishift_schur) reproduces the same fixtures, is bit-identical
   to its own unfused composition (GB_MS_UNFUSED=1, subprocess), and the C++ driver runs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from grid_b200 import synthetic as syn
from oracle import pyoracle as po
from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = np.load(os.path.join(HERE, "golden", "dirac_golden.npz"))
N = np.load(os.path.join(HERE, "golden", "next_golden.npz"))
DIMS, LS = tuple(int(x) for x in G["dims"]), int(G["Ls"])
POLES, TOLS = [float(x) for x in N["ms_poles"]], [float(x) for x in N["ms_tols"]]


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    nb = np.linalg.norm(b, axis=1)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-3 * np.sqrt(np.mean(nb ** 2)) + 1e-300)))


def oracle_case(name):
    if name == "stag":
        o = po.StagOracleOp(DIMS, 0.1, prec=1); o.import_gauge(G["U"])
        return o, po.pick_checkerboard_sites(DIMS, 1, G["src_stag"])
    o = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=1); o.import_gauge(G["U"])
    return o, po.pick_checkerboard(DIMS, LS, 1, G["src5"])


def check_against_fixture(name, xs, iters, true_resid, to_complete):
    ref_it = [int(x) for x in N[f"{name}/multishift/iterations"]]
    for s, (a, b) in enumerate(zip(iters, ref_it)):
        assert abs(a - b) <= max(1, 0.02 * b), (name, s, iters, ref_it)
    assert abs(to_complete - int(N[f"{name}/multishift/iterations_to_complete"])) <= max(1, 0.02 * ref_it[0])
    for s, (a, b) in enumerate(zip(true_resid, N[f"{name}/multishift/true_residual"])):
        assert 0.6 < a / float(b) < 1.6, (name, s, a, b)
    for s in range(len(POLES)):
        assert site_err(xs[s], N[f"{name}/multishift/solutions"][s]) < 1e-6, (name, s)


@pytest.mark.parametrize("name", ["mobius", "stag"])
def test_oracle_multishift_matches_reference_outputs(name):
    o, src = oracle_case(name)
    xs, info = o.multishift_cg(1, src, POLES, TOLS, 5000)
    assert info["converged"] == 1
    check_against_fixture(name, xs, info["iterations"], info["true_residual"], info["iterations_to_complete"])
    # defining property: (HermOp + pole_s) x_s = src to the requested tolerance
    for s, pole in enumerate(POLES):
        r = o.apply(po.OP_HERMOP, xs[s], cb_in=1) + pole * xs[s] - src
        assert np.linalg.norm(r) / np.linalg.norm(src) < 3 * TOLS[s]
    # heavier poles need fewer iterations; the lightest one sets IterationsToComplete
    assert info["iterations"] == sorted(info["iterations"], reverse=True) and info["iterations_to_complete"] == info["iterations"][0]


def test_oracle_multishift_single_shift_is_plain_cg_on_shifted_operator():
    """With one shift at pole 0 the recurrences reduce to ConjugateGradient on HermOp (same Krylov space, same stopping norm)."""
    o, src = oracle_case("mobius")
    xs, info = o.multishift_cg(1, src, [0.0], [1e-8], 5000)
    x, cg = o.cg(1, src, 1e-8, 5000)
    assert abs(info["iterations"][0] - cg["iterations"]) <= 1
    assert site_err(xs[0], x) < 1e-7


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("prec", [1, 0])
def test_oracle_vs_reference_multishift(prec):
    dims, Ls = (4, 6, 8, 4), 6
    U = syn.hot_gauge(dims, seed=21)
    o = po.OracleOp(1, dims, Ls, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=prec); o.import_gauge(U)
    r = pr.RefOp(1, dims, Ls, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=prec); r.import_gauge(U)
    src = po.pick_checkerboard(dims, Ls, 1, syn.random_fermion(dims, Ls, seed=3)).astype(po._cdtype(prec))
    poles, tols = [0.02, 0.2, 2.0], ([1e-8, 1e-8, 1e-6] if prec else [1e-5, 1e-5, 1e-4])
    a, ia = o.multishift_cg(1, src, poles, tols, 3000)
    b, ib = r.multishift_cg(1, src, poles, tols, 3000)
    for x, y in zip(ia["iterations"], ib["iterations"]):
        assert abs(x - y) <= max(1, 0.02 * y), (ia, ib)
    for s in range(3):
        assert site_err(a[s], b[s]) < (1e-7 if prec else 1e-3)


def check_mixed_against_fixture(xs, iters, true_resid, to_complete):
    ref_it = [int(x) for x in N["mobius/multishift_mixed/iterations"]]
    for s, (a, b) in enumerate(zip(iters, ref_it)):
        assert abs(a - b) <= max(1, 0.02 * b), (s, iters, ref_it)
    assert abs(to_complete - int(N["mobius/multishift_mixed/iterations_to_complete"])) <= max(1, 0.02 * ref_it[0])
    for s in range(len(POLES)):
        assert true_resid[s] < 1.3 * TOLS[s], (s, true_resid)                 # after the clean-up solves every shift meets its tolerance
        assert site_err(xs[s], N["mobius/multishift_mixed/solutions"][s]) < 1e-6, s


def test_oracle_multishift_mixed_prec_matches_reference_outputs():
    """ConjugateGradientMultiShiftMixedPrec (ref: ConjugateGradientMultiShiftMixedPrec.h:128-410): fp32 operator, fp64 vectors,
    true-residual replacement every 20 iterations, mixed-precision clean-up of the shifts that miss their tolerance."""
    od, src = oracle_case("mobius")
    of = po.OracleOp(1, DIMS, LS, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=0); of.import_gauge(G["U"])
    xs, info = po.multishift_mixed_cg(od, of, 1, src, POLES, TOLS, 5000, 20)
    check_mixed_against_fixture(xs, info["iterations"], info["true_residual"], info["iterations_to_complete"])
    x64, _ = od.multishift_cg(1, src, POLES, TOLS, 5000)
    for s in range(len(POLES)):
        assert site_err(xs[s], x64[s]) < 1e-6                                  # same answers as the all-fp64 multishift


@pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libgridref.so not built (needs /root/reference)")
@pytest.mark.parametrize("freq", [10, 50])
def test_oracle_vs_reference_multishift_mixed_prec(freq):
    dims, Ls = (4, 6, 8, 4), 6
    U = syn.hot_gauge(dims, seed=21)
    mk = lambda mod, prec: mod(1, dims, Ls, mass=0.1, M5=1.8, b=1.5, c=0.5, prec=prec)
    od, of, rd, rf = mk(po.OracleOp, 1), mk(po.OracleOp, 0), mk(pr.RefOp, 1), mk(pr.RefOp, 0)
    for o in (od, of, rd, rf):
        o.import_gauge(U)
    src = po.pick_checkerboard(dims, Ls, 1, syn.random_fermion(dims, Ls, seed=3))
    poles, tols = [0.02, 0.2, 2.0], [1e-9, 1e-9, 1e-8]
    a, ia = po.multishift_mixed_cg(od, of, 1, src, poles, tols, 3000, freq)
    b, ib = pr.multishift_mixed_cg(rd, rf, 1, src, poles, tols, 3000, freq)
    for x, y in zip(ia["iterations"], ib["iterations"]):
        assert abs(x - y) <= max(1, 0.02 * y), (ia, ib)
    for s in range(3):
        assert site_err(a[s], b[s]) < 1e-6


# ---------------------------------------------------------------------------------------------- GPU
def _device_case(gb, name, prec):
    ctx = gb.Context(0)
    grid = gb.GridCartesian(ctx, DIMS)
    Umu = gb.LatticeGaugeField(grid, prec).import_lex(G["U"])
    if name == "stag":
        D = gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
        full = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(G["src_stag"].astype(gb._cdtype(prec)))
        mk = lambda: gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
        lin = gb.SchurStaggeredOperator(D)
    else:
        D = gb.MobiusFermion(Umu, grid, LS, 0.1, 1.8, 1.5, 0.5)
        full = gb.LatticeFermion(grid, LS, prec).import_lex(G["src5"].astype(gb._cdtype(prec)))
        mk = lambda: gb.LatticeFermion(grid, LS, prec, gb.HALF)
        lin = gb.SchurDiagMooeeOperator(D)
    src = mk()
    gb.pickCheckerboard(gb.Odd, src, full)
    return ctx, D, lin, src, mk


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mobius", "stag"])
def test_cuda_multishift_matches_reference(name):
    import grid_b200 as gb
    ctx, D, lin, src, mk = _device_case(gb, name, gb.F64)
    results = [mk() for _ in POLES]
    shifts = gb.MultiShiftFunction(POLES, TOLS, residues=[0.5, 0.25, 0.125, 1.0], norm=2.0)
    MSCG = gb.ConjugateGradientMultiShift(5000, shifts)
    psi = mk()
    assert MSCG(lin, src, results, psi)
    xs = [r.export_lex() for r in results]
    check_against_fixture(name, xs, MSCG.IterationsToCompleteShift, MSCG.TrueResidualShift, MSCG.IterationsToComplete)
    assert all(r.Checkerboard() == gb.Odd for r in results)
    # psi = norm * src + sum_s residues[s] results[s]   (ref: ConjugateGradientMultiShift.h:69-82)
    want = 2.0 * src.export_lex() + sum(c * x for c, x in zip([0.5, 0.25, 0.125, 1.0], xs))
    assert site_err(psi.export_lex(), want) < 1e-12
    # each solution against the oracle's operator on the host
    o, src_h = oracle_case(name)
    for s, pole in enumerate(POLES):
        r = o.apply(po.OP_HERMOP, xs[s], cb_in=1) + pole * xs[s] - src_h
        assert np.linalg.norm(r) / np.linalg.norm(src_h) < 3 * TOLS[s]


_UNFUSED_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
sys.path.insert(0, {tests!r})
import os
import grid_b200 as gb
if os.environ.get("GB_TEST_MOCK_LIB"):      # tests/test_next_on_cpu_mock.py
    gb.LIB_PATH = os.environ["GB_TEST_MOCK_LIB"]
import test_next_multishift as t
ctx, D, lin, src, mk = t._device_case(gb, "mobius", gb.F32)
results = [mk() for _ in t.POLES]
MSCG = gb.ConjugateGradientMultiShift(5000, gb.MultiShiftFunction(t.POLES, [1e-5] * 4))
MSCG(lin, src, results)
np.savez({out!r}, it=np.array(MSCG.IterationsToCompleteShift), **{{f"x{{i}}": r.export_lex() for i, r in enumerate(results)}})
"""


@pytest.mark.gpu
def test_cuda_multishift_fused_updates_are_bit_identical_to_unfused(tmp_path):
    """The fused multi-field kernels perform the same FMAs as the single-field BLAS calls: identical iterates, bit for bit."""
    outs = []
    for unfused in (False, True):
        out = str(tmp_path / f"ms_{int(unfused)}.npz")
        env = dict(os.environ)
        env.pop("GB_MS_UNFUSED", None)
        if unfused:
            env["GB_MS_UNFUSED"] = "1"
        p = subprocess.run([sys.executable, "-c", _UNFUSED_SNIPPET.format(root=ROOT, tests=HERE, out=out)], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        outs.append(np.load(out))
    assert np.array_equal(outs[0]["it"], outs[1]["it"])
    for i in range(len(POLES)):
        assert np.array_equal(outs[0][f"x{i}"], outs[1][f"x{i}"]), i


@pytest.mark.gpu
def test_dwf_multishift_driver():
    exe = os.path.join(ROOT, "drivers", "Test_dwf_multishift")
    assert os.path.exists(exe), f"{exe} missing: run make -C grid_b200"
    p = subprocess.run([exe, "--grid", "8.8.8.8", "--Ls", "8"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "PASS" in p.stdout


@pytest.mark.gpu
def test_cuda_multishift_mixed_prec_matches_reference():
    import grid_b200 as gb
    ctx, Dd, lin_d, src, mk = _device_case(gb, "mobius", gb.F64)
    Df = gb.MobiusFermion(gb.LatticeGaugeField(src.grid, gb.F32).import_lex(G["U"]), src.grid, LS, 0.1, 1.8, 1.5, 0.5)
    results = [mk() for _ in POLES]
    mcg = gb.ConjugateGradientMultiShiftMixedPrec(5000, gb.MultiShiftFunction(POLES, TOLS), gb.SchurDiagMooeeOperator(Df), 20)
    mcg(lin_d, src, results)
    check_mixed_against_fixture([r.export_lex() for r in results], mcg.IterationsToCompleteShift, mcg.TrueResidualShift, mcg.IterationsToComplete)
