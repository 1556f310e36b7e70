#!/usr/bin/env python
"""N-rank parity on the CPU mock (tests/mock/README.md): every "rank" is a host thread with its own context, halo messages travel
through the mock's mailboxes.  The same checks scripts/mgpu_check.py makes on N GPUs, at sizes the mock runs in seconds:
decomposed Wilson / DWF hops (serial and overlapped orchestration of dhop.cu with the generic kernel, gauge-face exchange of
the double store), the improved staggered operator with three-deep halos of field and links (stag.cu, overlapped and serial),
global reductions, CG and the Schur solve -- against the oracle on the GLOBAL lattice.
usage: mgpu_on_mock.py <libgridb200_mock.so> ; exit code 0 = every check passed"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grid_b200 as gb                      # noqa: E402
from grid_b200 import synthetic as syn, decomp   # noqa: E402
from oracle import pyoracle as po           # noqa: E402

gb.LIB_PATH = sys.argv[1]
fails, lock = [], threading.Lock()


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    nb = np.linalg.norm(b, axis=1)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-6 * np.sqrt(np.mean(nb ** 2)) + 1e-300)))


def check(rank, name, err, tol):
    if not err < tol:
        with lock:
            fails.append(f"rank {rank} {name}: {err:.3e} (tol {tol:.0e})")


def run_ranks(world, body):
    errs = []

    def wrap(rank):
        try:
            body(rank)
        except Exception as e:      # a rank that dies would leave the others waiting in a collective: report and let the timeout reap
            with lock:
                errs.append(f"rank {rank}: {type(e).__name__}: {e}")
    ts = [threading.Thread(target=wrap, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    if any(t.is_alive() for t in ts):
        errs.append("ranks still waiting after 600 s (deadlock)")
    return errs


def wilson_like(mpi, gdims, kind, Ls, phases=None):
    world = int(np.prod(mpi))
    U = syn.hot_gauge(gdims, seed=3)
    src = syn.random_fermion(gdims, Ls, seed=4)
    orc = po.OracleOp(0 if kind == "wilson" else 1, gdims, Ls, mass=0.1, M5=1.8, b=1.5 if kind == "mobius" else 1.0, c=0.5 if kind == "mobius" else 0.0, prec=1)
    orc.import_gauge(U, phases)
    ref = {("dhop", d): orc.apply(po.OP_DHOP, src, dag=d) for d in (0, 1)}
    ref["M"] = orc.apply(po.OP_M, src)
    src_o = po.pick_checkerboard(gdims, Ls, 1, src)
    ref_e = np.zeros_like(src)
    po.set_checkerboard(gdims, Ls, 0, ref_e, orc.apply(po.OP_DHOP_EO, src_o))
    x_ref, info = orc.cg(1, src_o, 1e-8, 5000)
    n2ref = np.vdot(src, src).real

    def body(rank):
        ctx = gb.Context(rank)
        ctx.comm_init(rank, world, b"\0" * 128)
        grid = gb.GridCartesian(ctx, gdims, mpi)
        tag = f"mpi {mpi} {kind} Ls{Ls}"
        for prec, tol in ((gb.F32, 1e-6), (gb.F64, 1e-13)):
            Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gdims, mpi, rank))
            D = gb.WilsonFermion(Umu, grid, 0.1, phases) if kind == "wilson" else gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8, phases) if kind == "dwf" else \
                gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, 1.5, 0.5, phases)
            fin = gb.LatticeFermion(grid, Ls, prec).import_lex(decomp.scatter(src, gdims, mpi, rank, inner=Ls).astype(gb._cdtype(prec)))
            out = gb.LatticeFermion(grid, Ls, prec)
            for overlap in (True, False):
                D.set_overlap(overlap)
                for dag in (0, 1):
                    D.Dhop(fin, out, dag)
                    check(rank, f"{tag} prec{prec} overlap{int(overlap)} Dhop dag{dag}", site_err(out.export_lex(), decomp.scatter(ref[("dhop", dag)], gdims, mpi, rank, inner=Ls)), tol)
            D.set_overlap(True)
            he, ho = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
            gb.pickCheckerboard(gb.Odd, ho, fin)
            D.DhopEO(ho, he, 0)
            full = gb.LatticeFermion(grid, Ls, prec).zero()
            gb.setCheckerboard(full, he)
            refl = decomp.scatter(ref_e, gdims, mpi, rank, inner=Ls)
            mask = np.linalg.norm(refl.reshape(refl.shape[0], -1), axis=1) > 0
            check(rank, f"{tag} prec{prec} DhopEO", site_err(full.export_lex()[mask], refl[mask]), tol)
            D.M(fin, out)
            check(rank, f"{tag} prec{prec} M", site_err(out.export_lex(), decomp.scatter(ref["M"], gdims, mpi, rank, inner=Ls)), 4 * tol)
            check(rank, f"{tag} prec{prec} norm2", abs(gb.norm2(fin) - n2ref) / n2ref, 1e-6 if prec == gb.F32 else 1e-13)
            # a single leg across the rank boundary (force terms): the eight legs still sum to Dhop
            tot = gb.LatticeFermion(grid, Ls, prec).zero()
            for d in range(4):
                for s in (1, -1):
                    D.DhopDir(fin, out, d, s)
                    gb.axpy(tot, 1.0, out, tot)
            check(rank, f"{tag} prec{prec} sum of DhopDir legs", site_err(tot.export_lex(), decomp.scatter(ref[("dhop", 0)], gdims, mpi, rank, inner=Ls)), 4 * tol)
            # the face exchange on its own (scripts/halo_bench.py): bytes sent = every split face, both directions, both parities,
            # half spinors; and the epoch bookkeeping must leave the next hop intact
            ld = [g // m for g, m in zip(gdims, mpi)]
            want = sum(4 * (-(-(int(np.prod(ld)) // 2 // ld[mu] * Ls) // 16) * 16) * 6 * (8 if prec == gb.F32 else 16) for mu in range(4) if mpi[mu] > 1)
            check(rank, f"{tag} prec{prec} halo_exchange bytes {D.halo_exchange(fin)} vs {want}", abs(D.halo_exchange(fin) - want), 0.5)
            # host-resident fields through gb_op_dhop_host: on z / t splits the faces go in first, one exchange, then the slices stream
            hloc = decomp.scatter(src, gdims, mpi, rank, inner=Ls).astype(gb._cdtype(prec))
            for dag in (0, 1):
                got = D.Dhop_host(hloc, np.empty_like(hloc), dag)
                check(rank, f"{tag} prec{prec} Dhop_host dag{dag}", site_err(got, decomp.scatter(ref[("dhop", dag)], gdims, mpi, rank, inner=Ls)), tol)
            # compressed halos (fp32 operator: bf16 on the wire, fp64 operator: fp32): within the comms precision of the oracle, and back
            D.set_halo_compression(True)
            for overlap in (True, False):
                D.set_overlap(overlap)
                D.Dhop(fin, out, 1)
                check(rank, f"{tag} prec{prec} overlap{int(overlap)} Dhop dag1 compressed halos", site_err(out.export_lex(), decomp.scatter(ref[("dhop", 1)], gdims, mpi, rank, inner=Ls)), 8e-3 if prec == gb.F32 else 2e-6)
            got = D.Dhop_host(hloc, np.empty_like(hloc), 0)
            check(rank, f"{tag} prec{prec} Dhop_host compressed halos", site_err(got, decomp.scatter(ref[("dhop", 0)], gdims, mpi, rank, inner=Ls)), 8e-3 if prec == gb.F32 else 2e-6)
            D.set_halo_compression(False); D.set_overlap(True)
            # two-row links (third row rebuilt in registers; the folded-in factor and boundary phase go on by GLOBAL coordinate)
            D.set_link_reconstruct(12)
            for overlap in (True, False):
                D.set_overlap(overlap)
                D.Dhop(fin, out, 1)
                check(rank, f"{tag} prec{prec} overlap{int(overlap)} Dhop dag1 two-row links", site_err(out.export_lex(), decomp.scatter(ref[("dhop", 1)], gdims, mpi, rank, inner=Ls)), tol)
            D.set_link_reconstruct(18); D.set_overlap(True)
            D.Dhop(fin, out, 0)
            check(rank, f"{tag} prec{prec} Dhop after halo_exchange", site_err(out.export_lex(), decomp.scatter(ref[("dhop", 0)], gdims, mpi, rank, inner=Ls)), tol)
        so, sol = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
        gb.pickCheckerboard(gb.Odd, so, fin)
        cg = gb.ConjugateGradient(1e-8, 5000)
        cg(gb.SchurDiagMooeeOperator(D), so, sol)
        check(rank, f"{tag} CG iterations {cg.IterationsToComplete} vs {info['iterations']}", abs(cg.IterationsToComplete - info["iterations"]), max(1, 0.02 * info["iterations"]) + 0.5)
    return run_ranks(world, body)


def staggered(mpi, gdims):
    world = int(np.prod(mpi))
    V = int(np.prod(gdims))
    U = syn.hot_gauge(gdims, seed=5)
    rng = np.random.default_rng(6)
    src = rng.random((V, 3)) + 1j * rng.random((V, 3))
    orc = po.StagOracleOp(gdims, 0.1, prec=1)
    orc.import_gauge(U)
    ref = {("dhop", d): orc.apply(po.OP_DHOP, src, dag=d) for d in (0, 1)}
    ref["M"] = orc.apply(po.OP_M, src)
    ref_cb = {}
    for cb_in in (0, 1):
        f = np.zeros_like(src)
        po.set_checkerboard_sites(gdims, 1 - cb_in, f, orc.apply(po.OP_DHOP_EO if cb_in == 1 else po.OP_DHOP_OE, po.pick_checkerboard_sites(gdims, cb_in, src)))
        ref_cb[cb_in] = f
    x_ref, info = orc.cg(1, po.pick_checkerboard_sites(gdims, 1, src), 1e-8, 5000)
    xs_ref, _ = orc.schur_solve(src, 1e-8, 5000)

    def stag_err(got, want):
        d = np.linalg.norm(got.astype(np.complex128) - want, axis=1); nb = np.linalg.norm(want, axis=1)
        return float(np.max(d / np.maximum(nb, np.sqrt(np.mean(nb ** 2)))))

    def body(rank):
        ctx = gb.Context(rank)
        ctx.comm_init(rank, world, b"\0" * 128)
        grid = gb.GridCartesian(ctx, gdims, mpi)
        tag = f"mpi {mpi} staggered"
        for prec, tol in ((gb.F32, 1e-6), (gb.F64, 1e-13)):
            Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gdims, mpi, rank))
            D = gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
            fin = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(decomp.scatter(src, gdims, mpi, rank).astype(gb._cdtype(prec)))
            out = gb.LatticeStaggeredFermion(grid, 1, prec)
            for overlap in (True, False):
                D.set_overlap(overlap)
                for dag in (0, 1):
                    D.Dhop(fin, out, dag)
                    check(rank, f"{tag} prec{prec} overlap{int(overlap)} Dhop dag{dag}", stag_err(out.export_lex(), decomp.scatter(ref[("dhop", dag)], gdims, mpi, rank)), tol)
            D.set_overlap(True)
            D.M(fin, out)
            check(rank, f"{tag} prec{prec} M", stag_err(out.export_lex(), decomp.scatter(ref["M"], gdims, mpi, rank)), tol)
            he, ho = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
            for cb_in, hin, hout, meth in ((gb.Odd, ho, he, D.DhopEO), (gb.Even, he, ho, D.DhopOE)):
                gb.pickCheckerboard(cb_in, hin, fin)
                meth(hin, hout, 0)
                full = gb.LatticeStaggeredFermion(grid, 1, prec).zero()
                gb.setCheckerboard(full, hout)
                check(rank, f"{tag} prec{prec} Dhop{'EO' if cb_in == gb.Odd else 'OE'}", stag_err(full.export_lex(), decomp.scatter(ref_cb[cb_in], gdims, mpi, rank)), tol)
        so, sol = gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF).zero()
        gb.pickCheckerboard(gb.Odd, so, fin)
        cg = gb.ConjugateGradient(1e-8, 5000)
        cg(gb.SchurStaggeredOperator(D), so, sol)
        check(rank, f"{tag} CG iterations {cg.IterationsToComplete} vs {info['iterations']}", abs(cg.IterationsToComplete - info["iterations"]), max(1, 0.02 * info["iterations"]) + 0.5)
        xs = gb.LatticeStaggeredFermion(grid, 1, gb.F64)
        gb.SchurRedBlackStaggeredSolve(gb.ConjugateGradient(1e-8, 5000))(D, fin, xs)
        check(rank, f"{tag} SchurRedBlackStaggeredSolve", stag_err(xs.export_lex(), decomp.scatter(xs_ref, gdims, mpi, rank)), 1e-6)
    return run_ranks(world, body)


def main():
    errs = []
    for mpi in ((1, 1, 1, 2), (2, 1, 1, 1), (1, 2, 1, 1), (1, 1, 2, 2)):
        gd = tuple(4 * m if m > 1 else 4 for m in mpi)
        for kind, Ls in (("wilson", 1), ("dwf", 4), ("mobius", 6)):
            errs += wilson_like(mpi, gd, kind, Ls)
        errs += staggered(mpi, tuple(max(4 * m, 8) if m > 1 else (6 if d == 1 else 4) for d, m in enumerate(mpi)))
        print(f"mpi {mpi}: done, {len(fails) + len(errs)} problems so far", flush=True)
    errs += staggered((1, 1, 1, 4), (4, 4, 4, 16))       # distinct forward / backward neighbours
    # boundary phases on a decomposed lattice: they sit on the GLOBAL boundary links (full store: gauge-face exchange; two-row store: by coordinate)
    errs += wilson_like((1, 1, 2, 2), (4, 4, 8, 8), "dwf", 4, phases=[1, np.exp(0.4j), -1, np.exp(-0.9j)])
    errs += wilson_like((2, 1, 1, 1), (8, 4, 4, 4), "wilson", 1, phases=[-1, 1, 1, 1])
    # the tuned fp32 kernels on decomposed lattices (Ls = 8, local 8.4.4.4): pack_send into the neighbours' buffers, then the
    # semi-fused hop (z / t split: one launch, surface CTAs acquire the flags) or interior + exterior (x / y split)
    import ctypes
    lib = ctypes.CDLL(sys.argv[1])
    lib.gb_mock_coop_launches.restype = ctypes.c_long
    lib.gb_mock_coop_launches.argtypes = [ctypes.c_char_p]
    before = {k: lib.gb_mock_coop_launches(k) for k in (b"dhop_col2_kernel_fn", b"dhop_fast_kernel<LS, 0, 1>", b"pack_send_kernel", b"smat_kernel")}
    for mpi in ((1, 1, 2, 2), (1, 1, 1, 4), (1, 2, 1, 1), (2, 1, 1, 1)):
        errs += wilson_like(mpi, tuple(l * m for l, m in zip((8, 4, 4, 4), mpi)), "dwf", 8)
        print(f"mpi {mpi} tuned kernels: done, {len(fails) + len(errs)} problems so far", flush=True)
    for k, v in before.items():
        n = lib.gb_mock_coop_launches(k) - v
        print(f"cooperative launches of {k.decode()}: {n}", flush=True)
        if n == 0:
            errs.append(f"{k.decode()} never ran: the tuned multi-rank path was not exercised")
    for e in errs + fails:
        print("FAIL", e, flush=True)
    print("MGPU_ON_MOCK " + ("PASS" if not (errs or fails) else f"FAIL ({len(errs) + len(fails)})"), flush=True)
    os._exit(1 if (errs or fails) else 0)     # rank threads that are stuck must not keep the process alive


if __name__ == "__main__":
    main()
