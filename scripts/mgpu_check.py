"""Multi-GPU parity check, launched one rank per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/mgpu_check.py
Every rank builds the same global fields from fixed seeds, imports its local block, runs the decomposed operators
(halo pack + NCCL exchange + interior/exterior kernels) and compares with the CPU oracle on the GLOBAL lattice."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import grid_b200 as gb
from grid_b200 import synthetic as syn, decomp
from oracle import pyoracle as po

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ctx = gb.Context(lrank)
uid = [gb.Context.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])

MPIS = {2: [(1, 1, 1, 2), (2, 1, 1, 1), (1, 2, 1, 1)], 4: [(1, 1, 2, 2), (2, 2, 1, 1)], 8: [(1, 1, 2, 4), (2, 2, 2, 1)]}[world]
fails = []


def site_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)))


def check(name, err, tol):
    t = torch.tensor([err], dtype=torch.float64, device=f"cuda:{lrank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t.item() < tol
    if rank == 0:
        print(f"{'ok  ' if ok else 'FAIL'} {name}: max site err {t.item():.3e} (tol {tol:.0e})", flush=True)
    if not ok:
        fails.append(name)


if os.environ.get("MGPU_MPI_INDEX"):          # restrict to one processor grid (saves GPU time on 8-GPU boxes)
    MPIS = [MPIS[int(os.environ["MGPU_MPI_INDEX"])]]
CASES = []
for mpi in MPIS:
    gd = tuple(max(4 * m, 8) if m > 1 else 4 for m in mpi)
    CASES.append((mpi, gd, ((16, "dwf"), (6, "mobius"), (1, "wilson"))))
    # local x,y = 8: the tuned kernels apply, so a z/t decomposition takes the split-volume path (interior box + surface slabs)
    gd2 = tuple((8 if d < 2 else 4) * m if m == 1 else max((8 if d < 2 else 4) * m, 8) for d, m in enumerate(mpi))
    CASES.append((mpi, gd2, ((16, "dwf"), (8, "mobius"))))
for mpi, gdims, kinds in CASES:
    for Ls, kind in kinds:
        U = syn.hot_gauge(gdims, seed=3)
        src = syn.random_fermion(gdims, Ls, seed=4)
        grid = gb.GridCartesian(ctx, gdims, mpi)
        for prec, tol in ((gb.F32, 1e-6), (gb.F64, 1e-13)):
            orc = po.OracleOp(0 if kind == "wilson" else 1, gdims, Ls, mass=0.1, M5=1.8, b=1.5 if kind == "mobius" else 1.0,
                              c=0.5 if kind == "mobius" else 0.0, prec=1)
            orc.import_gauge(U)
            Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gdims, mpi, rank))
            if kind == "wilson":
                D = gb.WilsonFermion(Umu, grid, 0.1)
            elif kind == "dwf":
                D = gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8)
            else:
                D = gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, 1.5, 0.5)
            fin = gb.LatticeFermion(grid, Ls, prec).import_lex(decomp.scatter(src, gdims, mpi, rank, inner=Ls).astype(gb._cdtype(prec)))
            out = gb.LatticeFermion(grid, Ls, prec)
            for overlap in (True, False):
                D.set_overlap(overlap)
                for dag in (0, 1):
                    D.Dhop(fin, out, dag)
                    ref = decomp.scatter(orc.apply(po.OP_DHOP, src, dag=dag), gdims, mpi, rank, inner=Ls)
                    check(f"mpi {mpi} {kind} Ls{Ls} prec{prec} overlap{int(overlap)} Dhop dag{dag}", site_err(out.export_lex(), ref), tol)
            D.set_overlap(True)
            # checkerboard hop + full operator + Schur operator
            he, ho = gb.LatticeFermion(grid, Ls, prec, gb.HALF), gb.LatticeFermion(grid, Ls, prec, gb.HALF)
            gb.pickCheckerboard(gb.Odd, ho, fin)
            D.DhopEO(ho, he, 0)
            full = gb.LatticeFermion(grid, Ls, prec).zero()
            gb.setCheckerboard(full, he)
            ref_e = np.zeros_like(src)
            po.set_checkerboard(gdims, Ls, 0, ref_e, orc.apply(po.OP_DHOP_EO, po.pick_checkerboard(gdims, Ls, 1, src)))
            refl = decomp.scatter(ref_e, gdims, mpi, rank, inner=Ls)
            got = full.export_lex()
            mask = np.linalg.norm(refl.reshape(refl.shape[0], -1), axis=1) > 0
            check(f"mpi {mpi} {kind} Ls{Ls} prec{prec} DhopEO", site_err(got[mask], refl[mask]), tol)
            D.M(fin, out)
            check(f"mpi {mpi} {kind} Ls{Ls} prec{prec} M", site_err(out.export_lex(), decomp.scatter(orc.apply(po.OP_M, src), gdims, mpi, rank, inner=Ls)), 4 * tol)
            # reductions are global
            n2 = gb.norm2(fin)
            n2ref = np.vdot(src, src).real
            check(f"mpi {mpi} {kind} Ls{Ls} prec{prec} norm2", abs(n2 - n2ref) / n2ref, 1e-6 if prec == gb.F32 else 1e-13)
        # CG on the decomposed lattice vs the oracle on the global lattice (fp64)
        if kind != "wilson" or True:
            src_o = po.pick_checkerboard(gdims, Ls, 1, src)
            x_ref, info = orc.cg(1, src_o, 1e-8, 5000)
            so = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF)
            gb.pickCheckerboard(gb.Odd, so, gb.LatticeFermion(grid, Ls, gb.F64).import_lex(decomp.scatter(src, gdims, mpi, rank, inner=Ls)))
            sol = gb.LatticeFermion(grid, Ls, gb.F64, gb.HALF).zero()
            cg = gb.ConjugateGradient(1e-8, 5000)
            cg(gb.SchurDiagMooeeOperator(D), so, sol)
            ok = abs(cg.IterationsToComplete - info["iterations"]) <= max(1, 0.02 * info["iterations"])
            if rank == 0:
                print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} {kind} Ls{Ls} CG iterations {cg.IterationsToComplete} vs oracle {info['iterations']}, true resid {cg.TrueResidual:.3e} vs {info['true_residual']:.3e}", flush=True)
            if not ok:
                fails.append("cg")
# ---- improved staggered on the decomposed lattice (three-deep Naik halos of field and links), SURVEY 8 row a26 / 8e
for mpi in (MPIS if os.environ.get("MGPU_SKIP_STAG") is None else []):
    gdims = tuple(max(4 * m, 8) if m > 1 else (6 if d == 1 else 4) for d, m in enumerate(mpi))   # local extents 4 (6 in y when whole)
    V = int(np.prod(gdims))
    U = syn.hot_gauge(gdims, seed=5)
    rng = np.random.default_rng(6)
    src = rng.random((V, 3)) + 1j * rng.random((V, 3))
    grid = gb.GridCartesian(ctx, gdims, mpi)
    orc = po.StagOracleOp(gdims, 0.1, prec=1)
    orc.import_gauge(U)
    for prec, tol in ((gb.F32, 1e-6), (gb.F64, 1e-13)):
        Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gdims, mpi, rank))
        D = gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
        fin = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(decomp.scatter(src, gdims, mpi, rank).astype(gb._cdtype(prec)))
        out = gb.LatticeStaggeredFermion(grid, 1, prec)

        def stag_err(got, ref):   # relative to the rms site norm: single sites of a hop can cancel to ~0
            d = np.linalg.norm(got.astype(np.complex128) - ref, axis=1); nb = np.linalg.norm(ref, axis=1)
            return float(np.max(d / np.maximum(nb, np.sqrt(np.mean(nb ** 2)))))
        for overlap in (True, False):
            D.set_overlap(overlap)
            for dag in (0, 1):
                D.Dhop(fin, out, dag)
                check(f"mpi {mpi} staggered prec{prec} overlap{int(overlap)} Dhop dag{dag}", stag_err(out.export_lex(), decomp.scatter(orc.apply(po.OP_DHOP, src, dag=dag), gdims, mpi, rank)), tol)
        D.set_overlap(True)
        D.M(fin, out)
        check(f"mpi {mpi} staggered prec{prec} M", stag_err(out.export_lex(), decomp.scatter(orc.apply(po.OP_M, src), gdims, mpi, rank)), tol)
        he, ho = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
        for cb_in, hin, hout, meth in ((gb.Odd, ho, he, D.DhopEO), (gb.Even, he, ho, D.DhopOE)):
            gb.pickCheckerboard(cb_in, hin, fin)
            meth(hin, hout, 0)
            full = gb.LatticeStaggeredFermion(grid, 1, prec).zero()
            gb.setCheckerboard(full, hout)
            ref_f = np.zeros_like(src)
            po.set_checkerboard_sites(gdims, 1 - cb_in, ref_f, orc.apply(po.OP_DHOP_EO if cb_in == gb.Odd else po.OP_DHOP_OE, po.pick_checkerboard_sites(gdims, cb_in, src)))
            check(f"mpi {mpi} staggered prec{prec} Dhop{'EO' if cb_in == gb.Odd else 'OE'}", stag_err(full.export_lex(), decomp.scatter(ref_f, gdims, mpi, rank)), tol)
    # CG on SchurStaggeredOperator and the full SchurRedBlackStaggeredSolve vs the oracle on the global lattice (fp64 operator from the loop)
    x_ref, info = orc.cg(1, po.pick_checkerboard_sites(gdims, 1, src), 1e-8, 5000)
    so, sol = gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, gb.F64, gb.HALF).zero()
    gb.pickCheckerboard(gb.Odd, so, fin)
    cg = gb.ConjugateGradient(1e-8, 5000)
    cg(gb.SchurStaggeredOperator(D), so, sol)
    ok = abs(cg.IterationsToComplete - info["iterations"]) <= max(1, 0.02 * info["iterations"])
    if rank == 0:
        print(f"{'ok  ' if ok else 'FAIL'} mpi {mpi} staggered CG iterations {cg.IterationsToComplete} vs oracle {info['iterations']}, true resid {cg.TrueResidual:.3e} vs {info['true_residual']:.3e}", flush=True)
    if not ok:
        fails.append("stag cg")
    xs_ref, _ = orc.schur_solve(src, 1e-8, 5000)
    xs = gb.LatticeStaggeredFermion(grid, 1, gb.F64)
    gb.SchurRedBlackStaggeredSolve(gb.ConjugateGradient(1e-8, 5000))(D, fin, xs)
    check(f"mpi {mpi} staggered SchurRedBlackStaggeredSolve", stag_err(xs.export_lex(), decomp.scatter(xs_ref, gdims, mpi, rank)), 1e-6)
dist.barrier()
if rank == 0:
    print("MGPU_CHECK " + ("PASS" if not fails else f"FAIL {fails}"), flush=True)
dist.destroy_process_group()
sys.exit(1 if fails else 0)
