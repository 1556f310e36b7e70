#!/usr/bin/env python
"""Random decompositions and local shapes through the multi-rank hop on the CPU mock (ranks = host threads, peer-to-peer halos;
tests/mock/README.md).  The N-rank result (overlapped = semi-fused where it applies, overlapped as interior + exterior, serial) of
Dhop +-dag and DhopEO must agree per site with the SAME library run on one rank over the global lattice.  fp32, Ls 8 / 12 / 16 (tuned
kernels) and Ls 4 (generic kernel).  Not part of the test suite (open-ended); run by hand when the halo code or a tuned kernel changes.
The optional forms ride along: gb_op_dhop_host (pipelined on z / t splits), two-row links, compressed halos.
usage: fuzz_ranks.py <libgridb200_mock.so> <seed> <seconds>   (recorded runs: 3 seeds x 150 s = 905 cases, 0 disagreements; with the optional forms: see tests/mock/README.md)"""
import os
import random
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grid_b200 as gb                      # noqa: E402
from grid_b200 import synthetic as syn, decomp   # noqa: E402

gb.LIB_PATH = sys.argv[1]
random.seed(int(sys.argv[2]))
t_end = time.time() + float(sys.argv[3])
MPIS = [(1, 1, 1, 2), (1, 1, 2, 1), (1, 2, 1, 1), (2, 1, 1, 1), (1, 1, 2, 2), (2, 1, 1, 2), (1, 2, 2, 1), (1, 1, 1, 4), (1, 1, 4, 1)]
bad, lock, ncase = [], threading.Lock(), 0


def err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.complex128); b = b.reshape(b.shape[0], -1).astype(np.complex128)
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)))


def make_op(grid, U, Ls, kind):
    Umu = gb.LatticeGaugeField(grid, gb.F32).import_lex(U)
    return gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8) if kind == "dwf" else gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, 1.5, 0.5)


while time.time() < t_end:
    mpi = random.choice(MPIS)
    local = (random.choice([2, 4, 8, 16]), random.choice([2, 4, 8]), random.choice([2, 4, 6, 8]), random.choice([2, 4, 6]))
    Ls = random.choice([4, 8, 12, 16])
    gdims = tuple(l * m for l, m in zip(local, mpi))
    if np.prod(gdims) * Ls > 60000:
        continue
    kind = random.choice(["dwf", "mobius"])
    world = int(np.prod(mpi))
    U = syn.hot_gauge(gdims, seed=ncase + 1)
    src = syn.random_fermion(gdims, Ls, seed=500 + ncase, dtype=np.complex64)
    # one rank, global lattice: the reference of this comparison (itself checked against the oracle by the test suite)
    ctx1 = gb.Context(0)
    g1 = gb.GridCartesian(ctx1, gdims)
    D1 = make_op(g1, U, Ls, kind)
    f1 = gb.LatticeFermion(g1, Ls, gb.F32).import_lex(src)
    ref = {}
    for dag in (0, 1):
        o = gb.LatticeFermion(g1, Ls, gb.F32); D1.Dhop(f1, o, dag); ref[dag] = o.export_lex()
    ho = gb.LatticeFermion(g1, Ls, gb.F32, gb.HALF); he = gb.LatticeFermion(g1, Ls, gb.F32, gb.HALF)
    gb.pickCheckerboard(gb.Odd, ho, f1); D1.DhopEO(ho, he, 0)
    full = gb.LatticeFermion(g1, Ls, gb.F32).zero(); gb.setCheckerboard(full, he); ref["eo"] = full.export_lex()
    tag = f"mpi {mpi} local {local} Ls {Ls} {kind}"

    def body(rank):
        try:
            ctx = gb.Context(rank); ctx.comm_init(rank, world, b"\0" * 128)
            grid = gb.GridCartesian(ctx, gdims, mpi)
            D = make_op(grid, decomp.scatter(U, gdims, mpi, rank), Ls, kind)
            fin = gb.LatticeFermion(grid, Ls, gb.F32).import_lex(decomp.scatter(src, gdims, mpi, rank, inner=Ls))
            out = gb.LatticeFermion(grid, Ls, gb.F32)
            for overlap in (1, 2, 0):
                D.set_overlap(overlap)
                for dag in (0, 1):
                    D.Dhop(fin, out, dag)
                    e = err(out.export_lex(), decomp.scatter(ref[dag], gdims, mpi, rank, inner=Ls))
                    if not e < 4e-6:
                        with lock:
                            bad.append((tag, rank, f"overlap {overlap} dag {dag}", e))
                h_o, h_e = gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF), gb.LatticeFermion(grid, Ls, gb.F32, gb.HALF)
                gb.pickCheckerboard(gb.Odd, h_o, fin); D.DhopEO(h_o, h_e, 0)
                fl = gb.LatticeFermion(grid, Ls, gb.F32).zero(); gb.setCheckerboard(fl, h_e)
                want = decomp.scatter(ref["eo"], gdims, mpi, rank, inner=Ls)
                m = np.linalg.norm(want.reshape(want.shape[0], -1), axis=1) > 0
                e = err(fl.export_lex()[m], want[m])
                if not e < 4e-6:
                    with lock:
                        bad.append((tag, rank, f"overlap {overlap} DhopEO", e))
            # the optional forms on the same decomposition: host-buffer entry point (pipelined on z / t splits), two-row links
            # (same bar as the full store), compressed halos (bf16 on the wire: within 8e-3, and not better than fp32 roundoff
            # would be only if no leg left the rank)
            D.set_overlap(1)
            hloc = decomp.scatter(src, gdims, mpi, rank, inner=Ls)
            want = decomp.scatter(ref[0], gdims, mpi, rank, inner=Ls)
            for form, setup, tol in (("Dhop_host", lambda: None, 4e-6), ("two-row links", lambda: D.set_link_reconstruct(12), 4e-6),
                                     ("two-row links Dhop_host", lambda: None, 4e-6), ("compressed halos", lambda: (D.set_link_reconstruct(18), D.set_halo_compression(True)), 8e-3),
                                     ("compressed halos Dhop_host", lambda: None, 8e-3)):
                setup()
                if "Dhop_host" in form:
                    got = D.Dhop_host(hloc, np.empty_like(hloc), 0)
                else:
                    D.Dhop(fin, out, 0); got = out.export_lex()
                e = err(got, want)
                if not e < tol:
                    with lock:
                        bad.append((tag, rank, form, e))
            D.set_halo_compression(False)
        except Exception as ex:     # noqa: BLE001
            with lock:
                bad.append((tag, rank, f"{type(ex).__name__}: {ex}", 0.0))
    ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    if any(t.is_alive() for t in ts):
        bad.append((tag, -1, "ranks still waiting after 300 s (deadlock)", 0.0))
        break
    ncase += 1
import ctypes   # noqa: E402
lib = ctypes.CDLL(sys.argv[1])
lib.gb_mock_coop_launches.restype = ctypes.c_long
lib.gb_mock_coop_launches.argtypes = [ctypes.c_char_p]
print("cases", ncase, "bad", len(bad), bad[:8], "| launches:", {k.decode(): lib.gb_mock_coop_launches(k) for k in
      (b"dhop_fast_kernel<LS, 0, 2>", b"dhop_fast_kernel<LS, 0, 1>", b"dhop_col_kernel", b"pack_send_kernel", b"dhop_kernel")}, flush=True)
os._exit(1 if bad else 0)
