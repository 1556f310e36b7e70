#!/usr/bin/env python
"""Random decompositions and local shapes through the N-rank improved staggered operator on the CPU mock (ranks = host threads;
tests/mock/README.md): three-deep halos of the field and of the links, overlapped (interior / exterior split) and serial forms, both
precisions.  Dhop +-dag, DhopEO / DhopOE and M must agree per site with the SAME library run on one rank over the global lattice
(that single-rank kernel is measured parity-green on the B200).  Not part of the test suite (open-ended).
usage: fuzz_stag_ranks.py <libgridb200_mock.so> <seed> <seconds>   (last recorded run: 3 seeds x 120 s = 2088 cases, 0 disagreements)"""
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
MPIS = [(1, 1, 1, 2), (1, 1, 2, 1), (1, 2, 1, 1), (2, 1, 1, 1), (1, 1, 2, 2), (2, 1, 1, 2), (1, 2, 2, 1), (2, 2, 1, 1), (1, 1, 1, 4), (4, 1, 1, 1), (1, 1, 4, 1)]
bad, lock, ncase = [], threading.Lock(), 0


def err(a, b):
    d = np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128), axis=1); nb = np.linalg.norm(b.astype(np.complex128), axis=1)
    return float(np.max(d / np.maximum(nb, np.sqrt(np.mean(nb ** 2)) + 1e-300)))


while time.time() < t_end:
    mpi = random.choice(MPIS)
    local = tuple(random.choice([4, 6, 8]) for _ in range(4))
    gdims = tuple(l * m for l, m in zip(local, mpi))
    V = int(np.prod(gdims))
    if V > 20000:
        continue
    world = int(np.prod(mpi))
    prec = random.choice([gb.F32, gb.F64])
    tol = 2e-6 if prec == gb.F32 else 1e-13
    U = syn.hot_gauge(gdims, seed=ncase + 1)
    rng = np.random.default_rng(900 + ncase)
    src = (rng.random((V, 3)) + 1j * rng.random((V, 3))).astype(gb._cdtype(prec))
    ctx1 = gb.Context(0)
    g1 = gb.GridCartesian(ctx1, gdims)
    U1 = gb.LatticeGaugeField(g1, prec).import_lex(U)
    D1 = gb.ImprovedStaggeredFermion(U1, U1, g1, 0.1)
    f1 = gb.LatticeStaggeredFermion(g1, 1, prec).import_lex(src)
    ref = {}
    for dag in (0, 1):
        o = gb.LatticeStaggeredFermion(g1, 1, prec); D1.Dhop(f1, o, dag); ref[dag] = o.export_lex()
    o = gb.LatticeStaggeredFermion(g1, 1, prec); D1.M(f1, o); ref["M"] = o.export_lex()
    for cb_in, name in ((gb.Odd, "eo"), (gb.Even, "oe")):
        hi, ho = gb.LatticeStaggeredFermion(g1, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(g1, 1, prec, gb.HALF)
        gb.pickCheckerboard(cb_in, hi, f1)
        (D1.DhopEO if cb_in == gb.Odd else D1.DhopOE)(hi, ho, 0)
        full = gb.LatticeStaggeredFermion(g1, 1, prec).zero(); gb.setCheckerboard(full, ho); ref[name] = full.export_lex()
    tag = f"mpi {mpi} local {local} prec {prec}"

    def body(rank):
        try:
            ctx = gb.Context(rank); ctx.comm_init(rank, world, b"\0" * 128)
            grid = gb.GridCartesian(ctx, gdims, mpi)
            Umu = gb.LatticeGaugeField(grid, prec).import_lex(decomp.scatter(U, gdims, mpi, rank))
            D = gb.ImprovedStaggeredFermion(Umu, Umu, grid, 0.1)
            fin = gb.LatticeStaggeredFermion(grid, 1, prec).import_lex(decomp.scatter(src, gdims, mpi, rank))
            out = gb.LatticeStaggeredFermion(grid, 1, prec)

            def chk(what, got, want):
                e = err(got, decomp.scatter(want, gdims, mpi, rank))
                if not e < tol:
                    with lock:
                        bad.append((tag, rank, what, e))
            for overlap in (True, False):
                D.set_overlap(overlap)
                for dag in (0, 1):
                    D.Dhop(fin, out, dag); chk(f"overlap {overlap} Dhop dag {dag}", out.export_lex(), ref[dag])
                D.M(fin, out); chk(f"overlap {overlap} M", out.export_lex(), ref["M"])
                for cb_in, name in ((gb.Odd, "eo"), (gb.Even, "oe")):
                    hi, ho = gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF), gb.LatticeStaggeredFermion(grid, 1, prec, gb.HALF)
                    gb.pickCheckerboard(cb_in, hi, fin)
                    (D.DhopEO if cb_in == gb.Odd else D.DhopOE)(hi, ho, 0)
                    full = gb.LatticeStaggeredFermion(grid, 1, prec).zero(); gb.setCheckerboard(full, ho)
                    chk(f"overlap {overlap} Dhop{name}", full.export_lex(), ref[name])
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
print("cases", ncase, "bad", len(bad), bad[:8], flush=True)
os._exit(1 if bad else 0)
