#!/usr/bin/env python
"""Random lattice shapes through the tuned fp32 kernels on the CPU mock (tests/mock/README.md): local extents 2..18, Ls 8 / 12 / 16 / 24,
column heights 1..7, Shamir and Moebius; Dhop +-dag, M and Mdag through the default kernel selection and the micro-block kernel must
agree with the generic kernel per site to 8e-6.  Not part of the test suite (open-ended); run by hand when a tuned kernel changes.
usage: fuzz_shapes.py <libgridb200_mock.so> <seed> <seconds>   (last recorded run: 4 seeds x 150 s = 905 shapes, 0 disagreements)"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import grid_b200 as gb
from grid_b200 import synthetic as syn
gb.LIB_PATH = sys.argv[1]
random.seed(int(sys.argv[2]))
ctx = gb.Context(0)
t_end = time.time() + float(sys.argv[3])
nrun = 0; bad = []
while time.time() < t_end:
    dims = (random.choice([2, 4, 6, 8, 12, 16]), random.choice([2, 4, 6, 8, 12]), random.choice([2, 4, 6, 8, 10, 18]), random.choice([2, 4, 6]))
    Ls = random.choice([8, 12, 16, 24])
    if np.prod(dims) * Ls > 40000: continue
    kind = random.choice(["dwf", "mobius"])
    grid = gb.GridCartesian(ctx, dims)
    U = syn.hot_gauge(dims, seed=nrun + 1)
    Umu = gb.LatticeGaugeField(grid, gb.F32).import_lex(U)
    D = gb.DomainWallFermion(Umu, grid, Ls, 0.1, 1.8) if kind == "dwf" else gb.MobiusFermion(Umu, grid, Ls, 0.1, 1.8, 1.5, 0.5)
    coln = random.choice([0, 1, 2, 3, 4, 5, 7])
    if coln: D.set_tiling(0, coln, 0)
    h = syn.random_fermion(dims, Ls, seed=100 + nrun, dtype=np.complex64)
    fin = gb.LatticeFermion(grid, Ls, gb.F32).import_lex(h)
    res = {}
    for mode in (True, 2, False):
        D.set_fast_kernel(mode)
        for name, fn in (("dhop0", lambda o: D.Dhop(fin, o, 0)), ("dhop1", lambda o: D.Dhop(fin, o, 1)), ("M", lambda o: D.M(fin, o)), ("Mdag", lambda o: D.Mdag(fin, o))):
            o = gb.LatticeFermion(grid, Ls, gb.F32); fn(o); res[(mode, name)] = o.export_lex().astype(np.complex128)
    for name in ("dhop0", "dhop1", "M", "Mdag"):
        ref = res[(False, name)]
        for mode in (True, 2):
            a = res[(mode, name)]
            num = np.linalg.norm((a - ref).reshape(a.shape[0], -1), axis=1); den = np.linalg.norm(ref.reshape(a.shape[0], -1), axis=1)
            e = float(np.max(num / np.maximum(den, 1e-300)))
            if not e < 8e-6: bad.append((dims, Ls, kind, coln, mode, name, e))
    nrun += 1
print("shapes", nrun, "bad", len(bad), bad[:8], flush=True)
sys.exit(1 if bad else 0)
