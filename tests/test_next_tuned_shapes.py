"""Edge shapes of the tuned fp32 kernels (dhop_col_kernel, dhop_fast_kernel, smat_kernel) that the measured suite does not visit:
columns of one, two and three z-planes, a z extent whose largest divisor below 16 is odd, or 2, two sites along t (forward and backward
neighbour coincide), x extents of two and four sites (micro-blocks narrower than 4), Ls = 24 and 32 (micro-block kernel only),
z-chunks that do not divide the lattice, a 4D volume that is not a multiple of the 16-site tile (s-space kernel's ragged last tile).
Every case: Dhop +-dag through the default kernel selection, the micro-block kernel and the generic kernel against the fp64 oracle,
the checkerboard hops, and M (which runs the dense s-space kernel where Ls allows).

They passed on the B200 at the end of round 1 (GPUTEST_r01) and also run on the CPU mock of the
library, where the tuned kernels execute with one fibre per CUDA thread (tests/test_next_on_cpu_mock.py)."""
import numpy as np
import pytest

import grid_b200 as gb
from oracle import pyoracle as po
from test_gpu_parity import Setup, site_rel_err, TOL_HOP, TOL_COMPOSITE

pytestmark = [pytest.mark.gpu]

# dims, Ls, kind, column height passed to set_tiling (0 = default 16), what the shape is for
SHAPES = [
    ((8, 4, 6, 4), 8, "dwf", 0, "column = whole z extent (6)"),
    ((8, 4, 6, 4), 8, "dwf", 1, "columns of ONE plane"),
    ((8, 4, 6, 4), 8, "dwf", 2, "columns of two planes"),
    ((8, 4, 6, 4), 8, "dwf", 3, "columns of three planes (= ring depth)"),
    ((8, 4, 20, 2), 8, "dwf", 0, "Lz = 20 -> two columns of 10; Lt = 2"),
    ((8, 8, 22, 2), 12, "mobius", 0, "Lz = 22 -> two columns of 11 (odd height); Ls 12"),
    ((8, 4, 34, 2), 8, "dwf", 0, "Lz = 34 -> N falls to 2"),
    ((16, 4, 2, 2), 16, "dwf", 0, "Lz = 2, Lt = 2: every z and t neighbour is the same site"),
    ((4, 8, 4, 4), 8, "dwf", 0, "Lx = 4: micro-blocks two wide (no column kernel)"),
    ((2, 16, 4, 4), 8, "mobius", 0, "Lx = 2: one site per row, x neighbours wrap onto the same column"),
    ((16, 2, 4, 4), 8, "dwf", 0, "Ly = 2"),
    ((8, 4, 6, 4), 24, "dwf", 0, "Ls = 24 (micro-block kernel, 384 threads)"),
    ((8, 4, 4, 2), 32, "mobius", 0, "Ls = 32 (micro-block kernel, 512 threads)"),
    ((8, 4, 10, 4), 16, "dwf", 4, "z-chunk 4 does not divide Lz = 10"),
    ((6, 2, 2, 2), 8, "dwf", 0, "24 sites per parity: ragged 16-site tiles (generic hop, s-space kernel)"),
]


@pytest.fixture(scope="module")
def ctx():
    c = gb.Context(0)
    yield c
    c.synchronize()


@pytest.mark.parametrize("dims,Ls,kind,col_n,why", SHAPES, ids=[f"{'x'.join(map(str, s[0]))}_Ls{s[1]}_n{s[3]}" for s in SHAPES])
def test_tuned_kernels_at_edge_shapes(ctx, dims, Ls, kind, col_n, why):
    kw = dict(b=1.5, c=0.5) if kind == "mobius" else {}
    s = Setup(ctx, dims, Ls, kind, seed=21, **kw)
    prec = gb.F32
    op = s.dev[prec]
    if col_n:
        op.set_tiling(0, col_n, 0)
    h = s.host(22, prec)
    fin = s.field(prec).import_lex(h)
    h64 = h.astype(np.complex128)
    for dag in (0, 1):
        ref = s.oracle[gb.F64].apply(po.OP_DHOP, h64, dag=dag)
        outs = {}
        for mode, name in ((True, "default"), (2, "micro-block"), (False, "generic")):
            op.set_fast_kernel(mode)
            o = s.field(prec)
            op.Dhop(fin, o, dag)
            outs[name] = o.export_lex()
            assert site_rel_err(outs[name], ref) < TOL_HOP[prec], (why, name, dag)
        op.set_fast_kernel(True)
        assert site_rel_err(outs["default"], outs["generic"]) < 2 * TOL_HOP[prec], (why, dag)
    # checkerboard hops
    for cb_in, opc in ((gb.Odd, po.OP_DHOP_EO), (gb.Even, po.OP_DHOP_OE)):
        hh = po.pick_checkerboard(s.dims, s.Ls, cb_in, h)
        fi, fo = s.field(prec, gb.HALF).import_lex(hh), s.field(prec, gb.HALF)
        fi.set_checkerboard(cb_in)
        (op.DhopEO if cb_in == gb.Odd else op.DhopOE)(fi, fo, 0)
        ref = s.oracle[gb.F64].apply(opc, hh.astype(np.complex128), cb_in=cb_in)
        assert site_rel_err(fo.export_lex(), ref) < TOL_HOP[prec], (why, "cb", cb_in)
    # M and Mdag: hop + fifth-dimension operator (dense s-space kernel for Ls = 8, 12, 16)
    for dag, opc in ((0, po.OP_M), (1, po.OP_MDAG)):
        o = s.field(prec)
        (op.Mdag if dag else op.M)(fin, o)
        ref = s.oracle[gb.F64].apply(opc, h64)
        assert site_rel_err(o.export_lex(), ref) < TOL_COMPOSITE[prec], (why, "M", dag)
    # the Schur operator on the odd checkerboard: MooeeInv and Meooe through the s-space kernel
    ho = po.pick_checkerboard(s.dims, s.Ls, gb.Odd, h)
    fi, fo = s.field(prec, gb.HALF).import_lex(ho), s.field(prec, gb.HALF)
    fi.set_checkerboard(gb.Odd)
    gb.SchurDiagMooeeOperator(op).Mpc(fi, fo)
    ref = s.oracle[gb.F64].apply(po.OP_MPC, ho.astype(np.complex128), cb_in=gb.Odd)
    assert site_rel_err(fo.export_lex(), ref) < 4 * TOL_COMPOSITE[prec], (why, "Mpc")
