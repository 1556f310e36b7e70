#!/usr/bin/env python
"""Generates tests/golden/dirac_golden.npz from the COMPILED REFERENCE (oracle/_ref/libgridref.so: unmodified paboyle/Grid
built from /root/reference by oracle/Makefile.ref).  Run in the container that has /root/reference:

    make -C oracle -f Makefile.ref && python tests/golden/make_golden.py

Inputs are stored next to the outputs, so the fixtures do not depend on numpy's RNG streams.  Lattice 4^4, Ls = 4
(the reference needs z,t multiples of 4 under its AVX2 SIMD layout).  Everything fp64 unless the key ends in _f32.

Keys:  U [V4,4,3,3]; src4 [V4,4,3]; src5 [V4*Ls,4,3];
       <op>/<ENTRY>[/dag][/cbN]  for op in wilson, wilson_apbc (antiperiodic t), dwf, mobius (b=1.5,c=0.5);
       stag/... the same entries for ImprovedStaggeredFermion on src_stag [V4,3];
       <op>/cg/{iterations,true_residual,solution};  dwf/mixed_cg/{inner,outer,final,true_residual,solution}
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from grid_b200 import synthetic as syn          # noqa: E402
from oracle import pyref as pr                  # noqa: E402

DIMS, LS = (4, 4, 4, 4), 4
FULL = dict(DHOP=pr.OP_DHOP, M=pr.OP_M, MDAG=pr.OP_MDAG)
HALF = dict(MEOOE=pr.OP_MEOOE, MEOOE_DAG=pr.OP_MEOOE_DAG, MOOEE=pr.OP_MOOEE, MOOEE_DAG=pr.OP_MOOEE_DAG, MOOEE_INV=pr.OP_MOOEE_INV,
            MOOEE_INV_DAG=pr.OP_MOOEE_INV_DAG, MPC=pr.OP_MPC, MPC_DAG=pr.OP_MPC_DAG, HERMOP=pr.OP_HERMOP)


def main():
    out = {}
    U = syn.hot_gauge(DIMS, seed=101)
    src4 = syn.random_fermion(DIMS, 1, seed=102)
    src5 = syn.random_fermion(DIMS, LS, seed=103)
    out.update(U=U, src4=src4, src5=src5, dims=np.array(DIMS), Ls=np.array(LS), mass=np.array(0.1), M5=np.array(1.8))
    ops = {
        "wilson": (pr.RefOp(0, DIMS, 1, 0.1, prec=1), src4, 1, None),
        "wilson_apbc": (pr.RefOp(0, DIMS, 1, 0.1, prec=1), src4, 1, [1.0, 1.0, 1.0, -1.0]),
        "dwf": (pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.0, 0.0, prec=1), src5, LS, None),
        "mobius": (pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=1), src5, LS, None),
    }
    for name, (op, src, Ls, phases) in ops.items():
        op.import_gauge(U, phases)
        hop = name != "mobius"        # MobiusFermion's hopping term is DomainWallFermion's: not stored twice
        for k, code in FULL.items():
            if k == "DHOP" and not hop:
                continue
            for dag in ((0, 1) if k == "DHOP" else (0,)):
                out[f"{name}/{k}/dag{dag}"] = op.apply(code, src, dag=dag)
        if name == "dwf":
            for dag in (0, 1):
                out[f"{name}/DW/dag{dag}"] = op.apply(pr.OP_DW, src, dag=dag)
        he, ho = op.pick_checkerboard(0, src), op.pick_checkerboard(1, src)
        if Ls == 1 or name == "dwf":
            out[f"{name}/pick/cb0"], out[f"{name}/pick/cb1"] = he, ho
        for dag in ((0, 1) if hop else ()):
            out[f"{name}/DHOP_OE/dag{dag}"] = op.apply(pr.OP_DHOP_OE, he, dag=dag)
            out[f"{name}/DHOP_EO/dag{dag}"] = op.apply(pr.OP_DHOP_EO, ho, dag=dag)
        for k, code in HALF.items():
            for cb in ((0, 1) if k in ("MEOOE", "MEOOE_DAG") else (1,)):
                out[f"{name}/{k}/cb{cb}"] = op.apply(code, he if cb == 0 else ho, cb_in=cb)
        x, info = op.cg(1, ho, 1e-8, 5000)
        out[f"{name}/cg/solution"] = x
        out[f"{name}/cg/iterations"] = np.array(info["iterations"])
        out[f"{name}/cg/true_residual"] = np.array(info["true_residual"])
    # fp32 hop of the reference (DomainWallFermionF) and the mixed-precision solve of Test_dwf_mixedcg_prec.cc
    opf = pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.0, 0.0, prec=0)
    opf.import_gauge(U)
    out["dwf/DHOP/dag0_f32"] = opf.apply(pr.OP_DHOP, src5.astype(np.complex64))
    ho = ops["dwf"][0].pick_checkerboard(1, src5)
    x, info = pr.mixed_cg(ops["dwf"][0], opf, 1, ho, 1e-8, 10000, 50)
    out["dwf/mixed_cg/solution"] = x
    for k in ("inner", "outer", "final", "true_residual"):
        out[f"dwf/mixed_cg/{k}"] = np.array(info[k])
    # improved staggered (ImprovedStaggeredFermionD, fat = thin = U as in Benchmark_staggered.cc:92-96; c1=9/8, c2=-1/24, u0=1)
    rng = np.random.default_rng(104)
    srcs = rng.random((int(np.prod(DIMS)), 3)) + 1j * rng.random((int(np.prod(DIMS)), 3))
    out["src_stag"] = srcs
    st = pr.RefOp(2, DIMS, 1, 0.1, 9.0 / 8.0, -1.0 / 24.0, 1.0, prec=1)
    st.import_gauge(U)
    for dag in (0, 1):
        out[f"stag/DHOP/dag{dag}"] = st.apply(pr.OP_DHOP, srcs, dag=dag)
    out["stag/M/dag0"], out["stag/MDAG/dag0"] = st.apply(pr.OP_M, srcs), st.apply(pr.OP_MDAG, srcs)
    he, ho = st.pick_checkerboard(0, srcs), st.pick_checkerboard(1, srcs)
    out["stag/pick/cb0"], out["stag/pick/cb1"] = he, ho
    for dag in (0, 1):
        out[f"stag/DHOP_OE/dag{dag}"] = st.apply(pr.OP_DHOP_OE, he, dag=dag)
        out[f"stag/DHOP_EO/dag{dag}"] = st.apply(pr.OP_DHOP_EO, ho, dag=dag)
    for k in ("MEOOE", "MEOOE_DAG", "MOOEE", "MOOEE_INV", "MPC", "HERMOP"):
        for cb in (0, 1):
            out[f"stag/{k}/cb{cb}"] = st.apply(HALF[k], he if cb == 0 else ho, cb_in=cb)
    x, info = st.cg(1, ho, 1e-8, 5000)
    out["stag/cg/solution"], out["stag/cg/iterations"], out["stag/cg/true_residual"] = x, np.array(info["iterations"]), np.array(info["true_residual"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dirac_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB", file=sys.stderr)


if __name__ == "__main__":
    main()
