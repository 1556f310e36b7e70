#!/usr/bin/env python
"""Generates tests/golden/next_golden.npz from the COMPILED REFERENCE (oracle/_ref/libgridref.so, see make_golden.py) for the
SURVEY 8(f) "next" rows.  Inputs (U, src4, src5, src_stag) are those of dirac_golden.npz and are not stored again.

    make -C oracle -f Makefile.ref && python tests/golden/make_golden_next.py

Row f1 (full propagator solve, ref: SchurRedBlack.h:238-290,385-430 ; CayleyFermion5DImplementation.h:58-153):
  mobius/DMINUS, mobius/DMINUS_DAG                       CayleyFermion5D::Dminus / DminusDag on src5
  mobius/physical/{0,1}                                  ImportPhysicalFermionSource / ImportUnphysicalFermion of src4
  mobius/physical/{2,3}                                  ExportPhysicalFermionSolution / ExportPhysicalFermionSource of src5
  <op>/rb_source/{e,o}, <op>/rb_solution                 SchurRedBlack*Solve::RedBlackSource(src) ; RedBlackSolution(pick(Odd,src), e)
  <op>/schur_solve/{solution,iterations,true_residual,unprec_residual}   SchurRedBlack*Solve(ConjugateGradient(1e-8))(M, src, sol)
for op in wilson, dwf, mobius (b=1.5,c=0.5), stag.

Row f2 (single hop legs and force terms, ref: WilsonFermion5DImplementation.h:183-275 ; CayleyFermion5DImplementation.h:347-360), A = src,
B = the field stored as "src_b" (second random source):
  mobius/dhop_dir/{dir}_{disp}  for (1,+1), (3,-1);   <op>/deriv/{which}_{dag}  which 0 DhopDeriv, 1 MDeriv, for op in wilson, mobius
  <op>/mpc_deriv/{0,1}  SchurDifferentiableOperator::MpcDeriv / MpcDagDeriv(Force, pick(Odd, src), pick(Odd, src_b))

Row f3 (ConjugateGradientMultiShift, ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:84-343) on the Odd checkerboard of
the same sources, poles MS_POLES, tolerances MS_TOLS:
  <op>/multishift/{solutions [nshift, nsite_cb, ...], iterations [nshift], true_residual [nshift], iterations_to_complete}
for op in mobius, stag.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as pr                  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MS_POLES, MS_TOLS = [0.01, 0.05, 0.3, 1.0], [1e-8, 1e-8, 1e-7, 1e-6]


def main():
    G = np.load(os.path.join(HERE, "dirac_golden.npz"))
    DIMS, LS = tuple(int(x) for x in G["dims"]), int(G["Ls"])
    U, src4, src5, srcs = G["U"], G["src4"], G["src5"], G["src_stag"]
    out = {"ms_poles": np.array(MS_POLES), "ms_tols": np.array(MS_TOLS)}
    ops = {
        "wilson": (pr.RefOp(0, DIMS, 1, 0.1, prec=1), src4),
        "dwf": (pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.0, 0.0, prec=1), src5),
        "mobius": (pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=1), src5),
        "stag": (pr.RefOp(2, DIMS, 1, 0.1, 9.0 / 8.0, -1.0 / 24.0, 1.0, prec=1), srcs),
    }
    for name, (op, src) in ops.items():
        op.import_gauge(U)
        if name == "mobius":
            out["mobius/DMINUS"], out["mobius/DMINUS_DAG"] = op.apply(pr.OP_DMINUS, src), op.apply(pr.OP_DMINUS_DAG, src)
            for w in range(4):
                out[f"mobius/physical/{w}"] = op.physical(w, src4 if w < 2 else src5)
        if name in ("mobius", "stag"):
            e, o = op.redblack_source(src)
            out[f"{name}/rb_source/e"], out[f"{name}/rb_source/o"] = e, o
            out[f"{name}/rb_solution"] = op.redblack_solution(op.pick_checkerboard(1, src), e)
        if name in ("wilson", "mobius"):
            from grid_b200 import synthetic as syn
            srcb = syn.random_fermion(DIMS, op.Ls, seed=105)
            out[f"{name}/src_b"] = srcb
            if name == "mobius":
                out["mobius/dhop_dir/1_1"], out["mobius/dhop_dir/3_-1"] = op.dhop_dir(src, 1, 1), op.dhop_dir(src, 3, -1)
            for which in (0, 1):
                for dag in (0, 1):
                    out[f"{name}/deriv/{which}_{dag}"] = op.deriv(which, src, srcb, dag)
            uo, vo = op.pick_checkerboard(1, src), op.pick_checkerboard(1, srcb)
            out[f"{name}/mpc_deriv/0"], out[f"{name}/mpc_deriv/1"] = op.deriv_eo(2, uo, vo), op.deriv_eo(3, uo, vo)
        if name in ("mobius", "stag"):
            xs, info = op.multishift_cg(1, op.pick_checkerboard(1, src), MS_POLES, MS_TOLS, 5000)
            out[f"{name}/multishift/solutions"] = xs
            for k in ("iterations", "true_residual", "iterations_to_complete"):
                out[f"{name}/multishift/{k}"] = np.array(info[k])
        x, info = op.schur_solve(src, 1e-8, 5000)
        out[f"{name}/schur_solve/solution"] = x
        for k in ("iterations", "true_residual", "unprec_residual"):
            out[f"{name}/schur_solve/{k}"] = np.array(info[k])
    # Row f3, ConjugateGradientReliableUpdate (ref: ConjugateGradientReliableUpdate.h:80-270; Test_dwf_relupcg_prec.cc:88-104): Delta = 0.1
    opf = pr.RefOp(1, DIMS, LS, 0.1, 1.8, 1.5, 0.5, prec=0)
    opf.import_gauge(U)
    x, info = pr.relup_cg(ops["mobius"][0], opf, 1, ops["mobius"][0].pick_checkerboard(1, src5), 1e-8, 5000, 0.1)
    out["mobius/relup_cg/solution"] = x
    for k in ("iterations", "reliable_updates", "cleanup_iterations", "true_residual"):
        out[f"mobius/relup_cg/{k}"] = np.array(info[k])
    # Row f3, ConjugateGradientMultiShiftMixedPrec (ref: ConjugateGradientMultiShiftMixedPrec.h:128-410), reliable update every 20 iterations
    xs, info = pr.multishift_mixed_cg(ops["mobius"][0], opf, 1, ops["mobius"][0].pick_checkerboard(1, src5), MS_POLES, MS_TOLS, 5000, 20)
    out["mobius/multishift_mixed/solutions"] = xs
    for k in ("iterations", "true_residual", "iterations_to_complete"):
        out[f"mobius/multishift_mixed/{k}"] = np.array(info[k])
    path = os.path.join(HERE, "next_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB", file=sys.stderr)


if __name__ == "__main__":
    main()
