"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs.  The product package grid_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# op codes, identical to gb_opcode in include/gridb200.h
(OP_DHOP, OP_DHOP_OE, OP_DHOP_EO, OP_M, OP_MDAG, OP_MEOOE, OP_MEOOE_DAG, OP_MOOEE, OP_MOOEE_DAG, OP_MOOEE_INV,
 OP_MOOEE_INV_DAG, OP_MPC, OP_MPC_DAG, OP_HERMOP, OP_DW, OP_MEOOE5D, OP_MEOOEDAG5D, OP_DMINUS, OP_DMINUS_DAG) = range(19)
# physical 4D <-> 5D maps
IMPORT_PHYSICAL_SOURCE, IMPORT_UNPHYSICAL, EXPORT_PHYSICAL_SOLUTION, EXPORT_PHYSICAL_SOURCE = range(4)
EVEN, ODD = 0, 1


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_op_create.restype = C.c_void_p
        L.orc_op_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_op_destroy.argtypes = [C.c_void_p]
        L.orc_op_import_gauge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_op_export_doubled.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_op_coeffs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_apply.restype = C.c_int
        L.orc_dhop_naive.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_pick_checkerboard.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_checkerboard.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_inner_product.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_multishift_cg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_dhop_dir.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_deriv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_deriv_eo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_physical.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_physical.restype = C.c_int
        L.orc_redblack_source.argtypes = [C.c_void_p] * 4
        L.orc_redblack_solution.argtypes = [C.c_void_p] * 4
        L.orc_schur_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_relup_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_multishift_mixed_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mixed_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_mixed_cg_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_time_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_time_apply.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        L.orc_stag_create.restype = C.c_void_p
        L.orc_stag_create.argtypes = [C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_stag_destroy.argtypes = [C.c_void_p]
        L.orc_stag_import_gauge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stag_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_stag_apply.restype = C.c_int
        L.orc_stag_cg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_stag_redblack_source.argtypes = [C.c_void_p] * 4
        L.orc_stag_redblack_solution.argtypes = [C.c_void_p] * 4
        L.orc_stag_schur_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_stag_dhop_naive.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_pick_checkerboard_bytes.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_checkerboard_bytes.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_stag_time_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_stag_time_apply.restype = C.c_double
        _LIB = L
    return _LIB


def _cdtype(prec):
    return np.complex64 if prec == 0 else np.complex128


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _L(dims):
    return (C.c_int * 4)(*dims)


class OracleOp:
    """CPU restatement of WilsonFermion (kind=0) / DomainWallFermion / MobiusFermion (kind=1)."""

    def __init__(self, kind, dims, Ls, mass, M5=1.8, b=1.0, c=0.0, prec=1):
        self.kind, self.dims, self.Ls, self.prec = kind, tuple(dims), Ls, prec
        self.V4 = int(np.prod(dims))
        self.h = lib().orc_op_create(kind, _L(dims), Ls, mass, M5, b, c, prec)

    def __del__(self):
        try:
            if self.h:
                lib().orc_op_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def import_gauge(self, Umu, phases=None):
        """Umu: complex array [V4,4,3,3] (lexicographic, x fastest)."""
        U = np.ascontiguousarray(Umu, dtype=_cdtype(self.prec))
        assert U.shape == (self.V4, 4, 3, 3)
        ph = None
        if phases is not None:
            ph = np.ascontiguousarray(np.asarray(phases, dtype=np.complex128))
        lib().orc_op_import_gauge(self.h, _ptr(U), _ptr(ph) if ph is not None else None)

    def doubled(self):
        out = np.empty((self.V4, 8, 3, 3), dtype=_cdtype(self.prec))
        lib().orc_op_export_doubled(self.h, _ptr(out))
        return out

    def coeffs(self):
        out = np.empty((9, self.Ls), dtype=np.float64)
        lib().orc_op_coeffs(self.h, _ptr(out))
        return dict(zip(["bs", "cs", "bee", "cee", "dee", "lee", "leem", "uee", "ueem"], out))

    def apply(self, which, x, dag=0, cb_in=0):
        """x: complex [nsite5,4,3]; full (V4*Ls sites) or half (V4/2*Ls sites) field."""
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        n = x.shape[0]
        half = 1 if n == self.V4 * self.Ls // 2 else 0
        assert n in (self.V4 * self.Ls, self.V4 * self.Ls // 2)
        out = np.empty_like(x)
        rc = lib().orc_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, half)
        assert rc == 0
        return out

    def time_apply(self, which, x, ncall, dag=0, cb_in=0):
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        half = 1 if x.shape[0] == self.V4 * self.Ls // 2 else 0
        out = np.empty_like(x)
        return lib().orc_time_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, half, ncall)

    def cg(self, cb, src, tol, maxit, guess=None):
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src) if guess is None else np.ascontiguousarray(guess, dtype=_cdtype(self.prec)).copy()
        it = np.zeros(2, dtype=np.int32)
        tr = np.zeros(1, dtype=np.float64)
        lib().orc_cg(self.h, cb, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(tr))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(tr[0]))

    # ---- single hop legs and force terms (full grid)
    def dhop_dir(self, x, dir, disp):
        """FermionOperator::DhopDir(in, out, dir, disp): the leg of the hopping term that reads x + disp * dir (dir = 0..3)."""
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        out = np.empty_like(x)
        lib().orc_dhop_dir(self.h, _ptr(x), _ptr(out), dir, disp)
        return out

    def deriv(self, which, U, V, dag=0):
        """which = 0: DhopDeriv(mat, U, V, dag); 1: MDeriv(mat, U, V, dag).  Returns mat as [V4,4,3,3]."""
        U = np.ascontiguousarray(U, dtype=_cdtype(self.prec)); V = np.ascontiguousarray(V, dtype=_cdtype(self.prec))
        mat = np.zeros((self.V4, 4, 3, 3), dtype=U.dtype)
        lib().orc_deriv(self.h, which, _ptr(mat), _ptr(U), _ptr(V), dag)
        return mat

    def deriv_eo(self, which, U, V, dag=0):
        """which 0 = MeoDeriv (U Even, V Odd), 1 = MoeDeriv (U Odd, V Even): only the sites of U's parity of the returned full-lattice
        [V4,4,3,3] are written; 2 = SchurDifferentiableOperator::MpcDeriv, 3 = MpcDagDeriv (U, V Odd; the whole force)."""
        U = np.ascontiguousarray(U, dtype=_cdtype(self.prec)); V = np.ascontiguousarray(V, dtype=_cdtype(self.prec))
        mat = np.zeros((self.V4, 4, 3, 3), dtype=U.dtype)
        lib().orc_deriv_eo(self.h, which, _ptr(mat), _ptr(U), _ptr(V), dag)
        return mat

    def multishift_cg(self, cb, src, poles, tols, maxit):
        """ConjugateGradientMultiShift on the Schur operator of checkerboard cb: (A + poles[s]) x_s = src.
        Returns ([nshift, nsite, ...] solutions, dict(iterations=[...], true_residual=[...], iterations_to_complete, converged))."""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        poles = np.ascontiguousarray(poles, dtype=np.float64); tols = np.ascontiguousarray(tols, dtype=np.float64)
        n = len(poles)
        res = np.zeros((n,) + src.shape, dtype=src.dtype)
        it = np.zeros(n + 2, dtype=np.int32)
        tr = np.zeros(n, dtype=np.float64)
        lib().orc_multishift_cg(self.h, 0, cb, _ptr(src), n, _ptr(poles), _ptr(tols), maxit, _ptr(res), _ptr(it), _ptr(tr))
        return res, dict(iterations=[int(x) for x in it[:n]], true_residual=[float(x) for x in tr], iterations_to_complete=int(it[n]), converged=int(it[n + 1]))

    # ---- physical 4D <-> 5D maps, SchurRedBlackDiagMooeeSolve
    def physical(self, which, x):
        """which = IMPORT_PHYSICAL_SOURCE / IMPORT_UNPHYSICAL (x: [V4,4,3] -> [V4*Ls,4,3]) or EXPORT_PHYSICAL_SOLUTION /
        EXPORT_PHYSICAL_SOURCE ([V4*Ls,4,3] -> [V4,4,3])."""
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        n_out = self.V4 * self.Ls if which < 2 else self.V4
        assert x.shape[0] == (self.V4 if which < 2 else self.V4 * self.Ls), x.shape
        out = np.empty((n_out,) + x.shape[1:], dtype=x.dtype)
        rc = lib().orc_physical(self.h, which, _ptr(x), _ptr(out))
        assert rc == 0, rc
        return out

    def redblack_source(self, src):
        """SchurRedBlack*Solve::RedBlackSource: full-lattice src -> (src_e, src_o')"""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        e = np.empty((src.shape[0] // 2,) + src.shape[1:], dtype=src.dtype)
        o = np.empty_like(e)
        lib().orc_redblack_source(self.h, _ptr(src), _ptr(e), _ptr(o))
        return e, o

    def redblack_solution(self, sol_o, src_e):
        sol_o = np.ascontiguousarray(sol_o, dtype=_cdtype(self.prec)); src_e = np.ascontiguousarray(src_e, dtype=_cdtype(self.prec))
        sol = np.zeros((2 * sol_o.shape[0],) + sol_o.shape[1:], dtype=sol_o.dtype)
        lib().orc_redblack_solution(self.h, _ptr(sol_o), _ptr(src_e), _ptr(sol))
        return sol

    def schur_solve(self, src, tol, maxit):
        """M sol = src on the full lattice through the red-black Schur decomposition + CG (zero guess)."""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src)
        it = np.zeros(2, dtype=np.int32)
        rs = np.zeros(2, dtype=np.float64)
        lib().orc_schur_solve(self.h, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(rs))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(rs[0]), unprec_residual=float(rs[1]))


class StagOracleOp:
    """CPU restatement of ImprovedStaggeredFermion (oracle/stag_oracle.hpp); fields are [nsite,3] complex colour vectors."""

    def __init__(self, dims, mass, c1=9.0 / 8.0, c2=-1.0 / 24.0, u0=1.0, prec=1):
        self.dims, self.prec, self.Ls = tuple(dims), prec, 1
        self.V4 = int(np.prod(dims))
        self.c1, self.c2, self.u0, self.mass = c1, c2, u0, mass
        self.h = lib().orc_stag_create(_L(dims), mass, c1, c2, u0, prec)

    def __del__(self):
        try:
            if self.h:
                lib().orc_stag_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def import_gauge(self, Uthin, Ufat=None):
        Ut = np.ascontiguousarray(Uthin, dtype=_cdtype(self.prec))
        Uf = Ut if Ufat is None else np.ascontiguousarray(Ufat, dtype=_cdtype(self.prec))
        assert Ut.shape == (self.V4, 4, 3, 3) and Uf.shape == Ut.shape
        lib().orc_stag_import_gauge(self.h, _ptr(Ut), _ptr(Uf))

    def apply(self, which, x, dag=0, cb_in=0):
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        assert x.shape[1:] == (3,) and x.shape[0] in (self.V4, self.V4 // 2)
        out = np.empty_like(x)
        rc = lib().orc_stag_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, 1 if x.shape[0] == self.V4 // 2 else 0)
        assert rc == 0
        return out

    def time_apply(self, which, x, ncall, dag=0, cb_in=0):
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        out = np.empty_like(x)
        return lib().orc_stag_time_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, 1 if x.shape[0] == self.V4 // 2 else 0, ncall)

    def cg(self, cb, src, tol, maxit, guess=None):
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src) if guess is None else np.ascontiguousarray(guess, dtype=_cdtype(self.prec)).copy()
        it = np.zeros(2, dtype=np.int32)
        tr = np.zeros(1, dtype=np.float64)
        lib().orc_stag_cg(self.h, cb, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(tr))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(tr[0]))

    def multishift_cg(self, cb, src, poles, tols, maxit):
        """ConjugateGradientMultiShift on the Schur operator of checkerboard cb: (A + poles[s]) x_s = src.
        Returns ([nshift, nsite, ...] solutions, dict(iterations=[...], true_residual=[...], iterations_to_complete, converged))."""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        poles = np.ascontiguousarray(poles, dtype=np.float64); tols = np.ascontiguousarray(tols, dtype=np.float64)
        n = len(poles)
        res = np.zeros((n,) + src.shape, dtype=src.dtype)
        it = np.zeros(n + 2, dtype=np.int32)
        tr = np.zeros(n, dtype=np.float64)
        lib().orc_multishift_cg(self.h, 1, cb, _ptr(src), n, _ptr(poles), _ptr(tols), maxit, _ptr(res), _ptr(it), _ptr(tr))
        return res, dict(iterations=[int(x) for x in it[:n]], true_residual=[float(x) for x in tr], iterations_to_complete=int(it[n]), converged=int(it[n + 1]))

    # ---- SchurRedBlackStaggeredSolve (ref: SchurRedBlack.h:294-349)
    def redblack_source(self, src):
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        e = np.empty((src.shape[0] // 2,) + src.shape[1:], dtype=src.dtype)
        o = np.empty_like(e)
        lib().orc_stag_redblack_source(self.h, _ptr(src), _ptr(e), _ptr(o))
        return e, o

    def redblack_solution(self, sol_o, src_e):
        sol_o = np.ascontiguousarray(sol_o, dtype=_cdtype(self.prec)); src_e = np.ascontiguousarray(src_e, dtype=_cdtype(self.prec))
        sol = np.zeros((2 * sol_o.shape[0],) + sol_o.shape[1:], dtype=sol_o.dtype)
        lib().orc_stag_redblack_solution(self.h, _ptr(sol_o), _ptr(src_e), _ptr(sol))
        return sol

    def schur_solve(self, src, tol, maxit):
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src)
        it = np.zeros(2, dtype=np.int32)
        rs = np.zeros(2, dtype=np.float64)
        lib().orc_stag_schur_solve(self.h, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(rs))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(rs[0]), unprec_residual=float(rs[1]))


def stag_dhop_naive(dims, Uthin, Ufat, c1, c2, u0, x, dag=0, prec=1):
    Ut = np.ascontiguousarray(Uthin, dtype=_cdtype(prec)); Uf = np.ascontiguousarray(Ufat, dtype=_cdtype(prec))
    x = np.ascontiguousarray(x, dtype=_cdtype(prec))
    out = np.empty_like(x)
    lib().orc_stag_dhop_naive(_L(dims), prec, _ptr(Ut), _ptr(Uf), c1, c2, u0, _ptr(x), _ptr(out), dag)
    return out


def pick_checkerboard_sites(dims, cb, full):
    """pickCheckerboard for a 4D field of any site type ([V4, ...] array)."""
    full = np.ascontiguousarray(full)
    half = np.empty((full.shape[0] // 2,) + full.shape[1:], dtype=full.dtype)
    lib().orc_pick_checkerboard_bytes(_L(dims), full.itemsize * int(np.prod(full.shape[1:])), cb, _ptr(half), _ptr(full))
    return half


def set_checkerboard_sites(dims, cb, full, half):
    assert full.flags.c_contiguous and half.flags.c_contiguous and full.dtype == half.dtype
    lib().orc_set_checkerboard_bytes(_L(dims), full.itemsize * int(np.prod(full.shape[1:])), cb, _ptr(full), _ptr(half))


def mixed_cg(op_d, op_f, cb, src_d, tol, maxinner, maxouter):
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    sol = np.zeros_like(src)
    it = np.zeros(4, dtype=np.int32)
    tr = np.zeros(1, dtype=np.float64)
    lib().orc_mixed_cg(op_d.h, op_f.h, cb, _ptr(src), _ptr(sol), tol, maxinner, maxouter, _ptr(it), _ptr(tr))
    return sol, dict(inner=int(it[0]), outer=int(it[1]), final=int(it[2]), converged=int(it[3]), true_residual=float(tr[0]))


def mixed_cg_batched(op_d, op_f, cb, srcs_d, tol, maxinner, maxouter, maxpatch):
    """MixedPrecisionConjugateGradientBatched (ref: ConjugateGradientMixedPrecBatched.h:79-207) from zero guesses.
    srcs_d: [nbatch, nsite, ...] complex128.  Returns (solutions, dict(outer, inner=[...], final=[...], true_residual=[...]))."""
    srcs = np.ascontiguousarray(srcs_d, dtype=np.complex128)
    nb = srcs.shape[0]
    sols = np.zeros_like(srcs)
    it = np.zeros(1 + 2 * nb, dtype=np.int32)
    tr = np.zeros(nb, dtype=np.float64)
    lib().orc_mixed_cg_batched(op_d.h, op_f.h, cb, nb, _ptr(srcs), _ptr(sols), tol, maxinner, maxouter, maxpatch, _ptr(it), _ptr(tr))
    return sols, dict(outer=int(it[0]), inner=[int(v) for v in it[1:1 + nb]], final=[int(v) for v in it[1 + nb:]], true_residual=[float(v) for v in tr])


def dhop_naive(dims, Ls, Umu, x, dag=0, prec=1):
    U = np.ascontiguousarray(Umu, dtype=_cdtype(prec))
    x = np.ascontiguousarray(x, dtype=_cdtype(prec))
    out = np.empty_like(x)
    lib().orc_dhop_naive(_L(dims), Ls, prec, _ptr(U), _ptr(x), _ptr(out), dag)
    return out


def pick_checkerboard(dims, Ls, cb, full):
    prec = 0 if full.dtype == np.complex64 else 1
    full = np.ascontiguousarray(full)
    half = np.empty((full.shape[0] // 2,) + full.shape[1:], dtype=full.dtype)
    lib().orc_pick_checkerboard(_L(dims), Ls, prec, cb, _ptr(half), _ptr(full))
    return half


def set_checkerboard(dims, Ls, cb, full, half):
    prec = 0 if full.dtype == np.complex64 else 1
    assert full.flags.c_contiguous and half.flags.c_contiguous and full.dtype == half.dtype
    lib().orc_set_checkerboard(_L(dims), Ls, prec, cb, _ptr(full), _ptr(half))


def inner_product(l, r):
    prec = 0 if l.dtype == np.complex64 else 1
    l = np.ascontiguousarray(l); r = np.ascontiguousarray(r, dtype=l.dtype)
    out = np.zeros(2)
    lib().orc_inner_product(l.shape[0], prec, _ptr(l), _ptr(r), _ptr(out))
    return complex(out[0], out[1])


def num_threads():
    return lib().orc_num_threads()


def relup_cg(op_d, op_f, cb, src_d, tol, maxit, delta):
    """ConjugateGradientReliableUpdate(tol, maxit, delta, ..., Linop_f, Linop_d)(src, sol) with a zero guess."""
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    sol = np.zeros_like(src)
    it = np.zeros(4, dtype=np.int32)
    tr = np.zeros(1, dtype=np.float64)
    lib().orc_relup_cg(op_d.h, op_f.h, cb, _ptr(src), _ptr(sol), tol, maxit, delta, _ptr(it), _ptr(tr))
    return sol, dict(iterations=int(it[0]), reliable_updates=int(it[1]), cleanup_iterations=int(it[2]), converged=int(it[3]), true_residual=float(tr[0]))


def multishift_mixed_cg(op_d, op_f, cb, src_d, poles, tols, maxit, relup_freq):
    """ConjugateGradientMultiShiftMixedPrec(maxit, shifts, ..., Linop_f, relup_freq)(Linop_d, src, results)."""
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    poles = np.ascontiguousarray(poles, dtype=np.float64); tols = np.ascontiguousarray(tols, dtype=np.float64)
    n = len(poles)
    res = np.zeros((n,) + src.shape, dtype=src.dtype)
    it = np.zeros(n + 2, dtype=np.int32)
    tr = np.zeros(n, dtype=np.float64)
    lib().orc_multishift_mixed_cg(op_d.h, op_f.h, cb, _ptr(src), n, _ptr(poles), _ptr(tols), maxit, relup_freq, _ptr(res), _ptr(it), _ptr(tr))
    return res, dict(iterations=[int(x) for x in it[:n]], true_residual=[float(x) for x in tr], iterations_to_complete=int(it[n]), cleanups=int(it[n + 1]))
