"""ctypes binding of the COMPILED REFERENCE (oracle/_ref/libgridref.so = unmodified paboyle/Grid CPU code + oracle/gridref_capi.cc).

TEST INFRASTRUCTURE ONLY -- same rule as the oracle: only tests/, the fixture generator tests/golden/make_golden.py and
bench.py's cpu_baseline / --impl reference legs import this.  The product package grid_b200 never does.

The library is built by `make -C oracle -f Makefile.ref` in the container that has /root/reference (see the Makefile's header
for what stands in for autotools' Config.h and the downloaded Eigen tree).  The built .so travels to the GPU box with the
repository snapshot; nothing here reads /root/reference at run time.

RefOp has the interface of pyoracle.OracleOp, so every check can be run against either.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libgridref.so")
_LIB = None

(OP_DHOP, OP_DHOP_OE, OP_DHOP_EO, OP_M, OP_MDAG, OP_MEOOE, OP_MEOOE_DAG, OP_MOOEE, OP_MOOEE_DAG, OP_MOOEE_INV,
 OP_MOOEE_INV_DAG, OP_MPC, OP_MPC_DAG, OP_HERMOP, OP_DW, OP_MEOOE5D, OP_MEOOEDAG5D, OP_DMINUS, OP_DMINUS_DAG) = range(19)
# physical 4D <-> 5D maps
IMPORT_PHYSICAL_SOURCE, IMPORT_UNPHYSICAL, EXPORT_PHYSICAL_SOLUTION, EXPORT_PHYSICAL_SOURCE = range(4)
KIND_WILSON, KIND_CAYLEY, KIND_STAGGERED = 0, 1, 2
OPT_GENERIC, OPT_HAND_UNROLL = 0, 1


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError(f"{_PATH} not built (make -C oracle -f Makefile.ref needs /root/reference)")
        L = C.CDLL(_PATH)
        L.gref_op_create.restype = C.c_void_p
        L.gref_op_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.gref_op_destroy.argtypes = [C.c_void_p]
        L.gref_op_import_gauge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gref_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.gref_apply.restype = C.c_int
        L.gref_pick_checkerboard.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_set_checkerboard.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_multishift_cg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gref_dhop_dir.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.gref_deriv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.gref_deriv_eo.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gref_deriv_eo.restype = C.c_int
        L.gref_nersc_write.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_char_p, C.c_int]
        L.gref_nersc_read.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_char_p, C.c_void_p]
        L.gref_physical.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_physical.restype = C.c_int
        L.gref_redblack_source.argtypes = [C.c_void_p] * 4
        L.gref_redblack_solution.argtypes = [C.c_void_p] * 4
        L.gref_schur_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_cg.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_relup_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.gref_multishift_mixed_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gref_mixed_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gref_mixed_cg_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gref_time_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.gref_time_apply.restype = C.c_double
        L.gref_init.argtypes = [C.c_int]
        L.gref_set_kernel_opt.argtypes = [C.c_int]
        # Grid_init prints its banner on stdout; keep the caller's stdout clean (bench.py prints ONE JSON line there)
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            L.gref_init(int(os.environ.get("GRIDREF_THREADS", "0")))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        _LIB = L
    return _LIB


def _cdtype(prec):
    return np.complex64 if prec == 0 else np.complex128


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().gref_num_threads()


def nsimd(prec):
    return lib().gref_nsimd(prec)


def set_kernel_opt(opt):
    """0 = WilsonKernelsStatic::OptGeneric (the reference's default), 1 = OptHandUnroll (--dslash-unroll)."""
    lib().gref_set_kernel_opt(opt)


class RefOp:
    """The reference's own WilsonFermion (kind 0) / DomainWallFermion or MobiusFermion (kind 1) /
    ImprovedStaggeredFermion (kind 2: M5,b,c carry c1,c2,u0), fp32 (prec 0) or fp64 (prec 1)."""

    def __init__(self, kind, dims, Ls, mass, M5=1.8, b=1.0, c=0.0, prec=1):
        self.kind, self.dims, self.Ls, self.prec = kind, tuple(dims), Ls, prec
        self.V4 = int(np.prod(dims))
        self.h = lib().gref_op_create(kind, (C.c_int * 4)(*dims), Ls, mass, M5, b, c, prec)

    def __del__(self):
        try:
            if self.h:
                lib().gref_op_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def import_gauge(self, Umu, phases=None):
        """kind 2 (staggered): fat = thin = Umu, like benchmarks/Benchmark_staggered.cc:92-96"""
        U = np.ascontiguousarray(Umu, dtype=_cdtype(self.prec))
        assert U.shape == (self.V4, 4, 3, 3)
        ph = None if phases is None else np.ascontiguousarray(np.asarray(phases, dtype=np.complex128))
        lib().gref_op_import_gauge(self.h, _ptr(U), _ptr(ph) if ph is not None else None)

    def _half(self, n):
        assert n in (self.V4 * self.Ls, self.V4 * self.Ls // 2), n
        return 1 if n == self.V4 * self.Ls // 2 else 0

    def apply(self, which, x, dag=0, cb_in=0):
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        out = np.empty_like(x)
        rc = lib().gref_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, self._half(x.shape[0]))
        assert rc == 0, rc
        return out

    def time_apply(self, which, x, ncall, dag=0, cb_in=0):
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        out = np.empty_like(x)
        return lib().gref_time_apply(self.h, which, _ptr(x), _ptr(out), dag, cb_in, self._half(x.shape[0]), ncall)

    def pick_checkerboard(self, cb, full):
        full = np.ascontiguousarray(full, dtype=_cdtype(self.prec))
        half = np.empty((full.shape[0] // 2,) + full.shape[1:], dtype=full.dtype)
        lib().gref_pick_checkerboard(self.h, cb, _ptr(half), _ptr(full))
        return half

    def set_checkerboard(self, cb, full, half):
        full = np.ascontiguousarray(full, dtype=_cdtype(self.prec)).copy()
        half = np.ascontiguousarray(half, dtype=_cdtype(self.prec))
        lib().gref_set_checkerboard(self.h, cb, _ptr(full), _ptr(half))
        return full

    def cg(self, cb, src, tol, maxit, guess=None):
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src) if guess is None else np.ascontiguousarray(guess, dtype=_cdtype(self.prec)).copy()
        it = np.zeros(2, dtype=np.int32)
        tr = np.zeros(1, dtype=np.float64)
        lib().gref_cg(self.h, cb, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(tr))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(tr[0]))

    # ---- single hop legs and force terms (full grid)
    def dhop_dir(self, x, dir, disp):
        """FermionOperator::DhopDir(in, out, dir, disp): the leg of the hopping term that reads x + disp * dir (dir = 0..3)."""
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        out = np.empty_like(x)
        lib().gref_dhop_dir(self.h, _ptr(x), _ptr(out), dir, disp)
        return out

    def deriv(self, which, U, V, dag=0):
        """which = 0: DhopDeriv(mat, U, V, dag); 1: MDeriv(mat, U, V, dag).  Returns mat as [V4,4,3,3]."""
        U = np.ascontiguousarray(U, dtype=_cdtype(self.prec)); V = np.ascontiguousarray(V, dtype=_cdtype(self.prec))
        mat = np.zeros((self.V4, 4, 3, 3), dtype=U.dtype)
        lib().gref_deriv(self.h, which, _ptr(mat), _ptr(U), _ptr(V), dag)
        return mat

    def deriv_eo(self, which, U, V, dag=0):
        """which 0 = MeoDeriv (U Even, V Odd), 1 = MoeDeriv (U Odd, V Even): only the sites of U's parity of the returned full-lattice
        [V4,4,3,3] are written; 2 = SchurDifferentiableOperator::MpcDeriv, 3 = MpcDagDeriv (U, V Odd; the whole force)."""
        U = np.ascontiguousarray(U, dtype=_cdtype(self.prec)); V = np.ascontiguousarray(V, dtype=_cdtype(self.prec))
        mat = np.zeros((self.V4, 4, 3, 3), dtype=U.dtype)
        assert which in (2, 3)
        rc = lib().gref_deriv_eo(self.h, which, _ptr(mat), _ptr(U), _ptr(V))
        assert rc == 0, rc
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
        lib().gref_multishift_cg(self.h, cb, _ptr(src), n, _ptr(poles), _ptr(tols), maxit, _ptr(res), _ptr(it), _ptr(tr))
        return res, dict(iterations=[int(x) for x in it[:n]], true_residual=[float(x) for x in tr], iterations_to_complete=int(it[n]), converged=int(it[n + 1]))

    # ---- physical 4D <-> 5D maps, SchurRedBlackDiagMooeeSolve
    def physical(self, which, x):
        """which = IMPORT_PHYSICAL_SOURCE / IMPORT_UNPHYSICAL (x: [V4,4,3] -> [V4*Ls,4,3]) or EXPORT_PHYSICAL_SOLUTION /
        EXPORT_PHYSICAL_SOURCE ([V4*Ls,4,3] -> [V4,4,3])."""
        x = np.ascontiguousarray(x, dtype=_cdtype(self.prec))
        n_out = self.V4 * self.Ls if which < 2 else self.V4
        assert x.shape[0] == (self.V4 if which < 2 else self.V4 * self.Ls), x.shape
        out = np.empty((n_out,) + x.shape[1:], dtype=x.dtype)
        rc = lib().gref_physical(self.h, which, _ptr(x), _ptr(out))
        assert rc == 0, rc
        return out

    def redblack_source(self, src):
        """SchurRedBlack*Solve::RedBlackSource: full-lattice src -> (src_e, src_o')"""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        e = np.empty((src.shape[0] // 2,) + src.shape[1:], dtype=src.dtype)
        o = np.empty_like(e)
        lib().gref_redblack_source(self.h, _ptr(src), _ptr(e), _ptr(o))
        return e, o

    def redblack_solution(self, sol_o, src_e):
        sol_o = np.ascontiguousarray(sol_o, dtype=_cdtype(self.prec)); src_e = np.ascontiguousarray(src_e, dtype=_cdtype(self.prec))
        sol = np.zeros((2 * sol_o.shape[0],) + sol_o.shape[1:], dtype=sol_o.dtype)
        lib().gref_redblack_solution(self.h, _ptr(sol_o), _ptr(src_e), _ptr(sol))
        return sol

    def schur_solve(self, src, tol, maxit):
        """M sol = src on the full lattice through the red-black Schur decomposition + CG (zero guess)."""
        src = np.ascontiguousarray(src, dtype=_cdtype(self.prec))
        sol = np.zeros_like(src)
        it = np.zeros(2, dtype=np.int32)
        rs = np.zeros(2, dtype=np.float64)
        lib().gref_schur_solve(self.h, _ptr(src), _ptr(sol), tol, maxit, _ptr(it), _ptr(rs))
        return sol, dict(iterations=int(it[0]), converged=int(it[1]), true_residual=float(rs[0]), unprec_residual=float(rs[1]))


def mixed_cg(op_d, op_f, cb, src_d, tol, maxinner, maxouter):
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    sol = np.zeros_like(src)
    it = np.zeros(4, dtype=np.int32)
    tr = np.zeros(1, dtype=np.float64)
    lib().gref_mixed_cg(op_d.h, op_f.h, cb, _ptr(src), _ptr(sol), tol, maxinner, maxouter, _ptr(it), _ptr(tr))
    return sol, dict(inner=int(it[0]), outer=int(it[1]), final=int(it[2]), converged=int(it[3]), true_residual=float(tr[0]))


def mixed_cg_batched(op_d, op_f, cb, srcs_d, tol, maxinner, maxouter, maxpatch):
    """MixedPrecisionConjugateGradientBatched(tol, maxinner, maxouter, maxpatch, ..., Linop_f, Linop_d)(srcs, sols) from zero guesses;
    the iteration counts are the ones the class logs (it keeps no members for them)."""
    srcs = np.ascontiguousarray(srcs_d, dtype=np.complex128)
    nb = srcs.shape[0]
    sols = np.zeros_like(srcs)
    it = np.zeros(1 + 2 * nb, dtype=np.int32)
    lib().gref_mixed_cg_batched(op_d.h, op_f.h, cb, nb, _ptr(srcs), _ptr(sols), tol, maxinner, maxouter, maxpatch, _ptr(it))
    return sols, dict(outer=int(it[0]), inner=[int(v) for v in it[1:1 + nb]], final=[int(v) for v in it[1 + nb:]])


def relup_cg(op_d, op_f, cb, src_d, tol, maxit, delta):
    """ConjugateGradientReliableUpdate(tol, maxit, delta, ..., Linop_f, Linop_d)(src, sol) with a zero guess."""
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    sol = np.zeros_like(src)
    it = np.zeros(4, dtype=np.int32)
    tr = np.zeros(1, dtype=np.float64)
    lib().gref_relup_cg(op_d.h, op_f.h, cb, _ptr(src), _ptr(sol), tol, maxit, delta, _ptr(it), _ptr(tr))
    return sol, dict(iterations=int(it[0]), reliable_updates=int(it[1]), cleanup_iterations=int(it[2]), converged=int(it[3]), true_residual=float(tr[0]))


def nersc_write(dims, U, path, two_row=0):
    """NerscIO::writeConfiguration of the reference (IEEE64BIG; two_row drops the third row)."""
    U = np.ascontiguousarray(U, dtype=np.complex128)
    lib().gref_nersc_write((C.c_int * 4)(*dims), _ptr(U), str(path).encode(), two_row)


def nersc_read(dims, path):
    """NerscIO::readConfiguration of the reference (with its checksum / plaquette / link-trace QA). Returns (U, plaquette, link_trace)."""
    U = np.empty((int(np.prod(dims)), 4, 3, 3), dtype=np.complex128)
    pl = np.zeros(2)
    lib().gref_nersc_read((C.c_int * 4)(*dims), _ptr(U), str(path).encode(), _ptr(pl))
    return U, float(pl[0]), float(pl[1])


def multishift_mixed_cg(op_d, op_f, cb, src_d, poles, tols, maxit, relup_freq):
    """ConjugateGradientMultiShiftMixedPrec(maxit, shifts, ..., Linop_f, relup_freq)(Linop_d, src, results)."""
    src = np.ascontiguousarray(src_d, dtype=np.complex128)
    poles = np.ascontiguousarray(poles, dtype=np.float64); tols = np.ascontiguousarray(tols, dtype=np.float64)
    n = len(poles)
    res = np.zeros((n,) + src.shape, dtype=src.dtype)
    it = np.zeros(n + 2, dtype=np.int32)
    tr = np.zeros(n, dtype=np.float64)
    lib().gref_multishift_mixed_cg(op_d.h, op_f.h, cb, _ptr(src), n, _ptr(poles), _ptr(tols), maxit, relup_freq, _ptr(res), _ptr(it), _ptr(tr))
    return res, dict(iterations=[int(x) for x in it[:n]], true_residual=[float(x) for x in tr], iterations_to_complete=int(it[n]), cleanups=int(it[n + 1]))
