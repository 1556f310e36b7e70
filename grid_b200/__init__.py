"""grid_b200 -- Python host-side mirror of the reference's operator / solver interface over the C ABI.

The product is grid_b200/libgridb200.so (hand-written sm_100a CUDA behind include/gridb200.h).  This module only
binds it with ctypes and re-exposes the reference's class and method names so that tests and drivers read like
the reference's own (ref: Grid/qcd/action/fermion/FermionOperator.h:40-192, Grid/algorithms/LinearOperator.h:286-349,
Grid/algorithms/iterative/ConjugateGradient.h:42-258, ConjugateGradientMixedPrec.h:34-170).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked), but creating a Context
without a CUDA device raises, and a missing libgridb200.so raises at import of the library handle.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgridb200.so")

F32, F64 = 0, 1
Even, Odd = 0, 1
FULL, HALF = 0, 1
DaggerNo, DaggerYes = 0, 1
(OP_DHOP, OP_DHOP_OE, OP_DHOP_EO, OP_M, OP_MDAG, OP_MEOOE, OP_MEOOE_DAG, OP_MOOEE, OP_MOOEE_DAG, OP_MOOEE_INV,
 OP_MOOEE_INV_DAG, OP_MPC, OP_MPC_DAG, OP_HERMOP, OP_DW, OP_MEOOE5D, OP_MEOOEDAG5D, OP_DMINUS, OP_DMINUS_DAG) = range(19)
GB_OK, GB_ERR_INVALID, GB_ERR_CUDA, GB_ERR_NO_DEVICE, GB_ERR_NOT_CONVERGED, GB_ERR_COMM = 0, -1, -2, -3, -4, -5
UNIQUE_ID_BYTES = 128

# every symbol include/gridb200.h declares: (name, restype, argtypes)
_vp, _i, _d, _i64, _u64 = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_uint64
_pi, _pd, _pvp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p)
HERMOP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
SYMBOLS = [
    ("gb_context_create", _i, [_i, _pvp]), ("gb_context_destroy", _i, [_vp]), ("gb_last_error", C.c_char_p, []),
    ("gb_device_count", _i, []), ("gb_synchronize", _i, [_vp]), ("gb_timer_start", _i, [_vp]), ("gb_timer_stop", _i, [_vp, _pd]),
    ("gb_launch_count", _i64, [_vp]), ("gb_flush_l2", _i, [_vp]),
    ("gb_comm_unique_id", _i, [_vp]), ("gb_comm_init", _i, [_vp, _i, _i, _vp]), ("gb_comm_rank", _i, [_vp, _pi, _pi]),
    ("gb_comm_global_sum", _i, [_vp, _pd, _i]), ("gb_comm_barrier", _i, [_vp]),
    ("gb_grid_create", _i, [_vp, _pi, _pi, _pvp]), ("gb_geometry_query", _i, [_pi, _pi, _i, _pi, _pi, _pi]), ("gb_grid_destroy", _i, [_vp]), ("gb_grid_local_dims", _i, [_vp, _pi]),
    ("gb_grid_local_origin", _i, [_vp, _pi]),
    ("gb_fermion_create", _i, [_vp, _i, _i, _i, _pvp]), ("gb_staggered_fermion_create", _i, [_vp, _i, _i, _pvp]), ("gb_fermion_destroy", _i, [_vp]), ("gb_fermion_checkerboard", _i, [_vp]),
    ("gb_fermion_set_checkerboard_tag", _i, [_vp, _i]), ("gb_fermion_local_sites", _i64, [_vp]),
    ("gb_fermion_import", _i, [_vp, _vp, _i]), ("gb_fermion_export", _i, [_vp, _vp, _i]),
    ("gb_pick_checkerboard", _i, [_i, _vp, _vp]), ("gb_set_checkerboard", _i, [_vp, _vp]), ("gb_precision_change", _i, [_vp, _vp]),
    ("gb_fermion_random", _i, [_vp, _u64]),
    ("gb_zero", _i, [_vp]), ("gb_copy", _i, [_vp, _vp]), ("gb_scale", _i, [_vp, _d, _vp]), ("gb_axpy", _i, [_vp, _d, _vp, _vp]),
    ("gb_axpby", _i, [_vp, _d, _d, _vp, _vp]), ("gb_axpy_norm", _i, [_vp, _d, _vp, _vp, _pd]), ("gb_norm2", _i, [_vp, _pd]),
    ("gb_inner_product", _i, [_vp, _vp, _pd]),
    ("gb_gauge_create", _i, [_vp, _i, _pvp]), ("gb_gauge_destroy", _i, [_vp]), ("gb_gauge_import", _i, [_vp, _vp, _i]),
    ("gb_gauge_export", _i, [_vp, _vp, _i]), ("gb_gauge_random", _i, [_vp, _u64]), ("gb_gauge_unit", _i, [_vp]),
    ("gb_nersc_read_host", _i, [C.c_char_p, _vp, _vp]), ("gb_nersc_write_host", _i, [C.c_char_p, _vp, _pi, _i, C.c_char_p, C.c_char_p, _i]),
    ("gb_gauge_read_nersc", _i, [_vp, C.c_char_p, _vp]), ("gb_gauge_write_nersc", _i, [_vp, C.c_char_p, _i, C.c_char_p, C.c_char_p, _i]),
    ("gb_op_create_wilson", _i, [_vp, _vp, _d, _pd, _pvp]), ("gb_op_create_dwf", _i, [_vp, _vp, _i, _d, _d, _pd, _pvp]),
    ("gb_op_create_mobius", _i, [_vp, _vp, _i, _d, _d, _d, _d, _pd, _pvp]), ("gb_op_import_gauge", _i, [_vp, _vp]),
    ("gb_op_create_staggered", _i, [_vp, _vp, _vp, _d, _d, _d, _d, _pvp]), ("gb_op_import_gauge_staggered", _i, [_vp, _vp, _vp]),
    ("gb_op_destroy", _i, [_vp]), ("gb_op_Ls", _i, [_vp]), ("gb_op_apply", _i, [_vp, _i, _vp, _vp, _i]),
    ("gb_op_halo_exchange", _i, [_vp, _vp, _i, C.POINTER(C.c_int64)]),
    ("gb_op_dhop_host", _i, [_vp, _vp, _vp, _i, _i]), ("gb_op_set_tiling", _i, [_vp, _i, _i, _i]), ("gb_op_set_overlap", _i, [_vp, _i]), ("gb_op_set_halo_compression", _i, [_vp, _i]), ("gb_op_set_link_reconstruct", _i, [_vp, _i]), ("gb_op_set_fast_kernel", _i, [_vp, _i]),
    ("gb_cg_schur", _i, [_vp, _vp, _vp, _d, _i, _pi, _pd]), ("gb_cg", _i, [_vp, HERMOP_FN, _vp, _vp, _vp, _d, _i, _pi, _pd]),
    ("gb_mixed_cg_schur", _i, [_vp, _vp, _vp, _vp, _d, _i, _i, _pi, _pd]),
    ("gb_mixed_cg_schur_ex", _i, [_vp, _vp, _vp, _vp, _d, _d, _d, _i, _i, _pi, _pd]),
    ("gb_mixed_cg_batched_schur", _i, [_vp, _vp, _i, _vp, _vp, _d, _i, _i, _i, _i, _pi, _pd]),
    ("gb_op_dhop_dir", _i, [_vp, _vp, _vp, _i, _i]), ("gb_op_dhop_deriv", _i, [_vp, _vp, _vp, _vp, _i]), ("gb_op_mderiv", _i, [_vp, _vp, _vp, _vp, _i]),
    ("gb_op_meooe_deriv", _i, [_vp, _vp, _vp, _vp, _i]), ("gb_op_mpc_deriv", _i, [_vp, _vp, _vp, _vp, _i]),
    ("gb_relup_cg_schur", _i, [_vp, _vp, _vp, _vp, _d, _i, _d, _pi, _pd]),
    ("gb_cg_multishift_mixed_schur", _i, [_vp, _vp, _vp, _i, _pd, _pd, _i, _i, _pvp, _pi, _pd]),
    ("gb_cg_multishift_schur", _i, [_vp, _vp, _i, _pd, _pd, _i, _pvp, _pi, _pd]),
    ("gb_op_import_physical_fermion_source", _i, [_vp, _vp, _vp]), ("gb_op_import_unphysical_fermion", _i, [_vp, _vp, _vp]),
    ("gb_op_export_physical_fermion_solution", _i, [_vp, _vp, _vp]), ("gb_op_export_physical_fermion_source", _i, [_vp, _vp, _vp]),
    ("gb_schur_redblack_source", _i, [_vp, _vp, _vp, _vp]), ("gb_schur_redblack_solution", _i, [_vp, _vp, _vp, _vp]),
    ("gb_schur_solve", _i, [_vp, _vp, _vp, _d, _i, _i, _pi, _pd]),
    ("gb_schur_solve_mixed", _i, [_vp, _vp, _vp, _vp, _d, _i, _i, _pi, _pd]),
]

_LIB = None


class GridB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgridb200 error {code}: {msg}")
        self.code = code


def lib():
    """Load libgridb200.so; fails loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C grid_b200). grid_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != GB_OK:
        raise GridB200Error(rc, lib().gb_last_error().decode())


def _cdtype(prec):
    return np.complex64 if prec == F32 else np.complex128


def _prec_of(a):
    if a.dtype in (np.complex64, np.float32):
        return F32
    if a.dtype in (np.complex128, np.float64):
        return F64
    raise TypeError(f"unsupported dtype {a.dtype}")


def _i4(v):
    return (C.c_int * 4)(*[int(x) for x in v])


class Context:
    """Grid_init analogue: device, streams, communicator (ref: Grid/util/Init.cc:300-560)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().gb_context_create(device, C.byref(self.h)))
        self.device = device

    def synchronize(self):
        _chk(lib().gb_synchronize(self.h))

    def timer_start(self):
        _chk(lib().gb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _chk(lib().gb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib().gb_launch_count(self.h)

    def flush_l2(self):
        _chk(lib().gb_flush_l2(self.h))

    # ---- communicator
    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        _chk(lib().gb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, nranks, uid):
        buf = C.create_string_buffer(uid, UNIQUE_ID_BYTES) if uid is not None else None
        _chk(lib().gb_comm_init(self.h, rank, nranks, buf))
        self.rank, self.nranks = rank, nranks

    def global_sum(self, vals):
        arr = (C.c_double * len(vals))(*vals)
        _chk(lib().gb_comm_global_sum(self.h, arr, len(vals)))
        return list(arr)

    def barrier(self):
        _chk(lib().gb_comm_barrier(self.h))

    def close(self):
        if self.h:
            lib().gb_context_destroy(self.h)
            self.h = C.c_void_p()


def geometry_query(gdims, mpi, rank):
    """Host-only: (ldims, origin, [(fwd, bwd)] per dimension) of `rank` -- no device needed."""
    l, o, n = _i4([0] * 4), _i4([0] * 4), (C.c_int * 8)()
    _chk(lib().gb_geometry_query(_i4(gdims), _i4(mpi), rank, l, o, n))
    return tuple(l), tuple(o), [(n[2 * d], n[2 * d + 1]) for d in range(4)]


class GridCartesian:
    """4D grid + its red-black and 5D companions (ref: SpaceTimeGrid::makeFourDimGrid etc.)."""

    def __init__(self, ctx, gdims, mpi=(1, 1, 1, 1)):
        self.ctx, self.gdims, self.mpi = ctx, tuple(gdims), tuple(mpi)
        self.h = C.c_void_p()
        _chk(lib().gb_grid_create(ctx.h, _i4(gdims), _i4(mpi), C.byref(self.h)))
        l, o = _i4([0] * 4), _i4([0] * 4)
        lib().gb_grid_local_dims(self.h, l)
        lib().gb_grid_local_origin(self.h, o)
        self.ldims, self.origin = tuple(l), tuple(o)
        self.lsites = int(np.prod(self.ldims))

    def __del__(self):
        try:
            if self.h:
                lib().gb_grid_destroy(self.h)
        except Exception:
            pass


class LatticeFermion:
    """LatticeFermion{F,D} on the full (kind=FULL) or red-black (kind=HALF) 4D/5D grid."""
    SITE = (4, 3)   # SpinColourVector

    def __init__(self, grid, Ls=1, prec=F32, kind=FULL):
        self.grid, self.Ls, self.prec, self.kind = grid, Ls, prec, kind
        self.h = C.c_void_p()
        self._create()

    def _create(self):
        _chk(lib().gb_fermion_create(self.grid.h, self.Ls, self.prec, self.kind, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib().gb_fermion_destroy(self.h)
        except Exception:
            pass

    def like(self, kind=None, prec=None):
        cls = getattr(self, "_cls", type(self))     # a borrowed view (generic-CG callback) creates fields of the class it views
        return cls(self.grid, self.Ls, self.prec if prec is None else prec, self.kind if kind is None else kind)

    @property
    def local_sites(self):
        return lib().gb_fermion_local_sites(self.h)

    def Checkerboard(self):
        return lib().gb_fermion_checkerboard(self.h)

    def set_checkerboard(self, cb):
        lib().gb_fermion_set_checkerboard_tag(self.h, cb)

    def import_lex(self, host):
        """host: complex [nsites,4,3] in local lexicographic (full) or checkerboard-lexicographic (half) order."""
        host = np.ascontiguousarray(host)
        assert host.shape == (self.local_sites,) + self.SITE, (host.shape, self.local_sites)
        _chk(lib().gb_fermion_import(self.h, host.ctypes.data_as(C.c_void_p), _prec_of(host)))
        return self

    def export_lex(self, dtype=None):
        dt = _cdtype(self.prec) if dtype is None else dtype
        out = np.empty((self.local_sites,) + self.SITE, dtype=dt)
        _chk(lib().gb_fermion_export(self.h, out.ctypes.data_as(C.c_void_p), _prec_of(out)))
        return out

    def random(self, seed):
        _chk(lib().gb_fermion_random(self.h, seed))
        return self

    def zero(self):
        _chk(lib().gb_zero(self.h))
        return self


class LatticeStaggeredFermion(LatticeFermion):
    """LatticeStaggeredFermion{F,D}: one ColourVector per 4D site (ref: StaggeredImpl.h:60-75).  Host arrays are [nsites,3]."""
    SITE = (3,)

    def __init__(self, grid, Ls=1, prec=F32, kind=FULL):
        assert Ls == 1, "staggered fields are 4D"
        super().__init__(grid, 1, prec, kind)

    def _create(self):
        _chk(lib().gb_staggered_fermion_create(self.grid.h, self.prec, self.kind, C.byref(self.h)))


def pickCheckerboard(cb, half, full):
    _chk(lib().gb_pick_checkerboard(cb, half.h, full.h))


def setCheckerboard(full, half):
    _chk(lib().gb_set_checkerboard(full.h, half.h))


def precisionChange(out, inp):
    _chk(lib().gb_precision_change(out.h, inp.h))


def norm2(x):
    v = C.c_double()
    _chk(lib().gb_norm2(x.h, C.byref(v)))
    return v.value


def innerProduct(l, r):
    v = (C.c_double * 2)()
    _chk(lib().gb_inner_product(l.h, r.h, v))
    return complex(v[0], v[1])


def axpy(z, a, x, y):
    _chk(lib().gb_axpy(z.h, a, x.h, y.h))


def axpby(z, a, b, x, y):
    _chk(lib().gb_axpby(z.h, a, b, x.h, y.h))


def axpy_norm(z, a, x, y):
    v = C.c_double()
    _chk(lib().gb_axpy_norm(z.h, a, x.h, y.h, C.byref(v)))
    return v.value


def scale(z, a, x):
    _chk(lib().gb_scale(z.h, a, x.h))


def copy(z, x):
    _chk(lib().gb_copy(z.h, x.h))


class LatticeGaugeField:
    def __init__(self, grid, prec=F32):
        self.grid, self.prec = grid, prec
        self.h = C.c_void_p()
        _chk(lib().gb_gauge_create(grid.h, prec, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                lib().gb_gauge_destroy(self.h)
        except Exception:
            pass

    def import_lex(self, host):
        host = np.ascontiguousarray(host)
        assert host.shape == (self.grid.lsites, 4, 3, 3), host.shape
        _chk(lib().gb_gauge_import(self.h, host.ctypes.data_as(C.c_void_p), _prec_of(host)))
        return self

    def export_lex(self, dtype=np.complex128):
        out = np.empty((self.grid.lsites, 4, 3, 3), dtype=dtype)
        _chk(lib().gb_gauge_export(self.h, out.ctypes.data_as(C.c_void_p), _prec_of(out)))
        return out

    def random(self, seed):
        """SU<Nc>::HotConfiguration analogue, generated on the device."""
        _chk(lib().gb_gauge_random(self.h, seed))
        return self

    def unit(self):
        _chk(lib().gb_gauge_unit(self.h))
        return self


class NerscHeader(C.Structure):
    """gb_nersc_header (include/gridb200.h): FieldMetaData of a NERSC configuration (ref: Grid/parallelIO/MetaData.h:60-97)"""
    _fields_ = [("dimension", C.c_int * 4), ("link_trace", C.c_double), ("plaquette", C.c_double), ("checksum", C.c_uint32),
                ("data_type", C.c_char * 64), ("floating_point", C.c_char * 32), ("ensemble_id", C.c_char * 64), ("ensemble_label", C.c_char * 64),
                ("sequence_number", C.c_int), ("data_start", C.c_int64), ("computed_link_trace", C.c_double), ("computed_plaquette", C.c_double),
                ("computed_checksum", C.c_uint32)]


class NerscIO:
    """ref: Grid/parallelIO/NerscIO.h:43-290.  readConfiguration / writeConfiguration on device gauge fields; the *_host variants
    work on a global lexicographic [V,4,3,3] complex128 array and need no GPU."""

    @staticmethod
    def readHeader(path):
        h = NerscHeader()
        _chk(lib().gb_nersc_read_host(str(path).encode(), None, C.byref(h)))
        return h

    @staticmethod
    def read_host(path):
        h = NerscIO.readHeader(path)
        U = np.empty((int(np.prod(list(h.dimension))), 4, 3, 3), dtype=np.complex128)
        _chk(lib().gb_nersc_read_host(str(path).encode(), U.ctypes.data_as(C.c_void_p), C.byref(h)))
        return U, h

    @staticmethod
    def write_host(path, U, dims, two_row=0, ens_label="DWF", ens_id="UKQCD", sequence_number=1):
        U = np.ascontiguousarray(U, dtype=np.complex128)
        assert U.shape == (int(np.prod(dims)), 4, 3, 3)
        _chk(lib().gb_nersc_write_host(str(path).encode(), U.ctypes.data_as(C.c_void_p), _i4(dims), two_row, ens_label.encode(), ens_id.encode(), sequence_number))

    @staticmethod
    def readConfiguration(Umu, path):
        h = NerscHeader()
        _chk(lib().gb_gauge_read_nersc(Umu.h, str(path).encode(), C.byref(h)))
        return h

    @staticmethod
    def writeConfiguration(Umu, path, two_row=0, ens_label="DWF", ens_id="UKQCD", sequence_number=1):
        _chk(lib().gb_gauge_write_nersc(Umu.h, str(path).encode(), two_row, ens_label.encode(), ens_id.encode(), sequence_number))


class FermionOperator:
    """Mirror of FermionOperator<Impl> (ref: FermionOperator.h:40-192): same method names, (in, out[, dag])."""

    _cayley = False     # 5D Cayley operators apply Meo5D in front of the directional hops (Mdir, MdirAll)

    def __init__(self):
        self.h = C.c_void_p()
        self.grid = None
        self.Ls = 1

    def __del__(self):
        try:
            if self.h:
                lib().gb_op_destroy(self.h)
        except Exception:
            pass

    def _apply(self, which, i, o, dag=0):
        _chk(lib().gb_op_apply(self.h, which, i.h, o.h, dag))

    def ImportGauge(self, Umu):
        _chk(lib().gb_op_import_gauge(self.h, Umu.h))

    def M(self, i, o): self._apply(OP_M, i, o)
    def Mdag(self, i, o): self._apply(OP_MDAG, i, o)
    def Meooe(self, i, o): self._apply(OP_MEOOE, i, o)
    def MeooeDag(self, i, o): self._apply(OP_MEOOE_DAG, i, o)
    def Mooee(self, i, o): self._apply(OP_MOOEE, i, o)
    def MooeeDag(self, i, o): self._apply(OP_MOOEE_DAG, i, o)
    def MooeeInv(self, i, o): self._apply(OP_MOOEE_INV, i, o)
    def MooeeInvDag(self, i, o): self._apply(OP_MOOEE_INV_DAG, i, o)
    def Dhop(self, i, o, dag=0): self._apply(OP_DHOP, i, o, dag)
    def DhopOE(self, i, o, dag=0): self._apply(OP_DHOP_OE, i, o, dag)
    def DhopEO(self, i, o, dag=0): self._apply(OP_DHOP_EO, i, o, dag)
    def DW(self, i, o, dag=0): self._apply(OP_DW, i, o, dag)
    def Meooe5D(self, i, o): self._apply(OP_MEOOE5D, i, o)
    def MeooeDag5D(self, i, o): self._apply(OP_MEOOEDAG5D, i, o)

    # physical 4D <-> 5D maps (ref: FermionOperator.h:172-191, CayleyFermion5DImplementation.h:58-153)
    def Dminus(self, i, o): self._apply(OP_DMINUS, i, o)
    def DminusDag(self, i, o): self._apply(OP_DMINUS_DAG, i, o)
    def ImportPhysicalFermionSource(self, input4d, imported5d): _chk(lib().gb_op_import_physical_fermion_source(self.h, input4d.h, imported5d.h))
    def ImportUnphysicalFermion(self, input4d, imported5d): _chk(lib().gb_op_import_unphysical_fermion(self.h, input4d.h, imported5d.h))
    def ExportPhysicalFermionSolution(self, solution5d, exported4d): _chk(lib().gb_op_export_physical_fermion_solution(self.h, solution5d.h, exported4d.h))
    def ExportPhysicalFermionSource(self, source5d, exported4d): _chk(lib().gb_op_export_physical_fermion_source(self.h, source5d.h, exported4d.h))

    # single hop legs and force terms (ref: FermionOperator.h:79-93; dir = 0..3, disp = +-1; mat = LatticeGaugeField)
    def DhopDir(self, i, o, dir, disp): _chk(lib().gb_op_dhop_dir(self.h, i.h, o.h, dir, disp))
    def DhopDirAll(self, i, outs):
        """outs[0..3] = forward legs x,y,z,t; outs[4..7] = backward legs (ref: WilsonKernelsImplementation.h:343-372)"""
        assert len(outs) == 8
        for p, o in enumerate(outs):
            self.DhopDir(i, o, p & 3, 1 if p < 4 else -1)

    def Mdir(self, i, o, dir, disp):
        """ref: CayleyFermion5DImplementation.h:331-337 (Meo5D then DhopDir); WilsonFermionImplementation.h:345-348 (= DhopDir)"""
        if self._cayley:
            tmp = i.like()
            self.Meooe5D(i, tmp)
            self.DhopDir(tmp, o, dir, disp)
        else:
            self.DhopDir(i, o, dir, disp)

    def MdirAll(self, i, outs):
        if self._cayley:
            tmp = i.like()
            self.Meooe5D(i, tmp)
            self.DhopDirAll(tmp, outs)
        else:
            self.DhopDirAll(i, outs)

    def DhopDeriv(self, mat, U, V, dag): _chk(lib().gb_op_dhop_deriv(self.h, mat.h, U.h, V.h, dag))
    def MDeriv(self, mat, U, V, dag): _chk(lib().gb_op_mderiv(self.h, mat.h, U.h, V.h, dag))
    def MeoDeriv(self, mat, U, V, dag): assert U.Checkerboard() == Even; _chk(lib().gb_op_meooe_deriv(self.h, mat.h, U.h, V.h, dag))
    def MoeDeriv(self, mat, U, V, dag): assert U.Checkerboard() == Odd; _chk(lib().gb_op_meooe_deriv(self.h, mat.h, U.h, V.h, dag))

    def Dhop_host(self, host_in, host_out, dag=0):
        """Dhop on host-resident full-lattice arrays [V4*Ls,4,3] (lexicographic); pipelined H2D / hop / D2H (on z / t decomposed lattices: faces first, one halo exchange, then the slices stream)."""
        assert host_in.flags.c_contiguous and host_out.flags.c_contiguous and host_in.dtype == host_out.dtype and host_in.shape == host_out.shape
        _chk(lib().gb_op_dhop_host(self.h, host_in.ctypes.data_as(C.c_void_p), host_out.ctypes.data_as(C.c_void_p), _prec_of(host_in), dag))
        return host_out

    def halo_exchange(self, i, dag=0):
        """The face exchange of one full-lattice Dhop without the hop (halo microbenchmark); returns the bytes this rank sent."""
        b = C.c_int64()
        _chk(lib().gb_op_halo_exchange(self.h, i.h, dag, C.byref(b)))
        return b.value

    def set_tiling(self, by=0, bz=0, bt=0):
        lib().gb_op_set_tiling(self.h, by, bz, bt)

    def set_fast_kernel(self, on):
        """True/1: default kernel selection; 2: micro-block kernel instead of the column-sweep kernel; False/0: generic kernel"""
        lib().gb_op_set_fast_kernel(self.h, int(on))

    def set_link_reconstruct(self, nreal):
        """12: two rows per link, third row rebuilt in registers (needs special unitary links); 18: full store (default)"""
        _chk(lib().gb_op_set_link_reconstruct(self.h, int(nreal)))
        return self

    def set_halo_compression(self, on):
        """halos one precision down (fp32 operator: bf16, fp64 operator: fp32), the reference's ...FH / ...DF comms
        (FermionOperatorImpl.h:96-137 CoeffRealHalfComms, WilsonCompressor.h:244-306); arithmetic stays in the operator's precision"""
        _chk(lib().gb_op_set_halo_compression(self.h, int(on)))
        return self

    def set_overlap(self, on):
        """True/1: overlapped (semi-fused where it applies); 2: overlapped as interior + accumulate-exterior; False/0: serial"""
        lib().gb_op_set_overlap(self.h, int(on))


def _phases(ph):
    if ph is None:
        return None
    a = np.ascontiguousarray(np.asarray(ph, dtype=np.complex128)).view(np.float64)
    return (C.c_double * 8)(*a)


class WilsonFermion(FermionOperator):
    """ref: WilsonFermion.h:139-142"""

    def __init__(self, Umu, grid, mass, boundary_phases=None):
        super().__init__()
        self.grid = grid
        _chk(lib().gb_op_create_wilson(grid.h, Umu.h, mass, _phases(boundary_phases), C.byref(self.h)))


class DomainWallFermion(FermionOperator):
    """ref: DomainWallFermion.h:108-134 (Shamir: b=1, c=0)"""
    _cayley = True

    def __init__(self, Umu, grid, Ls, mass, M5, boundary_phases=None):
        super().__init__()
        self.grid, self.Ls = grid, Ls
        _chk(lib().gb_op_create_dwf(grid.h, Umu.h, Ls, mass, M5, _phases(boundary_phases), C.byref(self.h)))


class MobiusFermion(FermionOperator):
    """ref: MobiusFermion.h:45-71"""
    _cayley = True

    def __init__(self, Umu, grid, Ls, mass, M5, b, c, boundary_phases=None):
        super().__init__()
        self.grid, self.Ls = grid, Ls
        _chk(lib().gb_op_create_mobius(grid.h, Umu.h, Ls, mass, M5, b, c, _phases(boundary_phases), C.byref(self.h)))


class ImprovedStaggeredFermion(FermionOperator):
    """ref: ImprovedStaggeredFermion.h:115-121 -- (Uthin, Ufat, grid, mass, c1, c2, u0); fields are LatticeStaggeredFermion."""

    def __init__(self, Uthin, Ufat, grid, mass, c1=9.0 / 8.0, c2=-1.0 / 24.0, u0=1.0):
        super().__init__()
        self.grid, self.mass = grid, mass
        _chk(lib().gb_op_create_staggered(grid.h, Uthin.h, Ufat.h, mass, c1, c2, u0, C.byref(self.h)))

    def ImportGauge(self, Uthin, Ufat=None):
        _chk(lib().gb_op_import_gauge_staggered(self.h, Uthin.h, (Ufat or Uthin).h))

    def Mass(self):
        return self.mass


class LinearOperatorBase:
    """ref: Grid/algorithms/LinearOperator.h:44-56"""

    def Op(self, i, o): raise NotImplementedError
    def AdjOp(self, i, o): raise NotImplementedError
    def HermOp(self, i, o): raise NotImplementedError

    def HermOpAndNorm(self, i, o):
        self.HermOp(i, o)
        return innerProduct(i, o).real, norm2(o)


class MdagMLinearOperator(LinearOperatorBase):
    """ref: LinearOperator.h:74-105 -- the unpreconditioned normal operator on the full grid: HermOp = Mdag M.  ConjugateGradient
    drives it through the generic (virtual HermOp) path."""

    def __init__(self, Mat):
        self._Mat = Mat

    def Op(self, i, o): self._Mat.M(i, o)
    def AdjOp(self, i, o): self._Mat.Mdag(i, o)

    def HermOp(self, i, o):
        tmp = i.like()
        self._Mat.M(i, tmp)
        self._Mat.Mdag(tmp, o)


class SchurDiagMooeeOperator(LinearOperatorBase):
    """ref: LinearOperator.h:325-349.  Mpc = Mooee - Meooe MooeeInv Meooe ; HermOp = MpcDag Mpc."""

    def __init__(self, Mat):
        self._Mat = Mat

    def Mpc(self, i, o): self._Mat._apply(OP_MPC, i, o)
    def MpcDag(self, i, o): self._Mat._apply(OP_MPC_DAG, i, o)
    def MpcDagMpc(self, i, o): self._Mat._apply(OP_HERMOP, i, o)
    def Op(self, i, o): self.Mpc(i, o)
    def AdjOp(self, i, o): self.MpcDag(i, o)
    def HermOp(self, i, o): self.MpcDagMpc(i, o)


class SchurDifferentiableOperator(SchurDiagMooeeOperator):
    """ref: Grid/qcd/action/pseudofermion/EvenOddSchurDifferentiable.h:43-139 -- the Schur operator with its force terms."""

    def MpcDeriv(self, Force, U, V): _chk(lib().gb_op_mpc_deriv(self._Mat.h, Force.h, U.h, V.h, 0))
    def MpcDagDeriv(self, Force, U, V): _chk(lib().gb_op_mpc_deriv(self._Mat.h, Force.h, U.h, V.h, 1))


class SchurStaggeredOperator(SchurDiagMooeeOperator):
    """ref: LinearOperator.h:543-584.  Mpc = MpcDag = HermOp = mass^2 - Meooe Meooe (Hermitian: one application per CG step)."""

    def MpcDagMpc(self, i, o):
        raise AssertionError("never needed with staggered (ref: LinearOperator.h:581-583)")

    def HermOp(self, i, o): self.Mpc(i, o)


class ConjugateGradient:
    """ref: ConjugateGradient.h:42-258.  __call__(Linop, src, psi); psi is the initial guess on entry."""

    def __init__(self, tol, maxit, err_on_no_conv=True):
        self.Tolerance, self.MaxIterations, self.ErrorOnNoConverge = tol, maxit, err_on_no_conv
        self.IterationsToComplete, self.TrueResidual = 0, 0.0

    def __call__(self, Linop, src, psi):
        it, tr = C.c_int(), C.c_double()
        if isinstance(Linop, SchurDiagMooeeOperator):
            rc = lib().gb_cg_schur(Linop._Mat.h, src.h, psi.h, self.Tolerance, self.MaxIterations, C.byref(it), C.byref(tr))
        else:  # any user-written LinearOperatorBase: CG drives its HermOp through the callback path
            raised = []

            def cb(_user, hin, hout):
                # ctypes swallows exceptions raised inside a callback (and returns 0 = GB_OK): stash ANY exception, stop the
                # solver with an error code, and re-raise once gb_cg has returned
                try:
                    fin, fout = _Borrowed(hin, src), _Borrowed(hout, src)
                    Linop.HermOp(fin, fout)
                    return GB_OK
                except GridB200Error as e:
                    raised.append(e)
                    return e.code
                except BaseException as e:  # noqa: BLE001
                    raised.append(e)
                    return GB_ERR_INVALID

            fn = HERMOP_FN(cb)
            rc = lib().gb_cg(src.grid.ctx.h, fn, None, src.h, psi.h, self.Tolerance, self.MaxIterations, C.byref(it), C.byref(tr))
            if raised:
                raise raised[0]
        self.IterationsToComplete, self.TrueResidual = it.value, tr.value
        if rc == GB_ERR_NOT_CONVERGED:
            assert not self.ErrorOnNoConverge, "ConjugateGradient did NOT converge"  # ref: ConjugateGradient.h:254
            return
        _chk(rc)
        if self.ErrorOnNoConverge:
            assert self.TrueResidual / self.Tolerance < 10000.0  # ref: ConjugateGradient.h:225


class _Borrowed(LatticeFermion):
    """Non-owning view of a library-owned field handle (used inside the generic-CG callback)."""

    def __init__(self, handle, like):
        self.grid, self.Ls, self.prec, self.kind, self.SITE = like.grid, like.Ls, like.prec, like.kind, like.SITE
        self._cls = getattr(like, "_cls", type(like))
        self.h = C.c_void_p(handle)

    def __del__(self):
        pass


class MixedPrecisionConjugateGradient:
    """ref: ConjugateGradientMixedPrec.h:34-170."""

    def __init__(self, tol, maxinnerit, maxouterit, Linop_f, Linop_d):
        self.Tolerance, self.MaxInnerIterations, self.MaxOuterIterations = tol, maxinnerit, maxouterit
        self.Linop_f, self.Linop_d = Linop_f, Linop_d
        self.InnerTolerance, self.OuterLoopNormMult = tol, 100.0      # public tuning members, ref :41-45,:64-65
        self.TotalInnerIterations = self.TotalOuterIterations = self.TotalFinalStepIterations = 0
        self.TrueResidual = 0.0

    def __call__(self, src_d, sol_d):
        it, tr = (C.c_int * 3)(), C.c_double()
        rc = lib().gb_mixed_cg_schur_ex(self.Linop_f._Mat.h, self.Linop_d._Mat.h, src_d.h, sol_d.h, self.Tolerance, self.InnerTolerance,
                                        self.OuterLoopNormMult, self.MaxInnerIterations, self.MaxOuterIterations, it, C.byref(tr))
        self.TotalInnerIterations, self.TotalOuterIterations, self.TotalFinalStepIterations = it[0], it[1], it[2]
        self.TrueResidual = tr.value
        _chk(rc)


class SchurRedBlackDiagMooeeSolve:
    """ref: Grid/algorithms/iterative/SchurRedBlack.h:96-290,385-430.  SchurSolver = SchurRedBlackDiagMooeeSolve(CG);
    SchurSolver(Ddwf, src, result) solves M result = src on the full lattice (tests/solver/Test_dwf_cg_schur.cc:46-50).
    HermitianRBSolver is a ConjugateGradient (fused device path), a MixedPrecisionSchurCG (below) or any callable
    solver(LinearOperator, src_o, sol_o)."""
    _Operator = SchurDiagMooeeOperator

    def __init__(self, HermitianRBSolver, initSubGuess=False, solnAsInitGuess=False):
        self._HermitianRBSolver, self.subGuess, self.useSolnAsInitGuess = HermitianRBSolver, initSubGuess, solnAsInitGuess
        self.TrueUnprecResidual = None

    def subtractGuess(self, initSubGuess): self.subGuess = initSubGuess
    def isSubtractGuess(self): return self.subGuess

    def RedBlackSource(self, Matrix, src, src_e, src_o):
        _chk(lib().gb_schur_redblack_source(Matrix.h, src.h, src_e.h, src_o.h))

    def RedBlackSolution(self, Matrix, sol_o, src_e, sol):
        _chk(lib().gb_schur_redblack_solution(Matrix.h, sol_o.h, src_e.h, sol.h))

    def RedBlackSolve(self, Matrix, src_o, sol_o):
        self._HermitianRBSolver(self._Operator(Matrix), src_o, sol_o)
        assert sol_o.Checkerboard() == Odd

    def __call__(self, Matrix, src, out):
        S = self._HermitianRBSolver
        if isinstance(S, ConjugateGradient) and not self.subGuess:   # one call: source, device-resident CG, reconstruction
            it, rs = C.c_int(), (C.c_double * 2)()
            rc = lib().gb_schur_solve(Matrix.h, src.h, out.h, S.Tolerance, S.MaxIterations, int(self.useSolnAsInitGuess), C.byref(it), rs)
            S.IterationsToComplete, S.TrueResidual, self.TrueUnprecResidual = it.value, rs[0], rs[1]
            if rc == GB_ERR_NOT_CONVERGED:
                assert not S.ErrorOnNoConverge, "ConjugateGradient did NOT converge"
                return
            _chk(rc)
            return
        src_e, src_o, sol_o = src.like(kind=HALF), src.like(kind=HALF), src.like(kind=HALF)
        self.RedBlackSource(Matrix, src, src_e, src_o)
        if self.useSolnAsInitGuess:
            pickCheckerboard(Odd, sol_o, out)
        else:
            sol_o.zero().set_checkerboard(Odd)       # ZeroGuesser
        guess_save = None
        if self.subGuess:
            guess_save = src.like(kind=HALF)
            copy(guess_save, sol_o)
        self.RedBlackSolve(Matrix, src_o, sol_o)
        if self.subGuess:
            axpy(sol_o, -1.0, guess_save, sol_o)
        self.RedBlackSolution(Matrix, sol_o, src_e, out)
        if not self.subGuess:                       # "true unprec resid", ref: SchurRedBlack.h:277-285
            resid = src.like()
            Matrix.M(out, resid)
            axpy(resid, -1.0, src, resid)
            self.TrueUnprecResidual = (norm2(resid) / norm2(src)) ** 0.5


class SchurRedBlackStaggeredSolve(SchurRedBlackDiagMooeeSolve):
    """ref: SchurRedBlack.h:294-349 (source carries Mooee = mass instead of MpcDag; the C entry points dispatch on the operator)."""
    _Operator = SchurStaggeredOperator


SchurRedBlackStagSolve = SchurRedBlackStaggeredSolve


def schur_solve_mixed(Mat_f, Mat_d, src, out, tol, maxinnerit, maxouterit):
    """Full-lattice M out = src with MixedPrecisionConjugateGradient as the red-black solver (fp64 outside, fp32 inside).
    Returns dict(inner, outer, final, true_residual, unprec_residual)."""
    it, rs = (C.c_int * 3)(), (C.c_double * 2)()
    _chk(lib().gb_schur_solve_mixed(Mat_f.h, Mat_d.h, src.h, out.h, tol, maxinnerit, maxouterit, it, rs))
    return dict(inner=it[0], outer=it[1], final=it[2], true_residual=rs[0], unprec_residual=rs[1])


class MultiShiftFunction:
    """ref: Grid/algorithms/approx/MultiShiftFunction.h:34-44 -- order, poles, residues, tolerances, norm (no Remez here: the
    caller supplies the partial-fraction coefficients)."""

    def __init__(self, poles, tolerances, residues=None, norm=0.0):
        self.poles = [float(x) for x in poles]
        self.order = len(self.poles)
        self.tolerances = [float(tolerances)] * self.order if np.isscalar(tolerances) else [float(x) for x in tolerances]
        self.residues = [1.0] * self.order if residues is None else [float(x) for x in residues]
        self.norm = float(norm)
        assert len(self.tolerances) == self.order and len(self.residues) == self.order

    def approx(self, x):
        return self.norm + sum(r / (x + p) for r, p in zip(self.residues, self.poles))


class ConjugateGradientMultiShift:
    """ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:40-343.  MSCG = ConjugateGradientMultiShift(maxit, shifts);
    MSCG(HermOpEO, src_o, results) fills results[s] = (HermOp + poles[s])^-1 src_o; with a fourth argument psi it also forms
    psi = norm * src + sum_s residues[s] * results[s] (:69-82)."""

    def __init__(self, maxit, shifts):
        self.MaxIterations, self.shifts = maxit, shifts
        self.IterationsToComplete = 0
        self.IterationsToCompleteShift = [0] * shifts.order
        self.TrueResidualShift = [0.0] * shifts.order

    def __call__(self, Linop, src, results, psi=None):
        assert isinstance(Linop, SchurDiagMooeeOperator), "ConjugateGradientMultiShift runs on SchurDiagMooeeOperator / SchurStaggeredOperator"
        n = self.shifts.order
        assert len(results) == n
        poles, tols = (C.c_double * n)(*self.shifts.poles), (C.c_double * n)(*self.shifts.tolerances)
        handles = (C.c_void_p * n)(*[r.h.value for r in results])
        it, tr = (C.c_int * (n + 1))(), (C.c_double * n)()
        rc = lib().gb_cg_multishift_schur(Linop._Mat.h, src.h, n, poles, tols, self.MaxIterations, handles, it, tr)
        self.IterationsToCompleteShift, self.TrueResidualShift, self.IterationsToComplete = list(it[:n]), list(tr), it[n]
        if rc != GB_ERR_NOT_CONVERGED:     # the reference only logs "CG multi shift did not converge" (:336-338)
            _chk(rc)
        if psi is not None:
            scale(psi, self.shifts.norm, src)
            for res, r in zip(self.shifts.residues, results):
                axpy(psi, res, r, psi)
        return rc == GB_OK


class ConjugateGradientReliableUpdate:
    """ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:36-270.  mCG = ConjugateGradientReliableUpdate(tol, maxit,
    delta, Linop_f, Linop_d); mCG(src_d, psi_d)."""

    def __init__(self, tol, maxit, delta, Linop_f, Linop_d, err_on_no_conv=True):
        self.Tolerance, self.MaxIterations, self.Delta, self.ErrorOnNoConverge = tol, maxit, delta, err_on_no_conv
        self.Linop_f, self.Linop_d = Linop_f, Linop_d
        self.IterationsToComplete = self.ReliableUpdatesPerformed = self.IterationsToCleanup = 0
        self.TrueResidual = 0.0

    def __call__(self, src_d, psi_d):
        it, tr = (C.c_int * 3)(), C.c_double()
        rc = lib().gb_relup_cg_schur(self.Linop_f._Mat.h, self.Linop_d._Mat.h, src_d.h, psi_d.h, self.Tolerance, self.MaxIterations,
                                     self.Delta, it, C.byref(tr))
        self.IterationsToComplete, self.ReliableUpdatesPerformed, self.IterationsToCleanup, self.TrueResidual = it[0], it[1], it[2], tr.value
        if rc == GB_ERR_NOT_CONVERGED:
            assert not self.ErrorOnNoConverge, "ConjugateGradientReliableUpdate did NOT converge"
            return
        _chk(rc)


class ConjugateGradientMultiShiftMixedPrec(ConjugateGradientMultiShift):
    """ref: Grid/algorithms/iterative/ConjugateGradientMultiShiftMixedPrec.h:73-410.  mcg = ConjugateGradientMultiShiftMixedPrec(maxit,
    shifts, Linop_f, ReliableUpdateFreq); mcg(Linop_d, src_d, results_d[, psi_d])."""

    def __init__(self, maxit, shifts, Linop_f, ReliableUpdateFreq):
        super().__init__(maxit, shifts)
        self.Linop_f, self.ReliableUpdateFreq = Linop_f, ReliableUpdateFreq

    def __call__(self, Linop_d, src, results, psi=None):
        n = self.shifts.order
        assert len(results) == n
        poles, tols = (C.c_double * n)(*self.shifts.poles), (C.c_double * n)(*self.shifts.tolerances)
        handles = (C.c_void_p * n)(*[r.h.value for r in results])
        it, tr = (C.c_int * (n + 1))(), (C.c_double * n)()
        _chk(lib().gb_cg_multishift_mixed_schur(self.Linop_f._Mat.h, Linop_d._Mat.h, src.h, n, poles, tols, self.MaxIterations,
                                                self.ReliableUpdateFreq, handles, it, tr))
        self.IterationsToCompleteShift, self.TrueResidualShift, self.IterationsToComplete = list(it[:n]), list(tr), it[n]
        if psi is not None:
            scale(psi, self.shifts.norm, src)
            for res, r in zip(self.shifts.residues, results):
                axpy(psi, res, r, psi)
        return True


class MixedPrecisionConjugateGradientBatched:
    """ref: Grid/algorithms/iterative/ConjugateGradientMixedPrecBatched.h:36-213 -- the defect-correction loop of
    MixedPrecisionConjugateGradient over a batch of right-hand sides with ONE inner tolerance schedule (the largest residual of
    the batch sets it), then a double-precision patch-up CG per right-hand side: gb_mixed_cg_batched_schur.  The reference keeps its
    counts in locals and only logs them; here they are members."""

    def __init__(self, tol, maxinnerit, maxouterit, maxpatchit, Linop_f, Linop_d, updateResidual=True):
        self.Tolerance, self.InnerTolerance = tol, tol
        self.MaxInnerIterations, self.MaxOuterIterations, self.MaxPatchupIterations = maxinnerit, maxouterit, maxpatchit
        self.Linop_f, self.Linop_d, self.updateResidual, self.OuterLoopNormMult = Linop_f, Linop_d, updateResidual, 100.0
        self.TotalOuterIterations, self.TotalInnerIterations, self.TotalFinalStepIterations = 0, [], []

    def __call__(self, srcs_d, sols_d):
        if not isinstance(srcs_d, (list, tuple)):
            srcs_d, sols_d = [srcs_d], [sols_d]
        assert len(srcs_d) == len(sols_d)
        nb = len(srcs_d)
        S = (C.c_void_p * nb)(*[f.h for f in srcs_d])
        X = (C.c_void_p * nb)(*[f.h for f in sols_d])
        it, tr = (C.c_int * (1 + 2 * nb))(), (C.c_double * nb)()
        rc = lib().gb_mixed_cg_batched_schur(self.Linop_f._Mat.h, self.Linop_d._Mat.h, nb, S, X, self.Tolerance, self.MaxInnerIterations,
                                             self.MaxOuterIterations, self.MaxPatchupIterations, 1 if self.updateResidual else 0, it, tr)
        self.TotalOuterIterations = it[0]
        self.TotalInnerIterations, self.TotalFinalStepIterations = list(it[1:1 + nb]), list(it[1 + nb:])
        self.TrueResidual = list(tr)
        _chk(rc)
