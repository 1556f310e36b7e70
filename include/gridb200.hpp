// gridb200.hpp -- header-only C++ mirror of the reference's operator / solver interface over the C ABI.
//
// Class and method names, argument order and error behaviour (assert) follow paboyle/Grid so that drivers shaped
// like benchmarks/Benchmark_dwf_fp32.cc and tests/Test_dwf_mixedcg_prec.cc compile against this header with only
// the include and the namespace changed.  Everything here is a thin veneer: all arithmetic happens in
// libgridb200.so (hand-written sm_100a CUDA).  "ref:" paths are relative to the reference tree.
#pragma once
#include "gridb200.h"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

namespace gridb200 {

typedef double RealD;
typedef std::complex<double> ComplexD;
typedef int Integer;
enum { Even = 0, Odd = 1 };
enum { DaggerNo = 0, DaggerYes = 1 };
typedef std::vector<int> Coordinate;

#define GB_ASSERT_OK(expr)                                                              \
  do {                                                                                  \
    int _rc = (expr);                                                                   \
    if (_rc != GB_OK) { std::fprintf(stderr, "gridb200: %s -> %s\n", #expr, gb_last_error()); assert(_rc == GB_OK); std::abort(); } \
  } while (0)

// ---- Grid_init analogue (ref: Grid/util/Init.cc:300-560): one context per process / GPU
class Runtime {
public:
  static gb_context *&ctx() { static gb_context *c = nullptr; return c; }
  static void init(int device = 0) { if (!ctx()) GB_ASSERT_OK(gb_context_create(device, &ctx())); }
  static void finalize() { if (ctx()) { gb_context_destroy(ctx()); ctx() = nullptr; } }
};
inline void Grid_init(int * /*argc*/, char *** /*argv*/, int device = 0) { Runtime::init(device); }
inline void Grid_finalize() { Runtime::finalize(); }

// ---- grids (ref: Grid/cartesian/Cartesian_base.h, Cartesian_red_black.h, qcd/utils/SpaceTimeGrid.cc:36-78)
class GridBase {
public:
  gb_grid *h = nullptr;
  int Ls = 1;          // 1 for four-dimensional grids
  bool redblack = false;
  std::shared_ptr<gb_grid> owner;
  Coordinate dims, mpi;
  int64_t gSites() const { int64_t v = Ls; for (int d : dims) v *= d; return redblack ? v / 2 : v; }
};
typedef GridBase GridCartesian;
typedef GridBase GridRedBlackCartesian;
struct SpaceTimeGrid {
  static GridCartesian *makeFourDimGrid(const Coordinate &latt, const Coordinate & /*simd*/, const Coordinate &mpi) {
    Runtime::init();
    GridBase *g = new GridBase();
    gb_grid *raw = nullptr;
    GB_ASSERT_OK(gb_grid_create(Runtime::ctx(), latt.data(), mpi.data(), &raw));
    g->owner = std::shared_ptr<gb_grid>(raw, [](gb_grid *p) { gb_grid_destroy(p); });
    g->h = raw; g->dims = latt; g->mpi = mpi;
    return g;
  }
  static GridRedBlackCartesian *makeFourDimRedBlackGrid(const GridCartesian *g4) { GridBase *g = new GridBase(*g4); g->redblack = true; return g; }
  static GridCartesian *makeFiveDimGrid(int Ls, const GridCartesian *g4) { GridBase *g = new GridBase(*g4); g->Ls = Ls; return g; }
  static GridRedBlackCartesian *makeFiveDimRedBlackGrid(int Ls, const GridCartesian *g4) { GridBase *g = new GridBase(*g4); g->Ls = Ls; g->redblack = true; return g; }
};

// ---- lattice containers (ref: Grid/lattice/Lattice_base.h); Prec = GB_F32 / GB_F64
template <gb_precision Prec> class LatticeFermionT {
public:
  gb_fermion *h = nullptr;
  GridBase *_grid;
  bool _staggered = false;   // true: one ColourVector per site (LatticeStaggeredFermion), false: SpinColourVector
  LatticeFermionT(GridBase *g) : _grid(g) { create(); }   // not explicit: std::vector<Field> v(n, grid) as in the reference's drivers
  LatticeFermionT(const LatticeFermionT &o) : _grid(o._grid), _staggered(o._staggered) { create(); GB_ASSERT_OK(gb_copy(h, o.h)); }
  LatticeFermionT &operator=(const LatticeFermionT &o) { GB_ASSERT_OK(gb_copy(h, o.h)); return *this; }
  ~LatticeFermionT() { gb_fermion_destroy(h); }
  GridBase *Grid() const { return _grid; }
  int Checkerboard() const { return gb_fermion_checkerboard(h); }
  void SetCheckerboard(int cb) { gb_fermion_set_checkerboard_tag(h, cb); } // Grid: f.Checkerboard() = cb
  void Zero() { GB_ASSERT_OK(gb_zero(h)); }
  // import/export of the local lattice in lexicographic order (ref: Lattice_transfer.h:1123,1218)
  void ImportLex(const void *host, gb_precision hp) { GB_ASSERT_OK(gb_fermion_import(h, host, hp)); }
  void ExportLex(void *host, gb_precision hp) const { GB_ASSERT_OK(gb_fermion_export(h, host, hp)); }
protected:
  LatticeFermionT(GridBase *g, bool staggered) : _grid(g), _staggered(staggered) { create(); }
  void create() {
    const gb_gridkind kind = _grid->redblack ? GB_HALF : GB_FULL;
    if (_staggered) GB_ASSERT_OK(gb_staggered_fermion_create(_grid->h, Prec, kind, &h));
    else GB_ASSERT_OK(gb_fermion_create(_grid->h, _grid->Ls, Prec, kind, &h));
  }
};
typedef LatticeFermionT<GB_F32> LatticeFermionF;
typedef LatticeFermionT<GB_F64> LatticeFermionD;
// LatticeStaggeredFermion{F,D} (ref: StaggeredImpl.h:60-75): every algebra / checkerboard / solver entry point above and
// below takes it through its base class
template <gb_precision Prec> class LatticeStaggeredFermionT : public LatticeFermionT<Prec> {
public:
  LatticeStaggeredFermionT(GridBase *g) : LatticeFermionT<Prec>(g, true) {}
};
typedef LatticeStaggeredFermionT<GB_F32> LatticeStaggeredFermionF;
typedef LatticeStaggeredFermionT<GB_F64> LatticeStaggeredFermionD;

template <gb_precision Prec> class LatticeGaugeFieldT {
public:
  gb_gauge *h = nullptr;
  GridBase *_grid;
  explicit LatticeGaugeFieldT(GridBase *g) : _grid(g) { GB_ASSERT_OK(gb_gauge_create(g->h, Prec, &h)); }
  ~LatticeGaugeFieldT() { gb_gauge_destroy(h); }
  LatticeGaugeFieldT(const LatticeGaugeFieldT &) = delete;
  void ImportLex(const void *host, gb_precision hp) { GB_ASSERT_OK(gb_gauge_import(h, host, hp)); }
  void ExportLex(void *host, gb_precision hp) const { GB_ASSERT_OK(gb_gauge_export(h, host, hp)); }
};
typedef LatticeGaugeFieldT<GB_F32> LatticeGaugeFieldF;
typedef LatticeGaugeFieldT<GB_F64> LatticeGaugeFieldD;

// ---- NerscIO (ref: Grid/parallelIO/NerscIO.h:43-290): checksum / plaquette / link-trace validated reader, IEEE64BIG writer
typedef gb_nersc_header FieldMetaData;
class NerscIO {
public:
  template <gb_precision P> static void readConfiguration(LatticeGaugeFieldT<P> &Umu, FieldMetaData &header, const std::string &file) {
    GB_ASSERT_OK(gb_gauge_read_nersc(Umu.h, file.c_str(), &header));
  }
  template <gb_precision P> static void writeConfiguration(LatticeGaugeFieldT<P> &Umu, const std::string &file, int two_row = 0, int /*bits32*/ = 0,
                                                           const std::string &ens_label = "DWF", const std::string &ens_id = "UKQCD", unsigned int sequence_number = 1) {
    GB_ASSERT_OK(gb_gauge_write_nersc(Umu.h, file.c_str(), two_row, ens_label.c_str(), ens_id.c_str(), (int)sequence_number));
  }
};

// ---- RNG facade: synthetic fields are generated on the device, keyed by global site (decomposition independent)
class GridParallelRNG {
public:
  uint64_t seed = 0;
  explicit GridParallelRNG(GridBase *) {}
  void SeedFixedIntegers(const std::vector<int> &s) { seed = 0; for (int v : s) seed = seed * 1000003ull + (uint64_t)v; }
};
template <gb_precision P> void random(GridParallelRNG &rng, LatticeFermionT<P> &f) { GB_ASSERT_OK(gb_fermion_random(f.h, rng.seed)); }
template <int Nc> struct SU {
  template <gb_precision P> static void HotConfiguration(GridParallelRNG &rng, LatticeGaugeFieldT<P> &U) { GB_ASSERT_OK(gb_gauge_random(U.h, rng.seed)); }
  template <gb_precision P> static void ColdConfiguration(LatticeGaugeFieldT<P> &U) { GB_ASSERT_OK(gb_gauge_unit(U.h)); }
};

// ---- lattice algebra (ref: Lattice_arith.h:231-258, Lattice_reduction.h:256-372, Lattice_transfer.h:50-86,1461-1492)
template <gb_precision P> RealD norm2(const LatticeFermionT<P> &x) { double v; GB_ASSERT_OK(gb_norm2(x.h, &v)); return v; }
template <gb_precision P> ComplexD innerProduct(const LatticeFermionT<P> &l, const LatticeFermionT<P> &r) { double v[2]; GB_ASSERT_OK(gb_inner_product(l.h, r.h, v)); return ComplexD(v[0], v[1]); }
template <gb_precision P> void axpy(LatticeFermionT<P> &z, RealD a, const LatticeFermionT<P> &x, const LatticeFermionT<P> &y) { GB_ASSERT_OK(gb_axpy(z.h, a, x.h, y.h)); }
template <gb_precision P> void axpby(LatticeFermionT<P> &z, RealD a, RealD b, const LatticeFermionT<P> &x, const LatticeFermionT<P> &y) { GB_ASSERT_OK(gb_axpby(z.h, a, b, x.h, y.h)); }
template <gb_precision P> RealD axpy_norm(LatticeFermionT<P> &z, RealD a, const LatticeFermionT<P> &x, const LatticeFermionT<P> &y) { double v; GB_ASSERT_OK(gb_axpy_norm(z.h, a, x.h, y.h, &v)); return v; }
template <gb_precision P> void pickCheckerboard(int cb, LatticeFermionT<P> &half, const LatticeFermionT<P> &full) { GB_ASSERT_OK(gb_pick_checkerboard(cb, half.h, full.h)); }
template <gb_precision P> void setCheckerboard(LatticeFermionT<P> &full, const LatticeFermionT<P> &half) { GB_ASSERT_OK(gb_set_checkerboard(full.h, half.h)); }
template <gb_precision PO, gb_precision PI> void precisionChange(LatticeFermionT<PO> &out, const LatticeFermionT<PI> &in) { GB_ASSERT_OK(gb_precision_change(out.h, in.h)); }
// gauge precision change goes through the host layout (setup only)
inline void precisionChange(LatticeGaugeFieldF &out, const LatticeGaugeFieldD &in) {
  std::vector<double> tmp((size_t)in._grid->gSites() / in._grid->Ls * 72);
  in.ExportLex(tmp.data(), GB_F64);
  out.ImportLex(tmp.data(), GB_F64);
}

// ---- FermionOperator (ref: Grid/qcd/action/fermion/FermionOperator.h:40-192)
template <gb_precision Prec> class FermionOperator {
public:
  typedef LatticeFermionT<Prec> FermionField;
  typedef LatticeGaugeFieldT<Prec> GaugeField;
  gb_fermop *h = nullptr;
  virtual ~FermionOperator() { gb_op_destroy(h); }
  void apply(int which, const FermionField &in, FermionField &out, int dag = 0) { GB_ASSERT_OK(gb_op_apply(h, which, in.h, out.h, dag)); }
  virtual void M(const FermionField &in, FermionField &out) { apply(GB_OP_M, in, out); }
  virtual void Mdag(const FermionField &in, FermionField &out) { apply(GB_OP_MDAG, in, out); }
  virtual void Meooe(const FermionField &in, FermionField &out) { apply(GB_OP_MEOOE, in, out); }
  virtual void MeooeDag(const FermionField &in, FermionField &out) { apply(GB_OP_MEOOE_DAG, in, out); }
  virtual void Mooee(const FermionField &in, FermionField &out) { apply(GB_OP_MOOEE, in, out); }
  virtual void MooeeDag(const FermionField &in, FermionField &out) { apply(GB_OP_MOOEE_DAG, in, out); }
  virtual void MooeeInv(const FermionField &in, FermionField &out) { apply(GB_OP_MOOEE_INV, in, out); }
  virtual void MooeeInvDag(const FermionField &in, FermionField &out) { apply(GB_OP_MOOEE_INV_DAG, in, out); }
  virtual void Dhop(const FermionField &in, FermionField &out, int dag) { apply(GB_OP_DHOP, in, out, dag); }
  virtual void DhopOE(const FermionField &in, FermionField &out, int dag) { apply(GB_OP_DHOP_OE, in, out, dag); }
  virtual void DhopEO(const FermionField &in, FermionField &out, int dag) { apply(GB_OP_DHOP_EO, in, out, dag); }
  virtual void ImportGauge(const GaugeField &U) { GB_ASSERT_OK(gb_op_import_gauge(h, U.h)); }
  // physical 4D <-> 5D maps (ref: FermionOperator.h:172-191 ; CayleyFermion5DImplementation.h:58-153)
  virtual void Dminus(const FermionField &psi, FermionField &chi) { apply(GB_OP_DMINUS, psi, chi); }
  virtual void DminusDag(const FermionField &psi, FermionField &chi) { apply(GB_OP_DMINUS_DAG, psi, chi); }
  virtual void ImportPhysicalFermionSource(const FermionField &input4d, FermionField &imported5d) { GB_ASSERT_OK(gb_op_import_physical_fermion_source(h, input4d.h, imported5d.h)); }
  virtual void ImportUnphysicalFermion(const FermionField &input4d, FermionField &imported5d) { GB_ASSERT_OK(gb_op_import_unphysical_fermion(h, input4d.h, imported5d.h)); }
  virtual void ExportPhysicalFermionSolution(const FermionField &solution5d, FermionField &exported4d) { GB_ASSERT_OK(gb_op_export_physical_fermion_solution(h, solution5d.h, exported4d.h)); }
  virtual void ExportPhysicalFermionSource(const FermionField &solution5d, FermionField &exported4d) { GB_ASSERT_OK(gb_op_export_physical_fermion_source(h, solution5d.h, exported4d.h)); }
  // single hop legs and force terms on full-grid fields (ref: FermionOperator.h:79-93): dir = 0..3, disp = +-1
  virtual void DhopDir(const FermionField &in, FermionField &out, int dir, int disp) { GB_ASSERT_OK(gb_op_dhop_dir(h, in.h, out.h, dir, disp)); }
  virtual void DhopDirAll(const FermionField &in, std::vector<FermionField> &out) {   // out[0..3] forward legs, out[4..7] backward (ref: WilsonKernelsImplementation.h:343-372)
    assert(out.size() == 8);
    for (int p = 0; p < 8; p++) DhopDir(in, out[p], p & 3, p < 4 ? 1 : -1);
  }
  virtual void Mdir(const FermionField &in, FermionField &out, int dir, int disp) { DhopDir(in, out, dir, disp); }     // ref: WilsonFermionImplementation.h:345-353
  virtual void MdirAll(const FermionField &in, std::vector<FermionField> &out) { DhopDirAll(in, out); }
  virtual void DhopDeriv(GaugeField &mat, const FermionField &U, const FermionField &V, int dag) { GB_ASSERT_OK(gb_op_dhop_deriv(h, mat.h, U.h, V.h, dag)); }
  virtual void MDeriv(GaugeField &mat, const FermionField &U, const FermionField &V, int dag) { GB_ASSERT_OK(gb_op_mderiv(h, mat.h, U.h, V.h, dag)); }
  virtual void MeoDeriv(GaugeField &mat, const FermionField &U, const FermionField &V, int dag) { assert(U.Checkerboard() == Even); GB_ASSERT_OK(gb_op_meooe_deriv(h, mat.h, U.h, V.h, dag)); }
  virtual void MoeDeriv(GaugeField &mat, const FermionField &U, const FermionField &V, int dag) { assert(U.Checkerboard() == Odd); GB_ASSERT_OK(gb_op_meooe_deriv(h, mat.h, U.h, V.h, dag)); }
  // Dhop on host-resident full-lattice arrays in the reference's unvectorised layout (pipelined H2D / hop / D2H, also on z / t decomposed lattices)
  void DhopHost(const void *host_in, void *host_out, gb_precision host_prec, int dag) { GB_ASSERT_OK(gb_op_dhop_host(h, host_in, host_out, host_prec, dag)); }
  // halos one precision below the operator's (ref: CoeffRealHalfComms, FermionOperatorImpl.h:96-137); see the typedefs ...FH / ...DF below
  // 12: two-row link storage with the third row rebuilt in registers (special unitary links only); 18: the full store
  void SetLinkReconstruct(int nreal) { GB_ASSERT_OK(gb_op_set_link_reconstruct(h, nreal)); }
  void SetHaloCompression(bool on) { GB_ASSERT_OK(gb_op_set_halo_compression(h, on ? 1 : 0)); }
};
template <gb_precision Prec> class WilsonFermionT : public FermionOperator<Prec> {
public: // ref: WilsonFermion.h:139-142
  WilsonFermionT(LatticeGaugeFieldT<Prec> &Umu, GridCartesian &Fgrid, GridRedBlackCartesian &, RealD mass) {
    GB_ASSERT_OK(gb_op_create_wilson(Fgrid.h, Umu.h, mass, nullptr, &this->h));
  }
};
// CayleyFermion5D pieces shared by DomainWallFermion and MobiusFermion (ref: CayleyFermion5DImplementation.h:165-190,331-344)
template <gb_precision Prec> class CayleyFermion5DT : public FermionOperator<Prec> {
public:
  typedef LatticeFermionT<Prec> FermionField;
  void Meooe5D(const FermionField &in, FermionField &out) { this->apply(GB_OP_MEOOE5D, in, out); }
  void Meo5D(const FermionField &in, FermionField &out) { this->apply(GB_OP_MEOOE5D, in, out); }
  void MeooeDag5D(const FermionField &in, FermionField &out) { this->apply(GB_OP_MEOOEDAG5D, in, out); }
  void Mdir(const FermionField &psi, FermionField &chi, int dir, int disp) override { FermionField tmp(psi.Grid()); Meo5D(psi, tmp); this->DhopDir(tmp, chi, dir, disp); }
  void MdirAll(const FermionField &psi, std::vector<FermionField> &out) override { FermionField tmp(psi.Grid()); Meo5D(psi, tmp); this->DhopDirAll(tmp, out); }
};
template <gb_precision Prec> class DomainWallFermionT : public CayleyFermion5DT<Prec> {
public: // ref: DomainWallFermion.h:108-134
  DomainWallFermionT(LatticeGaugeFieldT<Prec> &Umu, GridCartesian &FGrid, GridRedBlackCartesian &, GridCartesian &UGrid, GridRedBlackCartesian &, RealD mass, RealD M5) {
    GB_ASSERT_OK(gb_op_create_dwf(UGrid.h, Umu.h, FGrid.Ls, mass, M5, nullptr, &this->h));
  }
};
template <gb_precision Prec> class MobiusFermionT : public CayleyFermion5DT<Prec> {
public: // ref: MobiusFermion.h:45-71
  MobiusFermionT(LatticeGaugeFieldT<Prec> &Umu, GridCartesian &FGrid, GridRedBlackCartesian &, GridCartesian &UGrid, GridRedBlackCartesian &, RealD mass, RealD M5, RealD b, RealD c) {
    GB_ASSERT_OK(gb_op_create_mobius(UGrid.h, Umu.h, FGrid.Ls, mass, M5, b, c, nullptr, &this->h));
  }
};
template <gb_precision Prec> class ImprovedStaggeredFermionT : public FermionOperator<Prec> {
public: // ref: ImprovedStaggeredFermion.h:115-121
  RealD mass;
  ImprovedStaggeredFermionT(LatticeGaugeFieldT<Prec> &Uthin, LatticeGaugeFieldT<Prec> &Ufat, GridCartesian &Fgrid, GridRedBlackCartesian &, RealD _mass,
                            RealD c1 = 9.0 / 8.0, RealD c2 = -1.0 / 24.0, RealD u0 = 1.0) : mass(_mass) {
    GB_ASSERT_OK(gb_op_create_staggered(Fgrid.h, Uthin.h, Ufat.h, _mass, c1, c2, u0, &this->h));
  }
  void ImportGauge(const LatticeGaugeFieldT<Prec> &Uthin, const LatticeGaugeFieldT<Prec> &Ufat) { GB_ASSERT_OK(gb_op_import_gauge_staggered(this->h, Uthin.h, Ufat.h)); }
  RealD Mass() const { return mass; }
  bool isTrivialEE() const { return true; }
};
typedef ImprovedStaggeredFermionT<GB_F32> ImprovedStaggeredFermionF; typedef ImprovedStaggeredFermionT<GB_F64> ImprovedStaggeredFermionD;
typedef WilsonFermionT<GB_F32> WilsonFermionF; typedef WilsonFermionT<GB_F64> WilsonFermionD;
typedef DomainWallFermionT<GB_F32> DomainWallFermionF; typedef DomainWallFermionT<GB_F64> DomainWallFermionD;
typedef MobiusFermionT<GB_F32> MobiusFermionF; typedef MobiusFermionT<GB_F64> MobiusFermionD;
// compressed-comms operators (ref: DomainWallFermionFH in tests/Test_dwf_mixedcg_prec_halfcomms.cc:71; impl typedefs ...ImplFH / ...ImplDF, DomainWallVec5dImpl.h:204-206):
// the same operator with SetHaloCompression(true) from construction
template <class Base> class CompressedComms : public Base {
public:
  template <class... A> CompressedComms(A &&...a) : Base(std::forward<A>(a)...) { this->SetHaloCompression(true); }
};
typedef CompressedComms<WilsonFermionF> WilsonFermionFH; typedef CompressedComms<WilsonFermionD> WilsonFermionDF;
typedef CompressedComms<DomainWallFermionF> DomainWallFermionFH; typedef CompressedComms<DomainWallFermionD> DomainWallFermionDF;
typedef CompressedComms<MobiusFermionF> MobiusFermionFH; typedef CompressedComms<MobiusFermionD> MobiusFermionDF;

// ---- linear operators (ref: Grid/algorithms/LinearOperator.h:44-56,286-349)
template <class Field> class LinearOperatorBase {
public:
  virtual ~LinearOperatorBase() {}
  virtual void Op(const Field &in, Field &out) = 0;
  virtual void AdjOp(const Field &in, Field &out) = 0;
  virtual void HermOp(const Field &in, Field &out) = 0;
  virtual void HermOpAndNorm(const Field &in, Field &out, RealD &n1, RealD &n2) {
    HermOp(in, out);
    n1 = innerProduct(in, out).real();
    n2 = norm2(out);
  }
  virtual gb_fermop *FusedSchurMatrix() { return nullptr; } // non-null: CG may use the fused device path
};
// MdagMLinearOperator (ref: LinearOperator.h:74-105): the unpreconditioned normal operator on the full grid, HermOp = Mdag M
template <class Matrix, class Field> class MdagMLinearOperator : public LinearOperatorBase<Field> {
public:
  Matrix &_Mat;
  explicit MdagMLinearOperator(Matrix &Mat) : _Mat(Mat) {}
  void Op(const Field &in, Field &out) override { _Mat.M(in, out); }
  void AdjOp(const Field &in, Field &out) override { _Mat.Mdag(in, out); }
  void HermOp(const Field &in, Field &out) override { Field tmp(in.Grid()); _Mat.M(in, tmp); _Mat.Mdag(tmp, out); }
};
template <class Matrix, class Field> class SchurDiagMooeeOperator : public LinearOperatorBase<Field> {
public:
  Matrix &_Mat;
  explicit SchurDiagMooeeOperator(Matrix &Mat) : _Mat(Mat) {}
  virtual void Mpc(const Field &in, Field &out) { _Mat.apply(GB_OP_MPC, in, out); }
  virtual void MpcDag(const Field &in, Field &out) { _Mat.apply(GB_OP_MPC_DAG, in, out); }
  virtual void MpcDagMpc(const Field &in, Field &out) { _Mat.apply(GB_OP_HERMOP, in, out); }
  void Op(const Field &in, Field &out) override { Mpc(in, out); }
  void AdjOp(const Field &in, Field &out) override { MpcDag(in, out); }
  void HermOp(const Field &in, Field &out) override { MpcDagMpc(in, out); }
  gb_fermop *FusedSchurMatrix() override { return _Mat.h; }
};

// SchurDifferentiableOperator (ref: Grid/qcd/action/pseudofermion/EvenOddSchurDifferentiable.h:43-139): Mpc with its force terms
template <class Matrix, class Field> class SchurDifferentiableOperator : public SchurDiagMooeeOperator<Matrix, Field> {
public:
  explicit SchurDifferentiableOperator(Matrix &Mat) : SchurDiagMooeeOperator<Matrix, Field>(Mat) {}
  template <class GaugeField> void MpcDeriv(GaugeField &Force, const Field &U, const Field &V) { GB_ASSERT_OK(gb_op_mpc_deriv(this->_Mat.h, Force.h, U.h, V.h, 0)); }
  template <class GaugeField> void MpcDagDeriv(GaugeField &Force, const Field &U, const Field &V) { GB_ASSERT_OK(gb_op_mpc_deriv(this->_Mat.h, Force.h, U.h, V.h, 1)); }
};

// SchurStaggeredOperator (ref: LinearOperator.h:543-584): Mpc = mass^2 - Meooe Meooe is Hermitian, HermOp = Mpc
template <class Matrix, class Field> class SchurStaggeredOperator : public SchurDiagMooeeOperator<Matrix, Field> {
public:
  explicit SchurStaggeredOperator(Matrix &Mat) : SchurDiagMooeeOperator<Matrix, Field>(Mat) { assert(Mat.isTrivialEE()); }
  void MpcDagMpc(const Field &, Field &) override { assert(0); } // never needed with staggered
  void HermOp(const Field &in, Field &out) override { this->Mpc(in, out); }
};

// ---- solvers (ref: Grid/algorithms/iterative/ConjugateGradient.h:42-258, ConjugateGradientMixedPrec.h:34-170)
// OperatorFunction (ref: Grid/algorithms/LinearOperator.h:590-600): what SchurRedBlack*Solve takes as its red-black solver
template <class Field> class OperatorFunction {
public:
  virtual ~OperatorFunction() {}
  virtual void operator()(LinearOperatorBase<Field> &Linop, const Field &in, Field &out) = 0;
  // non-null: SchurRedBlack*Solve may run source preparation, CG and reconstruction as one library call
  virtual struct FusedCG *Fused() { return nullptr; }
};
struct FusedCG { RealD *Tolerance; Integer *MaxIterations; Integer *IterationsToComplete; RealD *TrueResidual; bool *ErrorOnNoConverge; };
template <class Field> class ConjugateGradient : public OperatorFunction<Field> {
  FusedCG _fused{&Tolerance, &MaxIterations, &IterationsToComplete, &TrueResidual, &ErrorOnNoConverge};
public:
  FusedCG *Fused() override { return &_fused; }
  bool ErrorOnNoConverge;
  RealD Tolerance;
  Integer MaxIterations;
  Integer IterationsToComplete = 0;
  RealD TrueResidual = 0;
  ConjugateGradient(RealD tol, Integer maxit, bool err_on_no_conv = true) : ErrorOnNoConverge(err_on_no_conv), Tolerance(tol), MaxIterations(maxit) {}
  void operator()(LinearOperatorBase<Field> &Linop, const Field &src, Field &psi) override {
    int rc;
    if (gb_fermop *m = Linop.FusedSchurMatrix()) {
      rc = gb_cg_schur(m, src.h, psi.h, Tolerance, MaxIterations, &IterationsToComplete, &TrueResidual);
    } else { // user-written LinearOperatorBase: drive its virtual HermOp through the generic path
      Thunk t{&Linop, &src};
      rc = gb_cg(Runtime::ctx(), &Thunk::call, &t, src.h, psi.h, Tolerance, MaxIterations, &IterationsToComplete, &TrueResidual);
    }
    if (rc == GB_ERR_NOT_CONVERGED) { if (ErrorOnNoConverge) assert(0 && "ConjugateGradient did NOT converge"); return; }
    GB_ASSERT_OK(rc);
    if (ErrorOnNoConverge) assert(TrueResidual / Tolerance < 10000.0); // ref: ConjugateGradient.h:225
  }
private:
  struct Thunk {
    LinearOperatorBase<Field> *op; const Field *like;
    static int call(void *user, const gb_fermion *in, gb_fermion *out) {
      Thunk *t = (Thunk *)user;
      Borrow bi(t->like->Grid(), const_cast<gb_fermion *>(in)), bo(t->like->Grid(), out);
      t->op->HermOp(bi.f(), bo.f());
      return GB_OK;
    }
  };
  // wraps a library-owned handle in a Field without taking ownership
  struct Borrow {
    alignas(Field) unsigned char buf[sizeof(Field)];
    Borrow(GridBase *g, gb_fermion *hh) { Field *p = reinterpret_cast<Field *>(buf); p->h = hh; p->_grid = g; }
    Field &f() { return *reinterpret_cast<Field *>(buf); }
  };
};

template <class FieldD, class FieldF> class MixedPrecisionConjugateGradient {
public:
  RealD Tolerance, InnerTolerance;      // InnerTolerance: initial tolerance of the inner CG, defaults to Tolerance (ref :41)
  RealD OuterLoopNormMult = 100.;       // ref :45
  Integer MaxInnerIterations, MaxOuterIterations;
  GridBase *SinglePrecGrid;
  LinearOperatorBase<FieldF> &Linop_f;
  LinearOperatorBase<FieldD> &Linop_d;
  Integer TotalInnerIterations = 0, TotalOuterIterations = 0, TotalFinalStepIterations = 0;
  RealD TrueResidual = 0;
  MixedPrecisionConjugateGradient(RealD tol, Integer maxinnerit, Integer maxouterit, GridBase *sp_grid, LinearOperatorBase<FieldF> &lf, LinearOperatorBase<FieldD> &ld)
      : Tolerance(tol), InnerTolerance(tol), MaxInnerIterations(maxinnerit), MaxOuterIterations(maxouterit), SinglePrecGrid(sp_grid), Linop_f(lf), Linop_d(ld) {}
  void operator()(const FieldD &src, FieldD &sol) {
    gb_fermop *mf = Linop_f.FusedSchurMatrix(), *md = Linop_d.FusedSchurMatrix();
    assert(mf && md && "MixedPrecisionConjugateGradient needs SchurDiagMooeeOperator arguments");
    int it[3];
    int rc = gb_mixed_cg_schur_ex(mf, md, src.h, sol.h, Tolerance, InnerTolerance, OuterLoopNormMult, MaxInnerIterations, MaxOuterIterations, it, &TrueResidual);
    TotalInnerIterations = it[0]; TotalOuterIterations = it[1]; TotalFinalStepIterations = it[2];
    GB_ASSERT_OK(rc);
  }
};


// ---- MixedPrecisionConjugateGradientBatched (ref: Grid/algorithms/iterative/ConjugateGradientMixedPrecBatched.h:36-213)
// the defect-correction loop over a batch of right-hand sides with one inner-tolerance schedule, then a patch-up CG each: gb_mixed_cg_batched_schur
template <class FieldD, class FieldF> class MixedPrecisionConjugateGradientBatched {
public:
  RealD Tolerance, InnerTolerance;
  Integer MaxInnerIterations, MaxOuterIterations, MaxPatchupIterations;
  GridBase *SinglePrecGrid;
  RealD OuterLoopNormMult = 100.;
  LinearOperatorBase<FieldF> &Linop_f;
  LinearOperatorBase<FieldD> &Linop_d;
  bool updateResidual;
  MixedPrecisionConjugateGradientBatched(RealD tol, Integer maxinnerit, Integer maxouterit, Integer maxpatchit, GridBase *_sp_grid,
                                         LinearOperatorBase<FieldF> &_Linop_f, LinearOperatorBase<FieldD> &_Linop_d, bool _updateResidual = true)
      : Tolerance(tol), InnerTolerance(tol), MaxInnerIterations(maxinnerit), MaxOuterIterations(maxouterit), MaxPatchupIterations(maxpatchit),
        SinglePrecGrid(_sp_grid), Linop_f(_Linop_f), Linop_d(_Linop_d), updateResidual(_updateResidual) {}
  // counts the reference only logs (:201), kept here as members
  Integer TotalOuterIterations = 0;
  std::vector<Integer> TotalInnerIterations, TotalFinalStepIterations;
  std::vector<RealD> TrueResidual;
  void operator()(const std::vector<FieldD> &src_d_in, std::vector<FieldD> &sol_d) {
    assert(src_d_in.size() == sol_d.size());
    const int NBatch = (int)src_d_in.size();
    gb_fermop *mf = Linop_f.FusedSchurMatrix(), *md = Linop_d.FusedSchurMatrix();
    assert(mf && md && "MixedPrecisionConjugateGradientBatched needs SchurDiagMooeeOperator arguments");
    std::vector<const gb_fermion *> s(NBatch);
    std::vector<gb_fermion *> x(NBatch);
    for (int i = 0; i < NBatch; i++) { s[i] = src_d_in[i].h; x[i] = sol_d[i].h; }
    std::vector<int> it(1 + 2 * NBatch, 0);
    TrueResidual.assign(NBatch, 0.);
    int rc = gb_mixed_cg_batched_schur(mf, md, NBatch, s.data(), x.data(), Tolerance, MaxInnerIterations, MaxOuterIterations, MaxPatchupIterations,
                                       updateResidual ? 1 : 0, it.data(), TrueResidual.data());
    TotalOuterIterations = it[0];
    TotalInnerIterations.assign(it.begin() + 1, it.begin() + 1 + NBatch);
    TotalFinalStepIterations.assign(it.begin() + 1 + NBatch, it.end());
    GB_ASSERT_OK(rc);
  }
  void operator()(const FieldD &src_d_in, FieldD &sol_d) {        // ref :70-77
    std::vector<const gb_fermion *> s{src_d_in.h};
    std::vector<gb_fermion *> x{sol_d.h};
    int it[3] = {0, 0, 0};
    TrueResidual.assign(1, 0.);
    gb_fermop *mf = Linop_f.FusedSchurMatrix(), *md = Linop_d.FusedSchurMatrix();
    assert(mf && md && "MixedPrecisionConjugateGradientBatched needs SchurDiagMooeeOperator arguments");
    int rc = gb_mixed_cg_batched_schur(mf, md, 1, s.data(), x.data(), Tolerance, MaxInnerIterations, MaxOuterIterations, MaxPatchupIterations,
                                       updateResidual ? 1 : 0, it, TrueResidual.data());
    TotalOuterIterations = it[0]; TotalInnerIterations.assign(1, it[1]); TotalFinalStepIterations.assign(1, it[2]);
    GB_ASSERT_OK(rc);
  }
};

// ---- ConjugateGradientReliableUpdate (ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:36-270)
template <class FieldD, class FieldF> class ConjugateGradientReliableUpdate {
public:
  bool ErrorOnNoConverge;
  RealD Tolerance;
  Integer MaxIterations;
  Integer IterationsToComplete = 0, ReliableUpdatesPerformed = 0, IterationsToCleanup = 0;
  RealD TrueResidual = 0;
  LinearOperatorBase<FieldF> &Linop_f;
  LinearOperatorBase<FieldD> &Linop_d;
  GridBase *SinglePrecGrid;
  RealD Delta;
  ConjugateGradientReliableUpdate(RealD tol, Integer maxit, RealD _delta, GridBase *_sp_grid, LinearOperatorBase<FieldF> &_Linop_f,
                                  LinearOperatorBase<FieldD> &_Linop_d, bool err_on_no_conv = true)
      : ErrorOnNoConverge(err_on_no_conv), Tolerance(tol), MaxIterations(maxit), Linop_f(_Linop_f), Linop_d(_Linop_d), SinglePrecGrid(_sp_grid), Delta(_delta) {
    assert(Delta > 0. && Delta < 1. && "Expect  0 < Delta < 1");
  }
  void operator()(const FieldD &src, FieldD &psi) {
    gb_fermop *mf = Linop_f.FusedSchurMatrix(), *md = Linop_d.FusedSchurMatrix();
    assert(mf && md && "ConjugateGradientReliableUpdate needs SchurDiagMooeeOperator arguments");
    int it[3];
    int rc = gb_relup_cg_schur(mf, md, src.h, psi.h, Tolerance, MaxIterations, Delta, it, &TrueResidual);
    IterationsToComplete = it[0]; ReliableUpdatesPerformed = it[1]; IterationsToCleanup = it[2];
    if (rc == GB_ERR_NOT_CONVERGED) { if (ErrorOnNoConverge) assert(0 && "ConjugateGradientReliableUpdate did NOT converge"); return; }
    GB_ASSERT_OK(rc);
  }
};

// ---- ConjugateGradientMultiShift (ref: Grid/algorithms/approx/MultiShiftFunction.h:34-44 ; iterative/ConjugateGradientMultiShift.h:40-343)
class MultiShiftFunction {
public:
  int order = 0;
  std::vector<RealD> poles, residues, tolerances;
  RealD norm = 0, lo = 0, hi = 0;
  MultiShiftFunction() {}
  MultiShiftFunction(int n, RealD _lo, RealD _hi) : order(n), poles(n), residues(n), tolerances(n), lo(_lo), hi(_hi) {}
  RealD approx(RealD x) { RealD a = norm; for (size_t n = 0; n < poles.size(); n++) a += residues[n] / (x + poles[n]); return a; }
};
template <class Field> class ConjugateGradientMultiShift {
public:
  Integer MaxIterations;
  Integer IterationsToComplete = 0;
  std::vector<int> IterationsToCompleteShift;
  MultiShiftFunction shifts;
  std::vector<RealD> TrueResidualShift;
  ConjugateGradientMultiShift(Integer maxit, const MultiShiftFunction &_shifts) : MaxIterations(maxit), shifts(_shifts) {
    IterationsToCompleteShift.resize(_shifts.order); TrueResidualShift.resize(_shifts.order);
  }
  void operator()(LinearOperatorBase<Field> &Linop, const Field &src, std::vector<Field> &psi) {
    gb_fermop *m = Linop.FusedSchurMatrix();
    assert(m && "ConjugateGradientMultiShift needs a SchurDiagMooeeOperator / SchurStaggeredOperator");
    const int n = shifts.order;
    assert((int)psi.size() == n);
    std::vector<gb_fermion *> hs(n);
    for (int i = 0; i < n; i++) hs[i] = psi[i].h;
    std::vector<int> it(n + 1);
    int rc = gb_cg_multishift_schur(m, src.h, n, shifts.poles.data(), shifts.tolerances.data(), MaxIterations, hs.data(), it.data(), TrueResidualShift.data());
    for (int i = 0; i < n; i++) IterationsToCompleteShift[i] = it[i];
    IterationsToComplete = it[n];
    if (rc == GB_ERR_NOT_CONVERGED) { std::fprintf(stderr, "CG multi shift did not converge\n"); return; }   // ref :336-338
    GB_ASSERT_OK(rc);
  }
  void operator()(LinearOperatorBase<Field> &Linop, const Field &src, std::vector<Field> &results, Field &psi) {   // ref :69-82
    (*this)(Linop, src, results);
    GB_ASSERT_OK(gb_scale(psi.h, shifts.norm, src.h));
    for (int i = 0; i < shifts.order; i++) axpy(psi, shifts.residues[i], results[i], psi);
  }
};

// ---- ConjugateGradientMultiShiftMixedPrec (ref: Grid/algorithms/iterative/ConjugateGradientMultiShiftMixedPrec.h:73-410)
template <class FieldD, class FieldF> class ConjugateGradientMultiShiftMixedPrec {
public:
  Integer MaxIterationsMshift;
  Integer IterationsToComplete = 0;
  std::vector<int> IterationsToCompleteShift;
  MultiShiftFunction shifts;
  std::vector<RealD> TrueResidualShift;
  int ReliableUpdateFreq;
  GridBase *SinglePrecGrid;
  LinearOperatorBase<FieldF> &Linop_f;
  ConjugateGradientMultiShiftMixedPrec(Integer maxit, const MultiShiftFunction &_shifts, GridBase *_SinglePrecGrid, LinearOperatorBase<FieldF> &_Linop_f,
                                       int _ReliableUpdateFreq)
      : MaxIterationsMshift(maxit), shifts(_shifts), ReliableUpdateFreq(_ReliableUpdateFreq), SinglePrecGrid(_SinglePrecGrid), Linop_f(_Linop_f) {
    IterationsToCompleteShift.resize(_shifts.order); TrueResidualShift.resize(_shifts.order);
  }
  void operator()(LinearOperatorBase<FieldD> &Linop_d, const FieldD &src_d, std::vector<FieldD> &psi_d) {
    gb_fermop *mf = Linop_f.FusedSchurMatrix(), *md = Linop_d.FusedSchurMatrix();
    assert(mf && md && "ConjugateGradientMultiShiftMixedPrec needs SchurDiagMooeeOperator arguments");
    const int n = shifts.order;
    assert((int)psi_d.size() == n);
    std::vector<gb_fermion *> hs(n);
    for (int i = 0; i < n; i++) hs[i] = psi_d[i].h;
    std::vector<int> it(n + 1);
    GB_ASSERT_OK(gb_cg_multishift_mixed_schur(mf, md, src_d.h, n, shifts.poles.data(), shifts.tolerances.data(), MaxIterationsMshift, ReliableUpdateFreq,
                                              hs.data(), it.data(), TrueResidualShift.data()));
    for (int i = 0; i < n; i++) IterationsToCompleteShift[i] = it[i];
    IterationsToComplete = it[n];
  }
  void operator()(LinearOperatorBase<FieldD> &Linop, const FieldD &src, std::vector<FieldD> &results, FieldD &psi) {
    (*this)(Linop, src, results);
    GB_ASSERT_OK(gb_scale(psi.h, shifts.norm, src.h));
    for (int i = 0; i < shifts.order; i++) axpy(psi, shifts.residues[i], results[i], psi);
  }
};

// ---- SchurRedBlackDiagMooeeSolve / SchurRedBlackStaggeredSolve (ref: Grid/algorithms/iterative/SchurRedBlack.h:96-290,294-349,385-430)
//   SchurRedBlackDiagMooeeSolve<LatticeFermion> SchurSolver(CG);  SchurSolver(Ddwf, src, result);   solves M result = src
template <class Field, bool Staggered> class SchurRedBlackSolveT {
protected:
  OperatorFunction<Field> &_HermitianRBSolver;
  bool subGuess, useSolnAsInitGuess;
public:
  RealD TrueUnprecResidual = 0;   // the "true unprec resid" the reference logs (ref: :277-285)
  SchurRedBlackSolveT(OperatorFunction<Field> &HermitianRBSolver, const bool initSubGuess = false, const bool _solnAsInitGuess = false)
      : _HermitianRBSolver(HermitianRBSolver), subGuess(initSubGuess), useSolnAsInitGuess(_solnAsInitGuess) {}
  void subtractGuess(const bool initSubGuess) { subGuess = initSubGuess; }
  bool isSubtractGuess() { return subGuess; }
  template <class Matrix> void RedBlackSource(Matrix &_Matrix, const Field &src, Field &src_e, Field &src_o) { GB_ASSERT_OK(gb_schur_redblack_source(_Matrix.h, src.h, src_e.h, src_o.h)); }
  template <class Matrix> void RedBlackSolution(Matrix &_Matrix, const Field &sol_o, const Field &src_e, Field &sol) { GB_ASSERT_OK(gb_schur_redblack_solution(_Matrix.h, sol_o.h, src_e.h, sol.h)); }
  template <class Matrix> void RedBlackSolve(Matrix &_Matrix, const Field &src_o, Field &sol_o) {
    if constexpr (Staggered) { SchurStaggeredOperator<Matrix, Field> _HermOpEO(_Matrix); _HermitianRBSolver(_HermOpEO, src_o, sol_o); }
    else { SchurDiagMooeeOperator<Matrix, Field> _HermOpEO(_Matrix); _HermitianRBSolver(_HermOpEO, src_o, sol_o); }
    assert(sol_o.Checkerboard() == Odd);
  }
  template <class Matrix> void operator()(Matrix &_Matrix, const Field &in, Field &out) {
    if (FusedCG *f = subGuess ? nullptr : _HermitianRBSolver.Fused()) {
      double rs[2];
      int rc = gb_schur_solve(_Matrix.h, in.h, out.h, *f->Tolerance, *f->MaxIterations, useSolnAsInitGuess ? 1 : 0, f->IterationsToComplete, rs);
      *f->TrueResidual = rs[0]; TrueUnprecResidual = rs[1];
      if (rc == GB_ERR_NOT_CONVERGED) { if (*f->ErrorOnNoConverge) assert(0 && "ConjugateGradient did NOT converge"); return; }
      GB_ASSERT_OK(rc);
      return;
    }
    GridBase rb(*in.Grid()); rb.redblack = true;
    Field src_e(&rb), src_o(&rb), sol_o(&rb), guess_save(&rb);
    RedBlackSource(_Matrix, in, src_e, src_o);
    if (useSolnAsInitGuess) pickCheckerboard(Odd, sol_o, out);
    else { sol_o.Zero(); sol_o.SetCheckerboard(Odd); }                 // ZeroGuesser
    guess_save = sol_o;
    RedBlackSolve(_Matrix, src_o, sol_o);
    if (subGuess) axpy(sol_o, -1.0, guess_save, sol_o);
    RedBlackSolution(_Matrix, sol_o, src_e, out);
    if (!subGuess) {
      Field resid(in.Grid());
      _Matrix.M(out, resid);
      axpy(resid, -1.0, in, resid);
      TrueUnprecResidual = std::sqrt(norm2(resid) / norm2(in));
    }
  }
};
template <class Field> using SchurRedBlackDiagMooeeSolve = SchurRedBlackSolveT<Field, false>;
template <class Field> using SchurRedBlackStaggeredSolve = SchurRedBlackSolveT<Field, true>;
template <class Field> using SchurRedBlackStagSolve = SchurRedBlackSolveT<Field, true>;

} // namespace gridb200
