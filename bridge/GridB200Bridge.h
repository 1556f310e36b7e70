// GridB200Bridge.h -- the reference-side binding of libgridb200.so: what a paboyle/Grid maintainer adds to run the Wilson / domain-wall
// hopping term and the operators built on it on a B200 through the C ABI (include/gridb200.h), nothing else in Grid changing.
//
//   B200DomainWallFermion<Impl> : public DomainWallFermion<Impl>     (ref: Grid/qcd/action/fermion/DomainWallFermion.h:36-131)
//   B200MobiusFermion<Impl>     : public MobiusFermion<Impl>         (ref: Grid/qcd/action/fermion/MobiusFermion.h:37-74)
// override the FermionOperator virtuals of the hot path (ref: Grid/qcd/action/fermion/FermionOperator.h:63-77,144)
//   M Mdag Meooe MeooeDag Mooee MooeeDag MooeeInv MooeeInvDag Dhop DhopOE DhopEO ImportGauge
// and forward them to gb_op_apply / gb_op_import_gauge.  Fields cross the boundary in Grid's own lexicographic host layout
// (unvectorizeToLexOrdArray / vectorizeFromLexOrdArray, ref: Grid/lattice/Lattice_transfer.h:1123-1166,1218-1260), which is the layout
// gb_fermion_import / gb_gauge_import take -- the minimal, obviously correct binding (one host<->device round trip per call; a
// production port keeps the device handle inside the Lattice object).
//
// Use: compile an UNMODIFIED Grid program with  -include bridge/GridB200Bridge.h -DGRID_B200_REPLACE_OPERATORS : the header pulls in
// <Grid/Grid.h> first (the program's own #include then is a no-op) and afterwards maps the names DomainWallFermion{F,D} /
// MobiusFermion{F,D} onto the bridge classes, so tests/Test_dwf_mixedcg_prec.cc and benchmarks/Benchmark_dwf_fp32.cc drive the
// library as they are (bridge/Makefile builds both; tests/test_gpu_bridge.py runs them and checks the reference's own asserts).
#pragma once
#include <Grid/Grid.h>
#include <gridb200.h>
#include <memory>

NAMESPACE_BEGIN(Grid);

namespace b200 {

inline void check(int rc, const char *what) {
  if (rc != GB_OK) {
    std::cerr << "GridB200Bridge: " << what << " failed: " << gb_last_error() << std::endl;
    assert(rc == GB_OK);
    abort();
  }
}
inline gb_context *context() {   // one device context per process (ref: acceleratorInit picks the rank's device once, Accelerator.cc)
  static gb_context *ctx = nullptr;
  if (!ctx) check(gb_context_create(0, &ctx), "gb_context_create");
  return ctx;
}
template <class scalar> constexpr gb_precision precision_of() { return sizeof(scalar) == 8 ? GB_F32 : GB_F64; }   // complex<float> is 8 bytes

template <class Field> void to_device(const Field &f, gb_fermion *h) {
  typedef typename Field::vector_object::scalar_object sobj;
  std::vector<sobj> lex;
  unvectorizeToLexOrdArray(lex, f);
  check(gb_fermion_import(h, lex.data(), precision_of<typename sobj::scalar_type>()), "gb_fermion_import");
  if (f.Grid()->_isCheckerBoarded) check(gb_fermion_set_checkerboard_tag(h, f.Checkerboard()), "gb_fermion_set_checkerboard_tag");
}
template <class Field> void from_device(const gb_fermion *h, Field &f) {
  typedef typename Field::vector_object::scalar_object sobj;
  std::vector<sobj> lex(f.Grid()->lSites());
  check(gb_fermion_export(h, lex.data(), precision_of<typename sobj::scalar_type>()), "gb_fermion_export");
  vectorizeFromLexOrdArray(lex, f);
  if (f.Grid()->_isCheckerBoarded) f.Checkerboard() = gb_fermion_checkerboard(h);
}

// the device side of one Cayley operator: grid, links, operator and one in / out field pair per grid kind
template <class Impl> class Device {
public:
  INHERIT_IMPL_TYPES(Impl);
  typedef typename FermionField::vector_object::scalar_object::scalar_type scalar_type;
  static constexpr gb_precision prec = precision_of<scalar_type>();
  gb_grid *grid = nullptr;
  gb_gauge *links = nullptr;
  gb_fermop *op = nullptr;
  gb_fermion *in_h[2] = {nullptr, nullptr}, *out_h[2] = {nullptr, nullptr};   // [GB_FULL], [GB_HALF]
  int Ls = 1;
  double phases[8];

  Device(GridCartesian &UGrid, int Ls_, const ImplParams &p) : Ls(Ls_) {
    int g[4], m[4];
    for (int d = 0; d < 4; d++) { g[d] = UGrid.GlobalDimensions()[d]; m[d] = UGrid.ProcessorGrid()[d]; }
    for (int d = 0; d < 4; d++) assert(m[d] == 1 && "GridB200Bridge: one rank per process group in this binding (gb_comm_init for more)");
    check(gb_grid_create(context(), g, m, &grid), "gb_grid_create");
    check(gb_gauge_create(grid, prec, &links), "gb_gauge_create");
    for (int d = 0; d < 4; d++) { phases[2 * d] = real(p.boundary_phases[d]); phases[2 * d + 1] = imag(p.boundary_phases[d]); }
    for (int k = 0; k < 2; k++) {
      check(gb_fermion_create(grid, Ls, prec, k ? GB_HALF : GB_FULL, &in_h[k]), "gb_fermion_create");
      check(gb_fermion_create(grid, Ls, prec, k ? GB_HALF : GB_FULL, &out_h[k]), "gb_fermion_create");
    }
  }
  ~Device() {
    for (int k = 0; k < 2; k++) { gb_fermion_destroy(in_h[k]); gb_fermion_destroy(out_h[k]); }
    gb_op_destroy(op); gb_gauge_destroy(links); gb_grid_destroy(grid);
  }
  void upload_links(const GaugeField &Umu) {
    typedef typename GaugeField::vector_object::scalar_object sobj;   // LorentzColourMatrix: [mu][row][col]
    std::vector<sobj> lex;
    unvectorizeToLexOrdArray(lex, Umu);
    check(gb_gauge_import(links, lex.data(), prec), "gb_gauge_import");
  }
  void apply(int which, const FermionField &in, FermionField &out, int dag) {
    const int k = in.Grid()->_isCheckerBoarded ? 1 : 0;
    to_device(in, in_h[k]);
    check(gb_op_apply(op, which, in_h[k], out_h[k], dag), "gb_op_apply");
    from_device(out_h[k], out);
  }
};

// the overrides, shared by the two operator classes (Base = DomainWallFermion<Impl> or MobiusFermion<Impl>)
#define GRID_B200_FORWARD_OPERATOR_VIRTUALS                                                                                         \
  void M(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_M, in, out, 0); }                                  \
  void Mdag(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MDAG, in, out, 0); }                            \
  void Meooe(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MEOOE, in, out, 0); }                          \
  void MeooeDag(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MEOOE_DAG, in, out, 0); }                   \
  void Mooee(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MOOEE, in, out, 0); }                          \
  void MooeeDag(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MOOEE_DAG, in, out, 0); }                   \
  void MooeeInv(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MOOEE_INV, in, out, 0); }                   \
  void MooeeInvDag(const FermionField &in, FermionField &out) override { dev->apply(GB_OP_MOOEE_INV_DAG, in, out, 0); }            \
  void Dhop(const FermionField &in, FermionField &out, int dag) override { dev->apply(GB_OP_DHOP, in, out, dag); }                 \
  void DhopOE(const FermionField &in, FermionField &out, int dag) override { dev->apply(GB_OP_DHOP_OE, in, out, dag); }            \
  void DhopEO(const FermionField &in, FermionField &out, int dag) override { dev->apply(GB_OP_DHOP_EO, in, out, dag); }            \
  void ImportGauge(const GaugeField &Umu) override {                                                                               \
    Base::ImportGauge(Umu);                                        /* the reference's own copy stays valid (DhopDeriv etc.) */   \
    if (dev) { dev->upload_links(Umu); b200::check(gb_op_import_gauge(dev->op, dev->links), "gb_op_import_gauge"); }                     \
  }

}   // namespace b200

template <class Impl> class B200DomainWallFermion : public DomainWallFermion<Impl> {
public:
  INHERIT_IMPL_TYPES(Impl);
  typedef DomainWallFermion<Impl> Base;
  std::unique_ptr<b200::Device<Impl>> dev;
  // ref: DomainWallFermion.h:108-113 (same argument list)
  B200DomainWallFermion(GaugeField &Umu, GridCartesian &FiveDimGrid, GridRedBlackCartesian &FiveDimRedBlackGrid, GridCartesian &FourDimGrid,
                        GridRedBlackCartesian &FourDimRedBlackGrid, RealD mass, RealD M5, const ImplParams &p = ImplParams())
    : Base(Umu, FiveDimGrid, FiveDimRedBlackGrid, FourDimGrid, FourDimRedBlackGrid, mass, M5, p) {
    dev.reset(new b200::Device<Impl>(FourDimGrid, this->Ls, p));
    dev->upload_links(Umu);
    b200::check(gb_op_create_dwf(dev->grid, dev->links, this->Ls, mass, M5, dev->phases, &dev->op), "gb_op_create_dwf");
  }
  GRID_B200_FORWARD_OPERATOR_VIRTUALS
};

template <class Impl> class B200MobiusFermion : public MobiusFermion<Impl> {
public:
  INHERIT_IMPL_TYPES(Impl);
  typedef MobiusFermion<Impl> Base;
  std::unique_ptr<b200::Device<Impl>> dev;
  // ref: MobiusFermion.h:45-51 (same argument list)
  B200MobiusFermion(GaugeField &Umu, GridCartesian &FiveDimGrid, GridRedBlackCartesian &FiveDimRedBlackGrid, GridCartesian &FourDimGrid,
                    GridRedBlackCartesian &FourDimRedBlackGrid, RealD mass, RealD M5, RealD b, RealD c, const ImplParams &p = ImplParams())
    : Base(Umu, FiveDimGrid, FiveDimRedBlackGrid, FourDimGrid, FourDimRedBlackGrid, mass, M5, b, c, p) {
    dev.reset(new b200::Device<Impl>(FourDimGrid, this->Ls, p));
    dev->upload_links(Umu);
    b200::check(gb_op_create_mobius(dev->grid, dev->links, this->Ls, mass, M5, b, c, dev->phases, &dev->op), "gb_op_create_mobius");
  }
  GRID_B200_FORWARD_OPERATOR_VIRTUALS
};

typedef B200DomainWallFermion<WilsonImplF> B200DomainWallFermionF;
typedef B200DomainWallFermion<WilsonImplD> B200DomainWallFermionD;
typedef B200MobiusFermion<WilsonImplF> B200MobiusFermionF;
typedef B200MobiusFermion<WilsonImplD> B200MobiusFermionD;

NAMESPACE_END(Grid);

#ifdef GRID_B200_REPLACE_OPERATORS
// from here on (i.e. in the program that force-included this header) the reference's operator names mean the bridge classes
#define DomainWallFermionF B200DomainWallFermionF
#define DomainWallFermionD B200DomainWallFermionD
#define MobiusFermionF B200MobiusFermionF
#define MobiusFermionD B200MobiusFermionD
#endif
