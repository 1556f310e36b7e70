// gridref_capi.cc -- C entry points over the UNMODIFIED reference (paboyle/Grid) compiled from /root/reference.
//
// TEST INFRASTRUCTURE ONLY (same rule as the oracle): loaded with ctypes by tests/, by the fixture generator
// (tests/golden/make_golden.py) and by bench.py's cpu_baseline / --impl reference legs.  The product library never
// links or dlopens it.  This file contains no reference source: it only CALLS the reference's public classes
//   WilsonFermion / DomainWallFermion / MobiusFermion / ImprovedStaggeredFermion  (Grid/qcd/action/fermion/*.h)
//   SchurDiagMooeeOperator / SchurStaggeredOperator                                (Grid/algorithms/LinearOperator.h)
//   ConjugateGradient / MixedPrecisionConjugateGradient                            (Grid/algorithms/iterative/*.h)
//   vectorizeFromLexOrdArray / unvectorizeToLexOrdArray / pick/setCheckerboard     (Grid/lattice/Lattice_transfer.h)
// through the same host layouts the C ABI of the product uses (include/gridb200.h), so a test can feed both the
// same bytes.  Built by oracle/Makefile.ref into oracle/_ref/libgridref.so.
#include <Grid/Grid.h>
#include <Grid/parallelIO/NerscIO.h>
#include <chrono>
#include <memory>
#include <sstream>
#include <cstdio>

using namespace Grid;

namespace {

enum {
  OP_DHOP = 0, OP_DHOP_OE = 1, OP_DHOP_EO = 2, OP_M = 3, OP_MDAG = 4, OP_MEOOE = 5, OP_MEOOE_DAG = 6, OP_MOOEE = 7,
  OP_MOOEE_DAG = 8, OP_MOOEE_INV = 9, OP_MOOEE_INV_DAG = 10, OP_MPC = 11, OP_MPC_DAG = 12, OP_HERMOP = 13, OP_DW = 14,
  OP_MEOOE5D = 15, OP_MEOOEDAG5D = 16, OP_DMINUS = 17, OP_DMINUS_DAG = 18
};
enum { KIND_WILSON = 0, KIND_CAYLEY = 1, KIND_STAGGERED = 2 };

bool g_inited = false;

struct BoxBase {
  int kind, prec, Ls;
  virtual ~BoxBase() {}
  virtual void import_gauge(const void *Umu, const double *phases) = 0;
  virtual int apply(int which, const void *in, void *out, int dag, int cb_in, int half) = 0;
  virtual void cg(int cb, const void *src, void *sol, double tol, int maxit, int *iters, double *tr) = 0;
  virtual void pick(int cb, void *half, const void *full) = 0;
  virtual void set(int cb, void *full, const void *half) = 0;
  // physical 4D <-> 5D maps and SchurRedBlack*Solve
  virtual int physical(int which, const void *in, void *out) { return -1; }
  virtual void redblack_source(const void *src, void *src_e, void *src_o) = 0;
  virtual void redblack_solution(const void *sol_o, const void *src_e, void *sol) = 0;
  virtual void schur_solve(const void *src, void *sol, double tol, int maxit, int *iters, double *resid) = 0;
  // DhopDir, DhopDeriv (which 0) / MDeriv (which 1) on the full grid; mat = LatticeGaugeField
  virtual int dhop_dir(const void *in, void *out, int dir, int disp) { return -1; }
  virtual int deriv(int which, void *mat, const void *U, const void *V, int dag) { return -1; }
  // which 2 = SchurDifferentiableOperator::MpcDeriv, 3 = MpcDagDeriv (U, V on the Odd checkerboard, Force on the full grid)
  virtual int deriv_eo(int which, void *mat, const void *U, const void *V) { return -1; }
  // ConjugateGradientMultiShift on the Schur operator of checkerboard cb
  virtual void multishift(int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit, void *results, int *iters, double *tr) = 0;
};

template <class Field> void import_lex(Field &f, const void *host) {
  typedef typename Field::vector_object::scalar_object sobj;
  const size_t n = f.Grid()->lSites();
  std::vector<sobj> tmp(n);
  std::memcpy((void *)tmp.data(), host, n * sizeof(sobj));
  vectorizeFromLexOrdArray(tmp, f);
}
template <class Field> void export_lex(const Field &f, void *host) {
  typedef typename Field::vector_object::scalar_object sobj;
  std::vector<sobj> tmp;
  unvectorizeToLexOrdArray(tmp, f);
  std::memcpy(host, (const void *)tmp.data(), tmp.size() * sizeof(sobj));
}

// the common body of the three SchurRedBlack entry points, for any Solver = SchurRedBlack{DiagMooee,Staggered}Solve<Field>
template <class Solver, class Matrix, class Field>
void rb_source(Matrix &M, GridBase *fg, GridBase *rbg, const void *src, void *src_e, void *src_o) {
  ConjugateGradient<Field> CG(1e-8, 1, false);
  Solver S(CG);
  Field f(fg), e(rbg), o(rbg);
  import_lex(f, src);
  S.RedBlackSource(M, f, e, o);
  export_lex(e, src_e); export_lex(o, src_o);
}
template <class Solver, class Matrix, class Field>
void rb_solution(Matrix &M, GridBase *fg, GridBase *rbg, const void *sol_o, const void *src_e, void *sol) {
  ConjugateGradient<Field> CG(1e-8, 1, false);
  Solver S(CG);
  Field f(fg), e(rbg), o(rbg);
  import_lex(o, sol_o); import_lex(e, src_e);
  o.Checkerboard() = Odd; e.Checkerboard() = Even;
  f = Zero();
  S.RedBlackSolution(M, o, e, f);
  export_lex(f, sol);
}
template <class Solver, class Matrix, class Field>
void rb_solve(Matrix &M, GridBase *fg, const void *src, void *sol, double tol, int maxit, int *iters, double *resid) {
  ConjugateGradient<Field> CG(tol, maxit, false);
  Solver S(CG);
  Field f(fg), x(fg), r(fg);
  import_lex(f, src);
  x = Zero();
  S(M, f, x);
  iters[0] = CG.IterationsToComplete; iters[1] = CG.IterationsToComplete < maxit; resid[0] = CG.TrueResidual;
  M.M(x, r); r = r - f;
  resid[1] = std::sqrt(norm2(r) / norm2(f));
  export_lex(x, sol);
}

// ConjugateGradientMultiShift<Field>(maxit, MultiShiftFunction{poles, tolerances})(Linop, src, results)
// iters: [per-shift IterationsToCompleteShift..., IterationsToComplete, converged]
template <class Field, class Linop>
void ms_solve(Linop &S, GridBase *rbg, int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit, void *results, int *iters, double *tr) {
  typedef typename Field::vector_object::scalar_object sobj;
  MultiShiftFunction shifts(nshift, 0.0, 1.0);
  shifts.order = nshift; shifts.norm = 0.0;
  for (int s = 0; s < nshift; s++) { shifts.poles[s] = poles[s]; shifts.tolerances[s] = tols[s]; shifts.residues[s] = 1.0; }
  ConjugateGradientMultiShift<Field> MSCG(maxit, shifts);
  MSCG.IterationsToComplete = -1;
  Field s(rbg);
  import_lex(s, src);
  s.Checkerboard() = cb;
  std::vector<Field> res(nshift, rbg);
  for (auto &f : res) f.Checkerboard() = cb;
  MSCG(S, s, res);
  const size_t n = rbg->lSites();
  for (int i = 0; i < nshift; i++) {
    export_lex(res[i], (char *)results + (size_t)i * n * sizeof(sobj));
    iters[i] = MSCG.IterationsToCompleteShift[i]; tr[i] = MSCG.TrueResidualShift[i];
  }
  iters[nshift] = MSCG.IterationsToComplete; iters[nshift + 1] = MSCG.IterationsToComplete >= 0;
}

// Simd tag -> grids
template <class vComplexT> struct Grids {
  GridCartesian *UGrid = nullptr, *FGrid = nullptr;
  GridRedBlackCartesian *UrbGrid = nullptr, *FrbGrid = nullptr;
  void make(const int *L, int Ls, bool fiveD) {
    Coordinate latt({L[0], L[1], L[2], L[3]});
    Coordinate mpi({1, 1, 1, 1});
    UGrid = SpaceTimeGrid::makeFourDimGrid(latt, GridDefaultSimd(Nd, vComplexT::Nsimd()), mpi);
    UrbGrid = SpaceTimeGrid::makeFourDimRedBlackGrid(UGrid);
    if (fiveD) {
      FGrid = SpaceTimeGrid::makeFiveDimGrid(Ls, UGrid);
      FrbGrid = SpaceTimeGrid::makeFiveDimRedBlackGrid(Ls, UGrid);
    }
  }
  ~Grids() { delete FrbGrid; delete FGrid; delete UrbGrid; delete UGrid; }
};

// ---- Wilson-type operators (WilsonFermion, DomainWallFermion, MobiusFermion)
template <class Impl, class vComplexT> struct WilsonBox : BoxBase {
  typedef typename Impl::FermionField FermionField;
  typedef typename Impl::GaugeField GaugeField;
  typedef FermionOperator<Impl> OpBase;
  Grids<vComplexT> G;
  double mass, M5, b, c;
  std::unique_ptr<GaugeField> Umu;
  std::unique_ptr<OpBase> op;
  std::unique_ptr<WilsonFermion5D<Impl>> dummy;

  GridBase *fgrid() { return kind == KIND_WILSON ? (GridBase *)G.UGrid : (GridBase *)G.FGrid; }
  GridBase *frbgrid() { return kind == KIND_WILSON ? (GridBase *)G.UrbGrid : (GridBase *)G.FrbGrid; }

  void import_gauge(const void *U, const double *phases) override {
    Umu.reset(new GaugeField(G.UGrid));
    import_lex(*Umu, U);
    typename Impl::ImplParams p;
    if (phases) for (int mu = 0; mu < Nd; mu++) p.boundary_phases[mu] = Complex(phases[2 * mu], phases[2 * mu + 1]);
    if (kind == KIND_WILSON) op.reset(new WilsonFermion<Impl>(*Umu, *G.UGrid, *G.UrbGrid, mass, p));
    else if (b == 1.0 && c == 0.0) op.reset(new DomainWallFermion<Impl>(*Umu, *G.FGrid, *G.FrbGrid, *G.UGrid, *G.UrbGrid, mass, M5, p));
    else op.reset(new MobiusFermion<Impl>(*Umu, *G.FGrid, *G.FrbGrid, *G.UGrid, *G.UrbGrid, mass, M5, b, c, p));
  }
  int apply(int which, const void *in, void *out, int dag, int cb_in, int half) override {
    GridBase *gi = half ? frbgrid() : fgrid();
    FermionField x(gi), y(gi);
    import_lex(x, in);
    if (half) { x.Checkerboard() = cb_in; y.Checkerboard() = cb_in; }
    SchurDiagMooeeOperator<OpBase, FermionField> S(*op);
    CayleyFermion5D<Impl> *cay = dynamic_cast<CayleyFermion5D<Impl> *>(op.get());
    WilsonFermion5D<Impl> *w5 = dynamic_cast<WilsonFermion5D<Impl> *>(op.get());
    switch (which) {
    case OP_DHOP: op->Dhop(x, y, dag); break;
    case OP_DHOP_OE: x.Checkerboard() = Even; op->DhopOE(x, y, dag); break;
    case OP_DHOP_EO: x.Checkerboard() = Odd; op->DhopEO(x, y, dag); break;
    case OP_M: op->M(x, y); break;
    case OP_MDAG: op->Mdag(x, y); break;
    case OP_MEOOE: op->Meooe(x, y); break;
    case OP_MEOOE_DAG: op->MeooeDag(x, y); break;
    case OP_MOOEE: op->Mooee(x, y); break;
    case OP_MOOEE_DAG: op->MooeeDag(x, y); break;
    case OP_MOOEE_INV: op->MooeeInv(x, y); break;
    case OP_MOOEE_INV_DAG: op->MooeeInvDag(x, y); break;
    case OP_MPC: S.Mpc(x, y); break;
    case OP_MPC_DAG: S.MpcDag(x, y); break;
    case OP_HERMOP: S.HermOp(x, y); break;
    case OP_DW: if (!w5) return -1; w5->DW(x, y, dag); break;
    case OP_MEOOE5D: if (!cay) return -1; cay->Meooe5D(x, y); break;
    case OP_MEOOEDAG5D: if (!cay) return -1; cay->MeooeDag5D(x, y); break;
    case OP_DMINUS: op->Dminus(x, y); break;
    case OP_DMINUS_DAG: op->DminusDag(x, y); break;
    default: return -1;
    }
    export_lex(y, out);
    return 0;
  }
  void cg(int cb, const void *src, void *sol, double tol, int maxit, int *iters, double *tr) override {
    FermionField s(frbgrid()), x(frbgrid());
    import_lex(s, src); import_lex(x, sol);
    s.Checkerboard() = cb; x.Checkerboard() = cb;
    SchurDiagMooeeOperator<OpBase, FermionField> S(*op);
    ConjugateGradient<FermionField> CG(tol, maxit, false);
    CG(S, s, x);
    iters[0] = CG.IterationsToComplete; iters[1] = CG.IterationsToComplete < maxit; *tr = CG.TrueResidual;
    export_lex(x, sol);
  }
  void pick(int cb, void *half, const void *full) override {
    FermionField f(fgrid()), h(frbgrid());
    import_lex(f, full);
    pickCheckerboard(cb, h, f);
    export_lex(h, half);
  }
  void set(int cb, void *full, const void *half) override {
    FermionField f(fgrid()), h(frbgrid());
    import_lex(f, full); import_lex(h, half);
    h.Checkerboard() = cb;
    setCheckerboard(f, h);
    export_lex(f, full);
  }
  // which: 0 ImportPhysicalFermionSource, 1 ImportUnphysicalFermion (4D -> 5D), 2 ExportPhysicalFermionSolution,
  // 3 ExportPhysicalFermionSource (5D -> 4D).  For WilsonFermion both sides are the 4D grid.
  int physical(int which, const void *in, void *out) override {
    FermionField f4(G.UGrid), f5(fgrid());
    if (which < 2) {
      import_lex(f4, in);
      if (which == 0) op->ImportPhysicalFermionSource(f4, f5); else op->ImportUnphysicalFermion(f4, f5);
      export_lex(f5, out);
    } else {
      import_lex(f5, in);
      if (which == 2) op->ExportPhysicalFermionSolution(f5, f4); else op->ExportPhysicalFermionSource(f5, f4);
      export_lex(f4, out);
    }
    return 0;
  }
  int dhop_dir(const void *in, void *out, int dir, int disp) override {
    FermionField x(fgrid()), y(fgrid());
    import_lex(x, in);
    // WilsonFermion counts directions 0..3, WilsonFermion5D 1..4 (the fifth dimension is 0): ref WilsonFermion5DImplementation.h:185
    op->DhopDir(x, y, kind == KIND_WILSON ? dir : dir + 1, disp);
    export_lex(y, out);
    return 0;
  }
  int deriv(int which, void *mat, const void *U, const void *V, int dag) override {
    FermionField u(fgrid()), v(fgrid());
    import_lex(u, U); import_lex(v, V);
    GaugeField m(G.UGrid);
    m = Zero();
    if (which == 0) op->DhopDeriv(m, u, v, dag); else op->MDeriv(m, u, v, dag);
    export_lex(m, mat);
    return 0;
  }
  int deriv_eo(int which, void *mat, const void *U, const void *V) override {
    FermionField u(frbgrid()), v(frbgrid());
    import_lex(u, U); import_lex(v, V);
    u.Checkerboard() = Odd; v.Checkerboard() = Odd;
    GaugeField F(G.UGrid);
    F = Zero();
    SchurDifferentiableOperator<Impl> S(*op);
    if (which == 2) S.MpcDeriv(F, u, v); else if (which == 3) S.MpcDagDeriv(F, u, v); else return -1;
    export_lex(F, mat);
    return 0;
  }
  typedef SchurRedBlackDiagMooeeSolve<FermionField> RBSolver;
  void redblack_source(const void *src, void *src_e, void *src_o) override { rb_source<RBSolver, OpBase, FermionField>(*op, fgrid(), frbgrid(), src, src_e, src_o); }
  void redblack_solution(const void *sol_o, const void *src_e, void *sol) override { rb_solution<RBSolver, OpBase, FermionField>(*op, fgrid(), frbgrid(), sol_o, src_e, sol); }
  void schur_solve(const void *src, void *sol, double tol, int maxit, int *iters, double *resid) override {
    rb_solve<RBSolver, OpBase, FermionField>(*op, fgrid(), src, sol, tol, maxit, iters, resid);
  }
  void multishift(int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit, void *results, int *iters, double *tr) override {
    SchurDiagMooeeOperator<OpBase, FermionField> S(*op);
    ms_solve<FermionField>(S, frbgrid(), cb, src, nshift, poles, tols, maxit, results, iters, tr);
  }
};

// ---- improved staggered (fat = thin = the imported links, as Benchmark_staggered.cc:92-96 does)
template <class Impl, class vComplexT> struct StagBox : BoxBase {
  typedef typename Impl::FermionField FermionField;
  typedef typename Impl::GaugeField GaugeField;
  Grids<vComplexT> G;
  double mass, c1, c2, u0;
  std::unique_ptr<GaugeField> Umu;
  std::unique_ptr<ImprovedStaggeredFermion<Impl>> op;
  void import_gauge(const void *U, const double *phases) override {
    Umu.reset(new GaugeField(G.UGrid));
    import_lex(*Umu, U);
    typename Impl::ImplParams p;
    op.reset(new ImprovedStaggeredFermion<Impl>(*Umu, *Umu, *G.UGrid, *G.UrbGrid, mass, c1, c2, u0, p));
  }
  int apply(int which, const void *in, void *out, int dag, int cb_in, int half) override {
    GridBase *gi = half ? (GridBase *)G.UrbGrid : (GridBase *)G.UGrid;
    FermionField x(gi), y(gi);
    import_lex(x, in);
    if (half) { x.Checkerboard() = cb_in; y.Checkerboard() = cb_in; }
    SchurStaggeredOperator<ImprovedStaggeredFermion<Impl>, FermionField> S(*op);
    switch (which) {
    case OP_DHOP: op->Dhop(x, y, dag); break;
    case OP_DHOP_OE: x.Checkerboard() = Even; op->DhopOE(x, y, dag); break;
    case OP_DHOP_EO: x.Checkerboard() = Odd; op->DhopEO(x, y, dag); break;
    case OP_M: op->M(x, y); break;
    case OP_MDAG: op->Mdag(x, y); break;
    case OP_MEOOE: op->Meooe(x, y); break;
    case OP_MEOOE_DAG: op->MeooeDag(x, y); break;
    case OP_MOOEE: op->Mooee(x, y); break;
    case OP_MOOEE_DAG: op->MooeeDag(x, y); break;
    case OP_MOOEE_INV: op->MooeeInv(x, y); break;
    case OP_MOOEE_INV_DAG: op->MooeeInvDag(x, y); break;
    case OP_MPC: S.Mpc(x, y); break;
    case OP_MPC_DAG: S.MpcDag(x, y); break;
    case OP_HERMOP: S.HermOp(x, y); break;
    default: return -1;
    }
    export_lex(y, out);
    return 0;
  }
  void cg(int cb, const void *src, void *sol, double tol, int maxit, int *iters, double *tr) override {
    FermionField s(G.UrbGrid), x(G.UrbGrid);
    import_lex(s, src); import_lex(x, sol);
    s.Checkerboard() = cb; x.Checkerboard() = cb;
    SchurStaggeredOperator<ImprovedStaggeredFermion<Impl>, FermionField> S(*op);
    ConjugateGradient<FermionField> CG(tol, maxit, false);
    CG(S, s, x);
    iters[0] = CG.IterationsToComplete; iters[1] = CG.IterationsToComplete < maxit; *tr = CG.TrueResidual;
    export_lex(x, sol);
  }
  void pick(int cb, void *half, const void *full) override {
    FermionField f(G.UGrid), h(G.UrbGrid);
    import_lex(f, full); pickCheckerboard(cb, h, f); export_lex(h, half);
  }
  void set(int cb, void *full, const void *half) override {
    FermionField f(G.UGrid), h(G.UrbGrid);
    import_lex(f, full); import_lex(h, half); h.Checkerboard() = cb; setCheckerboard(f, h); export_lex(f, full);
  }
  typedef SchurRedBlackStaggeredSolve<FermionField> RBSolver;
  typedef ImprovedStaggeredFermion<Impl> Op;
  void redblack_source(const void *src, void *src_e, void *src_o) override { rb_source<RBSolver, Op, FermionField>(*op, G.UGrid, G.UrbGrid, src, src_e, src_o); }
  void redblack_solution(const void *sol_o, const void *src_e, void *sol) override { rb_solution<RBSolver, Op, FermionField>(*op, G.UGrid, G.UrbGrid, sol_o, src_e, sol); }
  void schur_solve(const void *src, void *sol, double tol, int maxit, int *iters, double *resid) override {
    rb_solve<RBSolver, Op, FermionField>(*op, G.UGrid, src, sol, tol, maxit, iters, resid);
  }
  void multishift(int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit, void *results, int *iters, double *tr) override {
    SchurStaggeredOperator<Op, FermionField> S(*op);
    ms_solve<FermionField>(S, G.UrbGrid, cb, src, nshift, poles, tols, maxit, results, iters, tr);
  }
};

} // namespace

extern "C" {

// Grid_init with a synthetic command line; threads <= 0 keeps the OpenMP default
int gref_init(int threads) {
  if (g_inited) return 0;
  // --device-mem: the reference's software cache for lattice fields defaults to 128 MB and asserts on a larger field
  // (MemoryManagerCache.cc:236); 64 GB lets it hold the 32^4 x 16 and 64.64.32.16 x 16 fields of BASELINE configs[1] and [3]
  // GridThread::SetThreads caps --threads at omp_get_max_threads(), which torchrun's OMP_NUM_THREADS=1 pins to one: raise it first
  if (threads > 0) omp_set_num_threads(threads);
  static std::string a0 = "gridref", a1 = "--threads", a2, a3 = "--grid", a4 = "8.8.8.8", a5 = "--device-mem", a6 = "65536";
  a2 = std::to_string(threads > 0 ? threads : omp_get_max_threads());
  static char *args[] = {(char *)a0.c_str(), (char *)a1.c_str(), (char *)a2.c_str(), (char *)a3.c_str(), (char *)a4.c_str(),
                         (char *)a5.c_str(), (char *)a6.c_str(), nullptr};
  int argc = 7;
  char **argv = args;
  Grid_init(&argc, &argv);
  g_inited = true;
  return 0;
}
int gref_num_threads() { return GridThread::GetThreads(); }
int gref_nsimd(int prec) { return prec == 0 ? vComplexF::Nsimd() : vComplexD::Nsimd(); }
// 0 generic, 1 hand-unrolled (the reference's default on CPU), 2 asm
void gref_set_kernel_opt(int opt) {
  WilsonKernelsStatic::Opt = opt == 0 ? WilsonKernelsStatic::OptGeneric : opt == 1 ? WilsonKernelsStatic::OptHandUnroll : WilsonKernelsStatic::OptInlineAsm;
  StaggeredKernelsStatic::Opt = opt == 0 ? StaggeredKernelsStatic::OptGeneric : StaggeredKernelsStatic::OptHandUnroll;
}

// kind 0: WilsonFermion (Ls = 1); 1: DomainWallFermion (b=1,c=0) or MobiusFermion; 2: ImprovedStaggeredFermion
// (M5 = c1, b = c2, c = u0).  prec 0 = fp32 (…F types), 1 = fp64 (…D types).
void *gref_op_create(int kind, const int *L, int Ls, double mass, double M5, double b, double c, int prec) {
  gref_init(0);
  BoxBase *r = nullptr;
  if (kind == KIND_STAGGERED) {
    if (prec == 0) { auto *x = new StagBox<StaggeredImplF, vComplexF>(); x->G.make(L, 1, false); x->mass = mass; x->c1 = M5; x->c2 = b; x->u0 = c; r = x; }
    else { auto *x = new StagBox<StaggeredImplD, vComplexD>(); x->G.make(L, 1, false); x->mass = mass; x->c1 = M5; x->c2 = b; x->u0 = c; r = x; }
  } else {
    if (prec == 0) { auto *x = new WilsonBox<WilsonImplF, vComplexF>(); x->G.make(L, Ls, kind == KIND_CAYLEY); x->mass = mass; x->M5 = M5; x->b = b; x->c = c; r = x; }
    else { auto *x = new WilsonBox<WilsonImplD, vComplexD>(); x->G.make(L, Ls, kind == KIND_CAYLEY); x->mass = mass; x->M5 = M5; x->b = b; x->c = c; r = x; }
  }
  r->kind = kind; r->prec = prec; r->Ls = Ls;
  return r;
}
void gref_op_destroy(void *h) { delete (BoxBase *)h; }
void gref_op_import_gauge(void *h, const void *Umu, const double *phases) { ((BoxBase *)h)->import_gauge(Umu, phases); }
int gref_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half) {
  return ((BoxBase *)h)->apply(which, in, out, dag, cb_in, half);
}
void gref_pick_checkerboard(void *h, int cb, void *half, const void *full) { ((BoxBase *)h)->pick(cb, half, full); }
void gref_set_checkerboard(void *h, int cb, void *full, const void *half) { ((BoxBase *)h)->set(cb, full, half); }
void gref_cg(void *h, int cb, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_true_resid) {
  ((BoxBase *)h)->cg(cb, src, sol, tol, maxit, out_iters, out_true_resid);
}

int gref_physical(void *h, int which, const void *in, void *out) { return ((BoxBase *)h)->physical(which, in, out); }
void gref_redblack_source(void *h, const void *src, void *src_e, void *src_o) { ((BoxBase *)h)->redblack_source(src, src_e, src_o); }
void gref_redblack_solution(void *h, const void *sol_o, const void *src_e, void *sol) { ((BoxBase *)h)->redblack_solution(sol_o, src_e, sol); }
// SchurRedBlackDiagMooeeSolve / SchurRedBlackStaggeredSolve with ConjugateGradient and the ZeroGuesser, as
// tests/solver/Test_dwf_cg_schur.cc:46-50 and Test_staggered_cg_schur.cc set it up.
// out_iters: [iterations, converged]; out_resid: [CG TrueResidual, |M sol - src| / |src|]
void gref_schur_solve(void *h, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_resid) {
  ((BoxBase *)h)->schur_solve(src, sol, tol, maxit, out_iters, out_resid);
}

int gref_dhop_dir(void *h, const void *in, void *out, int dir, int disp) { return ((BoxBase *)h)->dhop_dir(in, out, dir, disp); }
int gref_deriv(void *h, int which, void *mat, const void *U, const void *V, int dag) { return ((BoxBase *)h)->deriv(which, mat, U, V, dag); }
int gref_deriv_eo(void *h, int which, void *mat, const void *U, const void *V) { return ((BoxBase *)h)->deriv_eo(which, mat, U, V); }
// NerscIO::writeConfiguration / readConfiguration of a LatticeGaugeFieldD given as a lexicographic [V4][4][3][3] complex double array
// (ref: tests/IO/Test_nersc_io.cc).  gref_nersc_read returns {plaquette, link_trace} of the header after the reference's own QA.
void gref_nersc_write(const int *L, const void *Umu, const char *path, int two_row) {
  gref_init(0);
  Grids<vComplexD> G; G.make(L, 1, false);
  LatticeGaugeFieldD U(G.UGrid);
  import_lex(U, Umu);
  NerscIO::writeConfiguration(U, std::string(path), two_row, 0, std::string("DWF"), std::string("UKQCD"), 7);
}
void gref_nersc_read(const int *L, void *Umu_out, const char *path, double *plaq_link) {
  gref_init(0);
  Grids<vComplexD> G; G.make(L, 1, false);
  LatticeGaugeFieldD U(G.UGrid);
  FieldMetaData header;
  NerscIO::readConfiguration(U, header, std::string(path));
  plaq_link[0] = header.plaquette; plaq_link[1] = header.link_trace;
  export_lex(U, Umu_out);
}
// ConjugateGradientMultiShift as tests/solver/Test_staggered_multishift.cc:98-107 drives it, with explicit poles / tolerances
void gref_multishift_cg(void *h, int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit, void *results,
                        int *out_iters, double *out_true_resid) {
  ((BoxBase *)h)->multishift(cb, src, nshift, poles, tols, maxit, results, out_iters, out_true_resid);
}

// MixedPrecisionConjugateGradient exactly as tests/Test_dwf_mixedcg_prec.cc:113-196 sets it up.
// out_iters: [inner, outer, final, converged]
void gref_mixed_cg(void *h_d, void *h_f, int cb, const void *src_d, void *sol_d, double tol, int maxinner, int maxouter,
                   int *out_iters, double *out_true_resid) {
  auto *bd = dynamic_cast<WilsonBox<WilsonImplD, vComplexD> *>((BoxBase *)h_d);
  auto *bf = dynamic_cast<WilsonBox<WilsonImplF, vComplexF> *>((BoxBase *)h_f);
  assert(bd && bf);
  typedef FermionOperator<WilsonImplD> OpD;
  typedef FermionOperator<WilsonImplF> OpF;
  SchurDiagMooeeOperator<OpD, LatticeFermionD> Sd(*bd->op);
  SchurDiagMooeeOperator<OpF, LatticeFermionF> Sf(*bf->op);
  LatticeFermionD s(bd->frbgrid()), x(bd->frbgrid());
  import_lex(s, src_d); import_lex(x, sol_d);
  s.Checkerboard() = cb; x.Checkerboard() = cb;
  MixedPrecisionConjugateGradient<LatticeFermionD, LatticeFermionF> mCG(tol, maxinner, maxouter, bf->frbgrid(), Sf, Sd);
  mCG(s, x);
  out_iters[0] = mCG.TotalInnerIterations; out_iters[1] = mCG.TotalOuterIterations; out_iters[2] = mCG.TotalFinalStepIterations;
  out_iters[3] = 1;
  *out_true_resid = mCG.TrueResidual;
  export_lex(x, sol_d);
}

// MixedPrecisionConjugateGradientBatched (Grid/algorithms/iterative/ConjugateGradientMixedPrecBatched.h).  The class keeps its iteration
// counts in locals and only logs them, so they are read back from its own GridLogMessage lines
//   "MixedPrecisionConjugateGradientBatched: solve i Inner CG iterations N Restarts R Final CG iterations F"   (:201)
// srcs / sols: nbatch fields back to back; out_iters: [restarts, inner_0.., final_0..]
void gref_mixed_cg_batched(void *h_d, void *h_f, int cb, int nbatch, const void *srcs_d, void *sols_d, double tol, int maxinner, int maxouter,
                           int maxpatch, int *out_iters) {
  auto *bd = dynamic_cast<WilsonBox<WilsonImplD, vComplexD> *>((BoxBase *)h_d);
  auto *bf = dynamic_cast<WilsonBox<WilsonImplF, vComplexF> *>((BoxBase *)h_f);
  assert(bd && bf);
  typedef FermionOperator<WilsonImplD> OpD;
  typedef FermionOperator<WilsonImplF> OpF;
  SchurDiagMooeeOperator<OpD, LatticeFermionD> Sd(*bd->op);
  SchurDiagMooeeOperator<OpF, LatticeFermionF> Sf(*bf->op);
  typedef LatticeFermionD::vector_object::scalar_object sobj;
  const size_t n = bd->frbgrid()->lSites();
  std::vector<LatticeFermionD> s(nbatch, bd->frbgrid()), x(nbatch, bd->frbgrid());
  for (int i = 0; i < nbatch; i++) {
    import_lex(s[i], (const char *)srcs_d + (size_t)i * n * sizeof(sobj)); import_lex(x[i], (const char *)sols_d + (size_t)i * n * sizeof(sobj));
    s[i].Checkerboard() = cb; x[i].Checkerboard() = cb;
  }
  MixedPrecisionConjugateGradientBatched<LatticeFermionD, LatticeFermionF> mCG(tol, maxinner, maxouter, maxpatch, bf->frbgrid(), Sf, Sd);
  std::stringstream log;
  std::streambuf *saved = std::cout.rdbuf(log.rdbuf());
  mCG(s, x);
  std::cout.rdbuf(saved);
  for (int i = 0; i < 1 + 2 * nbatch; i++) out_iters[i] = -1;
  std::string line;
  while (std::getline(log, line)) {
    const size_t k = line.find("MixedPrecisionConjugateGradientBatched: solve ");
    int i, inner, restarts, fin;
    if (k != std::string::npos && std::sscanf(line.c_str() + k, "MixedPrecisionConjugateGradientBatched: solve %d Inner CG iterations %d Restarts %d Final CG iterations %d",
                                               &i, &inner, &restarts, &fin) == 4 && i >= 0 && i < nbatch) {
      out_iters[0] = restarts; out_iters[1 + i] = inner; out_iters[1 + nbatch + i] = fin;
    }
  }
  for (int i = 0; i < nbatch; i++) export_lex(x[i], (char *)sols_d + (size_t)i * n * sizeof(sobj));
}

// ConjugateGradientReliableUpdate as tests/solver/Test_dwf_relupcg_prec.cc:88-104 sets it up.
// out_iters: [IterationsToComplete, ReliableUpdatesPerformed, IterationsToCleanup, converged]
void gref_relup_cg(void *h_d, void *h_f, int cb, const void *src_d, void *sol_d, double tol, int maxit, double delta, int *out_iters, double *out_true_resid) {
  auto *bd = dynamic_cast<WilsonBox<WilsonImplD, vComplexD> *>((BoxBase *)h_d);
  auto *bf = dynamic_cast<WilsonBox<WilsonImplF, vComplexF> *>((BoxBase *)h_f);
  assert(bd && bf);
  typedef FermionOperator<WilsonImplD> OpD;
  typedef FermionOperator<WilsonImplF> OpF;
  SchurDiagMooeeOperator<OpD, LatticeFermionD> Sd(*bd->op);
  SchurDiagMooeeOperator<OpF, LatticeFermionF> Sf(*bf->op);
  LatticeFermionD s(bd->frbgrid()), x(bd->frbgrid());
  import_lex(s, src_d); import_lex(x, sol_d);
  s.Checkerboard() = cb; x.Checkerboard() = cb;
  ConjugateGradientReliableUpdate<LatticeFermionD, LatticeFermionF> mCG(tol, maxit, delta, bf->frbgrid(), Sf, Sd, false);
  mCG.IterationsToCleanup = 0;
  mCG(s, x);
  out_iters[0] = mCG.IterationsToComplete; out_iters[1] = mCG.ReliableUpdatesPerformed; out_iters[2] = mCG.IterationsToCleanup; out_iters[3] = 1;
  // the true residual, recomputed here (the class only logs it)
  LatticeFermionD mmp(bd->frbgrid());
  Sd.HermOp(x, mmp);
  mmp = mmp - s;
  *out_true_resid = std::sqrt(norm2(mmp) / norm2(s));
  export_lex(x, sol_d);
}

// ConjugateGradientMultiShiftMixedPrec as tests/solver/Test_dwf_multishift_mixedprec.cc:121-126 drives it, explicit poles / tolerances.
// out_iters: [per-shift IterationsToCompleteShift..., IterationsToComplete, 0]
void gref_multishift_mixed_cg(void *h_d, void *h_f, int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit,
                              int relup_freq, void *results, int *out_iters, double *out_true_resid) {
  auto *bd = dynamic_cast<WilsonBox<WilsonImplD, vComplexD> *>((BoxBase *)h_d);
  auto *bf = dynamic_cast<WilsonBox<WilsonImplF, vComplexF> *>((BoxBase *)h_f);
  assert(bd && bf);
  typedef FermionOperator<WilsonImplD> OpD;
  typedef FermionOperator<WilsonImplF> OpF;
  SchurDiagMooeeOperator<OpD, LatticeFermionD> Sd(*bd->op);
  SchurDiagMooeeOperator<OpF, LatticeFermionF> Sf(*bf->op);
  MultiShiftFunction shifts(nshift, 0.0, 1.0);
  shifts.order = nshift; shifts.norm = 0.0;
  for (int s = 0; s < nshift; s++) { shifts.poles[s] = poles[s]; shifts.tolerances[s] = tols[s]; shifts.residues[s] = 1.0; }
  ConjugateGradientMultiShiftMixedPrec<LatticeFermionD, LatticeFermionF> mcg(maxit, shifts, bf->frbgrid(), Sf, relup_freq);
  LatticeFermionD s(bd->frbgrid());
  import_lex(s, src);
  s.Checkerboard() = cb;
  std::vector<LatticeFermionD> res(nshift, bd->frbgrid());
  for (auto &f : res) f.Checkerboard() = cb;
  mcg(Sd, s, res);
  typedef LatticeFermionD::vector_object::scalar_object sobj;
  const size_t n = bd->frbgrid()->lSites();
  for (int i = 0; i < nshift; i++) {
    export_lex(res[i], (char *)results + (size_t)i * n * sizeof(sobj));
    out_iters[i] = mcg.IterationsToCompleteShift[i]; out_true_resid[i] = mcg.TrueResidualShift[i];
  }
  out_iters[nshift] = mcg.IterationsToComplete; out_iters[nshift + 1] = 0;
}

// Timed loop for the CPU baseline: fields stay resident in Grid's own layout; returns seconds for ncall applications.
double gref_time_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half, int ncall) {
  BoxBase *b = (BoxBase *)h;
  double secs = -1;
  auto run = [&](auto *box) {
    typedef typename std::remove_pointer<decltype(box)>::type Box;
    typedef typename Box::FermionField FermionField;
    GridBase *gi = half ? (GridBase *)box->frbgrid() : (GridBase *)box->fgrid();
    FermionField x(gi), y(gi);
    import_lex(x, in);
    if (half) x.Checkerboard() = which == OP_DHOP_OE ? Even : Odd;
    auto call = [&]() {
      if (which == OP_DHOP) box->op->Dhop(x, y, dag);
      else if (which == OP_DHOP_OE) box->op->DhopOE(x, y, dag);
      else box->op->DhopEO(x, y, dag);
    };
    call(); // warm-up (stencil tables, first-touch)
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < ncall; i++) call();
    auto t1 = std::chrono::steady_clock::now();
    secs = std::chrono::duration<double>(t1 - t0).count();
    export_lex(y, out);
  };
  if (auto *p = dynamic_cast<WilsonBox<WilsonImplF, vComplexF> *>(b)) run(p);
  else if (auto *p = dynamic_cast<WilsonBox<WilsonImplD, vComplexD> *>(b)) run(p);
  return secs;
}
}
