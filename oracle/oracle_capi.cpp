// oracle_capi.cpp -- C entry points of the CPU ORACLE (test infrastructure only; see dirac_oracle.hpp).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// The product library (grid_b200/libgridb200.so) never links or dlopens this file.
#include "dirac_oracle.hpp"
#include "stag_oracle.hpp"
#include "solvers_oracle.hpp"
#include <array>
#include <chrono>
#include <omp.h>

using namespace oracle;

namespace {
struct OpBox {
  int prec; // 0 = fp32, 1 = fp64
  FermOp<float> f;
  FermOp<double> d;
};
template <class T> FermOp<T> &get(OpBox *b);
template <> FermOp<float> &get<float>(OpBox *b) { return b->f; }
template <> FermOp<double> &get<double>(OpBox *b) { return b->d; }

// op codes shared with include/gridb200.h (gb_opcode)
enum {
  OP_DHOP = 0, OP_DHOP_OE = 1, OP_DHOP_EO = 2, OP_M = 3, OP_MDAG = 4, OP_MEOOE = 5, OP_MEOOE_DAG = 6, OP_MOOEE = 7,
  OP_MOOEE_DAG = 8, OP_MOOEE_INV = 9, OP_MOOEE_INV_DAG = 10, OP_MPC = 11, OP_MPC_DAG = 12, OP_HERMOP = 13, OP_DW = 14,
  OP_MEOOE5D = 15, OP_MEOOEDAG5D = 16, OP_DMINUS = 17, OP_DMINUS_DAG = 18
};

template <class T> int applyT(OpBox *box, int which, const void *vin, void *vout, int dag, int cb_in, int half) {
  FermOp<T> &op = get<T>(box);
  const Spinor<T> *in = (const Spinor<T> *)vin;
  Spinor<T> *out = (Spinor<T> *)vout;
  const int64_t n4 = half ? op.g.V4cb() : op.g.V4();
  switch (which) {
  case OP_DHOP: op.DhopFull(in, out, dag); break;
  case OP_DHOP_OE: op.DhopOE(in, out, dag); break;
  case OP_DHOP_EO: op.DhopEO(in, out, dag); break;
  case OP_M: op.M(in, out); break;
  case OP_MDAG: op.Mdag(in, out); break;
  case OP_MEOOE: op.Meooe(in, out, cb_in); break;
  case OP_MEOOE_DAG: op.MeooeDag(in, out, cb_in); break;
  case OP_MOOEE: op.Mooee(n4, in, out); break;
  case OP_MOOEE_DAG: op.MooeeDag(n4, in, out); break;
  case OP_MOOEE_INV: op.MooeeInv(n4, in, out); break;
  case OP_MOOEE_INV_DAG: op.MooeeInvDag(n4, in, out); break;
  case OP_MPC: op.Mpc(in, out, cb_in); break;
  case OP_MPC_DAG: op.MpcDag(in, out, cb_in); break;
  case OP_HERMOP: op.HermOp(in, out, cb_in); break;
  case OP_DW: op.DW(in, out, dag); break;
  case OP_MEOOE5D: op.Meooe5D(n4, in, out); break;
  case OP_MEOOEDAG5D: op.MeooeDag5D(n4, in, out); break;
  case OP_DMINUS: op.Dminus(in, out, 0); break;
  case OP_DMINUS_DAG: op.Dminus(in, out, 1); break;
  default: return -1;
  }
  return 0;
}
// Physical 4D <-> 5D maps.  which: 0 ImportPhysicalFermionSource (4D -> 5D), 1 ImportUnphysicalFermion
// (4D -> 5D), 2 ExportPhysicalFermionSolution (5D -> 4D), 3 ExportPhysicalFermionSource (5D -> 4D)
template <class T> static int physicalT(FermOp<T> &op, int which, const void *in, void *out) {
  const Spinor<T> *i = (const Spinor<T> *)in; Spinor<T> *o = (Spinor<T> *)out;
  switch (which) {
  case 0: op.ImportPhysicalFermionSource(i, o); break;
  case 1: op.ImportUnphysicalFermion(i, o); break;
  case 2: op.ExportPhysicalFermionSolution(i, o); break;
  case 3: op.ExportPhysicalFermionSource(i, o); break;
  default: return -1;
  }
  return 0;
}
} // namespace

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// kind: 0 = Wilson 4D (Ls must be 1), 1 = Cayley 5D (Shamir b=1,c=0 / Moebius). prec: 0 fp32, 1 fp64.
void *orc_op_create(int kind, const int *L, int Ls, double mass, double M5, double b, double c, int prec) {
  OpBox *box = new OpBox();
  box->prec = prec;
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = Ls;
  auto init = [&](auto &op) {
    op.kind = kind == 0 ? OpKind::Wilson4D : OpKind::Cayley5D;
    op.g = g; op.mass = mass;
    if (kind == 1) op.k = cayleyCoeffs(Ls, mass, M5, b, c);
  };
  if (prec == 0) init(box->f); else init(box->d);
  return box;
}
void orc_op_destroy(void *h) { delete (OpBox *)h; }

// Umu: [V4][4][3][3] complex in the operator's precision; phases: 4 complex doubles (re,im) or NULL for periodic.
void orc_op_import_gauge(void *h, const void *Umu, const double *phases) {
  OpBox *box = (OpBox *)h;
  cx<double> ph[4];
  for (int i = 0; i < 4; i++) ph[i] = phases ? cx<double>(phases[2 * i], phases[2 * i + 1]) : cx<double>(1.0, 0.0);
  if (box->prec == 0) box->f.importGauge((const ColourMatrix<float> *)Umu, ph);
  else box->d.importGauge((const ColourMatrix<double> *)Umu, ph);
}
// copies out the doubled links [V4][8][3][3] (for direct tests of DoubleStore)
void orc_op_export_doubled(void *h, void *out) {
  OpBox *box = (OpBox *)h;
  if (box->prec == 0) std::memcpy(out, box->f.Uds.data(), box->f.Uds.size() * sizeof(ColourMatrix<float>));
  else std::memcpy(out, box->d.Uds.data(), box->d.Uds.size() * sizeof(ColourMatrix<double>));
}
// Cayley coefficient vectors, each of length Ls: order bs,cs,bee,cee,dee,lee,leem,uee,ueem
void orc_op_coeffs(void *h, double *out) {
  OpBox *box = (OpBox *)h;
  const CayleyCoeffs &k = box->prec == 0 ? box->f.k : box->d.k;
  const std::vector<double> *v[9] = {&k.bs, &k.cs, &k.bee, &k.cee, &k.dee, &k.lee, &k.leem, &k.uee, &k.ueem};
  for (int i = 0; i < 9; i++) for (int s = 0; s < k.Ls; s++) out[i * k.Ls + s] = (*v[i])[s];
}

int orc_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half) {
  OpBox *box = (OpBox *)h;
  return box->prec == 0 ? applyT<float>(box, which, in, out, dag, cb_in, half) : applyT<double>(box, which, in, out, dag, cb_in, half);
}
// independent naive Cshift-style hopping term on the ORIGINAL links (periodic)
void orc_dhop_naive(const int *L, int Ls, int prec, const void *Umu, const void *in, void *out, int dag) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = Ls;
  if (prec == 0) DhopNaive(g, (const ColourMatrix<float> *)Umu, (const Spinor<float> *)in, (Spinor<float> *)out, dag);
  else DhopNaive(g, (const ColourMatrix<double> *)Umu, (const Spinor<double> *)in, (Spinor<double> *)out, dag);
}
void orc_pick_checkerboard(const int *L, int Ls, int prec, int cb, void *half, const void *full) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = Ls;
  if (prec == 0) pickCheckerboard(g, cb, (Spinor<float> *)half, (const Spinor<float> *)full);
  else pickCheckerboard(g, cb, (Spinor<double> *)half, (const Spinor<double> *)full);
}
void orc_set_checkerboard(const int *L, int Ls, int prec, int cb, void *full, const void *half) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = Ls;
  if (prec == 0) setCheckerboard(g, cb, (Spinor<float> *)full, (const Spinor<float> *)half);
  else setCheckerboard(g, cb, (Spinor<double> *)full, (const Spinor<double> *)half);
}
void orc_inner_product(int64_t nsites, int prec, const void *l, const void *r, double *out2) {
  cx<double> v = prec == 0 ? innerProduct(nsites, (const Spinor<float> *)l, (const Spinor<float> *)r)
                           : innerProduct(nsites, (const Spinor<double> *)l, (const Spinor<double> *)r);
  out2[0] = v.re; out2[1] = v.im;
}
// Schur-preconditioned CG on checkerboard cb. out: [iterations, converged], true_resid.
void orc_cg(void *h, int cb, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_true_resid) {
  OpBox *box = (OpBox *)h;
  CGResult r = box->prec == 0 ? ConjugateGradient(box->f, cb, (const Spinor<float> *)src, (Spinor<float> *)sol, tol, maxit)
                              : ConjugateGradient(box->d, cb, (const Spinor<double> *)src, (Spinor<double> *)sol, tol, maxit);
  out_iters[0] = r.iterations; out_iters[1] = r.converged; *out_true_resid = r.true_residual;
}
int orc_physical(void *h, int which, const void *in, void *out) {
  OpBox *box = (OpBox *)h;
  return box->prec == 0 ? physicalT(box->f, which, in, out) : physicalT(box->d, which, in, out);
}
// SURVEY 8 row f2.  orc_dhop_dir: FermionOperator::DhopDir(in, out, dir, disp).  orc_deriv: which 0 = DhopDeriv, 1 = MDeriv;
// mat is a LatticeGaugeField [V4][4][3][3] of the operator's precision.
void orc_dhop_dir(void *h, const void *in, void *out, int dir, int disp) {
  OpBox *box = (OpBox *)h;
  if (box->prec == 0) box->f.DhopDir((const Spinor<float> *)in, (Spinor<float> *)out, dir, disp);
  else box->d.DhopDir((const Spinor<double> *)in, (Spinor<double> *)out, dir, disp);
}
// one leg of the hopping term: point 0..3 forward mu, 4..7 backward; ocb < 0: full grid, else checkerboard hop into parity ocb
void orc_dhop_leg(void *h, const void *in, void *out, int point, int dag, int ocb) {
  OpBox *box = (OpBox *)h;
  auto run = [&](auto &op, auto *i, auto *o) { if (ocb < 0) op.DhopLeg(i, o, point, dag); else op.DhopLegCB(i, o, ocb, point, dag); };
  if (box->prec == 0) run(box->f, (const Spinor<float> *)in, (Spinor<float> *)out);
  else run(box->d, (const Spinor<double> *)in, (Spinor<double> *)out);
}
void orc_deriv(void *h, int which, void *mat, const void *U, const void *V, int dag) {
  OpBox *box = (OpBox *)h;
  auto run = [&](auto &op, auto *m, auto *u, auto *v) { if (which == 0) op.DhopDeriv(m, u, v, dag); else op.MDeriv(m, u, v, dag); };
  if (box->prec == 0) run(box->f, (ColourMatrix<float> *)mat, (const Spinor<float> *)U, (const Spinor<float> *)V);
  else run(box->d, (ColourMatrix<double> *)mat, (const Spinor<double> *)U, (const Spinor<double> *)V);
}
// which 0 = MeoDeriv (U Even, V Odd), 1 = MoeDeriv (U Odd, V Even): writes the sites of U's parity of the full-lattice mat;
// which 2 = SchurDifferentiableOperator::MpcDeriv, 3 = MpcDagDeriv (U, V Odd; the whole Force)
void orc_deriv_eo(void *h, int which, void *mat, const void *U, const void *V, int dag) {
  OpBox *box = (OpBox *)h;
  auto run = [&](auto &op, auto *m, auto *u, auto *v) {
    if (which < 2) op.MeooeDeriv(m, u, v, dag, which == 0 ? Even : Odd);
    else op.MpcDeriv(m, u, v, which == 3);
  };
  if (box->prec == 0) run(box->f, (ColourMatrix<float> *)mat, (const Spinor<float> *)U, (const Spinor<float> *)V);
  else run(box->d, (ColourMatrix<double> *)mat, (const Spinor<double> *)U, (const Spinor<double> *)V);
}
// SchurRedBlackDiagMooeeSolve pieces: full-lattice src -> (src_e, src_o') ; (sol_o, src_e) -> full-lattice sol
void orc_redblack_source(void *h, const void *src, void *src_e, void *src_o) {
  OpBox *box = (OpBox *)h;
  if (box->prec == 0) box->f.RedBlackSource((const Spinor<float> *)src, (Spinor<float> *)src_e, (Spinor<float> *)src_o);
  else box->d.RedBlackSource((const Spinor<double> *)src, (Spinor<double> *)src_e, (Spinor<double> *)src_o);
}
void orc_redblack_solution(void *h, const void *sol_o, const void *src_e, void *sol) {
  OpBox *box = (OpBox *)h;
  if (box->prec == 0) box->f.RedBlackSolution((const Spinor<float> *)sol_o, (const Spinor<float> *)src_e, (Spinor<float> *)sol);
  else box->d.RedBlackSolution((const Spinor<double> *)sol_o, (const Spinor<double> *)src_e, (Spinor<double> *)sol);
}
// the whole solve M sol = src with CG on the odd checkerboard. out_iters: [iterations, converged]; out_resid: [CG true
// residual, unpreconditioned residual |M sol - src| / |src|]
void orc_schur_solve(void *h, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_resid) {
  OpBox *box = (OpBox *)h;
  CGResult r = box->prec == 0 ? SchurRedBlackDiagMooeeSolve(box->f, (const Spinor<float> *)src, (Spinor<float> *)sol, tol, maxit, out_resid + 1)
                              : SchurRedBlackDiagMooeeSolve(box->d, (const Spinor<double> *)src, (Spinor<double> *)sol, tol, maxit, out_resid + 1);
  out_iters[0] = r.iterations; out_iters[1] = r.converged; out_resid[0] = r.true_residual;
}
// ConjugateGradientReliableUpdate.  out_iters: [IterationsToComplete, ReliableUpdatesPerformed, IterationsToCleanup, converged]
void orc_relup_cg(void *h_d, void *h_f, int cb, const void *src_d, void *sol_d, double tol, int maxit, double delta, int *out_iters, double *out_true_resid) {
  OpBox *bd = (OpBox *)h_d, *bf = (OpBox *)h_f;
  RelUpResult r = ReliableUpdateCG(bd->d, bf->f, cb, (const Spinor<double> *)src_d, (Spinor<double> *)sol_d, tol, maxit, delta);
  out_iters[0] = r.iterations; out_iters[1] = r.reliable_updates; out_iters[2] = r.cleanup_iterations; out_iters[3] = r.converged;
  *out_true_resid = r.true_residual;
}
// out_iters: [inner, outer, final, converged]
void orc_mixed_cg(void *h_d, void *h_f, int cb, const void *src_d, void *sol_d, double tol, int maxinner, int maxouter,
                  int *out_iters, double *out_true_resid) {
  OpBox *bd = (OpBox *)h_d, *bf = (OpBox *)h_f;
  MixedCGResult r = MixedPrecisionCG(bd->d, bf->f, cb, (const Spinor<double> *)src_d, (Spinor<double> *)sol_d, tol, maxinner, maxouter);
  out_iters[0] = r.inner_iterations; out_iters[1] = r.outer_iterations; out_iters[2] = r.final_iterations; out_iters[3] = r.converged;
  *out_true_resid = r.true_residual;
}
// srcs / sols: nbatch fields back to back; out_iters: [outer, inner_0..inner_{n-1}, final_0..final_{n-1}]; out_tr: [n]
void orc_mixed_cg_batched(void *h_d, void *h_f, int cb, int nbatch, const void *srcs_d, void *sols_d, double tol, int maxinner, int maxouter,
                          int maxpatch, int *out_iters, double *out_tr) {
  OpBox *bd = (OpBox *)h_d, *bf = (OpBox *)h_f;
  const int64_t n = bd->d.V5cb();
  std::vector<const Spinor<double> *> s(nbatch);
  std::vector<Spinor<double> *> x(nbatch);
  for (int i = 0; i < nbatch; i++) { s[i] = (const Spinor<double> *)srcs_d + (size_t)i * n; x[i] = (Spinor<double> *)sols_d + (size_t)i * n; }
  BatchedCGResult r = MixedPrecisionCGBatched(bd->d, bf->f, cb, nbatch, s.data(), x.data(), tol, maxinner, maxouter, maxpatch);
  out_iters[0] = r.outer_iterations;
  for (int i = 0; i < nbatch; i++) { out_iters[1 + i] = r.inner_iterations[i]; out_iters[1 + nbatch + i] = r.final_iterations[i]; out_tr[i] = r.true_residual[i]; }
}
// Timed loop for the CPU baseline: applies `which` ncall times, returns seconds.
double orc_time_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half, int ncall) {
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < ncall; i++) orc_apply(h, which, in, out, dag, cb_in, half);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---------------------------------------------------------------- improved staggered (stag_oracle.hpp)
struct StagBox { int prec; StagOp<float> f; StagOp<double> d; };
void *orc_stag_create(const int *L, double mass, double c1, double c2, double u0, int prec) {
  StagBox *b = new StagBox(); b->prec = prec;
  auto init = [&](auto &op) { for (int i = 0; i < 4; i++) op.g.L[i] = L[i]; op.g.Ls = 1; op.mass = mass; op.c1 = c1; op.c2 = c2; op.u0 = u0; };
  if (prec == 0) init(b->f); else init(b->d);
  return b;
}
void orc_stag_destroy(void *h) { delete (StagBox *)h; }
// Uthin, Ufat: [V4][4][3][3] complex in the operator's precision
void orc_stag_import_gauge(void *h, const void *Uthin, const void *Ufat) {
  StagBox *b = (StagBox *)h;
  if (b->prec == 0) b->f.importGauge((const ColourMatrix<float> *)Uthin, (const ColourMatrix<float> *)Ufat);
  else b->d.importGauge((const ColourMatrix<double> *)Uthin, (const ColourMatrix<double> *)Ufat);
}
} // extern "C"
template <class T> static int stagApply(StagOp<T> &op, int which, const void *vin, void *vout, int dag, int cb_in, int half) {
  const ColourVector<T> *in = (const ColourVector<T> *)vin; ColourVector<T> *out = (ColourVector<T> *)vout;
  const int64_t n = half ? op.g.V4cb() : op.g.V4();
  switch (which) {
  case OP_DHOP: op.Dhop(in, out, dag); break;
  case OP_DHOP_OE: op.DhopCB(in, out, 1, dag); break;
  case OP_DHOP_EO: op.DhopCB(in, out, 0, dag); break;
  case OP_M: op.M(in, out); break;
  case OP_MDAG: op.Mdag(in, out); break;
  case OP_MEOOE: op.Meooe(in, out, cb_in, 0); break;
  case OP_MEOOE_DAG: op.Meooe(in, out, cb_in, 1); break;
  case OP_MOOEE: case OP_MOOEE_DAG: op.scale(n, out, (T)op.mass, in); break;
  case OP_MOOEE_INV: case OP_MOOEE_INV_DAG: op.scale(n, out, (T)(1.0 / op.mass), in); break;
  case OP_MPC: case OP_MPC_DAG: case OP_HERMOP: op.Mpc(in, out, cb_in); break;
  default: return -1;
  }
  return 0;
}
extern "C" {
int orc_stag_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half) {
  StagBox *b = (StagBox *)h;
  return b->prec == 0 ? stagApply(b->f, which, in, out, dag, cb_in, half) : stagApply(b->d, which, in, out, dag, cb_in, half);
}
void orc_stag_cg(void *h, int cb, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_true_resid) {
  StagBox *b = (StagBox *)h;
  CGResult r = b->prec == 0 ? StagConjugateGradient(b->f, cb, (const ColourVector<float> *)src, (ColourVector<float> *)sol, tol, maxit)
                            : StagConjugateGradient(b->d, cb, (const ColourVector<double> *)src, (ColourVector<double> *)sol, tol, maxit);
  out_iters[0] = r.iterations; out_iters[1] = r.converged; *out_true_resid = r.true_residual;
}
void orc_stag_redblack_source(void *h, const void *src, void *src_e, void *src_o) {
  StagBox *b = (StagBox *)h;
  if (b->prec == 0) StagRedBlackSource(b->f, (const ColourVector<float> *)src, (ColourVector<float> *)src_e, (ColourVector<float> *)src_o);
  else StagRedBlackSource(b->d, (const ColourVector<double> *)src, (ColourVector<double> *)src_e, (ColourVector<double> *)src_o);
}
void orc_stag_redblack_solution(void *h, const void *sol_o, const void *src_e, void *sol) {
  StagBox *b = (StagBox *)h;
  if (b->prec == 0) StagRedBlackSolution(b->f, (const ColourVector<float> *)sol_o, (const ColourVector<float> *)src_e, (ColourVector<float> *)sol);
  else StagRedBlackSolution(b->d, (const ColourVector<double> *)sol_o, (const ColourVector<double> *)src_e, (ColourVector<double> *)sol);
}
void orc_stag_schur_solve(void *h, const void *src, void *sol, double tol, int maxit, int *out_iters, double *out_resid) {
  StagBox *b = (StagBox *)h;
  CGResult r = b->prec == 0 ? StagSchurSolve(b->f, (const ColourVector<float> *)src, (ColourVector<float> *)sol, tol, maxit, out_resid + 1)
                            : StagSchurSolve(b->d, (const ColourVector<double> *)src, (ColourVector<double> *)sol, tol, maxit, out_resid + 1);
  out_iters[0] = r.iterations; out_iters[1] = r.converged; out_resid[0] = r.true_residual;
}
void orc_stag_dhop_naive(const int *L, int prec, const void *Uthin, const void *Ufat, double c1, double c2, double u0, const void *in, void *out, int dag) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = 1;
  if (prec == 0) StagDhopNaive(g, (const ColourMatrix<float> *)Uthin, (const ColourMatrix<float> *)Ufat, c1, c2, u0, (const ColourVector<float> *)in, (ColourVector<float> *)out, dag);
  else StagDhopNaive(g, (const ColourMatrix<double> *)Uthin, (const ColourMatrix<double> *)Ufat, c1, c2, u0, (const ColourVector<double> *)in, (ColourVector<double> *)out, dag);
}
// pick / set checkerboard for any site size (bytes per site)
void orc_pick_checkerboard_bytes(const int *L, int site_bytes, int cb, void *half, const void *full) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = 1;
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    if (Geometry::parity(x) != cb) continue;
    std::memcpy((char *)half + g.cb4(x) * site_bytes, (const char *)full + i4 * site_bytes, site_bytes);
  }
}
void orc_set_checkerboard_bytes(const int *L, int site_bytes, int cb, void *full, const void *half) {
  Geometry g; for (int i = 0; i < 4; i++) g.L[i] = L[i]; g.Ls = 1;
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    if (Geometry::parity(x) != cb) continue;
    std::memcpy((char *)full + i4 * site_bytes, (const char *)half + g.cb4(x) * site_bytes, site_bytes);
  }
}
// ConjugateGradientMultiShift on the Schur operator of checkerboard cb (Wilson types: MpcDagMpc; staggered: Mpc).
// results: nshift fields back to back.  out_iters: [nshift per-shift iterations..., IterationsToComplete, converged]
void orc_multishift_cg(void *h, int staggered, int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit,
                       void *results, int *out_iters, double *out_true_resid) {
  std::vector<double> m(poles, poles + nshift), t(tols, tols + nshift);
  MultiShiftResult R;
  auto run = [&](auto &op, auto *srcp, auto *resp, int64_t n, auto A) {
    using F = std::remove_const_t<std::remove_pointer_t<decltype(resp)>>;
    using T = std::remove_reference_t<decltype(op.mass_word())>;
    std::vector<F *> psi(nshift);
    for (int s = 0; s < nshift; s++) psi[s] = resp + (size_t)s * n;
    R = MultiShiftCG<T, F>(A, n, srcp, psi, m, t, maxit);
  };
  if (staggered) {
    StagBox *b = (StagBox *)h;
    if (b->prec == 0) run(b->f, (const ColourVector<float> *)src, (ColourVector<float> *)results, b->f.g.V4cb(), [&](const ColourVector<float> *i, ColourVector<float> *o) { b->f.Mpc(i, o, cb); });
    else run(b->d, (const ColourVector<double> *)src, (ColourVector<double> *)results, b->d.g.V4cb(), [&](const ColourVector<double> *i, ColourVector<double> *o) { b->d.Mpc(i, o, cb); });
  } else {
    OpBox *b = (OpBox *)h;
    if (b->prec == 0) run(b->f, (const Spinor<float> *)src, (Spinor<float> *)results, b->f.V5cb(), [&](const Spinor<float> *i, Spinor<float> *o) { b->f.HermOp(i, o, cb); });
    else run(b->d, (const Spinor<double> *)src, (Spinor<double> *)results, b->d.V5cb(), [&](const Spinor<double> *i, Spinor<double> *o) { b->d.HermOp(i, o, cb); });
  }
  for (int s = 0; s < nshift; s++) { out_iters[s] = R.iterations[s]; out_true_resid[s] = R.true_residual[s]; }
  out_iters[nshift] = R.iterations_to_complete; out_iters[nshift + 1] = R.converged;
}
// ConjugateGradientMultiShiftMixedPrec.  out_iters: [per-shift iterations..., IterationsToComplete, number of clean-up solves]
void orc_multishift_mixed_cg(void *h_d, void *h_f, int cb, const void *src, int nshift, const double *poles, const double *tols, int maxit,
                             int relup_freq, void *results, int *out_iters, double *out_true_resid) {
  OpBox *bd = (OpBox *)h_d, *bf = (OpBox *)h_f;
  const int64_t n = bd->d.V5cb();
  std::vector<Spinor<double> *> psi(nshift);
  for (int s = 0; s < nshift; s++) psi[s] = (Spinor<double> *)results + (size_t)s * n;
  MultiShiftMixedResult R = MultiShiftMixedPrecCG(bd->d, bf->f, cb, (const Spinor<double> *)src, psi, std::vector<double>(poles, poles + nshift),
                                                  std::vector<double>(tols, tols + nshift), maxit, relup_freq);
  for (int s = 0; s < nshift; s++) { out_iters[s] = R.iterations[s]; out_true_resid[s] = R.true_residual[s]; }
  out_iters[nshift] = R.iterations_to_complete; out_iters[nshift + 1] = R.cleanups;
}
double orc_stag_time_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half, int ncall) {
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < ncall; i++) orc_stag_apply(h, which, in, out, dag, cb_in, half);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
}
