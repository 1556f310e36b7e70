// schur.cu -- the full propagator solve  M psi = eta  around the red-black CG.
//   SchurRedBlackDiagMooeeSolve / SchurRedBlackStaggeredSolve   ref: Grid/algorithms/iterative/SchurRedBlack.h:238-290,294-349,385-430
//   Import/ExportPhysicalFermion*, ImportUnphysicalFermion      ref: .../CayleyFermion5DImplementation.h:58-130
//   (Dminus / DminusDag live in fermop.cu as opcodes GB_OP_DMINUS / GB_OP_DMINUS_DAG, ref: :132-153)
// Everything here is composition of the operator entry points of fermop.cu plus one small kernel that moves the two chiral
// halves of a 4D spinor field into / out of the s = 0 and s = Ls-1 walls of a 5D field.
#include "fermop.hpp"
#include "kernels_common.cuh"
#include "next_kernels.cuh"
#include <cmath>

namespace gb {

// one thread per 16-byte vec of the 4D field; the body is chiral_wall_elem (next_kernels.cuh, shared with the CPU emulation)
template <class T, int DIR>
__global__ void chiral_wall_kernel(typename Prec<T>::vec *f4, typename Prec<T>::vec *f5, int nparity, uint32_t nsite4, uint32_t hblk4,
                                   uint32_t hblk5, int Ls, int s_up, int s_lo) {
  chiral_wall_elem<T, DIR>(blockIdx.x * blockDim.x + threadIdx.x, f4, f5, nparity, nsite4, hblk4, hblk5, Ls, s_up, s_lo);
}

static void chk(int rc) { if (rc != GB_OK) throw Error(rc, gb_last_error()); }

static void chiral_walls(gb_fermop *op, gb_fermion *f4, gb_fermion *f5, int dir, int s_up, int s_lo) {
  GB_REQUIRE(f4 && f5, "null field");
  GB_REQUIRE(f4->ncomplex == 12 && f5->ncomplex == 12, "spinor fields expected");
  GB_REQUIRE(f4->grid == op->grid && f5->grid == op->grid && f4->prec == op->prec && f5->prec == op->prec, "field is not conformable with the operator");
  GB_REQUIRE(f4->Ls == 1 && f5->Ls == op->Ls && f4->kind == f5->kind, "expected a 4D field and a 5D field of the operator's Ls on the same kind of grid");
  gb_context *ctx = op->ctx;
  const uint32_t n = (uint32_t)(f4->hblk * nv_of(op->prec) * W * f4->nparity);
  const unsigned blocks = (n + 255) / 256;
  if (dir == 0) GB_CUDA(cudaMemsetAsync(f5->data, 0, f5->bytes, ctx->stream));
#define LAUNCH(T, D)                                                                                                          \
  chiral_wall_kernel<T, D><<<blocks, 256, 0, ctx->stream>>>((typename Prec<T>::vec *)f4->data, (typename Prec<T>::vec *)f5->data, f4->nparity, \
                                                           (uint32_t)f4->nsite4, (uint32_t)f4->hblk, (uint32_t)f5->hblk, op->Ls, s_up, s_lo)
  if (op->prec == GB_F32) { if (dir == 0) LAUNCH(float, 0); else LAUNCH(float, 1); }
  else { if (dir == 0) LAUNCH(double, 0); else LAUNCH(double, 1); }
#undef LAUNCH
  count_launch(ctx);
  check_launch(ctx, "chiral_wall");
  if (dir == 0) f5->cb = f4->cb; else f4->cb = f5->cb;
}

// which: 0 ImportPhysicalFermionSource, 1 ImportUnphysicalFermion, 2 ExportPhysicalFermionSolution, 3 ExportPhysicalFermionSource
static void physical_map(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out) {
  GB_REQUIRE(op && in && out && in != out, "null or aliased argument");
  if (op->kind != GB_KIND_CAYLEY) { chk(gb_copy(out, in)); return; }   // ref: FermionOperator.h:172-191 (identity maps)
  const int Ls = op->Ls;
  switch (which) {
  case 0: {  // 5D = Dminus [ P+ in4 at s=0, P- in4 at s=Ls-1 ]     ref: :115-130
    GB_REQUIRE(in->kind == GB_FULL, "ImportPhysicalFermionSource works on full-grid fields");
    gb_fermion *tmp = op_tmp_full(op, 1);
    chiral_walls(op, const_cast<gb_fermion *>(in), tmp, 0, 0, Ls - 1);
    op_apply(op, GB_OP_DMINUS, tmp, out, 0);
    break;
  }
  case 1: chiral_walls(op, const_cast<gb_fermion *>(in), out, 0, 0, Ls - 1); break;      // ref: :100-113
  case 2: chiral_walls(op, out, const_cast<gb_fermion *>(in), 1, Ls - 1, 0); break;      // P- sol_0 + P+ sol_{Ls-1}   ref: :58-69
  case 3: chiral_walls(op, out, const_cast<gb_fermion *>(in), 1, 0, Ls - 1); break;      // P+ src_0 + P- src_{Ls-1}   ref: :88-99
  default: GB_REQUIRE(false, "bad map");
  }
}

struct Tmp {   // RAII solver temporary (recycled through the context's field pool)
  gb_fermion *f = nullptr;
  Tmp(const gb_fermion *like, int prec) { f = fermion_create_like(like, prec); }
  ~Tmp() { gb_fermion_destroy(f); }
  operator gb_fermion *() const { return f; }
};

static void check_rb_fields(gb_fermop *op, const gb_fermion *full, const gb_fermion *h1, const gb_fermion *h2) {
  GB_REQUIRE(op && full && h1 && h2, "null argument");
  GB_REQUIRE(full->kind == GB_FULL && h1->kind == GB_HALF && h2->kind == GB_HALF, "expected (full, red-black, red-black) fields");
  for (const gb_fermion *f : {full, h1, h2})
    GB_REQUIRE(f->grid == op->grid && f->Ls == op->Ls && f->prec == op->prec, "field is not conformable with the operator");
}

// src_e = src|Even ; src_o' = MpcDag (src|Odd - Meooe MooeeInv src_e)    ref: SchurRedBlack.h:396-414
// staggered: the last factor is Mooee (= mass) instead of MpcDag            ref: :307-326
static void redblack_source(gb_fermop *op, const gb_fermion *src, gb_fermion *src_e, gb_fermion *src_o) {
  check_rb_fields(op, src, src_e, src_o);
  Tmp tmp(src_e, op->prec), Mtmp(src_e, op->prec);
  chk(gb_pick_checkerboard(GB_EVEN, src_e, src));
  chk(gb_pick_checkerboard(GB_ODD, src_o, src));
  op_apply(op, GB_OP_MOOEE_INV, src_e, tmp, 0);      GB_REQUIRE(tmp.f->cb == GB_EVEN, "checkerboard");
  op_apply(op, GB_OP_MEOOE, tmp, Mtmp, 0);           GB_REQUIRE(Mtmp.f->cb == GB_ODD, "checkerboard");
  chk(gb_axpy(tmp, -1.0, Mtmp, src_o));              // tmp = src_o - Mtmp   (Odd)
  op_apply(op, op->kind == GB_KIND_STAGGERED ? GB_OP_MOOEE : GB_OP_MPC_DAG, tmp, src_o, 0);
  GB_REQUIRE(src_o->cb == GB_ODD, "checkerboard");
}

// sol_e = MooeeInv (src_e - Meooe sol_o) ; sol = [sol_e | sol_o]    ref: SchurRedBlack.h:415-432
static void redblack_solution(gb_fermop *op, const gb_fermion *sol_o, const gb_fermion *src_e, gb_fermion *sol) {
  check_rb_fields(op, sol, sol_o, src_e);
  GB_REQUIRE(sol_o->cb == GB_ODD && src_e->cb == GB_EVEN, "RedBlackSolution(sol_o [Odd], src_e [Even], sol)");
  Tmp tmp(src_e, op->prec), sol_e(src_e, op->prec);
  op_apply(op, GB_OP_MEOOE, sol_o, tmp, 0);          GB_REQUIRE(tmp.f->cb == GB_EVEN, "checkerboard");
  chk(gb_axpy(tmp, -1.0, tmp, src_e));               // tmp = src_e - Meooe sol_o
  op_apply(op, GB_OP_MOOEE_INV, tmp, sol_e, 0);      GB_REQUIRE(sol_e.f->cb == GB_EVEN, "checkerboard");
  chk(gb_set_checkerboard(sol, sol_e));
  chk(gb_set_checkerboard(sol, sol_o));
}

// |M sol - src| / |src| on the full lattice (the "true unprec resid" the reference logs, ref: SchurRedBlack.h:277-285)
static double unprec_residual(gb_fermop *op, const gb_fermion *src, const gb_fermion *sol) {
  Tmp r(src, op->prec);
  op_apply(op, GB_OP_M, sol, r, 0);
  double nr, ns;
  chk(gb_axpy_norm(r, -1.0, src, r, &nr));
  chk(gb_norm2(src, &ns));
  return std::sqrt(nr / ns);
}

} // namespace gb

using namespace gb;

extern "C" {
int gb_op_import_physical_fermion_source(gb_fermop *op, const gb_fermion *in4d, gb_fermion *out5d) {
  GB_API_BEGIN
  physical_map(op, 0, in4d, out5d);
  GB_API_END
}
int gb_op_import_unphysical_fermion(gb_fermop *op, const gb_fermion *in4d, gb_fermion *out5d) {
  GB_API_BEGIN
  physical_map(op, 1, in4d, out5d);
  GB_API_END
}
int gb_op_export_physical_fermion_solution(gb_fermop *op, const gb_fermion *sol5d, gb_fermion *out4d) {
  GB_API_BEGIN
  physical_map(op, 2, sol5d, out4d);
  GB_API_END
}
int gb_op_export_physical_fermion_source(gb_fermop *op, const gb_fermion *src5d, gb_fermion *out4d) {
  GB_API_BEGIN
  physical_map(op, 3, src5d, out4d);
  GB_API_END
}
int gb_schur_redblack_source(gb_fermop *op, const gb_fermion *src, gb_fermion *src_e, gb_fermion *src_o) {
  GB_API_BEGIN
  redblack_source(op, src, src_e, src_o);
  GB_API_END
}
int gb_schur_redblack_solution(gb_fermop *op, const gb_fermion *sol_o, const gb_fermion *src_e, gb_fermion *sol) {
  GB_API_BEGIN
  redblack_solution(op, sol_o, src_e, sol);
  GB_API_END
}

int gb_schur_solve(gb_fermop *op, const gb_fermion *src, gb_fermion *sol, double tol, int maxit, int use_sol_as_guess, int *iters_out,
                   double resid_out[2]) {
  GB_API_BEGIN
  GB_REQUIRE(op && src && sol && src != sol, "null or aliased argument");
  GB_REQUIRE(src->kind == GB_FULL && sol->kind == GB_FULL, "SchurRedBlackSolve works on full-grid fields");
  fermion_check_same(src, sol);
  gb_fermion *like = nullptr;
  chk(src->ncomplex == 3 ? gb_staggered_fermion_create(src->grid, (gb_precision)src->prec, GB_HALF, &like)
                         : gb_fermion_create(src->grid, src->Ls, (gb_precision)src->prec, GB_HALF, &like));
  Tmp src_e(like, op->prec), src_o(like, op->prec);
  gb_fermion *sol_o = like;                       // zeroed at creation = ZeroGuesser
  struct G { gb_fermion *f; ~G() { gb_fermion_destroy(f); } } guard{like};
  redblack_source(op, src, src_e, src_o);
  if (use_sol_as_guess) chk(gb_pick_checkerboard(GB_ODD, sol_o, sol));   // ref: useSolnAsInitGuess, SchurRedBlack.h:253-257
  else sol_o->cb = GB_ODD;
  int iters = 0; double tr = 0;
  const int rc = gb_cg_schur(op, src_o, sol_o, tol, maxit, &iters, &tr);
  if (rc != GB_OK && rc != GB_ERR_NOT_CONVERGED) throw Error(rc, gb_last_error());
  redblack_solution(op, sol_o, src_e, sol);
  if (iters_out) *iters_out = iters;
  if (resid_out) { resid_out[0] = tr; resid_out[1] = unprec_residual(op, src, sol); }
  if (rc == GB_ERR_NOT_CONVERGED) throw Error(rc, "ConjugateGradient did NOT converge");
  GB_API_END
}

// The same solve with MixedPrecisionConjugateGradient as the red-black solver: source preparation, reconstruction and the
// residual check in fp64 (op_d), inner CG in fp32 (op_f).
int gb_schur_solve_mixed(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src, gb_fermion *sol, double tol, int max_inner, int max_outer,
                         int iters_out[3], double resid_out[2]) {
  GB_API_BEGIN
  GB_REQUIRE(op_f && op_d && src && sol && src != sol, "null or aliased argument");
  GB_REQUIRE(src->kind == GB_FULL && sol->kind == GB_FULL && src->prec == GB_F64, "SchurRedBlackSolve (mixed) works on fp64 full-grid fields");
  GB_REQUIRE(op_d->kind != GB_KIND_STAGGERED, "mixed-precision Schur solve is defined for Wilson-type operators");
  fermion_check_same(src, sol);
  gb_fermion *like = nullptr;
  chk(gb_fermion_create(src->grid, src->Ls, GB_F64, GB_HALF, &like));
  Tmp src_e(like, GB_F64), src_o(like, GB_F64);
  gb_fermion *sol_o = like;
  struct G { gb_fermion *f; ~G() { gb_fermion_destroy(f); } } guard{like};
  redblack_source(op_d, src, src_e, src_o);
  sol_o->cb = GB_ODD;
  double tr = 0;
  int its[3] = {0, 0, 0};
  const int rc = gb_mixed_cg_schur(op_f, op_d, src_o, sol_o, tol, max_inner, max_outer, its, &tr);
  if (rc != GB_OK && rc != GB_ERR_NOT_CONVERGED) throw Error(rc, gb_last_error());
  redblack_solution(op_d, sol_o, src_e, sol);
  if (iters_out) for (int i = 0; i < 3; i++) iters_out[i] = its[i];
  if (resid_out) { resid_out[0] = tr; resid_out[1] = unprec_residual(op_d, src, sol); }
  if (rc == GB_ERR_NOT_CONVERGED) throw Error(rc, "MixedPrecisionConjugateGradient did NOT converge");
  GB_API_END
}
}
