// force.cu -- single hop legs and the fermion force terms HMC needs, full-grid fields.
//   FermionOperator::DhopDir(in, out, dir, disp)            ref: WilsonFermion5DImplementation.h:183-200 ; WilsonFermionImplementation.h:344-360
//   WilsonKernels::DhopDirKernel                            ref: implementation/WilsonKernelsImplementation.h:375-412
//   DhopDeriv / DerivInternal                               ref: WilsonFermion5DImplementation.h:212-275 ; WilsonFermionImplementation.h:238-278
//   Impl::InsertForce4D / InsertForce5D (spin-traced outer product summed over s)   ref: WilsonImpl.h:173-238
//   CayleyFermion5D::MDeriv                                 ref: CayleyFermion5DImplementation.h:347-360
// A single leg is the generic hopping kernel with a leg mask (dhop_kernel, csrc/dhop.cu): halo exchange, boundary phases
// and the -1/2 prefactor are the hopping term's own, so DhopDir summed over the eight legs IS Dhop.
#include "fermop.hpp"
#include "kernels_common.cuh"
#include "next_kernels.cuh"

namespace gb {

void op_dhop_leg(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int point, int dag) {
  GB_REQUIRE(op && in && out && in != out, "null or aliased argument");
  GB_REQUIRE(op->kind != GB_KIND_STAGGERED, "DhopDir is defined for Wilson-type operators here");
  GB_REQUIRE(point >= 0 && point < 8, "stencil point out of range");
  GB_REQUIRE(op->Uds != nullptr, "operator has no gauge field: call ImportGauge first");
  for (const gb_fermion *f : {in, (const gb_fermion *)out})
    GB_REQUIRE(f->grid == op->grid && f->Ls == op->Ls && f->prec == op->prec && f->kind == GB_FULL && f->ncomplex == 12, "DhopDir: field is not a conformable full-grid field");
  // the leg mask lives in the generic kernel only: route around the tuned kernels and the overlapped multi-GPU forms
  struct Restore {
    gb_fermop *op; bool df, oc; int lm;
    ~Restore() { op->disable_fast = df; op->overlap_comms = oc; op->leg_mask = lm; }
  } restore{op, op->disable_fast, op->overlap_comms, op->leg_mask};
  op->disable_fast = true; op->overlap_comms = false; op->leg_mask = 1 << point;
  const void *ib[2] = {in->block(0), in->block(1)};
  void *ob[2] = {out->block(0), out->block(1)};
  dhop_blocks(op, ib, ob, 0, 2, dag ? 1 : 0, nullptr, 1, 0);
}

// the same leg between checkerboards: in has parity ip, out gets parity 1 - ip
void op_dhop_leg_cb(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int point, int dag) {
  GB_REQUIRE(op && in && out && in != out, "null or aliased argument");
  GB_REQUIRE(op->kind != GB_KIND_STAGGERED && point >= 0 && point < 8 && op->Uds != nullptr, "bad operator or stencil point");
  for (const gb_fermion *f : {in, (const gb_fermion *)out})
    GB_REQUIRE(f->grid == op->grid && f->Ls == op->Ls && f->prec == op->prec && f->kind == GB_HALF && f->ncomplex == 12, "field is not a conformable red-black field");
  struct Restore {
    gb_fermop *op; bool df, oc; int lm;
    ~Restore() { op->disable_fast = df; op->overlap_comms = oc; op->leg_mask = lm; }
  } restore{op, op->disable_fast, op->overlap_comms, op->leg_mask};
  op->disable_fast = true; op->overlap_comms = false; op->leg_mask = 1 << point;
  const int ip = in->cb, po = 1 - ip;
  const void *ib[2] = {nullptr, nullptr};
  void *ob[2] = {nullptr, nullptr};
  ib[ip] = in->data; ob[po] = out->data;
  dhop_blocks(op, ib, ob, po, 1, dag ? 1 : 0, nullptr, 1, 0);
  out->cb = po;
}

// one thread per (parity, cb site, c1*3 + c2); the body is insert_force_elem (next_kernels.cuh, shared with the CPU emulation)
template <class T>
__global__ void insert_force_kernel(T *__restrict__ mat, const typename Prec<T>::vec *__restrict__ Bt, const typename Prec<T>::vec *__restrict__ A,
                                    int Ls, int Lx, int Ly, int Lz, int origin_parity, uint32_t V4cb, size_t parity_stride /* vecs */, int mu,
                                    int p0, int npar, T sign) {
  insert_force_elem<T>(blockIdx.x * blockDim.x + threadIdx.x, mat, Bt, A, Ls, Lx, Ly, Lz, origin_parity, V4cb, parity_stride, mu, p0, npar, sign);
}

static void insert_force(gb_fermop *op, gb_gauge *mat, const gb_fermion *Btilde, const gb_fermion *A, int mu, double sign = 1.0) {
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  const int npar = A->nparity, p0 = A->kind == GB_FULL ? 0 : A->cb;
  const uint32_t n = (uint32_t)npar * (uint32_t)g->V4cb * 9;
  const int op_ = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  const size_t pstride = (size_t)A->hblk * nv_of(op->prec) * W;
  if (op->prec == GB_F32)
    insert_force_kernel<float><<<(n + 255) / 256, 256, 0, ctx->stream>>>((float *)mat->data, (const float4 *)Btilde->data, (const float4 *)A->data, op->Ls, g->ldims[0],
                                                                         g->ldims[1], g->ldims[2], op_, (uint32_t)g->V4cb, pstride, mu, p0, npar, (float)sign);
  else
    insert_force_kernel<double><<<(n + 255) / 256, 256, 0, ctx->stream>>>((double *)mat->data, (const double2 *)Btilde->data, (const double2 *)A->data, op->Ls, g->ldims[0],
                                                                          g->ldims[1], g->ldims[2], op_, (uint32_t)g->V4cb, pstride, mu, p0, npar, sign);
  count_launch(ctx);
  check_launch(ctx, "insert_force");
}

static void dhop_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *A, const gb_fermion *B, int dag) {
  GB_REQUIRE(op && mat && A && B, "null argument");
  GB_REQUIRE(mat->grid == op->grid && mat->prec == op->prec, "force field lives on another grid or precision");
  fermion_check_same(A, B);
  gb_fermion *Btilde = op_tmp_full(op, 1);
  GB_REQUIRE(A != Btilde && B != Btilde, "operator temporaries cannot be arguments");
  for (int mu = 0; mu < 4; mu++) {
    op_dhop_leg(op, B, Btilde, mu, dag);     // forward leg mu with the projector of Dhop^(dag)  (ref: gamma = mu (+ Nd if !dag))
    insert_force(op, mat, Btilde, A, mu);
  }
}

// DhopDerivEO / DhopDerivOE: A on one checkerboard, B on the other; writes the sites of A's parity of the full-lattice mat
// ref: WilsonFermion5DImplementation.h:277-305
static void dhop_deriv_cb(gb_fermop *op, gb_gauge *mat, const gb_fermion *A, const gb_fermion *B, int dag, double sign) {
  GB_REQUIRE(op && mat && A && B, "null argument");
  GB_REQUIRE(mat->grid == op->grid && mat->prec == op->prec, "force field lives on another grid or precision");
  GB_REQUIRE(A->kind == GB_HALF && B->kind == GB_HALF && A->cb != B->cb, "DhopDerivEO/OE: A and B live on opposite checkerboards");
  gb_fermion *Btilde = op_tmp_half(op, 3);
  GB_REQUIRE(A != Btilde && B != Btilde, "operator temporaries cannot be arguments");
  for (int mu = 0; mu < 4; mu++) {
    op_dhop_leg_cb(op, B, Btilde, mu, dag);
    insert_force(op, mat, Btilde, A, mu, sign);
  }
}
// MeoDeriv (U Even) / MoeDeriv (U Odd)   ref: CayleyFermion5DImplementation.h:361-390 ; Wilson: = DhopDerivEO / OE
static void meooe_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *U, const gb_fermion *V, int dag, double sign) {
  if (op->kind != GB_KIND_CAYLEY) { dhop_deriv_cb(op, mat, U, V, dag, sign); return; }
  gb_fermion *Din = op_tmp_half(op, 2);
  GB_REQUIRE(U != Din && V != Din, "operator temporaries cannot be arguments");
  if (!dag) { op_apply(op, GB_OP_MEOOE5D, V, Din, 0); dhop_deriv_cb(op, mat, U, Din, 0, sign); }
  else { op_apply(op, GB_OP_MEOOE5D, U, Din, 0); dhop_deriv_cb(op, mat, Din, V, 1, sign); }
}

} // namespace gb

using namespace gb;

extern "C" {
int gb_op_dhop_dir(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dir, int disp) {
  GB_API_BEGIN
  GB_REQUIRE(dir >= 0 && dir < 4 && (disp == 1 || disp == -1), "DhopDir(in, out, dir in 0..3, disp = +-1)");
  op_dhop_leg(op, in, out, disp == 1 ? dir : dir + 4, 0);
  GB_API_END
}
int gb_op_dhop_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *A, const gb_fermion *B, int dag) {
  GB_API_BEGIN
  dhop_deriv(op, mat, A, B, dag ? 1 : 0);
  GB_API_END
}
int gb_op_meooe_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *U, const gb_fermion *V, int dag) {
  GB_API_BEGIN
  GB_REQUIRE(op && op->kind != GB_KIND_STAGGERED, "force terms are defined for Wilson-type operators here");
  meooe_deriv(op, mat, U, V, dag ? 1 : 0, 1.0);
  GB_API_END
}
// SchurDifferentiableOperator::MpcDeriv / MpcDagDeriv   ref: Grid/qcd/action/pseudofermion/EvenOddSchurDifferentiable.h:52-137
int gb_op_mpc_deriv(gb_fermop *op, gb_gauge *Force, const gb_fermion *U, const gb_fermion *V, int dagger) {
  GB_API_BEGIN
  GB_REQUIRE(op && Force && U && V && op->kind != GB_KIND_STAGGERED, "null argument or staggered operator");
  GB_REQUIRE(U->kind == GB_HALF && V->kind == GB_HALF && U->cb == GB_ODD && V->cb == GB_ODD, "MpcDeriv: U and V live on the Odd checkerboard");
  gb_fermion *tmp1 = fermion_create_like(U, op->prec), *tmp2 = fermion_create_like(U, op->prec);
  struct G { gb_fermion *a, *b; ~G() { gb_fermion_destroy(a); gb_fermion_destroy(b); } } guard{tmp1, tmp2};
  // Force = -(ForceE + ForceO): the sign rides in the outer-product kernel
  if (!dagger) {
    op_apply(op, GB_OP_MEOOE, V, tmp1, 0); op_apply(op, GB_OP_MOOEE_INV, tmp1, tmp2, 0);
    meooe_deriv(op, Force, U, tmp2, 0, -1.0);                       // MoeDeriv(ForceO, U, tmp2, DaggerNo)
    op_apply(op, GB_OP_MEOOE_DAG, U, tmp1, 0); op_apply(op, GB_OP_MOOEE_INV_DAG, tmp1, tmp2, 0);
    meooe_deriv(op, Force, tmp2, V, 0, -1.0);                       // MeoDeriv(ForceE, tmp2, V, DaggerNo)
  } else {
    op_apply(op, GB_OP_MEOOE_DAG, V, tmp1, 0); op_apply(op, GB_OP_MOOEE_INV_DAG, tmp1, tmp2, 0);
    meooe_deriv(op, Force, U, tmp2, 1, -1.0);
    op_apply(op, GB_OP_MEOOE, U, tmp1, 0); op_apply(op, GB_OP_MOOEE_INV, tmp1, tmp2, 0);
    meooe_deriv(op, Force, tmp2, V, 1, -1.0);
  }
  GB_API_END
}
int gb_op_mderiv(gb_fermop *op, gb_gauge *mat, const gb_fermion *U, const gb_fermion *V, int dag) {
  GB_API_BEGIN
  GB_REQUIRE(op && U && V, "null argument");
  if (op->kind != GB_KIND_CAYLEY) { dhop_deriv(op, mat, U, V, dag ? 1 : 0); return GB_OK; }   // FermionOperator default: MDeriv = DhopDeriv
  gb_fermion *Din = op_tmp_full(op, 0);
  if (!dag) { op_apply(op, GB_OP_MEOOE5D, V, Din, 0); dhop_deriv(op, mat, U, Din, 0); }
  else { op_apply(op, GB_OP_MEOOE5D, U, Din, 0); dhop_deriv(op, mat, Din, V, 1); }
  GB_API_END
}
}
