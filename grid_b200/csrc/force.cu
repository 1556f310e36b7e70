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

// mat[lex][mu][c1][c2] = sum_s sum_spin Btilde(x,s)[spin][c1] * conj(A(x,s)[spin][c2])
// one thread per (parity, cb site, c1*3 + c2); fields in the blocked layout (internal.hpp), mat lexicographic
template <class T>
__global__ void insert_force_kernel(T *__restrict__ mat, const typename Prec<T>::vec *__restrict__ Bt, const typename Prec<T>::vec *__restrict__ A,
                                    int Ls, int Lx, int Ly, int Lz, int origin_parity, uint32_t V4cb, size_t parity_stride /* vecs */, int mu) {
  using P = Prec<T>;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2u * V4cb * 9) return;
  const uint32_t k9 = e % 9, sp_ = e / 9;
  const uint32_t p = sp_ / V4cb, site = sp_ - p * V4cb;
  const int c1 = k9 / 3, c2 = k9 - 3 * c1;
  const int Lxh = Lx / 2;
  uint32_t r = site;
  const int xh = r % Lxh; r /= Lxh;
  const int y = r % Ly; r /= Ly;
  const int z = r % Lz;
  const int t = r / Lz;
  const int x = 2 * xh + ((p + origin_parity + y + z + t) & 1);
  const size_t lex = x + (size_t)Lx * (y + (size_t)Ly * (z + (size_t)Lz * t));
  const T *bs = (const T *)(Bt + (size_t)p * parity_stride);
  const T *as = (const T *)(A + (size_t)p * parity_stride);
  constexpr int CPV = sizeof(T) == 4 ? 2 : 1;          // complex numbers per 16-byte vec
  T re = 0, im = 0;
  for (int s = 0; s < Ls; s++) {
    const uint32_t i5 = site * Ls + s;
    const size_t base = ((size_t)(i5 >> LOGW) * P::NV) << LOGW;   // vec index of element (i5, k = 0) minus the lane
    const uint32_t lane = i5 & (W - 1);
#pragma unroll
    for (int spin = 0; spin < 4; spin++) {
      const int kb = spin * 3 + c1, ka = spin * 3 + c2;            // complex component index 0..11
      const size_t ob = ((base + ((size_t)(kb / CPV) << LOGW) + lane) * CPV + (kb % CPV)) * 2;
      const size_t oa = ((base + ((size_t)(ka / CPV) << LOGW) + lane) * CPV + (ka % CPV)) * 2;
      const T br = bs[ob], bi = bs[ob + 1], ar = as[oa], ai = as[oa + 1];
      re += br * ar + bi * ai;     // b * conj(a)
      im += bi * ar - br * ai;
    }
  }
  T *m = mat + ((lex * 4 + mu) * 9 + k9) * 2;
  m[0] = re; m[1] = im;
}

static void insert_force(gb_fermop *op, gb_gauge *mat, const gb_fermion *Btilde, const gb_fermion *A, int mu) {
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  const uint32_t n = 2u * (uint32_t)g->V4cb * 9;
  const int op_ = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  const size_t pstride = (size_t)A->hblk * nv_of(op->prec) * W;
  if (op->prec == GB_F32)
    insert_force_kernel<float><<<(n + 255) / 256, 256, 0, ctx->stream>>>((float *)mat->data, (const float4 *)Btilde->data, (const float4 *)A->data, op->Ls, g->ldims[0],
                                                                         g->ldims[1], g->ldims[2], op_, (uint32_t)g->V4cb, pstride, mu);
  else
    insert_force_kernel<double><<<(n + 255) / 256, 256, 0, ctx->stream>>>((double *)mat->data, (const double2 *)Btilde->data, (const double2 *)A->data, op->Ls, g->ldims[0],
                                                                          g->ldims[1], g->ldims[2], op_, (uint32_t)g->V4cb, pstride, mu);
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
