// cayley.cu -- fifth-dimension kernels of the Cayley (Shamir / Moebius) domain-wall operator.
//   m5d_kernel       replaces CayleyFermion5D::M5D / M5Ddag      ref: implementation/CayleyFermion5Dcache.h:43-115
//   mooee_inv_kernel replaces CayleyFermion5D::MooeeInv / ...Dag ref: implementation/CayleyFermion5Dcache.h:117-230
//   cayley_coeffs    restates SetCoefficientsInternal            ref: implementation/CayleyFermion5DImplementation.h:411-535
// Chirality decouples everything: upper spin components (vec index k < NV/2) only see P+ terms, lower only P-,
// so both kernels work on 16-byte vecs without ever assembling a spinor.
#include "fermop.hpp"
#include "kernels_common.cuh"

namespace gb {

CayleyCoeffs cayley_coeffs(int Ls, double mass, double M5, double b, double c) {
  CayleyCoeffs k; k.Ls = Ls; k.mass = mass; k.M5 = M5; k.b = b; k.c = c;
  auto rs = [&](std::vector<double> &v) { v.assign(Ls, 0.0); };
  rs(k.bs); rs(k.cs); rs(k.bee); rs(k.cee); rs(k.beo); rs(k.ceo); rs(k.aee); rs(k.dee); rs(k.lee); rs(k.leem); rs(k.uee); rs(k.ueem);
  const double bpc = b + c, bmc = b - c;
  for (int i = 0; i < Ls; i++) {
    const double omega = 1.0; // tanh approximation: gamma_s = 1 (ref: Grid/algorithms/approx/Zolotarev.cc:473), zolo_hi = 1
    k.bs[i] = 0.5 * (bpc / omega + bmc);
    k.cs[i] = 0.5 * (bpc / omega - bmc);
    k.bee[i] = k.bs[i] * (4.0 - M5) + 1.0;
    k.cee[i] = 1.0 - k.cs[i] * (4.0 - M5);
    k.beo[i] = k.bs[i];
    k.ceo[i] = -k.cs[i];
    k.aee[i] = k.cee[i];
  }
  for (int i = 0; i < Ls; i++) {
    k.dee[i] = k.bee[i];
    if (i < Ls - 1) {
      k.lee[i] = -k.cee[i + 1] / k.bee[i];
      k.leem[i] = mass * k.cee[Ls - 1] / k.bee[0];
      for (int j = 0; j < i; j++) k.leem[i] *= k.aee[j] / k.bee[j + 1];
      k.uee[i] = -k.aee[i] / k.bee[i];
      k.ueem[i] = mass;
      for (int j = 1; j <= i; j++) k.ueem[i] *= k.cee[j] / k.bee[j];
      k.ueem[i] *= k.aee[0] / k.bee[0];
    }
  }
  double delta_d = mass * k.cee[Ls - 1];
  for (int j = 0; j < Ls - 1; j++) delta_d *= k.cee[j] / k.bee[j];
  k.dee[Ls - 1] += delta_d;
  return k;
}

constexpr int MAXLS = 64;
template <class T> struct M5DCoef { T d[MAXLS], u[MAXLS], l[MAXLS]; };

// one thread per vec element of one parity block
template <class T, int DAG>
__global__ void m5d_kernel(const typename Prec<T>::vec *__restrict__ psi, const typename Prec<T>::vec *__restrict__ phi,
                           typename Prec<T>::vec *__restrict__ chi, const typename Prec<T>::vec *__restrict__ w, T alpha,
                           const M5DCoef<T> cf, int Ls, FastDiv dLs, uint32_t n5cb, size_t block_stride) {
  using P = Prec<T>;
  using V = typename P::vec;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nelem = ((n5cb + W - 1) / W) * P::NV * W;
  if (e >= nelem) return;
  const size_t boff = (size_t)blockIdx.y * block_stride;
  const uint32_t lane = e & (W - 1);
  const uint32_t r = e >> LOGW;
  const uint32_t blk = r / P::NV, k = r - blk * P::NV;
  const uint32_t i5 = blk * W + lane;
  if (i5 >= n5cb) return; // padding stays zero
  uint32_t site, s;
  dLs.divmod(i5, site, s);
  const bool upperSpin = k < P::NV / 2;
  // non-dag: upper_s multiplies P- psi_{s+1} (lower spins), lower_s multiplies P+ psi_{s-1} (upper spins); dag swaps
  const bool use_up = DAG ? upperSpin : !upperSpin;
  const uint32_t sn = use_up ? (s + 1 == (uint32_t)Ls ? 0 : s + 1) : (s == 0 ? Ls - 1 : s - 1);
  const uint32_t in = site * Ls + sn;
  const T cn = use_up ? cf.u[s] : cf.l[s];
  const V pn = __ldg(psi + boff + (((size_t)(in >> LOGW) * P::NV + k) << LOGW) + (in & (W - 1)));
  const V ph = __ldg(phi + boff + e);
  V res = vaxpby(cf.d[s], ph, cn, pn);
  if (w != nullptr) res = vaxpy(alpha, __ldg(w + boff + e), res);
  chi[boff + e] = res;
}

void m5d_apply(gb_fermop *op, const gb_fermion *psi, const gb_fermion *phi, gb_fermion *chi, const std::vector<double> &lower,
               const std::vector<double> &diag, const std::vector<double> &upper, int dag, const gb_fermion *w, double alpha) {
  gb_context *ctx = op->ctx;
  const int Ls = op->Ls;
  GB_REQUIRE(Ls <= MAXLS, "Ls > 64 not supported");
  GB_REQUIRE(psi->Ls == Ls && psi->prec == op->prec, "field does not match operator");
  fermion_check_same(psi, phi); fermion_check_same(psi, chi);
  if (w) fermion_check_same(psi, w);
  GB_REQUIRE(psi != chi, "M5D: psi and chi must be distinct fields");
  const uint32_t n5cb = (uint32_t)psi->n5cb;
  const uint32_t nelem = (uint32_t)(psi->hblk * nv_of(op->prec) * W);
  dim3 grid((nelem + 255) / 256, psi->nparity);
  const size_t bstride = (size_t)psi->hblk * nv_of(op->prec) * W;
  if (op->prec == GB_F32) {
    M5DCoef<float> cf;
    for (int s = 0; s < Ls; s++) { cf.d[s] = (float)diag[s]; cf.u[s] = (float)upper[s]; cf.l[s] = (float)lower[s]; }
    if (dag) m5d_kernel<float, 1><<<grid, 256, 0, ctx->stream>>>((const float4 *)psi->data, (const float4 *)phi->data, (float4 *)chi->data, w ? (const float4 *)w->data : nullptr, (float)alpha, cf, Ls, FastDiv(Ls), n5cb, bstride);
    else m5d_kernel<float, 0><<<grid, 256, 0, ctx->stream>>>((const float4 *)psi->data, (const float4 *)phi->data, (float4 *)chi->data, w ? (const float4 *)w->data : nullptr, (float)alpha, cf, Ls, FastDiv(Ls), n5cb, bstride);
  } else {
    M5DCoef<double> cf;
    for (int s = 0; s < Ls; s++) { cf.d[s] = diag[s]; cf.u[s] = upper[s]; cf.l[s] = lower[s]; }
    if (dag) m5d_kernel<double, 1><<<grid, 256, 0, ctx->stream>>>((const double2 *)psi->data, (const double2 *)phi->data, (double2 *)chi->data, w ? (const double2 *)w->data : nullptr, alpha, cf, Ls, FastDiv(Ls), n5cb, bstride);
    else m5d_kernel<double, 0><<<grid, 256, 0, ctx->stream>>>((const double2 *)psi->data, (const double2 *)phi->data, (double2 *)chi->data, w ? (const double2 *)w->data : nullptr, alpha, cf, Ls, FastDiv(Ls), n5cb, bstride);
  }
  count_launch(ctx);
  check_launch(ctx, "m5d");
  chi->cb = psi->cb;
}

// ------------------------------------------------------------------ MooeeInv
// Per chirality the LDU solve is one of two patterns (derived from CayleyFermion5Dcache.h:144-169 / :202-227):
//   type A: forward  chi_s = psi_s - a[s-1] chi_{s-1}              ; backward chi_s = chi_s/dee_s - bm[s] chi_{Ls-1}
//   type B: forward  chi_{Ls-1} = psi_{Ls-1} - sum_s am[s] psi_s   ; backward chi_s = chi_s/dee_s - b[s] chi_{s+1}
// non-dag: upper spins = A(a=lee, bm=ueem), lower spins = B(am=leem, b=uee)
// dag    : upper spins = B(am=ueem, b=lee), lower spins = A(a=uee, bm=leem)
template <class T> struct MooeeInvCoef { T a[MAXLS], bm[MAXLS], am[MAXLS], b[MAXLS], idee[MAXLS]; };

// one thread per (4D site, vec component k): walks s serially, everything stays in registers for Ls <= MAXLS
template <class T, int DAG>
__global__ void mooee_inv_kernel(const typename Prec<T>::vec *__restrict__ psi, typename Prec<T>::vec *__restrict__ chi,
                                 const MooeeInvCoef<T> cf, int Ls, uint32_t nsite4, size_t block_stride) {
  using P = Prec<T>;
  using V = typename P::vec;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsite4 * P::NV) return;
  const size_t boff = (size_t)blockIdx.y * block_stride;
  // consecutive threads walk consecutive k of the same site, then the next site
  const uint32_t site = e / P::NV, k = e - site * P::NV;
  const bool upperSpin = k < P::NV / 2;
  const bool typeA = DAG ? !upperSpin : upperSpin;
  auto addr = [&](uint32_t s) {
    const uint32_t i5 = site * Ls + s;
    return boff + (((size_t)(i5 >> LOGW) * P::NV + k) << LOGW) + (i5 & (W - 1));
  };
  if (typeA) {
    V prev = __ldg(psi + addr(0));
    chi[addr(0)] = prev;
    for (int s = 1; s < Ls; s++) {
      V cur = vaxpy(-cf.a[s - 1], prev, __ldg(psi + addr(s)));
      if (s < Ls - 1) chi[addr(s)] = cur;
      prev = cur;
    }
    const V last = vscale(cf.idee[Ls - 1], prev);
    chi[addr(Ls - 1)] = last;
    for (int s = Ls - 2; s >= 0; s--) {
      V c = chi[addr(s)];
      chi[addr(s)] = vaxpy(-cf.bm[s], last, vscale(cf.idee[s], c));
    }
  } else {
    V acc = vzero(V());
    for (int s = 0; s < Ls - 1; s++) acc = vaxpy(cf.am[s], __ldg(psi + addr(s)), acc);
    V last = __ldg(psi + addr(Ls - 1));
    last = vscale(cf.idee[Ls - 1], vaxpy((T)-1, acc, last));
    chi[addr(Ls - 1)] = last;
    V next = last;
    for (int s = Ls - 2; s >= 0; s--) {
      V c = vaxpy(-cf.b[s], next, vscale(cf.idee[s], __ldg(psi + addr(s))));
      chi[addr(s)] = c;
      next = c;
    }
  }
}

void mooee_inv_apply(gb_fermop *op, const gb_fermion *psi, gb_fermion *chi, int dag) {
  gb_context *ctx = op->ctx;
  const int Ls = op->Ls;
  GB_REQUIRE(Ls <= MAXLS && Ls >= 2, "MooeeInv needs 2 <= Ls <= 64");
  fermion_check_same(psi, chi);
  GB_REQUIRE(psi != chi, "MooeeInv: in and out must be distinct fields");
  GB_REQUIRE(psi->Ls == Ls && psi->prec == op->prec, "field does not match operator");
  const CayleyCoeffs &k = op->k;
  const uint32_t nsite4 = (uint32_t)psi->nsite4;
  const size_t bstride = (size_t)psi->hblk * nv_of(op->prec) * W;
  auto fill = [&](auto &cf) {
    using TT = std::remove_reference_t<decltype(cf.a[0])>;
    for (int s = 0; s < Ls; s++) {
      cf.idee[s] = (TT)(1.0 / k.dee[s]);
      if (!dag) { cf.a[s] = (TT)k.lee[s]; cf.bm[s] = (TT)k.ueem[s]; cf.am[s] = (TT)k.leem[s]; cf.b[s] = (TT)k.uee[s]; }
      else { cf.a[s] = (TT)k.uee[s]; cf.bm[s] = (TT)k.leem[s]; cf.am[s] = (TT)k.ueem[s]; cf.b[s] = (TT)k.lee[s]; }
    }
  };
  if (op->prec == GB_F32) {
    MooeeInvCoef<float> cf; fill(cf);
    dim3 grid((nsite4 * 6 + 127) / 128, psi->nparity);
    if (dag) mooee_inv_kernel<float, 1><<<grid, 128, 0, ctx->stream>>>((const float4 *)psi->data, (float4 *)chi->data, cf, Ls, nsite4, bstride);
    else mooee_inv_kernel<float, 0><<<grid, 128, 0, ctx->stream>>>((const float4 *)psi->data, (float4 *)chi->data, cf, Ls, nsite4, bstride);
  } else {
    MooeeInvCoef<double> cf; fill(cf);
    dim3 grid((nsite4 * 12 + 127) / 128, psi->nparity);
    if (dag) mooee_inv_kernel<double, 1><<<grid, 128, 0, ctx->stream>>>((const double2 *)psi->data, (double2 *)chi->data, cf, Ls, nsite4, bstride);
    else mooee_inv_kernel<double, 0><<<grid, 128, 0, ctx->stream>>>((const double2 *)psi->data, (double2 *)chi->data, cf, Ls, nsite4, bstride);
  }
  count_launch(ctx);
  check_launch(ctx, "mooee_inv");
  chi->cb = psi->cb;
}

} // namespace gb
