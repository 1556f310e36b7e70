// solver.cu -- ConjugateGradient and MixedPrecisionConjugateGradient on the device.
//   ref: Grid/algorithms/iterative/ConjugateGradient.h:68-257 (update order, stopping rule cp <= tol^2 |src|^2,
//        true residual via an extra HermOp), ConjugateGradientMixedPrec.h:71-167 (defect-correction restarts).
// The iteration is   d = <p, A p> ; a = c/d ; r -= a Ap ; cp = |r|^2 ; b = cp/c ; psi += a p ; p = b p + r
// with the norm fused into the r-update (axpy_norm) and the two axpys fused into one pass (ConjugateGradient.h:176-183).
#include "fermop.hpp"
#include <algorithm>
#include "kernels_common.cuh"
#include <cmath>
#include <cstdlib>
#include <functional>
#include <vector>

namespace gb {

// psi += a p ; p = b p + r    (ref: ConjugateGradient.h:176-183)
template <class V, class T> __global__ void cg_update_kernel(V *psi, V *p, const V *r, T a, T b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const V pv = p[i];
    psi[i] = vaxpy(a, pv, psi[i]);
    p[i] = vaxpy(b, pv, r[i]);
  }
}
static void cg_update(gb_context *ctx, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, double a, double b) {
  const int64_t n = psi->nvec();
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (psi->prec == GB_F32) cg_update_kernel<float4, float><<<blocks, 256, 0, ctx->stream>>>((float4 *)psi->data, (float4 *)p->data, (const float4 *)r->data, (float)a, (float)b, n);
  else cg_update_kernel<double2, double><<<blocks, 256, 0, ctx->stream>>>((double2 *)psi->data, (double2 *)p->data, (const double2 *)r->data, a, b, n);
  count_launch(ctx);
  check_launch(ctx, "cg_update");
}

static void chk(int rc) { if (rc != GB_OK) throw Error(rc, gb_last_error()); }

struct CGOut { int iters = 0; double true_resid = 0; bool converged = false; };
using HermOpFn = std::function<void(const gb_fermion *, gb_fermion *)>;

static CGOut cg_core(gb_context *ctx, const HermOpFn &A, const gb_fermion *src, gb_fermion *psi, double tol, int maxit) {
  GB_TRACE("ConjugateGradient");
  fermion_check_same(src, psi);
  gb_fermion *p = nullptr, *mmp = nullptr, *r = nullptr;
  auto mk = [&](gb_fermion **f) { *f = fermion_create_like(src, src->prec); };
  mk(&p); mk(&mmp); mk(&r);
  struct Guard { gb_fermion *a, *b, *c; ~Guard() { gb_fermion_destroy(a); gb_fermion_destroy(b); gb_fermion_destroy(c); } } guard{p, mmp, r};
  psi->cb = src->cb;
  CGOut out;
  double ssq, guess, a, cp, c, d, b, dd[2];
  chk(gb_norm2(src, &ssq));
  chk(gb_norm2(psi, &guess));
  GB_REQUIRE(!std::isnan(guess), "initial guess contains NaN");
  if (guess == 0.0) { chk(gb_copy(r, src)); chk(gb_copy(p, r)); a = ssq; }
  else {
    A(psi, mmp);
    chk(gb_axpy(r, -1.0, mmp, src));
    chk(gb_copy(p, r));
    chk(gb_norm2(p, &a));
  }
  cp = a;
  if (ssq == 0.0) { chk(gb_zero(psi)); out.iters = 1; out.true_resid = 0; out.converged = true; return out; }
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { out.true_resid = std::sqrt(a / ssq); out.iters = 0; out.converged = true; return out; }
  int k;
  for (k = 1; k <= maxit; k++) {
    c = cp;
    A(p, mmp);
    chk(gb_inner_product(p, mmp, dd));
    d = dd[0];
    a = c / d;
    chk(gb_axpy_norm(r, -a, mmp, r, &cp));
    b = cp / c;
    cg_update(ctx, psi, p, r, a, b);
    if (cp <= rsq) {
      A(psi, mmp);
      chk(gb_axpy(p, -1.0, src, mmp)); // p = mmp - src
      double rn;
      chk(gb_norm2(p, &rn));
      out.true_resid = std::sqrt(rn) / std::sqrt(ssq);
      out.iters = k; out.converged = true;
      return out;
    }
  }
  out.iters = k; out.converged = false;
  return out;
}

// fields.cu: device-resident reductions and updates
void reduce_inner_dev(gb_context *ctx, const gb_fermion *l, const gb_fermion *r, double *d_out);
void axpy_norm_dev(gb_context *ctx, gb_fermion *z, const gb_fermion *x, const gb_fermion *y, const double *d_c, const double *d_d, double *d_out);
void cg_update_dev(gb_context *ctx, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, const double *d_c, const double *d_d, const double *d_cp);

// Same algorithm and update order as cg_core (ref: ConjugateGradient.h:151-231), but the scalars live on the device:
//   d = <p,Ap> and cp = |r|^2 are reduced (and all-reduced over ranks) in-stream, a = c/d and b = cp/c are formed inside
//   the consuming kernels, and the next iteration's A p is enqueued before the host looks at cp, so the only host
//   involvement per iteration is one 8-byte read-back for the stopping test, off the critical path.  If that test says
//   "converged" the speculative A p has only overwritten the scratch field mmp.
// shift != 0: CG on HermOp + shift (ShiftedLinop, ref: ConjugateGradientMultiShiftMixedPrec.h:44-70)
static CGOut cg_schur_device_scalars(gb_fermop *op, const gb_fermion *src, gb_fermion *psi, double tol, int maxit, double shift = 0.0) {
  GB_TRACE("ConjugateGradient");
  gb_context *ctx = op->ctx;
  fermion_check_same(src, psi);
  gb_fermion *p = nullptr, *mmp = nullptr, *r = nullptr;
  auto mk = [&](gb_fermion **f) { *f = fermion_create_like(src, src->prec); };
  mk(&p); mk(&mmp); mk(&r);
  struct Guard { gb_fermion *a, *b, *c; ~Guard() { gb_fermion_destroy(a); gb_fermion_destroy(b); gb_fermion_destroy(c); } } guard{p, mmp, r};
  auto A = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op, GB_OP_HERMOP, in, out, 0); if (shift != 0.0) chk(gb_axpy(out, shift, in, out)); };
  psi->cb = src->cb;
  CGOut out;
  double ssq, guess, a;
  chk(gb_norm2(src, &ssq));
  chk(gb_norm2(psi, &guess));
  GB_REQUIRE(!std::isnan(guess), "initial guess contains NaN");
  if (guess == 0.0) { chk(gb_copy(r, src)); chk(gb_copy(p, r)); a = ssq; }
  else {
    A(psi, mmp);
    chk(gb_axpy(r, -1.0, mmp, src));
    chk(gb_copy(p, r));
    chk(gb_norm2(p, &a));
  }
  double cp = a;
  if (ssq == 0.0) { chk(gb_zero(psi)); out.iters = 1; out.true_resid = 0; out.converged = true; return out; }
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { out.true_resid = std::sqrt(a / ssq); out.iters = 0; out.converged = true; return out; }
  // device scalar slots: cv[0], cv[1] alternate as c (old |r|^2) and cp (new |r|^2); dd[0..1] = <p,Ap>
  double *cv = ctx->d_scalars, *dd = ctx->d_scalars + 2;
  GB_CUDA(cudaMemcpyAsync(cv, &cp, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  int k;
  const bool unfused = getenv("GB_CG_UNFUSED") != nullptr;   // read per solve: tests compare the two forms in one process
  if (shift == 0.0 && !unfused && cg_fused_available(op)) {
    // the linear algebra rides on the s-space passes of A p (fermop.cu: cg_fused_first / cg_fused_rest); same scalars, same
    // update order, same stopping test; mmp (= A p) is never materialised
    cg_fused_first(op, nullptr, p, nullptr, nullptr, nullptr, nullptr);
    for (k = 1; k <= maxit; k++) {
      double *d_c = cv + ((k - 1) & 1), *d_cp = cv + (k & 1);
      cg_fused_rest(op, p, r, d_c, dd, d_cp);            // d = <p, A p> ; r -= (c/d) A p ; cp = |r|^2
      GB_CUDA(cudaMemcpyAsync(ctx->h_result, d_cp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      GB_CUDA(cudaEventRecord(ctx->ev_scalar, ctx->stream));
      if (k < maxit) cg_fused_first(op, psi, p, r, d_c, dd, d_cp);   // psi += a p ; p = b p + r ; and the first pass of the next A p
      else cg_update_dev(ctx, psi, p, r, d_c, dd, d_cp);
      GB_CUDA(cudaEventSynchronize(ctx->ev_scalar));
      cp = ctx->h_result[0];
      GB_REQUIRE(!std::isnan(cp), "ConjugateGradient: residual is NaN");
      if (cp <= rsq) {
        A(psi, mmp);
        chk(gb_axpy(p, -1.0, src, mmp)); // p = mmp - src
        double rn;
        chk(gb_norm2(p, &rn));
        out.true_resid = std::sqrt(rn) / std::sqrt(ssq);
        out.iters = k; out.converged = true;
        return out;
      }
    }
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    out.iters = k; out.converged = false;
    return out;
  }
  A(p, mmp);
  for (k = 1; k <= maxit; k++) {
    double *d_c = cv + ((k - 1) & 1), *d_cp = cv + (k & 1);
    reduce_inner_dev(ctx, p, mmp, dd);                 // d = <p, A p>
    axpy_norm_dev(ctx, r, mmp, r, d_c, dd, d_cp);      // r -= (c/d) A p ; cp = |r|^2
    cg_update_dev(ctx, psi, p, r, d_c, dd, d_cp);      // psi += a p ; p = b p + r
    GB_CUDA(cudaMemcpyAsync(ctx->h_result, d_cp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaEventRecord(ctx->ev_scalar, ctx->stream));
    if (k < maxit) A(p, mmp);                          // speculative: needed unless this iteration converged
    GB_CUDA(cudaEventSynchronize(ctx->ev_scalar));
    cp = ctx->h_result[0];
    GB_REQUIRE(!std::isnan(cp), "ConjugateGradient: residual is NaN");
    if (cp <= rsq) {
      A(psi, mmp);
      chk(gb_axpy(p, -1.0, src, mmp)); // p = mmp - src
      double rn;
      chk(gb_norm2(p, &rn));
      out.true_resid = std::sqrt(rn) / std::sqrt(ssq);
      out.iters = k; out.converged = true;
      return out;
    }
  }
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  out.iters = k; out.converged = false;
  return out;
}

// =====================================================================================================
// ConjugateGradientMultiShift   ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:84-343
// One Krylov space for all shifts: (A + mass[s]) psi[s] = src, zero guess, primary shift = the lightest pole.  The
// recurrences for z_s, b_s and the stopping rule c z_s^2 < |src|^2 tol_s^2 are the reference's; the per-shift linear algebra
// is what the reference's own comments ask for (:213-218,:263-274): ONE launch updates every search direction (r is read
// once for all shifts) and ONE launch updates every solution.  GB_MS_UNFUSED=1 runs the same algorithm through the
// single-field BLAS entry points instead (tests compare the two bit for bit).
// =====================================================================================================
constexpr int MS_MAX = 12;   // fields per launch; more shifts go in chunks
struct MsEntries { void *y[MS_MAX]; const void *x[MS_MAX]; double a[MS_MAX], b[MS_MAX]; int plain[MS_MAX]; int n; };
// MODE 0 (search directions, shared r):  y_k = a_k y_k + r            (plain)   |   y_k = b_k r + a_k y_k
// MODE 1 (solutions):                    y_k = a_k x_k + y_k
template <class V, class T, int MODE> __global__ void ms_update_kernel(const MsEntries e, const V *__restrict__ r, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    V rv;
    if (MODE == 0) rv = r[i];
    for (int k = 0; k < e.n; k++) {
      V *y = (V *)e.y[k];
      if (MODE == 0) y[i] = e.plain[k] ? vaxpy((T)e.a[k], y[i], rv) : vaxpby((T)e.b[k], rv, (T)e.a[k], y[i]);
      else y[i] = vaxpy((T)e.a[k], ((const V *)e.x[k])[i], y[i]);
    }
  }
}
template <int MODE> static void ms_update(gb_context *ctx, const MsEntries &e, const gb_fermion *r, const gb_fermion *like) {
  if (e.n == 0) return;
  const int64_t n = like->nvec();
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (like->prec == GB_F32) ms_update_kernel<float4, float, MODE><<<blocks, 256, 0, ctx->stream>>>(e, r ? (const float4 *)r->data : nullptr, n);
  else ms_update_kernel<double2, double, MODE><<<blocks, 256, 0, ctx->stream>>>(e, r ? (const double2 *)r->data : nullptr, n);
  count_launch(ctx);
  check_launch(ctx, "ms_update");
}

struct MSOut { std::vector<int> iters; std::vector<double> true_resid; int iters_to_complete = 0; bool converged = false; };

static MSOut multishift_core(gb_context *ctx, const HermOpFn &A, const gb_fermion *src, gb_fermion *const *psi, int nshift, const double *mass,
                             const double *mresidual, int maxit) {
  static const bool unfused = getenv("GB_MS_UNFUSED") != nullptr;
  GB_REQUIRE(nshift >= 1, "ConjugateGradientMultiShift needs at least one shift");
  for (int s = 0; s < nshift; s++) {
    GB_REQUIRE(psi[s] != nullptr && psi[s] != src, "null or aliased result field");
    fermion_check_same(src, psi[s]);
    GB_REQUIRE(mass[s] >= mass[0], "the first pole must be the lightest (ref: ConjugateGradientMultiShift.h:125-128)");
  }
  MSOut R;
  R.iters.assign(nshift, 0); R.true_resid.assign(nshift, 0.0);
  std::vector<double> alpha(nshift, 1.0), bs(nshift), rsq(nshift), z0v(nshift, 1.0), z1v(nshift, 1.0);
  std::vector<int> converged(nshift, 0);
  std::vector<gb_fermion *> ps(nshift, nullptr);
  gb_fermion *r = nullptr, *p = nullptr, *tmp = nullptr, *mmp = nullptr;
  struct Guard {
    std::vector<gb_fermion *> &v; gb_fermion *&a, *&b, *&c, *&d;
    ~Guard() { for (auto *f : v) gb_fermion_destroy(f); gb_fermion_destroy(a); gb_fermion_destroy(b); gb_fermion_destroy(c); gb_fermion_destroy(d); }
  } guard{ps, r, p, tmp, mmp};
  for (int s = 0; s < nshift; s++) ps[s] = fermion_create_like(src, src->prec);
  r = fermion_create_like(src, src->prec); p = fermion_create_like(src, src->prec);
  tmp = fermion_create_like(src, src->prec); mmp = fermion_create_like(src, src->prec);
  double a, b, c, d, cp, bp, rn, dd[2];
  chk(gb_norm2(src, &cp));
  if (cp == 0.0) {   // ref :137-144
    for (int s = 0; s < nshift; s++) { chk(gb_zero(psi[s])); psi[s]->cb = src->cb; R.iters[s] = 1; }
    R.converged = true;
    return R;
  }
  for (int s = 0; s < nshift; s++) { rsq[s] = cp * mresidual[s] * mresidual[s]; chk(gb_copy(ps[s], src)); }
  chk(gb_copy(r, src)); chk(gb_copy(p, src));
  // z[s][iz] / z[s][1-iz] of the reference are z1v / z0v here ("current" / "previous")
  A(p, mmp);
  chk(gb_inner_product(p, mmp, dd)); d = dd[0];
  chk(gb_axpy(mmp, mass[0], p, mmp));
  chk(gb_norm2(p, &rn));
  d += rn * mass[0];
  b = -cp / d;
  bs[0] = b;
  for (int s = 1; s < nshift; s++) { z0v[s] = 1.0; z1v[s] = 1.0 / (1.0 - b * (mass[s] - mass[0])); bs[s] = b * z1v[s]; }
  chk(gb_axpy_norm(r, b, mmp, r, &c));
  for (int s = 0; s < nshift; s++) { chk(gb_axpby(psi[s], 0.0, -bs[s] * alpha[s], src, src)); psi[s]->cb = src->cb; }
  for (int k = 1; k <= maxit; k++) {
    a = c / cp;
    // search directions: p = a p + r ; ps[0] = a ps[0] + r ; ps[s] = z_s r + a_s ps[s]
    if (unfused) {
      chk(gb_axpy(p, a, p, r));
      for (int s = 0; s < nshift; s++) if (!converged[s]) {
        if (s == 0) chk(gb_axpy(ps[s], a, ps[s], r));
        else chk(gb_axpby(ps[s], z1v[s], a * z1v[s] * bs[s] / (z0v[s] * b), r, ps[s]));
      }
    } else {
      MsEntries e; e.n = 0;
      auto push = [&](gb_fermion *y, double ca, double cb_, int plain) {
        e.y[e.n] = y->data; e.x[e.n] = nullptr; e.a[e.n] = ca; e.b[e.n] = cb_; e.plain[e.n] = plain;
        if (++e.n == MS_MAX) { ms_update<0>(ctx, e, r, src); e.n = 0; }
      };
      push(p, a, 1.0, 1);
      for (int s = 0; s < nshift; s++) if (!converged[s]) {
        if (s == 0) push(ps[s], a, 1.0, 1);
        else push(ps[s], a * z1v[s] * bs[s] / (z0v[s] * b), z1v[s], 0);
      }
      ms_update<0>(ctx, e, r, src);
    }
    cp = c;
    A(p, mmp);
    chk(gb_inner_product(p, mmp, dd)); d = dd[0];
    chk(gb_axpy(mmp, mass[0], p, mmp));
    chk(gb_norm2(p, &rn));
    d += rn * mass[0];
    bp = b;
    b = -cp / d;
    chk(gb_axpy_norm(r, b, mmp, r, &c));
    GB_REQUIRE(!std::isnan(c), "ConjugateGradientMultiShift: residual is NaN");
    bs[0] = b;
    for (int s = 1; s < nshift; s++) if (!converged[s]) {   // toggle the recurrence history (ref :246-256)
      const double z0 = z1v[s], z1 = z0v[s];                // after the toggle: z0 = previous "current", z1 = the one before
      const double znew = z0 * z1 * bp / (b * a * (z1 - z0) + z1 * bp * (1 - (mass[s] - mass[0]) * b));
      z0v[s] = z0; z1v[s] = znew;
      bs[s] = b * znew / z0;
    }
    if (unfused) {
      for (int s = 0; s < nshift; s++) if (!converged[s]) chk(gb_axpy(psi[s], -bs[s] * alpha[s], ps[s], psi[s]));
    } else {
      MsEntries e; e.n = 0;
      for (int s = 0; s < nshift; s++) if (!converged[s]) {
        e.y[e.n] = psi[s]->data; e.x[e.n] = ps[s]->data; e.a[e.n] = -bs[s] * alpha[s]; e.b[e.n] = 0; e.plain[e.n] = 0;
        if (++e.n == MS_MAX) { ms_update<1>(ctx, e, nullptr, src); e.n = 0; }
      }
      ms_update<1>(ctx, e, nullptr, src);
    }
    bool all_converged = true;
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      R.iters[s] = k;
      const double zc = s == 0 ? 1.0 : z1v[s];
      if (c * zc * zc < rsq[s]) converged[s] = 1; else all_converged = false;
    }
    if (all_converged) {   // check the answers (ref :296-306)
      double cn;
      chk(gb_norm2(src, &cn));
      for (int s = 0; s < nshift; s++) {
        A(psi[s], mmp);
        chk(gb_axpy(tmp, mass[s], psi[s], mmp));
        chk(gb_axpy_norm(r, -alpha[s], src, tmp, &rn));
        R.true_resid[s] = std::sqrt(rn / cn);
      }
      R.iters_to_complete = k; R.converged = true;
      return R;
    }
  }
  R.iters_to_complete = maxit; R.converged = false;
  return R;
}

} // namespace gb

using namespace gb;

extern "C" {

int gb_cg_multishift_schur(gb_fermop *op, const gb_fermion *src, int nshift, const double *poles, const double *tolerances, int maxit,
                           gb_fermion *const *results, int *iters_out, double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(op && src && poles && tolerances && results, "null argument");
  GB_REQUIRE(src->kind == GB_HALF, "ConjugateGradientMultiShift on the Schur operator works on red-black fields");
  MSOut o = multishift_core(op->ctx, [&](const gb_fermion *in, gb_fermion *out) { op_apply(op, GB_OP_HERMOP, in, out, 0); }, src, results, nshift,
                            poles, tolerances, maxit);
  for (int s = 0; s < nshift; s++) { if (iters_out) iters_out[s] = o.iters[s]; if (true_resid_out) true_resid_out[s] = o.true_resid[s]; }
  if (iters_out) iters_out[nshift] = o.iters_to_complete;
  if (!o.converged) throw Error(GB_ERR_NOT_CONVERGED, "CG multi shift did not converge");
  GB_API_END
}

// ConjugateGradientReliableUpdate   ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:80-270
// fp32 iteration, fp64 reliable updates of solution and residual whenever |r|^2 has dropped by Delta since the last one, fp64
// clean-up CG at the end (DoFinalCleanup).  iters_out = {IterationsToComplete, ReliableUpdatesPerformed, IterationsToCleanup}.
int gb_relup_cg_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src, gb_fermion *psi, double tol, int maxit, double Delta,
                      int iters_out[3], double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(op_f && op_d && src && psi && src != psi, "null or aliased argument");
  GB_REQUIRE(op_f->prec == GB_F32 && op_d->prec == GB_F64, "reliable-update CG needs an fp32 and an fp64 operator");
  GB_REQUIRE(src->prec == GB_F64 && psi->prec == GB_F64 && src->kind == GB_HALF, "reliable-update CG works on fp64 red-black fields");
  GB_REQUIRE(Delta > 0. && Delta < 1., "Expect  0 < Delta < 1");   // ref :69
  fermion_check_same(src, psi);
  psi->cb = src->cb;
  auto Ad = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op_d, GB_OP_HERMOP, in, out, 0); };
  auto Af = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op_f, GB_OP_HERMOP, in, out, 0); };
  gb_fermion *p = fermion_create_like(src, GB_F64), *mmp = fermion_create_like(src, GB_F64), *r = fermion_create_like(src, GB_F64);
  gb_fermion *r_f = fermion_create_like(src, GB_F32), *psi_f = fermion_create_like(src, GB_F32), *p_f = fermion_create_like(src, GB_F32),
             *mmp_f = fermion_create_like(src, GB_F32);
  struct Guard { gb_fermion *f[7]; ~Guard() { for (auto *x : f) gb_fermion_destroy(x); } } guard{{p, mmp, r, r_f, psi_f, p_f, mmp_f}};
  int its[3] = {0, 0, 0};
  double true_resid = 0, cp, c, a, d, b, ssq, dd[2];
  auto finish = [&](bool converged) {
    if (iters_out) for (int i = 0; i < 3; i++) iters_out[i] = its[i];
    if (true_resid_out) *true_resid_out = true_resid;
    if (!converged) throw Error(GB_ERR_NOT_CONVERGED, "ConjugateGradientReliableUpdate did NOT converge");
  };
  Ad(psi, mmp);
  chk(gb_axpy(r, -1.0, mmp, src));            // r = src - mmp
  chk(gb_copy(p, r));
  chk(gb_norm2(p, &a)); cp = a;
  chk(gb_norm2(src, &ssq));
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { true_resid = std::sqrt(cp / ssq); finish(true); return GB_OK; }   // "guess was REALLY good" (ref :121-125)
  chk(gb_precision_change(r_f, r));
  chk(gb_zero(psi_f)); psi_f->cb = src->cb;
  chk(gb_copy(p_f, r_f));
  double MaxResidSinceLastRelUp = cp;
  int k, l = 0;
  for (k = 1; k <= maxit; k++) {
    c = cp;
    Af(p_f, mmp_f);
    chk(gb_inner_product(p_f, mmp_f, dd)); d = dd[0];
    a = c / d;
    chk(gb_axpy_norm(r_f, -a, mmp_f, r_f, &cp));
    GB_REQUIRE(!std::isnan(cp), "ConjugateGradientReliableUpdate: residual is NaN");
    b = cp / c;
    chk(gb_axpy(psi_f, a, p_f, psi_f));
    if (cp > MaxResidSinceLastRelUp) MaxResidSinceLastRelUp = cp;
    if (cp <= rsq) {
      chk(gb_precision_change(mmp, psi_f));
      chk(gb_axpy(psi, 1.0, mmp, psi));
      Ad(psi, mmp);
      double rn;
      chk(gb_axpy_norm(p, -1.0, src, mmp, &rn));          // p = mmp - src
      true_resid = std::sqrt(rn) / std::sqrt(ssq);
      its[0] = k; its[1] = l;
      CGOut fin = cg_schur_device_scalars(op_d, src, psi, tol, maxit);   // DoFinalCleanup (ref :190-197)
      its[2] = fin.iters;
      if (fin.iters > 0) true_resid = fin.true_resid;
      finish(fin.converged);
      return GB_OK;
    } else if (cp < Delta * MaxResidSinceLastRelUp) {    // reliable update (ref :203-228)
      chk(gb_precision_change(mmp, psi_f));
      chk(gb_axpy(psi, 1.0, mmp, psi));
      Ad(psi, mmp);
      chk(gb_axpy_norm(r, -1.0, mmp, src, &cp));          // r = src - mmp ; cp = |r|^2
      chk(gb_zero(psi_f));
      chk(gb_precision_change(r_f, r));
      MaxResidSinceLastRelUp = cp;
      b = cp / c;
      l++;
    }
    chk(gb_axpby(p_f, b, 1.0, p_f, r_f));                 // p_f = b p_f + r_f (after the update: ref :230)
  }
  its[0] = k; its[1] = l;
  finish(false);
  GB_API_END
}

int gb_cg_schur(gb_fermop *op, const gb_fermion *src, gb_fermion *sol, double tol, int maxit, int *iters_out, double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(op && src && sol, "null argument");
  CGOut o = cg_schur_device_scalars(op, src, sol, tol, maxit);
  if (iters_out) *iters_out = o.iters;
  if (true_resid_out) *true_resid_out = o.true_resid;
  if (!o.converged) throw Error(GB_ERR_NOT_CONVERGED, "ConjugateGradient did NOT converge");
  GB_API_END
}

int gb_cg(gb_context *ctx, gb_hermop_fn hermop, void *user, const gb_fermion *src, gb_fermion *sol, double tol, int maxit,
          int *iters_out, double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(ctx && hermop && src && sol, "null argument");
  CGOut o = cg_core(ctx, [&](const gb_fermion *in, gb_fermion *out) {
    int rc = hermop(user, in, out);
    if (rc != GB_OK) throw Error(rc, "user HermOp failed");
  }, src, sol, tol, maxit);
  if (iters_out) *iters_out = o.iters;
  if (true_resid_out) *true_resid_out = o.true_resid;
  if (!o.converged) throw Error(GB_ERR_NOT_CONVERGED, "ConjugateGradient did NOT converge");
  GB_API_END
}

struct MixedOut { int inner = 0, outer = 0, fin = 0; double true_resid = 0; bool converged = false; };
static MixedOut mixed_cg_core(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d_in, gb_fermion *sol_d, double tol, int max_inner, int max_outer,
                              double shift, double inner_tol0 = -1.0, double outer_loop_norm_mult = 100.0);
int gb_mixed_cg_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d_in, gb_fermion *sol_d, double tol, int max_inner,
                      int max_outer, int iters_out[3], double *true_resid_out) {
  GB_API_BEGIN
  MixedOut o = mixed_cg_core(op_f, op_d, src_d_in, sol_d, tol, max_inner, max_outer, 0.0);
  if (iters_out) { iters_out[0] = o.inner; iters_out[1] = o.outer; iters_out[2] = o.fin; }
  if (true_resid_out) *true_resid_out = o.true_resid;
  if (!o.converged) throw Error(GB_ERR_NOT_CONVERGED, "MixedPrecisionConjugateGradient final solve did NOT converge");
  GB_API_END
}
// the same with the class's public tuning members: InnerTolerance (<= 0: Tolerance) and OuterLoopNormMult   ref: ConjugateGradientMixedPrec.h:41-45
int gb_mixed_cg_schur_ex(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d_in, gb_fermion *sol_d, double tol, double inner_tol,
                         double outer_loop_norm_mult, int max_inner, int max_outer, int iters_out[3], double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(outer_loop_norm_mult > 0, "OuterLoopNormMult must be positive");
  MixedOut o = mixed_cg_core(op_f, op_d, src_d_in, sol_d, tol, max_inner, max_outer, 0.0, inner_tol, outer_loop_norm_mult);
  if (iters_out) { iters_out[0] = o.inner; iters_out[1] = o.outer; iters_out[2] = o.fin; }
  if (true_resid_out) *true_resid_out = o.true_resid;
  if (!o.converged) throw Error(GB_ERR_NOT_CONVERGED, "MixedPrecisionConjugateGradient final solve did NOT converge");
  GB_API_END
}
// MixedPrecisionConjugateGradientBatched   ref: ConjugateGradientMixedPrecBatched.h:79-207 (same order of operations)
int gb_mixed_cg_batched_schur(gb_fermop *op_f, gb_fermop *op_d, int nbatch, const gb_fermion *const *srcs_d, gb_fermion *const *sols_d,
                              double tol, int max_inner, int max_outer, int max_patchup, int update_residual, int *iters_out,
                              double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(op_f && op_d && srcs_d && sols_d && nbatch >= 1, "null argument");
  GB_TRACE("MixedPrecisionConjugateGradientBatched");
  GB_REQUIRE(op_f->prec == GB_F32 && op_d->prec == GB_F64, "mixed CG needs an fp32 and an fp64 operator");
  const int cb = srcs_d[0]->cb;
  std::vector<gb_fermion *> src_d(nbatch, nullptr), src_f(nbatch, nullptr), sol_f(nbatch, nullptr);
  gb_fermion *tmp_d = nullptr;
  struct Guard {
    std::vector<gb_fermion *> &a, &b, &c; gb_fermion *&t;
    ~Guard() { for (auto *f : a) gb_fermion_destroy(f); for (auto *f : b) gb_fermion_destroy(f); for (auto *f : c) gb_fermion_destroy(f); gb_fermion_destroy(t); }
  } guard{src_d, src_f, sol_f, tmp_d};
  std::vector<double> stop(nbatch), norm(nbatch, 0.0);
  for (int i = 0; i < nbatch; i++) {
    GB_REQUIRE(srcs_d[i] && sols_d[i] && srcs_d[i]->prec == GB_F64 && sols_d[i]->prec == GB_F64 && srcs_d[i]->kind == GB_HALF, "batched mixed CG works on fp64 red-black fields");
    GB_REQUIRE(srcs_d[i]->cb == cb, "the right-hand sides of a batch live on one checkerboard");
    fermion_check_same(srcs_d[0], srcs_d[i]); fermion_check_same(srcs_d[i], sols_d[i]);
    sols_d[i]->cb = cb;
    double n2;
    chk(gb_norm2(srcs_d[i], &n2));
    stop[i] = n2 * tol * tol;
    src_d[i] = fermion_create_like(srcs_d[i], GB_F64); src_f[i] = fermion_create_like(srcs_d[i], GB_F32); sol_f[i] = fermion_create_like(srcs_d[i], GB_F32);
    src_d[i]->cb = src_f[i]->cb = sol_f[i]->cb = cb;
    chk(gb_copy(src_d[i], srcs_d[i]));
  }
  tmp_d = fermion_create_like(srcs_d[0], GB_F64);
  tmp_d->cb = cb;
  const double OuterLoopNormMult = 100.0;
  double inner_tol = tol;
  std::vector<int> inner(nbatch, 0), fin(nbatch, 0);
  int outer;
  for (outer = 0; outer < max_outer; outer++) {
    bool all_converged = true;
    for (int i = 0; i < nbatch; i++) {
      op_apply(op_d, GB_OP_HERMOP, sols_d[i], tmp_d, 0);
      chk(gb_axpy_norm(src_d[i], -1.0, tmp_d, srcs_d[i], &norm[i]));       // src_d = residual
      chk(gb_precision_change(src_f[i], src_d[i]));
      chk(gb_zero(sol_f[i]));
      if (norm[i] > OuterLoopNormMult * stop[i]) all_converged = false;
    }
    if (all_converged) break;
    if (update_residual) {
      const double norm_max = *std::max_element(norm.begin(), norm.end()), stop_max = *std::max_element(stop.begin(), stop.end());
      while (norm_max * inner_tol * inner_tol < stop_max) inner_tol *= 2;
    }
    for (int i = 0; i < nbatch; i++) {
      CGOut in = cg_schur_device_scalars(op_f, src_f[i], sol_f[i], inner_tol, max_inner);   // ErrorOnNoConverge = false
      inner[i] += in.iters;
      chk(gb_precision_change(tmp_d, sol_f[i]));
      chk(gb_axpy(sols_d[i], 1.0, tmp_d, sols_d[i]));
    }
  }
  bool all_ok = true;
  for (int i = 0; i < nbatch; i++) {
    CGOut f = cg_schur_device_scalars(op_d, srcs_d[i], sols_d[i], tol, max_patchup);
    fin[i] = f.iters;
    if (true_resid_out) true_resid_out[i] = f.true_resid;
    all_ok = all_ok && f.converged;
  }
  if (iters_out) { iters_out[0] = outer; for (int i = 0; i < nbatch; i++) { iters_out[1 + i] = inner[i]; iters_out[1 + nbatch + i] = fin[i]; } }
  if (!all_ok) throw Error(GB_ERR_NOT_CONVERGED, "MixedPrecisionConjugateGradientBatched: a patch-up solve did NOT converge");
  GB_API_END
}
} // extern "C"
// MixedPrecisionConjugateGradient on HermOp (+ shift)   ref: ConjugateGradientMixedPrec.h:71-167
static MixedOut mixed_cg_core(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d_in, gb_fermion *sol_d, double tol, int max_inner, int max_outer,
                              double shift, double inner_tol0, double outer_loop_norm_mult) {
  GB_REQUIRE(op_f && op_d && src_d_in && sol_d, "null argument");
  GB_TRACE("MixedPrecisionConjugateGradient");
  GB_REQUIRE(op_f->prec == GB_F32 && op_d->prec == GB_F64, "mixed CG needs an fp32 and an fp64 operator");
  GB_REQUIRE(src_d_in->prec == GB_F64 && sol_d->prec == GB_F64, "mixed CG outer fields must be fp64");
  const int cb = src_d_in->cb;
  sol_d->cb = cb;
  gb_fermion *tmp_d = nullptr, *src_d = nullptr, *src_f = nullptr, *sol_f = nullptr;
  GB_REQUIRE(src_d_in->kind == GB_HALF, "mixed CG works on red-black fields");
  tmp_d = fermion_create_like(src_d_in, GB_F64);
  src_d = fermion_create_like(src_d_in, GB_F64);
  src_f = fermion_create_like(src_d_in, GB_F32);
  sol_f = fermion_create_like(src_d_in, GB_F32);
  struct Guard { gb_fermion *a, *b, *c, *d; ~Guard() { gb_fermion_destroy(a); gb_fermion_destroy(b); gb_fermion_destroy(c); gb_fermion_destroy(d); } } guard{tmp_d, src_d, src_f, sol_f};
  tmp_d->cb = src_d->cb = src_f->cb = sol_f->cb = cb;
  double src_norm;
  chk(gb_norm2(src_d_in, &src_norm));
  const double stop = src_norm * tol * tol;
  const double OuterLoopNormMult = outer_loop_norm_mult;     // public member of the reference's class, default 100 (ref :45,:65)
  double inner_tol = inner_tol0 > 0 ? inner_tol0 : tol;     // InnerTolerance, defaults to Tolerance (ref :41,:64)
  int total_inner = 0, outer;
  auto Ad = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op_d, GB_OP_HERMOP, in, out, 0); if (shift != 0.0) chk(gb_axpy(out, shift, in, out)); };
  chk(gb_copy(src_d, src_d_in));
  for (outer = 0; outer < max_outer; outer++) {
    Ad(sol_d, tmp_d);
    double norm;
    chk(gb_axpy_norm(src_d, -1.0, tmp_d, src_d_in, &norm));
    if (norm < OuterLoopNormMult * stop) break;
    while (norm * inner_tol * inner_tol < stop) inner_tol *= 2;
    chk(gb_precision_change(src_f, src_d));
    chk(gb_zero(sol_f));
    CGOut in = cg_schur_device_scalars(op_f, src_f, sol_f, inner_tol, max_inner, shift); // ErrorOnNoConverge = false
    total_inner += in.iters;
    chk(gb_precision_change(tmp_d, sol_f));
    chk(gb_axpy(sol_d, 1.0, tmp_d, sol_d));
  }
  CGOut fin = cg_schur_device_scalars(op_d, src_d_in, sol_d, tol, max_inner, shift);
  MixedOut R;
  R.inner = total_inner; R.outer = outer; R.fin = fin.iters; R.true_resid = fin.true_resid; R.converged = fin.converged;
  return R;
}

// ConjugateGradientMultiShiftMixedPrec   ref: Grid/algorithms/iterative/ConjugateGradientMultiShiftMixedPrec.h:128-410
// fp64 vectors and recurrences, fp32 operator per iteration, true-residual replacement every relup_freq iterations, clean-up of
// the shifts that miss their tolerance by MixedPrecisionConjugateGradient on HermOp + pole.  The multi-field updates are the fused
// kernels of the fp64 multishift solver.
extern "C" int gb_cg_multishift_mixed_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src, int nshift, const double *mass,
                                            const double *mresidual, int maxit, int relup_freq, gb_fermion *const *psi, int *iters_out,
                                            double *true_resid_out) {
  GB_API_BEGIN
  GB_REQUIRE(op_f && op_d && src && mass && mresidual && psi, "null argument");
  GB_REQUIRE(op_f->prec == GB_F32 && op_d->prec == GB_F64 && src->prec == GB_F64 && src->kind == GB_HALF, "needs an fp32 and an fp64 operator and fp64 red-black fields");
  GB_REQUIRE(nshift >= 1 && relup_freq >= 1, "bad shift count or reliable-update frequency");
  gb_context *ctx = op_d->ctx;
  for (int s = 0; s < nshift; s++) {
    GB_REQUIRE(psi[s] != nullptr && psi[s] != src, "null or aliased result field");
    fermion_check_same(src, psi[s]);
    GB_REQUIRE(mass[s] >= mass[0], "the first pole must be the lightest");
  }
  auto Ad = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op_d, GB_OP_HERMOP, in, out, 0); };
  auto Af = [&](const gb_fermion *in, gb_fermion *out) { op_apply(op_f, GB_OP_HERMOP, in, out, 0); };
  std::vector<double> alpha(nshift, 1.0), bs(nshift), rsq(nshift), z0v(nshift, 1.0), z1v(nshift, 1.0), tr(nshift, 0.0);
  std::vector<int> converged(nshift, 0), its(nshift, 0);
  std::vector<gb_fermion *> tmps;
  struct Guard { std::vector<gb_fermion *> &v; ~Guard() { for (auto *f : v) gb_fermion_destroy(f); } } guard{tmps};
  auto mk = [&](int prec) { gb_fermion *f = fermion_create_like(src, prec); tmps.push_back(f); return f; };
  std::vector<gb_fermion *> ps(nshift);
  for (int s = 0; s < nshift; s++) ps[s] = mk(GB_F64);
  gb_fermion *p = mk(GB_F64), *r = mk(GB_F64), *tmp = mk(GB_F64), *mmp = mk(GB_F64), *p_f = mk(GB_F32), *mmp_f = mk(GB_F32);
  auto report = [&](int k) {
    for (int s = 0; s < nshift; s++) { if (iters_out) iters_out[s] = its[s]; if (true_resid_out) true_resid_out[s] = tr[s]; }
    if (iters_out) iters_out[nshift] = k;
  };
  double a, b, c, d, cp, bp, rn, dd[2];
  chk(gb_norm2(src, &cp));
  if (cp == 0.0) {
    for (int s = 0; s < nshift; s++) { chk(gb_zero(psi[s])); psi[s]->cb = src->cb; its[s] = 1; }
    report(0);
    return GB_OK;
  }
  for (int s = 0; s < nshift; s++) { rsq[s] = cp * mresidual[s] * mresidual[s]; chk(gb_copy(ps[s], src)); }
  chk(gb_copy(p, src)); chk(gb_copy(r, src));
  Ad(p, mmp);
  chk(gb_inner_product(p, mmp, dd)); d = dd[0];
  chk(gb_axpy(mmp, mass[0], p, mmp));
  chk(gb_norm2(p, &rn));
  d += rn * mass[0];
  b = -cp / d;
  bs[0] = b;
  for (int s = 1; s < nshift; s++) { z0v[s] = 1.0; z1v[s] = 1.0 / (1.0 - b * (mass[s] - mass[0])); bs[s] = b * z1v[s]; }
  chk(gb_axpy_norm(r, b, mmp, r, &c));
  for (int s = 0; s < nshift; s++) { chk(gb_axpby(psi[s], 0.0, -bs[s] * alpha[s], src, src)); psi[s]->cb = src->cb; }
  for (int k = 1; k <= maxit; k++) {
    a = c / cp;
    {
      MsEntries e; e.n = 0;
      auto push = [&](gb_fermion *y, double ca, double cb_, int plain) {
        e.y[e.n] = y->data; e.x[e.n] = nullptr; e.a[e.n] = ca; e.b[e.n] = cb_; e.plain[e.n] = plain;
        if (++e.n == MS_MAX) { ms_update<0>(ctx, e, r, src); e.n = 0; }
      };
      push(p, a, 1.0, 1);
      for (int s = 0; s < nshift; s++) if (!converged[s]) {
        if (s == 0) push(ps[s], a, 1.0, 1);
        else push(ps[s], a * z1v[s] * bs[s] / (z0v[s] * b), z1v[s], 0);
      }
      ms_update<0>(ctx, e, r, src);
    }
    chk(gb_precision_change(p_f, p));
    cp = c;
    Af(p_f, mmp_f);
    chk(gb_precision_change(mmp, mmp_f));
    chk(gb_inner_product(p, mmp, dd)); d = dd[0];
    chk(gb_axpy(mmp, mass[0], p, mmp));
    chk(gb_norm2(p, &rn));
    d += rn * mass[0];
    bp = b;
    b = -cp / d;
    bs[0] = b;
    for (int s = 1; s < nshift; s++) if (!converged[s]) {
      const double z0 = z1v[s], z1 = z0v[s];
      const double znew = z0 * z1 * bp / (b * a * (z1 - z0) + z1 * bp * (1 - (mass[s] - mass[0]) * b));
      z0v[s] = z0; z1v[s] = znew;
      bs[s] = b * znew / z0;
    }
    {
      MsEntries e; e.n = 0;
      for (int s = 0; s < nshift; s++) if (!converged[s]) {
        e.y[e.n] = psi[s]->data; e.x[e.n] = ps[s]->data; e.a[e.n] = -bs[s] * alpha[s]; e.b[e.n] = 0; e.plain[e.n] = 0;
        if (++e.n == MS_MAX) { ms_update<1>(ctx, e, nullptr, src); e.n = 0; }
      }
      ms_update<1>(ctx, e, nullptr, src);
    }
    chk(gb_axpy_norm(r, b, mmp, r, &c));
    GB_REQUIRE(!std::isnan(c), "ConjugateGradientMultiShiftMixedPrec: residual is NaN");
    if (k % relup_freq == 0) {           // replace r with the true residual of the primary shift (ref :322-336)
      Ad(psi[0], mmp);
      chk(gb_axpy(mmp, mass[0], psi[0], mmp));
      chk(gb_axpy_norm(r, -1.0, mmp, src, &c));
    }
    bool all_converged = true;
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      its[s] = k;
      const double zc = s == 0 ? 1.0 : z1v[s];
      if (c * zc * zc < rsq[s]) converged[s] = 1; else all_converged = false;
    }
    if (all_converged || k == maxit - 1) {
      double cn;
      chk(gb_norm2(src, &cn));
      bool ok = true;
      for (int s = 0; s < nshift; s++) {
        Ad(psi[s], mmp);
        chk(gb_axpy(tmp, mass[s], psi[s], mmp));
        chk(gb_axpy_norm(r, -alpha[s], src, tmp, &rn));
        tr[s] = std::sqrt(rn / cn);
        if (rn >= rsq[s]) {              // clean-up (ref :382-396)
          MixedOut m = mixed_cg_core(op_f, op_d, src, psi[s], mresidual[s], 20000, 20000, mass[s]);
          tr[s] = m.true_resid;
          ok = ok && m.converged;
        }
      }
      report(k);
      if (!ok) throw Error(GB_ERR_NOT_CONVERGED, "ConjugateGradientMultiShiftMixedPrec: a clean-up solve did NOT converge");
      return GB_OK;
    }
  }
  report(maxit);
  throw Error(GB_ERR_NOT_CONVERGED, "ConjugateGradientMultiShiftMixedPrec did NOT converge");   // the reference asserts (:408)
  GB_API_END
}
