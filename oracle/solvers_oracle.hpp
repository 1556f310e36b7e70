// solvers_oracle.hpp -- CPU ORACLE restatement of the solver variants of SURVEY 8(f) row 3 that share the hot path's kernels.
// TEST INFRASTRUCTURE ONLY (same rule as dirac_oracle.hpp): never linked into the product.
//
//   ConjugateGradientMultiShift   ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:84-343
//   (shifts as MultiShiftFunction{order, poles, residues, tolerances, norm}, ref: Grid/algorithms/approx/MultiShiftFunction.h:34-44)
//
// Generic over the site type (Spinor<T> / ColourVector<T>): the lattice algebra is found by overload (dirac_oracle.hpp,
// and the ColourVector overloads below).
#pragma once
#include "dirac_oracle.hpp"
#include "stag_oracle.hpp"
#include <array>

namespace oracle {

// ---- ColourVector overloads of the lattice algebra (Spinor versions: dirac_oracle.hpp)
template <class T> cx<double> innerProduct(int64_t n, const ColourVector<T> *l, const ColourVector<T> *r) { return stagInner(n, l, r); }
template <class T> double norm2(int64_t n, const ColourVector<T> *x) { return stagInner(n, x, x).re; }
template <class T> void axpy(int64_t n, ColourVector<T> *z, T a, const ColourVector<T> *x, const ColourVector<T> *y) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++) for (int c = 0; c < Nc; c++) z[i].v[c] = a * x[i].v[c] + y[i].v[c];
}
template <class T> void axpby(int64_t n, ColourVector<T> *z, T a, T b, const ColourVector<T> *x, const ColourVector<T> *y) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++) for (int c = 0; c < Nc; c++) z[i].v[c] = a * x[i].v[c] + b * y[i].v[c];
}
template <class T> double axpy_norm(int64_t n, ColourVector<T> *z, T a, const ColourVector<T> *x, const ColourVector<T> *y) {
  axpy(n, z, a, x, y);
  return norm2(n, z);
}

struct MultiShiftResult {
  std::vector<int> iterations;          // IterationsToCompleteShift
  std::vector<double> true_residual;    // TrueResidualShift
  int iterations_to_complete = 0;       // IterationsToComplete
  int converged = 0;
};

// Solves (A + poles[s]) psi[s] = src for every shift with ONE Krylov space; A(in, out) is the Hermitian operator (MdagM).
// Same recurrences, update order and stopping rule as the reference (:158-343): zero guess, primary shift = the lightest pole,
// shift s stops when c z_s^2 < |src|^2 tol_s^2.
template <class T, class F, class HermOpFn>
MultiShiftResult MultiShiftCG(HermOpFn &&A, int64_t n, const F *src, std::vector<F *> psi, const std::vector<double> &mass,
                              const std::vector<double> &mresidual, int maxit) {
  const int nshift = (int)mass.size();
  MultiShiftResult R;
  R.iterations.assign(nshift, 0); R.true_residual.assign(nshift, 0.0);
  std::vector<double> alpha(nshift, 1.0), bs(nshift), rsq(nshift);
  std::vector<std::array<double, 2>> z(nshift);
  std::vector<int> converged(nshift, 0);
  std::vector<std::vector<F>> ps(nshift, std::vector<F>(n));
  std::vector<F> r(src, src + n), p(src, src + n), tmp(n), mmp(n);
  double a, b, c, d, cp, bp;
  cp = norm2(n, src);
  if (cp == 0.0) {
    for (int s = 0; s < nshift; s++) { std::memset((void *)psi[s], 0, sizeof(F) * n); R.iterations[s] = 1; }
    R.converged = 1;
    return R;
  }
  for (int s = 0; s < nshift; s++) { rsq[s] = cp * mresidual[s] * mresidual[s]; std::copy(src, src + n, ps[s].begin()); }
  A(p.data(), mmp.data());
  d = innerProduct(n, p.data(), mmp.data()).re;               // HermOpAndNorm: d = <p, A p>
  axpy(n, mmp.data(), (T)mass[0], p.data(), mmp.data());
  double rn = norm2(n, p.data());
  d += rn * mass[0];
  b = -cp / d;
  int iz = 0;
  z[0][1 - iz] = 1.0; z[0][iz] = 1.0; bs[0] = b;
  for (int s = 1; s < nshift; s++) { z[s][1 - iz] = 1.0; z[s][iz] = 1.0 / (1.0 - b * (mass[s] - mass[0])); bs[s] = b * z[s][iz]; }
  c = axpy_norm(n, r.data(), (T)b, mmp.data(), r.data());
  for (int s = 0; s < nshift; s++) axpby(n, psi[s], (T)0, (T)(-bs[s] * alpha[s]), src, src);
  for (int k = 1; k <= maxit; k++) {
    a = c / cp;
    axpy(n, p.data(), (T)a, p.data(), r.data());
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      if (s == 0) axpy(n, ps[s].data(), (T)a, ps[s].data(), r.data());
      else { const double as = a * z[s][iz] * bs[s] / (z[s][1 - iz] * b); axpby(n, ps[s].data(), (T)z[s][iz], (T)as, r.data(), ps[s].data()); }
    }
    cp = c;
    A(p.data(), mmp.data());
    d = innerProduct(n, p.data(), mmp.data()).re;
    axpy(n, mmp.data(), (T)mass[0], p.data(), mmp.data());
    rn = norm2(n, p.data());
    d += rn * mass[0];
    bp = b;
    b = -cp / d;
    c = axpy_norm(n, r.data(), (T)b, mmp.data(), r.data());
    bs[0] = b;
    iz = 1 - iz;
    for (int s = 1; s < nshift; s++) if (!converged[s]) {
      const double z0 = z[s][1 - iz], z1 = z[s][iz];
      z[s][iz] = z0 * z1 * bp / (b * a * (z1 - z0) + z1 * bp * (1 - (mass[s] - mass[0]) * b));
      bs[s] = b * z[s][iz] / z0;
    }
    for (int s = 0; s < nshift; s++) if (!converged[s]) axpy(n, psi[s], (T)(-bs[s] * alpha[s]), ps[s].data(), psi[s]);
    int all_converged = 1;
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      R.iterations[s] = k;
      const double css = c * z[s][iz] * z[s][iz];
      if (css < rsq[s]) converged[s] = 1; else all_converged = 0;
    }
    if (all_converged) {
      for (int s = 0; s < nshift; s++) {
        A(psi[s], mmp.data());
        axpy(n, tmp.data(), (T)mass[s], psi[s], mmp.data());
        axpy(n, r.data(), (T)(-alpha[s]), src, tmp.data());
        R.true_residual[s] = std::sqrt(norm2(n, r.data()) / norm2(n, src));
      }
      R.iterations_to_complete = k; R.converged = 1;
      return R;
    }
  }
  R.iterations_to_complete = maxit; R.converged = 0;   // "CG multi shift did not converge" (the reference only logs it, :336-338)
  return R;
}

// ConjugateGradientReliableUpdate   ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:80-270
// fp32 CG iteration on (r_f, p_f, psi_f); whenever the residual has dropped by Delta relative to its maximum since the last
// update, psi += psi_f in fp64, the true residual is recomputed with the fp64 operator and the fp32 iteration restarts from
// it (keeping the search direction).  On convergence: final accumulation, true residual, then a fp64 clean-up CG.
struct RelUpResult { int iterations = 0, reliable_updates = 0, cleanup_iterations = 0; double true_residual = 0; int converged = 0; };
inline RelUpResult ReliableUpdateCG(const FermOp<double> &op_d, const FermOp<float> &op_f, int cb, const Spinor<double> *src, Spinor<double> *psi,
                                    double tol, int maxit, double Delta) {
  const int64_t n = op_d.V5cb();
  RelUpResult R;
  std::vector<Spinor<double>> p(n), mmp(n), r(n);
  double cp, c, a, d, b, ssq;
  op_d.HermOp(psi, mmp.data(), cb);
  axpy(n, r.data(), -1.0, mmp.data(), src);               // r = src - mmp
  p = r;
  a = norm2(n, p.data()); cp = a; ssq = norm2(n, src);
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { R.converged = 1; R.true_residual = std::sqrt(cp / ssq); return R; }
  std::vector<Spinor<float>> r_f(n), psi_f(n), p_f(n), mmp_f(n);
  precisionChange(n, r_f.data(), r.data());
  std::memset((void *)psi_f.data(), 0, sizeof(Spinor<float>) * n);
  p_f = r_f;
  double MaxResidSinceLastRelUp = cp;
  int k, l = 0;
  for (k = 1; k <= maxit; k++) {
    c = cp;
    op_f.HermOp(p_f.data(), mmp_f.data(), cb);
    d = innerProduct(n, p_f.data(), mmp_f.data()).re;
    a = c / d;
    cp = axpy_norm(n, r_f.data(), (float)(-a), mmp_f.data(), r_f.data());
    b = cp / c;
    axpy(n, psi_f.data(), (float)a, p_f.data(), psi_f.data());
    if (cp > MaxResidSinceLastRelUp) MaxResidSinceLastRelUp = cp;
    if (cp <= rsq) {
      precisionChange(n, mmp.data(), psi_f.data());
      axpy(n, psi, 1.0, mmp.data(), psi);
      op_d.HermOp(psi, mmp.data(), cb);
      axpy(n, p.data(), -1.0, src, mmp.data());
      R.true_residual = std::sqrt(norm2(n, p.data())) / std::sqrt(ssq);
      R.iterations = k; R.reliable_updates = l;
      CGResult fin = ConjugateGradient(op_d, cb, src, psi, tol, maxit);   // DoFinalCleanup
      R.cleanup_iterations = fin.iterations; R.converged = fin.converged;
      if (fin.iterations > 0) R.true_residual = fin.true_residual;
      return R;
    } else if (cp < Delta * MaxResidSinceLastRelUp) {
      precisionChange(n, mmp.data(), psi_f.data());
      axpy(n, psi, 1.0, mmp.data(), psi);
      op_d.HermOp(psi, mmp.data(), cb);
      axpy(n, r.data(), -1.0, mmp.data(), src);
      std::memset((void *)psi_f.data(), 0, sizeof(Spinor<float>) * n);
      precisionChange(n, r_f.data(), r.data());
      cp = norm2(n, r.data());
      MaxResidSinceLastRelUp = cp;
      b = cp / c;
      l++;
    }
    axpby(n, p_f.data(), (float)b, 1.0f, p_f.data(), r_f.data());        // p_f = b p_f + r_f
  }
  R.iterations = k; R.reliable_updates = l; R.converged = 0;
  return R;
}

// ConjugateGradientMultiShiftMixedPrec   ref: Grid/algorithms/iterative/ConjugateGradientMultiShiftMixedPrec.h:128-410
// The multishift recurrences with fp64 vectors and an fp32 operator application per iteration; every ReliableUpdateFreq
// iterations the residual is replaced by the true fp64 residual of the primary shift; each shift that misses its tolerance at
// the end is cleaned up with MixedPrecisionConjugateGradient on HermOp + pole.
struct MultiShiftMixedResult { std::vector<int> iterations; std::vector<double> true_residual; int iterations_to_complete = 0; int cleanups = 0; };
inline MultiShiftMixedResult MultiShiftMixedPrecCG(const FermOp<double> &op_d, const FermOp<float> &op_f, int cb, const Spinor<double> *src,
                                                   std::vector<Spinor<double> *> psi, const std::vector<double> &mass,
                                                   const std::vector<double> &mresidual, int maxit, int relup_freq) {
  typedef Spinor<double> FD; typedef Spinor<float> FF;
  const int64_t n = op_d.V5cb();
  const int nshift = (int)mass.size();
  MultiShiftMixedResult R;
  R.iterations.assign(nshift, 0); R.true_residual.assign(nshift, 0.0);
  std::vector<double> alpha(nshift, 1.0), bs(nshift), rsq(nshift);
  std::vector<std::array<double, 2>> z(nshift);
  std::vector<int> converged(nshift, 0);
  std::vector<std::vector<FD>> ps(nshift, std::vector<FD>(n));
  std::vector<FD> p(src, src + n), r(src, src + n), tmp(n), mmp(n);
  std::vector<FF> p_f(n), mmp_f(n);
  double a, b, c, d, cp, bp;
  cp = norm2(n, src);
  if (cp == 0.0) { for (int s = 0; s < nshift; s++) { std::memset((void *)psi[s], 0, sizeof(FD) * n); R.iterations[s] = 1; } return R; }
  for (int s = 0; s < nshift; s++) { rsq[s] = cp * mresidual[s] * mresidual[s]; std::copy(src, src + n, ps[s].begin()); }
  op_d.HermOp(p.data(), mmp.data(), cb);                    // the reference also applies Linop_f here, only to compare (:203-209)
  d = innerProduct(n, p.data(), mmp.data()).re;
  axpy(n, mmp.data(), mass[0], p.data(), mmp.data());
  double rn = norm2(n, p.data());
  d += rn * mass[0];
  b = -cp / d;
  int iz = 0;
  z[0][1 - iz] = 1.0; z[0][iz] = 1.0; bs[0] = b;
  for (int s = 1; s < nshift; s++) { z[s][1 - iz] = 1.0; z[s][iz] = 1.0 / (1.0 - b * (mass[s] - mass[0])); bs[s] = b * z[s][iz]; }
  c = axpy_norm(n, r.data(), b, mmp.data(), r.data());
  for (int s = 0; s < nshift; s++) axpby(n, psi[s], 0.0, -bs[s] * alpha[s], src, src);
  for (int k = 1; k <= maxit; k++) {
    a = c / cp;
    axpy(n, p.data(), a, p.data(), r.data());
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      if (s == 0) axpy(n, ps[s].data(), a, ps[s].data(), r.data());
      else { const double as = a * z[s][iz] * bs[s] / (z[s][1 - iz] * b); axpby(n, ps[s].data(), z[s][iz], as, r.data(), ps[s].data()); }
    }
    precisionChange(n, p_f.data(), p.data());
    cp = c;
    op_f.HermOp(p_f.data(), mmp_f.data(), cb);
    precisionChange(n, mmp.data(), mmp_f.data());
    d = innerProduct(n, p.data(), mmp.data()).re;
    axpy(n, mmp.data(), mass[0], p.data(), mmp.data());
    rn = norm2(n, p.data());
    d += rn * mass[0];
    bp = b;
    b = -cp / d;
    bs[0] = b;
    iz = 1 - iz;
    for (int s = 1; s < nshift; s++) if (!converged[s]) {
      const double z0 = z[s][1 - iz], z1 = z[s][iz];
      z[s][iz] = z0 * z1 * bp / (b * a * (z1 - z0) + z1 * bp * (1 - (mass[s] - mass[0]) * b));
      bs[s] = b * z[s][iz] / z0;
    }
    for (int s = 0; s < nshift; s++) if (!converged[s]) axpy(n, psi[s], -bs[s] * alpha[s], ps[s].data(), psi[s]);
    c = axpy_norm(n, r.data(), b, mmp.data(), r.data());
    if (k % relup_freq == 0) {           // replace r with the true residual of the primary shift
      op_d.HermOp(psi[0], mmp.data(), cb);
      axpy(n, mmp.data(), mass[0], psi[0], mmp.data());
      c = axpy_norm(n, r.data(), -1.0, mmp.data(), src);
    }
    int all_converged = 1;
    for (int s = 0; s < nshift; s++) if (!converged[s]) {
      R.iterations[s] = k;
      if (c * z[s][iz] * z[s][iz] < rsq[s]) converged[s] = 1; else all_converged = 0;
    }
    if (all_converged || k == maxit - 1) {
      const double cn = norm2(n, src);
      for (int s = 0; s < nshift; s++) {
        op_d.HermOp(psi[s], mmp.data(), cb);
        axpy(n, tmp.data(), mass[s], psi[s], mmp.data());
        axpy(n, r.data(), -alpha[s], src, tmp.data());
        rn = norm2(n, r.data());
        R.true_residual[s] = std::sqrt(rn / cn);
        if (rn >= rsq[s]) {              // clean up with mixed-precision CG on HermOp + pole (:382-396)
          MixedCGResult m = MixedPrecisionCG(op_d, op_f, cb, src, psi[s], mresidual[s], 20000, 20000, -1.0, mass[s]);
          R.true_residual[s] = m.true_residual;
          R.cleanups++;
        }
      }
      R.iterations_to_complete = k;
      return R;
    }
  }
  R.iterations_to_complete = maxit;      // the reference asserts here (:408)
  return R;
}

} // namespace oracle
