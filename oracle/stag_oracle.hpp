// =====================================================================================
// stag_oracle.hpp -- CPU ORACLE for the improved staggered (ASQTAD/HISQ-shaped, one-link + Naik three-link) hopping
// term.  TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as dirac_oracle.hpp).
//
// PARITY STATUS: pinned -- tests/test_oracle_vs_reference.py compares every entry with the reference's own
// ImprovedStaggeredFermion{F,D} compiled from /root/reference (oracle/Makefile.ref), and tests/golden holds its outputs.
//
// Restates (paths relative to the reference tree):
//   StaggeredImpl::DoubleStore                     Grid/qcd/action/fermion/StaggeredImpl.h:105-162
//       eta_x = 1, eta_y = (-1)^x, eta_z = (-1)^(x+y), eta_t = (-1)^(x+y+z)            (:121-129)
//       Uds[mu]   = eta_mu(x) Ufat_mu(x)        Uds[mu+4]   = eta_mu(x) Ufat_mu(x-mu)^dag            (:131-141)
//       UUUds[mu] = eta_mu(x) U(x)U(x+mu)U(x+2mu)   UUUds[mu+4] = eta_mu(x) [U(x-3mu)U(x-2mu)U(x-mu)]^dag  (:145-156, thin links)
//   ImprovedStaggeredFermion::ImportGauge          implementation/ImprovedStaggeredFermionImplementation.h:137-167
//       Uds[mu] *= 0.5 c1/u0 ; Uds[mu+4] *= -0.5 c1/u0 ; UUUds[mu] *= 0.5 c2/u0^3 ; UUUds[mu+4] *= -0.5 c2/u0^3
//   StaggeredKernels::DhopSiteGeneric              implementation/StaggeredKernelsImplementation.h:73-121
//       out(x) = sum_mu Uds[mu] in(x+mu) + Uds[mu+4] in(x-mu) + UUUds[mu] in(x+3mu) + UUUds[mu+4] in(x-3mu); dag: out = -out
//       stencil displacements {+1 x4, -1 x4, +3 x4, -3 x4}: instantiation/ImprovedStaggeredFermionInstantiation.cc:33-34
//   M = Dhop + mass, Mdag = -Dhop + mass, Mooee = mass, MooeeInv = 1/mass, Meooe = DhopEO|OE   (:173-247)
//   SchurStaggeredOperator::Mpc = mass^2 - Meooe Meooe ; HermOp = Mpc                 Grid/algorithms/LinearOperator.h:543-584
//   ConjugateGradient                                                                Grid/algorithms/iterative/ConjugateGradient.h:68-257
// Host layouts as in dirac_oracle.hpp; the staggered site object is chi[colour 3] complex (ColourVector).
// =====================================================================================
#pragma once
#include "dirac_oracle.hpp"

namespace oracle {

template <class T> struct ColourVector { cx<T> v[Nc]; };

template <class T> inline ColourMatrix<T> matmul(const ColourMatrix<T> &a, const ColourMatrix<T> &b) {
  ColourMatrix<T> c;
  for (int i = 0; i < Nc; i++) for (int j = 0; j < Nc; j++) {
    cx<T> s(0, 0);
    for (int k = 0; k < Nc; k++) s += a.m[i][k] * b.m[k][j];
    c.m[i][j] = s;
  }
  return c;
}
template <class T> inline ColourMatrix<T> adjoint(const ColourMatrix<T> &a) {
  ColourMatrix<T> c;
  for (int i = 0; i < Nc; i++) for (int j = 0; j < Nc; j++) c.m[i][j] = conj(a.m[j][i]);
  return c;
}
template <class T> inline ColourMatrix<T> scaled(const ColourMatrix<T> &a, T f) {
  ColourMatrix<T> c;
  for (int i = 0; i < Nc; i++) for (int j = 0; j < Nc; j++) c.m[i][j] = f * a.m[i][j];
  return c;
}
template <class T> inline void multAdd(ColourVector<T> &o, const ColourMatrix<T> &U, const ColourVector<T> &x) {
  for (int r = 0; r < Nc; r++) for (int c = 0; c < Nc; c++) o.v[r] += U.m[r][c] * x.v[c];
}

template <class T> struct StagOp {
  Geometry g;
  double mass = 0, c1 = 1, c2 = 1, u0 = 1;
  std::vector<ColourMatrix<T>> Uds, UUUds; // [V4][8]
  T mass_word() const { return (T)mass; }   // the working-precision type, for generic callers

  int64_t shifted(const int x[4], int mu, int d) const {
    int y[4] = {x[0], x[1], x[2], x[3]};
    y[mu] = ((x[mu] + d) % g.L[mu] + g.L[mu]) % g.L[mu];
    return g.lex4(y);
  }
  void importGauge(const ColourMatrix<T> *Uthin, const ColourMatrix<T> *Ufat) {
    const int64_t V = g.V4();
    Uds.assign(V * 8, ColourMatrix<T>()); UUUds.assign(V * 8, ColourMatrix<T>());
    std::vector<ColourMatrix<T>> UUU(V * 4);
#pragma omp parallel for
    for (int64_t i = 0; i < V; i++) {
      int x[4]; g.coor4(i, x);
      for (int mu = 0; mu < 4; mu++) {
        const ColourMatrix<T> &u0m = Uthin[i * 4 + mu], &u1m = Uthin[shifted(x, mu, 1) * 4 + mu], &u2m = Uthin[shifted(x, mu, 2) * 4 + mu];
        UUU[i * 4 + mu] = matmul(u0m, matmul(u1m, u2m)); // CovShiftForward(U,mu,CovShiftForward(U,mu,U))
      }
    }
    const T f1 = (T)(0.5 * c1 / u0), f3 = (T)(0.5 * c2 / u0 / u0 / u0);
#pragma omp parallel for
    for (int64_t i = 0; i < V; i++) {
      int x[4]; g.coor4(i, x);
      const int eta[4] = {1, (x[0] & 1) ? -1 : 1, ((x[0] + x[1]) & 1) ? -1 : 1, ((x[0] + x[1] + x[2]) & 1) ? -1 : 1};
      for (int mu = 0; mu < 4; mu++) {
        const T e = (T)eta[mu];
        Uds[i * 8 + mu] = scaled(scaled(Ufat[i * 4 + mu], e), f1);
        Uds[i * 8 + mu + 4] = scaled(scaled(adjoint(Ufat[shifted(x, mu, -1) * 4 + mu]), e), -f1);
        UUUds[i * 8 + mu] = scaled(scaled(UUU[i * 4 + mu], e), f3);
        UUUds[i * 8 + mu + 4] = scaled(scaled(adjoint(UUU[shifted(x, mu, -3) * 4 + mu]), e), -f3);
      }
    }
  }
  // one output site at coordinate x; idx(y) maps a neighbour coordinate to its index in `in`
  template <class Idx> ColourVector<T> site(const int x[4], const ColourVector<T> *in, int dag, Idx idx) const {
    const int64_t i4 = g.lex4(x);
    ColourVector<T> o; std::memset((void *)&o, 0, sizeof(o));
    for (int pass = 0; pass < 2; pass++) {
      const ColourMatrix<T> *U = (pass == 0 ? Uds.data() : UUUds.data()) + i4 * 8;
      const int d = pass == 0 ? 1 : 3;
      for (int mu = 0; mu < 4; mu++) { int y[4] = {x[0], x[1], x[2], x[3]}; y[mu] = (x[mu] + d) % g.L[mu]; multAdd(o, U[mu], in[idx(y)]); }
      for (int mu = 0; mu < 4; mu++) { int y[4] = {x[0], x[1], x[2], x[3]}; y[mu] = ((x[mu] - d) % g.L[mu] + g.L[mu]) % g.L[mu]; multAdd(o, U[mu + 4], in[idx(y)]); }
    }
    if (dag) for (int c = 0; c < Nc; c++) o.v[c] = -o.v[c];
    return o;
  }
  void Dhop(const ColourVector<T> *in, ColourVector<T> *out, int dag) const {
#pragma omp parallel for
    for (int64_t i = 0; i < g.V4(); i++) { int x[4]; g.coor4(i, x); out[i] = site(x, in, dag, [&](const int *y) { return g.lex4(y); }); }
  }
  // input parity 1-ocb, output parity ocb, both in checkerboard order
  void DhopCB(const ColourVector<T> *in, ColourVector<T> *out, int ocb, int dag) const {
#pragma omp parallel for
    for (int64_t ic = 0; ic < g.V4cb(); ic++) { int x[4]; g.cbcoor4(ic, ocb, x); out[ic] = site(x, in, dag, [&](const int *y) { return g.cb4(y); }); }
  }
  void axpby(int64_t n, ColourVector<T> *z, T a, T b, const ColourVector<T> *x, const ColourVector<T> *y) const {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) for (int c = 0; c < Nc; c++) z[i].v[c] = a * x[i].v[c] + b * y[i].v[c];
  }
  void M(const ColourVector<T> *in, ColourVector<T> *out) const { Dhop(in, out, 0); axpby(g.V4(), out, (T)mass, (T)1, in, out); }
  void Mdag(const ColourVector<T> *in, ColourVector<T> *out) const { Dhop(in, out, 1); axpby(g.V4(), out, (T)mass, (T)1, in, out); }
  void Meooe(const ColourVector<T> *in, ColourVector<T> *out, int cb_in, int dag) const { DhopCB(in, out, 1 - cb_in, dag); }
  void scale(int64_t n, ColourVector<T> *out, T a, const ColourVector<T> *in) const {
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) for (int c = 0; c < Nc; c++) out[i].v[c] = a * in[i].v[c];
  }
  void Mpc(const ColourVector<T> *in, ColourVector<T> *out, int cb) const {
    std::vector<ColourVector<T>> t1(g.V4cb()), t2(g.V4cb());
    Meooe(in, t1.data(), cb, 0);
    Meooe(t1.data(), t2.data(), 1 - cb, 0);
    axpby(g.V4cb(), out, (T)-1, (T)(mass * mass), t2.data(), in);
  }
};

template <class T> inline cx<double> stagInner(int64_t n, const ColourVector<T> *l, const ColourVector<T> *r) {
  double re = 0, im = 0;
#pragma omp parallel for reduction(+ : re, im)
  for (int64_t i = 0; i < n; i++) {
    cx<T> d(0, 0);
    for (int c = 0; c < Nc; c++) d += conj(l[i].v[c]) * r[i].v[c];
    re += (double)d.re; im += (double)d.im;
  }
  return cx<double>(re, im);
}

// ConjugateGradient on SchurStaggeredOperator (HermOp = Mpc), same update order as dirac_oracle.hpp's
template <class T>
CGResult StagConjugateGradient(const StagOp<T> &op, int cb, const ColourVector<T> *src, ColourVector<T> *psi, double tol, int maxit) {
  const int64_t n = op.g.V4cb();
  std::vector<ColourVector<T>> p(n), mmp(n), r(n);
  CGResult res;
  auto norm2 = [&](const ColourVector<T> *x) { return stagInner(n, x, x).re; };
  double ssq = norm2(src), guess = norm2(psi), a, cp, c, d, b;
  if (guess == 0.0) { std::copy(src, src + n, r.begin()); p = r; a = ssq; }
  else { op.Mpc(psi, mmp.data(), cb); op.axpby(n, r.data(), (T)-1, (T)1, mmp.data(), src); p = r; a = norm2(p.data()); }
  cp = a;
  if (ssq == 0.0) { std::memset((void *)psi, 0, sizeof(ColourVector<T>) * n); res.iterations = 1; res.converged = 1; return res; }
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { res.true_residual = std::sqrt(a / ssq); res.converged = 1; return res; }
  int k;
  for (k = 1; k <= maxit; k++) {
    c = cp;
    op.Mpc(p.data(), mmp.data(), cb);
    d = stagInner(n, p.data(), mmp.data()).re;
    a = c / d;
    op.axpby(n, r.data(), (T)(-a), (T)1, mmp.data(), r.data());
    cp = norm2(r.data());
    b = cp / c;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) for (int col = 0; col < Nc; col++) {
      psi[i].v[col] = (T)a * p[i].v[col] + psi[i].v[col];
      p[i].v[col] = (T)b * p[i].v[col] + r[i].v[col];
    }
    if (cp <= rsq) {
      op.Mpc(psi, mmp.data(), cb);
      op.axpby(n, p.data(), (T)1, (T)-1, mmp.data(), src);
      res.true_residual = std::sqrt(norm2(p.data())) / std::sqrt(ssq);
      res.iterations = k; res.converged = 1;
      return res;
    }
  }
  res.iterations = k; res.converged = 0;
  return res;
}

// SchurRedBlackStaggeredSolve pieces and the whole solve (ZeroGuesser)   ref: Grid/algorithms/iterative/SchurRedBlack.h:294-349
//   src_o' = Mooee (src_o - Meooe MooeeInv src_e) ; sol_e = MooeeInv (src_e - Meooe sol_o)
template <class T> void StagPick(const Geometry &g, int cb, ColourVector<T> *half, const ColourVector<T> *full) {
  for (int64_t i = 0; i < g.V4(); i++) { int x[4]; g.coor4(i, x); if (Geometry::parity(x) == cb) half[g.cb4(x)] = full[i]; }
}
template <class T> void StagSet(const Geometry &g, int cb, ColourVector<T> *full, const ColourVector<T> *half) {
  for (int64_t i = 0; i < g.V4(); i++) { int x[4]; g.coor4(i, x); if (Geometry::parity(x) == cb) full[i] = half[g.cb4(x)]; }
}
template <class T> void StagRedBlackSource(const StagOp<T> &op, const ColourVector<T> *src, ColourVector<T> *src_e, ColourVector<T> *src_o) {
  const int64_t n = op.g.V4cb();
  std::vector<ColourVector<T>> tmp(n), Mtmp(n), so(n);
  StagPick(op.g, 0, src_e, src); StagPick(op.g, 1, so.data(), src);
  op.scale(n, tmp.data(), (T)(1.0 / op.mass), src_e);
  op.Meooe(tmp.data(), Mtmp.data(), 0, 0);
  op.axpby(n, tmp.data(), (T)1, (T)-1, so.data(), Mtmp.data());
  op.scale(n, src_o, (T)op.mass, tmp.data());
}
template <class T> void StagRedBlackSolution(const StagOp<T> &op, const ColourVector<T> *sol_o, const ColourVector<T> *src_e, ColourVector<T> *sol) {
  const int64_t n = op.g.V4cb();
  std::vector<ColourVector<T>> tmp(n), sol_e(n);
  op.Meooe(sol_o, tmp.data(), 1, 0);
  op.axpby(n, tmp.data(), (T)1, (T)-1, src_e, tmp.data());
  op.scale(n, sol_e.data(), (T)(1.0 / op.mass), tmp.data());
  StagSet(op.g, 0, sol, sol_e.data()); StagSet(op.g, 1, sol, sol_o);
}
template <class T>
CGResult StagSchurSolve(const StagOp<T> &op, const ColourVector<T> *src, ColourVector<T> *sol, double tol, int maxit, double *unprec_resid) {
  const int64_t n = op.g.V4cb(), nf = op.g.V4();
  std::vector<ColourVector<T>> src_e(n), src_o(n), sol_o(n), r(nf);
  StagRedBlackSource(op, src, src_e.data(), src_o.data());
  std::memset((void *)sol_o.data(), 0, sizeof(ColourVector<T>) * n);
  CGResult res = StagConjugateGradient(op, 1, src_o.data(), sol_o.data(), tol, maxit);
  StagRedBlackSolution(op, sol_o.data(), src_e.data(), sol);
  op.M(sol, r.data());
  op.axpby(nf, r.data(), (T)1, (T)-1, r.data(), src);
  if (unprec_resid) *unprec_resid = std::sqrt(stagInner(nf, r.data(), r.data()).re / stagInner(nf, src, src).re);
  return res;
}

// independent naive form for the identity test (ref: tests/core/Test_staggered.cc:92-156 builds the same sum from
// Cshift / CovShift of the ORIGINAL links): phases and coefficients applied on the fly, no double store
template <class T>
void StagDhopNaive(const Geometry &g, const ColourMatrix<T> *Uthin, const ColourMatrix<T> *Ufat, double c1, double c2, double u0,
                   const ColourVector<T> *in, ColourVector<T> *out, int dag) {
  auto sh = [&](const int x[4], int mu, int d) { int y[4] = {x[0], x[1], x[2], x[3]}; y[mu] = ((x[mu] + d) % g.L[mu] + g.L[mu]) % g.L[mu]; return g.lex4(y); };
#pragma omp parallel for
  for (int64_t i = 0; i < g.V4(); i++) {
    int x[4]; g.coor4(i, x);
    const int eta[4] = {1, (x[0] & 1) ? -1 : 1, ((x[0] + x[1]) & 1) ? -1 : 1, ((x[0] + x[1] + x[2]) & 1) ? -1 : 1};
    ColourVector<T> o; std::memset((void *)&o, 0, sizeof(o));
    for (int mu = 0; mu < 4; mu++) {
      ColourVector<T> acc; std::memset((void *)&acc, 0, sizeof(acc));
      // one link: c1/u0 * 1/2 [ U(x) chi(x+mu) - U(x-mu)^dag chi(x-mu) ]
      ColourVector<T> t; std::memset((void *)&t, 0, sizeof(t));
      multAdd(t, Ufat[i * 4 + mu], in[sh(x, mu, 1)]);
      ColourVector<T> m; std::memset((void *)&m, 0, sizeof(m));
      multAdd(m, adjoint(Ufat[sh(x, mu, -1) * 4 + mu]), in[sh(x, mu, -1)]);
      for (int c = 0; c < Nc; c++) acc.v[c] = (T)(0.5 * c1 / u0) * (t.v[c] - m.v[c]);
      // three link
      ColourMatrix<T> fwd = matmul(Uthin[i * 4 + mu], matmul(Uthin[sh(x, mu, 1) * 4 + mu], Uthin[sh(x, mu, 2) * 4 + mu]));
      ColourMatrix<T> bwd = adjoint(matmul(Uthin[sh(x, mu, -3) * 4 + mu], matmul(Uthin[sh(x, mu, -2) * 4 + mu], Uthin[sh(x, mu, -1) * 4 + mu])));
      std::memset((void *)&t, 0, sizeof(t)); std::memset((void *)&m, 0, sizeof(m));
      multAdd(t, fwd, in[sh(x, mu, 3)]); multAdd(m, bwd, in[sh(x, mu, -3)]);
      for (int c = 0; c < Nc; c++) acc.v[c] += (T)(0.5 * c2 / (u0 * u0 * u0)) * (t.v[c] - m.v[c]);
      for (int c = 0; c < Nc; c++) o.v[c] += (T)eta[mu] * acc.v[c];
    }
    if (dag) for (int c = 0; c < Nc; c++) o.v[c] = -o.v[c];
    out[i] = o;
  }
}

} // namespace oracle
