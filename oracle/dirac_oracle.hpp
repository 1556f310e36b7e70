// =====================================================================================
// dirac_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17/OpenMP restatement of the arithmetic of paboyle/Grid's Wilson / domain-wall
// hopping term and even-odd Schur CG.  It exists only so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs can check and time the CUDA path against
// an independent CPU implementation.  Nothing under grid_b200/ may include, link or call it.
//
// PARITY STATUS: PINNED against the reference itself.  oracle/Makefile.ref compiles the unmodified reference (CPU, AVX2,
// OpenMP, comms none) from /root/reference into oracle/_ref/libgridref.so; tests/test_oracle_vs_reference.py compares every
// entry point of this file with the reference's own WilsonFermion / DomainWallFermion / MobiusFermion (fp64 agreement to
// rounding, same CG and mixed-CG iteration counts), and tests/golden/dirac_golden.npz stores reference outputs generated
// from that library (tests/golden/make_golden.py) for machines without it.  The reference's executable identities
// (Dhop == naive Cshift form, Deo+Doe == D, adjointness, MooeeInv*Mooee == 1, Hermiticity, free-field plane waves) are
// kept as a second, independent pin in tests/test_oracle_identities.py.
//
// All "ref:" citations are paths relative to the reference tree.
//
// Host layouts (identical to the C-ABI import/export layouts of include/gridb200.h):
//   4D lexicographic site index   i4 = x + Lx*(y + Ly*(z + Lz*t))      ref: Grid/util/Lexicographic.h:18-27
//   5D lexicographic site index   i5 = s + Ls*i4  (s is coordinate 0)  ref: Grid/qcd/utils/SpaceTimeGrid.cc:49-63
//   fermion site object           psi[spin 4][colour 3] complex        ref: Grid/qcd/QCD.h (SpinColourVector)
//   gauge site object             U[mu 4][row 3][col 3] complex        ref: Grid/qcd/QCD.h:106 (LorentzColourMatrix)
//   checkerboarded index          icb = s + Ls*((x>>1) + (Lx/2)*(y + Ly*(z + Lz*t)))
//                                                                      ref: Grid/cartesian/Cartesian_red_black.h:271-286
//   parity                        (x+y+z+t)&1, Even=0, Odd=1, s ignored ref: Cartesian_red_black.h:68-76
// =====================================================================================
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace oracle {

// ------------------------------------------------------------------ complex scalar
template <class T> struct cx {
  T re, im;
  cx() = default;
  constexpr cx(T r, T i = T(0)) : re(r), im(i) {}
};
template <class T> inline cx<T> operator+(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <class T> inline cx<T> operator-(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <class T> inline cx<T> operator-(cx<T> a) { return {-a.re, -a.im}; }
template <class T> inline cx<T> operator*(cx<T> a, cx<T> b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <class T> inline cx<T> operator*(T a, cx<T> b) { return {a * b.re, a * b.im}; }
template <class T> inline cx<T> conj(cx<T> a) { return {a.re, -a.im}; }
template <class T> inline cx<T> timesI(cx<T> a) { return {-a.im, a.re}; }
template <class T> inline cx<T> timesMinusI(cx<T> a) { return {a.im, -a.re}; }
template <class T> inline cx<T> &operator+=(cx<T> &a, cx<T> b) { a.re += b.re; a.im += b.im; return a; }
template <class T> inline cx<T> &operator-=(cx<T> &a, cx<T> b) { a.re -= b.re; a.im -= b.im; return a; }

constexpr int Nc = 3, Ns = 4, Nhs = 2, Nd = 4;
// stencil point ids, ref: Grid/qcd/QCD.h:36-48
enum { Xp = 0, Yp = 1, Zp = 2, Tp = 3, Xm = 4, Ym = 5, Zm = 6, Tm = 7 };
enum { Even = 0, Odd = 1 };
enum { DaggerNo = 0, DaggerYes = 1 };

template <class T> struct Spinor { cx<T> v[Ns][Nc]; };
template <class T> struct HalfSpinor { cx<T> v[Nhs][Nc]; };
template <class T> struct ColourMatrix { cx<T> m[Nc][Nc]; };

template <class T> inline void zero(Spinor<T> &a) { std::memset(&a, 0, sizeof(a)); }

// ------------------------------------------------------------------ spin projection / reconstruction
// ref: Grid/qcd/spin/TwoSpinor.h:75-133 (spProj{X,Y,Z,T}{p,m}) and :193-354 (accumRecon*)
// dir: 0..3 = X,Y,Z,T ; sign=+1 -> (1+gamma_mu), sign=-1 -> (1-gamma_mu)
template <class T> inline void spProj(HalfSpinor<T> &h, const Spinor<T> &f, int dir, int sign) {
  for (int c = 0; c < Nc; c++) {
    const cx<T> f0 = f.v[0][c], f1 = f.v[1][c], f2 = f.v[2][c], f3 = f.v[3][c];
    switch (dir) {
    case 0: // X : h0 = f0 +- i f3 ; h1 = f1 +- i f2
      if (sign > 0) { h.v[0][c] = f0 + timesI(f3); h.v[1][c] = f1 + timesI(f2); }
      else          { h.v[0][c] = f0 - timesI(f3); h.v[1][c] = f1 - timesI(f2); }
      break;
    case 1: // Y : h0 = f0 -+ f3 ; h1 = f1 +- f2
      if (sign > 0) { h.v[0][c] = f0 - f3; h.v[1][c] = f1 + f2; }
      else          { h.v[0][c] = f0 + f3; h.v[1][c] = f1 - f2; }
      break;
    case 2: // Z : h0 = f0 +- i f2 ; h1 = f1 -+ i f3
      if (sign > 0) { h.v[0][c] = f0 + timesI(f2); h.v[1][c] = f1 - timesI(f3); }
      else          { h.v[0][c] = f0 - timesI(f2); h.v[1][c] = f1 + timesI(f3); }
      break;
    default: // T : h0 = f0 +- f2 ; h1 = f1 +- f3
      if (sign > 0) { h.v[0][c] = f0 + f2; h.v[1][c] = f1 + f3; }
      else          { h.v[0][c] = f0 - f2; h.v[1][c] = f1 - f3; }
      break;
    }
  }
}
template <class T> inline void accumRecon(Spinor<T> &f, const HalfSpinor<T> &h, int dir, int sign) {
  for (int c = 0; c < Nc; c++) {
    const cx<T> h0 = h.v[0][c], h1 = h.v[1][c];
    f.v[0][c] += h0;
    f.v[1][c] += h1;
    switch (dir) {
    case 0: // Xp: f2 -= i h1, f3 -= i h0 ; Xm: +=
      if (sign > 0) { f.v[2][c] -= timesI(h1); f.v[3][c] -= timesI(h0); }
      else          { f.v[2][c] += timesI(h1); f.v[3][c] += timesI(h0); }
      break;
    case 1: // Yp: f2 += h1, f3 -= h0 ; Ym: f2 -= h1, f3 += h0
      if (sign > 0) { f.v[2][c] += h1; f.v[3][c] -= h0; }
      else          { f.v[2][c] -= h1; f.v[3][c] += h0; }
      break;
    case 2: // Zp: f2 -= i h0, f3 += i h1 ; Zm: f2 += i h0, f3 -= i h1
      if (sign > 0) { f.v[2][c] -= timesI(h0); f.v[3][c] += timesI(h1); }
      else          { f.v[2][c] += timesI(h0); f.v[3][c] -= timesI(h1); }
      break;
    default: // Tp: f2 += h0, f3 += h1 ; Tm: -=
      if (sign > 0) { f.v[2][c] += h0; f.v[3][c] += h1; }
      else          { f.v[2][c] -= h0; f.v[3][c] -= h1; }
      break;
    }
  }
}
// gamma_mu applied to a full spinor, ref: Grid/qcd/spin/Gamma.h:552-558 (X), :288-294.. (Y,Z,T), :90-96 (5)
// Used only by the *independent* naive Cshift form below.
template <class T> inline Spinor<T> gammaMul(int mu, const Spinor<T> &p) {
  Spinor<T> r;
  for (int c = 0; c < Nc; c++) {
    const cx<T> p0 = p.v[0][c], p1 = p.v[1][c], p2 = p.v[2][c], p3 = p.v[3][c];
    switch (mu) {
    case 0: r.v[0][c] = timesI(p3); r.v[1][c] = timesI(p2); r.v[2][c] = timesMinusI(p1); r.v[3][c] = timesMinusI(p0); break;
    case 1: r.v[0][c] = -p3; r.v[1][c] = p2; r.v[2][c] = p1; r.v[3][c] = -p0; break;
    case 2: r.v[0][c] = timesI(p2); r.v[1][c] = timesMinusI(p3); r.v[2][c] = timesMinusI(p0); r.v[3][c] = timesI(p1); break;
    case 3: r.v[0][c] = p2; r.v[1][c] = p3; r.v[2][c] = p0; r.v[3][c] = p1; break;
    default: r.v[0][c] = p0; r.v[1][c] = p1; r.v[2][c] = -p2; r.v[3][c] = -p3; break; // gamma5
    }
  }
  return r;
}

// (U chi)_row = sum_col U[row][col] chi_col, ref: WilsonImpl.h:84-91, WilsonKernelsHandImplementation.h:120-147
template <class T> inline void multLink(HalfSpinor<T> &out, const ColourMatrix<T> &U, const HalfSpinor<T> &in) {
  for (int s = 0; s < Nhs; s++)
    for (int r = 0; r < Nc; r++) {
      cx<T> acc = U.m[r][0] * in.v[s][0];
      acc += U.m[r][1] * in.v[s][1];
      acc += U.m[r][2] * in.v[s][2];
      out.v[s][r] = acc;
    }
}
template <class T> inline Spinor<T> matMulSpinor(const ColourMatrix<T> &U, const Spinor<T> &in, bool adj) {
  Spinor<T> out;
  for (int s = 0; s < Ns; s++)
    for (int r = 0; r < Nc; r++) {
      cx<T> acc(0, 0);
      for (int c = 0; c < Nc; c++) acc += (adj ? conj(U.m[c][r]) : U.m[r][c]) * in.v[s][c];
      out.v[s][r] = acc;
    }
  return out;
}

// ------------------------------------------------------------------ geometry
struct Geometry {
  int L[4];   // local == global (single rank oracle)
  int Ls;     // 1 for 4D fields
  int64_t V4() const { return (int64_t)L[0] * L[1] * L[2] * L[3]; }
  int64_t V4cb() const { return V4() / 2; }
  int64_t lex4(const int x[4]) const { return x[0] + (int64_t)L[0] * (x[1] + (int64_t)L[1] * (x[2] + (int64_t)L[2] * x[3])); }
  void coor4(int64_t i, int x[4]) const {
    x[0] = i % L[0]; i /= L[0]; x[1] = i % L[1]; i /= L[1]; x[2] = i % L[2]; i /= L[2]; x[3] = (int)i;
  }
  static int parity(const int x[4]) { return (x[0] + x[1] + x[2] + x[3]) & 1; }
  // checkerboarded 4D index; ref Cartesian_red_black.h:271-286 (checker dim = x)
  int64_t cb4(const int x[4]) const { return (x[0] >> 1) + (int64_t)(L[0] / 2) * (x[1] + (int64_t)L[1] * (x[2] + (int64_t)L[2] * x[3])); }
  // inverse: coordinates of cb site icb with parity cb
  void cbcoor4(int64_t icb, int cb, int x[4]) const {
    int xh = icb % (L[0] / 2); icb /= (L[0] / 2);
    x[1] = icb % L[1]; icb /= L[1]; x[2] = icb % L[2]; icb /= L[2]; x[3] = (int)icb;
    x[0] = 2 * xh + ((cb + x[1] + x[2] + x[3]) & 1);
  }
};

// ------------------------------------------------------------------ pick / set checkerboard
// ref: Grid/lattice/Lattice_transfer.h:50-86
template <class T> void pickCheckerboard(const Geometry &g, int cb, Spinor<T> *half, const Spinor<T> *full) {
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    if (Geometry::parity(x) != cb) continue;
    int64_t ic = g.cb4(x);
    for (int s = 0; s < g.Ls; s++) half[ic * g.Ls + s] = full[i4 * g.Ls + s];
  }
}
template <class T> void setCheckerboard(const Geometry &g, int cb, Spinor<T> *full, const Spinor<T> *half) {
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    if (Geometry::parity(x) != cb) continue;
    int64_t ic = g.cb4(x);
    for (int s = 0; s < g.Ls; s++) full[i4 * g.Ls + s] = half[ic * g.Ls + s];
  }
}

// ------------------------------------------------------------------ gauge preparation
// ref: WilsonFermion5DImplementation.h:149-181 (HUmu = -0.5*Umu, DoubleStore, pickCheckerboard)
//      WilsonImpl.h:127-171 (DoubleStore incl. boundary phases)
// Uds[i4][mu]   = -1/2 * phase_mu(if x_mu==L-1) * U_mu(x)
// Uds[i4][mu+4] = -1/2 * conj(phase_mu)(if x_mu==0) * U_mu(x-mu)^dagger
template <class T>
void doubleStore(const Geometry &g, const ColourMatrix<T> *Umu /*[V4][4]*/, ColourMatrix<T> *Uds /*[V4][8]*/,
                 const cx<double> phases[4], double prefactor = -0.5) {
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    for (int mu = 0; mu < Nd; mu++) {
      cx<T> ph((T)phases[mu].re, (T)phases[mu].im);
      const ColourMatrix<T> &U = Umu[i4 * 4 + mu];
      ColourMatrix<T> &F = Uds[i4 * 8 + mu];
      bool edge = (x[mu] == g.L[mu] - 1);
      for (int r = 0; r < Nc; r++) for (int c = 0; c < Nc; c++) {
        cx<T> u = U.m[r][c];
        if (edge) u = ph * u;
        F.m[r][c] = (T)prefactor * u;
      }
      int xm[4] = {x[0], x[1], x[2], x[3]};
      xm[mu] = (x[mu] + g.L[mu] - 1) % g.L[mu];
      const ColourMatrix<T> &Ub = Umu[g.lex4(xm) * 4 + mu];
      ColourMatrix<T> &B = Uds[i4 * 8 + mu + 4];
      bool edge0 = (x[mu] == 0);
      for (int r = 0; r < Nc; r++) for (int c = 0; c < Nc; c++) {
        cx<T> u = conj(Ub.m[c][r]);
        if (edge0) u = conj(ph) * u;
        B.m[r][c] = (T)prefactor * u;
      }
    }
  }
}

// ------------------------------------------------------------------ hopping term
// One output site, all 8 legs.  ref: WilsonKernelsImplementation.h:57-68 (leg), :112-163 (site, dag / non-dag)
// non-dag: legs Xm..Tm use (1+gamma), legs Xp..Tp use (1-gamma); dag swaps the projectors, links unchanged.
template <class T>
inline Spinor<T> dhopSite(const Spinor<T> *const nbr[8], const ColourMatrix<T> *U8, int dag) {
  Spinor<T> result; zero(result);
  HalfSpinor<T> chi, Uchi;
  const int sgn_m = dag ? -1 : +1; // projector sign on the "minus" legs
  for (int mu = 0; mu < 4; mu++) { // Xm,Ym,Zm,Tm
    spProj(chi, *nbr[mu + 4], mu, sgn_m);
    multLink(Uchi, U8[mu + 4], chi);
    accumRecon(result, Uchi, mu, sgn_m);
  }
  for (int mu = 0; mu < 4; mu++) { // Xp,Yp,Zp,Tp
    spProj(chi, *nbr[mu], mu, -sgn_m);
    multLink(Uchi, U8[mu], chi);
    accumRecon(result, Uchi, mu, -sgn_m);
  }
  return result;
}

// Full-lattice Dhop (ref: WilsonFermion5DImplementation.h:437-445, WilsonFermionImplementation.h:310-318)
template <class T>
void Dhop(const Geometry &g, const ColourMatrix<T> *Uds, const Spinor<T> *in, Spinor<T> *out, int dag) {
  const int Ls = g.Ls;
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    int64_t nb[8];
    for (int mu = 0; mu < 4; mu++) {
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = (x[mu] + 1) % g.L[mu]; nb[mu] = g.lex4(y);
      y[mu] = (x[mu] + g.L[mu] - 1) % g.L[mu]; nb[mu + 4] = g.lex4(y);
    }
    for (int s = 0; s < Ls; s++) {
      const Spinor<T> *n[8];
      for (int p = 0; p < 8; p++) n[p] = &in[nb[p] * Ls + s];
      out[i4 * Ls + s] = dhopSite(n, &Uds[i4 * 8], dag);
    }
  }
}
// Checkerboarded hop: input has parity (1-ocb), output parity ocb; both in cb-lex order.
// DhopOE: in Even -> out Odd (ocb=Odd). DhopEO: in Odd -> out Even. ref: WilsonFermion5DImplementation.h:415-435
template <class T>
void DhopCB(const Geometry &g, const ColourMatrix<T> *Uds, const Spinor<T> *in, Spinor<T> *out, int ocb, int dag) {
  const int Ls = g.Ls;
#pragma omp parallel for
  for (int64_t ic = 0; ic < g.V4cb(); ic++) {
    int x[4]; g.cbcoor4(ic, ocb, x);
    int64_t i4 = g.lex4(x);
    int64_t nb[8];
    for (int mu = 0; mu < 4; mu++) {
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = (x[mu] + 1) % g.L[mu]; nb[mu] = g.cb4(y);
      y[mu] = (x[mu] + g.L[mu] - 1) % g.L[mu]; nb[mu + 4] = g.cb4(y);
    }
    for (int s = 0; s < Ls; s++) {
      const Spinor<T> *n[8];
      for (int p = 0; p < 8; p++) n[p] = &in[nb[p] * Ls + s];
      out[ic * Ls + s] = dhopSite(n, &Uds[i4 * 8], dag);
    }
  }
}

// Independent naive form (NOT via projectors), restating the reference's own check:
// ref: benchmarks/Benchmark_dwf_fp32.cc:214-245 (Dhop) and :324-364 (dagger); Benchmark_wilson.cc:122-145
//   ref(x) = -1/2 sum_mu [ (1 -+ gamma_mu) U_mu(x) src(x+mu) + (1 +- gamma_mu) U_mu(x-mu)^dag src(x-mu) ]
// Takes the *original* single-stored links Umu (periodic, no phases).
template <class T>
void DhopNaive(const Geometry &g, const ColourMatrix<T> *Umu, const Spinor<T> *in, Spinor<T> *out, int dag) {
  const int Ls = g.Ls;
#pragma omp parallel for
  for (int64_t i4 = 0; i4 < g.V4(); i4++) {
    int x[4]; g.coor4(i4, x);
    for (int s = 0; s < Ls; s++) {
      Spinor<T> acc; zero(acc);
      for (int mu = 0; mu < 4; mu++) {
        int y[4] = {x[0], x[1], x[2], x[3]};
        y[mu] = (x[mu] + 1) % g.L[mu];
        Spinor<T> t1 = matMulSpinor(Umu[i4 * 4 + mu], in[g.lex4(y) * Ls + s], false);
        Spinor<T> g1 = gammaMul(mu, t1);
        y[mu] = (x[mu] + g.L[mu] - 1) % g.L[mu];
        int64_t jm = g.lex4(y);
        Spinor<T> t2 = matMulSpinor(Umu[jm * 4 + mu], in[jm * Ls + s], true);
        Spinor<T> g2 = gammaMul(mu, t2);
        for (int a = 0; a < Ns; a++) for (int c = 0; c < Nc; c++) {
          if (!dag) acc.v[a][c] += (t1.v[a][c] - g1.v[a][c]) + (t2.v[a][c] + g2.v[a][c]);
          else      acc.v[a][c] += (t1.v[a][c] + g1.v[a][c]) + (t2.v[a][c] - g2.v[a][c]);
        }
      }
      for (int a = 0; a < Ns; a++) for (int c = 0; c < Nc; c++) out[i4 * Ls + s].v[a][c] = T(-0.5) * acc.v[a][c];
    }
  }
}

// ------------------------------------------------------------------ BLAS-1 and reductions
// ref: Grid/lattice/Lattice_arith.h:231-258 ; Lattice_reduction.h:256-311 (innerProduct = sum conj(l)*r,
// per-site in working precision, lattice sum in double), :321-372 (axpy_norm)
template <class T> inline cx<T> siteInner(const Spinor<T> &l, const Spinor<T> &r) {
  cx<T> acc(0, 0);
  for (int a = 0; a < Ns; a++) for (int c = 0; c < Nc; c++) acc += conj(l.v[a][c]) * r.v[a][c];
  return acc;
}
template <class T> cx<double> innerProduct(int64_t n, const Spinor<T> *l, const Spinor<T> *r) {
  double re = 0, im = 0;
#pragma omp parallel for reduction(+ : re, im)
  for (int64_t i = 0; i < n; i++) { cx<T> d = siteInner(l[i], r[i]); re += (double)d.re; im += (double)d.im; }
  return {re, im};
}
template <class T> double norm2(int64_t n, const Spinor<T> *x) { return innerProduct(n, x, x).re; }
template <class T> void axpy(int64_t n, Spinor<T> *z, T a, const Spinor<T> *x, const Spinor<T> *y) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < Ns; k++) for (int c = 0; c < Nc; c++) z[i].v[k][c] = a * x[i].v[k][c] + y[i].v[k][c];
}
template <class T> void axpby(int64_t n, Spinor<T> *z, T a, T b, const Spinor<T> *x, const Spinor<T> *y) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < Ns; k++) for (int c = 0; c < Nc; c++) z[i].v[k][c] = a * x[i].v[k][c] + b * y[i].v[k][c];
}
template <class T> double axpy_norm(int64_t n, Spinor<T> *z, T a, const Spinor<T> *x, const Spinor<T> *y) {
  axpy(n, z, a, x, y);
  return norm2(n, z);
}

// ------------------------------------------------------------------ Cayley (Shamir / Moebius) 5D pieces
struct CayleyCoeffs {
  int Ls = 0;
  double mass = 0, M5 = 0, b = 1, c = 0;
  std::vector<double> bs, cs, bee, cee, beo, ceo, aee, dee, lee, leem, uee, ueem;
};
// ref: CayleyFermion5DImplementation.h:411-535 (SetCoefficientsInternal), with gamma_s == 1 from the tanh
// (higham) approximation: Grid/algorithms/approx/Zolotarev.cc:473 ; DomainWallFermion.h:125-131 (b=1,c=0);
// MobiusFermion.h:63-66 (user b,c).  mass_plus = mass_minus = mass (CayleyFermion5DImplementation.h:50).
inline CayleyCoeffs cayleyCoeffs(int Ls, double mass, double M5, double b, double c) {
  CayleyCoeffs k; k.Ls = Ls; k.mass = mass; k.M5 = M5; k.b = b; k.c = c;
  auto rs = [&](std::vector<double> &v) { v.assign(Ls, 0.0); };
  rs(k.bs); rs(k.cs); rs(k.bee); rs(k.cee); rs(k.beo); rs(k.ceo); rs(k.aee); rs(k.dee); rs(k.lee); rs(k.leem); rs(k.uee); rs(k.ueem);
  const double bpc = b + c, bmc = b - c;
  for (int i = 0; i < Ls; i++) {
    const double omega = 1.0; // gamma[i]*zolo_hi, both 1
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

// chi_s = diag_s phi_s + upper_s P- psi_{s+1} + lower_s P+ psi_{s-1}   ref: CayleyFermion5Dcache.h:43-78
// dag:  chi_s = diag_s phi_s + upper_s P+ psi_{s+1} + lower_s P- psi_{s-1}   ref: :82-115
// P+ keeps spins 0,1 ; P- keeps spins 2,3  ref: TwoSpinor.h:140-182
template <class T>
void M5Dgen(int64_t nsite4, int Ls, const Spinor<T> *psi, const Spinor<T> *phi, Spinor<T> *chi,
            const std::vector<double> &lower, const std::vector<double> &diag, const std::vector<double> &upper, int dag) {
#pragma omp parallel for
  for (int64_t ss = 0; ss < nsite4; ss++) {
    std::vector<Spinor<T>> tmp(Ls);
    for (int s = 0; s < Ls; s++) {
      const Spinor<T> &pu = psi[ss * Ls + (s + 1) % Ls];
      const Spinor<T> &pl = psi[ss * Ls + (s + Ls - 1) % Ls];
      const Spinor<T> &ph = phi[ss * Ls + s];
      const T d = (T)diag[s], u = (T)upper[s], l = (T)lower[s];
      for (int a = 0; a < Ns; a++) for (int c = 0; c < Nc; c++) {
        cx<T> r = d * ph.v[a][c];
        bool upperSpin = (a < 2);
        // non-dag: upper term carries P- (spins 2,3), lower term carries P+ (spins 0,1)
        if (!dag) { if (!upperSpin) r += u * pu.v[a][c]; else r += l * pl.v[a][c]; }
        else      { if (upperSpin) r += u * pu.v[a][c]; else r += l * pl.v[a][c]; }
        tmp[s].v[a][c] = r;
      }
    }
    for (int s = 0; s < Ls; s++) chi[ss * Ls + s] = tmp[s]; // safe for chi aliasing phi/psi
  }
}

// ref: CayleyFermion5Dcache.h:117-172 (MooeeInv) and :174-230 (MooeeInvDag); real coefficients.
template <class T>
void MooeeInvGen(int64_t nsite4, const CayleyCoeffs &k, const Spinor<T> *psi, Spinor<T> *chi, int dag) {
  const int Ls = k.Ls;
#pragma omp parallel for
  for (int64_t ss4 = 0; ss4 < nsite4; ss4++) {
    const int64_t ss = ss4 * Ls;
    std::vector<Spinor<T>> out(Ls);
    // chirality roles: non-dag uses (acc: P-, carry: P+) in the forward sweep ; dag swaps
    auto projA = [&](const Spinor<T> &r, bool minus) { // returns P-(r) if minus else P+(r)
      Spinor<T> t; zero(t);
      for (int a = (minus ? 2 : 0); a < (minus ? 4 : 2); a++) for (int c = 0; c < Nc; c++) t.v[a][c] = r.v[a][c];
      return t;
    };
    auto axp = [&](Spinor<T> &r, T a, const Spinor<T> &x) { for (int i = 0; i < Ns; i++) for (int c = 0; c < Nc; c++) r.v[i][c] += a * x.v[i][c]; };
    auto scal = [&](Spinor<T> &r, T a) { for (int i = 0; i < Ns; i++) for (int c = 0; c < Nc; c++) r.v[i][c] = a * r.v[i][c]; };
    const std::vector<double> &Lm = dag ? k.ueem : k.leem; // accumulated "m" column
    const std::vector<double> &Ls1 = dag ? k.uee : k.lee;  // sub-diagonal carried term
    const std::vector<double> &Um = dag ? k.leem : k.ueem;
    const std::vector<double> &Us1 = dag ? k.lee : k.uee;
    const bool accMinus = !dag; // forward sweep: acc uses P- (non-dag) / P+ (dag)
    Spinor<T> res = psi[ss], tmp, acc;
    if (Ls == 1) { // degenerate: only the diagonal survives
      scal(res, (T)(1.0 / k.dee[0])); chi[ss] = res; continue;
    }
    tmp = projA(res, accMinus); acc = tmp; scal(acc, (T)Lm[0]);
    tmp = projA(res, !accMinus);
    out[0] = res;
    for (int s = 1; s < Ls - 1; s++) {
      res = psi[ss + s];
      axp(res, (T)(-Ls1[s - 1]), tmp);
      tmp = projA(res, accMinus);
      axp(acc, (T)Lm[s], tmp);
      tmp = projA(res, !accMinus);
      out[s] = res;
    }
    res = psi[ss + Ls - 1];
    axp(res, (T)(-Ls1[Ls - 2]), tmp);
    axp(res, (T)(-1), acc);
    // backward
    scal(res, (T)(1.0 / k.dee[Ls - 1]));
    out[Ls - 1] = res;
    acc = projA(res, !accMinus); // non-dag: spProj5p(acc,res)
    tmp = projA(res, accMinus);  // non-dag: spProj5m(tmp,res)
    for (int s = Ls - 2; s >= 0; s--) {
      res = out[s];
      scal(res, (T)(1.0 / k.dee[s]));
      axp(res, (T)(-Us1[s]), tmp);
      axp(res, (T)(-Um[s]), acc);
      tmp = projA(res, accMinus);
      out[s] = res;
    }
    for (int s = 0; s < Ls; s++) chi[ss + s] = out[s];
  }
}

// ------------------------------------------------------------------ operator objects
enum class OpKind { Wilson4D, Cayley5D };

template <class T> struct FermOp {
  OpKind kind;
  Geometry g;          // g.Ls = 1 for Wilson4D
  double mass = 0;     // Wilson mass (diag = 4+mass) or DWF mass
  CayleyCoeffs k;      // Cayley only
  std::vector<ColourMatrix<T>> Uds; // [V4][8], -1/2 and phases folded in

  T mass_word() const { return (T)mass; }   // the working-precision type, for generic callers
  int64_t V5() const { return g.V4() * g.Ls; }
  int64_t V5cb() const { return g.V4cb() * g.Ls; }
  using F = Spinor<T>;
  using Vec = std::vector<F>;

  void importGauge(const ColourMatrix<T> *Umu, const cx<double> phases[4]) {
    Uds.resize(g.V4() * 8);
    doubleStore(g, Umu, Uds.data(), phases);
  }
  // ---- hopping
  void DhopFull(const F *in, F *out, int dag) const { Dhop(g, Uds.data(), in, out, dag); }
  void DhopOE(const F *in, F *out, int dag) const { DhopCB(g, Uds.data(), in, out, Odd, dag); }
  void DhopEO(const F *in, F *out, int dag) const { DhopCB(g, Uds.data(), in, out, Even, dag); }
  // DW = Dhop + (4-M5)  ref: WilsonFermion5DImplementation.h:447-452
  void DW(const F *in, F *out, int dag) const {
    DhopFull(in, out, dag);
    axpy(V5(), out, (T)(4.0 - k.M5), in, out);
  }
  // ---- 5D pieces (n4 = number of 4D sites in the field: V4 or V4cb)
  void Meooe5D(int64_t n4, const F *psi, F *Din) const { // ref: CayleyFermion5DImplementation.h:165-174
    std::vector<double> diag = k.bs, upper = k.cs, lower = k.cs;
    upper[k.Ls - 1] = -mass * upper[k.Ls - 1]; lower[0] = -mass * lower[0];
    M5Dgen(n4, k.Ls, psi, psi, Din, lower, diag, upper, 0);
  }
  void MeooeDag5D(int64_t n4, const F *psi, F *Din) const { // ref: :248-271
    int Ls = k.Ls; std::vector<double> diag = k.bs, upper(Ls), lower(Ls);
    for (int s = 0; s < Ls; s++) {
      if (s == 0) { upper[s] = k.cs[(s + 1) % Ls]; lower[s] = -mass * k.cs[Ls - 1]; }
      else if (s == Ls - 1) { upper[s] = -mass * k.cs[0]; lower[s] = k.cs[s - 1]; }
      else { upper[s] = k.cs[s + 1]; lower[s] = k.cs[s - 1]; }
    }
    if (Ls == 1) { upper[0] = -mass * k.cs[0]; lower[0] = -mass * k.cs[0]; }
    M5Dgen(n4, Ls, psi, psi, Din, lower, diag, upper, 1);
  }
  void Mooee(int64_t n4, const F *psi, F *chi) const { // ref: :191-204 ; Wilson: WilsonFermionImplementation.h:152-157
    if (kind == OpKind::Wilson4D) { axpby(n4, chi, (T)(4.0 + mass), (T)0, psi, psi); return; }
    int Ls = k.Ls; std::vector<double> diag = k.bee, upper(Ls), lower(Ls);
    for (int i = 0; i < Ls; i++) { upper[i] = -k.cee[i]; lower[i] = -k.cee[i]; }
    upper[Ls - 1] = -mass * upper[Ls - 1]; lower[0] = -mass * lower[0];
    M5Dgen(n4, Ls, psi, psi, chi, lower, diag, upper, 0);
  }
  void MooeeDag(int64_t n4, const F *psi, F *chi) const { // ref: :206-233
    if (kind == OpKind::Wilson4D) { Mooee(n4, psi, chi); return; }
    int Ls = k.Ls; std::vector<double> diag = k.bee, upper(Ls), lower(Ls);
    for (int s = 0; s < Ls; s++) {
      if (s == 0) { upper[s] = -k.cee[(s + 1) % Ls]; lower[s] = mass * k.cee[Ls - 1]; }
      else if (s == Ls - 1) { upper[s] = mass * k.cee[0]; lower[s] = -k.cee[s - 1]; }
      else { upper[s] = -k.cee[s + 1]; lower[s] = -k.cee[s - 1]; }
    }
    if (Ls == 1) { upper[0] = mass * k.cee[0]; lower[0] = mass * k.cee[0]; }
    M5Dgen(n4, Ls, psi, psi, chi, lower, diag, upper, 1);
  }
  void MooeeInv(int64_t n4, const F *psi, F *chi) const { // Wilson: WilsonFermionImplementation.h:165-170
    if (kind == OpKind::Wilson4D) { axpby(n4, chi, (T)(1.0 / (4.0 + mass)), (T)0, psi, psi); return; }
    MooeeInvGen(n4, k, psi, chi, 0);
  }
  void MooeeInvDag(int64_t n4, const F *psi, F *chi) const {
    if (kind == OpKind::Wilson4D) { MooeeInv(n4, psi, chi); return; }
    MooeeInvGen(n4, k, psi, chi, 1);
  }
  // ---- cb operators; cb_in = checkerboard of the input field
  void Meooe(const F *psi, F *chi, int cb_in) const { // ref: :308-317 ; Wilson: WilsonFermionImplementation.h:134-141
    if (kind == OpKind::Wilson4D) { if (cb_in == Odd) DhopEO(psi, chi, 0); else DhopOE(psi, chi, 0); return; }
    Vec tmp(V5cb()); Meooe5D(g.V4cb(), psi, tmp.data());
    if (cb_in == Odd) DhopEO(tmp.data(), chi, 0); else DhopOE(tmp.data(), chi, 0);
  }
  void MeooeDag(const F *psi, F *chi, int cb_in) const { // ref: :320-329
    if (kind == OpKind::Wilson4D) { if (cb_in == Odd) DhopEO(psi, chi, 1); else DhopOE(psi, chi, 1); return; }
    Vec tmp(V5cb());
    if (cb_in == Odd) DhopEO(psi, tmp.data(), 1); else DhopOE(psi, tmp.data(), 1);
    MeooeDag5D(g.V4cb(), tmp.data(), chi);
  }
  // ---- unpreconditioned operator on the full lattice
  void M(const F *psi, F *chi) const {
    if (kind == OpKind::Wilson4D) { // ref: WilsonFermionImplementation.h:114-119  M = Dhop + (4+m)
      DhopFull(psi, chi, 0); axpy(V5(), chi, (T)(4.0 + mass), psi, chi); return;
    }
    // ref: CayleyFermion5DImplementation.h:274-286
    Vec Din(V5()); Meooe5D(g.V4(), psi, Din.data());
    DW(Din.data(), chi, 0);
    axpby(V5(), chi, (T)1, (T)1, chi, psi);
    int Ls = k.Ls; std::vector<double> diag(Ls, 1.0), upper(Ls, -1.0), lower(Ls, -1.0);
    upper[Ls - 1] = mass; lower[0] = mass; // ref: :156-163
    M5Dgen(g.V4(), Ls, psi, chi, chi, lower, diag, upper, 0);
  }
  void Mdag(const F *psi, F *chi) const {
    if (kind == OpKind::Wilson4D) { DhopFull(psi, chi, 1); axpy(V5(), chi, (T)(4.0 + mass), psi, chi); return; }
    // ref: CayleyFermion5DImplementation.h:289-304 and :236-245
    Vec Din(V5()); DW(psi, Din.data(), 1);
    MeooeDag5D(g.V4(), Din.data(), chi);
    int Ls = k.Ls; std::vector<double> diag(Ls, 1.0), upper(Ls, -1.0), lower(Ls, -1.0);
    upper[Ls - 1] = -mass * upper[Ls - 1]; lower[0] = -mass * lower[0];
    M5Dgen(g.V4(), Ls, psi, chi, chi, lower, diag, upper, 1);
    axpby(V5(), chi, (T)1, (T)1, chi, psi);
  }
  // ---- Schur even-odd operator on one checkerboard (cb = parity of in/out)
  // ref: Grid/algorithms/LinearOperator.h:325-349
  void Mpc(const F *in, F *out, int cb) const {
    Vec tmp(V5cb());
    Meooe(in, tmp.data(), cb);
    MooeeInv(g.V4cb(), tmp.data(), out);
    Meooe(out, tmp.data(), 1 - cb);
    Mooee(g.V4cb(), in, out);
    axpy(V5cb(), out, (T)-1, tmp.data(), out);
  }
  void MpcDag(const F *in, F *out, int cb) const {
    Vec tmp(V5cb());
    MeooeDag(in, tmp.data(), cb);
    MooeeInvDag(g.V4cb(), tmp.data(), out);
    MeooeDag(out, tmp.data(), 1 - cb);
    MooeeDag(g.V4cb(), in, out);
    axpy(V5cb(), out, (T)-1, tmp.data(), out);
  }
  void HermOp(const F *in, F *out, int cb) const { // MpcDagMpc, ref: LinearOperator.h:291-307
    Vec tmp(V5cb()); Mpc(in, tmp.data(), cb); MpcDag(tmp.data(), out, cb);
  }
  // ---- single hop legs and force terms
  // One leg of the hopping term on the full lattice.  point 0..3 = forward leg mu (reads x+mu, link U_mu(x)), 4..7 = backward
  // leg mu; projector as in dhopSite (non-dag: forward legs carry (1-gamma), backward (1+gamma); dag swaps).
  // ref: WilsonKernelsImplementation.h:375-412 (DhopDirKernel), :57-68 (leg)
  void DhopLeg(const F *in, F *out, int point, int dag) const {
    const int Ls = g.Ls, mu = point & 3, fwd = point < 4;
    const int sgn = (fwd ? -1 : +1) * (dag ? -1 : +1);
#pragma omp parallel for
    for (int64_t i4 = 0; i4 < g.V4(); i4++) {
      int x[4]; g.coor4(i4, x);
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = fwd ? (x[mu] + 1) % g.L[mu] : (x[mu] + g.L[mu] - 1) % g.L[mu];
      const int64_t nb = g.lex4(y);
      for (int s = 0; s < Ls; s++) {
        HalfSpinor<T> chi, Uchi;
        Spinor<T> r; zero(r);
        spProj(chi, in[nb * Ls + s], mu, sgn);
        multLink(Uchi, Uds[i4 * 8 + point], chi);
        accumRecon(r, Uchi, mu, sgn);
        out[i4 * Ls + s] = r;
      }
    }
  }
  // DhopDir(in, out, dir, disp = +-1): that leg of the non-dagger hop   ref: WilsonFermion5DImplementation.h:183-200
  void DhopDir(const F *in, F *out, int dir, int disp) const { DhopLeg(in, out, disp == 1 ? dir : dir + 4, 0); }
  // DhopDeriv: mat_mu(x) = sum_s trace_spin [ Btilde_mu(x,s) (x) A(x,s)^dagger ], Btilde_mu = forward leg mu of Dhop^(dag) on B
  // ref: WilsonFermion5DImplementation.h:212-262 (DerivInternal), WilsonImpl.h:193-238 (InsertForce5D),
  //      Grid/tensors/Tensor_outer.h (outerProduct(l, r) = l conj(r))
  void DhopDeriv(ColourMatrix<T> *mat /* [V4][4] */, const F *A, const F *B, int dag) const {
    const int Ls = g.Ls;
    Vec Btilde(V5());
    for (int mu = 0; mu < 4; mu++) {
      DhopLeg(B, Btilde.data(), mu, dag);
#pragma omp parallel for
      for (int64_t i4 = 0; i4 < g.V4(); i4++) {
        ColourMatrix<T> m; std::memset((void *)&m, 0, sizeof(m));
        for (int s = 0; s < Ls; s++) {
          const F &b = Btilde[i4 * Ls + s], &a = A[i4 * Ls + s];
          for (int sp = 0; sp < Ns; sp++) for (int c1 = 0; c1 < Nc; c1++) for (int c2 = 0; c2 < Nc; c2++) m.m[c1][c2] += b.v[sp][c1] * conj(a.v[sp][c2]);
        }
        mat[i4 * 4 + mu] = m;
      }
    }
  }
  // MDeriv(mat, U, V, dag): d/dU of U^dagger M V (dag: U^dagger M^dagger V); the 5D factor is applied to the side it acts on
  // ref: CayleyFermion5DImplementation.h:347-360 ; WilsonFermion: MDeriv = DhopDeriv (FermionOperator default)
  void MDeriv(ColourMatrix<T> *mat, const F *U, const F *V, int dag) const {
    if (kind == OpKind::Wilson4D) { DhopDeriv(mat, U, V, dag); return; }
    Vec Din(V5());
    if (!dag) { Meooe5D(g.V4(), V, Din.data()); DhopDeriv(mat, U, Din.data(), dag); }
    else { Meooe5D(g.V4(), U, Din.data()); DhopDeriv(mat, Din.data(), V, dag); }
  }
  // ---- even-odd force terms.  Fields are red-black; mat is the FULL lexicographic gauge field, of which only the sites of
  // parity ocb (= A's checkerboard) are written, the way SchurDifferentiableOperator assembles ForceE / ForceO.
  // ref: WilsonFermion5DImplementation.h:277-305 (DhopDerivEO / OE), CayleyFermion5DImplementation.h:361-390 (MoeDeriv / MeoDeriv)
  void DhopLegCB(const F *in, F *out, int ocb, int point, int dag) const {   // in has parity 1-ocb
    const int Ls = g.Ls, mu = point & 3, fwd = point < 4;
    const int sgn = (fwd ? -1 : +1) * (dag ? -1 : +1);
#pragma omp parallel for
    for (int64_t ic = 0; ic < g.V4cb(); ic++) {
      int x[4]; g.cbcoor4(ic, ocb, x);
      const int64_t i4 = g.lex4(x);
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = fwd ? (x[mu] + 1) % g.L[mu] : (x[mu] + g.L[mu] - 1) % g.L[mu];
      const int64_t nb = g.cb4(y);
      for (int s = 0; s < Ls; s++) {
        HalfSpinor<T> chi, Uchi;
        Spinor<T> r; zero(r);
        spProj(chi, in[nb * Ls + s], mu, sgn);
        multLink(Uchi, Uds[i4 * 8 + point], chi);
        accumRecon(r, Uchi, mu, sgn);
        out[ic * Ls + s] = r;
      }
    }
  }
  // A on parity ocb, B on parity 1-ocb
  void DhopDerivCB(ColourMatrix<T> *mat, const F *A, const F *B, int dag, int ocb) const {
    const int Ls = g.Ls;
    Vec Btilde(V5cb());
    for (int mu = 0; mu < 4; mu++) {
      DhopLegCB(B, Btilde.data(), ocb, mu, dag);
#pragma omp parallel for
      for (int64_t ic = 0; ic < g.V4cb(); ic++) {
        int x[4]; g.cbcoor4(ic, ocb, x);
        ColourMatrix<T> m; std::memset((void *)&m, 0, sizeof(m));
        for (int s = 0; s < Ls; s++) {
          const F &b = Btilde[ic * Ls + s], &a = A[ic * Ls + s];
          for (int sp = 0; sp < Ns; sp++) for (int c1 = 0; c1 < Nc; c1++) for (int c2 = 0; c2 < Nc; c2++) m.m[c1][c2] += b.v[sp][c1] * conj(a.v[sp][c2]);
        }
        mat[g.lex4(x) * 4 + mu] = m;
      }
    }
  }
  // MeoDeriv (ocb = Even: U Even, V Odd) / MoeDeriv (ocb = Odd)
  void MeooeDeriv(ColourMatrix<T> *mat, const F *U, const F *V, int dag, int ocb) const {
    if (kind == OpKind::Wilson4D) { DhopDerivCB(mat, U, V, dag, ocb); return; }
    Vec Din(V5cb());
    if (!dag) { Meooe5D(g.V4cb(), V, Din.data()); DhopDerivCB(mat, U, Din.data(), dag, ocb); }
    else { Meooe5D(g.V4cb(), U, Din.data()); DhopDerivCB(mat, Din.data(), V, dag, ocb); }
  }
  // SchurDifferentiableOperator::MpcDeriv / MpcDagDeriv: U, V on the Odd checkerboard, Force on the full lattice
  // ref: Grid/qcd/action/pseudofermion/EvenOddSchurDifferentiable.h:52-137
  void MpcDeriv(ColourMatrix<T> *Force, const F *U, const F *V, int dagger) const {
    Vec tmp1(V5cb()), tmp2(V5cb());
    if (!dagger) {
      Meooe(V, tmp1.data(), Odd); MooeeInv(g.V4cb(), tmp1.data(), tmp2.data());
      MeooeDeriv(Force, U, tmp2.data(), 0, Odd);                     // MoeDeriv(ForceO, U, tmp2, DaggerNo)
      MeooeDag(U, tmp1.data(), Odd); MooeeInvDag(g.V4cb(), tmp1.data(), tmp2.data());
      MeooeDeriv(Force, tmp2.data(), V, 0, Even);                    // MeoDeriv(ForceE, tmp2, V, DaggerNo)
    } else {
      MeooeDag(V, tmp1.data(), Odd); MooeeInvDag(g.V4cb(), tmp1.data(), tmp2.data());
      MeooeDeriv(Force, U, tmp2.data(), 1, Odd);
      Meooe(U, tmp1.data(), Odd); MooeeInv(g.V4cb(), tmp1.data(), tmp2.data());
      MeooeDeriv(Force, tmp2.data(), V, 1, Even);
    }
    for (int64_t i = 0; i < g.V4() * 4; i++) for (int a = 0; a < Nc; a++) for (int b = 0; b < Nc; b++) Force[i].m[a][b] = -Force[i].m[a][b];
  }
  // ---- physical 4D <-> 5D maps.  4D fields have V4 sites, 5D fields V4*Ls (s fastest).
  // Wilson4D: every map is the identity (ref: FermionOperator.h:172-191).
  // Dminus: chi_s = psi_s - cs[s] DW(psi)_s ; DminusDag uses DW^dag   ref: CayleyFermion5DImplementation.h:132-153
  void Dminus(const F *psi, F *chi, int dag) const {
    if (kind == OpKind::Wilson4D) { std::copy(psi, psi + V5(), chi); return; }
    Vec tmp(V5()); DW(psi, tmp.data(), dag);
    const int Ls = k.Ls;
#pragma omp parallel for
    for (int64_t i = 0; i < V5(); i++) {
      const T c = (T)(-k.cs[i % Ls]);
      for (int a = 0; a < Ns; a++) for (int col = 0; col < Nc; col++) chi[i].v[a][col] = psi[i].v[a][col] + c * tmp[i].v[a][col];
    }
  }
  // imported5d_{s=0} = P+ in4d, imported5d_{s=Ls-1} = P- in4d, zero elsewhere   ref: :100-113
  void ImportUnphysicalFermion(const F *in4, F *out5) const {
    if (kind == OpKind::Wilson4D) { std::copy(in4, in4 + V5(), out5); return; }
    const int Ls = k.Ls;
    std::memset((void *)out5, 0, sizeof(F) * V5());
#pragma omp parallel for
    for (int64_t i4 = 0; i4 < g.V4(); i4++)
      for (int c = 0; c < Nc; c++) {
        for (int a = 0; a < 2; a++) out5[i4 * Ls].v[a][c] = in4[i4].v[a][c];           // Ls >= 2 (Cayley operators)
        for (int a = 2; a < 4; a++) out5[i4 * Ls + Ls - 1].v[a][c] = in4[i4].v[a][c];
      }
  }
  // ref: :115-130  Dminus of the above
  void ImportPhysicalFermionSource(const F *in4, F *out5) const {
    if (kind == OpKind::Wilson4D) { std::copy(in4, in4 + V5(), out5); return; }
    Vec tmp(V5()); ImportUnphysicalFermion(in4, tmp.data()); Dminus(tmp.data(), out5, 0);
  }
  // exported4d = P- sol_{s=0} + P+ sol_{s=Ls-1}   ref: :58-69
  void ExportPhysicalFermionSolution(const F *sol5, F *out4) const {
    if (kind == OpKind::Wilson4D) { std::copy(sol5, sol5 + V5(), out4); return; }
    const int Ls = k.Ls;
#pragma omp parallel for
    for (int64_t i4 = 0; i4 < g.V4(); i4++)
      for (int c = 0; c < Nc; c++) {
        for (int a = 0; a < 2; a++) out4[i4].v[a][c] = sol5[i4 * Ls + Ls - 1].v[a][c];
        for (int a = 2; a < 4; a++) out4[i4].v[a][c] = sol5[i4 * Ls].v[a][c];
      }
  }
  // exported4d = P+ src_{s=0} + P- src_{s=Ls-1}   ref: :88-99
  void ExportPhysicalFermionSource(const F *src5, F *out4) const {
    if (kind == OpKind::Wilson4D) { std::copy(src5, src5 + V5(), out4); return; }
    const int Ls = k.Ls;
#pragma omp parallel for
    for (int64_t i4 = 0; i4 < g.V4(); i4++)
      for (int c = 0; c < Nc; c++) {
        for (int a = 0; a < 2; a++) out4[i4].v[a][c] = src5[i4 * Ls].v[a][c];
        for (int a = 2; a < 4; a++) out4[i4].v[a][c] = src5[i4 * Ls + Ls - 1].v[a][c];
      }
  }
  // ---- SchurRedBlackDiagMooeeSolve pieces   ref: Grid/algorithms/iterative/SchurRedBlack.h:385-430
  // src_o' = MpcDag (src_o - Meooe MooeeInv src_e)
  void RedBlackSource(const F *src, F *src_e, F *src_o) const {
    Geometry g5 = g;
    Vec tmp(V5cb()), Mtmp(V5cb()), so(V5cb());
    pickCheckerboard(g5, Even, src_e, src);
    pickCheckerboard(g5, Odd, so.data(), src);
    MooeeInv(g.V4cb(), src_e, tmp.data());
    Meooe(tmp.data(), Mtmp.data(), Even);
    axpy(V5cb(), tmp.data(), (T)-1, Mtmp.data(), so.data());
    MpcDag(tmp.data(), src_o, Odd);
  }
  // sol_e = MooeeInv (src_e - Meooe sol_o) ; sol = [sol_e, sol_o]
  void RedBlackSolution(const F *sol_o, const F *src_e, F *sol) const {
    Vec tmp(V5cb()), sol_e(V5cb());
    Meooe(sol_o, tmp.data(), Odd);
    axpy(V5cb(), tmp.data(), (T)-1, tmp.data(), src_e);
    MooeeInv(g.V4cb(), tmp.data(), sol_e.data());
    setCheckerboard(g, Even, sol, sol_e.data());
    setCheckerboard(g, Odd, sol, sol_o);
  }
};

// ------------------------------------------------------------------ solvers
struct CGResult { int iterations = 0; double true_residual = 0; int converged = 0; };

// ref: Grid/algorithms/iterative/ConjugateGradient.h:68-257 ; the operator is SchurDiagMooee HermOp on `cb`.
template <class T>
CGResult ConjugateGradient(const FermOp<T> &op, int cb, const Spinor<T> *src, Spinor<T> *psi, double tol, int maxit, double shift = 0.0) {
  // shift != 0: the operator is HermOp + shift (ShiftedLinop of ConjugateGradientMultiShiftMixedPrec.h:44-70)
  const int64_t n = op.V5cb();
  std::vector<Spinor<T>> p(n), mmp(n), r(n);
  CGResult res;
  double ssq = norm2(n, src), guess = norm2(n, psi), a, cp, c, d, b;
  if (guess == 0.0) { std::copy(src, src + n, r.begin()); p = r; a = ssq; }
  else {
    op.HermOp(psi, mmp.data(), cb);
    if (shift != 0.0) axpy(n, mmp.data(), (T)shift, psi, mmp.data());
    axpy(n, r.data(), (T)-1, mmp.data(), src);
    p = r; a = norm2(n, p.data());
  }
  cp = a;
  if (ssq == 0.0) { std::memset((void *)psi, 0, sizeof(Spinor<T>) * n); res.iterations = 1; res.true_residual = 0; res.converged = 1; return res; }
  const double rsq = tol * tol * ssq;
  if (cp <= rsq) { res.true_residual = std::sqrt(a / ssq); res.iterations = 0; res.converged = 1; return res; }
  int k;
  for (k = 1; k <= maxit; k++) {
    c = cp;
    op.HermOp(p.data(), mmp.data(), cb);
    if (shift != 0.0) axpy(n, mmp.data(), (T)shift, p.data(), mmp.data());
    d = innerProduct(n, p.data(), mmp.data()).re;
    a = c / d;
    cp = axpy_norm(n, r.data(), (T)(-a), mmp.data(), r.data());
    b = cp / c;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++)
      for (int s = 0; s < Ns; s++) for (int col = 0; col < Nc; col++) {
        psi[i].v[s][col] = (T)a * p[i].v[s][col] + psi[i].v[s][col];
        p[i].v[s][col] = (T)b * p[i].v[s][col] + r[i].v[s][col];
      }
    if (cp <= rsq) {
      op.HermOp(psi, mmp.data(), cb);
      if (shift != 0.0) axpy(n, mmp.data(), (T)shift, psi, mmp.data());
      axpy(n, p.data(), (T)-1, src, mmp.data()); // p = mmp - src
      res.true_residual = std::sqrt(norm2(n, p.data())) / std::sqrt(ssq);
      res.iterations = k; res.converged = 1;
      return res;
    }
  }
  res.iterations = k; res.converged = 0;
  return res;
}

// SchurRedBlackDiagMooeeSolve<Field>(ConjugateGradient)(op, src, sol) with the ZeroGuesser: red-black source, CG on
// MpcDagMpc for the odd checkerboard, even-site reconstruction, then the unpreconditioned residual |M sol - src| / |src|
// the reference prints.   ref: Grid/algorithms/iterative/SchurRedBlack.h:238-290,385-430
template <class T>
CGResult SchurRedBlackDiagMooeeSolve(const FermOp<T> &op, const Spinor<T> *src, Spinor<T> *sol, double tol, int maxit, double *unprec_resid) {
  const int64_t n = op.V5cb(), nf = op.V5();
  std::vector<Spinor<T>> src_e(n), src_o(n), sol_o(n), resid(nf);
  op.RedBlackSource(src, src_e.data(), src_o.data());
  std::memset((void *)sol_o.data(), 0, sizeof(Spinor<T>) * n);
  CGResult res = ConjugateGradient(op, Odd, src_o.data(), sol_o.data(), tol, maxit);
  op.RedBlackSolution(sol_o.data(), src_e.data(), sol);
  op.M(sol, resid.data());
  axpy(nf, resid.data(), (T)-1, src, resid.data());
  if (unprec_resid) *unprec_resid = std::sqrt(norm2(nf, resid.data()) / norm2(nf, src));
  return res;
}

struct MixedCGResult { int inner_iterations = 0, outer_iterations = 0, final_iterations = 0; double true_residual = 0; int converged = 0; };

template <class TD, class TF> void precisionChange(int64_t n, Spinor<TD> *out, const Spinor<TF> *in) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++)
    for (int s = 0; s < Ns; s++) for (int c = 0; c < Nc; c++) { out[i].v[s][c].re = (TD)in[i].v[s][c].re; out[i].v[s][c].im = (TD)in[i].v[s][c].im; }
}

// ref: Grid/algorithms/iterative/ConjugateGradientMixedPrec.h:71-167
inline MixedCGResult MixedPrecisionCG(const FermOp<double> &op_d, const FermOp<float> &op_f, int cb, const Spinor<double> *src_d_in,
                                      Spinor<double> *sol_d, double tol, int maxinner, int maxouter, double inner_tol0 = -1.0, double shift = 0.0) {
  const int64_t n = op_d.V5cb();
  MixedCGResult R;
  const double src_norm = norm2(n, src_d_in), stop = src_norm * tol * tol, OuterLoopNormMult = 100.0;
  std::vector<Spinor<double>> tmp_d(n), src_d(src_d_in, src_d_in + n);
  std::vector<Spinor<float>> src_f(n), sol_f(n);
  double inner_tol = inner_tol0 > 0 ? inner_tol0 : tol;
  int outer;
  for (outer = 0; outer < maxouter; outer++) {
    op_d.HermOp(sol_d, tmp_d.data(), cb);
    if (shift != 0.0) axpy(n, tmp_d.data(), shift, sol_d, tmp_d.data());
    double norm = axpy_norm(n, src_d.data(), -1.0, tmp_d.data(), src_d_in);
    if (norm < OuterLoopNormMult * stop) break;
    while (norm * inner_tol * inner_tol < stop) inner_tol *= 2;
    precisionChange(n, src_f.data(), src_d.data());
    std::memset((void *)sol_f.data(), 0, sizeof(Spinor<float>) * n);
    CGResult in = ConjugateGradient(op_f, cb, src_f.data(), sol_f.data(), inner_tol, maxinner, shift);
    R.inner_iterations += in.iterations;
    precisionChange(n, tmp_d.data(), sol_f.data());
    axpy(n, sol_d, 1.0, tmp_d.data(), sol_d);
  }
  R.outer_iterations = outer;
  CGResult fin = ConjugateGradient(op_d, cb, src_d_in, sol_d, tol, maxinner, shift);
  R.final_iterations = fin.iterations; R.true_residual = fin.true_residual; R.converged = fin.converged;
  return R;
}

// ref: Grid/algorithms/iterative/ConjugateGradientMixedPrecBatched.h:79-207 -- the defect-correction loop over a BATCH of right-hand
// sides with one restart schedule: every outer iteration recomputes all residuals in double precision, stops when ALL are below
// OuterLoopNormMult * stop, loosens the common inner tolerance by the LARGEST residual / stop of the batch (:158-163), runs one
// single-precision CG per right-hand side, and finishes with a double-precision patch-up CG per right-hand side (:190-195).
struct BatchedCGResult { int outer_iterations = 0; std::vector<int> inner_iterations, final_iterations; std::vector<double> true_residual; };
inline BatchedCGResult MixedPrecisionCGBatched(const FermOp<double> &op_d, const FermOp<float> &op_f, int cb, int nbatch, const Spinor<double> *const *src_d_in,
                                               Spinor<double> *const *sol_d, double tol, int maxinner, int maxouter, int maxpatch) {
  const int64_t n = op_d.V5cb();
  BatchedCGResult R;
  R.inner_iterations.assign(nbatch, 0); R.final_iterations.assign(nbatch, 0); R.true_residual.assign(nbatch, 0.0);
  const double OuterLoopNormMult = 100.0;
  std::vector<double> stop(nbatch), norm(nbatch, 0.0);
  std::vector<std::vector<Spinor<double>>> src_d(nbatch);
  std::vector<std::vector<Spinor<float>>> src_f(nbatch), sol_f(nbatch);
  std::vector<Spinor<double>> tmp_d(n);
  for (int i = 0; i < nbatch; i++) {
    stop[i] = norm2(n, src_d_in[i]) * tol * tol;
    src_d[i].assign(src_d_in[i], src_d_in[i] + n); src_f[i].resize(n); sol_f[i].resize(n);
  }
  double inner_tol = tol;
  int outer;
  for (outer = 0; outer < maxouter; outer++) {
    bool all_converged = true;
    for (int i = 0; i < nbatch; i++) {
      op_d.HermOp(sol_d[i], tmp_d.data(), cb);
      norm[i] = axpy_norm(n, src_d[i].data(), -1.0, tmp_d.data(), src_d_in[i]);
      precisionChange(n, src_f[i].data(), src_d[i].data());
      std::memset((void *)sol_f[i].data(), 0, sizeof(Spinor<float>) * n);
      if (norm[i] > OuterLoopNormMult * stop[i]) all_converged = false;
    }
    if (all_converged) break;
    const double norm_max = *std::max_element(norm.begin(), norm.end()), stop_max = *std::max_element(stop.begin(), stop.end());
    while (norm_max * inner_tol * inner_tol < stop_max) inner_tol *= 2;
    for (int i = 0; i < nbatch; i++) {
      CGResult in = ConjugateGradient(op_f, cb, src_f[i].data(), sol_f[i].data(), inner_tol, maxinner, 0.0);
      R.inner_iterations[i] += in.iterations;
      precisionChange(n, tmp_d.data(), sol_f[i].data());
      axpy(n, sol_d[i], 1.0, tmp_d.data(), sol_d[i]);
    }
  }
  R.outer_iterations = outer;
  for (int i = 0; i < nbatch; i++) {
    CGResult fin = ConjugateGradient(op_d, cb, src_d_in[i], sol_d[i], tol, maxpatch, 0.0);
    R.final_iterations[i] = fin.iterations; R.true_residual[i] = fin.true_residual;
  }
  return R;
}

} // namespace oracle
