// Test_dwf_force-shaped driver on the B200-native library (ref: tests/forces/Test_dwf_force.cc:33-141).
// S = |M phi|^2 for DomainWallFermionD; the links are moved by U' = exp(dt P) U with Gaussian traceless anti-Hermitian P, and
// the change of S is compared with the prediction dt * sum_x,mu tr(P_mu 2 Ta(UdSdU_mu)), UdSdU = MDeriv(Mphi, phi, DaggerNo) +
// MDeriv(phi, Mphi, DaggerYes).  Grid's lattice-matrix algebra (Ta, exponentiation, trace) is outside this library's scope
// (DESIGN.md section 7), so the 3x3 work on the links is done here on the host, on the lexicographic arrays the C ABI exchanges.
// usage: Test_dwf_force [--grid x.y.z.t] [--Ls n]
#include "../include/gridb200.hpp"
#include <cstring>
#include <iostream>
#include <random>
using namespace gridb200;
typedef std::complex<double> cd;
struct M3 { cd m[9]; };
static M3 mul(const M3 &a, const M3 &b) { M3 c; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { cd s = 0; for (int k = 0; k < 3; k++) s += a.m[3 * i + k] * b.m[3 * k + j]; c.m[3 * i + j] = s; } return c; }
static M3 ta(const M3 &a) {   // traceless anti-Hermitian part, ref: Grid/tensors/Tensor_Ta.h
  M3 r; cd tr = 0;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = 0.5 * (a.m[3 * i + j] - std::conj(a.m[3 * j + i]));
  for (int i = 0; i < 3; i++) tr += r.m[4 * i];
  for (int i = 0; i < 3; i++) r.m[4 * i] -= tr / 3.0;
  return r;
}

int main(int argc, char **argv) {
  Grid_init(&argc, &argv);
  Coordinate latt4 = {8, 8, 8, 8}, mpi = {1, 1, 1, 1}, simd = {1, 1, 1, 1};
  int Ls = 8;
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "--grid")) sscanf(argv[i + 1], "%d.%d.%d.%d", &latt4[0], &latt4[1], &latt4[2], &latt4[3]);
    if (!strcmp(argv[i], "--Ls")) Ls = atoi(argv[i + 1]);
  }
  GridCartesian *UGrid = SpaceTimeGrid::makeFourDimGrid(latt4, simd, mpi);
  GridRedBlackCartesian *UrbGrid = SpaceTimeGrid::makeFourDimRedBlackGrid(UGrid);
  GridCartesian *FGrid = SpaceTimeGrid::makeFiveDimGrid(Ls, UGrid);
  GridRedBlackCartesian *FrbGrid = SpaceTimeGrid::makeFiveDimRedBlackGrid(Ls, UGrid);
  GridParallelRNG RNG5(FGrid); RNG5.SeedFixedIntegers(std::vector<int>({5, 6, 7, 8}));
  GridParallelRNG RNG4(UGrid); RNG4.SeedFixedIntegers(std::vector<int>({1, 2, 3, 5}));

  LatticeFermionD phi(FGrid); random(RNG5, phi);
  LatticeFermionD Mphi(FGrid), MphiPrime(FGrid);
  LatticeGaugeFieldD U(UGrid);
  SU<3>::HotConfiguration(RNG4, U);

  RealD mass = 0.01, M5 = 1.8;
  DomainWallFermionD Ddwf(U, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);
  Ddwf.M(phi, Mphi);
  const RealD S = innerProduct(Mphi, Mphi).real();

  // the derivative of phi^dag Mdag M phi with respect to U
  LatticeGaugeFieldD UdSdU(UGrid), tmp(UGrid);
  const size_t V = (size_t)UGrid->gSites();
  std::vector<M3> f1(V * 4), f2(V * 4), Uh(V * 4), mom(V * 4);
  Ddwf.MDeriv(tmp, Mphi, phi, DaggerNo);  tmp.ExportLex(f1.data(), GB_F64);
  Ddwf.MDeriv(tmp, phi, Mphi, DaggerYes); tmp.ExportLex(f2.data(), GB_F64);
  U.ExportLex(Uh.data(), GB_F64);

  // modify the gauge field a little: U' = exp(dt P) U, sixth-order series as the reference does (:95-117)
  const RealD dt = 0.0001;
  std::mt19937_64 gen(99);
  std::normal_distribution<double> gauss(0.0, 1.0);
  RealD dSpred = 0;
  for (size_t i = 0; i < V * 4; i++) {
    M3 g; for (auto &c : g.m) c = cd(gauss(gen), gauss(gen));
    const M3 P = ta(g);
    mom[i] = P;
    M3 term = Uh[i], Up = Uh[i];
    for (int n = 1; n <= 6; n++) { term = mul(P, term); for (auto &c : term.m) c *= dt / n; for (int k = 0; k < 9; k++) Up.m[k] += term.m[k]; }
    M3 F; for (int k = 0; k < 9; k++) F.m[k] = f1[i].m[k] + f2[i].m[k];
    const M3 T = ta(F);                                      // mommu = Ta(UdSdU) * 2.0 ; dS += trace(mom * force) dt   (:127-141)
    cd tr = 0; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) tr += P.m[3 * a + b] * 2.0 * T.m[3 * b + a];
    dSpred += tr.real() * dt;
    Uh[i] = Up;
  }
  LatticeGaugeFieldD Uprime(UGrid);
  Uprime.ImportLex(Uh.data(), GB_F64);
  Ddwf.ImportGauge(Uprime);
  Ddwf.M(phi, MphiPrime);
  const RealD Sprime = innerProduct(MphiPrime, MphiPrime).real();

  std::cout.precision(14);
  std::cout << " S      " << S << std::endl;
  std::cout << " Sprime " << Sprime << std::endl;
  std::cout << "dS      " << Sprime - S << std::endl;
  std::cout << "predict dS    " << dSpred << std::endl;
  assert(std::fabs(Sprime - S - dSpred) < 1.0);                                    // the reference's criterion (:154)
  assert(std::fabs(Sprime - S - dSpred) < 1e-2 * std::fabs(dSpred) + 1e-8 * S);     // and a meaningful one: O(dt) relative agreement
  std::cout << "Test_dwf_force: PASS" << std::endl;
  Grid_finalize();
  return 0;
}
