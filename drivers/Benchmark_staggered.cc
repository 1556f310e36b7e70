// Benchmark_staggered-shaped driver on the B200-native library.
// Same flow as the reference's benchmarks/Benchmark_staggered.cc:36-120 (grid, random source, random gauge field used as both
// thin and fat links, ImprovedStaggeredFermion with c1 = 9/8, c2 = -1/24, u0 = 1, timed Dhop calls), plus the even-odd check
// of tests/core/Test_staggered.cc, written against include/gridb200.hpp.
// usage: Benchmark_staggered [--grid x.y.z.t] [--ncall n]
#include "../include/gridb200.hpp"
#include <chrono>
#include <cstring>
#include <iostream>
using namespace gridb200;

static double usecond() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  Grid_init(&argc, &argv);
  Coordinate latt = {16, 16, 16, 16}, mpi = {1, 1, 1, 1}, simd = {1, 1, 1, 1};
  int ncall = 1000;
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "--grid")) sscanf(argv[i + 1], "%d.%d.%d.%d", &latt[0], &latt[1], &latt[2], &latt[3]);
    if (!strcmp(argv[i], "--ncall")) ncall = atoi(argv[i + 1]);
  }
  GridCartesian *Grid = SpaceTimeGrid::makeFourDimGrid(latt, simd, mpi);
  GridRedBlackCartesian *RBGrid = SpaceTimeGrid::makeFourDimRedBlackGrid(Grid);
  GridParallelRNG pRNG(Grid); pRNG.SeedFixedIntegers({1, 2, 3, 4});

  LatticeStaggeredFermionF src(Grid); random(pRNG, src);
  LatticeStaggeredFermionF result(Grid); result.Zero();
  LatticeGaugeFieldF Umu(Grid);
  SU<3>::HotConfiguration(pRNG, Umu);

  double volume = 1;
  for (int mu = 0; mu < 4; mu++) volume *= latt[mu];
  RealD mass = 0.1, c1 = 9.0 / 8.0, c2 = -1.0 / 24.0, u0 = 1.0;   // ref: Benchmark_staggered.cc:92-95
  ImprovedStaggeredFermionF Ds(Umu, Umu, *Grid, *RBGrid, mass, c1, c2, u0);

  std::cout << "Calling Ds" << std::endl;
  Ds.Dhop(src, result, 0);
  gb_synchronize(Runtime::ctx());
  double t0 = usecond();
  for (int i = 0; i < ncall; i++) Ds.Dhop(src, result, 0);
  gb_synchronize(Runtime::ctx());
  double t1 = usecond();
  double flops = (16 * (3 * (6 + 8 + 8)) + 15 * 3 * 2) * volume * ncall; // == 1146 per site, ref: Benchmark_staggered.cc:105
  std::cout << "Called Ds" << std::endl;
  std::cout << "norm result " << norm2(result) << std::endl;
  std::cout << "mflop/s =   " << flops / (t1 - t0) << std::endl;

  // Deo + Doe == D and anti-Hermiticity (ref: tests/core/Test_staggered.cc)
  LatticeStaggeredFermionF src_e(RBGrid), src_o(RBGrid), r_e(RBGrid), r_o(RBGrid), r_eo(Grid), err(Grid);
  pickCheckerboard(Even, src_e, src);
  pickCheckerboard(Odd, src_o, src);
  Ds.DhopEO(src_o, r_e, DaggerNo);
  Ds.DhopOE(src_e, r_o, DaggerNo);
  setCheckerboard(r_eo, r_o);
  setCheckerboard(r_eo, r_e);
  axpy(err, -1.0, result, r_eo);
  std::cout << "norm diff (Deo+Doe - D)   " << norm2(err) << std::endl;
  assert(norm2(err) < 1.0e-8);
  ComplexD a = innerProduct(src, result);
  std::cout << "Re <src, D src> / |.| = " << a.real() / std::abs(a) << " (anti-Hermitian: 0)" << std::endl;
  assert(std::abs(a.real()) < 1e-5 * std::abs(a) + 1e-6 * norm2(src));

  // even-odd CG on SchurStaggeredOperator
  LatticeStaggeredFermionD srcd(Grid); random(pRNG, srcd);
  LatticeGaugeFieldD Ud(Grid);
  SU<3>::HotConfiguration(pRNG, Ud);
  ImprovedStaggeredFermionD Dsd(Ud, Ud, *Grid, *RBGrid, mass, c1, c2, u0);
  LatticeStaggeredFermionD so(RBGrid), sol(RBGrid);
  pickCheckerboard(Odd, so, srcd);
  sol.Zero();
  SchurStaggeredOperator<ImprovedStaggeredFermionD, LatticeFermionD> HermOp(Dsd);
  ConjugateGradient<LatticeFermionD> CG(1.0e-8, 10000);
  CG(HermOp, so, sol);
  std::cout << "CG iterations " << CG.IterationsToComplete << " true residual " << CG.TrueResidual << std::endl;
  std::cout << "Benchmark_staggered (gridb200) done" << std::endl;
  Grid_finalize();
  return 0;
}
