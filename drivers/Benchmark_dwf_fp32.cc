// Benchmark_dwf_fp32-shaped driver on the B200-native library.
// Same flow as the reference's benchmarks/Benchmark_dwf_fp32.cc:151-447 (grids, random source, hot gauge field,
// DomainWallFermionF, 300 timed Dhop calls, 300 timed DhopEO calls, Deo+Doe == Dunprec), written against
// include/gridb200.hpp.  usage: Benchmark_dwf_fp32 [--grid x.y.z.t] [--Ls n] [--ncall n]
#include "../include/gridb200.hpp"
#include <chrono>
#include <cstring>
#include <iostream>
#include <string>
using namespace gridb200;

static double usecond() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  Grid_init(&argc, &argv);
  Coordinate latt4 = {16, 16, 16, 16}, mpi = {1, 1, 1, 1}, simd = {1, 1, 1, 1};
  int Ls = 16, ncall = 300;
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "--grid")) sscanf(argv[i + 1], "%d.%d.%d.%d", &latt4[0], &latt4[1], &latt4[2], &latt4[3]);
    if (!strcmp(argv[i], "--Ls")) Ls = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "--ncall")) ncall = atoi(argv[i + 1]);
  }
  const double single_site_flops = 8 * 3 * (7 + 16 * 3); // ref: Benchmark_dwf_fp32.cc:124

  GridCartesian *UGrid = SpaceTimeGrid::makeFourDimGrid(latt4, simd, mpi);
  GridRedBlackCartesian *UrbGrid = SpaceTimeGrid::makeFourDimRedBlackGrid(UGrid);
  GridCartesian *FGrid = SpaceTimeGrid::makeFiveDimGrid(Ls, UGrid);
  GridRedBlackCartesian *FrbGrid = SpaceTimeGrid::makeFiveDimRedBlackGrid(Ls, UGrid);

  GridParallelRNG RNG4(UGrid); RNG4.SeedFixedIntegers({1, 2, 3, 4});
  GridParallelRNG RNG5(FGrid); RNG5.SeedFixedIntegers({5, 6, 7, 8});

  LatticeFermionF src(FGrid); random(RNG5, src);
  {
    RealD N2 = 1.0 / std::sqrt(norm2(src));
    LatticeFermionF tmp(src);
    axpby(src, N2, 0.0, tmp, tmp);
  }
  LatticeFermionF result(FGrid); result.Zero();
  LatticeFermionF err(FGrid);
  LatticeGaugeFieldF Umu(UGrid);
  SU<3>::HotConfiguration(RNG4, Umu);
  std::cout << "Random gauge initialised" << std::endl;

  RealD mass = 0.1, M5 = 1.8;
  DomainWallFermionF Dw(Umu, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);
  Dw.ImportGauge(Umu);

  double volume = Ls;
  for (int mu = 0; mu < 4; mu++) volume *= latt4[mu];
  {
    Dw.Dhop(src, result, 0);
    gb_synchronize(Runtime::ctx());
    double t0 = usecond();
    for (int i = 0; i < ncall; i++) Dw.Dhop(src, result, 0);
    gb_synchronize(Runtime::ctx());
    double t1 = usecond();
    double flops = single_site_flops * volume * ncall;
    std::cout << "Called Dw " << ncall << " times in " << t1 - t0 << " us" << std::endl;
    std::cout << "mflop/s =   " << flops / (t1 - t0) << std::endl;
    std::cout << "norm result " << norm2(result) << std::endl;
  }

  LatticeFermionF src_e(FrbGrid), src_o(FrbGrid), r_e(FrbGrid), r_o(FrbGrid), r_eo(FGrid);
  pickCheckerboard(Even, src_e, src);
  pickCheckerboard(Odd, src_o, src);
  std::cout << "src_e " << norm2(src_e) << "  src_o " << norm2(src_o) << std::endl;
  {
    Dw.DhopEO(src_o, r_e, DaggerNo);
    gb_synchronize(Runtime::ctx());
    double t0 = usecond();
    for (int i = 0; i < ncall; i++) Dw.DhopEO(src_o, r_e, DaggerNo);
    gb_synchronize(Runtime::ctx());
    double t1 = usecond();
    double flops = (single_site_flops * volume * ncall) / 2.0;
    std::cout << "Deo mflop/s =   " << flops / (t1 - t0) << std::endl;
  }
  Dw.DhopEO(src_o, r_e, DaggerNo);
  Dw.DhopOE(src_e, r_o, DaggerNo);
  Dw.Dhop(src, result, DaggerNo);
  std::cout << "r_e " << norm2(r_e) << "  r_o " << norm2(r_o) << "  res " << norm2(result) << std::endl;
  setCheckerboard(r_eo, r_o);
  setCheckerboard(r_eo, r_e);
  axpy(err, -1.0, result, r_eo); // err = r_eo - result
  RealD n2e = norm2(err);
  std::cout << "norm diff   " << n2e << std::endl;
  assert(n2e < 1.0e-4); // ref: Benchmark_dwf_fp32.cc:438
  std::cout << "Benchmark_dwf_fp32 (gridb200) done" << std::endl;
  Grid_finalize();
  return 0;
}
