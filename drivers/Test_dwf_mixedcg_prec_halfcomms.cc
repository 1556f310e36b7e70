// Test_dwf_mixedcg_prec_halfcomms-shaped driver on the B200-native library (ref: tests/Test_dwf_mixedcg_prec_halfcomms.cc:40-122; the
// reference compiles its program out at :33-34).  The single-precision inner operator is DomainWallFermionFH -- fp32 arithmetic, halos
// one precision down -- inside MixedPrecisionConjugateGradient (InnerTolerance 3e-5, :87-88) and ConjugateGradientReliableUpdate
// (delta 0.1, :94-95); both must land on the plain double-precision CG solution, because every correction step is taken with the
// uncompressed fp64 operator.  On one rank the halo path (and with it the compression) only exists with GB_SELF_HALO=<dimension mask>.
// usage: [GB_SELF_HALO=12] Test_dwf_mixedcg_prec_halfcomms [--grid x.y.z.t] [--Ls n]
#include "../include/gridb200.hpp"
#include <cstring>
#include <iostream>
using namespace gridb200;

int main(int argc, char **argv) {
  Grid_init(&argc, &argv);
  Coordinate latt4 = {8, 8, 8, 8}, mpi = {1, 1, 1, 1}, simd = {1, 1, 1, 1};
  int Ls = 24;   // ref :44
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "--grid")) sscanf(argv[i + 1], "%d.%d.%d.%d", &latt4[0], &latt4[1], &latt4[2], &latt4[3]);
    if (!strcmp(argv[i], "--Ls")) Ls = atoi(argv[i + 1]);
  }
  GridCartesian *UGrid = SpaceTimeGrid::makeFourDimGrid(latt4, simd, mpi);
  GridRedBlackCartesian *UrbGrid = SpaceTimeGrid::makeFourDimRedBlackGrid(UGrid);
  GridCartesian *FGrid = SpaceTimeGrid::makeFiveDimGrid(Ls, UGrid);
  GridRedBlackCartesian *FrbGrid = SpaceTimeGrid::makeFiveDimRedBlackGrid(Ls, UGrid);
  GridParallelRNG RNG5(FGrid); RNG5.SeedFixedIntegers({5, 6, 7, 8});
  GridParallelRNG RNG4(UGrid); RNG4.SeedFixedIntegers({1, 2, 3, 4});

  LatticeFermionD src(FGrid); random(RNG5, src);
  LatticeGaugeFieldD Umu(UGrid);
  LatticeGaugeFieldF Umu_f(UGrid);
  SU<3>::HotConfiguration(RNG4, Umu);
  precisionChange(Umu_f, Umu);

  RealD mass = 0.1, M5 = 1.8;
  DomainWallFermionD Ddwf(Umu, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);
  DomainWallFermionFH Ddwf_f(Umu_f, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);

  LatticeFermionD src_o(FrbGrid), result_cg(FrbGrid), result_mcg(FrbGrid), result_rlcg(FrbGrid);
  pickCheckerboard(Odd, src_o, src);
  for (LatticeFermionD *r : {&result_cg, &result_mcg, &result_rlcg}) { r->SetCheckerboard(Odd); r->Zero(); }

  SchurDiagMooeeOperator<DomainWallFermionD, LatticeFermionD> HermOpEO(Ddwf);
  SchurDiagMooeeOperator<DomainWallFermionFH, LatticeFermionF> HermOpEO_f(Ddwf_f);

  std::cout << "Starting mixed CG" << std::endl;
  MixedPrecisionConjugateGradient<LatticeFermionD, LatticeFermionF> mCG(1.0e-8, 10000, 50, FrbGrid, HermOpEO_f, HermOpEO);
  mCG.InnerTolerance = 3.0e-5;
  mCG(src_o, result_mcg);
  std::cout << " MixedCG: inner " << mCG.TotalInnerIterations << " outer " << mCG.TotalOuterIterations << " final " << mCG.TotalFinalStepIterations
            << " true residual " << mCG.TrueResidual << std::endl;

  std::cout << "Starting reliable update CG" << std::endl;
  ConjugateGradientReliableUpdate<LatticeFermionD, LatticeFermionF> rlCG(1.e-8, 10000, 0.1, FrbGrid, HermOpEO_f, HermOpEO);
  rlCG(src_o, result_rlcg);
  std::cout << " ReliableUpdateCG: iterations " << rlCG.IterationsToComplete << " reliable updates " << rlCG.ReliableUpdatesPerformed
            << " true residual " << rlCG.TrueResidual << std::endl;

  std::cout << "Starting regular CG" << std::endl;
  ConjugateGradient<LatticeFermionD> CG(1.0e-8, 10000);
  CG(HermOpEO, src_o, result_cg);

  LatticeFermionD diff(FrbGrid);
  RealD vdiff_mcg = axpy_norm(diff, -1.0, result_cg, result_mcg);
  std::cout << "Diff between mixed and regular CG: " << vdiff_mcg << std::endl;       // ref :107-108
  RealD vdiff_rlcg = axpy_norm(diff, -1.0, result_cg, result_rlcg);
  std::cout << "Diff between reliable update and regular CG: " << vdiff_rlcg << std::endl;   // ref :113-114
  assert(mCG.TrueResidual < 1e-7 && rlCG.TrueResidual < 1e-7);
  assert(vdiff_mcg < 1e-4 && vdiff_rlcg < 1e-4);   // the bar of Test_dwf_mixedcg_prec.cc:212-215
  std::cout << "Test_dwf_mixedcg_prec_halfcomms (gridb200) done" << std::endl;
  Grid_finalize();
  return 0;
}
