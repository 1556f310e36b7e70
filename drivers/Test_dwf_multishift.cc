// ConjugateGradientMultiShift driver on the B200-native library, shaped like the double-precision leg of
// tests/solver/Test_dwf_multishift_mixedprec.cc:104-128 and tests/solver/Test_staggered_multishift.cc:98-107:
// one Krylov space solves (MpcDagMpc + pole_s) x_s = src_o for every pole, then each solution is checked by hand.
// The reference builds its poles with AlgRemez (GMP, not on this path); here they are given explicitly.
// usage: Test_dwf_multishift [--grid x.y.z.t] [--Ls n]
#include "../include/gridb200.hpp"
#include <cstring>
#include <iostream>
using namespace gridb200;
typedef LatticeFermionD FermionFieldD;

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
  GridParallelRNG RNG4(UGrid); RNG4.SeedFixedIntegers(std::vector<int>({1, 2, 3, 4}));
  FermionFieldD src(FGrid); random(RNG5, src);
  LatticeGaugeFieldD Umu(UGrid); SU<3>::HotConfiguration(RNG4, Umu);
  RealD mass = 0.01, M5 = 1.8;
  DomainWallFermionD Ddwf(Umu, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);
  FermionFieldD src_o(FrbGrid);
  pickCheckerboard(Odd, src_o, src);
  SchurDiagMooeeOperator<DomainWallFermionD, FermionFieldD> HermOpEO(Ddwf);

  const int order = 5;
  MultiShiftFunction shifts(order, 1e-4, 64.0);
  shifts.order = order; shifts.norm = 0.5;
  const double poles[order] = {1e-3, 1e-2, 0.1, 1.0, 10.0};
  for (int s = 0; s < order; s++) { shifts.poles[s] = poles[s]; shifts.residues[s] = 1.0 / (s + 1); shifts.tolerances[s] = 1e-8; }
  ConjugateGradientMultiShift<FermionFieldD> MSCG(10000, shifts);
  std::vector<FermionFieldD> results_o(order, FrbGrid);
  FermionFieldD psi(FrbGrid);
  MSCG(HermOpEO, src_o, results_o, psi);
  std::cout << "CGMultiShift: All shifts have converged iteration " << MSCG.IterationsToComplete << std::endl;

  // check each shift and the partial-fraction sum by hand
  FermionFieldD tmp(FrbGrid), sum(FrbGrid);
  GB_ASSERT_OK(gb_scale(sum.h, shifts.norm, src_o.h));
  const RealD ns = norm2(src_o);
  for (int s = 0; s < order; s++) {
    HermOpEO.HermOp(results_o[s], tmp);
    axpy(tmp, shifts.poles[s], results_o[s], tmp);
    axpy(tmp, -1.0, src_o, tmp);
    const RealD r = std::sqrt(norm2(tmp) / ns);
    std::cout << "CGMultiShift: shift[" << s << "] iterations " << MSCG.IterationsToCompleteShift[s] << " true residual " << MSCG.TrueResidualShift[s] << " by hand " << r << std::endl;
    assert(r < 1e-7 && std::fabs(r - MSCG.TrueResidualShift[s]) <= 1e-3 * r + 1e-14);
    if (s > 0) assert(MSCG.IterationsToCompleteShift[s] <= MSCG.IterationsToCompleteShift[s - 1]);   // heavier poles converge first
    axpy(sum, shifts.residues[s], results_o[s], sum);
  }
  axpy(sum, -1.0, psi, sum);
  assert(norm2(sum) <= 1e-24 * norm2(psi));
  std::cout << "Test_dwf_multishift: PASS" << std::endl;
  Grid_finalize();
  return 0;
}
