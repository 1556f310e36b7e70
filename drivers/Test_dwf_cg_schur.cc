// Test_dwf_cg_schur-shaped driver on the B200-native library (ref: tests/solver/Test_dwf_cg_schur.cc:46-77).
// Full-lattice solve  Ddwf result = src  through SchurRedBlackDiagMooeeSolve(ConjugateGradient), then the checks a
// propagator code makes around it: the unpreconditioned residual, and one 4D -> 5D -> solve -> 4D column
// (ImportPhysicalFermionSource / ExportPhysicalFermionSolution).  usage: Test_dwf_cg_schur [--grid x.y.z.t] [--Ls n]
#include "../include/gridb200.hpp"
#include <cstring>
#include <iostream>
using namespace gridb200;
typedef LatticeFermionD LatticeFermion;
typedef LatticeGaugeFieldD LatticeGaugeField;

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

  std::vector<int> seeds4({1, 2, 3, 4});
  std::vector<int> seeds5({5, 6, 7, 8});
  GridParallelRNG RNG5(FGrid); RNG5.SeedFixedIntegers(seeds5);
  GridParallelRNG RNG4(UGrid); RNG4.SeedFixedIntegers(seeds4);

  LatticeFermion src(FGrid); random(RNG5, src);
  LatticeFermion result(FGrid); result.Zero();
  LatticeGaugeField Umu(UGrid); SU<3>::HotConfiguration(RNG4, Umu);

  RealD mass = 0.1;
  RealD M5 = 1.8;
  DomainWallFermionD Ddwf(Umu, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);

  ConjugateGradient<LatticeFermion> CG(1.0e-8, 10000);
  SchurRedBlackDiagMooeeSolve<LatticeFermion> SchurSolver(CG);
  SchurSolver(Ddwf, src, result);
  std::cout << "ConjugateGradient Converged on iteration " << CG.IterationsToComplete << " True residual " << CG.TrueResidual << std::endl;
  std::cout << "SchurRedBlackBase solver true unprec resid " << SchurSolver.TrueUnprecResidual << std::endl;
  assert(SchurSolver.TrueUnprecResidual < 1.0e-6);

  // the residual again, by hand
  LatticeFermion resid(FGrid);
  Ddwf.M(result, resid);
  axpy(resid, -1.0, src, resid);
  RealD r = std::sqrt(norm2(resid) / norm2(src));
  std::cout << "|M result - src| / |src| = " << r << std::endl;
  assert(std::fabs(r - SchurSolver.TrueUnprecResidual) <= 1e-3 * r);

  // one propagator column: 4D source in, 4D solution out
  LatticeFermion src4(UGrid); random(RNG4, src4);
  LatticeFermion src5(FGrid), sol5(FGrid), sol4(UGrid);
  Ddwf.ImportPhysicalFermionSource(src4, src5);
  sol5.Zero();
  SchurSolver(Ddwf, src5, sol5);
  Ddwf.ExportPhysicalFermionSolution(sol5, sol4);
  std::cout << "propagator column: iterations " << CG.IterationsToComplete << " |sol4|^2 = " << norm2(sol4) << std::endl;
  assert(norm2(sol4) > 0.0);
  std::cout << "Test_dwf_cg_schur: PASS" << std::endl;
  Grid_finalize();
  return 0;
}
