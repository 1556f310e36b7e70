// Test_dwf_mixedcg_prec-shaped driver on the B200-native library (ref: tests/Test_dwf_mixedcg_prec.cc:76-215).
// Mixed-precision CG against pure double CG on Shamir DWF, Ls=12, m=0.1, M5=1.8, tol 1e-8; asserts the reference's
// own criterion |x_mixed - x_double|^2 < 1e-4.  A user-written "Paranoid" Schur operator (ref :39-73) also drives the
// generic CG path through virtual HermOp.  usage: Test_dwf_mixedcg_prec [--grid x.y.z.t] [--Ls n]
#include "../include/gridb200.hpp"
#include <chrono>
#include <cstring>
#include <iostream>
using namespace gridb200;

template <class Matrix, class Field> class SchurDiagMooeeOperatorParanoid : public LinearOperatorBase<Field> {
public:
  Matrix &_Mat;
  explicit SchurDiagMooeeOperatorParanoid(Matrix &Mat) : _Mat(Mat) {}
  void Mpc(const Field &in, Field &out) {
    Field tmp(in.Grid()), tmp2(in.Grid());
    _Mat.Meooe(in, tmp); _Mat.MooeeInv(tmp, tmp2); _Mat.Meooe(tmp2, tmp); _Mat.Mooee(in, out);
    axpy(out, -1.0, tmp, out);
  }
  void MpcDag(const Field &in, Field &out) {
    Field tmp(in.Grid()), tmp2(in.Grid());
    _Mat.MeooeDag(in, tmp); _Mat.MooeeInvDag(tmp, tmp2); _Mat.MeooeDag(tmp2, tmp); _Mat.MooeeDag(in, out);
    axpy(out, -1.0, tmp, out);
  }
  void Op(const Field &in, Field &out) override { Mpc(in, out); }
  void AdjOp(const Field &in, Field &out) override { MpcDag(in, out); }
  void HermOp(const Field &in, Field &out) override { Field tmp(in.Grid()); Mpc(in, tmp); MpcDag(tmp, out); }
};

static double usecond() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  Grid_init(&argc, &argv);
  Coordinate latt4 = {8, 8, 8, 8}, mpi = {1, 1, 1, 1}, simd = {1, 1, 1, 1};
  int Ls = 12;
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
  DomainWallFermionF Ddwf_f(Umu_f, *FGrid, *FrbGrid, *UGrid, *UrbGrid, mass, M5);

  LatticeFermionD src_o(FrbGrid), result_o(FrbGrid), result_o_2(FrbGrid), result_o_3(FrbGrid);
  pickCheckerboard(Odd, src_o, src);
  result_o.SetCheckerboard(Odd); result_o.Zero();
  result_o_2.SetCheckerboard(Odd); result_o_2.Zero();
  result_o_3.SetCheckerboard(Odd); result_o_3.Zero();

  SchurDiagMooeeOperator<DomainWallFermionD, LatticeFermionD> HermOpEO(Ddwf);
  SchurDiagMooeeOperator<DomainWallFermionF, LatticeFermionF> HermOpEO_f(Ddwf_f);
  SchurDiagMooeeOperatorParanoid<DomainWallFermionD, LatticeFermionD> HermOpEO_paranoid(Ddwf);

  const double MdagMsiteflops = 1452, CGsiteflops = (8 + 4 + 8 + 4 + 4) * 3 * 4; // ref: Test_dwf_mixedcg_prec.cc:138-142
  std::cout << "::::::::::::: Starting mixed CG" << std::endl;
  MixedPrecisionConjugateGradient<LatticeFermionD, LatticeFermionF> mCG(1.0e-8, 10000, 50, FrbGrid, HermOpEO_f, HermOpEO);
  double t1 = usecond();
  mCG(src_o, result_o);
  double t2 = usecond();
  int iters = mCG.TotalInnerIterations;
  double flops = (MdagMsiteflops * 4 * FrbGrid->gSites() + CGsiteflops * FrbGrid->gSites()) * iters;
  std::cout << " MixedCG: inner " << iters << " outer " << mCG.TotalOuterIterations << " final " << mCG.TotalFinalStepIterations
            << " true residual " << mCG.TrueResidual << "  time " << (t2 - t1) * 1e-6 << " s  SinglePrecision GF/s " << flops / (t2 - t1) / 1000. << std::endl;

  std::cout << "::::::::::::: Starting regular CG" << std::endl;
  ConjugateGradient<LatticeFermionD> CG(1.0e-8, 10000);
  t1 = usecond();
  CG(HermOpEO, src_o, result_o_2);
  t2 = usecond();
  iters = CG.IterationsToComplete;
  flops = (MdagMsiteflops * 4 * FrbGrid->gSites() + CGsiteflops * FrbGrid->gSites()) * iters;
  std::cout << " DoubleCG: iterations " << iters << " true residual " << CG.TrueResidual << "  time " << (t2 - t1) * 1e-6
            << " s  DoublePrecision GF/s " << flops / (t2 - t1) / 1000. << std::endl;

  std::cout << "::::::::::::: CG through a user-written LinearOperatorBase (Paranoid)" << std::endl;
  ConjugateGradient<LatticeFermionD> CG2(1.0e-8, 10000);
  CG2(HermOpEO_paranoid, src_o, result_o_3);
  std::cout << " ParanoidCG: iterations " << CG2.IterationsToComplete << " true residual " << CG2.TrueResidual << std::endl;
  assert(CG2.IterationsToComplete == CG.IterationsToComplete);

  LatticeFermionD diff_o(FrbGrid);
  axpy(diff_o, -1.0, result_o_2, result_o); // diff = result_o - result_o_2
  RealD diff = norm2(diff_o);
  std::cout << "::::::::::::: Diff between mixed and regular CG: " << diff << std::endl;
  assert(diff < 1e-4); // ref: Test_dwf_mixedcg_prec.cc:212-215
  std::cout << "Test_dwf_mixedcg_prec (gridb200) done" << std::endl;
  Grid_finalize();
  return 0;
}
