"""Every class template of the C++ mirror (include/gridb200.hpp) compiles when fully instantiated -- the drivers only touch part
of it.  Syntax / type check only (g++ -fsyntax-only): nothing is linked or run, so no GPU is needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
#include "gridb200.hpp"
using namespace gridb200;
typedef LatticeFermionD FD; typedef LatticeFermionF FF;
template class gridb200::FermionOperator<GB_F64>;
template class gridb200::CayleyFermion5DT<GB_F32>;
template class gridb200::DomainWallFermionT<GB_F64>;
template class gridb200::MobiusFermionT<GB_F32>;
template class gridb200::WilsonFermionT<GB_F64>;
template class gridb200::ImprovedStaggeredFermionT<GB_F32>;
template class gridb200::SchurDiagMooeeOperator<MobiusFermionD, FD>;
template class gridb200::MdagMLinearOperator<WilsonFermionD, FD>;
template class gridb200::SchurStaggeredOperator<ImprovedStaggeredFermionD, LatticeStaggeredFermionD>;
template class gridb200::ConjugateGradient<FD>;
template class gridb200::MixedPrecisionConjugateGradient<FD, FF>;
template class gridb200::MixedPrecisionConjugateGradientBatched<FD, FF>;
template class gridb200::ConjugateGradientReliableUpdate<FD, FF>;
template class gridb200::ConjugateGradientMultiShift<FD>;
template class gridb200::ConjugateGradientMultiShiftMixedPrec<FD, FF>;
void use(MobiusFermionD &D, ImprovedStaggeredFermionD &S, LatticeGaugeFieldD &U, FD &a, FD &b, LatticeStaggeredFermionD &c, LatticeStaggeredFermionD &d) {
  ConjugateGradient<FD> CG(1e-8, 100);
  SchurRedBlackDiagMooeeSolve<FD> s1(CG, true); s1(D, a, b);
  SchurRedBlackDiagMooeeSolve<FD> s2(CG); s2(D, a, b);
  ConjugateGradient<LatticeStaggeredFermionD> CGs(1e-8, 100);
  SchurRedBlackStaggeredSolve<LatticeStaggeredFermionD> s3(CGs, true); s3(S, c, d);
  SchurDifferentiableOperator<MobiusFermionD, FD> Sd(D); Sd.MpcDeriv(U, a, b); Sd.MpcDagDeriv(U, a, b);
  std::vector<FD> outs(8, a.Grid()); D.MdirAll(a, outs); D.Mdir(a, b, 1, -1); D.MDeriv(U, a, b, DaggerNo);
  FieldMetaData h; NerscIO::readConfiguration(U, h, "f"); NerscIO::writeConfiguration(U, "f", 1);
  D.ImportPhysicalFermionSource(a, b); D.ExportPhysicalFermionSolution(a, b); D.Dminus(a, b);
}
"""


def test_every_mirror_template_compiles(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = tmp_path / "mirror_check.cc"
    src.write_text(SNIPPET)
    p = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", str(src)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-4000:]
