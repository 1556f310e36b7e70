// next_kernels_emul.cu -- CPU emulation of the small kernels behind the SURVEY 8(f) rows (test infrastructure; built and run by
// tests/test_next_kernels_host.py with nvcc as a HOST program: no kernel is launched, no device is needed).
//
// It runs the per-element bodies the GPU kernels execute (grid_b200/csrc/next_kernels.cuh: chiral_wall_elem of schur.cu,
// insert_force_elem of force.cu) for every global thread index, on fields held in the DEVICE layout (blocked, checkerboarded;
// grid_b200/csrc/internal.hpp), and compares with expectations a Python test wrote from the oracle.  The lexicographic <-> device
// layout conversion below is an independent restatement of that layout (include/gridb200.h header comment + internal.hpp).
//
// usage: next_kernels_emul <f32|f64> Lx Ly Lz Lt Ls dir       with files in the directory `dir`:
//   src4.bin [V4][12] complex128, src5.bin / a5.bin [V4*Ls][12] complex128 (lexicographic, s fastest)
//   want_unphys.bin [V4*Ls][12], want_sol.bin, want_src.bin [V4][12]   (ImportUnphysicalFermion, ExportPhysical{Solution,Source})
//   want_force.bin [V4][4][9] complex128: force[x][mu] = sum_s trace_spin src5(x,s) a5(x,s)^dagger  (every mu the same input)
#include "../../grid_b200/csrc/next_kernels.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace gb;
typedef std::complex<double> cd;

static std::vector<cd> load(const std::string &p, size_t n) {
  std::vector<cd> v(n);
  FILE *f = std::fopen(p.c_str(), "rb");
  if (!f || std::fread(v.data(), sizeof(cd), n, f) != n) { std::fprintf(stderr, "cannot read %s\n", p.c_str()); std::exit(2); }
  std::fclose(f);
  return v;
}

template <class T> struct Layout {
  using P = Prec<T>;
  using V = typename P::vec;
  int L[4], Ls;
  int64_t V4cb, n5cb, hblk;
  Layout(const int *l, int ls) : Ls(ls) { for (int d = 0; d < 4; d++) L[d] = l[d]; V4cb = (int64_t)L[0] * L[1] * L[2] * L[3] / 2; n5cb = V4cb * Ls; hblk = (n5cb + W - 1) / W; }
  size_t nvec() const { return (size_t)2 * hblk * P::NV * W; }
  // scalar (T) offset of complex component c12 of 5D site (parity p, cb site, s)
  size_t off(int p, int64_t site, int s, int c12) const {
    constexpr int CPV = sizeof(T) == 4 ? 2 : 1;
    const int64_t i5 = site * Ls + s;
    const size_t vec = (size_t)p * hblk * P::NV * W + (((size_t)(i5 / W) * P::NV + c12 / CPV) * W) + i5 % W;
    return (vec * CPV + c12 % CPV) * 2;
  }
  // full-grid lexicographic [V4*Ls][12] -> device layout [even block][odd block]
  std::vector<V> to_device(const std::vector<cd> &lex) const {
    std::vector<V> d(nvec());
    std::fill((T *)d.data(), (T *)d.data() + d.size() * (16 / sizeof(T)), (T)0);
    T *raw = (T *)d.data();
    for (int t = 0; t < L[3]; t++) for (int z = 0; z < L[2]; z++) for (int y = 0; y < L[1]; y++) for (int x = 0; x < L[0]; x++) {
      const int p = (x + y + z + t) & 1;
      const int64_t site = (x >> 1) + (int64_t)(L[0] / 2) * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
      const int64_t i4 = x + (int64_t)L[0] * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
      for (int s = 0; s < Ls; s++) for (int c = 0; c < 12; c++) {
        const cd v = lex[(size_t)(i4 * Ls + s) * 12 + c];
        raw[off(p, site, s, c)] = (T)v.real(); raw[off(p, site, s, c) + 1] = (T)v.imag();
      }
    }
    return d;
  }
  std::vector<cd> to_lex(const std::vector<V> &d) const {
    std::vector<cd> lex((size_t)2 * V4cb * Ls * 12);
    const T *raw = (const T *)d.data();
    for (int t = 0; t < L[3]; t++) for (int z = 0; z < L[2]; z++) for (int y = 0; y < L[1]; y++) for (int x = 0; x < L[0]; x++) {
      const int p = (x + y + z + t) & 1;
      const int64_t site = (x >> 1) + (int64_t)(L[0] / 2) * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
      const int64_t i4 = x + (int64_t)L[0] * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
      for (int s = 0; s < Ls; s++) for (int c = 0; c < 12; c++) lex[(size_t)(i4 * Ls + s) * 12 + c] = cd(raw[off(p, site, s, c)], raw[off(p, site, s, c) + 1]);
    }
    return lex;
  }
};

static long g_bad = 0;
static void compare(const char *what, const std::vector<cd> &got, const std::vector<cd> &want, double tol) {
  double worst = 0;
  for (size_t i = 0; i < want.size(); i++) worst = std::max(worst, std::abs(got[i] - want[i]));
  std::printf("%-28s max |diff| %.3e (tol %.1e)\n", what, worst, tol);
  if (!(worst <= tol)) g_bad++;
}

template <class T> int run(const int *L, int Ls, const std::string &dir) {
  using V = typename Prec<T>::vec;
  const int64_t V4 = (int64_t)L[0] * L[1] * L[2] * L[3];
  const double eps = sizeof(T) == 4 ? 1e-6 : 1e-14;
  Layout<T> l4(L, 1), l5(L, Ls);
  auto src4 = load(dir + "/src4.bin", V4 * 12), src5 = load(dir + "/src5.bin", V4 * Ls * 12), a5 = load(dir + "/a5.bin", V4 * Ls * 12);
  // ---- chiral walls: ImportUnphysicalFermion (DIR 0, walls 0 / Ls-1), ExportPhysicalFermionSolution (DIR 1, Ls-1 / 0), ...Source (DIR 1, 0 / Ls-1)
  {
    std::vector<V> f4 = l4.to_device(src4), f5(l5.nvec());
    std::fill((T *)f5.data(), (T *)f5.data() + f5.size() * (16 / sizeof(T)), (T)0);
    const uint32_t n = (uint32_t)(l4.hblk * Prec<T>::NV * W * 2);
    for (uint32_t e = 0; e < n + 7; e++) chiral_wall_elem<T, 0>(e, f4.data(), f5.data(), 2, (uint32_t)l4.V4cb, (uint32_t)l4.hblk, (uint32_t)l5.hblk, Ls, 0, Ls - 1);
    compare("ImportUnphysicalFermion", l5.to_lex(f5), load(dir + "/want_unphys.bin", V4 * Ls * 12), 0.0 + (sizeof(T) == 4 ? 1e-7 : 0.0));
    std::vector<V> g5 = l5.to_device(src5), o4(l4.nvec());
    std::fill((T *)o4.data(), (T *)o4.data() + o4.size() * (16 / sizeof(T)), (T)0);
    for (uint32_t e = 0; e < n + 7; e++) chiral_wall_elem<T, 1>(e, o4.data(), g5.data(), 2, (uint32_t)l4.V4cb, (uint32_t)l4.hblk, (uint32_t)l5.hblk, Ls, Ls - 1, 0);
    compare("ExportPhysicalFermionSolution", l4.to_lex(o4), load(dir + "/want_sol.bin", V4 * 12), sizeof(T) == 4 ? 1e-7 : 0.0);
    for (uint32_t e = 0; e < n + 7; e++) chiral_wall_elem<T, 1>(e, o4.data(), g5.data(), 2, (uint32_t)l4.V4cb, (uint32_t)l4.hblk, (uint32_t)l5.hblk, Ls, 0, Ls - 1);
    compare("ExportPhysicalFermionSource", l4.to_lex(o4), load(dir + "/want_src.bin", V4 * 12), sizeof(T) == 4 ? 1e-7 : 0.0);
  }
  // ---- insert force: full grid (both parities in one pass), then parity by parity (the even-odd force terms) with sign -1
  {
    std::vector<V> b = l5.to_device(src5), a = l5.to_device(a5);
    auto want = load(dir + "/want_force.bin", V4 * 36);
    const size_t pstride = (size_t)l5.hblk * Prec<T>::NV * W;
    std::vector<T> mat((size_t)V4 * 72, (T)7);
    for (int mu = 0; mu < 4; mu++)
      for (uint32_t e = 0; e < 2u * (uint32_t)l5.V4cb * 9 + 5; e++) insert_force_elem<T>(e, mat.data(), b.data(), a.data(), Ls, L[0], L[1], L[2], 0, (uint32_t)l5.V4cb, pstride, mu, 0, 2, (T)1);
    std::vector<cd> got((size_t)V4 * 36);
    for (size_t i = 0; i < got.size(); i++) got[i] = cd(mat[2 * i], mat[2 * i + 1]);
    double scale = 0; for (auto &w : want) scale = std::max(scale, std::abs(w));
    compare("InsertForce5D (full grid)", got, want, eps * 10 * scale);
    std::fill(mat.begin(), mat.end(), (T)7);
    for (int p = 0; p < 2; p++) for (int mu = 0; mu < 4; mu++)
      for (uint32_t e = 0; e < (uint32_t)l5.V4cb * 9 + 5; e++)
        insert_force_elem<T>(e, mat.data(), b.data() + p * pstride, a.data() + p * pstride, Ls, L[0], L[1], L[2], 0, (uint32_t)l5.V4cb, pstride, mu, p, 1, (T)-1);
    for (size_t i = 0; i < got.size(); i++) got[i] = -cd(mat[2 * i], mat[2 * i + 1]);
    compare("InsertForce5D (per parity)", got, want, eps * 10 * scale);
  }
  return g_bad ? 1 : 0;
}

int main(int argc, char **argv) {
  if (argc != 8) { std::fprintf(stderr, "usage: %s <f32|f64> Lx Ly Lz Lt Ls dir\n", argv[0]); return 2; }
  int L[4]; for (int d = 0; d < 4; d++) L[d] = std::atoi(argv[2 + d]);
  const int Ls = std::atoi(argv[6]);
  return std::string(argv[1]) == "f32" ? run<float>(L, Ls, argv[7]) : run<double>(L, Ls, argv[7]);
}
