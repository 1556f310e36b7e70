// next_kernels.cuh -- per-element bodies of the small kernels behind the SURVEY 8(f) rows (schur.cu, force.cu), written as
// __host__ __device__ functions of the global thread index so that tests/host/next_kernels_emul.cu can run exactly the code
// the GPU runs, on the CPU, against independently computed expectations (the same idea as stag_halo.cuh).
#pragma once
#include "internal.hpp"

#ifndef GB_HD
#ifdef __CUDACC__
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif
#endif

namespace gb {

// One element per 16-byte vec of the 4D field (all parity blocks).  Vec k of a spinor holds upper spin components (P+) for
// k < NV/2 and lower ones (P-) otherwise, so the chiral projectors are a choice of wall per vec, never arithmetic.
//   DIR 0: f5[s = s_up] <- upper half of f4, f5[s = s_lo] <- lower half   (f5 zeroed beforehand)
//   DIR 1: f4 upper half <- f5[s = s_up], f4 lower half <- f5[s = s_lo]
template <class T, int DIR>
GB_HD void chiral_wall_elem(uint32_t e, typename Prec<T>::vec *f4, typename Prec<T>::vec *f5, int nparity, uint32_t nsite4, uint32_t hblk4, uint32_t hblk5,
                            int Ls, int s_up, int s_lo) {
  using P = Prec<T>;
  const uint32_t per_block = hblk4 * P::NV * W;
  if (e >= per_block * (uint32_t)nparity) return;
  const uint32_t p = e / per_block, ep = e - p * per_block;
  const uint32_t lane = ep & (W - 1);
  const uint32_t r = ep >> LOGW;
  const uint32_t blk = r / P::NV, k = r - blk * P::NV;
  const uint32_t i4 = blk * W + lane;
  if (i4 >= nsite4) return;
  const uint32_t i5 = i4 * (uint32_t)Ls + (uint32_t)(k < P::NV / 2 ? s_up : s_lo);
  const size_t a5 = (size_t)p * hblk5 * P::NV * W + ((((size_t)(i5 >> LOGW)) * P::NV + k) << LOGW) + (i5 & (W - 1));
  if (DIR == 0) f5[a5] = f4[e];
  else f4[e] = f5[a5];
}

// mat[lex][mu][c1][c2] = sign * sum_s sum_spin Btilde(x,s)[spin][c1] * conj(A(x,s)[spin][c2])
// one element per (parity, cb site, c1*3 + c2) for npar parities starting at p0; fields in the blocked layout (internal.hpp):
// Bt / A point at the block of parity p0; mat is the full lexicographic gauge field (sites of other parities untouched)
template <class T>
GB_HD void insert_force_elem(uint32_t e, T *mat, const typename Prec<T>::vec *Bt, const typename Prec<T>::vec *A, int Ls, int Lx, int Ly, int Lz,
                             int origin_parity, uint32_t V4cb, size_t parity_stride /* vecs */, int mu, int p0, int npar, T sign) {
  using P = Prec<T>;
  if (e >= (uint32_t)npar * V4cb * 9) return;
  const uint32_t k9 = e % 9, sp_ = e / 9;
  const uint32_t pj = sp_ / V4cb, site = sp_ - pj * V4cb;
  const uint32_t p = (uint32_t)p0 + pj;
  const int c1 = k9 / 3, c2 = k9 - 3 * c1;
  const int Lxh = Lx / 2;
  uint32_t r = site;
  const int xh = r % Lxh; r /= Lxh;
  const int y = r % Ly; r /= Ly;
  const int z = r % Lz;
  const int t = r / Lz;
  const int x = 2 * xh + ((p + origin_parity + y + z + t) & 1);
  const size_t lex = x + (size_t)Lx * (y + (size_t)Ly * (z + (size_t)Lz * t));
  const T *bs = (const T *)(Bt + (size_t)pj * parity_stride);
  const T *as = (const T *)(A + (size_t)pj * parity_stride);
  constexpr int CPV = sizeof(T) == 4 ? 2 : 1;          // complex numbers per 16-byte vec
  T re = 0, im = 0;
  for (int s = 0; s < Ls; s++) {
    const uint32_t i5 = site * Ls + s;
    const size_t base = ((size_t)(i5 >> LOGW) * P::NV) << LOGW;   // vec index of element (i5, k = 0) minus the lane
    const uint32_t lane = i5 & (W - 1);
    for (int spin = 0; spin < 4; spin++) {
      const int kb = spin * 3 + c1, ka = spin * 3 + c2;            // complex component index 0..11
      const size_t ob = ((base + ((size_t)(kb / CPV) << LOGW) + lane) * CPV + (kb % CPV)) * 2;
      const size_t oa = ((base + ((size_t)(ka / CPV) << LOGW) + lane) * CPV + (ka % CPV)) * 2;
      const T br = bs[ob], bi = bs[ob + 1], ar = as[oa], ai = as[oa + 1];
      re += br * ar + bi * ai;     // b * conj(a)
      im += bi * ar - br * ai;
    }
  }
  T *m = mat + ((lex * 4 + mu) * 9 + k9) * 2;
  m[0] = sign * re; m[1] = sign * im;
}

} // namespace gb
