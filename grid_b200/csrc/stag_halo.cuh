// stag_halo.cuh -- geometry and halo index arithmetic of the improved staggered operator (stag.cu).
//
// The Naik term reaches three sites along every direction (ref: displacements +-1 and +-3, Grid/qcd/action/fermion/
// instantiation/ImprovedStaggeredFermionInstantiation.cc:33-34), so on a decomposed lattice every split dimension carries
// three-deep halos of the input colour-vector field and of the gauge links that enter the double store
// (ref: StaggeredImpl.h:105-162 uses Cshift(U, mu, +1|+2|-1|-2|-3); Cshift is where the reference communicates).
//
// Everything here is __host__ __device__ on purpose: tests/host/stag_halo_emul.cu runs the very same functions on the CPU
// for emulated ranks (pack -> exchange -> neighbour lookup against global site ids), so the index logic of the kernels is
// checked where no GPU exists.
#pragma once
#include "internal.hpp"

#ifdef __CUDACC__
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

namespace gb {

constexpr int STAG_DEPTH = 3;   // halo depth in sites (Naik three-link term)

struct StagGeom {
  int L[4], Lxh, origin_parity;
  int64_t V4cb, hblk;
  FastDiv dLxh, dLy, dLz;
};
inline StagGeom stag_geom_of(const int ldims[4], const int origin[4]) {
  StagGeom G;
  for (int d = 0; d < 4; d++) G.L[d] = ldims[d];
  G.Lxh = G.L[0] / 2;
  G.origin_parity = (origin[0] + origin[1] + origin[2] + origin[3]) & 1;
  G.V4cb = (int64_t)ldims[0] * ldims[1] * ldims[2] * ldims[3] / 2; G.hblk = (G.V4cb + W - 1) / W;
  G.dLxh = FastDiv(G.Lxh); G.dLy = FastDiv(G.L[1]); G.dLz = FastDiv(G.L[2]);
  return G;
}
// (parity p, cb site) -> local coordinates.  ref: Cartesian_red_black.h:271-286 (x/2 fastest)
GB_HD void stag_coor(const StagGeom &G, int p, uint32_t site, int &x, int &y, int &z, int &t) {
  uint32_t r, xh, yy, zz;
  G.dLxh.divmod(site, r, xh); G.dLy.divmod(r, r, yy); G.dLz.divmod(r, r, zz);
  y = yy; z = zz; t = r;
  x = 2 * xh + ((p + G.origin_parity + y + z + t) & 1);
}
GB_HD uint32_t stag_cb(const StagGeom &G, int x, int y, int z, int t) {
  return (uint32_t)(x >> 1) + (uint32_t)G.Lxh * (y + G.L[1] * (z + G.L[2] * t));
}
// complex index of (site, colour) inside one parity block of a ColourVector field (also the layout of the halo buffers)
GB_HD size_t cv_index(uint32_t site, int c) { return ((size_t)(site >> LOGW) * 3 + c) * W + (site & (W - 1)); }

// ---- colour-vector halos.  One buffer per (mu, dir, input parity); element i = depth * nface + face index.
//   dir 0: consumed by forward legs  (receiver x_mu = L + depth   <- forward neighbour's slice depth)
//   dir 1: consumed by backward legs (receiver x_mu = depth - 3   <- backward neighbour's slice L - 3 + depth)
GB_HD uint32_t stag_nface(const StagGeom &G, int mu) { return (uint32_t)(G.V4cb / G.L[mu]); }
// cb sites of one parity inside a slice orthogonal to mu; for mu == 0 the parity constraint halves y
GB_HD uint32_t stag_face_index(const StagGeom &G, int mu, int x, int y, int z, int t) {
  if (mu == 0) return (uint32_t)(y >> 1) + (uint32_t)(G.L[1] >> 1) * (z + G.L[2] * t);
  if (mu == 1) return (uint32_t)(x >> 1) + (uint32_t)G.Lxh * (z + G.L[2] * t);
  if (mu == 2) return (uint32_t)(x >> 1) + (uint32_t)G.Lxh * (y + G.L[1] * t);
  return (uint32_t)(x >> 1) + (uint32_t)G.Lxh * (y + G.L[1] * z);
}
// inverse on the sending side: (slice sl along mu, face index, parity ip of the field being packed) -> coordinates
GB_HD void stag_face_coor(const StagGeom &G, int mu, int sl, uint32_t fi, int ip, int &x, int &y, int &z, int &t) {
  uint32_t r = fi;
  if (mu == 0) {
    const uint32_t Lyh = (uint32_t)(G.L[1] >> 1);
    const int yh = r % Lyh; r /= Lyh; z = r % G.L[2]; t = r / G.L[2];
    x = sl; y = 2 * yh + ((sl + ip + G.origin_parity + z + t) & 1);
    return;
  }
  const int xh = r % G.Lxh; r /= G.Lxh;
  if (mu == 1) { z = r % G.L[2]; t = r / G.L[2]; y = sl; }
  else if (mu == 2) { y = r % G.L[1]; t = r / G.L[1]; z = sl; }
  else { y = r % G.L[1]; z = r / G.L[1]; t = sl; }
  x = 2 * xh + ((ip + G.origin_parity + y + z + t) & 1);
}
// slice of the SENDER that fills depth d of the halo consumed in direction dir
GB_HD int stag_send_slice(const StagGeom &G, int mu, int dir, int d) { return dir == 0 ? d : G.L[mu] - STAG_DEPTH + d; }
// Where the neighbour of the site at c, displaced by disp (+-1, +-3) along mu, lives.  Returns -1 and the cb site index of the
// (periodically wrapped) local neighbour, or dir (0 forward / 1 backward) and the element index in halo (mu, dir).
GB_HD int stag_neighbour(const StagGeom &G, int comm_dim_mask, const int c[4], int mu, int disp, uint32_t &index) {
  int n[4] = {c[0], c[1], c[2], c[3]};
  const int L = G.L[mu];
  int nm = c[mu] + disp;
  if ((comm_dim_mask >> mu) & 1) {
    if (nm >= L) { index = (uint32_t)(nm - L) * stag_nface(G, mu) + stag_face_index(G, mu, c[0], c[1], c[2], c[3]); return 0; }
    if (nm < 0) { index = (uint32_t)(nm + STAG_DEPTH) * stag_nface(G, mu) + stag_face_index(G, mu, c[0], c[1], c[2], c[3]); return 1; }
  }
  if (nm >= L) nm -= L;
  if (nm < 0) nm += L;
  n[mu] = nm;
  index = stag_cb(G, n[0], n[1], n[2], n[3]);
  return -1;
}

// A site is "exterior" when at least one of its 16 neighbours lives in a halo: within STAG_DEPTH of a decomposed boundary.
// The overlapped hop computes the other ("interior") sites while the faces travel (ref: the interior / exterior split of
// ImprovedStaggeredFermion::DhopInternalOverlappedComms, ImprovedStaggeredFermionImplementation.h:283-335).
GB_HD bool stag_is_exterior(const StagGeom &G, int comm_dim_mask, const int c[4]) {
  bool ext = false;
  for (int mu = 0; mu < 4; mu++)
    if (((comm_dim_mask >> mu) & 1) && (c[mu] < STAG_DEPTH || c[mu] >= G.L[mu] - STAG_DEPTH)) ext = true;
  return ext;
}

// ---- gauge halos for the double store.  Links stay lexicographic; only U_mu is needed beyond the mu faces.
// One buffer per (mu, dir): [depth][lexicographic face, dimension mu removed][18 reals];
//   dir 0: slices 0..2 of the forward neighbour (x_mu = L + depth), dir 1: slices L-3..L-1 of the backward one (x_mu = depth - 3)
GB_HD uint32_t stag_gface_sites(const int L[4], int mu) { return (uint32_t)((int64_t)L[0] * L[1] * L[2] * L[3] / L[mu]); }
GB_HD uint32_t stag_gface_index(const int L[4], int mu, const int x[4]) {
  uint32_t fi = 0, st = 1;
  for (int d = 0; d < 4; d++) if (d != mu) { fi += st * (uint32_t)x[d]; st *= (uint32_t)L[d]; }
  return fi;
}
GB_HD void stag_gface_coor(const int L[4], int mu, int sl, uint32_t fi, int x[4]) {
  uint32_t r = fi;
  for (int d = 0; d < 4; d++) if (d != mu) { x[d] = r % L[d]; r /= L[d]; }
  x[mu] = sl;
}
// scalar offset (in reals) of link U_mu(x + d mu) for the site at local coordinate x: *where = -1 -> offset into the
// lexicographic field [V4][4][18]; 0 / 1 -> offset into gauge halo (mu, dir)
GB_HD size_t stag_link_offset(const int L[4], int comm_dim_mask, const int x[4], int mu, int d, int *where) {
  int q[4] = {x[0], x[1], x[2], x[3]};
  int nm = x[mu] + d;
  if ((comm_dim_mask >> mu) & 1) {
    if (nm >= L[mu]) { *where = 0; return ((size_t)(nm - L[mu]) * stag_gface_sites(L, mu) + stag_gface_index(L, mu, x)) * 18; }
    if (nm < 0) { *where = 1; return ((size_t)(nm + STAG_DEPTH) * stag_gface_sites(L, mu) + stag_gface_index(L, mu, x)) * 18; }
  }
  q[mu] = ((nm % L[mu]) + L[mu]) % L[mu];
  *where = -1;
  const size_t lex = q[0] + (size_t)L[0] * (q[1] + (size_t)L[1] * (q[2] + (size_t)L[2] * q[3]));
  return (lex * 4 + mu) * 18;
}

} // namespace gb
