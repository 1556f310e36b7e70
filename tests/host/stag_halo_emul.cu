// stag_halo_emul.cu -- CPU emulation of the three-deep staggered halos (test infrastructure; built and run by
// tests/test_stag_halo_host.py with nvcc as a HOST program: no kernel is launched, no device is needed).
//
// It drives the very __host__ __device__ index functions the CUDA kernels of grid_b200/csrc/stag.cu are built from
// (grid_b200/csrc/stag_halo.cuh) for every rank of an emulated processor grid:
//   colour-vector field:  pack (stag_face_coor) -> exchange (neighbour table of gb_geometry_query) -> 16-point neighbour lookup
//                         (stag_neighbour) must return the GLOBAL periodic neighbour of every site, both parities;
//   gauge links:          face gather (stag_gface_coor) -> exchange -> stag_link_offset for d = -3..+2 must return the global
//                         link U_mu(x + d mu)   (what StaggeredImpl::DoubleStore reads through Cshift, ref: StaggeredImpl.h:105-162).
// usage: stag_halo_emul GX GY GZ GT PX PY PZ PT     exit code 0 = every lookup correct
#include "../../grid_b200/csrc/stag_halo.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace gb;

struct Rank {
  int ldims[4], origin[4], nbr[8];
  StagGeom G;
  std::vector<double> field[2];                 // [parity][cv_index]
  std::vector<double> send[4][2][2], recv[4][2][2]; // [mu][dir][input parity]
  std::vector<double> U;                        // [lex][4][18]
  std::vector<double> gsend[4][2], grecv[4][2]; // [mu][dir]
};

int main(int argc, char **argv) {
  if (argc != 9) { std::fprintf(stderr, "usage: %s GX GY GZ GT PX PY PZ PT\n", argv[0]); return 2; }
  int gd[4], mpi[4];
  for (int d = 0; d < 4; d++) { gd[d] = std::atoi(argv[1 + d]); mpi[d] = std::atoi(argv[5 + d]); }
  const int nranks = mpi[0] * mpi[1] * mpi[2] * mpi[3];
  int mask = 0;
  for (int d = 0; d < 4; d++) if (mpi[d] > 1) mask |= 1 << d;
  auto gid = [&](const int g[4]) { return (double)(g[0] + gd[0] * (g[1] + gd[1] * (g[2] + (long)gd[2] * g[3]))); };
  std::vector<Rank> R(nranks);
  for (int r = 0; r < nranks; r++) {
    Rank &k = R[r];
    if (gb_geometry_query(gd, mpi, r, k.ldims, k.origin, k.nbr) != GB_OK) { std::fprintf(stderr, "geometry_query: %s\n", gb_last_error()); return 2; }
    k.G = stag_geom_of(k.ldims, k.origin);
    const StagGeom &G = k.G;
    // ---- colour-vector field: value = 3 * global site id + colour
    for (int p = 0; p < 2; p++) {
      k.field[p].assign((size_t)G.hblk * 3 * W, -1.0);
      for (uint32_t s = 0; s < (uint32_t)G.V4cb; s++) {
        int c[4]; stag_coor(G, p, s, c[0], c[1], c[2], c[3]);
        int g[4]; for (int d = 0; d < 4; d++) g[d] = c[d] + k.origin[d];
        if (((g[0] + g[1] + g[2] + g[3]) & 1) != p) { std::fprintf(stderr, "parity of site %u on rank %d\n", s, r); return 1; }
        if (stag_cb(G, c[0], c[1], c[2], c[3]) != s) { std::fprintf(stderr, "stag_cb o stag_coor != id\n"); return 1; }
        for (int col = 0; col < 3; col++) k.field[p][cv_index(s, col)] = 3 * gid(g) + col;
      }
    }
    // ---- pack
    for (int mu = 0; mu < 4; mu++) if ((mask >> mu) & 1) {
      const uint32_t nf = stag_nface(G, mu);
      for (int dir = 0; dir < 2; dir++) for (int ip = 0; ip < 2; ip++) {
        const size_t blocks = ((size_t)STAG_DEPTH * nf + W - 1) / W;
        k.send[mu][dir][ip].assign(blocks * 3 * W, -2.0);
        for (uint32_t i = 0; i < STAG_DEPTH * nf; i++) {
          const int d = i / nf; const uint32_t fi = i - d * nf;
          int x, y, z, t;
          stag_face_coor(G, mu, stag_send_slice(G, mu, dir, d), fi, ip, x, y, z, t);
          const uint32_t s = stag_cb(G, x, y, z, t);
          for (int col = 0; col < 3; col++) k.send[mu][dir][ip][cv_index(i, col)] = k.field[ip][cv_index(s, col)];
        }
      }
    }
    // ---- gauge: value = 72 * global site id + 18 * mu + k
    const int *L = k.ldims;
    const size_t V4 = (size_t)L[0] * L[1] * L[2] * L[3];
    k.U.assign(V4 * 72, -1.0);
    for (size_t lex = 0; lex < V4; lex++) {
      int x[4]; size_t q = lex;
      for (int d = 0; d < 4; d++) { x[d] = q % L[d]; q /= L[d]; }
      int g[4]; for (int d = 0; d < 4; d++) g[d] = x[d] + k.origin[d];
      for (int m = 0; m < 4; m++) for (int e = 0; e < 18; e++) k.U[(lex * 4 + m) * 18 + e] = 72 * gid(g) + 18 * m + e;
    }
    for (int mu = 0; mu < 4; mu++) if ((mask >> mu) & 1) {
      const uint32_t nf = stag_gface_sites(L, mu);
      for (int dir = 0; dir < 2; dir++) {
        k.gsend[mu][dir].assign((size_t)STAG_DEPTH * nf * 18, -2.0);
        for (uint32_t i = 0; i < STAG_DEPTH * nf; i++) {
          const int d = i / nf; const uint32_t fi = i - d * nf;
          int x[4]; stag_gface_coor(L, mu, dir == 0 ? d : L[mu] - STAG_DEPTH + d, fi, x);
          const size_t lex = x[0] + (size_t)L[0] * (x[1] + (size_t)L[1] * (x[2] + (size_t)L[2] * x[3]));
          for (int e = 0; e < 18; e++) k.gsend[mu][dir][(size_t)i * 18 + e] = k.U[(lex * 4 + mu) * 18 + e];
        }
      }
    }
  }
  // ---- exchange: halo dir 0 (forward legs) comes from the forward neighbour nbr[2 mu], dir 1 from the backward one nbr[2 mu + 1]
  for (int r = 0; r < nranks; r++)
    for (int mu = 0; mu < 4; mu++) if ((mask >> mu) & 1)
      for (int dir = 0; dir < 2; dir++) {
        const int from = R[r].nbr[2 * mu + dir];
        for (int ip = 0; ip < 2; ip++) R[r].recv[mu][dir][ip] = R[from].send[mu][dir][ip];
        R[r].grecv[mu][dir] = R[from].gsend[mu][dir];
      }
  // ---- consume
  long checked = 0, bad = 0;
  const int disps[4] = {1, -1, 3, -3};
  for (int r = 0; r < nranks; r++) {
    const Rank &k = R[r]; const StagGeom &G = k.G;
    for (int p = 0; p < 2; p++) for (uint32_t s = 0; s < (uint32_t)G.V4cb; s++) {
      int c[4]; stag_coor(G, p, s, c[0], c[1], c[2], c[3]);
      bool any_halo = false;
      for (int mu = 0; mu < 4; mu++) for (int di = 0; di < 4; di++) { uint32_t idx; if (stag_neighbour(G, mask, c, mu, disps[di], idx) >= 0) any_halo = true; }
      checked++;
      if (any_halo != stag_is_exterior(G, mask, c)) { if (bad++ < 5) std::fprintf(stderr, "exterior flag: rank %d p %d site %u\n", r, p, s); }
      for (int mu = 0; mu < 4; mu++) for (int di = 0; di < 4; di++) {
        uint32_t idx;
        const int where = stag_neighbour(G, mask, c, mu, disps[di], idx);
        int g[4]; for (int d = 0; d < 4; d++) g[d] = c[d] + k.origin[d];
        g[mu] = ((g[mu] + disps[di]) % gd[mu] + gd[mu]) % gd[mu];
        for (int col = 0; col < 3; col++) {
          const double got = where < 0 ? k.field[1 - p][cv_index(idx, col)] : k.recv[mu][where][1 - p][cv_index(idx, col)];
          checked++;
          if (got != 3 * gid(g) + col) { if (bad++ < 5) std::fprintf(stderr, "field: rank %d p %d site %u mu %d disp %d: got %.0f want %.0f\n", r, p, s, mu, disps[di], got, 3 * gid(g) + col); }
        }
      }
    }
    const int *L = k.ldims;
    const size_t V4 = (size_t)L[0] * L[1] * L[2] * L[3];
    for (size_t lex = 0; lex < V4; lex++) {
      int x[4]; size_t q = lex;
      for (int d = 0; d < 4; d++) { x[d] = q % L[d]; q /= L[d]; }
      for (int mu = 0; mu < 4; mu++) for (int d = -3; d <= 2; d++) {
        int where;
        const size_t off = stag_link_offset(L, mask, x, mu, d, &where);
        int g[4]; for (int e = 0; e < 4; e++) g[e] = x[e] + k.origin[e];
        g[mu] = ((g[mu] + d) % gd[mu] + gd[mu]) % gd[mu];
        for (int e = 0; e < 18; e += 17) {
          const double got = where < 0 ? k.U[off + e] : k.grecv[mu][where][off + e];
          checked++;
          if (got != 72 * gid(g) + 18 * mu + e) { if (bad++ < 10) std::fprintf(stderr, "gauge: rank %d lex %zu mu %d d %d: got %.0f want %.0f\n", r, lex, mu, d, got, 72 * gid(g) + 18 * mu + e); }
        }
      }
    }
  }
  std::printf("stag_halo_emul %d.%d.%d.%d on %d.%d.%d.%d: %ld lookups, %ld wrong\n", gd[0], gd[1], gd[2], gd[3], mpi[0], mpi[1], mpi[2], mpi[3], checked, bad);
  return bad ? 1 : 0;
}
