// stag.cu -- improved staggered fermions (one-link + Naik three-link hopping term), hand-written for sm_100a.
//
// Replaces (paths relative to the reference tree):
//   StaggeredImpl::DoubleStore (phases eta_mu, U and UUU, forward and backward)   Grid/qcd/action/fermion/StaggeredImpl.h:105-162
//   ImprovedStaggeredFermion::ImportGauge (c1/u0, c2/u0^3 scale factors)          implementation/ImprovedStaggeredFermionImplementation.h:137-167
//   StaggeredKernels::DhopSiteGeneric / DhopImproved (16-point stencil)            implementation/StaggeredKernelsImplementation.h:73-121,260-330
//   ImprovedStaggeredFermion::Dhop/DhopOE/DhopEO/M/Mdag/Meooe/Mooee/MooeeInv       implementation/ImprovedStaggeredFermionImplementation.h:173-247,390-470
//   SchurStaggeredOperator::Mpc / HermOp                                           Grid/algorithms/LinearOperator.h:543-584
// Design (not a port): the kernel is a STREAM over the 16 double-stored links of each output site (1152 B/site in fp32
// against 24 B of colour vector in and out), so links are laid out per output parity as
// [block of 32 sites][72 float4 | 144 double2][32 lanes]: every load of a warp is one 512-byte row, each link byte is read
// exactly once, and the 16 neighbour colour vectors (which fit in L2: 64 MB per parity at 48^4) are gathered with arithmetic
// addressing -- no stencil table (ref: Grid/stencil/Stencil.h:79-136).
#include "comm.hpp"
#include "fermop.hpp"
#include "kernels_common.cuh"
#include "stag_halo.cuh"
#include <cstdlib>
#include <vector>

namespace gb {

constexpr int SW = 32;      // sites per link block
constexpr int SLOG = 5;

template <class T> struct CT;
template <> struct CT<float> { using c = float2; using lv = float4; static constexpr int LVN = 72; };
template <> struct CT<double> { using c = double2; using lv = double2; static constexpr int LVN = 144; };
__device__ __forceinline__ float2 mkc(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 mkc(double a, double b) { return make_double2(a, b); }

static StagGeom stag_geom(const gb_grid *g) { return stag_geom_of(g->ldims, g->origin); }   // StagGeom and its index functions: stag_halo.cuh

// =====================================================================================================
// host <-> device, random, precision change for ColourVector fields
// =====================================================================================================
template <class TD, class TH, int DIR>
__global__ void stag_transfer_kernel(typename CT<TD>::c *dev, TH *host, StagGeom G, int nparity, int full) {
  const int64_t n = (int64_t)nparity * G.hblk * W * 3;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  const int c = r % 3;
  const int64_t blk = r / 3;
  const int p = blk / G.hblk;
  const int64_t site = (blk - (int64_t)p * G.hblk) * W + lane;
  if (site >= G.V4cb) { if (DIR == 0) dev[e] = mkc((TD)0, (TD)0); return; }
  int64_t hidx = site;
  if (full) {
    int x, y, z, t;
    stag_coor(G, p, (uint32_t)site, x, y, z, t);
    hidx = x + (int64_t)G.L[0] * (y + (int64_t)G.L[1] * (z + (int64_t)G.L[2] * t));
  }
  TH *h = host + (hidx * 3 + c) * 2;
  if (DIR == 0) dev[e] = mkc((TD)h[0], (TD)h[1]);
  else { auto v = dev[e]; h[0] = (TH)v.x; h[1] = (TH)v.y; }
}
void stag_transfer(const gb_fermion *f, void *stage, int host_prec, int dir) {
  gb_context *ctx = f->grid->ctx;
  StagGeom G = stag_geom(f->grid);
  const int64_t n = (int64_t)f->nparity * G.hblk * W * 3;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const int full = f->kind == GB_FULL;
#define GB_L(TD, TH)                                                                                                               \
  do {                                                                                                                             \
    if (dir == 0) stag_transfer_kernel<TD, TH, 0><<<blocks, 256, 0, ctx->stream>>>((typename CT<TD>::c *)f->data, (TH *)stage, G, f->nparity, full); \
    else stag_transfer_kernel<TD, TH, 1><<<blocks, 256, 0, ctx->stream>>>((typename CT<TD>::c *)f->data, (TH *)stage, G, f->nparity, full);          \
  } while (0)
  if (f->prec == GB_F32 && host_prec == GB_F32) GB_L(float, float);
  else if (f->prec == GB_F32) GB_L(float, double);
  else if (host_prec == GB_F32) GB_L(double, float);
  else GB_L(double, double);
#undef GB_L
  count_launch(ctx);
  check_launch(ctx, "stag_transfer");
}

template <class TD>
__global__ void stag_random_kernel(typename CT<TD>::c *dev, StagGeom G, int nparity, int cb_half, int4 go, int4 gd, uint64_t seed) {
  const int64_t n = (int64_t)nparity * G.hblk * W * 3;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  const int c = r % 3;
  const int64_t blk = r / 3;
  const int p = blk / G.hblk;
  const int64_t site = (blk - (int64_t)p * G.hblk) * W + lane;
  if (site >= G.V4cb) { dev[e] = mkc((TD)0, (TD)0); return; }
  int x, y, z, t;
  stag_coor(G, nparity == 2 ? p : cb_half, (uint32_t)site, x, y, z, t);
  const uint64_t g4 = (uint64_t)(x + go.x) + (uint64_t)gd.x * ((y + go.y) + (uint64_t)gd.y * ((z + go.z) + (uint64_t)gd.z * (t + go.w)));
  const uint64_t key = splitmix64(seed ^ 0x5741474745524544ull);
  dev[e] = mkc((TD)u01(splitmix64(key ^ (g4 * 6 + 2 * c))), (TD)u01(splitmix64(key ^ (g4 * 6 + 2 * c + 1))));
}
void stag_random(gb_fermion *f, uint64_t seed) {
  gb_context *ctx = f->grid->ctx;
  const gb_grid *g = f->grid;
  StagGeom G = stag_geom(g);
  int4 go = make_int4(g->origin[0], g->origin[1], g->origin[2], g->origin[3]);
  int4 gd = make_int4(g->gdims[0], g->gdims[1], g->gdims[2], g->gdims[3]);
  const int64_t n = (int64_t)f->nparity * G.hblk * W * 3;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (f->prec == GB_F32) stag_random_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float2 *)f->data, G, f->nparity, f->cb, go, gd, seed);
  else stag_random_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double2 *)f->data, G, f->nparity, f->cb, go, gd, seed);
  count_launch(ctx);
  check_launch(ctx, "stag_random");
}
template <class CO, class CI> __global__ void stag_prec_kernel(CO *o, const CI *i, int64_t n) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) { CI v = i[e]; CO w; w.x = v.x; w.y = v.y; o[e] = w; }
}
void stag_precision_change(gb_fermion *out, const gb_fermion *in) {
  gb_context *ctx = in->grid->ctx;
  const int64_t n = (int64_t)in->nparity * in->hblk * W * 3;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (out->prec == GB_F32) stag_prec_kernel<float2, double2><<<blocks, 256, 0, ctx->stream>>>((float2 *)out->data, (const double2 *)in->data, n);
  else stag_prec_kernel<double2, float2><<<blocks, 256, 0, ctx->stream>>>((double2 *)out->data, (const float2 *)in->data, n);
  count_launch(ctx);
  check_launch(ctx, "stag_precision_change");
}

// =====================================================================================================
// DoubleStore: lexicographic thin/fat links -> per-parity streams of 16 scaled, phased links per site
// record of one site: link l = pass*8 + mu*2 + dirbit (pass 0 fat one-link, 1 Naik; dirbit 0 forward, 1 backward), complex
// index l*9 + row*3 + col; fp32 packs two complex per float4.
// =====================================================================================================
template <class T> struct M3 { T re[9], im[9]; };
template <class T> __device__ __forceinline__ M3<T> m3_load_p(const T *p) {
  M3<T> m;
#pragma unroll
  for (int k = 0; k < 9; k++) { m.re[k] = p[2 * k]; m.im[k] = p[2 * k + 1]; }
  return m;
}
template <class T> __device__ __forceinline__ M3<T> m3_mul(const M3<T> &a, const M3<T> &b) {
  M3<T> c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      T re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) { re += a.re[3 * i + k] * b.re[3 * k + j] - a.im[3 * i + k] * b.im[3 * k + j]; im += a.re[3 * i + k] * b.im[3 * k + j] + a.im[3 * i + k] * b.re[3 * k + j]; }
      c.re[3 * i + j] = re; c.im[3 * i + j] = im;
    }
  return c;
}
template <class T> __device__ __forceinline__ M3<T> m3_adj(const M3<T> &a) {
  M3<T> c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { c.re[3 * i + j] = a.re[3 * j + i]; c.im[3 * i + j] = -a.im[3 * j + i]; }
  return c;
}
template <class T> __device__ __forceinline__ void link_store(T *rec_base /* site record, element stride = SW lanes */, int l, const M3<T> &m, T f) {
  // scalar view of the record: complex k of the site lives at scalar offsets ((k*2)/VW)*SW*VW + (k*2)%VW (+1), VW = scalars per vec
  constexpr int VW = sizeof(T) == 4 ? 4 : 2;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int s = (l * 9 + k) * 2;
    T *q = rec_base + (size_t)(s / VW) * SW * VW + (s % VW);
    q[0] = f * m.re[k]; q[1] = f * m.im[k];
  }
}
// three-deep faces of U_mu beyond the mu boundaries of a decomposed lattice (layout: stag_halo.cuh); all null on one rank
template <class T> struct StagGaugeHalo { const T *thin[4][2], *fat[4][2]; int comm_dim_mask; };
template <class T>
__global__ void stag_double_store_kernel(T *links, size_t parity_stride /* scalars */, const T *Uthin, const T *Ufat, const StagGaugeHalo<T> H, StagGeom G, int4 go, T f1, T f3) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= G.V4cb * 2 * 4) return;
  const int mu = e & 3;
  const int64_t sp = e >> 2;
  const int p = sp / G.V4cb;
  const uint32_t site = (uint32_t)(sp - (int64_t)p * G.V4cb);
  int c[4];
  stag_coor(G, p, site, c[0], c[1], c[2], c[3]);
  // U_mu(x + d mu): local (periodic wrap in undecomposed dimensions) or from the neighbour's faces
  auto thin = [&](int d) { int w; const size_t o = stag_link_offset(G.L, H.comm_dim_mask, c, mu, d, &w); return m3_load_p(w < 0 ? Uthin + o : H.thin[mu][w] + o); };
  auto fat = [&](int d) { int w; const size_t o = stag_link_offset(G.L, H.comm_dim_mask, c, mu, d, &w); return m3_load_p(w < 0 ? Ufat + o : H.fat[mu][w] + o); };
  const int gx = c[0] + go.x, gy = c[1] + go.y, gz = c[2] + go.z;
  const int par = mu == 0 ? 0 : mu == 1 ? gx : mu == 2 ? gx + gy : gx + gy + gz;
  const T eta = (par & 1) ? (T)-1 : (T)1;
  constexpr int VW = sizeof(T) == 4 ? 4 : 2;
  constexpr int LVN = CT<T>::LVN;
  T *rec = links + (size_t)p * parity_stride + ((size_t)(site >> SLOG) * LVN * SW + (site & (SW - 1))) * VW;
  link_store(rec, mu * 2 + 0, fat(0), eta * f1);
  link_store(rec, mu * 2 + 1, m3_adj(fat(-1)), -eta * f1);
  const M3<T> fwd = m3_mul(thin(0), m3_mul(thin(1), thin(2)));
  const M3<T> bwd = m3_adj(m3_mul(thin(-3), m3_mul(thin(-2), thin(-1))));
  link_store(rec, 8 + mu * 2 + 0, fwd, eta * f3);
  link_store(rec, 8 + mu * 2 + 1, bwd, -eta * f3);
}

// One halo message: my `send` buffer goes to rank `to`, `recv` is filled by rank `from`.  A neighbour that is this rank itself
// (undecomposed dimension whose halos are forced on, see GB_STAG_SELF_HALO) is a device-to-device copy, like the reference's
// comms-to-self path (ref: Grid/communicator/Communicator_none.cc SendToRecvFrom, Cshift_common.h local branch).
struct HaloMsg { const void *send; void *recv; size_t bytes; int to, from; };
static void stag_sendrecv(gb_context *ctx, const std::vector<HaloMsg> &msgs, cudaStream_t st) {
  bool remote = false;
  for (const HaloMsg &m : msgs) {
    if (m.to == ctx->rank && m.from == ctx->rank) GB_CUDA(cudaMemcpyAsync(m.recv, m.send, m.bytes, cudaMemcpyDeviceToDevice, st));
    else remote = true;
  }
  if (!remote) return;
  GB_REQUIRE(ctx->nccl != nullptr, "decomposed lattice: call gb_comm_init first");
  NcclApi &N = nccl();
  nccl_check(N.GroupStart(), "ncclGroupStart");
  for (const HaloMsg &m : msgs) if (!(m.to == ctx->rank && m.from == ctx->rank)) {
    nccl_check(N.Send(m.send, m.bytes, ncclChar, m.to, ctx->nccl, st), "ncclSend");
    nccl_check(N.Recv(m.recv, m.bytes, ncclChar, m.from, ctx->nccl, st), "ncclRecv");
  }
  nccl_check(N.GroupEnd(), "ncclGroupEnd");
}

// gather the three slices of U_mu next to a mu boundary (dir 0: x_mu = 0..2, dir 1: x_mu = L-3..L-1) in the halo layout
template <class T> __global__ void stag_gauge_face_kernel(const T *Ulex, T *face, StagGeom G, int mu, int dir, uint32_t nface) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (uint32_t)STAG_DEPTH * nface * 18) return;
  const uint32_t i = e / 18, k = e - i * 18;
  const uint32_t d = i / nface, fi = i - d * nface;
  int x[4];
  stag_gface_coor(G.L, mu, dir == 0 ? (int)d : G.L[mu] - STAG_DEPTH + (int)d, fi, x);
  const size_t lex = x[0] + (size_t)G.L[0] * (x[1] + (size_t)G.L[1] * (x[2] + (size_t)G.L[2] * x[3]));
  face[e] = Ulex[(lex * 4 + mu) * 18 + k];
}

void stag_import_gauge(gb_fermop *op, const gb_gauge *Uthin, const gb_gauge *Ufat) {
  GB_REQUIRE(Uthin && Ufat && Uthin->grid == op->grid && Ufat->grid == op->grid, "gauge fields live on another grid");
  GB_REQUIRE(Uthin->prec == op->prec && Ufat->prec == op->prec, "gauge precision differs from the operator's");
  const gb_grid *g = op->grid;
  for (int d = 0; d < 4; d++) GB_REQUIRE(g->ldims[d] >= 4, "the Naik term needs local extents >= 4 (ref: Stencil.h:709 assert(abs(displacement)<ld))");
  gb_context *ctx = op->ctx;
  GB_CUDA(cudaSetDevice(ctx->device));
  const size_t nblk = (size_t)((g->V4cb + SW - 1) / SW);
  const size_t parity_bytes = nblk * SW * 16 * (op->prec == GB_F32 ? CT<float>::LVN : CT<double>::LVN);
  if (op->stag_links == nullptr) {
    GB_CUDA(cudaMalloc(&op->stag_links, 2 * parity_bytes));
    GB_CUDA(cudaMemsetAsync(op->stag_links, 0, 2 * parity_bytes, ctx->stream));
    op->stag_parity_bytes = parity_bytes;
  }
  StagGeom G = stag_geom(g);
  int4 go = make_int4(g->origin[0], g->origin[1], g->origin[2], g->origin[3]);
  // decomposed lattice: three slices of U_mu (thin and fat) from both mu neighbours, the reference's Cshift(U, mu, -3..+2)
  // [which: 0 thin, 1 fat][mu][dir]; dir 0 = my slices 0..2, consumed by the backward neighbour's x + d lookups
  void *gsend[2][4][2] = {}, *grecv[2][4][2] = {};
  const size_t esz = op->prec == GB_F32 ? 4 : 8;
  if (op->comm_dim_mask) {
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
      const uint32_t nface = stag_gface_sites(G.L, mu);
      const size_t bytes = (size_t)STAG_DEPTH * nface * 18 * esz;
      const unsigned fb = (unsigned)(((size_t)STAG_DEPTH * nface * 18 + 255) / 256);
      for (int which = 0; which < 2; which++) for (int dir = 0; dir < 2; dir++) {
        GB_CUDA(cudaMalloc(&gsend[which][mu][dir], bytes));
        GB_CUDA(cudaMalloc(&grecv[which][mu][dir], bytes));
        const void *U = which == 0 ? Uthin->data : Ufat->data;
        if (op->prec == GB_F32) stag_gauge_face_kernel<float><<<fb, 256, 0, ctx->stream>>>((const float *)U, (float *)gsend[which][mu][dir], G, mu, dir, nface);
        else stag_gauge_face_kernel<double><<<fb, 256, 0, ctx->stream>>>((const double *)U, (double *)gsend[which][mu][dir], G, mu, dir, nface);
        count_launch(ctx);
      }
    }
    check_launch(ctx, "stag_gauge_face");
    std::vector<HaloMsg> msgs;
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
      const size_t bytes = (size_t)STAG_DEPTH * stag_gface_sites(G.L, mu) * 18 * esz;
      for (int which = 0; which < 2; which++) {
        // halo dir 0 (x_mu >= L) holds the FORWARD neighbour's first slices: mine travel backward, and vice versa
        msgs.push_back({gsend[which][mu][0], grecv[which][mu][0], bytes, g->nbr_rank[mu][1], g->nbr_rank[mu][0]});
        msgs.push_back({gsend[which][mu][1], grecv[which][mu][1], bytes, g->nbr_rank[mu][0], g->nbr_rank[mu][1]});
      }
    }
    stag_sendrecv(ctx, msgs, ctx->stream);
  }
  const int64_t n = g->V4cb * 2 * 4;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  const double f1 = 0.5 * op->stag_c1 / op->stag_u0, f3 = 0.5 * op->stag_c2 / (op->stag_u0 * op->stag_u0 * op->stag_u0);
  auto run = [&](auto tag) {
    using T = decltype(tag);
    StagGaugeHalo<T> H;
    H.comm_dim_mask = op->comm_dim_mask;
    for (int mu = 0; mu < 4; mu++) for (int dir = 0; dir < 2; dir++) { H.thin[mu][dir] = (const T *)grecv[0][mu][dir]; H.fat[mu][dir] = (const T *)grecv[1][mu][dir]; }
    stag_double_store_kernel<T><<<blocks, 128, 0, ctx->stream>>>((T *)op->stag_links, parity_bytes / sizeof(T), (const T *)Uthin->data, (const T *)Ufat->data, H, G, go, (T)f1, (T)f3);
  };
  if (op->prec == GB_F32) run(float()); else run(double());
  count_launch(ctx);
  check_launch(ctx, "stag_double_store");
  if (op->comm_dim_mask) {
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int which = 0; which < 2; which++) for (int mu = 0; mu < 4; mu++) for (int dir = 0; dir < 2; dir++) { cudaFree(gsend[which][mu][dir]); cudaFree(grecv[which][mu][dir]); }
  }
}

// =====================================================================================================
// the hopping term
// =====================================================================================================
template <class T> struct StagArgs {
  const typename CT<T>::c *in[2];
  typename CT<T>::c *out[2];
  const typename CT<T>::lv *U[2];
  const typename CT<T>::c *ax[2]; // optional epilogue: out = hop + axb * ax   (M = Dhop + mass)
  T axb;
  StagGeom G;
  int first_parity;
  // decomposed lattices: received three-deep halos of the input field, point = mu + 4 * dir (layout: stag_halo.cuh)
  const typename CT<T>::c *halo[8];
  size_t halo_parity_stride[4];   // complex numbers between the two input-parity slots of a buffer
  int comm_dim_mask;
};

template <class T> struct CV { T re[3], im[3]; };
__device__ __forceinline__ void cv_load(CV<float> &v, const float2 *f, uint32_t site) {
#pragma unroll
  for (int c = 0; c < 3; c++) { float2 q = __ldg(f + cv_index(site, c)); v.re[c] = q.x; v.im[c] = q.y; }
}
__device__ __forceinline__ void cv_load(CV<double> &v, const double2 *f, uint32_t site) {
#pragma unroll
  for (int c = 0; c < 3; c++) { double2 q = __ldg(f + cv_index(site, c)); v.re[c] = q.x; v.im[c] = q.y; }
}
// o += M x, M given as 9 complex (re,im) pairs
template <class T> __device__ __forceinline__ void mv_add(CV<T> &o, const T (&m)[18], const CV<T> &x) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const T mr = m[2 * (3 * r + c)], mi = m[2 * (3 * r + c) + 1];
      o.re[r] = fma(mr, x.re[c], o.re[r]); o.re[r] = fma(-mi, x.im[c], o.re[r]);
      o.im[r] = fma(mr, x.im[c], o.im[r]); o.im[r] = fma(mi, x.re[c], o.im[r]);
    }
}
// the two links (forward, backward) of leg pair q = pass*4 + mu: 18 complex = 9 float4 | 18 double2
__device__ __forceinline__ void pair_load(float (&a)[18], float (&b)[18], const float4 *rec, int q) {
  float s[36];
#pragma unroll
  for (int j = 0; j < 9; j++) { const float4 v = __ldcs(rec + (size_t)(q * 9 + j) * SW); s[4 * j] = v.x; s[4 * j + 1] = v.y; s[4 * j + 2] = v.z; s[4 * j + 3] = v.w; }
#pragma unroll
  for (int k = 0; k < 18; k++) { a[k] = s[k]; b[k] = s[18 + k]; }
}
__device__ __forceinline__ void pair_load(double (&a)[18], double (&b)[18], const double2 *rec, int q) {
#pragma unroll
  for (int j = 0; j < 9; j++) { const double2 v = __ldcs(rec + (size_t)(q * 18 + j) * SW); a[2 * j] = v.x; a[2 * j + 1] = v.y; }
#pragma unroll
  for (int j = 0; j < 9; j++) { const double2 v = __ldcs(rec + (size_t)(q * 18 + 9 + j) * SW); b[2 * j] = v.x; b[2 * j + 1] = v.y; }
}

// one thread per halo element: copy the colour vectors of the three boundary slices into the send buffer (no projection to
// do for staggered fields: the compressor of the reference is the identity here, ref: Grid/stencil/SimpleCompressor.h:22-37)
template <class T>
__global__ void stag_pack_kernel(const typename CT<T>::c *__restrict__ in, typename CT<T>::c *__restrict__ buf, StagGeom G, int mu, int dir, int ip, uint32_t nface) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)STAG_DEPTH * nface) return;
  const uint32_t d = i / nface, fi = i - d * nface;
  int x, y, z, t;
  stag_face_coor(G, mu, stag_send_slice(G, mu, dir, (int)d), fi, ip, x, y, z, t);
  const uint32_t site = stag_cb(G, x, y, z, t);
#pragma unroll
  for (int c = 0; c < 3; c++) buf[cv_index(i, c)] = __ldg(in + cv_index(site, c));
}

// COMM 0: one rank, periodic wrap.  COMM 1: every site, halo lookups (serial comms).  COMM 2 / 3: the interior / exterior sites
// only (overlapped comms: COMM 2 runs while the faces travel and touches no halo, COMM 3 after they have arrived).
template <class T, int DAG, int AX, int COMM>
__global__ void __launch_bounds__(128) stag_dhop_kernel(const StagArgs<T> a) {
  const StagGeom &G = a.G;
  const int p = a.first_parity ^ (int)blockIdx.y;
  const uint32_t site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= G.V4cb) return;
  int c[4];
  stag_coor(G, p, site, c[0], c[1], c[2], c[3]);
  if (COMM >= 2 && stag_is_exterior(G, a.comm_dim_mask, c) != (COMM == 3)) return;
  const typename CT<T>::c *__restrict__ in = a.in[1 - p];
  const typename CT<T>::lv *__restrict__ rec = a.U[p] + (size_t)(site >> SLOG) * CT<T>::LVN * SW + (site & (SW - 1));
  CV<T> o;
#pragma unroll
  for (int k = 0; k < 3; k++) { o.re[k] = 0; o.im[k] = 0; }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int mu = q & 3, d = q < 4 ? 1 : 3;
    T uf[18], ub[18];
    pair_load(uf, ub, rec, q);
    CV<T> xf, xb;
    if (COMM) {   // neighbours beyond a decomposed boundary come from the received halos
      uint32_t idx;
      int w = stag_neighbour(G, a.comm_dim_mask, c, mu, d, idx);
      cv_load(xf, w < 0 ? in : a.halo[mu + 4 * w] + (size_t)(1 - p) * a.halo_parity_stride[mu], idx);
      w = stag_neighbour(G, a.comm_dim_mask, c, mu, -d, idx);
      cv_load(xb, w < 0 ? in : a.halo[mu + 4 * w] + (size_t)(1 - p) * a.halo_parity_stride[mu], idx);
    } else {
      int n[4] = {c[0], c[1], c[2], c[3]};
      const int L = G.L[mu];
      n[mu] = c[mu] + d; if (n[mu] >= L) n[mu] -= L;
      cv_load(xf, in, stag_cb(G, n[0], n[1], n[2], n[3]));
      n[mu] = c[mu] - d; if (n[mu] < 0) n[mu] += L;
      cv_load(xb, in, stag_cb(G, n[0], n[1], n[2], n[3]));
    }
    mv_add(o, uf, xf);
    mv_add(o, ub, xb);
  }
  typename CT<T>::c *__restrict__ out = a.out[p];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T re = DAG ? -o.re[k] : o.re[k], im = DAG ? -o.im[k] : o.im[k];
    if (AX) { const auto w = a.ax[p][cv_index(site, k)]; re = fma(a.axb, w.x, re); im = fma(a.axb, w.y, im); }
    out[cv_index(site, k)] = mkc(re, im);
  }
}

// Halo buffers: per (mu, dir) two input-parity slots of STAG_DEPTH * nface colour vectors (cv_index layout, padded to W).
static void stag_ensure_halo(gb_fermop *op) {
  if (op->halo_ready) return;
  const gb_grid *g = op->grid;
  StagGeom G = stag_geom(g);
  const size_t csz = op->prec == GB_F32 ? 8 : 16;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    const size_t blocks = ((size_t)STAG_DEPTH * stag_nface(G, mu) + W - 1) / W;
    op->halo_parity_stride[mu] = blocks * 3 * W;
    const size_t bytes = 2 * op->halo_parity_stride[mu] * csz;
    for (int dir = 0; dir < 2; dir++) {
      GB_CUDA(cudaMalloc(&op->halo_send[mu + 4 * dir], bytes));
      GB_CUDA(cudaMalloc(&op->halo_recv[mu + 4 * dir], bytes));
      GB_CUDA(cudaMemsetAsync(op->halo_send[mu + 4 * dir], 0, bytes, op->ctx->stream));
    }
  }
  op->halo_ready = true;
}
// pack the three boundary slices of every input parity of this hop and exchange them with the mu neighbours
// (replaces CartesianStencil::HaloExchange for the 16-point staggered stencil, ref: Grid/stencil/Stencil.h:367-430,
//  ImprovedStaggeredFermionImplementation.h:337-388 DhopInternalSerialComms)
template <class T>
static void stag_pack(gb_fermop *op, const void *const in[2], int first_parity, int nparity) {
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  StagGeom G = stag_geom(g);
  using C = typename CT<T>::c;
  // input parities: the opposite of each output parity
  const int ip0 = 1 - first_parity;
  for (int k = 0; k < nparity; k++) {
    const int ip = k == 0 ? ip0 : 1 - ip0;
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
      const uint32_t nface = stag_nface(G, mu);
      const unsigned blocks = (STAG_DEPTH * nface + 127) / 128;
      for (int dir = 0; dir < 2; dir++) {
        C *buf = (C *)op->halo_send[mu + 4 * dir] + (size_t)ip * op->halo_parity_stride[mu];
        stag_pack_kernel<T><<<blocks, 128, 0, ctx->stream>>>((const C *)in[ip], buf, G, mu, dir, ip, nface);
        count_launch(ctx);
      }
    }
  }
  check_launch(ctx, "stag_pack");
}
template <class T>
static void stag_exchange_packed(gb_fermop *op, int first_parity, int nparity, cudaStream_t st) {
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  using C = typename CT<T>::c;
  const int ip0 = 1 - first_parity;
  std::vector<HaloMsg> msgs;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    const size_t stride = op->halo_parity_stride[mu] * sizeof(C);
    const size_t off = nparity == 2 ? 0 : (size_t)ip0 * stride, bytes = nparity == 2 ? 2 * stride : stride;
    // halo dir 0 (the receiver's forward legs) is filled from MY first slices: it travels to my backward neighbour
    msgs.push_back({(const char *)op->halo_send[mu] + off, (char *)op->halo_recv[mu] + off, bytes, g->nbr_rank[mu][1], g->nbr_rank[mu][0]});
    msgs.push_back({(const char *)op->halo_send[mu + 4] + off, (char *)op->halo_recv[mu + 4] + off, bytes, g->nbr_rank[mu][0], g->nbr_rank[mu][1]});
  }
  stag_sendrecv(ctx, msgs, st);
}

// in[p]/out[p]: parity blocks.  Output parities first_parity (and the other one when nparity == 2).
template <class T>
static void stag_launch(gb_fermop *op, const void *const in[2], void *const out[2], int first_parity, int nparity, int dag, const void *const ax[2], double axb) {
  gb_context *ctx = op->ctx;
  StagArgs<T> a;
  for (int p = 0; p < 2; p++) {
    a.in[p] = (const typename CT<T>::c *)in[p]; a.out[p] = (typename CT<T>::c *)out[p];
    a.U[p] = (const typename CT<T>::lv *)((const char *)op->stag_links + (size_t)p * op->stag_parity_bytes);
    a.ax[p] = ax ? (const typename CT<T>::c *)ax[p] : nullptr;
  }
  a.axb = (T)axb; a.G = stag_geom(op->grid); a.first_parity = first_parity;
  a.comm_dim_mask = op->comm_dim_mask;
  for (int i = 0; i < 8; i++) a.halo[i] = nullptr;
  for (int mu = 0; mu < 4; mu++) a.halo_parity_stride[mu] = 0;
  dim3 grid((unsigned)((op->grid->V4cb + 127) / 128), nparity);
#define GB_SK(D, A, C) stag_dhop_kernel<T, D, A, C><<<grid, 128, 0, ctx->stream>>>(a)
#define GB_SKC(C) do { if (ax) { if (dag) GB_SK(1, 1, C); else GB_SK(0, 1, C); } else { if (dag) GB_SK(1, 0, C); else GB_SK(0, 0, C); } } while (0)
  if (op->comm_dim_mask) {
    stag_ensure_halo(op);
    for (int i = 0; i < 8; i++) a.halo[i] = (const typename CT<T>::c *)op->halo_recv[i];
    for (int mu = 0; mu < 4; mu++) a.halo_parity_stride[mu] = op->halo_parity_stride[mu];
    if (op->overlap_comms) {
      // pack on the compute stream, exchange on the comm stream while the interior sites are computed, then the exterior
      // sites (ref: ImprovedStaggeredFermion::DhopInternalOverlappedComms, ImprovedStaggeredFermionImplementation.h:283-335)
      stag_pack<T>(op, in, first_parity, nparity);
      GB_CUDA(cudaEventRecord(ctx->ev_comp, ctx->stream));
      GB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comp, 0));
      stag_exchange_packed<T>(op, first_parity, nparity, ctx->comm_stream);
      GB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
      GB_SKC(2);
      count_launch(ctx);
      GB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
      GB_SKC(3);
    } else {   // serial comms (ref: DhopInternalSerialComms, :337-388)
      stag_pack<T>(op, in, first_parity, nparity);
      stag_exchange_packed<T>(op, first_parity, nparity, ctx->stream);
      GB_SKC(1);
    }
  } else {
    if (ax) { if (dag) GB_SK(1, 1, 0); else GB_SK(0, 1, 0); }
    else { if (dag) GB_SK(1, 0, 0); else GB_SK(0, 0, 0); }
  }
#undef GB_SKC
#undef GB_SK
  count_launch(ctx);
  check_launch(ctx, "stag_dhop");
}
static void stag_hop(gb_fermop *op, const void *const in[2], void *const out[2], int first_parity, int nparity, int dag, const void *const ax[2] = nullptr, double axb = 0) {
  if (op->prec == GB_F32) stag_launch<float>(op, in, out, first_parity, nparity, dag, ax, axb);
  else stag_launch<double>(op, in, out, first_parity, nparity, dag, ax, axb);
}

static void stag_check(const gb_fermop *op, const gb_fermion *f, int kind, const char *what) {
  GB_REQUIRE(f != nullptr, "null field");
  if (!(f->grid == op->grid && f->ncomplex == 3 && f->Ls == 1 && f->prec == op->prec && f->kind == kind))
    throw Error(GB_ERR_INVALID, std::string(what) + ": field is not a conformable staggered (ColourVector) field of the operator's grid / precision");
}
static gb_fermion *stag_tmp(gb_fermop *op, int i, const gb_fermion *like) {
  if (!op->tmp_h[i]) op->tmp_h[i] = fermion_create_like(like, op->prec);
  return op->tmp_h[i];
}
static void stag_dhop_full(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dag, double mass_term = 0, bool with_mass = false) {
  const void *ib[2] = {in->block(0), in->block(1)};
  void *ob[2] = {out->block(0), out->block(1)};
  stag_hop(op, ib, ob, 0, 2, dag, with_mass ? ib : nullptr, mass_term);
}
static void stag_dhop_cb(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dag) {
  const int ip = in->cb, po = 1 - ip;
  const void *ib[2] = {nullptr, nullptr};
  void *ob[2] = {nullptr, nullptr};
  ib[ip] = in->data; ob[po] = out->data;
  stag_hop(op, ib, ob, po, 1, dag);
  out->cb = po;
}

void stag_op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag) {
  GB_REQUIRE(op->stag_links != nullptr, "operator has no gauge field: call ImportGauge first");
  auto chk = [](int rc) { if (rc != GB_OK) throw Error(rc, gb_last_error()); };
  switch (which) {
  case GB_OP_DHOP:
    stag_check(op, in, GB_FULL, "Dhop"); stag_check(op, out, GB_FULL, "Dhop");
    stag_dhop_full(op, in, out, dag);
    break;
  case GB_OP_DHOP_OE:
    stag_check(op, in, GB_HALF, "DhopOE"); stag_check(op, out, GB_HALF, "DhopOE");
    GB_REQUIRE(in->cb == GB_EVEN, "DhopOE needs an Even-checkerboard input");
    stag_dhop_cb(op, in, out, dag);
    break;
  case GB_OP_DHOP_EO:
    stag_check(op, in, GB_HALF, "DhopEO"); stag_check(op, out, GB_HALF, "DhopEO");
    GB_REQUIRE(in->cb == GB_ODD, "DhopEO needs an Odd-checkerboard input");
    stag_dhop_cb(op, in, out, dag);
    break;
  case GB_OP_M: case GB_OP_MDAG:   // M = Dhop + mass ; Mdag = -Dhop + mass: mass term fused into the hop's epilogue
    stag_check(op, in, GB_FULL, "M"); stag_check(op, out, GB_FULL, "M");
    stag_dhop_full(op, in, out, which == GB_OP_MDAG, op->mass, true);
    break;
  case GB_OP_MEOOE: case GB_OP_MEOOE_DAG:
    stag_check(op, in, GB_HALF, "Meooe"); stag_check(op, out, GB_HALF, "Meooe");
    stag_dhop_cb(op, in, out, which == GB_OP_MEOOE_DAG);
    break;
  case GB_OP_MOOEE: case GB_OP_MOOEE_DAG:
    stag_check(op, in, in->kind, "Mooee"); stag_check(op, out, in->kind, "Mooee");
    chk(gb_scale(out, op->mass, in));
    break;
  case GB_OP_MOOEE_INV: case GB_OP_MOOEE_INV_DAG:
    stag_check(op, in, in->kind, "MooeeInv"); stag_check(op, out, in->kind, "MooeeInv");
    chk(gb_scale(out, 1.0 / op->mass, in));
    break;
  case GB_OP_MPC: case GB_OP_MPC_DAG: case GB_OP_HERMOP: { // SchurStaggeredOperator: mass^2 - Meooe Meooe, Hermitian
    stag_check(op, in, GB_HALF, "Mpc"); stag_check(op, out, GB_HALF, "Mpc");
    gb_fermion *t = stag_tmp(op, 0, in);
    stag_dhop_cb(op, in, t, 0);
    // out = -Dhop t + mass^2 in : the dagger hop is minus the hop, the axpy rides in its epilogue
    const int ip = t->cb, po = 1 - ip;
    const void *ib[2] = {nullptr, nullptr}, *ab[2] = {nullptr, nullptr};
    void *ob[2] = {nullptr, nullptr};
    ib[ip] = t->data; ob[po] = out->data; ab[po] = in->data;
    stag_hop(op, ib, ob, po, 1, 1, ab, op->mass * op->mass);
    out->cb = in->cb;
    break;
  }
  case GB_OP_DMINUS: case GB_OP_DMINUS_DAG:   // ref: FermionOperator.h:172-173 (chi = psi)
    stag_check(op, in, in->kind, "Dminus"); stag_check(op, out, in->kind, "Dminus");
    chk(gb_copy(out, in));
    break;
  default:
    GB_REQUIRE(false, "opcode not defined for staggered operators");
  }
}

} // namespace gb

using namespace gb;

extern "C" {
// ref: ImprovedStaggeredFermion(Uthin,Ufat,Fgrid,Hgrid,mass,c1,c2,u0), ImprovedStaggeredFermion.h:115-121
int gb_op_create_staggered(gb_grid *g, const gb_gauge *Uthin, const gb_gauge *Ufat, double mass, double c1, double c2, double u0, gb_fermop **out) {
  GB_API_BEGIN
  GB_REQUIRE(g && Uthin && Ufat && out, "null argument");
  gb_fermop *op = new gb_fermop();
  op->grid = g; op->ctx = g->ctx; op->kind = GB_KIND_STAGGERED; op->prec = Uthin->prec; op->Ls = 1; op->mass = mass;
  // three-deep halos in the decomposed dimensions; GB_STAG_SELF_HALO=<bitmask> also routes undecomposed dimensions through
  // the pack / exchange-with-self / halo-lookup path (what the reference does for every dimension when it is told to treat
  // local wraps as comms; here it lets one GPU exercise the whole multi-rank code path)
  const int self_mask = getenv("GB_STAG_SELF_HALO") ? atoi(getenv("GB_STAG_SELF_HALO")) & 15 : 0;
  bool decomposed = false;
  for (int d = 0; d < 4; d++) { if (g->mpi[d] > 1) decomposed = true; if (g->mpi[d] > 1 || ((self_mask >> d) & 1)) op->comm_dim_mask |= 1 << d; }
  if (decomposed && g->ctx->nccl == nullptr) { delete op; GB_REQUIRE(false, "decomposed lattice: call gb_comm_init before creating operators"); }
  op->stag_c1 = c1; op->stag_c2 = c2; op->stag_u0 = u0;
  try { stag_import_gauge(op, Uthin, Ufat); } catch (...) { delete op; throw; }
  *out = op;
  GB_API_END
}
int gb_op_import_gauge_staggered(gb_fermop *op, const gb_gauge *Uthin, const gb_gauge *Ufat) {
  GB_API_BEGIN
  GB_REQUIRE(op && op->kind == GB_KIND_STAGGERED, "not a staggered operator");
  stag_import_gauge(op, Uthin, Ufat);
  GB_API_END
}
}
