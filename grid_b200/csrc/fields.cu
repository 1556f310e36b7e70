// fields.cu -- grids, lattice containers, import/export, checkerboards, precision change, BLAS-1, reductions.
// Semantics follow Grid/lattice (ref: Lattice_transfer.h:50-86,1123-1260,1461-1492 ; Lattice_arith.h:231-258 ;
// Lattice_reduction.h:256-372) but the container is re-designed: device-resident, blocked SoA (internal.hpp).
#include "internal.hpp"
#include "fermop.hpp"
#include "kernels_common.cuh"
#include <cstring>
#include <cmath>

using namespace gb;

// =====================================================================================================
// grids
// =====================================================================================================
// pure host geometry (no device needed): local extents, origin and neighbour ranks of `rank` in the processor grid
extern "C" int gb_geometry_query(const int gdims[4], const int mpi[4], int rank, int ldims[4], int origin[4], int nbr[8]) {
  GB_API_BEGIN
  GB_REQUIRE(gdims && mpi && ldims && origin && nbr, "null argument");
  int np = 1, pc[4], r = rank;
  for (int d = 0; d < 4; d++) {
    GB_REQUIRE(mpi[d] >= 1 && gdims[d] % mpi[d] == 0, "processor grid must divide the lattice");
    ldims[d] = gdims[d] / mpi[d];
    np *= mpi[d];
  }
  GB_REQUIRE(rank >= 0 && rank < np, "rank outside the processor grid");
  for (int d = 0; d < 4; d++) { pc[d] = r % mpi[d]; r /= mpi[d]; origin[d] = pc[d] * ldims[d]; }
  for (int d = 0; d < 4; d++)
    for (int dir = 0; dir < 2; dir++) {
      int q[4] = {pc[0], pc[1], pc[2], pc[3]};
      q[d] = (q[d] + (dir == 0 ? 1 : mpi[d] - 1)) % mpi[d];
      nbr[2 * d + dir] = q[0] + mpi[0] * (q[1] + mpi[1] * (q[2] + mpi[2] * q[3]));
    }
  GB_API_END
}

extern "C" int gb_grid_create(gb_context *ctx, const int gdims[4], const int mpi[4], gb_grid **out) {
  GB_API_BEGIN
  GB_REQUIRE(ctx && gdims && out, "null argument");
  gb_grid *g = new gb_grid();
  g->ctx = ctx;
  int np = 1;
  for (int d = 0; d < 4; d++) {
    g->gdims[d] = gdims[d];
    g->mpi[d] = mpi ? mpi[d] : 1;
    GB_REQUIRE(g->mpi[d] >= 1 && gdims[d] % g->mpi[d] == 0, "processor grid must divide the lattice");
    g->ldims[d] = gdims[d] / g->mpi[d];
    GB_REQUIRE(g->ldims[d] >= 2 && g->ldims[d] % 2 == 0, "local extents must be even and >= 2");
    np *= g->mpi[d];
  }
  GB_REQUIRE(np == ctx->nranks, "product of mpi[] must equal the number of ranks of the communicator");
  // rank -> processor coordinate, dimension 0 fastest (ref: Lexicographic::CoorFromIndex)
  int r = ctx->rank;
  for (int d = 0; d < 4; d++) { g->pcoor[d] = r % g->mpi[d]; r /= g->mpi[d]; g->origin[d] = g->pcoor[d] * g->ldims[d]; }
  for (int d = 0; d < 4; d++)
    for (int dir = 0; dir < 2; dir++) {
      int pc[4] = {g->pcoor[0], g->pcoor[1], g->pcoor[2], g->pcoor[3]};
      pc[d] = (pc[d] + (dir == 0 ? 1 : g->mpi[d] - 1)) % g->mpi[d];
      g->nbr_rank[d][dir] = pc[0] + g->mpi[0] * (pc[1] + g->mpi[1] * (pc[2] + g->mpi[2] * pc[3]));
    }
  g->V4 = (int64_t)g->ldims[0] * g->ldims[1] * g->ldims[2] * g->ldims[3];
  g->V4cb = g->V4 / 2;
  GB_REQUIRE(g->V4 < (1ll << 31), "local 4D volume must be < 2^31");
  *out = g;
  GB_API_END
}
extern "C" int gb_grid_destroy(gb_grid *g) { delete g; return GB_OK; }
extern "C" int gb_grid_local_dims(const gb_grid *g, int l[4]) { for (int d = 0; d < 4; d++) l[d] = g->ldims[d]; return GB_OK; }
extern "C" int gb_grid_local_origin(const gb_grid *g, int o[4]) { for (int d = 0; d < 4; d++) o[d] = g->origin[d]; return GB_OK; }

// =====================================================================================================
// fermion containers
// =====================================================================================================
static int fermion_create_impl(gb_grid *g, int Ls, int ncomplex, gb_precision prec, gb_gridkind kind, gb_fermion **out) {
  GB_API_BEGIN
  GB_REQUIRE(g && out && Ls >= 1, "bad argument");
  GB_REQUIRE(prec == GB_F32 || prec == GB_F64, "bad precision");
  gb_fermion *f = new gb_fermion();
  f->ctx = g->ctx; f->grid = g; f->Ls = Ls; f->prec = prec; f->kind = kind; f->cb = GB_EVEN; f->ncomplex = ncomplex;
  f->nsite4 = g->V4cb;
  f->n5cb = g->V4cb * Ls;
  GB_REQUIRE(f->n5cb * 2 < (1ll << 31), "local 5D volume must be < 2^31");
  f->hblk = (f->n5cb + W - 1) / W;
  f->nparity = kind == GB_HALF ? 1 : 2;
  f->bytes = (size_t)f->nvec() * 16;
  GB_CUDA(cudaSetDevice(g->ctx->device));
  f->data = nullptr;
  auto &pool = g->ctx->field_pool;
  for (size_t i = 0; i < pool.size(); i++)
    if (pool[i].first == f->bytes) { f->data = pool[i].second; g->ctx->field_pool_bytes -= f->bytes; pool.erase(pool.begin() + i); break; }
  if (!f->data) {
    cudaError_t e = cudaMalloc(&f->data, f->bytes);
    if (e != cudaSuccess && !pool.empty()) {              // out of memory with parked buffers: release them and retry
      cudaGetLastError();
      for (auto &b : pool) cudaFree(b.second);
      pool.clear(); g->ctx->field_pool_bytes = 0;
      e = cudaMalloc(&f->data, f->bytes);
    }
    if (e != cudaSuccess) { delete f; GB_CUDA(e); }
  }
  GB_CUDA(cudaMemsetAsync(f->data, 0, f->bytes, g->ctx->stream));
  *out = f;
  GB_API_END
}
extern "C" int gb_fermion_create(gb_grid *g, int Ls, gb_precision prec, gb_gridkind kind, gb_fermion **out) {
  return fermion_create_impl(g, Ls, 12, prec, kind, out);
}
extern "C" int gb_staggered_fermion_create(gb_grid *g, gb_precision prec, gb_gridkind kind, gb_fermion **out) {
  return fermion_create_impl(g, 1, 3, prec, kind, out);
}
extern "C" int gb_fermion_destroy(gb_fermion *f) {
  if (!f) return GB_OK;
  gb_context *ctx = context_alive(f->ctx) ? f->ctx : nullptr;   // a field may be destroyed after its context
  constexpr size_t POOL_MAX_BYTES = (size_t)48 << 30;     // a quarter of the 180 GB of HBM
  if (ctx && f->data && ctx->field_pool.size() < 64 && ctx->field_pool_bytes + f->bytes <= POOL_MAX_BYTES) {
    ctx->field_pool.emplace_back(f->bytes, f->data);
    ctx->field_pool_bytes += f->bytes;
  } else {
    cudaFree(f->data);
  }
  delete f;
  return GB_OK;
}
extern "C" int gb_fermion_checkerboard(const gb_fermion *f) { return f->cb; }
extern "C" int gb_fermion_set_checkerboard_tag(gb_fermion *f, int cb) { f->cb = cb & 1; return GB_OK; }
extern "C" int64_t gb_fermion_local_sites(const gb_fermion *f) { return f->n5cb * f->nparity; }

namespace gb {
void fermion_check_same(const gb_fermion *a, const gb_fermion *b) {
  GB_REQUIRE(a && b, "null field");
  GB_REQUIRE(a->grid == b->grid && a->Ls == b->Ls && a->kind == b->kind && a->prec == b->prec && a->ncomplex == b->ncomplex, "fields are not conformable");
}
// two INPUT fields of a binary operation on the red-black grid must live on the same checkerboard
// (ref: conformable(), Grid/lattice/Lattice_conformable.h: assert(lhs.Checkerboard() == rhs.Checkerboard()))
void fermion_check_same_cb(const gb_fermion *a, const gb_fermion *b) {
  if (a->kind == GB_HALF && a->cb != b->cb)
    throw Error(GB_ERR_INVALID, "fields live on different checkerboards (Even vs Odd): not conformable");
}
gb_fermion *fermion_create_like(const gb_fermion *like, int prec) {
  gb_fermion *f = nullptr;
  int rc = fermion_create_impl(like->grid, like->Ls, like->ncomplex, (gb_precision)prec, (gb_gridkind)like->kind, &f);
  if (rc != GB_OK) throw Error(rc, gb_last_error());
  f->cb = like->cb;
  return f;
}
} // namespace gb

// grow-only device staging buffer for host<->device layout changes (kept across calls: cudaMalloc of GBs is slow)
static void *ctx_staging(gb_context *ctx, size_t bytes) {
  if (ctx->staging_bytes < bytes) {
    if (ctx->staging) { GB_CUDA(cudaStreamSynchronize(ctx->stream)); GB_CUDA(cudaFree(ctx->staging)); ctx->staging = nullptr; ctx->staging_bytes = 0; }
    GB_CUDA(cudaMalloc(&ctx->staging, bytes));
    ctx->staging_bytes = bytes;
  }
  return ctx->staging;
}

struct LatGeom {
  int L[4];      // local dims
  int Lxh;       // L[0]/2
  int Ls;
  int origin_parity;
  int64_t n5cb, hblk;
};
static LatGeom geom_of(const gb_fermion *f) {
  LatGeom G;
  for (int d = 0; d < 4; d++) G.L[d] = f->grid->ldims[d];
  G.Lxh = G.L[0] / 2; G.Ls = f->Ls;
  G.origin_parity = (f->grid->origin[0] + f->grid->origin[1] + f->grid->origin[2] + f->grid->origin[3]) & 1;
  G.n5cb = f->n5cb; G.hblk = f->hblk;
  return G;
}
// (parity block p, cb site index) -> local lexicographic 4D index.  ref: Cartesian_red_black.h:271-286
__device__ __forceinline__ int64_t cb_to_lex(const LatGeom &G, int p, int64_t site) {
  int xh = site % G.Lxh; site /= G.Lxh;
  int y = site % G.L[1]; site /= G.L[1];
  int z = site % G.L[2];
  int t = site / G.L[2];
  int x = 2 * xh + ((p + G.origin_parity + y + z + t) & 1);
  return x + (int64_t)G.L[0] * (y + (int64_t)G.L[1] * (z + (int64_t)G.L[2] * t));
}

// One thread per device vec element. DIR=0: host->device, DIR=1: device->host
template <class TD, class TH, int DIR>
__global__ void fermion_transfer_kernel(typename Prec<TD>::vec *dev, TH *host, LatGeom G, int nparity, int full, int cb_half) {
  using P = Prec<TD>;
  const int64_t nelem = (int64_t)nparity * G.hblk * P::NV * W;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  int k = r % P::NV;
  int64_t blk = r / P::NV;
  int p = blk / G.hblk;
  int64_t i5cb = (blk - (int64_t)p * G.hblk) * W + lane;
  if (i5cb >= G.n5cb) {
    if (DIR == 0) {
      if constexpr (sizeof(TD) == 4) dev[e] = make_float4(0, 0, 0, 0); else dev[e] = make_double2(0, 0);
    }
    return;
  }
  int64_t site = i5cb / G.Ls;
  int s = i5cb - site * G.Ls;
  int64_t hidx; // host 5D site index
  if (full) hidx = s + (int64_t)G.Ls * cb_to_lex(G, p, site);
  else hidx = i5cb;
  TH *h = host + hidx * 24;
  if constexpr (sizeof(TD) == 4) {
    if (DIR == 0) dev[e] = make_float4((float)h[4 * k], (float)h[4 * k + 1], (float)h[4 * k + 2], (float)h[4 * k + 3]);
    else { float4 v = dev[e]; h[4 * k] = (TH)v.x; h[4 * k + 1] = (TH)v.y; h[4 * k + 2] = (TH)v.z; h[4 * k + 3] = (TH)v.w; }
  } else {
    if (DIR == 0) dev[e] = make_double2((double)h[2 * k], (double)h[2 * k + 1]);
    else { double2 v = dev[e]; h[2 * k] = (TH)v.x; h[2 * k + 1] = (TH)v.y; }
  }
}

template <int DIR> static void fermion_transfer(const gb_fermion *f, void *host, int host_prec) {
  gb_context *ctx = f->grid->ctx;
  GB_CUDA(cudaSetDevice(ctx->device));
  const int64_t nsites = f->n5cb * f->nparity;
  const size_t hbytes = (size_t)nsites * 2 * f->ncomplex * (host_prec == GB_F32 ? 4 : 8);
  void *stage = ctx_staging(ctx, hbytes);
  if (DIR == 0) GB_CUDA(cudaMemcpyAsync(stage, host, hbytes, cudaMemcpyHostToDevice, ctx->stream));
  if (f->ncomplex == 3) {
    stag_transfer(f, stage, host_prec, DIR);
    if (DIR == 1) GB_CUDA(cudaMemcpyAsync(host, stage, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  LatGeom G = geom_of(f);
  const int64_t nelem = f->nvec();
  const int threads = 256;
  const unsigned blocks = (unsigned)((nelem + threads - 1) / threads);
  const int full = f->kind == GB_FULL;
#define LAUNCH(TD, TH)                                                                                         \
  fermion_transfer_kernel<TD, TH, DIR><<<blocks, threads, 0, ctx->stream>>>((typename Prec<TD>::vec *)f->data, (TH *)stage, G, f->nparity, full, f->cb)
  if (f->prec == GB_F32 && host_prec == GB_F32) LAUNCH(float, float);
  else if (f->prec == GB_F32) LAUNCH(float, double);
  else if (host_prec == GB_F32) LAUNCH(double, float);
  else LAUNCH(double, double);
#undef LAUNCH
  count_launch(ctx);
  check_launch(ctx, "fermion_transfer");
  if (DIR == 1) GB_CUDA(cudaMemcpyAsync(host, stage, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
}

extern "C" int gb_fermion_import(gb_fermion *f, const void *host, gb_precision host_prec) {
  GB_API_BEGIN
  GB_REQUIRE(f && host, "null argument");
  fermion_transfer<0>(f, const_cast<void *>(host), host_prec);
  GB_API_END
}
extern "C" int gb_fermion_export(const gb_fermion *f, void *host, gb_precision host_prec) {
  GB_API_BEGIN
  GB_REQUIRE(f && host, "null argument");
  fermion_transfer<1>(f, host, host_prec);
  GB_API_END
}

// pick/set checkerboard: a full field is stored as [even block][odd block], so these are block copies.
extern "C" int gb_pick_checkerboard(int cb, gb_fermion *half, const gb_fermion *full) {
  GB_API_BEGIN
  GB_REQUIRE(half && full && half->kind == GB_HALF && full->kind == GB_FULL, "pickCheckerboard(cb, half, full)");
  GB_REQUIRE(half->grid == full->grid && half->Ls == full->Ls && half->prec == full->prec && half->ncomplex == full->ncomplex, "fields are not conformable");
  gb_context *ctx = full->grid->ctx;
  GB_CUDA(cudaMemcpyAsync(half->data, full->block(cb & 1), half->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  half->cb = cb & 1;
  GB_API_END
}
extern "C" int gb_set_checkerboard(gb_fermion *full, const gb_fermion *half) {
  GB_API_BEGIN
  GB_REQUIRE(half && full && half->kind == GB_HALF && full->kind == GB_FULL, "setCheckerboard(full, half)");
  GB_REQUIRE(half->grid == full->grid && half->Ls == full->Ls && half->prec == full->prec && half->ncomplex == full->ncomplex, "fields are not conformable");
  gb_context *ctx = full->grid->ctx;
  GB_CUDA(cudaMemcpyAsync(full->block(half->cb), half->data, half->bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  GB_API_END
}

// ------------------------------------------------------------------ precision change
__global__ void prec_d2f_kernel(float4 *out, const double2 *in, int64_t nblk) {
  // one thread per output float4: (blk, k6, lane) <- two double2 (blk, 2*k6 [+1], lane)
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nblk * 6 * W) return;
  int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  int k = r % 6; int64_t blk = r / 6;
  double2 a = in[((blk * 12 + 2 * k) << LOGW) + lane], b = in[((blk * 12 + 2 * k + 1) << LOGW) + lane];
  out[e] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
}
__global__ void prec_f2d_kernel(double2 *out, const float4 *in, int64_t nblk) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nblk * 6 * W) return;
  int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  int k = r % 6; int64_t blk = r / 6;
  float4 v = in[e];
  out[((blk * 12 + 2 * k) << LOGW) + lane] = make_double2(v.x, v.y);
  out[((blk * 12 + 2 * k + 1) << LOGW) + lane] = make_double2(v.z, v.w);
}
extern "C" int gb_precision_change(gb_fermion *out, const gb_fermion *in) {
  GB_API_BEGIN
  GB_REQUIRE(out && in && out->grid == in->grid && out->Ls == in->Ls && out->kind == in->kind && out->ncomplex == in->ncomplex, "fields are not conformable");
  gb_context *ctx = in->grid->ctx;
  out->cb = in->cb;
  if (out->prec == in->prec) { GB_CUDA(cudaMemcpyAsync(out->data, in->data, in->bytes, cudaMemcpyDeviceToDevice, ctx->stream)); return GB_OK; }
  if (in->ncomplex == 3) { stag_precision_change(out, in); return GB_OK; }
  const int64_t nblk = in->hblk * in->nparity;
  const int64_t n = nblk * 6 * W;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (out->prec == GB_F32) prec_d2f_kernel<<<blocks, 256, 0, ctx->stream>>>((float4 *)out->data, (const double2 *)in->data, nblk);
  else prec_f2d_kernel<<<blocks, 256, 0, ctx->stream>>>((double2 *)out->data, (const float4 *)in->data, nblk);
  count_launch(ctx);
  check_launch(ctx, "precision_change");
  GB_API_END
}

// ------------------------------------------------------------------ synthetic source
template <class TD>
__global__ void fermion_random_kernel(typename Prec<TD>::vec *dev, LatGeom G, int nparity, int cb_half, int4 gorigin, int4 gdims, uint64_t seed) {
  using P = Prec<TD>;
  const int64_t nelem = (int64_t)nparity * G.hblk * P::NV * W;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  int k = r % P::NV;
  int64_t blk = r / P::NV;
  int p = blk / G.hblk;
  int64_t i5cb = (blk - (int64_t)p * G.hblk) * W + lane;
  if (i5cb >= G.n5cb) { if constexpr (sizeof(TD) == 4) dev[e] = make_float4(0, 0, 0, 0); else dev[e] = make_double2(0, 0); return; }
  int64_t site = i5cb / G.Ls;
  int s = i5cb - site * G.Ls;
  int par = nparity == 2 ? p : cb_half;
  int64_t lex = cb_to_lex(G, par, site);
  int x = lex % G.L[0]; lex /= G.L[0]; int y = lex % G.L[1]; lex /= G.L[1]; int z = lex % G.L[2]; int t = lex / G.L[2];
  // global 5D lexicographic index keys the stream => decomposition independent
  uint64_t g4 = (uint64_t)(x + gorigin.x) + (uint64_t)gdims.x * ((y + gorigin.y) + (uint64_t)gdims.y * ((z + gorigin.z) + (uint64_t)gdims.z * (t + gorigin.w)));
  uint64_t g5 = (uint64_t)s + (uint64_t)G.Ls * g4;
  uint64_t key = splitmix64(seed);
  if constexpr (sizeof(TD) == 4) {
    float v[4];
    for (int j = 0; j < 4; j++) v[j] = (float)u01(splitmix64(key ^ (g5 * 24 + 4 * k + j)));
    dev[e] = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    dev[e] = make_double2(u01(splitmix64(key ^ (g5 * 24 + 2 * k))), u01(splitmix64(key ^ (g5 * 24 + 2 * k + 1))));
  }
}
extern "C" int gb_fermion_random(gb_fermion *f, uint64_t seed) {
  GB_API_BEGIN
  if (f->ncomplex == 3) { stag_random(f, seed); return GB_OK; }
  gb_context *ctx = f->grid->ctx;
  LatGeom G = geom_of(f);
  const gb_grid *g = f->grid;
  int4 go = make_int4(g->origin[0], g->origin[1], g->origin[2], g->origin[3]);
  int4 gd = make_int4(g->gdims[0], g->gdims[1], g->gdims[2], g->gdims[3]);
  const int64_t nelem = f->nvec();
  const unsigned blocks = (unsigned)((nelem + 255) / 256);
  if (f->prec == GB_F32) fermion_random_kernel<float><<<blocks, 256, 0, ctx->stream>>>((float4 *)f->data, G, f->nparity, f->cb, go, gd, seed);
  else fermion_random_kernel<double><<<blocks, 256, 0, ctx->stream>>>((double2 *)f->data, G, f->nparity, f->cb, go, gd, seed);
  count_launch(ctx);
  check_launch(ctx, "fermion_random");
  GB_API_END
}

// =====================================================================================================
// BLAS-1 (elementwise over the blocked storage; padding lanes stay zero)
// =====================================================================================================
enum { BL_ZERO, BL_COPY, BL_SCALE, BL_AXPY, BL_AXPBY };
template <class V, class T, int OP> __global__ void blas_kernel(V *z, const V *x, const V *y, T a, T b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    V r;
    if (OP == BL_SCALE) r = vscale(a, x[i]);
    else if (OP == BL_AXPY) r = vaxpy(a, x[i], y[i]);
    else r = vaxpby(a, x[i], b, y[i]);
    z[i] = r;
  }
}
template <int OP> static void blas_launch(gb_fermion *z, double a, double b, const gb_fermion *x, const gb_fermion *y) {
  gb_context *ctx = z->grid->ctx;
  const int64_t n = z->nvec();
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (z->prec == GB_F32)
    blas_kernel<float4, float, OP><<<blocks, 256, 0, ctx->stream>>>((float4 *)z->data, (const float4 *)x->data, y ? (const float4 *)y->data : nullptr, (float)a, (float)b, n);
  else
    blas_kernel<double2, double, OP><<<blocks, 256, 0, ctx->stream>>>((double2 *)z->data, (const double2 *)x->data, y ? (const double2 *)y->data : nullptr, a, b, n);
  count_launch(ctx);
  check_launch(ctx, "blas");
}
extern "C" int gb_zero(gb_fermion *z) {
  GB_API_BEGIN
  GB_CUDA(cudaMemsetAsync(z->data, 0, z->bytes, z->grid->ctx->stream));
  GB_API_END
}
extern "C" int gb_copy(gb_fermion *z, const gb_fermion *x) {
  GB_API_BEGIN
  fermion_check_same(z, x);
  GB_CUDA(cudaMemcpyAsync(z->data, x->data, x->bytes, cudaMemcpyDeviceToDevice, z->grid->ctx->stream));
  z->cb = x->cb;
  GB_API_END
}
extern "C" int gb_scale(gb_fermion *z, double a, const gb_fermion *x) {
  GB_API_BEGIN
  fermion_check_same(z, x);
  blas_launch<BL_SCALE>(z, a, 0, x, nullptr);
  z->cb = x->cb;
  GB_API_END
}
extern "C" int gb_axpy(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y); fermion_check_same_cb(x, y);
  blas_launch<BL_AXPY>(z, a, 0, x, y);
  z->cb = x->cb;
  GB_API_END
}
extern "C" int gb_axpby(gb_fermion *z, double a, double b, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y); fermion_check_same_cb(x, y);
  blas_launch<BL_AXPBY>(z, a, b, x, y);
  z->cb = x->cb;
  GB_API_END
}

// =====================================================================================================
// reductions: per-thread double accumulation, fixed-shape block tree, fixed-order second stage
// ref semantics: Lattice_reduction.h:256-311 (innerProduct), :321-372 (axpy_norm) + GlobalSum
// =====================================================================================================
constexpr int RED_THREADS = 256;
template <int NOUT> __device__ __forceinline__ void block_reduce_store(double (&acc)[NOUT], double *partials) {
  __shared__ double sm[NOUT][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NOUT; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NOUT; j++) {
      double v = 0;
      for (int w = 0; w < RED_THREADS / 32; w++) v += sm[j][w];
      partials[blockIdx.x * NOUT + j] = v;
    }
  }
}
template <int NOUT> __global__ void reduce_final_kernel(const double *partials, int nblocks, double *result) {
  // single block; strided per-thread sums then a fixed tree: deterministic for a given nblocks
  double acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; j++) acc[j] = 0;
  for (int i = threadIdx.x; i < nblocks; i += RED_THREADS)
#pragma unroll
    for (int j = 0; j < NOUT; j++) acc[j] += partials[i * NOUT + j];
  __shared__ double sm[NOUT][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NOUT; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int j = 0; j < NOUT; j++) {
      double v = 0;
      for (int w = 0; w < RED_THREADS / 32; w++) v += sm[j][w];
      result[j] = v;
    }
}

template <class V> __global__ void norm2_kernel(const V *x, int64_t n, double *partials) {
  double acc[1] = {0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) acc[0] += vnorm2(x[i]);
  block_reduce_store<1>(acc, partials);
}
template <class V> __global__ void inner_kernel(const V *l, const V *r, int64_t n, double *partials) {
  double acc[2] = {0, 0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) vinner(l[i], r[i], acc[0], acc[1]);
  block_reduce_store<2>(acc, partials);
}
template <class V, class T> __global__ void axpy_norm_kernel(V *z, const V *x, const V *y, T a, int64_t n, double *partials) {
  double acc[1] = {0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    V r = vaxpy(a, x[i], y[i]);
    z[i] = r;
    acc[0] += vnorm2(r);
  }
  block_reduce_store<1>(acc, partials);
}

static unsigned red_blocks(gb_context *ctx, int64_t n) {
  return (unsigned)std::min<int64_t>(std::min<int64_t>((n + RED_THREADS - 1) / RED_THREADS, (int64_t)ctx->sm_count * 8), ctx->max_partials);
}
template <int NOUT> static void red_finish(gb_context *ctx, unsigned blocks, double *out) {
  reduce_final_kernel<NOUT><<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partials, (int)blocks, ctx->d_result);
  count_launch(ctx);
  check_launch(ctx, "reduce_final");
  GB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, NOUT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < NOUT; j++) out[j] = ctx->h_result[j];
  global_sum(ctx, out, NOUT);
}
// ---- device-resident variants used by the fused CG: results stay on the device (summed over ranks in-stream),
//      scalars a = c/d and b = cp/c are formed inside the consuming kernels, so the host never sits between launches
template <class V, class T> __global__ void axpy_norm_dev_kernel(V *z, const V *x, const V *y, const double *c, const double *d, int64_t n, double *partials) {
  const T a = (T)(-(*c) / (*d)); // r -= (c/d) q
  double acc[1] = {0};
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    V r = vaxpy(a, x[i], y[i]);
    z[i] = r;
    acc[0] += vnorm2(r);
  }
  block_reduce_store<1>(acc, partials);
}
template <class V, class T> __global__ void cg_update_dev_kernel(V *psi, V *p, const V *r, const double *c, const double *d, const double *cp, int64_t n) {
  const T a = (T)((*c) / (*d)), b = (T)((*cp) / (*c)); // psi += a p ; p = b p + r   (ref: ConjugateGradient.h:176-183)
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const V pv = p[i];
    psi[i] = vaxpy(a, pv, psi[i]);
    p[i] = vaxpy(b, pv, r[i]);
  }
}
template <int NOUT> __global__ void reduce_final_to_kernel(const double *partials, int nblocks, double *result) {
  double acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; j++) acc[j] = 0;
  for (int i = threadIdx.x; i < nblocks; i += RED_THREADS)
#pragma unroll
    for (int j = 0; j < NOUT; j++) acc[j] += partials[i * NOUT + j];
  __shared__ double sm[NOUT][RED_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NOUT; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int j = 0; j < NOUT; j++) {
      double v = 0;
      for (int w = 0; w < RED_THREADS / 32; w++) v += sm[j][w];
      result[j] = v;
    }
}
namespace gb {
void device_global_sum(gb_context *ctx, double *d_vals, int n); // context.cu
// d_out[0..1] = sum conj(l) r over all ranks, left on the device
void reduce_inner_dev(gb_context *ctx, const gb_fermion *l, const gb_fermion *r, double *d_out) {
  const int64_t n = l->nvec();
  const unsigned blocks = red_blocks(ctx, n);
  if (l->prec == GB_F32) inner_kernel<float4><<<blocks, RED_THREADS, 0, ctx->stream>>>((const float4 *)l->data, (const float4 *)r->data, n, ctx->d_partials);
  else inner_kernel<double2><<<blocks, RED_THREADS, 0, ctx->stream>>>((const double2 *)l->data, (const double2 *)r->data, n, ctx->d_partials);
  reduce_final_to_kernel<2><<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partials, (int)blocks, d_out);
  count_launch(ctx, 2);
  check_launch(ctx, "inner_dev");
  device_global_sum(ctx, d_out, 2);
}
// z = y - (c/d) x ; d_out[0] = |z|^2 over all ranks (device)
void axpy_norm_dev(gb_context *ctx, gb_fermion *z, const gb_fermion *x, const gb_fermion *y, const double *d_c, const double *d_d, double *d_out) {
  const int64_t n = z->nvec();
  const unsigned blocks = red_blocks(ctx, n);
  if (z->prec == GB_F32)
    axpy_norm_dev_kernel<float4, float><<<blocks, RED_THREADS, 0, ctx->stream>>>((float4 *)z->data, (const float4 *)x->data, (const float4 *)y->data, d_c, d_d, n, ctx->d_partials);
  else
    axpy_norm_dev_kernel<double2, double><<<blocks, RED_THREADS, 0, ctx->stream>>>((double2 *)z->data, (const double2 *)x->data, (const double2 *)y->data, d_c, d_d, n, ctx->d_partials);
  reduce_final_to_kernel<1><<<1, RED_THREADS, 0, ctx->stream>>>(ctx->d_partials, (int)blocks, d_out);
  count_launch(ctx, 2);
  check_launch(ctx, "axpy_norm_dev");
  device_global_sum(ctx, d_out, 1);
}
void cg_update_dev(gb_context *ctx, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, const double *d_c, const double *d_d, const double *d_cp) {
  const int64_t n = psi->nvec();
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (psi->prec == GB_F32) cg_update_dev_kernel<float4, float><<<blocks, 256, 0, ctx->stream>>>((float4 *)psi->data, (float4 *)p->data, (const float4 *)r->data, d_c, d_d, d_cp, n);
  else cg_update_dev_kernel<double2, double><<<blocks, 256, 0, ctx->stream>>>((double2 *)psi->data, (double2 *)p->data, (const double2 *)r->data, d_c, d_d, d_cp, n);
  count_launch(ctx);
  check_launch(ctx, "cg_update_dev");
}

void reduce_norm2(gb_context *ctx, const gb_fermion *x, double *out) {
  const int64_t n = x->nvec();
  const unsigned blocks = red_blocks(ctx, n);
  if (x->prec == GB_F32) norm2_kernel<float4><<<blocks, RED_THREADS, 0, ctx->stream>>>((const float4 *)x->data, n, ctx->d_partials);
  else norm2_kernel<double2><<<blocks, RED_THREADS, 0, ctx->stream>>>((const double2 *)x->data, n, ctx->d_partials);
  count_launch(ctx);
  check_launch(ctx, "norm2");
  red_finish<1>(ctx, blocks, out);
}
void reduce_inner(gb_context *ctx, const gb_fermion *l, const gb_fermion *r, double out[2]) {
  const int64_t n = l->nvec();
  const unsigned blocks = red_blocks(ctx, n);
  if (l->prec == GB_F32) inner_kernel<float4><<<blocks, RED_THREADS, 0, ctx->stream>>>((const float4 *)l->data, (const float4 *)r->data, n, ctx->d_partials);
  else inner_kernel<double2><<<blocks, RED_THREADS, 0, ctx->stream>>>((const double2 *)l->data, (const double2 *)r->data, n, ctx->d_partials);
  count_launch(ctx);
  check_launch(ctx, "inner");
  red_finish<2>(ctx, blocks, out);
}
} // namespace gb

extern "C" int gb_norm2(const gb_fermion *x, double *out) {
  GB_API_BEGIN
  GB_REQUIRE(x && out, "null argument");
  reduce_norm2(x->grid->ctx, x, out);
  GB_API_END
}
extern "C" int gb_inner_product(const gb_fermion *l, const gb_fermion *r, double out[2]) {
  GB_API_BEGIN
  fermion_check_same(l, r); fermion_check_same_cb(l, r);
  reduce_inner(l->grid->ctx, l, r, out);
  GB_API_END
}
extern "C" int gb_axpy_norm(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y, double *norm2_z) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y); fermion_check_same_cb(x, y);
  gb_context *ctx = z->grid->ctx;
  const int64_t n = z->nvec();
  const unsigned blocks = red_blocks(ctx, n);
  if (z->prec == GB_F32)
    axpy_norm_kernel<float4, float><<<blocks, RED_THREADS, 0, ctx->stream>>>((float4 *)z->data, (const float4 *)x->data, (const float4 *)y->data, (float)a, n, ctx->d_partials);
  else
    axpy_norm_kernel<double2, double><<<blocks, RED_THREADS, 0, ctx->stream>>>((double2 *)z->data, (const double2 *)x->data, (const double2 *)y->data, a, n, ctx->d_partials);
  count_launch(ctx);
  check_launch(ctx, "axpy_norm");
  z->cb = x->cb;
  red_finish<1>(ctx, blocks, norm2_z);
  GB_API_END
}

// =====================================================================================================
// gauge containers (lexicographic AoS on the device; only touched at import / DoubleStore time)
// =====================================================================================================
extern "C" int gb_gauge_create(gb_grid *g, gb_precision prec, gb_gauge **out) {
  GB_API_BEGIN
  GB_REQUIRE(g && out, "null argument");
  gb_gauge *u = new gb_gauge();
  u->grid = g; u->prec = prec;
  u->bytes = (size_t)g->V4 * 72 * (prec == GB_F32 ? 4 : 8);
  GB_CUDA(cudaSetDevice(g->ctx->device));
  GB_CUDA(cudaMalloc(&u->data, u->bytes));
  GB_CUDA(cudaMemsetAsync(u->data, 0, u->bytes, g->ctx->stream));
  *out = u;
  GB_API_END
}
extern "C" int gb_gauge_destroy(gb_gauge *u) {
  if (u) { cudaFree(u->data); delete u; }
  return GB_OK;
}
template <class TO, class TI> __global__ void convert_kernel(TO *o, const TI *i, int64_t n) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n) o[e] = (TO)i[e];
}
static void gauge_transfer(const gb_gauge *u, void *host, int host_prec, bool to_device) {
  gb_context *ctx = u->grid->ctx;
  GB_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = u->grid->V4 * 72;
  if (host_prec == u->prec) {
    if (to_device) GB_CUDA(cudaMemcpyAsync(u->data, host, u->bytes, cudaMemcpyHostToDevice, ctx->stream));
    else GB_CUDA(cudaMemcpyAsync(host, u->data, u->bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  const size_t hbytes = (size_t)n * (host_prec == GB_F32 ? 4 : 8);
  void *stage;
  GB_CUDA(cudaMalloc(&stage, hbytes));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (to_device) {
    GB_CUDA(cudaMemcpyAsync(stage, host, hbytes, cudaMemcpyHostToDevice, ctx->stream));
    if (u->prec == GB_F32) convert_kernel<float, double><<<blocks, 256, 0, ctx->stream>>>((float *)u->data, (const double *)stage, n);
    else convert_kernel<double, float><<<blocks, 256, 0, ctx->stream>>>((double *)u->data, (const float *)stage, n);
  } else {
    if (u->prec == GB_F32) convert_kernel<double, float><<<blocks, 256, 0, ctx->stream>>>((double *)stage, (const float *)u->data, n);
    else convert_kernel<float, double><<<blocks, 256, 0, ctx->stream>>>((float *)stage, (const double *)u->data, n);
    GB_CUDA(cudaMemcpyAsync(host, stage, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  count_launch(ctx);
  check_launch(ctx, "gauge convert");
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  GB_CUDA(cudaFree(stage));
}
extern "C" int gb_gauge_import(gb_gauge *u, const void *host, gb_precision host_prec) {
  GB_API_BEGIN
  GB_REQUIRE(u && host, "null argument");
  gauge_transfer(u, const_cast<void *>(host), host_prec, true);
  GB_API_END
}
extern "C" int gb_gauge_export(const gb_gauge *u, void *host, gb_precision host_prec) {
  GB_API_BEGIN
  GB_REQUIRE(u && host, "null argument");
  gauge_transfer(u, host, host_prec, false);
  GB_API_END
}

// random SU(3): gaussian -> Ta -> exp (scaling & squaring Taylor) -> reunitarise. ref: GaugeGroup.h:332-349
struct cd { double x, y; };
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ void mat_mul(cd (&c)[3][3], const cd (&a)[3][3], const cd (&b)[3][3]) {
  cd t[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    cd s = {0, 0};
    for (int k = 0; k < 3; k++) { cd p = cmul(a[i][k], b[k][j]); s.x += p.x; s.y += p.y; }
    t[i][j] = s;
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c[i][j] = t[i][j];
}
template <class T>
__global__ void gauge_random_kernel(T *U, int4 L, int4 gorigin, int4 gdims, uint64_t seed, int unit) {
  const int64_t V4 = (int64_t)L.x * L.y * L.z * L.w;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= V4 * 4) return;
  int mu = e & 3;
  int64_t lex = e >> 2;
  int x = lex % L.x; int64_t r = lex / L.x; int y = r % L.y; r /= L.y; int z = r % L.z; int t = r / L.z;
  uint64_t g4 = (uint64_t)(x + gorigin.x) + (uint64_t)gdims.x * ((y + gorigin.y) + (uint64_t)gdims.y * ((z + gorigin.z) + (uint64_t)gdims.z * (t + gorigin.w)));
  cd u[3][3];
  if (unit) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) u[i][j] = {i == j ? 1.0 : 0.0, 0.0};
  } else {
    uint64_t key = splitmix64(seed ^ 0xA5A5A5A5DEADBEEFull);
    cd g[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      uint64_t c = (g4 * 4 + mu) * 9 + i * 3 + j;
      double u1 = u01(splitmix64(key ^ (2 * c))), u2 = u01(splitmix64(key ^ (2 * c + 1)));
      double rad = sqrt(-2.0 * log(1.0 - u1)); // 1-u1 in (0,1]
      g[i][j] = {rad * cos(2.0 * M_PI * u2), rad * sin(2.0 * M_PI * u2)};
    }
    // Ta: anti-hermitian traceless part
    cd a[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = {0.5 * (g[i][j].x - g[j][i].x), 0.5 * (g[i][j].y + g[j][i].y)};
    double tr = (a[0][0].y + a[1][1].y + a[2][2].y) / 3.0;
    for (int i = 0; i < 3; i++) { a[i][i].x = 0; a[i][i].y -= tr; }
    // exp(a) = (exp(a/2^6))^(2^6), Taylor to order 12
    const double sc = 1.0 / 64.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j].x *= sc; a[i][j].y *= sc; }
    cd term[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { u[i][j] = {i == j ? 1.0 : 0.0, 0.0}; term[i][j] = u[i][j]; }
    for (int n = 1; n <= 12; n++) {
      mat_mul(term, term, a);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { term[i][j].x /= n; term[i][j].y /= n; u[i][j].x += term[i][j].x; u[i][j].y += term[i][j].y; }
    }
    for (int q = 0; q < 6; q++) mat_mul(u, u, u);
    // reunitarise: Gram-Schmidt rows 0,1 ; row 2 = conj(row0 x row1)
    double n0 = 0; for (int j = 0; j < 3; j++) n0 += u[0][j].x * u[0][j].x + u[0][j].y * u[0][j].y;
    n0 = 1.0 / sqrt(n0); for (int j = 0; j < 3; j++) { u[0][j].x *= n0; u[0][j].y *= n0; }
    cd dot = {0, 0}; // <row0,row1> = sum conj(u0) u1
    for (int j = 0; j < 3; j++) { dot.x += u[0][j].x * u[1][j].x + u[0][j].y * u[1][j].y; dot.y += u[0][j].x * u[1][j].y - u[0][j].y * u[1][j].x; }
    for (int j = 0; j < 3; j++) { cd p = cmul(dot, u[0][j]); u[1][j].x -= p.x; u[1][j].y -= p.y; }
    double n1 = 0; for (int j = 0; j < 3; j++) n1 += u[1][j].x * u[1][j].x + u[1][j].y * u[1][j].y;
    n1 = 1.0 / sqrt(n1); for (int j = 0; j < 3; j++) { u[1][j].x *= n1; u[1][j].y *= n1; }
    for (int j = 0; j < 3; j++) {
      int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      cd p = cmul(u[0][j1], u[1][j2]), q = cmul(u[0][j2], u[1][j1]);
      u[2][j] = {p.x - q.x, -(p.y - q.y)};
    }
  }
  T *o = U + e * 18;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { o[(i * 3 + j) * 2] = (T)u[i][j].x; o[(i * 3 + j) * 2 + 1] = (T)u[i][j].y; }
}
static void gauge_fill(gb_gauge *u, uint64_t seed, int unit) {
  gb_context *ctx = u->grid->ctx;
  const gb_grid *g = u->grid;
  int4 L = make_int4(g->ldims[0], g->ldims[1], g->ldims[2], g->ldims[3]);
  int4 go = make_int4(g->origin[0], g->origin[1], g->origin[2], g->origin[3]);
  int4 gd = make_int4(g->gdims[0], g->gdims[1], g->gdims[2], g->gdims[3]);
  const int64_t n = g->V4 * 4;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (u->prec == GB_F32) gauge_random_kernel<float><<<blocks, 128, 0, ctx->stream>>>((float *)u->data, L, go, gd, seed, unit);
  else gauge_random_kernel<double><<<blocks, 128, 0, ctx->stream>>>((double *)u->data, L, go, gd, seed, unit);
  count_launch(ctx);
  check_launch(ctx, "gauge_random");
}
extern "C" int gb_gauge_random(gb_gauge *u, uint64_t seed) {
  GB_API_BEGIN
  gauge_fill(u, seed, 0);
  GB_API_END
}
extern "C" int gb_gauge_unit(gb_gauge *u) {
  GB_API_BEGIN
  gauge_fill(u, 0, 1);
  GB_API_END
}
