// mock_backend.cpp -- TEST-ONLY CPU backend behind the C ABI (tests/mock/README.md).
//
// Purpose: run the product's source files (tests/mock/build_mock.py lists them: operators, generic and tuned hopping kernels, dense
// s-space kernel, peer-to-peer halos, staggered operator, solvers, Schur solve, force terms, NERSC I/O) on a machine without a GPU,
// built for the host through tests/mock/shim/ and tests/mock/transform.py, together with the Python mirror and the GPU tests
// themselves.  What is mocked HERE is the rest: the context, field containers, import / export, BLAS-1 and reductions (plain host
// loops over the same blocked layout; the real ones, fields.cu, use shared memory and warp shuffles) and NCCL (mailboxes and
// barriers between rank threads).  No oracle in here: the tests compare with it.
// Nothing here ships: the product library has no CPU path (tests/test_abi.py::test_no_cpu_fallback_without_a_device).
#include "fermop.hpp"
#include <complex>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

namespace gb_mock {
thread_local uint3 t_blockIdx, t_threadIdx;
thread_local dim3 t_blockDim, t_gridDim;
}

#include "comm.hpp"

using namespace gb;

namespace {
thread_local std::string g_err;
typedef std::complex<double> cd;

template <class T> size_t scalars(const gb_fermion *f) { return (size_t)f->nvec() * (16 / sizeof(T)); }

// scalar (T) offset of complex component c of 5D site i5cb in parity block p (layout: internal.hpp; staggered: stag_halo.cuh cv_index)
template <class T> size_t off(const gb_fermion *f, int p, int64_t i5, int c) {
  if (f->ncomplex == 3) return ((size_t)p * f->hblk * 3 * W + ((size_t)(i5 / W) * 3 + c) * W + i5 % W) * 2;
  constexpr int CPV = sizeof(T) == 4 ? 2 : 1;
  const int NV = sizeof(T) == 4 ? 6 : 12;
  const size_t vec = (size_t)p * f->hblk * NV * W + ((size_t)(i5 / W) * NV + c / CPV) * W + i5 % W;
  return (vec * CPV + c % CPV) * 2;
}
// device layout <-> host order of the C ABI (full: lexicographic, half: checkerboard-lexicographic), as complex<T> arrays
template <class T, class H> void transfer(const gb_fermion *f, H *host, bool to_device) {
  const int *L = f->grid->ldims;
  T *raw = (T *)f->data;
  const int nc = f->ncomplex, Ls = f->Ls;
  if (to_device) std::fill(raw, raw + scalars<T>(f), (T)0);
  for (int t = 0; t < L[3]; t++) for (int z = 0; z < L[2]; z++) for (int y = 0; y < L[1]; y++) for (int x = 0; x < L[0]; x++) {
    const int par = (x + y + z + t + f->grid->origin[0] + f->grid->origin[1] + f->grid->origin[2] + f->grid->origin[3]) & 1;   // parity of the GLOBAL site
    if (f->kind == GB_HALF && par != f->cb) continue;
    const int64_t site = (x >> 1) + (int64_t)(L[0] / 2) * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
    const int64_t i4 = x + (int64_t)L[0] * (y + (int64_t)L[1] * (z + (int64_t)L[2] * t));
    const int p = f->kind == GB_HALF ? 0 : par;
    const int64_t hsite = f->kind == GB_HALF ? site : i4;
    for (int s = 0; s < Ls; s++) for (int c = 0; c < nc; c++) {
      H *h = host + ((size_t)(hsite * Ls + s) * nc + c) * 2;
      const size_t o = off<T>(f, p, site * Ls + s, c);
      if (to_device) { raw[o] = (T)h[0]; raw[o + 1] = (T)h[1]; } else { h[0] = (H)raw[o]; h[1] = (H)raw[o + 1]; }
    }
  }
}
template <class T> std::vector<T> to_host(const gb_fermion *f) {
  std::vector<T> h((size_t)f->n5cb * f->nparity * f->ncomplex * 2);
  transfer<T, T>(f, h.data(), false);
  return h;
}
int fail(int code, const std::string &m) { g_err = m; return code; }
#define MOCK_UNSUPPORTED(name) return fail(GB_ERR_INVALID, std::string(name) + ": not part of the CPU mock backend")
} // namespace

namespace gb {
void set_last_error(const std::string &m) { g_err = m; }
bool context_alive(const gb_context *) { return true; }
void check_launch(gb_context *, const char *) {}
void fermion_check_same(const gb_fermion *a, const gb_fermion *b) {
  GB_REQUIRE(a && b, "null field");
  GB_REQUIRE(a->grid == b->grid && a->Ls == b->Ls && a->kind == b->kind && a->prec == b->prec && a->ncomplex == b->ncomplex, "fields are not conformable");
}
static gb_fermion *create(gb_grid *g, int Ls, int ncomplex, int prec, int kind) {
  gb_fermion *f = new gb_fermion();
  f->ctx = g->ctx; f->grid = g; f->Ls = Ls; f->prec = prec; f->kind = kind; f->cb = GB_EVEN; f->ncomplex = ncomplex;
  f->nsite4 = g->V4cb; f->n5cb = g->V4cb * Ls; f->hblk = (f->n5cb + W - 1) / W; f->nparity = kind == GB_HALF ? 1 : 2;
  f->bytes = (size_t)f->nvec() * 16;
  f->data = std::calloc(f->bytes, 1);
  return f;
}
gb_fermion *fermion_create_like(const gb_fermion *like, int prec) {
  gb_fermion *f = create(like->grid, like->Ls, like->ncomplex, prec, like->kind);
  f->cb = like->cb;
  return f;
}
// ---- nothing is switched off any more: the tuned kernels, the dense s-space kernel and the peer-to-peer halos are the product's own
//      (dhop_fast.cu, smat.cu, halo_p2p.cu); cooperative kernels run block by block on fibres (simt.cpp), "peer mappings" are plain
//      pointers because the ranks are threads of one process (shim: cudaIpc*)
} // namespace gb
// ---- "ranks" are host threads of one process (each drives its own context); send / recv are copies through mailboxes
struct ncclComm { int rank, nranks; };
namespace {
struct World {
  std::mutex m;
  std::condition_variable cv;
  std::map<std::pair<int, int>, std::deque<std::vector<char>>> box;   // (source, destination) -> messages in order
  int arrived = 0; long gen = 0;
  std::vector<double> acc, res;
  int ag_arrived = 0; long ag_gen = 0;
  std::vector<char> ag_buf, ag_res;
} g_world;
ncclResult_t mock_send(const void *buf, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  const size_t bytes = n * (t == ncclDouble ? 8 : 1);
  std::lock_guard<std::mutex> l(g_world.m);
  g_world.box[{c->rank, peer}].emplace_back((const char *)buf, (const char *)buf + bytes);
  g_world.cv.notify_all();
  return 0;
}
ncclResult_t mock_recv(void *buf, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  const size_t bytes = n * (t == ncclDouble ? 8 : 1);
  std::unique_lock<std::mutex> l(g_world.m);
  auto &q = g_world.box[{peer, c->rank}];
  g_world.cv.wait(l, [&] { return !q.empty(); });
  if (q.front().size() != bytes) return 1;
  std::memcpy(buf, q.front().data(), bytes);
  q.pop_front();
  return 0;
}
ncclResult_t mock_allgather(const void *send, void *recv, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
  const size_t bytes = n * (t == ncclDouble ? 8 : 1);
  std::unique_lock<std::mutex> l(g_world.m);
  if (g_world.ag_arrived == 0) g_world.ag_buf.assign(bytes * c->nranks, 0);
  std::memcpy(g_world.ag_buf.data() + (size_t)c->rank * bytes, send, bytes);   // (send may point into recv: copied before recv is written)
  const long mygen = g_world.ag_gen;
  if (++g_world.ag_arrived == c->nranks) { g_world.ag_res = g_world.ag_buf; g_world.ag_arrived = 0; g_world.ag_gen++; g_world.cv.notify_all(); }
  else g_world.cv.wait(l, [&] { return g_world.ag_gen != mygen; });
  std::memcpy(recv, g_world.ag_res.data(), bytes * c->nranks);
  return 0;
}
ncclResult_t mock_group() { return 0; }
const char *mock_errstr(ncclResult_t) { return "mock nccl error (message size mismatch)"; }
} // namespace
namespace gb {
NcclApi &nccl() {
  static NcclApi api;
  if (!api.ok) { api.ok = true; api.Send = mock_send; api.Recv = mock_recv; api.AllGather = mock_allgather; api.GroupStart = mock_group; api.GroupEnd = mock_group; api.GetErrorString = mock_errstr; }
  return api;
}
void nccl_check(int r, const char *what) { if (r != 0) throw Error(GB_ERR_COMM, what); }
// sum over the rank threads (all of them call it the same number of times, like an all-reduce)
void global_sum(gb_context *ctx, double *v, int n) {
  if (ctx->nranks <= 1) return;
  std::unique_lock<std::mutex> l(g_world.m);
  if (g_world.arrived == 0) g_world.acc.assign(n, 0.0);
  for (int i = 0; i < n; i++) g_world.acc[i] += v[i];
  const long mygen = g_world.gen;
  if (++g_world.arrived == ctx->nranks) { g_world.res = g_world.acc; g_world.arrived = 0; g_world.gen++; g_world.cv.notify_all(); }
  else g_world.cv.wait(l, [&] { return g_world.gen != mygen; });
  for (int i = 0; i < n; i++) v[i] = g_world.res[i];
}
void device_global_sum(gb_context *ctx, double *d_vals, int n) { global_sum(ctx, d_vals, n); }   // "device" scalars are host memory here
template <class T> static void inner_T(const gb_fermion *l, const gb_fermion *r, double out[2]) {
  const T *a = (const T *)l->data, *b = (const T *)r->data;
  long double re = 0, im = 0;   // site products in working precision, lattice sum in extended precision (the real reductions are trees)
  for (size_t i = 0; i < scalars<T>(l); i += 2) { re += (double)(a[i] * b[i] + a[i + 1] * b[i + 1]); im += (double)(a[i] * b[i + 1] - a[i + 1] * b[i]); }
  out[0] = (double)re; out[1] = (double)im;
}
void reduce_inner_dev(gb_context *ctx, const gb_fermion *l, const gb_fermion *r, double *d_out) { if (l->prec == GB_F32) inner_T<float>(l, r, d_out); else inner_T<double>(l, r, d_out); global_sum(ctx, d_out, 2); }
void axpy_norm_dev(gb_context *, gb_fermion *z, const gb_fermion *x, const gb_fermion *y, const double *d_c, const double *d_d, double *d_out) {
  double n2;
  gb_axpy_norm(z, -(*d_c) / (*d_d), x, y, &n2);
  d_out[0] = n2;
}
void cg_update_dev(gb_context *, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, const double *d_c, const double *d_d, const double *d_cp) {
  const double a = (*d_c) / (*d_d), b = (*d_cp) / (*d_c);
  gb_axpy(psi, a, p, psi);
  gb_axpy(p, b, p, r);
}
} // namespace gb

template <class T, class F> static void each(gb_fermion *z, F f) { T *p = (T *)z->data; for (size_t i = 0; i < scalars<T>(z); i++) p[i] = f(i); }
#define BY_PREC(f, expr_f, expr_d) do { if ((f)->prec == GB_F32) { expr_f; } else { expr_d; } } while (0)

extern "C" {
const char *gb_last_error(void) { return g_err.c_str(); }
int gb_device_count(void) { return 1; }
int gb_context_create(int device, gb_context **out) {
  gb_context *c = new gb_context();
  if (getenv("GB_MOCK_SM_COUNT")) c->sm_count = std::max(1, atoi(getenv("GB_MOCK_SM_COUNT")));   // small values make persistent kernels loop over several tiles
  c->device = device;
  c->d_scalars = (double *)std::calloc(8, sizeof(double)); c->h_result = (double *)std::calloc(8, sizeof(double));
  *out = c;
  return GB_OK;
}
int gb_context_destroy(gb_context *c) { if (c) { std::free(c->d_scalars); std::free(c->h_result); delete c; } return GB_OK; }
int gb_synchronize(gb_context *) { return GB_OK; }
int gb_timer_start(gb_context *) { return GB_OK; }
int gb_timer_stop(gb_context *, double *ms) { if (ms) *ms = 0; return GB_OK; }
int64_t gb_launch_count(gb_context *c) { return c->launches; }
int gb_flush_l2(gb_context *) { return GB_OK; }
int gb_comm_unique_id(void *id) { std::memset(id, 0, GB_UNIQUE_ID_BYTES); return GB_OK; }
int gb_comm_init(gb_context *c, int rank, int nranks, const void *) {
  c->rank = rank; c->nranks = nranks;
  if (nranks > 1) c->nccl = new ncclComm{rank, nranks};     // the "communicator" of a rank thread
  return GB_OK;
}
int gb_comm_rank(gb_context *c, int *r, int *n) { if (r) *r = c->rank; if (n) *n = c->nranks; return GB_OK; }
int gb_comm_global_sum(gb_context *c, double *v, int n) { global_sum(c, v, n); return GB_OK; }
int gb_comm_barrier(gb_context *c) { double v = 0; global_sum(c, &v, 1); return GB_OK; }
// rank -> processor coordinate lexicographic with dimension 0 fastest, the library's rule (include/gridb200.h)
int gb_geometry_query(const int gdims[4], const int mpi[4], int rank, int ldims[4], int origin[4], int nbr[8]) {
  int pc[4], r = rank;
  for (int d = 0; d < 4; d++) { if (mpi[d] < 1 || gdims[d] % mpi[d] || (gdims[d] / mpi[d]) % 2) return fail(GB_ERR_INVALID, "bad decomposition"); pc[d] = r % mpi[d]; r /= mpi[d]; }
  auto rk = [&](const int *p) { return p[0] + mpi[0] * (p[1] + mpi[1] * (p[2] + mpi[2] * p[3])); };
  for (int d = 0; d < 4; d++) {
    ldims[d] = gdims[d] / mpi[d]; origin[d] = pc[d] * ldims[d];
    int q[4] = {pc[0], pc[1], pc[2], pc[3]};
    q[d] = (pc[d] + 1) % mpi[d]; nbr[2 * d] = rk(q);
    q[d] = (pc[d] + mpi[d] - 1) % mpi[d]; nbr[2 * d + 1] = rk(q);
  }
  return GB_OK;
}
int gb_grid_create(gb_context *ctx, const int gdims[4], const int mpi[4], gb_grid **out) {
  gb_grid *g = new gb_grid();
  g->ctx = ctx;
  int nbr[8];
  if (gb_geometry_query(gdims, mpi, ctx->rank, g->ldims, g->origin, nbr) != GB_OK) { delete g; return GB_ERR_INVALID; }
  for (int d = 0; d < 4; d++) { g->gdims[d] = gdims[d]; g->mpi[d] = mpi[d]; g->pcoor[d] = g->origin[d] / g->ldims[d]; g->nbr_rank[d][0] = nbr[2 * d]; g->nbr_rank[d][1] = nbr[2 * d + 1]; }
  g->V4 = (int64_t)g->ldims[0] * g->ldims[1] * g->ldims[2] * g->ldims[3]; g->V4cb = g->V4 / 2;
  *out = g;
  return GB_OK;
}
int gb_grid_destroy(gb_grid *g) { delete g; return GB_OK; }
int gb_grid_local_dims(const gb_grid *g, int l[4]) { for (int d = 0; d < 4; d++) l[d] = g->ldims[d]; return GB_OK; }
int gb_grid_local_origin(const gb_grid *g, int o[4]) { for (int d = 0; d < 4; d++) o[d] = g->origin[d]; return GB_OK; }
int gb_fermion_create(gb_grid *g, int Ls, gb_precision prec, gb_gridkind kind, gb_fermion **out) { *out = create(g, Ls, 12, prec, kind); return GB_OK; }
int gb_staggered_fermion_create(gb_grid *g, gb_precision prec, gb_gridkind kind, gb_fermion **out) { *out = create(g, 1, 3, prec, kind); return GB_OK; }
int gb_fermion_destroy(gb_fermion *f) { if (f) { std::free(f->data); delete f; } return GB_OK; }
int gb_fermion_checkerboard(const gb_fermion *f) { return f->cb; }
int gb_fermion_set_checkerboard_tag(gb_fermion *f, int cb) { f->cb = cb & 1; return GB_OK; }
int64_t gb_fermion_local_sites(const gb_fermion *f) { return f->n5cb * f->nparity; }
int gb_fermion_import(gb_fermion *f, const void *host, gb_precision hp) {
  if (f->prec == GB_F32) { if (hp == GB_F32) transfer<float, float>(f, (float *)host, true); else transfer<float, double>(f, (double *)host, true); }
  else { if (hp == GB_F32) transfer<double, float>(f, (float *)host, true); else transfer<double, double>(f, (double *)host, true); }
  return GB_OK;
}
int gb_fermion_export(const gb_fermion *f, void *host, gb_precision hp) {
  if (f->prec == GB_F32) { if (hp == GB_F32) transfer<float, float>(f, (float *)host, false); else transfer<float, double>(f, (double *)host, false); }
  else { if (hp == GB_F32) transfer<double, float>(f, (float *)host, false); else transfer<double, double>(f, (double *)host, false); }
  return GB_OK;
}
int gb_pick_checkerboard(int cb, gb_fermion *half, const gb_fermion *full) {
  GB_API_BEGIN
  GB_REQUIRE(half->kind == GB_HALF && full->kind == GB_FULL && half->Ls == full->Ls && half->prec == full->prec && half->ncomplex == full->ncomplex, "pickCheckerboard(cb, half, full)");
  std::memcpy(half->data, full->block(cb & 1), half->bytes);
  half->cb = cb & 1;
  GB_API_END
}
int gb_set_checkerboard(gb_fermion *full, const gb_fermion *half) {
  GB_API_BEGIN
  GB_REQUIRE(half->kind == GB_HALF && full->kind == GB_FULL && half->Ls == full->Ls && half->prec == full->prec && half->ncomplex == full->ncomplex, "setCheckerboard(full, half)");
  std::memcpy(full->block(half->cb), half->data, half->bytes);
  GB_API_END
}
int gb_precision_change(gb_fermion *out, const gb_fermion *in) {
  GB_API_BEGIN
  GB_REQUIRE(out->grid == in->grid && out->Ls == in->Ls && out->kind == in->kind && out->ncomplex == in->ncomplex, "fields are not conformable");
  out->cb = in->cb;
  if (in->prec == GB_F64) { std::vector<double> h = to_host<double>(in); gb_fermion_import(out, h.data(), GB_F64); }
  else { std::vector<float> h = to_host<float>(in); gb_fermion_import(out, h.data(), GB_F32); }
  GB_API_END
}
int gb_fermion_random(gb_fermion *f, uint64_t seed) {
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  std::vector<double> h((size_t)f->n5cb * f->nparity * f->ncomplex * 2);
  for (auto &v : h) v = u(gen);
  return gb_fermion_import(f, h.data(), GB_F64);
}
int gb_zero(gb_fermion *z) { std::memset(z->data, 0, z->bytes); return GB_OK; }
int gb_copy(gb_fermion *z, const gb_fermion *x) { GB_API_BEGIN fermion_check_same(z, x); std::memcpy(z->data, x->data, x->bytes); z->cb = x->cb; GB_API_END }
int gb_scale(gb_fermion *z, double a, const gb_fermion *x) {
  GB_API_BEGIN
  fermion_check_same(z, x);
  BY_PREC(z, (each<float>(z, [&](size_t i) { return (float)a * ((const float *)x->data)[i]; })), (each<double>(z, [&](size_t i) { return a * ((const double *)x->data)[i]; })));
  z->cb = x->cb;
  GB_API_END
}
static void check_same_cb(const gb_fermion *a, const gb_fermion *b) {   // as fields.cu: the reference's conformable()
  if (a->kind == GB_HALF && a->cb != b->cb) throw Error(GB_ERR_INVALID, "fields live on different checkerboards (Even vs Odd): not conformable");
}
int gb_axpby(gb_fermion *z, double a, double b, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y); check_same_cb(x, y);
  BY_PREC(z, (each<float>(z, [&](size_t i) { return std::fmaf((float)a, ((const float *)x->data)[i], (float)b * ((const float *)y->data)[i]); })),
          (each<double>(z, [&](size_t i) { return std::fma(a, ((const double *)x->data)[i], b * ((const double *)y->data)[i]); })));
  z->cb = x->cb;
  GB_API_END
}
int gb_axpy(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y); check_same_cb(x, y);
  BY_PREC(z, (each<float>(z, [&](size_t i) { return std::fmaf((float)a, ((const float *)x->data)[i], ((const float *)y->data)[i]); })),
          (each<double>(z, [&](size_t i) { return std::fma(a, ((const double *)x->data)[i], ((const double *)y->data)[i]); })));
  z->cb = x->cb;
  GB_API_END
}
int gb_norm2(const gb_fermion *x, double *out) {
  long double s = 0;
  BY_PREC(x, { const float *p = (const float *)x->data; for (size_t i = 0; i < scalars<float>(x); i++) s += (double)(p[i] * p[i]); },
          { const double *p = (const double *)x->data; for (size_t i = 0; i < scalars<double>(x); i++) s += p[i] * p[i]; });
  *out = (double)s;
  global_sum(x->grid->ctx, out, 1);
  return GB_OK;
}
int gb_axpy_norm(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y, double *n2) { int rc = gb_axpy(z, a, x, y); if (rc != GB_OK) return rc; return gb_norm2(z, n2); }
int gb_inner_product(const gb_fermion *l, const gb_fermion *r, double out[2]) {
  GB_API_BEGIN
  check_same_cb(l, r);
  if (l->prec == GB_F32) inner_T<float>(l, r, out); else inner_T<double>(l, r, out);
  global_sum(l->grid->ctx, out, 2);
  GB_API_END
}
// gauge fields: lexicographic host arrays of the field's precision
int gb_gauge_create(gb_grid *g, gb_precision prec, gb_gauge **out) {
  gb_gauge *u = new gb_gauge();
  u->grid = g; u->prec = prec; u->bytes = (size_t)g->V4 * 72 * (prec == GB_F32 ? 4 : 8); u->data = std::calloc(u->bytes, 1);
  *out = u;
  return GB_OK;
}
int gb_gauge_destroy(gb_gauge *u) { if (u) { std::free(u->data); delete u; } return GB_OK; }
int gb_gauge_import(gb_gauge *u, const void *host, gb_precision hp) {
  const size_t n = (size_t)u->grid->V4 * 72;
  for (size_t i = 0; i < n; i++) {
    const double v = hp == GB_F32 ? ((const float *)host)[i] : ((const double *)host)[i];
    if (u->prec == GB_F32) ((float *)u->data)[i] = (float)v; else ((double *)u->data)[i] = v;
  }
  return GB_OK;
}
int gb_gauge_export(const gb_gauge *u, void *host, gb_precision hp) {
  const size_t n = (size_t)u->grid->V4 * 72;
  for (size_t i = 0; i < n; i++) {
    const double v = u->prec == GB_F32 ? ((const float *)u->data)[i] : ((const double *)u->data)[i];
    if (hp == GB_F32) ((float *)host)[i] = (float)v; else ((double *)host)[i] = v;
  }
  return GB_OK;
}
// some random SU(3) per link (Gram-Schmidt of a Gaussian matrix, determinant rotated to one); not the library's generator
int gb_gauge_random(gb_gauge *u, uint64_t seed) {
  std::mt19937_64 gen(seed * 2654435761ull + 17);
  std::normal_distribution<double> g(0.0, 1.0);
  std::vector<double> h((size_t)u->grid->V4 * 72);
  for (size_t l = 0; l < (size_t)u->grid->V4 * 4; l++) {
    cd m[3][3];
    for (auto &r : m) for (auto &c : r) c = cd(g(gen), g(gen));
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < i; j++) { cd d = 0; for (int k = 0; k < 3; k++) d += std::conj(m[j][k]) * m[i][k]; for (int k = 0; k < 3; k++) m[i][k] -= d * m[j][k]; }
      double n = 0; for (int k = 0; k < 3; k++) n += std::norm(m[i][k]);
      for (int k = 0; k < 3; k++) m[i][k] /= std::sqrt(n);
    }
    const cd det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    for (int k = 0; k < 3; k++) m[2][k] /= det;
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { h[(l * 9 + 3 * i + k) * 2] = m[i][k].real(); h[(l * 9 + 3 * i + k) * 2 + 1] = m[i][k].imag(); }
  }
  return gb_gauge_import(u, h.data(), GB_F64);
}
int gb_gauge_unit(gb_gauge *u) {
  std::vector<double> h((size_t)u->grid->V4 * 72, 0.0);
  for (size_t l = 0; l < (size_t)u->grid->V4 * 4; l++) for (int i = 0; i < 3; i++) h[(l * 9 + 4 * i) * 2] = 1.0;
  return gb_gauge_import(u, h.data(), GB_F64);
}
}
