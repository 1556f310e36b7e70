// mock_backend.cpp -- TEST-ONLY CPU backend behind the C ABI (tests/mock/README.md).
//
// Purpose: run the product's HOST ORCHESTRATION of the SURVEY 8(f) rows -- the actual source files grid_b200/csrc/solver.cu,
// schur.cu, force.cu, nersc.cu, built for the host through tests/mock/shim/cuda_runtime.h and tests/mock/transform.py -- on a
// machine without a GPU, together with the Python mirror and the GPU tests themselves.  What is mocked is everything those
// files CALL: field containers and BLAS / reductions (plain host loops over the same blocked layout), and the operator entry
// points op_apply / dhop_blocks, which are served by the CPU oracle (oracle/liboracle.so, test infrastructure).
// Nothing here ships: the product library has no CPU path (tests/test_abi.py::test_no_cpu_fallback_without_a_device).
#include "fermop.hpp"
#include <complex>
#include <map>
#include <random>
#include <string>
#include <vector>

namespace gb_mock {
thread_local uint3 t_blockIdx, t_threadIdx;
thread_local dim3 t_blockDim, t_gridDim;
}

// ---- the oracle's C API (oracle/oracle_capi.cpp)
extern "C" {
void *orc_op_create(int kind, const int *L, int Ls, double mass, double M5, double b, double c, int prec);
void orc_op_destroy(void *h);
void orc_op_import_gauge(void *h, const void *Umu, const double *phases);
int orc_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half);
void orc_dhop_leg(void *h, const void *in, void *out, int point, int dag, int ocb);
void *orc_stag_create(const int *L, double mass, double c1, double c2, double u0, int prec);
void orc_stag_destroy(void *h);
void orc_stag_import_gauge(void *h, const void *Uthin, const void *Ufat);
int orc_stag_apply(void *h, int which, const void *in, void *out, int dag, int cb_in, int half);
}

using namespace gb;

namespace {
thread_local std::string g_err;
std::map<const gb_fermop *, void *> g_oracle;     // operator -> oracle handle
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
    const int par = (x + y + z + t) & 1;
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
// a half field viewed through raw parity-block pointers (dhop_blocks hands those over)
gb_fermion block_view(const gb_fermop *op, const void *block, int cb) {
  gb_fermion f;
  f.grid = op->grid; f.Ls = op->Ls; f.prec = op->prec; f.kind = GB_HALF; f.cb = cb; f.ncomplex = 12;
  f.nsite4 = op->grid->V4cb; f.n5cb = f.nsite4 * op->Ls; f.hblk = (f.n5cb + W - 1) / W; f.nparity = 1;
  f.data = const_cast<void *>(block); f.bytes = (size_t)f.nvec() * 16;
  return f;
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
gb_fermion *op_tmp_half(gb_fermop *op, int i) {
  if (!op->tmp_h[i]) op->tmp_h[i] = create(op->grid, op->Ls, op->kind == GB_KIND_STAGGERED ? 3 : 12, op->prec, GB_HALF);
  return op->tmp_h[i];
}
gb_fermion *op_tmp_full(gb_fermop *op, int i) {
  if (!op->tmp_f[i]) op->tmp_f[i] = create(op->grid, op->Ls, op->kind == GB_KIND_STAGGERED ? 3 : 12, op->prec, GB_FULL);
  return op->tmp_f[i];
}
// every operator entry point = the oracle on host copies (the mock tests ORCHESTRATION, not these)
void op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag) {
  GB_REQUIRE(op && in && out && in != out, "null or aliased argument");
  GB_REQUIRE(in->grid == op->grid && in->prec == op->prec && in->Ls == op->Ls && out->kind == in->kind && out->prec == in->prec, "field is not conformable with the operator");
  const bool half = in->kind == GB_HALF;
  if ((which == GB_OP_DHOP_OE) && in->cb != GB_EVEN) throw Error(GB_ERR_INVALID, "DhopOE needs an Even-checkerboard input");
  if ((which == GB_OP_DHOP_EO) && in->cb != GB_ODD) throw Error(GB_ERR_INVALID, "DhopEO needs an Odd-checkerboard input");
  const bool flips = which == GB_OP_DHOP_OE || which == GB_OP_DHOP_EO || which == GB_OP_MEOOE || which == GB_OP_MEOOE_DAG;
  out->cb = half ? (flips ? 1 - in->cb : in->cb) : in->cb;
  void *h = g_oracle.at(op);
  auto run = [&](auto tag) {
    using T = decltype(tag);
    std::vector<T> x = to_host<T>(in), y(x.size());
    int rc;
    if (op->kind == GB_KIND_STAGGERED) {
      if (which == GB_OP_DMINUS || which == GB_OP_DMINUS_DAG) { y = x; rc = 0; }
      else rc = orc_stag_apply(h, which, x.data(), y.data(), dag, in->cb, half);
    } else rc = orc_apply(h, which, x.data(), y.data(), dag, in->cb, half);
    GB_REQUIRE(rc == 0, "opcode not served by the oracle");
    transfer<T, T>(out, y.data(), true);
  };
  if (op->prec == GB_F32) run(float()); else run(double());
}
// force.cu drives single legs through dhop_blocks with op->leg_mask = 1 << point
void dhop_blocks(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag, const void *const ax[2], double, double) {
  GB_REQUIRE(ax == nullptr, "mock dhop_blocks: no epilogue");
  int point = -1;
  for (int p = 0; p < 8; p++) if (op->leg_mask == (1 << p)) point = p;
  GB_REQUIRE(point >= 0, "mock dhop_blocks serves single legs only");
  void *h = g_oracle.at(op);
  auto run = [&](auto tag) {
    using T = decltype(tag);
    for (int j = 0; j < nparity; j++) {
      const int po = parity_out_first ^ j, ip = 1 - po;
      gb_fermion fi = block_view(op, in[ip], ip), fo = block_view(op, out[po], po);
      std::vector<T> x = to_host<T>(&fi), y(x.size());
      orc_dhop_leg(h, x.data(), y.data(), point, dag, po);
      transfer<T, T>(&fo, y.data(), true);
    }
  };
  if (op->prec == GB_F32) run(float()); else run(double());
}
template <class T> static void inner_T(const gb_fermion *l, const gb_fermion *r, double out[2]) {
  const T *a = (const T *)l->data, *b = (const T *)r->data;
  double re = 0, im = 0;
  for (size_t i = 0; i < scalars<T>(l); i += 2) { re += (double)(a[i] * b[i] + a[i + 1] * b[i + 1]); im += (double)(a[i] * b[i + 1] - a[i + 1] * b[i]); }
  out[0] = re; out[1] = im;
}
void reduce_inner_dev(gb_context *, const gb_fermion *l, const gb_fermion *r, double *d_out) { if (l->prec == GB_F32) inner_T<float>(l, r, d_out); else inner_T<double>(l, r, d_out); }
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
int gb_comm_unique_id(void *) { MOCK_UNSUPPORTED("gb_comm_unique_id"); }
int gb_comm_init(gb_context *c, int rank, int nranks, const void *) { if (nranks != 1) MOCK_UNSUPPORTED("gb_comm_init (nranks > 1)"); c->rank = rank; c->nranks = 1; return GB_OK; }
int gb_comm_rank(gb_context *, int *r, int *n) { if (r) *r = 0; if (n) *n = 1; return GB_OK; }
int gb_comm_global_sum(gb_context *, double *, int) { return GB_OK; }
int gb_comm_barrier(gb_context *) { return GB_OK; }
int gb_grid_create(gb_context *ctx, const int gdims[4], const int mpi[4], gb_grid **out) {
  for (int d = 0; d < 4; d++) if (mpi[d] != 1 || gdims[d] % 2) return fail(GB_ERR_INVALID, "mock grid: one rank, even extents");
  gb_grid *g = new gb_grid();
  g->ctx = ctx;
  for (int d = 0; d < 4; d++) { g->gdims[d] = g->ldims[d] = gdims[d]; g->mpi[d] = 1; g->pcoor[d] = g->origin[d] = 0; g->nbr_rank[d][0] = g->nbr_rank[d][1] = 0; }
  g->V4 = (int64_t)gdims[0] * gdims[1] * gdims[2] * gdims[3]; g->V4cb = g->V4 / 2;
  *out = g;
  return GB_OK;
}
int gb_geometry_query(const int *, const int *, int, int *, int *, int *) { MOCK_UNSUPPORTED("gb_geometry_query"); }
int gb_grid_destroy(gb_grid *g) { delete g; return GB_OK; }
int gb_grid_local_dims(const gb_grid *g, int l[4]) { for (int d = 0; d < 4; d++) l[d] = g->ldims[d]; return GB_OK; }
int gb_grid_local_origin(const gb_grid *g, int o[4]) { for (int d = 0; d < 4; d++) o[d] = 0; return GB_OK; }
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
int gb_axpby(gb_fermion *z, double a, double b, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y);
  BY_PREC(z, (each<float>(z, [&](size_t i) { return std::fmaf((float)a, ((const float *)x->data)[i], (float)b * ((const float *)y->data)[i]); })),
          (each<double>(z, [&](size_t i) { return std::fma(a, ((const double *)x->data)[i], b * ((const double *)y->data)[i]); })));
  z->cb = x->cb;
  GB_API_END
}
int gb_axpy(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y) {
  GB_API_BEGIN
  fermion_check_same(z, x); fermion_check_same(z, y);
  BY_PREC(z, (each<float>(z, [&](size_t i) { return std::fmaf((float)a, ((const float *)x->data)[i], ((const float *)y->data)[i]); })),
          (each<double>(z, [&](size_t i) { return std::fma(a, ((const double *)x->data)[i], ((const double *)y->data)[i]); })));
  z->cb = x->cb;
  GB_API_END
}
int gb_norm2(const gb_fermion *x, double *out) {
  double s = 0;
  BY_PREC(x, { const float *p = (const float *)x->data; for (size_t i = 0; i < scalars<float>(x); i++) s += (double)(p[i] * p[i]); },
          { const double *p = (const double *)x->data; for (size_t i = 0; i < scalars<double>(x); i++) s += p[i] * p[i]; });
  *out = s;
  return GB_OK;
}
int gb_axpy_norm(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y, double *n2) { int rc = gb_axpy(z, a, x, y); if (rc != GB_OK) return rc; return gb_norm2(z, n2); }
int gb_inner_product(const gb_fermion *l, const gb_fermion *r, double out[2]) { if (l->prec == GB_F32) inner_T<float>(l, r, out); else inner_T<double>(l, r, out); return GB_OK; }
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
int gb_gauge_random(gb_gauge *, uint64_t) { MOCK_UNSUPPORTED("gb_gauge_random"); }
int gb_gauge_unit(gb_gauge *) { MOCK_UNSUPPORTED("gb_gauge_unit"); }
// operators
static gb_fermop *make(gb_grid *g, const gb_gauge *U, int kind, int Ls, double mass, double M5, double b, double c, const double *ph) {
  gb_fermop *op = new gb_fermop();
  op->grid = g; op->ctx = g->ctx; op->kind = kind; op->prec = U->prec; op->Ls = Ls; op->mass = mass; op->M5 = M5;
  op->Uds = (void *)1;   // "has a gauge field"
  void *h = orc_op_create(kind == GB_KIND_WILSON ? 0 : 1, g->ldims, Ls, mass, M5, b, c, U->prec == GB_F32 ? 0 : 1);
  orc_op_import_gauge(h, U->data, ph);
  g_oracle[op] = h;
  return op;
}
int gb_op_create_wilson(gb_grid *g, const gb_gauge *U, double mass, const double *ph, gb_fermop **out) { *out = make(g, U, GB_KIND_WILSON, 1, mass, 0, 1, 0, ph); return GB_OK; }
int gb_op_create_dwf(gb_grid *g, const gb_gauge *U, int Ls, double mass, double M5, const double *ph, gb_fermop **out) { *out = make(g, U, GB_KIND_CAYLEY, Ls, mass, M5, 1, 0, ph); return GB_OK; }
int gb_op_create_mobius(gb_grid *g, const gb_gauge *U, int Ls, double mass, double M5, double b, double c, const double *ph, gb_fermop **out) { *out = make(g, U, GB_KIND_CAYLEY, Ls, mass, M5, b, c, ph); return GB_OK; }
int gb_op_import_gauge(gb_fermop *op, const gb_gauge *U) { orc_op_import_gauge(g_oracle.at(op), U->data, nullptr); return GB_OK; }
int gb_op_create_staggered(gb_grid *g, const gb_gauge *Ut, const gb_gauge *Uf, double mass, double c1, double c2, double u0, gb_fermop **out) {
  gb_fermop *op = new gb_fermop();
  op->grid = g; op->ctx = g->ctx; op->kind = GB_KIND_STAGGERED; op->prec = Ut->prec; op->Ls = 1; op->mass = mass;
  void *h = orc_stag_create(g->ldims, mass, c1, c2, u0, Ut->prec == GB_F32 ? 0 : 1);
  orc_stag_import_gauge(h, Ut->data, Uf->data);
  g_oracle[op] = h;
  *out = op;
  return GB_OK;
}
int gb_op_import_gauge_staggered(gb_fermop *op, const gb_gauge *Ut, const gb_gauge *Uf) { orc_stag_import_gauge(g_oracle.at(op), Ut->data, Uf->data); return GB_OK; }
int gb_op_destroy(gb_fermop *op) {
  if (!op) return GB_OK;
  if (op->kind == GB_KIND_STAGGERED) orc_stag_destroy(g_oracle.at(op)); else orc_op_destroy(g_oracle.at(op));
  g_oracle.erase(op);
  for (auto *f : op->tmp_h) gb_fermion_destroy(f);
  for (auto *f : op->tmp_f) gb_fermion_destroy(f);
  delete op;
  return GB_OK;
}
int gb_op_Ls(const gb_fermop *op) { return op->Ls; }
int gb_op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag) { GB_API_BEGIN op_apply(op, which, in, out, dag); GB_API_END }
int gb_op_dhop_host(gb_fermop *, const void *, void *, gb_precision, int) { MOCK_UNSUPPORTED("gb_op_dhop_host"); }
int gb_op_halo_exchange(gb_fermop *, const gb_fermion *, int, int64_t *) { MOCK_UNSUPPORTED("gb_op_halo_exchange"); }
int gb_op_set_tiling(gb_fermop *, int, int, int) { return GB_OK; }
int gb_op_set_overlap(gb_fermop *, int) { return GB_OK; }
int gb_op_set_fast_kernel(gb_fermop *, int) { return GB_OK; }
}
