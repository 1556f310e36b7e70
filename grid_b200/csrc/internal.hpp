// internal.hpp -- private structures of libgridb200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/gridb200.h"

namespace gb {

// ------------------------------------------------------------------ errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string &m);
#define GB_CUDA(expr)                                                                                        \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess)                                                                                   \
      throw gb::Error(GB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
#define GB_REQUIRE(cond, msg)                                                                                \
  do {                                                                                                       \
    if (!(cond)) throw gb::Error(GB_ERR_INVALID, std::string(msg) + " [" #cond "] at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
// wraps the body of an extern "C" entry point
#define GB_API_BEGIN try {
#define GB_API_END                                                                                           \
  return GB_OK;                                                                                              \
  }                                                                                                          \
  catch (const gb::Error &e) { gb::set_last_error(e.what()); return e.code; }                               \
  catch (const std::exception &e) { gb::set_last_error(e.what()); return GB_ERR_INVALID; }

// ------------------------------------------------------------------ NVTX ranges (ref: Grid/perfmon/Tracing.h:5-70 GRID_TRACE; names follow
// the reference's: Dhop, DhopDag, HaloExchange, Gather, ConjugateGradient, ...).  Header-only NVTX v3: free unless a tool is attached.
struct TraceRange {
  explicit TraceRange(const char *name) { nvtxRangePushA(name); }
  ~TraceRange() { nvtxRangePop(); }
  TraceRange(const TraceRange &) = delete;
};
#define GB_TRACE_CAT2(a, b) a##b
#define GB_TRACE_CAT(a, b) GB_TRACE_CAT2(a, b)
#define GB_TRACE(name) gb::TraceRange GB_TRACE_CAT(gb_trace_, __LINE__)(name)

// ------------------------------------------------------------------ device layout constants
// A fermion field is an array of "vector sites" i5 = site4*Ls + s.  Storage is blocked:
//   element(i5, k) lives at vec index ((i5/W)*NV + k)*W + (i5%W)
// W = 16 lanes; fp32: vec = float4 = two complex, NV = 6; fp64: vec = double2 = one complex, NV = 12.
// With Ls = 16 one block is exactly the 16 fifth-dimension slices of one 4D site, so a half warp reads
// 256 contiguous bytes per spin-colour pair and the gauge link of that site is shared by all 16 lanes.
constexpr int W = 16;
constexpr int LOGW = 4;
template <class T> struct Prec;
template <> struct Prec<float> {
  using real = float; using vec = float4;
  static constexpr int NV = 6;        // vecs per spinor
  static constexpr int LV = 5;        // vecs per stored link (18 reals + 2 pad)
  static constexpr int id = GB_F32;
};
template <> struct Prec<double> {
  using real = double; using vec = double2;
  static constexpr int NV = 12;
  static constexpr int LV = 9;
  static constexpr int id = GB_F64;
};
inline size_t vec_bytes(int prec) { return 16; }
inline int nv_of(int prec) { return prec == GB_F32 ? 6 : 12; }
inline int lv_of(int prec) { return prec == GB_F32 ? 5 : 9; }

// exact unsigned division by a runtime constant (n < 2^32): q = (t + ((n - t) >> s1)) >> s2, t = umulhi(m, n)
struct FastDiv {
  uint32_t d, m, s1, s2;
  FastDiv() : d(1), m(1), s1(0), s2(0) {}
  explicit FastDiv(uint32_t div) : d(div) {
    uint32_t l = 0;
    while ((1ull << l) < div) l++;
    m = (uint32_t)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
    s1 = l < 1 ? l : 1;
    s2 = l - s1;
  }
#ifdef __CUDACC__
  __host__ __device__
#endif
  inline uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    uint32_t t = __umulhi(m, n);
#else
    uint32_t t = (uint32_t)(((uint64_t)m * n) >> 32);
#endif
    return (t + ((n - t) >> s1)) >> s2;
  }
#ifdef __CUDACC__
  __host__ __device__
#endif
  inline void divmod(uint32_t n, uint32_t &q, uint32_t &r) const { q = div(n); r = n - q * d; }
};

} // namespace gb

// ------------------------------------------------------------------ opaque handle definitions
struct ncclComm;
struct gb_context {
  int device = 0;
  cudaStream_t stream = nullptr;       // compute stream (ref: computeStream)
  cudaStream_t comm_stream = nullptr;  // halo / copy stream (ref: copyStream)
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_comm = nullptr, ev_comp = nullptr;
  int64_t launches = 0;
  int sm_count = 148;
  // reductions
  double *d_partials = nullptr; // [max_partials][4]
  double *d_result = nullptr;   // [8]
  double *h_result = nullptr;   // pinned [8]
  double *d_scalars = nullptr;  // [8] device-resident CG scalars
  cudaEvent_t ev_scalar = nullptr;
  int max_partials = 0;
  // L2 flush scratch
  void *l2_scratch = nullptr;
  size_t l2_scratch_bytes = 0;
  // host<->device staging (import/export)
  void *staging = nullptr;
  size_t staging_bytes = 0;
  // recycled field storage (ref: Grid's MemoryManager allocation cache, Grid/allocator/MemoryManagerCache.cc): solver
  // temporaries of 0.4-1.6 GB are created and destroyed per solve; cudaMalloc / cudaFree of that size costs milliseconds and
  // synchronises the device, so freed field buffers are parked here (exact-size reuse, stream ordered on `stream`)
  std::vector<std::pair<size_t, void *>> field_pool;
  size_t field_pool_bytes = 0;
  // communicator
  int rank = 0, nranks = 1;
  ncclComm *nccl = nullptr;
};

struct gb_grid {
  gb_context *ctx;
  int gdims[4], mpi[4], ldims[4], pcoor[4], origin[4];
  int64_t V4, V4cb;
  int nbr_rank[4][2]; // rank of the neighbour in direction mu, [0]=forward(+), [1]=backward(-)
};

struct gb_fermion {
  gb_context *ctx = nullptr; // owner of the storage (checked against the live-context registry on destroy)
  gb_grid *grid;
  int Ls, prec, kind, cb;
  int64_t nsite4;  // 4D sites per parity block (V4cb)
  int64_t n5cb;    // 5D sites per parity block = V4cb*Ls
  int64_t hblk;    // blocks per parity block = ceil(n5cb / W)
  int nparity;     // 1 (half) or 2 (full: [even block][odd block])
  int ncomplex = 12; // complex numbers per site: 12 = SpinColourVector (Wilson types), 3 = ColourVector (staggered, Ls = 1)
  void *data;
  size_t bytes;
  // 16-byte vecs per block of W sites.  Spinors: NV*W.  Staggered colour vectors are stored one complex per element
  // (float2 / double2): element (i, c) at complex index ((i/W)*3 + c)*W + i%W, so a block is 3*W complex = 24 / 48 vecs and
  // the elementwise BLAS / reduction kernels, which only see an array of vecs, serve both site types.
  int64_t vecs_per_block() const { return (int64_t)gb::W * ncomplex * (prec == GB_F32 ? 8 : 16) / 16; }
  int64_t nvec() const { return (int64_t)nparity * hblk * vecs_per_block(); }
  // pointer to the start of parity block p
  void *block(int p) const { return (char *)data + (size_t)p * hblk * vecs_per_block() * 16; }
};

struct gb_gauge {
  gb_grid *grid;
  int prec;
  void *data; // lexicographic [V4][4][3][3] complex of `prec`
  size_t bytes;
};

namespace gb {
// contexts that have been created and not yet destroyed (fields may outlive their context in a host program's teardown)
bool context_alive(const gb_context *ctx);
// launch bookkeeping
inline void count_launch(gb_context *ctx, int n = 1) { ctx->launches += n; }
void check_launch(gb_context *ctx, const char *what);

// ---- fields.cu
void fermion_check_same(const gb_fermion *a, const gb_fermion *b);
// a new zeroed field of the same grid / Ls / site type / full-vs-redblack as `like`, in precision `prec`
gb_fermion *fermion_create_like(const gb_fermion *like, int prec);
// ---- stag.cu: layout-aware pieces for ColourVector fields (ncomplex == 3)
void stag_transfer(const gb_fermion *f, void *stage, int host_prec, int dir);
void stag_random(gb_fermion *f, uint64_t seed);
void stag_precision_change(gb_fermion *out, const gb_fermion *in);
// deterministic reductions; results land in ctx->h_result after sync. n_out doubles.
void reduce_norm2(gb_context *ctx, const gb_fermion *x, double *out);
void reduce_inner(gb_context *ctx, const gb_fermion *l, const gb_fermion *r, double out[2]);
void global_sum(gb_context *ctx, double *v, int n);

// ---- cayley / dhop / fermop
struct CayleyCoeffs {
  int Ls = 0;
  double mass = 0, M5 = 0, b = 1, c = 0;
  std::vector<double> bs, cs, bee, cee, beo, ceo, aee, dee, lee, leem, uee, ueem;
};
CayleyCoeffs cayley_coeffs(int Ls, double mass, double M5, double b, double c);
} // namespace gb

