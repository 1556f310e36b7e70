// smat.cu -- fifth-dimension operators as dense Ls x Ls real matrices per chirality, applied at HBM speed.
//
// Every s-space operator of the Cayley action (ref: CayleyFermion5Dcache.h:43-230 M5D / M5Ddag / MooeeInv / MooeeInvDag
// and the coefficient sets of CayleyFermion5DImplementation.h:156-271) is linear in s, diagonal in 4D and in colour, and
// block-diagonal in chirality.  It is therefore two real Ls x Ls matrices (upper spins = P+ block, lower spins = P-
// block) -- the MatpInv/MatmInv view the reference keeps commented out (CayleyFermion5DImplementation.h:532-534).
// Products such as Meooe5D * MooeeInv are folded on the host, so the Schur operator needs one pass over the field per
// hop instead of the reference's M5D + MooeeInv + M5D launches.
//
//   out = M x  [+ N y]  [+ alpha z]        M, N : (Mp, Mm) pairs;  x, y, z, out : fields on the same grid
//
// Kernel: CTA = 16 sites x LS lanes.  The CTA's x (and y) tiles are brought into shared memory by per-site TMA bulk
// copies (site stride padded by 16 B so the two sites of a warp sit in different banks); thread (site, s) keeps row s of
// the matrices in registers and accumulates with packed f32x2 FMAs, reading x[site][k][s'] as a broadcast LDS.128.
#include "fermop.hpp"
#include "dhop_fast.cuh"
#include "kernels_common.cuh"
#include <algorithm>

namespace gb {

constexpr int SM_NSITE = 16;

template <class T, int LS> struct SMatRows { T mp[LS], mm[LS]; };

// Epilogues that fold the conjugate-gradient linear algebra of the Schur solve into the s-space passes (solver.cu):
//   EPI_NORM  : out = M x [+ N y] [+ alpha z], and |out|^2 is reduced into *d_out (this is d = |Mpc p|^2 = <p, MpcDag Mpc p>)
//   EPI_RUPD  : q = M x [+ N y] is NOT stored: out = z - (c/d) q (the residual update r -= a A p), |out|^2 reduced into *d_out
//   EPI_CGUPD : x = r, y = p, N = M: b = cp/c, a = c/d; out = M (r + b p) (the first s-space pass of the next A p), and on the way
//               psi += a p, p = r + b p (ref: ConjugateGradient.h:176-183) for the thread's own element
enum { EPI_NONE = 0, EPI_NORM = 1, EPI_RUPD = 2, EPI_CGUPD = 3 };
template <class T> struct SMatArgs {
  const typename Prec<T>::vec *x, *y, *z;
  typename Prec<T>::vec *out;
  const T *M, *N;        // device: [2][LS][LS] (chirality, row, col), row-major
  T alpha;
  uint32_t nsite;        // 4D sites in this parity block
  size_t block_stride;   // vecs between parity blocks
  // epilogues
  const double *d_c = nullptr, *d_d = nullptr, *d_cp = nullptr;   // device-resident CG scalars
  double *partials = nullptr;                                    // one slot per CTA
  typename Prec<T>::vec *psi = nullptr, *p = nullptr;            // EPI_CGUPD
};

template <class V> __device__ __forceinline__ V v_fma(float a, V x, V acc);
template <> __device__ __forceinline__ float4 v_fma<float4>(float a, float4 x, float4 acc) {
  const f2 aa = pk(a, a);
  f2 lo = fma2(aa, pk(x.x, x.y), pk(acc.x, acc.y)), hi = fma2(aa, pk(x.z, x.w), pk(acc.z, acc.w));
  float4 r; upk(lo, r.x, r.y); upk(hi, r.z, r.w); return r;
}
__device__ __forceinline__ double2 v_fmad(double a, double2 x, double2 acc) { return make_double2(fma(a, x.x, acc.x), fma(a, x.y, acc.y)); }

template <class T, int LS, int NIN, int EPI>
__global__ void __launch_bounds__(SM_NSITE *LS) smat_kernel(const SMatArgs<T> a) {
  using P = Prec<T>;
  using V = typename P::vec;
  constexpr int SITE_VECS = P::NV * LS;          // vecs per 4D site
  constexpr int SSTRIDE = SITE_VECS + 1;         // padded smem stride (one extra 16-byte vec)
  constexpr int STAGE_VECS = NIN * SM_NSITE * SSTRIDE;
  // the layout is site-major only when LS is a multiple of the 16-lane block; then tiles stream through a
  // two-stage TMA pipeline in a persistent CTA, otherwise one tile per CTA with element loads
  constexpr bool CONTIG = (LS % W) == 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V *stage0 = reinterpret_cast<V *>(smem_raw);
  __shared__ uint64_t bar[2];
  const int sl = threadIdx.x / LS, s = threadIdx.x % LS;
  const size_t boff = (size_t)blockIdx.y * a.block_stride;
  const uint32_t ntiles = (a.nsite + SM_NSITE - 1) / SM_NSITE;
  // matrix rows of this thread's s (both chiralities)
  T mp[LS], mm[LS], np_[NIN == 2 ? LS : 1], nm_[NIN == 2 ? LS : 1];
#pragma unroll
  for (int j = 0; j < LS; j++) {
    mp[j] = a.M[s * LS + j]; mm[j] = a.M[LS * LS + s * LS + j];
    if (NIN == 2) { np_[j] = a.N[s * LS + j]; nm_[j] = a.N[LS * LS + s * LS + j]; }
  }
  T cg_a = 0, cg_b = 0;
  T alpha = a.alpha;
  double nrm = 0;
  if (EPI == EPI_RUPD) alpha = (T)(-(*a.d_c) / (*a.d_d));
  if (EPI == EPI_CGUPD) {
    cg_a = (T)((*a.d_c) / (*a.d_d)); cg_b = (T)((*a.d_cp) / (*a.d_c));
    if (NIN == 2) {
#pragma unroll
      for (int j = 0; j < LS; j++) { np_[j] *= cg_b; nm_[j] *= cg_b; }
    }
  }
  auto issue = [&](uint32_t tile, int st) { // called by all threads; thread 0 arms the barrier, lane s==0 of each site copies
    uint32_t site = tile * SM_NSITE + sl;
    if (site >= a.nsite) site = a.nsite - 1;
    V *sx = stage0 + st * STAGE_VECS;
    if (threadIdx.x == 0) mbar_expect_tx(&bar[st], SM_NSITE * SITE_VECS * 16 * NIN);
    __syncwarp();
    if (s == 0) {
      bulk_g2s(sx + sl * SSTRIDE, a.x + boff + (size_t)site * SITE_VECS, SITE_VECS * 16, &bar[st]);
      if (NIN == 2) bulk_g2s(sx + SM_NSITE * SSTRIDE + sl * SSTRIDE, a.y + boff + (size_t)site * SITE_VECS, SITE_VECS * 16, &bar[st]);
    }
  };
  if (CONTIG) {
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    __syncthreads();
  }
  uint32_t it = 0;
  if (CONTIG && blockIdx.x < ntiles) {
    if (threadIdx.x == 0) mbar_expect_tx(&bar[0], SM_NSITE * SITE_VECS * 16 * NIN);
    __syncthreads(); // expect_tx precedes every copy of this phase
    uint32_t site0 = blockIdx.x * SM_NSITE + sl;
    if (site0 >= a.nsite) site0 = a.nsite - 1;
    if (s == 0) {
      bulk_g2s(stage0 + sl * SSTRIDE, a.x + boff + (size_t)site0 * SITE_VECS, SITE_VECS * 16, &bar[0]);
      if (NIN == 2) bulk_g2s(stage0 + SM_NSITE * SSTRIDE + sl * SSTRIDE, a.y + boff + (size_t)site0 * SITE_VECS, SITE_VECS * 16, &bar[0]);
    }
  }
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int st = CONTIG ? (it & 1) : 0;
    V *sx = stage0 + st * STAGE_VECS;
    V *sy = sx + SM_NSITE * SSTRIDE;
    uint32_t site = tile * SM_NSITE + sl;
    const bool active = site < a.nsite;
    if (!active) site = a.nsite - 1;
    const uint32_t i5 = site * LS + s;
    const size_t g = boff + ((size_t)(i5 >> LOGW) * P::NV << LOGW) + (i5 & (W - 1));
    if (CONTIG) {
      // prefetch the next tile into the other stage (its readers finished at the __syncthreads closing the previous iteration)
      const uint32_t nxt = tile + gridDim.x;
      if (nxt < ntiles) {
        if (threadIdx.x == 0) mbar_expect_tx(&bar[st ^ 1], SM_NSITE * SITE_VECS * 16 * NIN);
        __syncthreads();
        uint32_t sn = nxt * SM_NSITE + sl;
        if (sn >= a.nsite) sn = a.nsite - 1;
        V *nx = stage0 + (st ^ 1) * STAGE_VECS;
        if (s == 0) {
          bulk_g2s(nx + sl * SSTRIDE, a.x + boff + (size_t)sn * SITE_VECS, SITE_VECS * 16, &bar[st ^ 1]);
          if (NIN == 2) bulk_g2s(nx + SM_NSITE * SSTRIDE + sl * SSTRIDE, a.y + boff + (size_t)sn * SITE_VECS, SITE_VECS * 16, &bar[st ^ 1]);
        }
      }
      mbar_wait(&bar[st], (it >> 1) & 1);
    } else {
#pragma unroll
      for (int k = 0; k < P::NV; k++) {
        sx[sl * SSTRIDE + k * LS + s] = a.x[g + ((size_t)k << LOGW)];
        if (NIN == 2) sy[sl * SSTRIDE + k * LS + s] = a.y[g + ((size_t)k << LOGW)];
      }
      __syncthreads();
    }
    // smem element (site, k, s'): CONTIG tiles keep the global order [blk][k][lane] with blk = LS/16 blocks per site
    auto sidx = [&](int k, int sp) -> int {
      if (CONTIG) return sl * SSTRIDE + ((sp >> LOGW) * P::NV + k) * W + (sp & (W - 1));
      return sl * SSTRIDE + k * LS + sp;
    };
#pragma unroll
    for (int k = 0; k < P::NV; k++) {
      const bool upper = k < P::NV / 2;
      V acc;
      if constexpr (sizeof(T) == 4) acc = make_float4(0.f, 0.f, 0.f, 0.f); else acc = make_double2(0., 0.);
#pragma unroll
      for (int j = 0; j < LS; j++) {
        const V xv = sx[sidx(k, j)];
        if constexpr (sizeof(T) == 4) acc = v_fma<float4>(upper ? mp[j] : mm[j], xv, acc); else acc = v_fmad(upper ? mp[j] : mm[j], xv, acc);
        if (NIN == 2) {
          const V yv = sy[sidx(k, j)];
          if constexpr (sizeof(T) == 4) acc = v_fma<float4>(upper ? np_[j] : nm_[j], yv, acc); else acc = v_fmad(upper ? np_[j] : nm_[j], yv, acc);
        }
      }
      if (EPI == EPI_RUPD) {
        const V zv = a.z[g + ((size_t)k << LOGW)];
        if constexpr (sizeof(T) == 4) acc = v_fma<float4>(alpha, acc, zv); else acc = v_fmad(alpha, acc, zv);
      } else if (a.z != nullptr) {
        const V zv = a.z[g + ((size_t)k << LOGW)];
        if constexpr (sizeof(T) == 4) acc = v_fma<float4>(alpha, zv, acc); else acc = v_fmad(alpha, zv, acc);
      }
      if (active) {
        a.out[g + ((size_t)k << LOGW)] = acc;
        if (EPI == EPI_NORM || EPI == EPI_RUPD) nrm += vnorm2(acc);
        if (EPI == EPI_CGUPD && NIN == 2) {
          const V rv = sx[sidx(k, s)], pv = sy[sidx(k, s)];
          a.psi[g + ((size_t)k << LOGW)] = vaxpy(cg_a, pv, a.psi[g + ((size_t)k << LOGW)]);
          a.p[g + ((size_t)k << LOGW)] = vaxpy(cg_b, pv, rv);
        }
      }
    }
    __syncthreads(); // every thread is done with this stage before it is refilled
  }
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    // fixed-shape block tree, one partial per CTA; the second stage (smat_reduce_kernel) adds them in a fixed order
    __shared__ double red[SM_NSITE * LS / 32 + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = nrm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int w = 0; w < (SM_NSITE * LS + 31) / 32; w++) t += red[w];
      a.partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}
__global__ void smat_reduce_kernel(const double *partials, int n, double *result) {
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partials[i];
  __shared__ double sm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) sm[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; w++) t += sm[w];
    *result = t;
  }
}

// ------------------------------------------------------------------ host: dense matrices of the s-space operators
// A matrix is [2][Ls][Ls]: block 0 acts on the upper (P+) spin components, block 1 on the lower (P-) ones.
SMat smat_identity(int Ls) {
  SMat m; m.Ls = Ls; m.a.assign(2 * Ls * Ls, 0.0);
  for (int c = 0; c < 2; c++) for (int s = 0; s < Ls; s++) m.a[(c * Ls + s) * Ls + s] = 1.0;
  return m;
}
// chi_s = diag_s phi_s + upper_s P-/+ psi_{s+1} + lower_s P+/- psi_{s-1} with phi == psi  (ref: CayleyFermion5Dcache.h:67-77,104-114)
SMat smat_m5d(int Ls, const std::vector<double> &lower, const std::vector<double> &diag, const std::vector<double> &upper, int dag) {
  SMat m; m.Ls = Ls; m.a.assign(2 * Ls * Ls, 0.0);
  for (int s = 0; s < Ls; s++) {
    const int su = (s + 1) % Ls, sd = (s + Ls - 1) % Ls;
    for (int c = 0; c < 2; c++) m.a[(c * Ls + s) * Ls + s] += diag[s];
    // non-dag: upper term carries P- (block 1), lower term carries P+ (block 0); dag swaps
    m.a[((dag ? 0 : 1) * Ls + s) * Ls + su] += upper[s];
    m.a[((dag ? 1 : 0) * Ls + s) * Ls + sd] += lower[s];
  }
  return m;
}
// columns of MooeeInv / MooeeInvDag by running the reference's LDU sweeps on unit vectors (ref: CayleyFermion5Dcache.h:136-170,194-228)
SMat smat_mooee_inv(const CayleyCoeffs &k, int dag) {
  const int Ls = k.Ls;
  SMat m; m.Ls = Ls; m.a.assign(2 * Ls * Ls, 0.0);
  for (int c = 0; c < 2; c++) {
    const bool upperSpin = c == 0;
    const bool typeA = dag ? !upperSpin : upperSpin;
    const std::vector<double> &a = dag ? k.uee : k.lee, &bm = dag ? k.leem : k.ueem, &am = dag ? k.ueem : k.leem, &b = dag ? k.lee : k.uee;
    for (int col = 0; col < Ls; col++) {
      std::vector<double> psi(Ls, 0.0), chi(Ls, 0.0);
      psi[col] = 1.0;
      if (typeA) {
        chi[0] = psi[0];
        for (int s = 1; s < Ls; s++) chi[s] = psi[s] - a[s - 1] * chi[s - 1];
        chi[Ls - 1] /= k.dee[Ls - 1];
        for (int s = Ls - 2; s >= 0; s--) chi[s] = chi[s] / k.dee[s] - bm[s] * chi[Ls - 1];
      } else {
        double acc = 0;
        for (int s = 0; s < Ls - 1; s++) acc += am[s] * psi[s];
        chi[Ls - 1] = (psi[Ls - 1] - acc) / k.dee[Ls - 1];
        for (int s = Ls - 2; s >= 0; s--) chi[s] = psi[s] / k.dee[s] - b[s] * chi[s + 1];
      }
      for (int s = 0; s < Ls; s++) m.a[(c * Ls + s) * Ls + col] = chi[s];
    }
  }
  return m;
}
SMat smat_mul(const SMat &A, const SMat &B) { // A * B (apply B first)
  const int Ls = A.Ls;
  SMat m; m.Ls = Ls; m.a.assign(2 * Ls * Ls, 0.0);
  for (int c = 0; c < 2; c++)
    for (int i = 0; i < Ls; i++) for (int j = 0; j < Ls; j++) {
      double acc = 0;
      for (int l = 0; l < Ls; l++) acc += A.a[(c * Ls + i) * Ls + l] * B.a[(c * Ls + l) * Ls + j];
      m.a[(c * Ls + i) * Ls + j] = acc;
    }
  return m;
}
SMat smat_scale(const SMat &A, double f) { SMat m = A; for (auto &v : m.a) v *= f; return m; }

// device copy in the operator's precision, cached by the operator
const void *smat_device(gb_fermop *op, const SMat &m) {
  const size_t n = m.a.size();
  const size_t bytes = n * (op->prec == GB_F32 ? 4 : 8);
  void *d = nullptr;
  GB_CUDA(cudaMalloc(&d, bytes));
  if (op->prec == GB_F32) {
    std::vector<float> f(n);
    for (size_t i = 0; i < n; i++) f[i] = (float)m.a[i];
    GB_CUDA(cudaMemcpy(d, f.data(), bytes, cudaMemcpyHostToDevice));
  } else {
    GB_CUDA(cudaMemcpy(d, m.a.data(), bytes, cudaMemcpyHostToDevice));
  }
  op->smat_allocs.push_back(d);
  op->smat_host[d] = m;
  return d;
}

template <class T, int LS, int NIN, int EPI> static void smat_launch_k(gb_fermop *op, SMatArgs<T> &a, int nparity, double *d_out) {
  using P = Prec<T>;
  gb_context *ctx = op->ctx;
  constexpr bool CONTIG = (LS % W) == 0;
  const size_t stage = (size_t)NIN * SM_NSITE * (P::NV * LS + 1) * 16;
  const size_t smem = stage * (CONTIG ? 2 : 1);
  const uint32_t ntiles = (a.nsite + SM_NSITE - 1) / SM_NSITE;
  static int per_sm = 0;                    // per instantiation: attribute set and occupancy asked once
  if (per_sm == 0) {
    GB_CUDA(cudaFuncSetAttribute(smat_kernel<T, LS, NIN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smat_kernel<T, LS, NIN, EPI>, SM_NSITE * LS, smem));
    if (per_sm < 1) per_sm = 1;
  }
  uint32_t gx = CONTIG ? std::min<uint32_t>(ntiles, (uint32_t)(ctx->sm_count * per_sm + nparity - 1) / nparity) : ntiles;
  if (gx < 1) gx = 1;
  const size_t nctas = (size_t)gx * nparity;
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    if (op->smat_partials_n < nctas) {
      if (op->smat_partials) cudaFree(op->smat_partials);
      op->smat_partials = nullptr; op->smat_partials_n = 0;
      GB_CUDA(cudaMalloc(&op->smat_partials, nctas * sizeof(double)));
      op->smat_partials_n = nctas;
    }
    a.partials = op->smat_partials;
  }
  dim3 grid(gx, nparity);
  smat_kernel<T, LS, NIN, EPI><<<grid, SM_NSITE * LS, smem, ctx->stream>>>(a);
  count_launch(ctx);
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    smat_reduce_kernel<<<1, 256, 0, ctx->stream>>>(op->smat_partials, (int)nctas, d_out);
    count_launch(ctx);
  }
}
template <class T, int LS> static void smat_launch_ls(gb_fermop *op, SMatArgs<T> &a, int nin, int epi, int nparity, double *d_out) {
  if (epi == EPI_NONE) { if (nin == 2) smat_launch_k<T, LS, 2, EPI_NONE>(op, a, nparity, d_out); else smat_launch_k<T, LS, 1, EPI_NONE>(op, a, nparity, d_out); }
  else if (epi == EPI_NORM) { GB_REQUIRE(nin == 1, "smat: EPI_NORM takes one input"); smat_launch_k<T, LS, 1, EPI_NORM>(op, a, nparity, d_out); }
  else if (epi == EPI_RUPD) { GB_REQUIRE(nin == 2, "smat: EPI_RUPD takes two inputs"); smat_launch_k<T, LS, 2, EPI_RUPD>(op, a, nparity, d_out); }
  else { GB_REQUIRE(nin == 2, "smat: EPI_CGUPD takes two inputs"); smat_launch_k<T, LS, 2, EPI_CGUPD>(op, a, nparity, d_out); }
}
template <class T> static bool smat_launch_T(gb_fermop *op, int Ls, SMatArgs<T> &a, int nin, int epi, int nparity, double *d_out) {
  switch (Ls) {
  case 8: smat_launch_ls<T, 8>(op, a, nin, epi, nparity, d_out); return true;
  case 12: smat_launch_ls<T, 12>(op, a, nin, epi, nparity, d_out); return true;
  case 16: smat_launch_ls<T, 16>(op, a, nin, epi, nparity, d_out); return true;
  default: return false;
  }
}


// ------------------------------------------------------------------ tridiagonal (cyclic) s-space operators with the CG epilogues
// Mooee, MooeeDag, Meooe5D and MeooeDag5D couple s only to s-1 and s+1 (cyclically: the corner carries the mass term), per chirality
// (ref: CayleyFermion5Dcache.h:43-114 M5D / M5Ddag).  For Ls = 16 one 16-lane block of the field layout is one 4D site, so the
// neighbours in s are two warp shuffles away and the fused CG passes become pure streaming kernels (no shared-memory staging, no
// broadcast LDS: the dense form spends 2 x 16 LDS.128 per output vec on matrices that have three non-zero entries per row).
constexpr int STRI_THREADS = 192;      // a multiple of 16 lanes x NV vecs (96 fp32, 192 fp64): (k, s) is fixed per thread
template <class T> struct STriArgs {
  using V = typename Prec<T>::vec;
  const V *x, *y, *z;
  V *out, *psi, *p;
  T dM[2][16], lM[2][16], uM[2][16], dN[2][16], lN[2][16], uN[2][16];   // [chirality][s]: out_s = d x_s + l x_{s-1} + u x_{s+1}
  T alpha;
  int64_t n;                           // vecs in the field (all parity blocks)
  const double *d_c, *d_d, *d_cp;
  double *partials;
};
__device__ __forceinline__ float4 shfl_vec(float4 v, int src) {
  return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}
__device__ __forceinline__ double2 shfl_vec(double2 v, int src) { return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
__device__ __forceinline__ float4 tri3(float d, float4 x, float l, float4 xm, float u, float4 xp) {
  return make_float4(fmaf(u, xp.x, fmaf(l, xm.x, d * x.x)), fmaf(u, xp.y, fmaf(l, xm.y, d * x.y)), fmaf(u, xp.z, fmaf(l, xm.z, d * x.z)), fmaf(u, xp.w, fmaf(l, xm.w, d * x.w)));
}
__device__ __forceinline__ double2 tri3(double d, double2 x, double l, double2 xm, double u, double2 xp) {
  return make_double2(fma(u, xp.x, fma(l, xm.x, d * x.x)), fma(u, xp.y, fma(l, xm.y, d * x.y)));
}

template <class T, int EPI>
__global__ void __launch_bounds__(STRI_THREADS) stri_kernel(const STriArgs<T> a) {
  using V = typename Prec<T>::vec;
  constexpr int NV = Prec<T>::NV;
  const int t = threadIdx.x, s = t & 15, k = (t >> 4) % NV, c = k >= NV / 2 ? 1 : 0;
  const int lane = t & 31, src_m = (lane & 16) | ((s + 15) & 15), src_p = (lane & 16) | ((s + 1) & 15);
  const T dM = a.dM[c][s], lM = a.lM[c][s], uM = a.uM[c][s];
  T dN = 0, lN = 0, uN = 0;
  if (EPI == EPI_RUPD) { dN = a.dN[c][s]; lN = a.lN[c][s]; uN = a.uN[c][s]; }
  T alpha = a.alpha, cg_a = 0, cg_b = 0;
  if (EPI == EPI_RUPD) alpha = (T)(-(*a.d_c) / (*a.d_d));
  if (EPI == EPI_CGUPD) { cg_a = (T)((*a.d_c) / (*a.d_d)); cg_b = (T)((*a.d_cp) / (*a.d_c)); }
  double nrm = 0;
  const int64_t stride = (int64_t)gridDim.x * STRI_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * STRI_THREADS + t; i < a.n; i += stride) {   // whole warps drop out together (n is a multiple of 32)
    V r;
    if (EPI == EPI_CGUPD) {
      const V rv = a.x[i], pv = a.y[i];
      const V pn = vaxpy(cg_b, pv, rv);                         // p = r + b p
      a.psi[i] = vaxpy(cg_a, pv, a.psi[i]);                     // psi += a p
      a.p[i] = pn;
      r = tri3(dM, pn, lM, shfl_vec(pn, src_m), uM, shfl_vec(pn, src_p));
    } else {
      const V xv = a.x[i];
      r = tri3(dM, xv, lM, shfl_vec(xv, src_m), uM, shfl_vec(xv, src_p));
      if (EPI == EPI_RUPD) {
        const V yv = a.y[i];
        r = vadd(r, tri3(dN, yv, lN, shfl_vec(yv, src_m), uN, shfl_vec(yv, src_p)));
        r = vaxpy(alpha, r, a.z[i]);                            // r_new = r_old - (c/d) q
      } else if (a.z != nullptr) r = vaxpy(alpha, a.z[i], r);
    }
    a.out[i] = r;
    if (EPI == EPI_NORM || EPI == EPI_RUPD) nrm += vnorm2(r);
  }
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    __shared__ double red[STRI_THREADS / 32];
    double v = nrm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
      double tot = 0;
      for (int w = 0; w < STRI_THREADS / 32; w++) tot += red[w];
      a.partials[blockIdx.x] = tot;
    }
  }
}

// d / l / u of a cyclic-tridiagonal [2][Ls][Ls] matrix; false if it has any other non-zero entry
static bool tri_extract(const SMat &m, double d[2][16], double l[2][16], double u[2][16]) {
  const int Ls = m.Ls;
  if (Ls != 16) return false;
  for (int c = 0; c < 2; c++)
    for (int i = 0; i < Ls; i++) {
      const int im = (i + Ls - 1) % Ls, ip = (i + 1) % Ls;
      for (int j = 0; j < Ls; j++) {
        const double v = m.a[(c * Ls + i) * Ls + j];
        if (j == i) d[c][i] = v; else if (j == im) l[c][i] = v; else if (j == ip) u[c][i] = v; else if (v != 0.0) return false;
      }
    }
  return true;
}
bool smat_tri_onesided(const gb_fermop *op, const void *dM, float d[2][16], float o[2][16], int dir[2]) {
  auto h = op->smat_host.find(dM);
  if (h == op->smat_host.end()) return false;
  double dd[2][16] = {}, l[2][16] = {}, u[2][16] = {};
  if (!tri_extract(h->second, dd, l, u)) return false;
  for (int c = 0; c < 2; c++) {
    bool has_l = false, has_u = false;
    for (int i = 0; i < 16; i++) { has_l |= l[c][i] != 0.0; has_u |= u[c][i] != 0.0; }
    if (has_l && has_u) return false;
    dir[c] = has_u ? +1 : -1;
    for (int i = 0; i < 16; i++) { d[c][i] = (float)dd[c][i]; o[c][i] = (float)(has_u ? u[c][i] : l[c][i]); }
  }
  return true;
}
double *smat_partials_ensure(gb_fermop *op, size_t n) {
  if (op->smat_partials_n < n) {
    if (op->smat_partials) cudaFree(op->smat_partials);
    op->smat_partials = nullptr; op->smat_partials_n = 0;
    GB_CUDA(cudaMalloc(&op->smat_partials, n * sizeof(double)));
    op->smat_partials_n = n;
  }
  return op->smat_partials;
}
void smat_reduce_partials(gb_fermop *op, size_t n, double *d_out) {
  smat_reduce_kernel<<<1, 256, 0, op->ctx->stream>>>(op->smat_partials, (int)n, d_out);
  count_launch(op->ctx);
}
template <class T, int EPI> static void stri_launch(gb_fermop *op, STriArgs<T> &a, double *d_out) {
  gb_context *ctx = op->ctx;
  const unsigned blocks = (unsigned)std::min<int64_t>((a.n + STRI_THREADS - 1) / STRI_THREADS, (int64_t)ctx->sm_count * 10);
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    if (op->smat_partials_n < blocks) {
      if (op->smat_partials) cudaFree(op->smat_partials);
      op->smat_partials = nullptr; op->smat_partials_n = 0;
      GB_CUDA(cudaMalloc(&op->smat_partials, blocks * sizeof(double)));
      op->smat_partials_n = blocks;
    }
    a.partials = op->smat_partials;
  }
  stri_kernel<T, EPI><<<blocks, STRI_THREADS, 0, ctx->stream>>>(a);
  count_launch(ctx);
  if (EPI == EPI_NORM || EPI == EPI_RUPD) {
    smat_reduce_kernel<<<1, 256, 0, ctx->stream>>>(op->smat_partials, (int)blocks, d_out);
    count_launch(ctx);
  }
  check_launch(ctx, "stri");
}
// the fused CG passes through the streaming kernel; false = not applicable (Ls != 16, a dense matrix, or switched off): use the dense form
template <class T> static bool stri_apply_T(gb_fermop *op, int epi, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y,
                                            double alpha, const gb_fermion *z, gb_fermion *out, const SMatCG *cg) {
  using V = typename Prec<T>::vec;
  auto hm = op->smat_host.find(dM);
  if (hm == op->smat_host.end()) return false;
  double d[2][16] = {}, l[2][16] = {}, u[2][16] = {}, dn[2][16] = {}, ln[2][16] = {}, un[2][16] = {};
  if (!tri_extract(hm->second, d, l, u)) return false;
  if (epi == EPI_RUPD) {
    auto hn = op->smat_host.find(dN);
    if (hn == op->smat_host.end() || !tri_extract(hn->second, dn, ln, un)) return false;
  }
  STriArgs<T> a;
  a.x = (const V *)x->data; a.y = y ? (const V *)y->data : nullptr; a.z = z ? (const V *)z->data : nullptr; a.out = (V *)out->data;
  a.psi = cg && cg->psi ? (V *)cg->psi->data : nullptr; a.p = cg && cg->p ? (V *)cg->p->data : nullptr;
  for (int c = 0; c < 2; c++) for (int i = 0; i < 16; i++) {
    a.dM[c][i] = (T)d[c][i]; a.lM[c][i] = (T)l[c][i]; a.uM[c][i] = (T)u[c][i];
    a.dN[c][i] = (T)dn[c][i]; a.lN[c][i] = (T)ln[c][i]; a.uN[c][i] = (T)un[c][i];
  }
  a.alpha = (T)alpha; a.n = x->nvec();
  a.d_c = cg ? cg->d_c : nullptr; a.d_d = cg ? cg->d_d : nullptr; a.d_cp = cg ? cg->d_cp : nullptr; a.partials = nullptr;
  double *d_out = cg ? cg->d_out : nullptr;
  if (epi == EPI_NORM) stri_launch<T, EPI_NORM>(op, a, d_out);
  else if (epi == EPI_RUPD) stri_launch<T, EPI_RUPD>(op, a, d_out);
  else if (epi == EPI_CGUPD) stri_launch<T, EPI_CGUPD>(op, a, d_out);
  else return false;
  out->cb = x->cb;
  return true;
}
static bool stri_apply(gb_fermop *op, int epi, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, double alpha,
                       const gb_fermion *z, gb_fermion *out, const SMatCG *cg) {
  if (op->Ls != 16 || getenv("GB_NO_STRI") != nullptr) return false;
  fermion_check_same(x, out);
  if (y) fermion_check_same(x, y);
  if (z) fermion_check_same(x, z);
  if (cg && cg->psi) { fermion_check_same(x, cg->psi); fermion_check_same(x, cg->p); }
  return op->prec == GB_F32 ? stri_apply_T<float>(op, epi, dM, x, dN, y, alpha, z, out, cg) : stri_apply_T<double>(op, epi, dM, x, dN, y, alpha, z, out, cg);
}

// out = M x [+ N y] [+ alpha z]; returns false when Ls is outside the instantiated set (caller falls back)
static bool smat_apply_epi(gb_fermop *op, int epi, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, double alpha,
                           const gb_fermion *z, gb_fermion *out, const SMatCG *cg) {
  GB_TRACE("SSpaceDense");
  gb_context *ctx = op->ctx;
  const int Ls = op->Ls;
  if (!(Ls == 8 || Ls == 12 || Ls == 16)) return false;
  fermion_check_same(x, out);
  if (y) fermion_check_same(x, y);
  if (z) fermion_check_same(x, z);
  if (cg && cg->psi) { fermion_check_same(x, cg->psi); fermion_check_same(x, cg->p); }
  // aliasing x/y/z with out is safe: a CTA stages its whole tile in shared memory before it writes, tiles are disjoint
  const int nin = y ? 2 : 1;
  const size_t bstride = (size_t)x->hblk * nv_of(op->prec) * W;
  double *d_out = cg ? cg->d_out : nullptr;
  bool ok;
  if (op->prec == GB_F32) {
    SMatArgs<float> a{(const float4 *)x->data, y ? (const float4 *)y->data : nullptr, z ? (const float4 *)z->data : nullptr, (float4 *)out->data,
                      (const float *)dM, (const float *)dN, (float)alpha, (uint32_t)x->nsite4, bstride};
    if (cg) { a.d_c = cg->d_c; a.d_d = cg->d_d; a.d_cp = cg->d_cp; a.psi = cg->psi ? (float4 *)cg->psi->data : nullptr; a.p = cg->p ? (float4 *)cg->p->data : nullptr; }
    ok = smat_launch_T<float>(op, Ls, a, nin, epi, x->nparity, d_out);
  } else {
    SMatArgs<double> a{(const double2 *)x->data, y ? (const double2 *)y->data : nullptr, z ? (const double2 *)z->data : nullptr, (double2 *)out->data,
                       (const double *)dM, (const double *)dN, alpha, (uint32_t)x->nsite4, bstride};
    if (cg) { a.d_c = cg->d_c; a.d_d = cg->d_d; a.d_cp = cg->d_cp; a.psi = cg->psi ? (double2 *)cg->psi->data : nullptr; a.p = cg->p ? (double2 *)cg->p->data : nullptr; }
    ok = smat_launch_T<double>(op, Ls, a, nin, epi, x->nparity, d_out);
  }
  if (ok) { check_launch(ctx, "smat"); out->cb = x->cb; }
  return ok;
}
bool smat_apply(gb_fermop *op, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, double alpha, const gb_fermion *z,
                gb_fermion *out) {
  return smat_apply_epi(op, EPI_NONE, dM, x, dN, y, alpha, z, out, nullptr);
}
// out = M x + alpha z ; *d_out = |out|^2 on this rank (device)
bool smat_apply_norm(gb_fermop *op, const void *dM, const gb_fermion *x, double alpha, const gb_fermion *z, gb_fermion *out, double *d_out) {
  SMatCG cg; cg.d_out = d_out;
  if (stri_apply(op, EPI_NORM, dM, x, nullptr, nullptr, alpha, z, out, &cg)) return true;
  return smat_apply_epi(op, EPI_NORM, dM, x, nullptr, nullptr, alpha, z, out, &cg);
}
// r = r - (c/d) (M x + N y) ; *d_out = |r|^2 on this rank (device)
bool smat_apply_rupd(gb_fermop *op, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, gb_fermion *r, const double *d_c,
                     const double *d_d, double *d_out) {
  SMatCG cg; cg.d_c = d_c; cg.d_d = d_d; cg.d_out = d_out;
  const int cb = r->cb;
  const bool ok = stri_apply(op, EPI_RUPD, dM, x, dN, y, 0.0, r, r, &cg) || smat_apply_epi(op, EPI_RUPD, dM, x, dN, y, 0.0, r, r, &cg);
  r->cb = cb;
  return ok;
}
// psi += (c/d) p ; p = r + (cp/c) p ; out = M p(new)
bool smat_apply_cgupd(gb_fermop *op, const void *dM, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, gb_fermion *out, const double *d_c,
                      const double *d_d, const double *d_cp) {
  SMatCG cg; cg.d_c = d_c; cg.d_d = d_d; cg.d_cp = d_cp; cg.psi = psi; cg.p = p;
  if (stri_apply(op, EPI_CGUPD, dM, r, dM, p, 0.0, nullptr, out, &cg)) return true;
  return smat_apply_epi(op, EPI_CGUPD, dM, r, dM, p, 0.0, nullptr, out, &cg);
}

} // namespace gb
