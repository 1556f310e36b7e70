// dhop_col2.cuh -- column-sweep fp32 hopping kernel, second generation (round 2): the kernel bench.py times.
//
// Same sweep as dhop_col.cuh (a CTA owns a 4x4 (x/2, y) micro-block and walks z; the other-parity "central column" lives in
// a three-plane ring in shared memory and serves 6 of the 8 legs), with what the round-1 ncu capture asked for:
//   * the ring planes are filled by TMA bulk copies (cp.async.bulk -> UBLKCP), not by 6 per-thread LDGSTS per step: the ring
//     uses the field's own blocked layout ([block of 16 i5][6 vecs][16 lanes]), in which one x-row of the micro-block (4 sites x
//     Ls) is ONE contiguous run of 4*Ls*96 bytes in global memory, so a plane is 4 bulk copies issued by 4 threads.  That takes
//     24-48 of the ~320 L1/LSU wavefronts per warp-step (the pipe that was 67 % busy) and 7 instructions per thread-step away;
//   * full / free mbarrier pairs instead of "everybody's cp.async arrival": one "plane landed" barrier per ring slot (three
//     slots, so a waiter can never miss a phase) and one "z- leg done" barrier that tells the issuing threads that the oldest
//     slot and the oldest link buffer may be overwritten;
//   * decomposed lattices (z and / or t split over ranks) run THIS kernel too: columns sweep only the planes whose z legs are
//     local, CTAs on the t surface are rasterised last, acquire the neighbours' epoch flags and take their off-node t leg from
//     the receive buffer (projected half spinors, written by the neighbour's pack kernel over NVLink) -- a CTA-uniform branch.
//     The two z-surface planes are left to the micro-block semi-fused kernel (dhop_fast.cuh, box launch).
// Arithmetic per leg is dhop_fast.cuh's (packed f32x2 projection / SU(3) multiply / reconstruction).
// ref (what it computes): WilsonKernelsImplementation.h:57-68,112-163 (site), :167-285 (interior / exterior legs).
#pragma once
#include "dhop_col.cuh"

namespace gb {

struct Col2Args {
  const float4 *in[2];
  float4 *out[2];
  const float4 *U[2];
  const float4 *axpy[2];
  float axpy_a, axpy_b;
  int Lxh, Ly, Lz, Lt;
  int z0, N, nzc;               // column c of a (x,y,t) block sweeps planes z0 + c*N ... z0 + c*N + N-1 (mod Lz)
  // 1-D grid = [interior-t CTAs of parity slot 0][... slot 1][surface-t CTAs slot 0][... slot 1]; inside a segment t runs
  // fastest (t neighbours side by side in L2), then the x blocks, then the y blocks (raster 1: y blocks before x blocks), then z chunks
  uint32_t n_int, n_surf;
  int t_int0, nt_int, nt_surf, raster;
  int tb;                       // t block: inside the interior segment only tb consecutive t run fastest (0 = all of them)
  int l2pf;                     // issue an L2 prefetch of ring plane z+2 at the top of step z (cp.async.bulk.prefetch.L2)
  FastDiv dnt_int, dnt_surf, dNxo, dNyo, dtb, dNzc;
  int nparity, first_parity, origin_parity;
  int cta_sync;                 // debugging switch (GB_COL2_SYNC=1): a CTA-wide barrier per step instead of the "z- leg done" mbarrier
  int hints;                    // cache hints (GB_COL_HINTS bit mask): 1 = streaming (evict-first) stores of the result, 2 = evict-first
                                // L2 policy on the link copies (each link is read once per hop), 4 = evict-last on the ring-plane copies
  uint32_t zero;                // always 0, but only the host knows: lets the kernel build register dependencies ptxas cannot fold
  // off-node t legs (MODE 1): receive buffers of the backward (point 7) and forward (point 3) t leg, epoch flags
  int t_comm;
  const float4 *halo_tm, *halo_tp;
  size_t hstride;               // float4 between the parity-0 and parity-1 faces
  // off-node z legs (MODE 1, z_inkernel): the columns sweep ALL planes; the z- leg of plane 0 and the z+ leg of plane Lz-1 come from
  // the receive buffers of points 6 / 2 (written by the neighbours' pack kernels, which precede their hop launches)
  int z_inkernel;
  const float4 *halo_zm, *halo_zp;
  size_t hstride_z;
  const unsigned long long *flags;
  unsigned long long epoch;
  // ---- t faces sent by the hop itself (MODE 1, t decomposed, Lt >= 4): the CTAs of slice t = 1 hold every input spinor of plane 0 in
  //      registers (their backward t neighbour), those of slice Lt-2 every spinor of plane Lt-1: they project them with the RECEIVER's
  //      projector and store the half spinors into the neighbour rank's receive buffer, so the pack kernel never re-reads the t faces.
  //      The last of these sender CTAs publishes the two t epoch flags.  Sender CTAs are interior CTAs (they wait for nobody) and
  //      precede the surface CTAs in the grid, so a surface CTA spinning on a flag never keeps a sender from being scheduled.
  int send_on;
  float4 *send_dst[2];          // [0]: plane 0 -> the backward neighbour's point-3 buffer ; [1]: plane Lt-1 -> the forward neighbour's point-7 buffer
  size_t send_pstride;          // float4 between the parity slots of a face buffer (full-lattice hops)
  unsigned long long *send_flag[2];
  unsigned int *send_counter;
  uint32_t n_senders;               // sender CTAs of the launch
  // ---- optional epilogue (template parameter EPI != 0, Ls = 16, one parity): the s-space pass of the Schur CG that follows this
  //      hop is applied to the result while it is still in registers (fermop.cu: cg_fused_rest).  The operators are the cyclic
  //      bidiagonal-per-chirality ones (Mooee, MooeeDag, MeooeDag5D; ref: CayleyFermion5Dcache.h:43-114): y_s = d_s x_s + o_s x_{s+dir},
  //      dir = -1 or +1 per chirality, the 16 s of a site being the 16 lanes of a half warp.
  //      EPI 1:  w = T_aux(aux) - hop            -> out ; |w|^2 -> partials[cta]      (w = Mpc p, d = |w|^2)
  //      EPI 2:  r += -(c/d) (T_aux(aux) + T_hop(hop))  -> e_r (out is not written) ; |r|^2 -> partials[cta]
  const float4 *e_aux;
  float4 *e_r;
  float e_ad[2][16], e_ao[2][16], e_hd[2][16], e_ho[2][16];
  int e_adir[2], e_hdir[2];
  const double *e_c, *e_d;
  double *e_partials;
};
__device__ __forceinline__ f2 shfl_f2(f2 v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ float norm2_f2(f2 v) { float x, y; upk(v, x, y); return x * x + y * y; }

__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void store_spinor_stream(const SpinorP &f, float4 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { float4 v; upk(f.c[2 * k], v.x, v.y); upk(f.c[2 * k + 1], v.z, v.w); __stcs(p + (k << LOGW), v); }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// float4 offset, inside one ring plane (or one parity block of a field), of vec 0 of vector site i
__device__ __forceinline__ int blk_off(int i) { return (((i >> LOGW) * 6) << LOGW) + (i & (W - 1)); }

// one leg whose source spinor is vec k at p[k * 16] (ring plane or global field: same blocked layout)
template <int DAG, int MU, int FWD>
__device__ __forceinline__ void col2_leg(const float4 *p, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  SpinorP f; HalfP chi, Uchi; LinkS u;
#pragma unroll
  for (int k = 0; k < 6; k++) { const float4 v = p[k << LOGW]; f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
  proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}
// The z- leg, which reads ring slot bm for the last time: its six shared-memory loads must have RETURNED before this thread
// tells the issuers that the slot may be refilled.  mbarrier.arrive orders the loads' issue, not their completion (the SASS had
// the arrive ahead of the FFMA2 that first consumes the last two loads), and a bulk copy into the slot can overtake a load that
// still waits in the shared-memory pipe: measured on the B200 as one stale spinor in ~2 of 1000 launches at 32^4 x 16
// (scripts/hop_stress.py; zero with a CTA-wide barrier instead).  So the arrive's ADDRESS is made to depend on one register of
// each load (an LDS.128 delivers its four registers together): (bits & a.zero) is 0, which the compiler cannot know.
// a z leg whose source is the already projected half spinor of an off-node neighbour (3 vecs in the halo blocking)
template <int DAG, int FWD>
__device__ __forceinline__ void col2_leg_zhalo(const float4 *h, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfP chi, Uchi; LinkS u;
#pragma unroll
  for (int q = 0; q < 3; q++) { const float4 v = h[q << LOGW]; chi.c[2 * q] = pk(v.x, v.y); chi.c[2 * q + 1] = pk(v.z, v.w); }
  lds_link(u, Usm + (FWD ? 2 : 6) * 5);
  mult_p(Uchi, u, chi);
  recon_p<2, SIGN>(res, Uchi);
}
template <int DAG>
__device__ __forceinline__ void col2_leg_zm_arrive(const float4 *p, const float4 *Usm, SpinorP &res, uint64_t *bar, uint32_t zero, bool arrive) {
  constexpr int SIGN = DAG ? -1 : +1;
  SpinorP f; HalfP chi, Uchi; LinkS u;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) { const float4 v = p[k << LOGW]; f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); bits |= __float_as_uint(v.x); }
  if (arrive) mbar_arrive(reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(bar) + (bits & zero)));
  proj_p<2, SIGN>(chi, f);
  lds_link(u, Usm + 6 * 5);
  mult_p(Uchi, u, chi);
  recon_p<2, SIGN>(res, Uchi);
}
// t leg from registers: a full spinor (local neighbour), or the already projected half spinor of an off-node neighbour in c[0..5]
template <int DAG, int MU, int FWD>
__device__ __forceinline__ void col2_leg_reg(const SpinorP &f, bool is_half, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfP chi, Uchi; LinkS u;
  if (is_half) {
#pragma unroll
    for (int q = 0; q < 6; q++) chi.c[q] = f.c[q];
  } else proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}

template <int LS> constexpr size_t col2_smem_bytes() { return (size_t)(3 * COL_NSITE * 6 * LS + 3 * COL_NSITE * FAST_USTRIDE) * 16 + 64; }

// MODE 0: single rank (every leg local, periodic wrap inside the local volume).  MODE 1: decomposed in z and / or t (see above).
template <int LS, int DAG, int MODE, int EPI = 0>
__global__ void __launch_bounds__(COL_NSITE *LS, 2) dhop_col2_kernel(const Col2Args a) {
  static_assert(EPI == 0 || LS == W, "the s-space epilogues need one 4D site per 16-lane block");
  extern __shared__ __align__(128) unsigned char col_smem[];
  constexpr int PLANE = COL_NSITE * 6 * LS;                  // float4 per ring plane, field layout [block][vec k][lane]
  constexpr int UBUF = COL_NSITE * FAST_USTRIDE;
  constexpr int NTHR = COL_NSITE * LS;
  constexpr uint32_t ROW_BYTES = 4u * LS * 6 * 16;           // one x-row of the micro-block: contiguous in the field
  float4 *ring = reinterpret_cast<float4 *>(col_smem);
  float4 *Usm = ring + 3 * PLANE;                            // 3 x [16 sites][41]
  uint64_t *bars = reinterpret_cast<uint64_t *>(Usm + 3 * UBUF);   // [0..2] links, [3..5] ring slot landed, [6] z- leg done
  const int sl = threadIdx.x / LS, s = threadIdx.x % LS;
  const int xl = sl & 3, yl = sl >> 2;
  // ---- which column
  uint32_t b = blockIdx.x;
  int slot = 0;
  bool surf = false;
  {
    const uint32_t tot_int = a.n_int * (uint32_t)a.nparity;
    if (b >= tot_int) { surf = true; b -= tot_int; if (b >= a.n_surf) { b -= a.n_surf; slot = 1; } }
    else if (b >= a.n_int) { b -= a.n_int; slot = 1; }
  }
  const int p = a.first_parity ^ slot;
  uint32_t t, xo, yo, zc;
  uint32_t tg = 0;
  bool sender_seg = false;
  if (!surf) {
    if (MODE == 1 && a.send_on) {
      // hop-sent t faces: t still runs fastest (the t neighbours of a column block stay side by side in the schedule, which is what
      // lets their re-reads hit L2 -- putting the sender slices in a segment of their own cost 0.8 GB of extra DRAM reads per hop at
      // 64.64.32.16 x 16), but inside each column block the two sender slices t = 1 and t = Lt-2 come first
      a.dnt_int.divmod(b, b, t);
      sender_seg = t < 2;
    } else if (a.tb) a.dtb.divmod(b, b, t);
    else a.dnt_int.divmod(b, b, t);
  } else { a.dnt_surf.divmod(b, b, t); t = t == 0 ? a.Lt - 1 : 0; }        // surface segment: t = Lt-1, then t = 0
  if (a.raster == 0) { a.dNxo.divmod(b, b, xo); a.dNyo.divmod(b, zc, yo); }
  else if (a.raster == 1) { a.dNyo.divmod(b, b, yo); a.dNxo.divmod(b, zc, xo); }
  else { a.dNzc.divmod(b, b, zc); a.dNyo.divmod(b, b, yo); a.dNxo.divmod(b, tg, xo); }   // z chunks of a column side by side
  if (!surf) {
    if (a.tb && a.raster != 2) { const uint32_t ntg = (uint32_t)(a.nt_int / a.tb); tg = zc % ntg; zc /= ntg; }   // t groups outside x / y, inside z chunks
    if (MODE == 1 && a.send_on) t = sender_seg ? (t ? (uint32_t)a.Lt - 2 : 1u) : t;      // slices 1, Lt-2, 2, 3, ..., Lt-3
    else t += (a.tb ? tg * a.tb : 0) + a.t_int0;
  }
  const int xh = xo * 4 + xl, y = yo * 4 + yl, zfirst = a.z0 + zc * a.N;
  const float4 *__restrict__ in = a.in[1 - p];
  const uint32_t zstride = (uint32_t)a.Lxh * a.Ly, tstride = zstride * a.Lz;
  const uint32_t site_xyt = xh + a.Lxh * y + tstride * t;
  auto gptr = [&](uint32_t site) { const uint32_t i = site * LS + s; return in + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1)); };
  const int i0 = sl * LS + s;
  float4 *const mine = ring + blk_off(i0);                     // this thread's element in ring slot 0
  // in-block neighbours: slot +-1 along x, +-4 along y (float4 offsets from the own element; compile-time constants for Ls = 16,
  // where a 16-lane block is exactly one 4D site)
  auto noff = [&](int dslot) { return LS == W ? dslot * 6 * W : blk_off(i0 + dslot * LS) - blk_off(i0); };
  const int off_x1 = xl < 3 ? noff(1) : 0, off_xm1 = xl > 0 ? noff(-1) : 0;
  const int off_y1 = yl < 3 ? noff(4) : 0, off_ym1 = yl > 0 ? noff(-4) : 0;

  const bool issuer = s == 0;                                  // 16 threads: each stages the links of its site; 4 of them a ring row
  const bool row_issuer = issuer && xl == 0;
  // source / destination of this thread's ring row (row yl of the block)
  const uint32_t row_site = site_xyt - xl;                      // x-row start (xl == 0 for row issuers)
  auto row_src = [&](int z) { return in + ((size_t)(((row_site + zstride * (uint32_t)z) * LS) >> LOGW) * 6 << LOGW); };
  float4 *const row_dst = ring + blk_off(4 * yl * LS);
  auto wrapz = [&](int z) { return z >= a.Lz ? z - a.Lz : (z < 0 ? z + a.Lz : z); };

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1); mbar_init(&bars[4], 1); mbar_init(&bars[5], 1);
    mbar_init(&bars[6], NTHR);
  }
  const bool tm_halo = MODE == 1 && a.t_comm && t == 0, tp_halo = MODE == 1 && a.t_comm && (int)t == a.Lt - 1;
  // columns that start at plane 0 / end at plane Lz-1 of a z-decomposed lattice take that plane's outward z leg from the receive buffer
  const bool zm_halo = MODE == 1 && a.z_inkernel && zfirst == 0, zp_halo = MODE == 1 && a.z_inkernel && zfirst + a.N == a.Lz;
  if (MODE == 1 && (surf || zm_halo || zp_halo)) {
    // acquire the neighbours' epoch flags (peer-written, system scope) before any thread touches the receive buffers: points 3 / 7
    // (t; set by the neighbours' sender CTAs or pack kernels) in the t-surface CTAs, points 2 / 6 (z; pack kernels) where a z halo is read
    const int pt = threadIdx.x;
    const bool mine_to_wait = (surf && (pt == 3 || pt == 7)) || (zp_halo && pt == 2) || (zm_halo && pt == 6);
    if (mine_to_wait) {
      unsigned long long v;
      do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + pt) : "memory"); } while (v < a.epoch);
    }
  }
  __syncthreads();
  // ---- prologue: planes zfirst-1 (slot 0) and zfirst (slot 1), links of step 0
  const int zfirst_w = wrapz(zfirst);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[3], 4 * ROW_BYTES); mbar_expect_tx(&bars[4], 4 * ROW_BYTES);
    mbar_expect_tx(&bars[0], COL_NSITE * 640);
  }
  if (row_issuer) {
    bulk_g2s(row_dst, row_src(wrapz(zfirst - 1)), ROW_BYTES, &bars[3]);
    bulk_g2s(row_dst + PLANE, row_src(zfirst_w), ROW_BYTES, &bars[4]);
  }
  if (issuer) {
    if (a.hints & 2) bulk_g2s_hint(Usm + sl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * zfirst_w) * 40, 640, &bars[0], l2_policy_evict_first());
    else bulk_g2s(Usm + sl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * zfirst_w) * 40, 640, &bars[0]);
  }

  const uint32_t site_tm = site_xyt + (t == 0 ? tstride * (a.Lt - 1) : 0u - tstride);
  const uint32_t site_tp = site_xyt + ((int)t == a.Lt - 1 ? 0u - tstride * (a.Lt - 1) : tstride);
  // global fall-backs of the in-plane legs that leave the 4x4 block (periodic wrap inside the local volume)
  const uint32_t site_xm = site_xyt - xh + (xh == 0 ? a.Lxh - 1 : xh - 1), site_xp = site_xyt - xh + (xh + 1 == a.Lxh ? 0 : xh + 1);
  const uint32_t site_ym = site_xyt + (y == 0 ? a.Lxh * (a.Ly - 1) : 0u - a.Lxh), site_yp = site_xyt + (y + 1 == a.Ly ? 0u - a.Lxh * (a.Ly - 1) : a.Lxh);
  // off-node t legs: face index = cb index with t removed; the neighbour stored half spinors in the same 16-lane blocking
  const int ip = 1 - p;
  const float4 *hb_tm = nullptr, *hb_tp = nullptr;
  if (MODE == 1) {
    if (tm_halo) hb_tm = a.halo_tm + (size_t)ip * a.hstride;
    if (tp_halo) hb_tp = a.halo_tp + (size_t)ip * a.hstride;
  }
  const bool snd_m = MODE == 1 && a.send_on && t == 1, snd_p = MODE == 1 && a.send_on && (int)t == a.Lt - 2;
  const uint32_t face_xy = xh + a.Lxh * y;
  auto hptr = [&](const float4 *base, int z) { const uint32_t i = (face_xy + zstride * (uint32_t)z) * LS + s; return base + ((size_t)(i >> LOGW) * 3 << LOGW) + (i & (W - 1)); };

  // off-node z legs: face index = cb index with z removed; this thread's half spinor in the point-6 / point-2 receive buffer
  const float4 *hz_m = nullptr, *hz_p = nullptr;
  if (MODE == 1 && (zm_halo || zp_halo)) {
    const uint32_t i = (face_xy + zstride * t) * LS + s;                 // (xh + Lxh (y + Ly t)) Ls + s
    const size_t ho = ((size_t)(i >> LOGW) * 3 << LOGW) + (i & (W - 1)) + (size_t)ip * a.hstride_z;
    if (zm_halo) hz_m = a.halo_zm + ho;
    if (zp_halo) hz_p = a.halo_zp + ho;
  }
  // epilogue constants of this thread's s (both chiralities), the lanes of its s neighbours, the CG scalar, the norm accumulator
  f2 e_ad[2], e_ao[2], e_hd[2], e_ho[2], e_alpha = pk(0.f, 0.f);
  int e_anb[2] = {0, 0}, e_hnb[2] = {0, 0};
  double e_nrm = 0;
  if (EPI != 0) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      e_ad[c] = pk(a.e_ad[c][s], a.e_ad[c][s]); e_ao[c] = pk(a.e_ao[c][s], a.e_ao[c][s]);
      e_anb[c] = (lane & 16) | ((s + a.e_adir[c]) & 15);
      if (EPI == 2) {
        e_hd[c] = pk(a.e_hd[c][s], a.e_hd[c][s]); e_ho[c] = pk(a.e_ho[c][s], a.e_ho[c][s]);
        e_hnb[c] = (lane & 16) | ((s + a.e_hdir[c]) & 15);
      }
    }
    if (EPI == 2) { const float al = (float)(-(*a.e_c) / (*a.e_d)); e_alpha = pk(al, al); }
  }
  mbar_wait(&bars[3], 0);
  mbar_wait(&bars[4], 0);
  int ub = 0, bm = 0;                                          // link buffer of this step; ring slot of plane z-1
#pragma unroll 1
  for (int k = 0; k < a.N; k++) {
    const int z = wrapz(zfirst + k);
    const int b0 = bm == 2 ? 0 : bm + 1, bp = b0 == 2 ? 0 : b0 + 1;   // ring slots of planes z, z+1
    const int un = ub == 2 ? 0 : ub + 1;
    const uint32_t zoff = zstride * z;
    const int zp = z + 1 == a.Lz ? 0 : z + 1;
    // Every thread observes the completion of every phase of bars[6] (an arrive-on of the next phase by a thread that has not
    // seen the previous one complete is undefined -- it faulted on the B200), which also bounds the run-ahead of any warp to
    // one step, as the everybody-arrives barrier of the round-1 kernel did.
    if (a.cta_sync) __syncthreads();
    else if (k > 0) mbar_wait(&bars[6], (uint32_t)(k - 1) & 1);
    if (issuer) {
      // ---- asynchronous: plane z+1 into ring slot bp, the links of the next step.  Slot bp held plane z-2 (read as the z- leg
      //      of step k-1) and link buffer un the links of step k-2: free once everybody arrived in step k-1.
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[3 + bp], 4 * ROW_BYTES);
        if (k + 1 < a.N) mbar_expect_tx(&bars[un], COL_NSITE * 640);
      }
      if (row_issuer) {
        if (a.hints & 4) bulk_g2s_hint(row_dst + bp * PLANE, row_src(zp), ROW_BYTES, &bars[3 + bp], l2_policy_evict_last());
        else bulk_g2s(row_dst + bp * PLANE, row_src(zp), ROW_BYTES, &bars[3 + bp]);
        if (a.l2pf && k + 1 < a.N) {
          const int zpp = zp + 1 == a.Lz ? 0 : zp + 1;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row_src(zpp)), "r"(ROW_BYTES) : "memory");
        }
      }
      if (k + 1 < a.N) {
        if (a.hints & 2) bulk_g2s_hint(Usm + un * UBUF + sl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * zp) * 40, 640, &bars[un], l2_policy_evict_first());
        else bulk_g2s(Usm + un * UBUF + sl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * zp) * 40, 640, &bars[un]);
      }
    }
    // ---- t neighbours into registers now, used after the shared-memory legs
    SpinorP ftm, ftp;
    if (MODE == 1 && tm_halo) {
      const float4 *h = hptr(hb_tm, z);
#pragma unroll
      for (int q = 0; q < 3; q++) { const float4 v = h[q << LOGW]; ftm.c[2 * q] = pk(v.x, v.y); ftm.c[2 * q + 1] = pk(v.z, v.w); }
    } else load_spinor_p(ftm, gptr(site_tm + zoff));
    if (MODE == 1 && tp_halo) {
      const float4 *h = hptr(hb_tp, z);
#pragma unroll
      for (int q = 0; q < 3; q++) { const float4 v = h[q << LOGW]; ftp.c[2 * q] = pk(v.x, v.y); ftp.c[2 * q + 1] = pk(v.z, v.w); }
    } else load_spinor_p(ftp, gptr(site_tp + zoff));
    if (MODE == 1 && (snd_m || snd_p)) {
      // the neighbour rank's t leg of this face site: forward leg (projector sign of FWD = 1) for plane 0, backward leg for plane Lt-1
      const uint32_t i = (face_xy + zstride * (uint32_t)z) * LS + s;
      const size_t ho = ((size_t)(i >> LOGW) * 3 << LOGW) + (i & (W - 1)) + (a.nparity == 1 ? 0 : (size_t)ip * a.send_pstride);
      if (snd_m) {
        HalfP h; proj_p<3, (DAG ? +1 : -1)>(h, ftm);
        float4 *d = a.send_dst[0] + ho;
#pragma unroll
        for (int q = 0; q < 3; q++) { float4 v; upk(h.c[2 * q], v.x, v.y); upk(h.c[2 * q + 1], v.z, v.w); d[q << LOGW] = v; }
      }
      if (snd_p) {
        HalfP h; proj_p<3, (DAG ? -1 : +1)>(h, ftp);
        float4 *d = a.send_dst[1] + ho;
#pragma unroll
        for (int q = 0; q < 3; q++) { float4 v; upk(h.c[2 * q], v.x, v.y); upk(h.c[2 * q + 1], v.z, v.w); d[q << LOGW] = v; }
      }
    }
    const int pb = (p + a.origin_parity + y + z + (int)t) & 1;
    SpinorP res;
#pragma unroll
    for (int q = 0; q < 12; q++) res.c[q] = pk(0.f, 0.f);
    mbar_wait(&bars[ub], (uint32_t)(k / 3) & 1);
    const float4 *Us = Usm + ub * UBUF + sl * FAST_USTRIDE;
    const float4 *cur = mine + b0 * PLANE;                       // own element, plane z
    // ---- z- : own element of plane z-1; then tell the issuers that this thread is done with that slot
    if (MODE == 1 && zm_halo && k == 0) {                          // (CTA-uniform) plane 0 of a z-decomposed lattice: the halo leg; nothing
      col2_leg_zhalo<DAG, 0>(hz_m, Us, res);                       // of ring slot bm is read, so the arrive needs no dependency
      if (!a.cta_sync) mbar_arrive(&bars[6]);
    } else col2_leg_zm_arrive<DAG>(mine + bm * PLANE, Us, res, &bars[6], a.zero, !a.cta_sync);
    // ---- x legs: the neighbour with the same x/2 index is this thread's own ring element; the other one is the adjacent
    //      slot or, at the block edge, a global load
    if (pb) {
      col2_leg<DAG, 0, 0>(cur, Us, res);
      col2_leg<DAG, 0, 1>(xl < 3 ? cur + off_x1 : gptr(site_xp + zoff), Us, res);
    } else {
      col2_leg<DAG, 0, 0>(xl > 0 ? cur + off_xm1 : gptr(site_xm + zoff), Us, res);
      col2_leg<DAG, 0, 1>(cur, Us, res);
    }
    // ---- y legs: slots +-4 inside the block
    col2_leg<DAG, 1, 0>(yl > 0 ? cur + off_ym1 : gptr(site_ym + zoff), Us, res);
    col2_leg<DAG, 1, 1>(yl < 3 ? cur + off_y1 : gptr(site_yp + zoff), Us, res);
    // ---- t legs from registers
    col2_leg_reg<DAG, 3, 0>(ftm, MODE == 1 && tm_halo, Us, res);
    col2_leg_reg<DAG, 3, 1>(ftp, MODE == 1 && tp_halo, Us, res);
    // ---- z+ : wait for plane z+1 (bulk copies issued at the top of this step), read the own element
    mbar_wait(&bars[3 + bp], (uint32_t)((k + 2) / 3) & 1);        // (observed in every step, also when the plane is not used)
    if (MODE == 1 && zp_halo && k == a.N - 1) col2_leg_zhalo<DAG, 1>(hz_p, Us, res);
    else col2_leg<DAG, 2, 1>(mine + bp * PLANE, Us, res);
    // ---- epilogue
    const uint32_t i = (site_xyt + zoff) * LS + s;
    const size_t offs = ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1));
    if (EPI != 0) {
      SpinorP aux;
      load_spinor_p(aux, a.e_aux + offs);
      float n2 = 0.f;
      if (EPI == 1) {
#pragma unroll
        for (int q = 0; q < 12; q++) {
          const int c = q >= 6 ? 1 : 0;
          const f2 y = fma2(e_ao[c], shfl_f2(aux.c[q], e_anb[c]), mul2(e_ad[c], aux.c[q]));
          res.c[q] = fma2(pk(-1.f, -1.f), res.c[q], y);                   // w = T(aux) - hop
          n2 += norm2_f2(res.c[q]);
        }
        store_spinor_p(res, a.out[p] + offs);
      } else {
        SpinorP r;
        load_spinor_rw(r, a.e_r + offs);
#pragma unroll
        for (int q = 0; q < 12; q++) {
          const int c = q >= 6 ? 1 : 0;
          f2 y = fma2(e_ao[c], shfl_f2(aux.c[q], e_anb[c]), mul2(e_ad[c], aux.c[q]));
          y = fma2(e_hd[c], res.c[q], y);
          y = fma2(e_ho[c], shfl_f2(res.c[q], e_hnb[c]), y);               // q = T_aux(aux) + T_hop(hop)
          r.c[q] = fma2(e_alpha, y, r.c[q]);                               // r -= (c/d) q
          n2 += norm2_f2(r.c[q]);
        }
        store_spinor_p(r, a.e_r + offs);
      }
      e_nrm += (double)n2;
    } else {
      if (a.axpy[p] != nullptr) {
        SpinorP ax;
        load_spinor_p(ax, a.axpy[p] + offs);
        const f2 sa = pk(a.axpy_a, a.axpy_a), sb = pk(a.axpy_b, a.axpy_b);
#pragma unroll
        for (int q = 0; q < 12; q++) res.c[q] = fma2(sa, res.c[q], mul2(sb, ax.c[q]));
      }
      if (a.hints & 1) store_spinor_stream(res, a.out[p] + offs);
      else store_spinor_p(res, a.out[p] + offs);
    }
    bm = b0; ub = un;
  }
  if (MODE == 1 && (snd_m || snd_p)) {
    // publish: fence this CTA's peer stores; the last sender CTA of the launch writes the t epoch flags into the neighbours' memory
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(a.send_counter, 1u);
      if (prev == a.n_senders - 1) {
        *a.send_counter = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.send_flag[0]), "l"(a.epoch) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.send_flag[1]), "l"(a.epoch) : "memory");
      }
    }
  }
  if (EPI != 0) {
    // one partial per CTA (fixed-shape tree); the caller's second stage adds the partials in a fixed order
    __shared__ double e_red[NTHR / 32];
    double v = e_nrm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) e_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t2 = 0;
      for (int w = 0; w < NTHR / 32; w++) t2 += e_red[w];
      a.e_partials[blockIdx.x] = t2;
    }
  }
}

} // namespace gb
