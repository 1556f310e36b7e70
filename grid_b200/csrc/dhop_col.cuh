// dhop_col.cuh -- column-sweep fp32 hopping kernel for sm_100a: the kernel bench.py times since round 1c.
//
// Why: the micro-block kernel (dhop_fast.cuh) is bound by L2->SM throughput, not by HBM.  ncu: 10.1 GB of L2->L1 sectors
// + 1.6 GB of stores per full 32^4 x 16 hop in 0.97 ms = 12 TB/s, which IS the chip's L2 slice cap (~6300 B/clk,
// B300_MICROARCH.md "LTS throughput cap"), because every input spinor is fetched 6.1 times from L2 (8 legs, 24 % L1 hits).
// The only way below that ceiling is to fetch each neighbour fewer times, i.e. reuse inside the SM.
//
// How: a CTA owns a 4x4 (x/2, y) micro-block and WALKS a column of N z-planes.  The other-parity spinors that sit on the
// same (x/2, y) indices ("the central column") serve, for the same thread, as the z+ neighbour at step k, as one of the two
// x neighbours at step k+1 and as the z- neighbour at step k+2; for the neighbouring threads they are the in-block x and y
// neighbours at step k+1.  So each thread loads ONE new column element per step from global memory, parks it in a
// three-plane ring in shared memory (24 KB per plane at Ls = 16), and 6 of the 8 legs read shared memory.  Only the t legs
// and the legs that leave the 4x4 block touch L2: 3.75 fetches per site instead of 6.1 (496 instead of 722 L2 bytes/site).
// Links are staged per step by TMA bulk copies into a double buffer, one step ahead.
// Arithmetic per leg is dhop_fast.cuh's (packed f32x2 projection / SU(3) multiply / reconstruction).
#pragma once
#include "dhop_fast.cuh"

namespace gb {

constexpr int COL_NSITE = 16;

struct ColArgs {
  const float4 *in[2];
  float4 *out[2];
  const float4 *U[2];
  const float4 *axpy[2];
  float axpy_a, axpy_b;
  int comm_dim_mask;            // interior pass of a decomposed lattice: z/t legs leaving the local volume are skipped
  int Lxh, Ly, Lz, Lt;
  int N;                        // z-planes per column (divides Lz)
  FastDiv dLt, dNxo, dNyo;      // CTA index -> (t fastest, x block, y block, z chunk): t neighbours run side by side in L2
  int first_parity, origin_parity;
};

// one leg whose source spinor is reachable through a generic pointer (shared ring or global field), vec k at p[k*stride]
template <int DAG, int MU, int FWD>
__device__ __forceinline__ void col_leg(const float4 *p, int stride, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  SpinorP f; HalfP chi, Uchi; LinkS u;
#pragma unroll
  for (int k = 0; k < 6; k++) { const float4 v = p[k * stride]; f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
  proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}
template <int DAG, int MU, int FWD>
__device__ __forceinline__ void col_leg_reg(const SpinorP &f, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfP chi, Uchi; LinkS u;
  proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}

// Ampere-style asynchronous copy global -> shared (LDGSTS), 16 bytes, bypassing registers and L1
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// the mbarrier receives this thread's arrival when all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <int LS, int NT = 1> constexpr size_t col_smem_bytes() { return (size_t)NT * (3 * COL_NSITE * 6 * LS + 3 * COL_NSITE * FAST_USTRIDE) * 16 + 64; }

// Synchronisation without a CTA-wide barrier per step:
//  * plane z+1 of the ring is filled by cp.async issued at the TOP of step k; each thread's arrival on the plane mbarrier
//    fires when its copy lands, and every thread waits for the phase only just before its own z+ leg at the END of the step.
//    Nobody waits for another thread's arithmetic, only for copies issued a whole step earlier.
//  * passing that wait proves every thread has begun step k, i.e. finished step k-1: the ring slot overwritten next
//    (plane z-1, read by other threads during step k-1) and the link buffer of step k-1 (3 buffers) are free.
// NT = 2: the CTA (512 threads at Ls = 16, one per SM) owns the same micro-block on TWO adjacent t-slices; the t leg between
// them reads the partner slice's ring slot, so only one t leg per site is a global load (2.75 L2 fetches per site).
template <int LS, int DAG, int INTERIOR, int NT = 1>
__global__ void __launch_bounds__(NT *COL_NSITE *LS, NT == 1 ? 2 : 1) dhop_col_kernel(const ColArgs a) {
  extern __shared__ __align__(16) unsigned char col_smem[];
  constexpr int TSLOT = COL_NSITE * 6 * LS;                 // float4 per t-slice of a ring plane
  constexpr int PLANE = NT * TSLOT;                         // float4 per plane of the ring: [t-slice][slot][vec k][s]
  constexpr int UBUF = NT * COL_NSITE * FAST_USTRIDE;
  float4 *ring = reinterpret_cast<float4 *>(col_smem);
  float4 *Usm = ring + 3 * PLANE;                            // 3 x [16 sites][41]
  uint64_t *bars = reinterpret_cast<uint64_t *>(Usm + 3 * UBUF);   // [0..2] links, [3] ring plane
  const int tt = NT == 1 ? 0 : threadIdx.x / (COL_NSITE * LS);
  const int tid1 = NT == 1 ? threadIdx.x : threadIdx.x % (COL_NSITE * LS);
  const int sl = tid1 / LS, s = tid1 % LS;
  const int xl = sl & 3, yl = sl >> 2;
  const int p = a.first_parity ^ (int)blockIdx.y;
  uint32_t b = blockIdx.x, t, xo, yo, zc;
  a.dLt.divmod(b, b, t); a.dNxo.divmod(b, b, xo); a.dNyo.divmod(b, zc, yo);   // dLt divides by Lt / NT
  t = t * NT + tt;
  const int xh = xo * 4 + xl, y = yo * 4 + yl, z0 = zc * a.N;
  const float4 *__restrict__ in = a.in[1 - p];
  const uint32_t zstride = (uint32_t)a.Lxh * a.Ly, tstride = zstride * a.Lz;
  const uint32_t site_xyt = xh + a.Lxh * y + tstride * t;
  auto gptr = [&](uint32_t site) { const uint32_t i = site * LS + s; return in + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1)); };
  float4 *const mine = ring + (tt * COL_NSITE + sl) * 6 * LS + s;   // this thread's slot in plane buffer 0

  // ---- prologue: planes z0-1 and z0 of the central column, links of step 0
  {
    const int zm = z0 == 0 ? a.Lz - 1 : z0 - 1;
    const float4 *g0 = gptr(site_xyt + zstride * zm), *g1 = gptr(site_xyt + zstride * z0);
#pragma unroll
    for (int k = 0; k < 6; k++) { mine[k * LS] = __ldg(g0 + (k << LOGW)); mine[PLANE + k * LS] = __ldg(g1 + (k << LOGW)); }
  }
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], NT * COL_NSITE * LS); }
  __syncthreads();
  const int usl = tt * COL_NSITE + sl;                        // this site's slot in a link buffer
  if (threadIdx.x == 0) mbar_expect_tx(&bars[0], NT * COL_NSITE * 640);
  if (s == 0) bulk_g2s(Usm + usl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * z0) * 40, 640, &bars[0]);

  const bool skip_tm = INTERIOR && ((a.comm_dim_mask >> 3) & 1) && t == 0;
  const bool skip_tp = INTERIOR && ((a.comm_dim_mask >> 3) & 1) && (int)t == a.Lt - 1;
  const uint32_t site_tm = site_xyt + (t == 0 ? tstride * (a.Lt - 1) : 0u - tstride);
  const uint32_t site_tp = site_xyt + ((int)t == a.Lt - 1 ? 0u - tstride * (a.Lt - 1) : tstride);
  // global fall-backs of the in-plane legs that leave the 4x4 block (periodic wrap inside the local volume)
  const uint32_t site_xm = site_xyt - xh + (xh == 0 ? a.Lxh - 1 : xh - 1), site_xp = site_xyt - xh + (xh + 1 == a.Lxh ? 0 : xh + 1);
  const uint32_t site_ym = site_xyt + (y == 0 ? a.Lxh * (a.Ly - 1) : 0u - a.Lxh), site_yp = site_xyt + (y + 1 == a.Ly ? 0u - a.Lxh * (a.Ly - 1) : a.Lxh);

  int ub = 0, bm = 0;                                          // link buffer of this step; ring slot of plane z-1
#pragma unroll 1
  for (int k = 0; k < a.N; k++) {
    const int z = z0 + k;
    const int b0 = bm == 2 ? 0 : bm + 1, bp = b0 == 2 ? 0 : b0 + 1;   // ring slots of planes z, z+1
    const int un = ub == 2 ? 0 : ub + 1;
    const uint32_t zoff = zstride * z;
    const int zp = z + 1 == a.Lz ? 0 : z + 1;
    // ---- asynchronous: the new column element (plane z+1) into the ring, the links of the next step
    {
      const float4 *g = gptr(site_xyt + zstride * zp);
      float4 *dst = mine + bp * PLANE;
#pragma unroll
      for (int q = 0; q < 6; q++) cp_async16(dst + q * LS, g + (q << LOGW));
      cp_async_arrive(&bars[3]);
    }
    if (k + 1 < a.N) {
      if (threadIdx.x == 0) mbar_expect_tx(&bars[un], NT * COL_NSITE * 640);
      if (s == 0) bulk_g2s(Usm + un * UBUF + usl * FAST_USTRIDE, a.U[p] + (size_t)(site_xyt + zstride * (z + 1)) * 40, 640, &bars[un]);
    }
    // ---- t neighbours into registers now, used after the shared-memory legs
    SpinorP ftm, ftp;
    if (NT == 1) {
      if (!skip_tm) load_spinor_p(ftm, gptr(site_tm + zoff));
      if (!skip_tp) load_spinor_p(ftp, gptr(site_tp + zoff));
    } else {                                                     // only the leg that leaves the pair of slices is global
      if (tt == 0) { if (!skip_tm) load_spinor_p(ftm, gptr(site_tm + zoff)); }
      else if (!skip_tp) load_spinor_p(ftm, gptr(site_tp + zoff));
    }
    const int pb = (p + a.origin_parity + y + z + (int)t) & 1;
    SpinorP res;
#pragma unroll
    for (int q = 0; q < 12; q++) res.c[q] = pk(0.f, 0.f);
    mbar_wait(&bars[ub], (uint32_t)(k / 3) & 1);
    const float4 *Us = Usm + ub * UBUF + usl * FAST_USTRIDE;
    const float4 *cur = mine + b0 * PLANE;                       // own slot, plane z
    // ---- z- : own slot of plane z-1
    if (!(INTERIOR && ((a.comm_dim_mask >> 2) & 1) && z == 0)) col_leg<DAG, 2, 0>(mine + bm * PLANE, LS, Us, res);
    // ---- x legs: the neighbour with the same x/2 index is this thread's own ring slot; the other one is the adjacent slot
    //      or, at the block edge, a global load
    if (pb) {
      col_leg<DAG, 0, 0>(cur, LS, Us, res);
      const bool inside = xl < 3;
      col_leg<DAG, 0, 1>(inside ? cur + 6 * LS : gptr(site_xp + zoff), inside ? LS : W, Us, res);
    } else {
      const bool inside = xl > 0;
      col_leg<DAG, 0, 0>(inside ? cur - 6 * LS : gptr(site_xm + zoff), inside ? LS : W, Us, res);
      col_leg<DAG, 0, 1>(cur, LS, Us, res);
    }
    // ---- y legs: slots +-4 inside the block
    col_leg<DAG, 1, 0>(yl > 0 ? cur - 4 * 6 * LS : gptr(site_ym + zoff), yl > 0 ? LS : W, Us, res);
    col_leg<DAG, 1, 1>(yl < 3 ? cur + 4 * 6 * LS : gptr(site_yp + zoff), yl < 3 ? LS : W, Us, res);
    // ---- t legs from registers
    if (NT == 1) {
      if (!skip_tm) col_leg_reg<DAG, 3, 0>(ftm, Us, res);
      if (!skip_tp) col_leg_reg<DAG, 3, 1>(ftp, Us, res);
    } else if (tt == 0) {                                        // t- global, t+ = the partner slice's slot of plane z
      if (!skip_tm) col_leg_reg<DAG, 3, 0>(ftm, Us, res);
      col_leg<DAG, 3, 1>(cur + TSLOT, LS, Us, res);
    } else {
      col_leg<DAG, 3, 0>(cur - TSLOT, LS, Us, res);
      if (!skip_tp) col_leg_reg<DAG, 3, 1>(ftm, Us, res);
    }
    // ---- z+ : wait for plane z+1 (every thread's copy, issued at the top of this step), read the own slot
    mbar_wait(&bars[3], (uint32_t)k & 1);
    if (!(INTERIOR && ((a.comm_dim_mask >> 2) & 1) && z == a.Lz - 1)) col_leg<DAG, 2, 1>(mine + bp * PLANE, LS, Us, res);
    // ---- epilogue
    const uint32_t i = (site_xyt + zoff) * LS + s;
    const size_t offs = ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1));
    if (a.axpy[p] != nullptr) {
      SpinorP ax;
      load_spinor_p(ax, a.axpy[p] + offs);
      const f2 sa = pk(a.axpy_a, a.axpy_a), sb = pk(a.axpy_b, a.axpy_b);
#pragma unroll
      for (int q = 0; q < 12; q++) res.c[q] = fma2(sa, res.c[q], mul2(sb, ax.c[q]));
    }
    store_spinor_p(res, a.out[p] + offs);
    bm = b0; ub = un;
  }
}

} // namespace gb
