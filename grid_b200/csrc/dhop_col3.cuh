// dhop_col3.cuh -- column-sweep fp32 hopping kernel, third generation (round 2).
//
// What the round-2 ncu capture of dhop_col2 said (profiles/r2_col2_*): 35 % of the warp samples are long-scoreboard stalls and
// only 16 warps live on an SM (128 registers x 256 threads x 2 CTAs beside 103 KB of shared memory), so the kernel is bound by
// latency it cannot hide, not by a pipe: (a) 14 % sit in mbarrier waits -- the ring plane z+1 is requested at the top of the step
// that needs it at its end, by compute threads that must first see every thread's "z- leg done" arrival, so a fast warp waits
// for the slowest warp of the CTA to come round and issue; (b) 12 % wait for the t neighbours, (c) 8 % for the legs that leave
// the 4x4 block.  This kernel keeps the sweep (a CTA owns a 4x4 (x/2, y) micro-block and walks z, the other-parity "central
// column" sits in a three-slot ring filled by TMA bulk copies) and changes who issues what, when:
//   * no thread ever waits before issuing a copy.  Each warp counts itself done with the oldest ring plane of a step on a
//     shared-memory counter; the warp whose count completes the step -- whichever it is -- issues the bulk copies of the plane
//     that replaces it and of the links two steps ahead.  Copies are in flight for almost two full steps instead of one, there
//     is no "z- leg done" barrier any more, and the only waits left are on data.
//   * the own-column element is read from the ring ONCE: the spinor loaded for the z+ leg of step k stays in registers over the
//     loop edge and serves the same thread's x leg of step k+1; its z- projection (12 registers) is carried into step k+2
//     (CARRY = 1), which frees the ring slot a step earlier (the ring holds planes z, z+1, z+2) and takes the shared-memory
//     spinor loads per thread-step from 36 to 24.  CARRY = 0 keeps the z- leg on the ring (planes z-1, z, z+1; 30 loads).
//   * global legs are requested one leg ahead of their use: t- at the top of the step, t+ when t- has been consumed, so that one
//     spinor, not two, is parked in registers.
// Decomposed lattices (MODE 1) as in dhop_col2.cuh.  Arithmetic per leg is dhop_fast.cuh's.
// ref (what it computes): WilsonKernelsImplementation.h:57-68,112-163 (site), :167-285 (interior / exterior legs).
#pragma once
#include "dhop_col2.cuh"
#include <type_traits>

namespace gb {

template <int MU, int SIGN> __device__ __forceinline__ void recon_init_p(SpinorP &r, const HalfP &h) {
#pragma unroll
  for (int q = 0; q < 12; q++) r.c[q] = pk(0.f, 0.f);
  recon_p<MU, SIGN>(r, h);
}
__device__ __forceinline__ void load_ring_p(SpinorP &f, const float4 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { const float4 v = p[k << LOGW]; f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
}
template <int DAG, int MU, int FWD> __device__ __forceinline__ void col3_leg(const SpinorP &f, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfP chi, Uchi; LinkS u;
  proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}
template <int DAG, int MU, int FWD> __device__ __forceinline__ void col3_leg_half(const HalfP &chi, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfP Uchi; LinkS u;
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}

// shared memory: ring[3][PLANE], links[3][16 sites][41], 6 mbarriers (links 0..2, ring slots 3..5), 3 step counters
template <int LS> constexpr size_t col3_smem_bytes() { return (size_t)(3 * COL_NSITE * 6 * LS + 3 * COL_NSITE * FAST_USTRIDE) * 16 + 64; }

template <int LS, int DAG, int MODE, int CARRY>
__global__ void __launch_bounds__(COL_NSITE *LS, 2) dhop_col3_kernel(const Col2Args a) {
  extern __shared__ __align__(128) unsigned char col_smem[];
  constexpr int PLANE = COL_NSITE * 6 * LS;                  // float4 per ring plane, field layout [block][vec k][lane]
  constexpr int UBUF = COL_NSITE * FAST_USTRIDE;
  constexpr int NTHR = COL_NSITE * LS;
  constexpr unsigned NWARP = NTHR / 32;
  constexpr uint32_t ROW_BYTES = 4u * LS * 6 * 16;           // one x-row of the micro-block: contiguous in the field
  float4 *ring = reinterpret_cast<float4 *>(col_smem);
  float4 *Usm = ring + 3 * PLANE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(Usm + 3 * UBUF);   // [0..2] links of step k % 3, [3..5] ring slot landed
  unsigned int *cnt = reinterpret_cast<unsigned int *>(bars + 6);  // [0..2] warps done with the oldest plane of step k % 3
  const int sl = threadIdx.x / LS, s = threadIdx.x % LS;
  const int xl = sl & 3, yl = sl >> 2;
  const int lane = threadIdx.x & 31;
  // ---- which column (as in dhop_col2_kernel)
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
  if (!surf) { a.dnt_int.divmod(b, b, t); t += a.t_int0; }
  else { a.dnt_surf.divmod(b, b, t); t = t == 0 ? a.Lt - 1 : 0; }        // surface segment: t = Lt-1, then t = 0
  if (a.raster == 0) { a.dNxo.divmod(b, b, xo); a.dNyo.divmod(b, zc, yo); }
  else { a.dNyo.divmod(b, b, yo); a.dNxo.divmod(b, zc, xo); }
  const int xh = xo * 4 + xl, y = yo * 4 + yl, zfirst = a.z0 + zc * a.N;
  const float4 *__restrict__ in = a.in[1 - p];
  const uint32_t zstride = (uint32_t)a.Lxh * a.Ly, tstride = zstride * a.Lz;
  const uint32_t site_blk = xo * 4 + a.Lxh * (yo * 4) + tstride * t;      // first site of the micro-block at z = 0
  const uint32_t site_xyt = site_blk + xl + a.Lxh * yl;
  auto goff = [&](uint32_t site) { const uint32_t i = site * LS + s; return (((i >> LOGW) * 6) << LOGW) + (i & (W - 1)); };   // float4 offset in a parity block
  const int i0 = sl * LS + s;
  const float4 *const mine = ring + blk_off(i0);               // this thread's element in ring slot 0
  auto noff = [&](int dslot) { return LS == W ? dslot * 6 * W : blk_off(i0 + dslot * LS) - blk_off(i0); };
  const int off_x1 = xl < 3 ? noff(1) : 0, off_xm1 = xl > 0 ? noff(-1) : 0;
  const int off_y1 = yl < 3 ? noff(4) : 0, off_ym1 = yl > 0 ? noff(-4) : 0;
  auto wrapz = [&](int z) { return z >= a.Lz ? z - a.Lz : (z < 0 ? z + a.Lz : z); };
  const uint32_t dz = zstride * LS * 6;                        // float4 per z step in the blocked layout (zstride % 16 == 0)

  // ring plane j (j = 0 ...) is plane zfirst - (1 - CARRY) + j and lives in slot j % 3; step k reads planes [k, k + 2 - CARRY]
  constexpr int JOFF = 1 - CARRY;
  const int jmax = a.N + JOFF;                                 // last ring plane this column reads
  // the copies of ring plane j and of the links of step ks, issued by one converged warp (any)
  auto issue = [&](int j, int ks) {
    if (lane == 0) {
      if (j <= jmax) mbar_expect_tx(&bars[3 + j % 3], 4 * ROW_BYTES);
      if (ks < a.N) mbar_expect_tx(&bars[ks % 3], COL_NSITE * 640);
    }
    __syncwarp();
    if (lane < 16) {
      if (ks < a.N) {
        const uint32_t site = site_blk + (lane & 3) + a.Lxh * (lane >> 2) + zstride * (uint32_t)wrapz(zfirst + ks);
        bulk_g2s(Usm + (ks % 3) * UBUF + lane * FAST_USTRIDE, a.U[p] + (size_t)site * 40, 640, &bars[ks % 3]);
      }
    } else if (lane < 20) {
      if (j <= jmax) {
        const int r = lane - 16;
        const uint32_t row_site = site_blk + a.Lxh * r + zstride * (uint32_t)wrapz(zfirst - JOFF + j);
        bulk_g2s(ring + (j % 3) * PLANE + blk_off(4 * r * LS), in + ((size_t)((row_site * LS) >> LOGW) * 6 << LOGW), ROW_BYTES, &bars[3 + j % 3]);
      }
    } else if (a.l2pf && lane < 28 && j <= jmax) {
      // L2 prefetch of the t neighbours' rows of the same plane (requested by this CTA's t legs when the plane is current)
      const int r = (lane - 20) & 3, up = (lane - 20) >> 2;
      const uint32_t tt = up ? ((int)t == a.Lt - 1 ? 0 : t + 1) : (t == 0 ? a.Lt - 1 : t - 1);
      const uint32_t row_site = site_blk - tstride * t + tstride * tt + a.Lxh * r + zstride * (uint32_t)wrapz(zfirst - JOFF + j);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + ((size_t)((row_site * LS) >> LOGW) * 6 << LOGW)), "r"(ROW_BYTES) : "memory");
    }
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 6; q++) mbar_init(&bars[q], 1);
    cnt[0] = cnt[1] = cnt[2] = 0;
  }
  const bool tm_halo = MODE == 1 && a.t_comm && t == 0, tp_halo = MODE == 1 && a.t_comm && (int)t == a.Lt - 1;
  if (MODE == 1 && surf) {
    // acquire the t neighbours' epoch flags (peer-written, system scope) before any thread touches the receive buffers
    if (threadIdx.x == 3 || threadIdx.x == 7) {
      unsigned long long v;
      do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + threadIdx.x) : "memory"); } while (v < a.epoch);
    }
  }
  __syncthreads();
  // ---- prologue: ring planes 0, 1, 2 and the links of steps 0, 1 (warp 0; the links of step 2 follow at the first count)
  if (threadIdx.x < 32) { issue(0, 0); issue(1, 1); issue(2, a.N); }

  // float4 offsets (parity block) of the global legs at z = 0; the plane offset zo = z * dz is added per step
  const uint32_t o_tm = goff(site_xyt + (t == 0 ? tstride * (a.Lt - 1) : 0u - tstride));
  const uint32_t o_tp = goff(site_xyt + ((int)t == a.Lt - 1 ? 0u - tstride * (a.Lt - 1) : tstride));
  const uint32_t o_xm = goff(site_xyt - xh + (xh == 0 ? a.Lxh - 1 : xh - 1)), o_xp = goff(site_xyt - xh + (xh + 1 == a.Lxh ? 0 : xh + 1));
  const uint32_t o_ym = goff(site_xyt + (y == 0 ? a.Lxh * (a.Ly - 1) : 0u - a.Lxh)), o_yp = goff(site_xyt + (y + 1 == a.Ly ? 0u - a.Lxh * (a.Ly - 1) : a.Lxh));
  const uint32_t o_me = goff(site_xyt);
  // off-node t legs: face index = cb index with t removed; the neighbour stored half spinors in the same 16-lane blocking
  const int ip = 1 - p;
  const float4 *hb_tm = nullptr, *hb_tp = nullptr;
  uint32_t o_h = 0;
  if (MODE == 1) {
    if (tm_halo) hb_tm = a.halo_tm + (size_t)ip * a.hstride;
    if (tp_halo) hb_tp = a.halo_tp + (size_t)ip * a.hstride;
    const uint32_t i = (uint32_t)(xh + a.Lxh * y) * LS + s;
    o_h = (((i >> LOGW) * 3) << LOGW) + (i & (W - 1));
  }
  const uint32_t dzh = zstride * LS * 3;
  auto load_t = [&](SpinorP &f, bool halo, const float4 *hb, uint32_t o_t, int z) {
    if (MODE == 1 && halo) {
      const float4 *h = hb + o_h + (size_t)dzh * z;
#pragma unroll
      for (int q = 0; q < 3; q++) { const float4 v = h[q << LOGW]; f.c[2 * q] = pk(v.x, v.y); f.c[2 * q + 1] = pk(v.z, v.w); }
    } else load_spinor_p(f, in + o_t + (size_t)dz * z);
  };
  auto leg_t = [&](const SpinorP &f, bool halo, const float4 *Us, SpinorP &res, auto fwd) {
    constexpr int FWD = decltype(fwd)::value;
    if (MODE == 1 && halo) {
      HalfP chi;
#pragma unroll
      for (int q = 0; q < 6; q++) chi.c[q] = f.c[q];
      col3_leg_half<DAG, 3, FWD>(chi, Us, res);
    } else col3_leg<DAG, 3, FWD>(f, Us, res);
  };

  // ---- own column: plane zfirst in registers; CARRY: the z- projection of plane zfirst - 1 (one global load per column)
  SpinorP fo;
  HalfP hc;
  constexpr int SIGN_ZM = DAG ? -1 : +1;
  if (CARRY) {
    SpinorP fm;
    load_spinor_p(fm, in + o_me + (size_t)dz * wrapz(zfirst - 1));
    proj_p<2, SIGN_ZM>(hc, fm);
    mbar_wait(&bars[3], 0);
    load_ring_p(fo, mine);
  } else {
    mbar_wait(&bars[3], 0);
    mbar_wait(&bars[4], 0);
    load_ring_p(fo, mine + PLANE);
  }
  int sc = CARRY ? 0 : 1;                                      // ring slot of plane z
#pragma unroll 1
  for (int k = 0; k < a.N; k++) {
    const int z = wrapz(zfirst + k);
    const int sn = sc == 2 ? 0 : sc + 1, sm = sc == 0 ? 2 : sc - 1;   // slots of planes z+1 and z-1 (CARRY: z+2)
    const int ub = CARRY ? sc : sm;                              // k % 3
    const int pb = (p + a.origin_parity + y + z + (int)t) & 1;
    const float4 *Us = Usm + ub * UBUF + sl * FAST_USTRIDE;
    const float4 *cur = mine + sc * PLANE;                       // own element, plane z
    // ---- t- neighbour requested now, used two legs later
    SpinorP ft;
    load_t(ft, tm_halo, hb_tm, o_tm, z);
    mbar_wait(&bars[ub], (uint32_t)(k / 3) & 1);
    SpinorP res;
    // ---- z- : carried projection, or the own element of plane z-1 on the ring
    if (CARRY) {
      HalfP Uchi; LinkS u;
      lds_link(u, Us + 6 * 5);
      mult_p(Uchi, u, hc);
      recon_init_p<2, SIGN_ZM>(res, Uchi);
    } else {
      SpinorP fm;
      load_ring_p(fm, mine + sm * PLANE);
#pragma unroll
      for (int q = 0; q < 12; q++) res.c[q] = pk(0.f, 0.f);
      col3_leg<DAG, 2, 0>(fm, Us, res);
    }
    // ---- x leg on the own element (same x/2 index), then its z- projection for the step after next
    if (pb) col3_leg<DAG, 0, 0>(fo, Us, res); else col3_leg<DAG, 0, 1>(fo, Us, res);
    if (CARRY) proj_p<2, SIGN_ZM>(hc, fo);
    // ---- the other x neighbour: adjacent ring element or, at the block edge, a global load
    SpinorP fx;
    {
      const float4 *px = pb ? (xl < 3 ? cur + off_x1 : in + o_xp + (size_t)dz * z) : (xl > 0 ? cur + off_xm1 : in + o_xm + (size_t)dz * z);
      load_ring_p(fx, px);
    }
    // ---- t- leg, then request t+
    leg_t(ft, tm_halo, Us, res, std::integral_constant<int, 0>());
    load_t(ft, tp_halo, hb_tp, o_tp, z);
    if (pb) col3_leg<DAG, 0, 1>(fx, Us, res); else col3_leg<DAG, 0, 0>(fx, Us, res);
    // ---- y legs: slots +-4 inside the block
    load_ring_p(fx, yl > 0 ? cur + off_ym1 : in + o_ym + (size_t)dz * z);
    if (!CARRY) {
      // the oldest ring plane (z-1) and the links of step k-1 are dead for this warp
    }
    col3_leg<DAG, 1, 0>(fx, Us, res);
    load_ring_p(fx, yl < 3 ? cur + off_y1 : in + o_yp + (size_t)dz * z);
    // ---- this warp has made its last read of the oldest ring plane (CARRY: plane z; else plane z-1, read by the z- leg): count
    //      it; the warp that completes the count requests the plane that replaces it and the links of step k+2
    {
      __syncwarp();
      unsigned int old = 0;
      if (lane == 0) { __threadfence_block(); old = atomicAdd(&cnt[ub], 1u); }
      old = __shfl_sync(0xffffffffu, old, 0);
      if ((old + 1) % NWARP == 0) issue(k + 3, k + 2);
    }
    col3_leg<DAG, 1, 1>(fx, Us, res);
    // ---- t+ leg
    leg_t(ft, tp_halo, Us, res, std::integral_constant<int, 1>());
    // ---- z+ : wait for the plane (requested almost two steps ago), read the own element; it stays in registers for step k+1
    {
      const int jn = k + 1 + JOFF;                               // ring plane index of plane z+1
      mbar_wait(&bars[3 + sn], (uint32_t)(jn / 3) & 1);
    }
    load_ring_p(fo, mine + sn * PLANE);
    col3_leg<DAG, 2, 1>(fo, Us, res);
    // ---- epilogue
    const size_t offs = o_me + (size_t)dz * z;
    if (a.axpy[p] != nullptr) {
      SpinorP ax;
      load_spinor_p(ax, a.axpy[p] + offs);
      const f2 sa = pk(a.axpy_a, a.axpy_a), sb = pk(a.axpy_b, a.axpy_b);
#pragma unroll
      for (int q = 0; q < 12; q++) res.c[q] = fma2(sa, res.c[q], mul2(sb, ax.c[q]));
    }
    store_spinor_p(res, a.out[p] + offs);
    sc = sn;
  }
}

} // namespace gb
