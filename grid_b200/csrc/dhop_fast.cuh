// dhop_fast.cuh -- the tuned fp32 hopping-term kernel for sm_100a (the kernel bench.py times).
//
// Same arithmetic as dhop_kernel (ref: WilsonKernelsImplementation.h:57-68,112-163) with three Blackwell-specific
// changes, each measured in experiments/dhopx.cu and recorded in DESIGN.md / profiles/:
//   1. complex arithmetic on packed f32x2 registers (PTX fma/add/mul .f32x2 -> SASS FFMA2/FADD2/FMUL2): a complex
//      number is the natural (re,im) register pair a float4 load delivers, the link element is a broadcast scalar
//      operand and every "times +-i" is an operand swizzle (LO_HI) + per-lane negate, so spin projection, SU(3)
//      multiply and reconstruction take 60 instead of 108 FP instructions per leg with no packing moves;
//   2. the CTA's 16 x 8 links are staged into shared memory by per-site TMA bulk copies (cp.async.bulk + mbarrier)
//      with a 41-float4 site stride, so that the two sites a warp covers sit in different banks and the
//      broadcast LDS.128 is a single conflict-free wavefront (it is 2-4 wavefronts through L1 or unpadded smem);
//   3. legs are visited in +-mu pairs and CTAs cover 4x4 (x,y) micro-blocks so neighbours shared inside the CTA
//      are re-read from L1 while still resident.
// One thread per (checkerboard site, s): 16 sites x LS slices per CTA.
#pragma once
#include "dhop_kernel.cuh"

namespace gb {

constexpr int FAST_NSITE = 16;
constexpr int FAST_USTRIDE = 41; // float4 per site in shared memory (40 used)

struct FastArgs {
  const float4 *in[2];
  float4 *out[2];
  const float4 *U[2];
  const float4 *axpy[2];
  float axpy_a, axpy_b;
  int comm_dim_mask;
  int Lxh, Ly, Lz, Lt;
  int ibx, iby, Bz;            // micro-block (ibx x iby) inside rows, z-chunk for the L2 sweep
  FastDiv dibx, diby, dNxo, dNyo, dBz, dLt;
  uint32_t V4cb;                // sites visited per parity (the box volume when a box is set)
  int first_parity, origin_parity;
  int zo, to;                   // box origin in z and t: the kernel visits z in [zo, zo + nz), t in [to, to + nt) (dBz / dLt divide nz / nt)
  // fused multi-GPU mode: this rank's receive buffers (this epoch), epoch flags, rasterisation rotation
  const float4 *halo[8];
  size_t hstride[4];               // float4 between the parity-0 and parity-1 faces
  const unsigned long long *flags;
  unsigned long long epoch;
  int rot_z, rot_t;                // 1: visit coordinate (i+1) mod L at step i, so the surface planes come last
  // fused mode: the same launch also projects this rank's boundary slices and stores them into the neighbours'
  // receive buffers (peer-mapped).  Pack CTAs are interleaved 1-in-pack_ratio with the hop CTAs at the start of the grid.
  struct PackItemF { const float4 *src; float4 *dst; uint32_t nface; int mu, fwd, ip; uint32_t cta_start; } pack[16];
  int npack_items;
  uint32_t npack_ctas, pack_ratio, nhop_ctas_per_parity;
  unsigned int *pack_counter;
  unsigned long long *peer_flag[8];
};

// ------------------------------------------------------------------ packed f32x2 helpers
typedef unsigned long long f2; // lo = re, hi = im
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void upk(f2 d, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 swp(f2 a) { float lo, hi; upk(a, lo, hi); return pk(hi, lo); }
__device__ __forceinline__ f2 addi(f2 a, f2 z) { return fma2(swp(z), pk(-1.f, 1.f), a); } // a + i z
__device__ __forceinline__ f2 subi(f2 a, f2 z) { return fma2(swp(z), pk(1.f, -1.f), a); } // a - i z

struct SpinorP { f2 c[12]; }; // index = spin*3 + colour
struct HalfP { f2 c[6]; };
struct LinkS { float re[9], im[9]; };

__device__ __forceinline__ void load_spinor_p(SpinorP &f, const float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { float4 v = __ldg(p + (k << LOGW)); f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
}
__device__ __forceinline__ void load_spinor_rw(SpinorP &f, const float4 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { float4 v = p[k << LOGW]; f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
}
__device__ __forceinline__ void store_spinor_p(const SpinorP &f, float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { float4 v; upk(f.c[2 * k], v.x, v.y); upk(f.c[2 * k + 1], v.z, v.w); p[k << LOGW] = v; }
}
// ref: Grid/qcd/spin/TwoSpinor.h:75-133
template <int MU, int SIGN> __device__ __forceinline__ void proj_p(HalfP &h, const SpinorP &f) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const f2 f0 = f.c[c], f1 = f.c[3 + c], f2_ = f.c[6 + c], f3 = f.c[9 + c];
    if (MU == 0) { if (SIGN > 0) { h.c[c] = addi(f0, f3); h.c[3 + c] = addi(f1, f2_); } else { h.c[c] = subi(f0, f3); h.c[3 + c] = subi(f1, f2_); } }
    else if (MU == 1) { if (SIGN > 0) { h.c[c] = sub2(f0, f3); h.c[3 + c] = add2(f1, f2_); } else { h.c[c] = add2(f0, f3); h.c[3 + c] = sub2(f1, f2_); } }
    else if (MU == 2) { if (SIGN > 0) { h.c[c] = addi(f0, f2_); h.c[3 + c] = subi(f1, f3); } else { h.c[c] = subi(f0, f2_); h.c[3 + c] = addi(f1, f3); } }
    else { if (SIGN > 0) { h.c[c] = add2(f0, f2_); h.c[3 + c] = add2(f1, f3); } else { h.c[c] = sub2(f0, f2_); h.c[3 + c] = sub2(f1, f3); } }
  }
}
// ref: Grid/qcd/spin/TwoSpinor.h:193-354
template <int MU, int SIGN> __device__ __forceinline__ void recon_p(SpinorP &r, const HalfP &h) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const f2 h0 = h.c[c], h1 = h.c[3 + c];
    r.c[c] = add2(r.c[c], h0); r.c[3 + c] = add2(r.c[3 + c], h1);
    if (MU == 0) { if (SIGN > 0) { r.c[6 + c] = subi(r.c[6 + c], h1); r.c[9 + c] = subi(r.c[9 + c], h0); } else { r.c[6 + c] = addi(r.c[6 + c], h1); r.c[9 + c] = addi(r.c[9 + c], h0); } }
    else if (MU == 1) { if (SIGN > 0) { r.c[6 + c] = add2(r.c[6 + c], h1); r.c[9 + c] = sub2(r.c[9 + c], h0); } else { r.c[6 + c] = sub2(r.c[6 + c], h1); r.c[9 + c] = add2(r.c[9 + c], h0); } }
    else if (MU == 2) { if (SIGN > 0) { r.c[6 + c] = subi(r.c[6 + c], h0); r.c[9 + c] = addi(r.c[9 + c], h1); } else { r.c[6 + c] = addi(r.c[6 + c], h0); r.c[9 + c] = subi(r.c[9 + c], h1); } }
    else { if (SIGN > 0) { r.c[6 + c] = add2(r.c[6 + c], h0); r.c[9 + c] = add2(r.c[9 + c], h1); } else { r.c[6 + c] = sub2(r.c[6 + c], h0); r.c[9 + c] = sub2(r.c[9 + c], h1); } }
  }
}
__device__ __forceinline__ void lds_link(LinkS &u, const float4 *up) {
  const float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
}
// (U h)_r = sum_c (ur + i ui)(hr + i hi):  A = sum ur*(hr,hi), B = sum ui*(hi,hr), result = (A.lo - B.lo, A.hi + B.hi)
__device__ __forceinline__ void mult_p(HalfP &o, const LinkS &u, const HalfP &h) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      f2 A = mul2(pk(u.re[3 * r], u.re[3 * r]), h.c[3 * s]);
      f2 B = mul2(pk(u.im[3 * r], u.im[3 * r]), swp(h.c[3 * s]));
#pragma unroll
      for (int c = 1; c < 3; c++) {
        A = fma2(pk(u.re[3 * r + c], u.re[3 * r + c]), h.c[3 * s + c], A);
        B = fma2(pk(u.im[3 * r + c], u.im[3 * r + c]), swp(h.c[3 * s + c]), B);
      }
      o.c[3 * s + r] = fma2(B, pk(-1.f, 1.f), A);
    }
}

// ------------------------------------------------------------------ mbarrier / TMA bulk copy
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)),
               "r"(parity)
               : "memory");
}

// ------------------------------------------------------------------ the kernel
struct FastSite { int xh, y, z, t, pb; uint32_t site; };

template <int MU, int FWD> __device__ __forceinline__ bool fast_nbr(const FastArgs &a, const FastSite &c, uint32_t &nsite) {
  // returns true when the leg leaves the local volume in a decomposed dimension (handled by the exterior pass)
  const int Lmu = MU == 0 ? 2 * a.Lxh : MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt;
  const int coord = MU == 0 ? 2 * c.xh + c.pb : MU == 1 ? c.y : MU == 2 ? c.z : c.t;
  const bool at_edge = FWD ? (coord == Lmu - 1) : (coord == 0);
  if (MU == 0) {
    int nx;
    if (FWD) nx = c.pb ? (c.xh + 1 == a.Lxh ? 0 : c.xh + 1) : c.xh;
    else nx = c.pb ? c.xh : (c.xh == 0 ? a.Lxh - 1 : c.xh - 1);
    nsite = c.site - c.xh + nx;
  } else {
    const uint32_t stride = MU == 1 ? a.Lxh : MU == 2 ? a.Lxh * a.Ly : a.Lxh * a.Ly * a.Lz;
    if (FWD) nsite = at_edge ? c.site - (Lmu - 1) * stride : c.site + stride;
    else nsite = at_edge ? c.site + (Lmu - 1) * stride : c.site - stride;
  }
  return at_edge && ((a.comm_dim_mask >> MU) & 1);
}

template <int LS, int DAG, int MU, int FWD>
__device__ __forceinline__ void fast_leg(const float4 *__restrict__ in, uint32_t nsite, int s, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  const uint32_t i = nsite * LS + s;
  SpinorP f; HalfP chi, Uchi; LinkS u;
  load_spinor_p(f, in + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1)));
  proj_p<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}

// off-node leg of the fused mode: the neighbour rank already projected the spinor; read the half spinor it stored
// into this rank's receive buffer (face index = checkerboard index with dimension MU removed, as in pack_body)
template <int LS, int DAG, int MU, int FWD>
__device__ __forceinline__ void fast_halo_leg(const FastArgs &a, const FastSite &c, int s, int ip, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  uint32_t fi;
  if (MU == 0) fi = (uint32_t)(c.y >> 1) + (uint32_t)(a.Ly >> 1) * (c.z + a.Lz * c.t);
  else if (MU == 1) fi = (uint32_t)c.xh + (uint32_t)a.Lxh * (c.z + a.Lz * c.t);
  else if (MU == 2) fi = (uint32_t)c.xh + (uint32_t)a.Lxh * (c.y + a.Ly * c.t);
  else fi = (uint32_t)c.xh + (uint32_t)a.Lxh * (c.y + a.Ly * c.z);
  const uint32_t i = fi * LS + s;
  const float4 *hp = a.halo[FWD ? MU : MU + 4] + (size_t)ip * a.hstride[MU] + ((size_t)(i >> LOGW) * 3 << LOGW) + (i & (W - 1));
  HalfP chi, Uchi; LinkS u;
#pragma unroll
  for (int k = 0; k < 3; k++) { const float4 v = hp[k << LOGW]; chi.c[2 * k] = pk(v.x, v.y); chi.c[2 * k + 1] = pk(v.z, v.w); }
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}

// INTERIOR = 0: all legs are local (single rank).  INTERIOR = 1: skip legs that leave the local volume (a separate
// exterior pass adds them).  INTERIOR = 2: fused -- local legs first, then CTAs that own surface sites acquire the
// neighbours' epoch flags (peer-written over NVLink by pack_send_kernel) and add the off-node legs from the receive
// buffers: one kernel does the whole decomposed hop, and the surface CTAs are rasterised last.
// pack CTA of the fused mode: project one chunk of a boundary slice with the receiving leg's projector and store the
// half spinors into the neighbour's receive buffer over NVLink (same indexing as pack_body in halo_p2p.cu)
template <int LS, int DAG, int MU, int FWD>
__device__ __forceinline__ void fast_pack(const FastArgs &a, const FastArgs::PackItemF &it, uint32_t q) {
  const uint32_t fi = q / LS, s = q - fi * LS;
  int xh, y, z, t;
  const int slice = FWD ? 0 : (MU == 0 ? 2 * a.Lxh : MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt) - 1;
  uint32_t r = fi;
  if (MU == 0) {
    int yhalf = r % (a.Ly >> 1); r /= (a.Ly >> 1); z = r % a.Lz; t = r / a.Lz;
    int ypar = (slice + it.ip + a.origin_parity + z + t) & 1;
    y = 2 * yhalf + ypar; xh = slice >> 1;
  } else if (MU == 1) { xh = r % a.Lxh; r /= a.Lxh; z = r % a.Lz; t = r / a.Lz; y = slice; }
  else if (MU == 2) { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; t = r / a.Ly; z = slice; }
  else { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; z = r / a.Ly; t = slice; }
  const uint32_t site = xh + a.Lxh * (y + a.Ly * (z + a.Lz * t));
  const uint32_t i = site * LS + s;
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  SpinorP f; HalfP h;
  load_spinor_p(f, it.src + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1)));
  proj_p<MU, SIGN>(h, f);
  float4 *d = it.dst + ((size_t)(q >> LOGW) * 3 << LOGW) + (q & (W - 1));
#pragma unroll
  for (int k = 0; k < 3; k++) { float4 v; upk(h.c[2 * k], v.x, v.y); upk(h.c[2 * k + 1], v.z, v.w); d[k << LOGW] = v; }
}

template <int LS, int DAG, int INTERIOR>
__global__ void __launch_bounds__(FAST_NSITE *LS, (FAST_NSITE * LS <= 256) ? 3 : 1) dhop_fast_kernel(const FastArgs a) {
  __shared__ __align__(16) float4 Usm[FAST_NSITE * FAST_USTRIDE];
  __shared__ uint64_t bar;
  const int sl = threadIdx.x / LS, s = threadIdx.x % LS;
  int p = a.first_parity ^ (int)blockIdx.y;
  uint32_t hop_cta = blockIdx.x;
  if (INTERIOR == 2) {
    // 1D grid: [pack CTAs interleaved 1-in-ratio] + [hop CTAs of parity slot 0] + [parity slot 1]
    const uint32_t b = blockIdx.x;
    if (b < a.npack_ctas * a.pack_ratio && b % a.pack_ratio == 0) {
      const uint32_t pc = b / a.pack_ratio;
      int it = 0;
#pragma unroll 1
      for (int j = 1; j < a.npack_items; j++) if (pc >= a.pack[j].cta_start) it = j;
      const FastArgs::PackItemF &item = a.pack[it];
      const uint32_t q = (pc - item.cta_start) * (FAST_NSITE * LS) + threadIdx.x;
      if (q < item.nface * LS) {
        switch (item.mu * 2 + item.fwd) {
        case 0: fast_pack<LS, DAG, 0, 0>(a, item, q); break;
        case 1: fast_pack<LS, DAG, 0, 1>(a, item, q); break;
        case 2: fast_pack<LS, DAG, 1, 0>(a, item, q); break;
        case 3: fast_pack<LS, DAG, 1, 1>(a, item, q); break;
        case 4: fast_pack<LS, DAG, 2, 0>(a, item, q); break;
        case 5: fast_pack<LS, DAG, 2, 1>(a, item, q); break;
        case 6: fast_pack<LS, DAG, 3, 0>(a, item, q); break;
        default: fast_pack<LS, DAG, 3, 1>(a, item, q); break;
        }
      }
      // publish: fence the peer stores; the last pack CTA writes the epoch flags into the neighbours' memory
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(a.pack_counter, 1u);
        if (prev == a.npack_ctas - 1) {
          *a.pack_counter = 0;
          __threadfence_system();
#pragma unroll
          for (int k = 0; k < 8; k++)
            if (a.peer_flag[k]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flag[k]), "l"(a.epoch) : "memory");
        }
      }
      return;
    }
    const uint32_t before = min(a.npack_ctas, (b + a.pack_ratio - 1) / a.pack_ratio);
    uint32_t h = b - before;
    const uint32_t slot = h / a.nhop_ctas_per_parity;
    hop_cta = h - slot * a.nhop_ctas_per_parity;
    p = a.first_parity ^ (int)slot;
  }
  uint32_t r = hop_cta * FAST_NSITE + sl;
  const bool active = r < a.V4cb;
  if (!active) r = a.V4cb - 1;
  FastSite c;
  {
    uint32_t xl, yl, xo, yo, zl, t, zh;
    a.dibx.divmod(r, r, xl); a.diby.divmod(r, r, yl); a.dNxo.divmod(r, r, xo); a.dNyo.divmod(r, r, yo);
    a.dBz.divmod(r, r, zl); a.dLt.divmod(r, zh, t);
    c.xh = xo * a.ibx + xl; c.y = yo * a.iby + yl; c.z = a.zo + zh * a.Bz + zl; c.t = a.to + t;
    if (INTERIOR == 2) {
      if (a.rot_z) { c.z += 1; if (c.z == a.Lz) c.z = 0; }
      if (a.rot_t) { c.t += 1; if (c.t == a.Lt) c.t = 0; }
    }
    c.site = c.xh + a.Lxh * (c.y + a.Ly * (c.z + a.Lz * c.t));
    c.pb = (p + a.origin_parity + c.y + c.z + c.t) & 1;
  }
  uint32_t nb[8];
  bool off[8];
  off[0] = fast_nbr<0, 0>(a, c, nb[0]); off[1] = fast_nbr<0, 1>(a, c, nb[1]); off[2] = fast_nbr<1, 0>(a, c, nb[2]); off[3] = fast_nbr<1, 1>(a, c, nb[3]);
  off[4] = fast_nbr<2, 0>(a, c, nb[4]); off[5] = fast_nbr<2, 1>(a, c, nb[5]); off[6] = fast_nbr<3, 0>(a, c, nb[6]); off[7] = fast_nbr<3, 1>(a, c, nb[7]);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  // the barrier that publishes the mbarrier also tells every thread whether this CTA owns surface sites (fused mode),
  // so that interior CTAs never synchronise again
  int cta_has_off = 0;
  if (INTERIOR == 2) cta_has_off = __syncthreads_or(off[0] | off[1] | off[2] | off[3] | off[4] | off[5] | off[6] | off[7]);
  else __syncthreads();
  if (threadIdx.x == 0) mbar_expect_tx(&bar, FAST_NSITE * 640);
  if (s == 0) bulk_g2s(Usm + sl * FAST_USTRIDE, a.U[p] + (size_t)c.site * 40, 640, &bar);
  const float4 *__restrict__ in = a.in[1 - p];
  mbar_wait(&bar, 0);
  const float4 *Us = Usm + sl * FAST_USTRIDE;
  SpinorP res;
#pragma unroll
  for (int k = 0; k < 12; k++) res.c[k] = pk(0.f, 0.f);
  // CTAs without surface sites (the great majority) take the unconditional path: with every leg guarded by a per-thread
  // predicate the compiler cannot batch the loads of a +-mu pair, which costs ~15 % (measured: 1.16 vs 1.0 ms per hop)
  if (INTERIOR == 2 && !cta_has_off) {
#define GB_LEG(I, M, F) fast_leg<LS, DAG, M, F>(in, nb[I], s, Us, res)
    GB_LEG(0, 0, 0); GB_LEG(1, 0, 1); GB_LEG(2, 1, 0); GB_LEG(3, 1, 1);
    GB_LEG(4, 2, 0); GB_LEG(5, 2, 1); GB_LEG(6, 3, 0); GB_LEG(7, 3, 1);
#undef GB_LEG
  } else {
#define GB_LEG(I, M, F) if (!INTERIOR || !off[I]) fast_leg<LS, DAG, M, F>(in, nb[I], s, Us, res)
    GB_LEG(0, 0, 0); GB_LEG(1, 0, 1); GB_LEG(2, 1, 0); GB_LEG(3, 1, 1);
    GB_LEG(4, 2, 0); GB_LEG(5, 2, 1); GB_LEG(6, 3, 0); GB_LEG(7, 3, 1);
#undef GB_LEG
  }
  if (INTERIOR == 2) {
    if (cta_has_off) {
      if (threadIdx.x < 8 && ((a.comm_dim_mask >> (threadIdx.x & 3)) & 1)) {
        unsigned long long v;
        do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + threadIdx.x) : "memory"); } while (v < a.epoch);
      }
      __syncthreads();
      const int ip = 1 - p;
#define GB_HLEG(I, M, F) if (off[I]) fast_halo_leg<LS, DAG, M, F>(a, c, s, ip, Us, res)
      GB_HLEG(0, 0, 0); GB_HLEG(1, 0, 1); GB_HLEG(2, 1, 0); GB_HLEG(3, 1, 1);
      GB_HLEG(4, 2, 0); GB_HLEG(5, 2, 1); GB_HLEG(6, 3, 0); GB_HLEG(7, 3, 1);
#undef GB_HLEG
    }
  }
  if (!active) return;
  const uint32_t i = c.site * LS + s;
  const size_t offs = ((size_t)(i >> LOGW) * 6 << LOGW) + (i & (W - 1));
  if (a.axpy[p] != nullptr) {
    SpinorP ax;
    load_spinor_p(ax, a.axpy[p] + offs);
    const f2 sa = pk(a.axpy_a, a.axpy_a), sb = pk(a.axpy_b, a.axpy_b);
#pragma unroll
    for (int k = 0; k < 12; k++) res.c[k] = fma2(sa, res.c[k], mul2(sb, ax.c[k]));
  }
  store_spinor_p(res, a.out[p] + offs);
}

} // namespace gb
