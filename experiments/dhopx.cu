// dhopx.cu -- kernel-variant laboratory for the fp32 DWF hopping term (development tool, not shipped).
// Builds a standalone binary; every variant is checked against V0 (the library kernel's algorithm) and timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I grid_b200/csrc experiments/dhopx.cu -o experiments/dhopx
//   ./dhopx [L=32] [Ls=16] [iters=20] [variant mask]
#include "dhop_kernel.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

using namespace gb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct XArgs {
  const float4 *in; float4 *out; const float4 *U;
  int Ls, Lx, Lxh, Ly, Lz, Lt, By, Bz, Bt;
  FastDiv dLs, dLxh, dBy, dBz, dBt, dNy, dNz;
  uint32_t n5cb; int p; int fake; int pfd;
  int ibx, iby, ibz, ibt; FastDiv dibx, diby, dibz, dibt, dNxo, dNyo, dNzo, dNto; int nzlo;
};

__global__ void fill_kernel(float *p, size_t n, uint64_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (float)(u01(splitmix64(seed ^ i)) - 0.5);
}

struct Site { int xh, y, z, t, pb; uint32_t site; };
__device__ __forceinline__ Site decode_site(const XArgs &a, uint32_t r /* tile-ordered cb site index */) {
  Site c;
  if (a.ibx) { // inner 4D micro-block (ibx x iby x ibz x ibt cb-sites) fastest, then x,y,z,t block indices
    uint32_t xl, yl, zl, tl, xo, yo, zo, to;
    a.dibx.divmod(r, r, xl); a.diby.divmod(r, r, yl); a.dibz.divmod(r, r, zl); a.dibt.divmod(r, r, tl);
    uint32_t zh; a.dNxo.divmod(r, r, xo); a.dNyo.divmod(r, r, yo); a.dNzo.divmod(r, r, zo); a.dNto.divmod(r, zh, to); zo += zh * a.nzlo;
    c.xh = xo * a.ibx + xl; c.y = yo * a.iby + yl; c.z = zo * a.ibz + zl; c.t = to * a.ibt + tl;
  } else {
    uint32_t yl, zl, tl, yh, zh, th, xh;
    a.dLxh.divmod(r, r, xh); a.dBy.divmod(r, r, yl); a.dBz.divmod(r, r, zl); a.dBt.divmod(r, r, tl); a.dNy.divmod(r, r, yh); a.dNz.divmod(r, th, zh);
    c.xh = xh; c.y = yh * a.By + yl; c.z = zh * a.Bz + zl; c.t = th * a.Bt + tl;
  }
  c.site = c.xh + a.Lxh * (c.y + a.Ly * (c.z + a.Lz * c.t));
  c.pb = (a.p + c.y + c.z + c.t) & 1;
  return c;
}
template <int MU, int FWD> __device__ __forceinline__ uint32_t nbr_site(const XArgs &a, const Site &c) {
  if (a.fake) return c.site;
  if (MU == 0) {
    int nx;
    if (FWD) nx = c.pb ? (c.xh + 1 == a.Lxh ? 0 : c.xh + 1) : c.xh;
    else nx = c.pb ? c.xh : (c.xh == 0 ? a.Lxh - 1 : c.xh - 1);
    return c.site - c.xh + nx;
  }
  const int Lmu = MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt;
  const int coord = MU == 1 ? c.y : MU == 2 ? c.z : c.t;
  const uint32_t stride = MU == 1 ? a.Lxh : MU == 2 ? a.Lxh * a.Ly : a.Lxh * a.Ly * a.Lz;
  if (FWD) return coord == Lmu - 1 ? c.site - (Lmu - 1) * stride : c.site + stride;
  return coord == 0 ? c.site + (Lmu - 1) * stride : c.site - stride;
}

// ---------------------------------------------------------------- V0 / V1: one thread per (site, s)
template <int MU, int FWD> __device__ __forceinline__ void leg1(const XArgs &a, const Site &c, int s, const float4 *Usite, SpinorReg<float> &res) {
  constexpr int SIGN = FWD ? -1 : +1;
  const uint32_t i = nbr_site<MU, FWD>(a, c) * a.Ls + s;
  SpinorReg<float> f; HalfReg<float> chi, Uchi; LinkReg<float> u;
  load_spinor(f, a.in + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & 15));
  sp_proj<MU, SIGN>(chi, f);
  load_link(u, Usite + (FWD ? MU : MU + 4) * 5);
  mult_link(Uchi, u, chi);
  accum_recon<MU, SIGN>(res, Uchi);
}
template <int ORDER, int MINB> __global__ void __launch_bounds__(256, MINB) k_s1(const XArgs a) {
  const uint32_t q = blockIdx.x * 256 + threadIdx.x;
  if (q >= a.n5cb) return;
  uint32_t r, s; a.dLs.divmod(q, r, s);
  Site c = decode_site(a, r);
  const float4 *Usite = a.U + (size_t)c.site * 40;
  SpinorReg<float> res;
#pragma unroll
  for (int k = 0; k < 12; k++) { res.re[k] = 0; res.im[k] = 0; }
  if (ORDER == 0) {
    leg1<0, 0>(a, c, s, Usite, res); leg1<1, 0>(a, c, s, Usite, res); leg1<2, 0>(a, c, s, Usite, res); leg1<3, 0>(a, c, s, Usite, res);
    leg1<0, 1>(a, c, s, Usite, res); leg1<1, 1>(a, c, s, Usite, res); leg1<2, 1>(a, c, s, Usite, res); leg1<3, 1>(a, c, s, Usite, res);
  } else {
    leg1<0, 0>(a, c, s, Usite, res); leg1<0, 1>(a, c, s, Usite, res); leg1<1, 0>(a, c, s, Usite, res); leg1<1, 1>(a, c, s, Usite, res);
    leg1<2, 0>(a, c, s, Usite, res); leg1<2, 1>(a, c, s, Usite, res); leg1<3, 0>(a, c, s, Usite, res); leg1<3, 1>(a, c, s, Usite, res);
  }
  const uint32_t i = c.site * a.Ls + s;
  store_spinor(res, a.out + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & 15));
}

// ---------------------------------------------------------------- V3: two s per thread (s, s+8), Ls == 16
template <int MU, int FWD> __device__ __forceinline__ void leg2(const XArgs &a, const Site &c, int s, const float4 *Usite, SpinorReg<float> &r0, SpinorReg<float> &r1) {
  constexpr int SIGN = FWD ? -1 : +1;
  const float4 *p = a.in + ((size_t)nbr_site<MU, FWD>(a, c) * 6 << LOGW) + s;
  LinkReg<float> u;
  load_link(u, Usite + (FWD ? MU : MU + 4) * 5);
  SpinorReg<float> f0, f1;
  load_spinor(f0, p);
  load_spinor(f1, p + 8);
  HalfReg<float> chi, Uchi;
  sp_proj<MU, SIGN>(chi, f0); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r0, Uchi);
  sp_proj<MU, SIGN>(chi, f1); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r1, Uchi);
}
template <int MINB> __global__ void __launch_bounds__(256, MINB) k_s2(const XArgs a) {
  const uint32_t r = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int s = threadIdx.x & 7;
  if (r * 16 >= a.n5cb) return;
  Site c = decode_site(a, r);
  const float4 *Usite = a.U + (size_t)c.site * 40;
  SpinorReg<float> r0, r1;
#pragma unroll
  for (int k = 0; k < 12; k++) { r0.re[k] = 0; r0.im[k] = 0; r1.re[k] = 0; r1.im[k] = 0; }
  leg2<0, 0>(a, c, s, Usite, r0, r1); leg2<0, 1>(a, c, s, Usite, r0, r1); leg2<1, 0>(a, c, s, Usite, r0, r1); leg2<1, 1>(a, c, s, Usite, r0, r1);
  leg2<2, 0>(a, c, s, Usite, r0, r1); leg2<2, 1>(a, c, s, Usite, r0, r1); leg2<3, 0>(a, c, s, Usite, r0, r1); leg2<3, 1>(a, c, s, Usite, r0, r1);
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
  store_spinor(r0, o);
  store_spinor(r1, o + 8);
}

// ---------------------------------------------------------------- V4: S2 + the CTA's links staged in shared memory
template <int MU, int FWD> __device__ __forceinline__ void leg2s(const XArgs &a, const Site &c, int s, const float4 *Usm, SpinorReg<float> &r0, SpinorReg<float> &r1) {
  constexpr int SIGN = FWD ? -1 : +1;
  const float4 *p = a.in + ((size_t)nbr_site<MU, FWD>(a, c) * 6 << LOGW) + s;
  SpinorReg<float> f0, f1;
  load_spinor(f0, p);
  load_spinor(f1, p + 8);
  LinkReg<float> u;
  const float4 *up = Usm + (FWD ? MU : MU + 4) * 5;
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
  HalfReg<float> chi, Uchi;
  sp_proj<MU, SIGN>(chi, f0); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r0, Uchi);
  sp_proj<MU, SIGN>(chi, f1); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r1, Uchi);
}
template <int MINB> __global__ void __launch_bounds__(256, MINB) k_s2smem(const XArgs a) {
  __shared__ float4 Usm[32 * 40];
  const uint32_t r = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int s = threadIdx.x & 7;
  Site c = decode_site(a, r);
  // the 32 sites of this CTA are contiguous in x within rows; each site's 40 float4 are contiguous: 8 lanes copy 5 each
  {
    const float4 *g = a.U + (size_t)c.site * 40;
    float4 *d = Usm + (threadIdx.x >> 3) * 40;
#pragma unroll
    for (int j = 0; j < 5; j++) d[s + 8 * j] = __ldg(g + s + 8 * j);
  }
  __syncthreads();
  const float4 *Us = Usm + (threadIdx.x >> 3) * 40;
  SpinorReg<float> r0, r1;
#pragma unroll
  for (int k = 0; k < 12; k++) { r0.re[k] = 0; r0.im[k] = 0; r1.re[k] = 0; r1.im[k] = 0; }
  leg2s<0, 0>(a, c, s, Us, r0, r1); leg2s<0, 1>(a, c, s, Us, r0, r1); leg2s<1, 0>(a, c, s, Us, r0, r1); leg2s<1, 1>(a, c, s, Us, r0, r1);
  leg2s<2, 0>(a, c, s, Us, r0, r1); leg2s<2, 1>(a, c, s, Us, r0, r1); leg2s<3, 0>(a, c, s, Us, r0, r1); leg2s<3, 1>(a, c, s, Us, r0, r1);
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
  store_spinor(r0, o);
  store_spinor(r1, o + 8);
}

// S1 with links in shared memory (16 sites x 16 s per CTA)
template <int MU, int FWD> __device__ __forceinline__ void leg1s(const XArgs &a, const Site &c, int s, const float4 *Usm, SpinorReg<float> &res) {
  constexpr int SIGN = FWD ? -1 : +1;
  const uint32_t i = nbr_site<MU, FWD>(a, c) * a.Ls + s;
  SpinorReg<float> f; HalfReg<float> chi, Uchi; LinkReg<float> u;
  load_spinor(f, a.in + ((size_t)(i >> LOGW) * 6 << LOGW) + (i & 15));
  sp_proj<MU, SIGN>(chi, f);
  const float4 *up = Usm + (FWD ? MU : MU + 4) * 5;
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
  mult_link(Uchi, u, chi);
  accum_recon<MU, SIGN>(res, Uchi);
}
template <int MINB> __global__ void __launch_bounds__(256, MINB) k_s1smem(const XArgs a) {
  __shared__ float4 Usm[16 * 40];
  const uint32_t r = blockIdx.x * 16 + (threadIdx.x >> 4);
  const int s = threadIdx.x & 15;
  Site c = decode_site(a, r);
  {
    const float4 *g = a.U + (size_t)c.site * 40;
    float4 *d = Usm + (threadIdx.x >> 4) * 40;
    d[s] = __ldg(g + s); d[s + 16] = __ldg(g + s + 16);
    if (s < 8) d[s + 32] = __ldg(g + s + 32);
  }
  __syncthreads();
  const float4 *Us = Usm + (threadIdx.x >> 4) * 40;
  SpinorReg<float> res;
#pragma unroll
  for (int k = 0; k < 12; k++) { res.re[k] = 0; res.im[k] = 0; }
  leg1s<0, 0>(a, c, s, Us, res); leg1s<0, 1>(a, c, s, Us, res); leg1s<1, 0>(a, c, s, Us, res); leg1s<1, 1>(a, c, s, Us, res);
  leg1s<2, 0>(a, c, s, Us, res); leg1s<2, 1>(a, c, s, Us, res); leg1s<3, 0>(a, c, s, Us, res); leg1s<3, 1>(a, c, s, Us, res);
  store_spinor(res, a.out + ((size_t)c.site * 6 << LOGW) + s);
}


// ---------------------------------------------------------------- V5: S2 + smem links + L1 prefetch (CCTL.PF1) DIST legs ahead
__device__ __forceinline__ void pf_site(const float4 *base, int lane8) {
  // a neighbour site's 16 slices x 96 B = 1536 B = 12 lines; lane j touches line j, lanes 0..3 also line 8+j
  asm volatile("prefetch.global.L1 [%0];" ::"l"((const char *)base + lane8 * 128));
  if (lane8 < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char *)base + (8 + lane8) * 128));
}
template <int MU, int FWD> __device__ __forceinline__ void leg2p(const float4 *p, const float4 *Usm, SpinorReg<float> &r0, SpinorReg<float> &r1) {
  constexpr int SIGN = FWD ? -1 : +1;
  SpinorReg<float> f0, f1;
  load_spinor(f0, p);
  load_spinor(f1, p + 8);
  LinkReg<float> u;
  const float4 *up = Usm + (FWD ? MU : MU + 4) * 5;
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
  HalfReg<float> chi, Uchi;
  sp_proj<MU, SIGN>(chi, f0); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r0, Uchi);
  sp_proj<MU, SIGN>(chi, f1); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r1, Uchi);
}
template <int MINB, int DIST> __global__ void __launch_bounds__(256, MINB) k_s2pf(const XArgs a) {
  __shared__ float4 Usm[32 * 40];
  const uint32_t r = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int s = threadIdx.x & 7;
  Site c = decode_site(a, r);
  uint32_t nb[8];
  nb[0] = nbr_site<0, 0>(a, c); nb[1] = nbr_site<0, 1>(a, c); nb[2] = nbr_site<1, 0>(a, c); nb[3] = nbr_site<1, 1>(a, c);
  nb[4] = nbr_site<2, 0>(a, c); nb[5] = nbr_site<2, 1>(a, c); nb[6] = nbr_site<3, 0>(a, c); nb[7] = nbr_site<3, 1>(a, c);
#define NBP(i) (a.in + ((size_t)nb[i] * 6 << LOGW))
#pragma unroll
  for (int i = 0; i < DIST && i < 8; i++) pf_site(NBP(i), s);
  {
    const float4 *g = a.U + (size_t)c.site * 40;
    float4 *d = Usm + (threadIdx.x >> 3) * 40;
#pragma unroll
    for (int j = 0; j < 5; j++) d[s + 8 * j] = __ldg(g + s + 8 * j);
  }
  __syncthreads();
  const float4 *Us = Usm + (threadIdx.x >> 3) * 40;
  SpinorReg<float> r0, r1;
#pragma unroll
  for (int k = 0; k < 12; k++) { r0.re[k] = 0; r0.im[k] = 0; r1.re[k] = 0; r1.im[k] = 0; }
  if (DIST + 0 < 8) pf_site(NBP(DIST + 0), s); leg2p<0, 0>(NBP(0) + s, Us, r0, r1);
  if (DIST + 1 < 8) pf_site(NBP(DIST + 1), s); leg2p<0, 1>(NBP(1) + s, Us, r0, r1);
  if (DIST + 2 < 8) pf_site(NBP(DIST + 2), s); leg2p<1, 0>(NBP(2) + s, Us, r0, r1);
  if (DIST + 3 < 8) pf_site(NBP(DIST + 3), s); leg2p<1, 1>(NBP(3) + s, Us, r0, r1);
  if (DIST + 4 < 8) pf_site(NBP(DIST + 4), s); leg2p<2, 0>(NBP(4) + s, Us, r0, r1);
  if (DIST + 5 < 8) pf_site(NBP(DIST + 5), s); leg2p<2, 1>(NBP(5) + s, Us, r0, r1);
  if (DIST + 6 < 8) pf_site(NBP(DIST + 6), s); leg2p<3, 0>(NBP(6) + s, Us, r0, r1);
  if (DIST + 7 < 8) pf_site(NBP(DIST + 7), s); leg2p<3, 1>(NBP(7) + s, Us, r0, r1);
#undef NBP
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
  store_spinor(r0, o);
  store_spinor(r1, o + 8);
}

// ---------------------------------------------------------------- V6: four s per thread (s, s+4, s+8, s+12), smem links (padded stride 41)
template <int MU, int FWD> __device__ __forceinline__ void leg4(const float4 *p, const float4 *Usm, SpinorReg<float> (&r)[4]) {
  constexpr int SIGN = FWD ? -1 : +1;
  LinkReg<float> u;
  const float4 *up = Usm + (FWD ? MU : MU + 4) * 5;
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    SpinorReg<float> f; HalfReg<float> chi, Uchi;
    load_spinor(f, p + 4 * j);
    sp_proj<MU, SIGN>(chi, f); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r[j], Uchi);
  }
}
template <int MINB, int NT> __global__ void __launch_bounds__(NT, MINB) k_s4(const XArgs a) {
  constexpr int NSITE = NT / 4;
  __shared__ float4 Usm[NSITE * 41];
  const uint32_t r = blockIdx.x * NSITE + (threadIdx.x >> 2);
  const int s = threadIdx.x & 3;
  Site c = decode_site(a, r);
  {
    const float4 *g = a.U + (size_t)c.site * 40;
    float4 *d = Usm + (threadIdx.x >> 2) * 41;
#pragma unroll
    for (int j = 0; j < 10; j++) d[s + 4 * j] = __ldg(g + s + 4 * j);
  }
  __syncthreads();
  const float4 *Us = Usm + (threadIdx.x >> 2) * 41;
  SpinorReg<float> rr[4];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int k = 0; k < 12; k++) { rr[j].re[k] = 0; rr[j].im[k] = 0; }
#define NBP(M, F) (a.in + ((size_t)nbr_site<M, F>(a, c) * 6 << LOGW) + s)
  leg4<0, 0>(NBP(0, 0), Us, rr); leg4<0, 1>(NBP(0, 1), Us, rr); leg4<1, 0>(NBP(1, 0), Us, rr); leg4<1, 1>(NBP(1, 1), Us, rr);
  leg4<2, 0>(NBP(2, 0), Us, rr); leg4<2, 1>(NBP(2, 1), Us, rr); leg4<3, 0>(NBP(3, 0), Us, rr); leg4<3, 1>(NBP(3, 1), Us, rr);
#undef NBP
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
#pragma unroll
  for (int j = 0; j < 4; j++) store_spinor(rr[j], o + 4 * j);
}


// ---------------------------------------------------------------- V7: links staged by TMA bulk copies into bank-padded smem
// per-site stride 41 float4 (164 words): different sites land in different banks, so a warp-wide LDS.128 with one
// address per site is a single conflict-free wavefront.
constexpr int USTRIDE = 41;
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
__device__ __forceinline__ void lds_link(LinkReg<float> &u, const float4 *up) {
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
}
template <int MU, int FWD> __device__ __forceinline__ void leg1t(const float4 *p, const float4 *Usm, SpinorReg<float> &res) {
  constexpr int SIGN = FWD ? -1 : +1;
  SpinorReg<float> f; HalfReg<float> chi, Uchi; LinkReg<float> u;
  load_spinor(f, p);
  sp_proj<MU, SIGN>(chi, f);
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_link(Uchi, u, chi);
  accum_recon<MU, SIGN>(res, Uchi);
}
template <int MU, int FWD> __device__ __forceinline__ void leg2t(const float4 *p, const float4 *Usm, SpinorReg<float> &r0, SpinorReg<float> &r1) {
  constexpr int SIGN = FWD ? -1 : +1;
  SpinorReg<float> f0, f1;
  load_spinor(f0, p);
  load_spinor(f1, p + 8);
  LinkReg<float> u;
  lds_link(u, Usm + (FWD ? MU : MU + 4) * 5);
  HalfReg<float> chi, Uchi;
  sp_proj<MU, SIGN>(chi, f0); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r0, Uchi);
  sp_proj<MU, SIGN>(chi, f1); mult_link(Uchi, u, chi); accum_recon<MU, SIGN>(r1, Uchi);
}
// SPT = s-values per thread (1 or 2); TMA = 1: bulk-copy staging, 0: LDG+STS staging
template <int MINB, int SPT, int TMA> __global__ void __launch_bounds__(256, MINB) k_tma(const XArgs a) {
  constexpr int LPS = 16 / SPT;          // lanes per site
  constexpr int NSITE = 256 / LPS;
  __shared__ __align__(16) float4 Usm[NSITE * USTRIDE];
  __shared__ uint64_t bar;
  const int sl = threadIdx.x / LPS;      // site slot in the CTA
  const int s = threadIdx.x % LPS;
  const uint32_t r = blockIdx.x * NSITE + sl;
  Site c = decode_site(a, r);
  if (TMA) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, NSITE * 640);
    if (s == 0) bulk_g2s(Usm + sl * USTRIDE, a.U + (size_t)c.site * 40, 640, &bar);
  } else {
    const float4 *g = a.U + (size_t)c.site * 40;
    float4 *d = Usm + sl * USTRIDE;
    for (int j = s; j < 40; j += LPS) d[j] = __ldg(g + j);
  }
  uint32_t nb[8];
  nb[0] = nbr_site<0, 0>(a, c); nb[1] = nbr_site<0, 1>(a, c); nb[2] = nbr_site<1, 0>(a, c); nb[3] = nbr_site<1, 1>(a, c);
  nb[4] = nbr_site<2, 0>(a, c); nb[5] = nbr_site<2, 1>(a, c); nb[6] = nbr_site<3, 0>(a, c); nb[7] = nbr_site<3, 1>(a, c);
  if (TMA) mbar_wait(&bar, 0); else __syncthreads();
  const float4 *Us = Usm + sl * USTRIDE;
#define NBP(i) (a.in + ((size_t)nb[i] * 6 << LOGW) + s)
  if (SPT == 1) {
    SpinorReg<float> res;
#pragma unroll
    for (int k = 0; k < 12; k++) { res.re[k] = 0; res.im[k] = 0; }
    leg1t<0, 0>(NBP(0), Us, res); leg1t<0, 1>(NBP(1), Us, res); leg1t<1, 0>(NBP(2), Us, res); leg1t<1, 1>(NBP(3), Us, res);
    leg1t<2, 0>(NBP(4), Us, res); leg1t<2, 1>(NBP(5), Us, res); leg1t<3, 0>(NBP(6), Us, res); leg1t<3, 1>(NBP(7), Us, res);
    store_spinor(res, a.out + ((size_t)c.site * 6 << LOGW) + s);
  } else {
    SpinorReg<float> r0, r1;
#pragma unroll
    for (int k = 0; k < 12; k++) { r0.re[k] = 0; r0.im[k] = 0; r1.re[k] = 0; r1.im[k] = 0; }
    leg2t<0, 0>(NBP(0), Us, r0, r1); leg2t<0, 1>(NBP(1), Us, r0, r1); leg2t<1, 0>(NBP(2), Us, r0, r1); leg2t<1, 1>(NBP(3), Us, r0, r1);
    leg2t<2, 0>(NBP(4), Us, r0, r1); leg2t<2, 1>(NBP(5), Us, r0, r1); leg2t<3, 0>(NBP(6), Us, r0, r1); leg2t<3, 1>(NBP(7), Us, r0, r1);
    float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
    store_spinor(r0, o);
    store_spinor(r1, o + 8);
  }
#undef NBP
}


// ---------------------------------------------------------------- V8: packed f32x2 complex arithmetic (FFMA2/FADD2), TMA-staged padded links
typedef unsigned long long f2; // (lo = re, hi = im)
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void upk(f2 d, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 swp(f2 a) { float lo, hi; upk(a, lo, hi); return pk(hi, lo); }
__device__ __forceinline__ f2 addi(f2 a, f2 z) { return fma2(swp(z), pk(-1.f, 1.f), a); }  // a + i z
__device__ __forceinline__ f2 subi(f2 a, f2 z) { return fma2(swp(z), pk(1.f, -1.f), a); }  // a - i z

struct SpinorP { f2 c[12]; };
struct HalfP { f2 c[6]; };
__device__ __forceinline__ void load_spinor_p(SpinorP &f, const float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { float4 v = __ldg(p + (k << LOGW)); f.c[2 * k] = pk(v.x, v.y); f.c[2 * k + 1] = pk(v.z, v.w); }
}
__device__ int g_stream_store;
__device__ __forceinline__ void store_spinor_p(const SpinorP &f, float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float4 v; upk(f.c[2 * k], v.x, v.y); upk(f.c[2 * k + 1], v.z, v.w);
#ifdef STREAM_STORE
    __stcs(p + (k << LOGW), v);
#else
    p[k << LOGW] = v;
#endif
  }
}
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
struct LinkP { float re[9], im[9]; };
__device__ __forceinline__ void lds_link_p(LinkP &u, const float4 *up) {
  float4 v0 = up[0], v1 = up[1], v2 = up[2], v3 = up[3], v4 = up[4];
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w; u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w; u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
}
// (U h)_r = sum_c (ur + i ui)(hr + i hi):  A = sum ur*(hr,hi) ; B = sum ui*(hi,hr) ; result = (A.lo - B.lo, A.hi + B.hi)
// every multiplier is a broadcast scalar and every swap / sign is an operand modifier of FFMA2: no packing moves.
__device__ __forceinline__ void mult_p(HalfP &o, const LinkP &u, const HalfP &h) {
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
template <int MU, int FWD> __device__ __forceinline__ void leg1k(const float4 *p, const float4 *Usm, SpinorP &res) {
  constexpr int SIGN = FWD ? -1 : +1;
  SpinorP f; HalfP chi, Uchi; LinkP u;
  load_spinor_p(f, p);
  proj_p<MU, SIGN>(chi, f);
  lds_link_p(u, Usm + (FWD ? MU : MU + 4) * 5);
  mult_p(Uchi, u, chi);
  recon_p<MU, SIGN>(res, Uchi);
}
template <int MU, int FWD> __device__ __forceinline__ void leg2k(const float4 *p, const float4 *Usm, SpinorP &r0, SpinorP &r1) {
  constexpr int SIGN = FWD ? -1 : +1;
  SpinorP f0, f1;
  load_spinor_p(f0, p);
  load_spinor_p(f1, p + 8);
  LinkP u;
  lds_link_p(u, Usm + (FWD ? MU : MU + 4) * 5);
  HalfP chi, Uchi;
  proj_p<MU, SIGN>(chi, f0); mult_p(Uchi, u, chi); recon_p<MU, SIGN>(r0, Uchi);
  proj_p<MU, SIGN>(chi, f1); mult_p(Uchi, u, chi); recon_p<MU, SIGN>(r1, Uchi);
}
template <int MINB, int SPT> __global__ void __launch_bounds__(256, MINB) k_pk(const XArgs a) {
  constexpr int LPS = 16 / SPT;
  constexpr int NSITE = 256 / LPS;
  __shared__ __align__(16) float4 Usm[NSITE * USTRIDE];
  __shared__ uint64_t bar;
  const int sl = threadIdx.x / LPS;
  const int s = threadIdx.x % LPS;
  const uint32_t r = blockIdx.x * NSITE + sl;
  Site c = decode_site(a, r);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) mbar_expect_tx(&bar, NSITE * 640);
  if (s == 0) bulk_g2s(Usm + sl * USTRIDE, a.U + (size_t)c.site * 40, 640, &bar);
  if (a.pfd && s == 1) {
    // L2 prefetch of the first-touch (t+1) neighbour block of the site slot handled pfd CTAs from now, plus its links
    const uint32_t rf = r + (uint32_t)a.pfd * NSITE;
    if (rf * 16 < a.n5cb) {
      Site cf = decode_site(a, rf);
      const uint32_t nf = nbr_site<3, 1>(a, cf);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.in + ((size_t)nf * 6 << LOGW)), "r"(1536) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.U + (size_t)cf.site * 40), "r"(640) : "memory");
    }
  }
  uint32_t nb[8];
  nb[0] = nbr_site<0, 0>(a, c); nb[1] = nbr_site<0, 1>(a, c); nb[2] = nbr_site<1, 0>(a, c); nb[3] = nbr_site<1, 1>(a, c);
  nb[4] = nbr_site<2, 0>(a, c); nb[5] = nbr_site<2, 1>(a, c); nb[6] = nbr_site<3, 0>(a, c); nb[7] = nbr_site<3, 1>(a, c);
  mbar_wait(&bar, 0);
  const float4 *Us = Usm + sl * USTRIDE;
#define NBP(i) (a.in + ((size_t)nb[i] * 6 << LOGW) + s)
  if (SPT == 1) {
    SpinorP res;
#pragma unroll
    for (int k = 0; k < 12; k++) res.c[k] = pk(0.f, 0.f);
    leg1k<0, 0>(NBP(0), Us, res); leg1k<0, 1>(NBP(1), Us, res); leg1k<1, 0>(NBP(2), Us, res); leg1k<1, 1>(NBP(3), Us, res);
    leg1k<2, 0>(NBP(4), Us, res); leg1k<2, 1>(NBP(5), Us, res); leg1k<3, 0>(NBP(6), Us, res); leg1k<3, 1>(NBP(7), Us, res);
    store_spinor_p(res, a.out + ((size_t)c.site * 6 << LOGW) + s);
  } else {
    SpinorP r0, r1;
#pragma unroll
    for (int k = 0; k < 12; k++) { r0.c[k] = pk(0.f, 0.f); r1.c[k] = pk(0.f, 0.f); }
    leg2k<0, 0>(NBP(0), Us, r0, r1); leg2k<0, 1>(NBP(1), Us, r0, r1); leg2k<1, 0>(NBP(2), Us, r0, r1); leg2k<1, 1>(NBP(3), Us, r0, r1);
    leg2k<2, 0>(NBP(4), Us, r0, r1); leg2k<2, 1>(NBP(5), Us, r0, r1); leg2k<3, 0>(NBP(6), Us, r0, r1); leg2k<3, 1>(NBP(7), Us, r0, r1);
    float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
    store_spinor_p(r0, o);
    store_spinor_p(r1, o + 8);
  }
#undef NBP
}


// ---------------------------------------------------------------- V9: memory-pattern-only probe: same loads/stores as the hop, trivial math
template <int MINB, int NLEG> __global__ void __launch_bounds__(256, MINB) k_mem(const XArgs a) {
  const int sl = threadIdx.x >> 4, s = threadIdx.x & 15;
  const uint32_t r = blockIdx.x * 16 + sl;
  Site c = decode_site(a, r);
  uint32_t nb[8];
  nb[0] = nbr_site<0, 0>(a, c); nb[1] = nbr_site<0, 1>(a, c); nb[2] = nbr_site<1, 0>(a, c); nb[3] = nbr_site<1, 1>(a, c);
  nb[4] = nbr_site<2, 0>(a, c); nb[5] = nbr_site<2, 1>(a, c); nb[6] = nbr_site<3, 0>(a, c); nb[7] = nbr_site<3, 1>(a, c);
  float4 acc[6];
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int l = 0; l < NLEG; l++) {
    const float4 *p = a.in + ((size_t)nb[l] * 6 << LOGW) + s;
#pragma unroll
    for (int k = 0; k < 6; k++) { float4 v = __ldg(p + (k << LOGW)); acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w; }
  }
  const float4 *g = a.U + (size_t)c.site * 40;
  float4 u = __ldg(g + s), u2 = __ldg(g + 16 + s);
  acc[0].x += u.x + u2.y;
  if (s < 8) { float4 u3 = __ldg(g + 32 + s); acc[1].x += u3.x; }
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
#pragma unroll
  for (int k = 0; k < 6; k++) o[k << LOGW] = acc[k];
}


// ---------------------------------------------------------------- V10: memory probe with L2 eviction-policy hints
__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ float4 ldg_hint(const float4 *p, uint64_t pol) {
  float4 v; asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void stg_hint(float4 *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// HIN: 0 none, 1 evict_last on inputs ; HOUT: 0 none, 1 evict_first on outputs+gauge
template <int MINB, int HIN, int HOUT> __global__ void __launch_bounds__(256, MINB) k_memh(const XArgs a) {
  const int sl = threadIdx.x >> 4, s = threadIdx.x & 15;
  const uint32_t r = blockIdx.x * 16 + sl;
  Site c = decode_site(a, r);
  const uint64_t pl = policy_evict_last(), pf = policy_evict_first();
  uint32_t nb[8];
  nb[0] = nbr_site<0, 0>(a, c); nb[1] = nbr_site<0, 1>(a, c); nb[2] = nbr_site<1, 0>(a, c); nb[3] = nbr_site<1, 1>(a, c);
  nb[4] = nbr_site<2, 0>(a, c); nb[5] = nbr_site<2, 1>(a, c); nb[6] = nbr_site<3, 0>(a, c); nb[7] = nbr_site<3, 1>(a, c);
  float4 acc[6];
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int l = 0; l < 8; l++) {
    const float4 *p = a.in + ((size_t)nb[l] * 6 << LOGW) + s;
#pragma unroll
    for (int k = 0; k < 6; k++) { float4 v = HIN ? ldg_hint(p + (k << LOGW), pl) : __ldg(p + (k << LOGW)); acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w; }
  }
  const float4 *g = a.U + (size_t)c.site * 40;
  float4 u = HOUT ? ldg_hint(g + s, pf) : __ldg(g + s), u2 = HOUT ? ldg_hint(g + 16 + s, pf) : __ldg(g + 16 + s);
  acc[0].x += u.x + u2.y;
  if (s < 8) { float4 u3 = HOUT ? ldg_hint(g + 32 + s, pf) : __ldg(g + 32 + s); acc[1].x += u3.x; }
  float4 *o = a.out + ((size_t)c.site * 6 << LOGW) + s;
#pragma unroll
  for (int k = 0; k < 6; k++) { if (HOUT) stg_hint(o + (k << LOGW), acc[k], pf); else o[k << LOGW] = acc[k]; }
}

// ---------------------------------------------------------------- driver
int main(int argc, char **argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 32, Ls = argc > 2 ? atoi(argv[2]) : 16, iters = argc > 3 ? atoi(argv[3]) : 20;
  const unsigned mask = argc > 4 ? strtoul(argv[4], 0, 0) : 0xffffffffu;
  const int bz = argc > 5 ? atoi(argv[5]) : 8;
  const size_t V4cb = (size_t)L * L * L * L / 2, n5 = V4cb * Ls;
  const size_t fvec = n5 * 6, uvec = V4cb * 40;
  float4 *in, *out, *ref, *U;
  CK(cudaMalloc(&in, fvec * 16)); CK(cudaMalloc(&out, fvec * 16)); CK(cudaMalloc(&ref, fvec * 16)); CK(cudaMalloc(&U, uvec * 16));
  fill_kernel<<<(fvec * 4 + 255) / 256, 256>>>((float *)in, fvec * 4, 1);
  fill_kernel<<<(uvec * 4 + 255) / 256, 256>>>((float *)U, uvec * 4, 2);
  XArgs a; a.in = in; a.out = ref; a.U = U; a.Ls = Ls; a.Lx = L; a.Lxh = L / 2; a.Ly = a.Lz = a.Lt = L;
  a.By = getenv("BY") ? atoi(getenv("BY")) : L; a.Bz = bz; a.Bt = getenv("BT") ? atoi(getenv("BT")) : L;
  a.dLs = FastDiv(Ls); a.dLxh = FastDiv(a.Lxh); a.dBy = FastDiv(a.By); a.dBz = FastDiv(a.Bz); a.dBt = FastDiv(a.Bt);
  a.dNy = FastDiv(a.Ly / a.By); a.dNz = FastDiv(a.Lz / a.Bz); a.n5cb = (uint32_t)n5; a.p = 0; a.fake = getenv("FAKE") ? 1 : 0; a.pfd = getenv("PFD") ? atoi(getenv("PFD")) : 0;
  a.ibx = 0;
  if (getenv("IB")) { sscanf(getenv("IB"), "%d,%d,%d,%d", &a.ibx, &a.iby, &a.ibz, &a.ibt);
    a.dibx = FastDiv(a.ibx); a.diby = FastDiv(a.iby); a.dibz = FastDiv(a.ibz); a.dibt = FastDiv(a.ibt);
    a.dNxo = FastDiv(a.Lxh / a.ibx); a.dNyo = FastDiv(a.Ly / a.iby); a.nzlo = a.Bz / a.ibz; a.dNzo = FastDiv(a.nzlo); a.dNto = FastDiv(a.Lt / a.ibt); }
  const int carve = getenv("CARVE") ? atoi(getenv("CARVE")) : -1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_s1<0, 1><<<(n5 + 255) / 256, 256>>>(a);
  CK(cudaDeviceSynchronize());
  a.out = out;
  std::vector<float> href(fvec * 4), hout(fvec * 4);
  CK(cudaMemcpy(href.data(), ref, fvec * 16, cudaMemcpyDeviceToHost));
  auto run = [&](const char *name, int bit, auto launch) {
    if (!((mask >> bit) & 1)) return;
    CK(cudaMemset(out, 0, fvec * 16));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hout.data(), out, fvec * 16, cudaMemcpyDeviceToHost));
    double md = 0; for (size_t i = 0; i < hout.size(); i++) md = fmax(md, fabs((double)hout[i] - href[i]));
    for (int i = 0; i < 3; i++) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) launch();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("%-28s %8.4f ms  %7.1f GF/s  alg %6.1f GB/s  maxdiff %.2e\n", name, ms, 1320.0 * n5 / ms / 1e6, (192.0 + 576.0 / Ls) * n5 / ms / 1e6, md);
    fflush(stdout);
  };
  const unsigned g1 = (n5 + 255) / 256, g2 = (n5 / 2 + 255) / 256;
  run("mem probe 8 legs", 26, [&] { k_mem<3, 8><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("mem probe 8 legs occ8", 27, [&] { k_mem<8, 8><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("mem probe 1 leg", 28, [&] { k_mem<8, 1><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("memh in=last", 29, [&] { k_memh<3, 1, 0><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("memh out=first", 30, [&] { k_memh<3, 0, 1><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("memh both", 31, [&] { k_memh<3, 1, 1><<<(unsigned)((n5 + 255) / 256), 256>>>(a); });
  run("s1 order0 (library)", 0, [&] { k_s1<0, 1><<<g1, 256>>>(a); });
  run("s1 order1 (pairs)", 1, [&] { k_s1<1, 1><<<g1, 256>>>(a); });
  run("s1 order1 minb3", 2, [&] { k_s1<1, 3><<<g1, 256>>>(a); });
  run("s1 order1 minb4", 3, [&] { k_s1<1, 4><<<g1, 256>>>(a); });
  run("s1 smem-links minb3", 4, [&] { k_s1smem<3><<<g1, 256>>>(a); });
  run("s1 smem-links minb4", 5, [&] { k_s1smem<4><<<g1, 256>>>(a); });
  if (Ls == 16) {
    run("s2 minb1", 6, [&] { k_s2<1><<<g2, 256>>>(a); });
    run("s2 minb2", 7, [&] { k_s2<2><<<g2, 256>>>(a); });
    run("s2 minb3", 8, [&] { k_s2<3><<<g2, 256>>>(a); });
    run("s2 smem-links minb2", 9, [&] { k_s2smem<2><<<g2, 256>>>(a); });
    run("s2 smem-links minb3", 10, [&] { k_s2smem<3><<<g2, 256>>>(a); });
    run("s2 smem pf1 minb2", 11, [&] { k_s2pf<2, 1><<<g2, 256>>>(a); });
    run("s2 smem pf2 minb2", 12, [&] { k_s2pf<2, 2><<<g2, 256>>>(a); });
    run("s2 smem pf3 minb2", 13, [&] { k_s2pf<2, 3><<<g2, 256>>>(a); });
    run("s2 smem pf8 minb2", 14, [&] { k_s2pf<2, 8><<<g2, 256>>>(a); });
    run("s1 padded smem minb3", 18, [&] { k_tma<3, 1, 0><<<g1, 256>>>(a); });
    run("s1 padded TMA minb3", 19, [&] { k_tma<3, 1, 1><<<g1, 256>>>(a); });
    run("s2 padded smem minb2", 20, [&] { k_tma<2, 2, 0><<<g2, 256>>>(a); });
    run("s2 padded TMA minb2", 21, [&] { k_tma<2, 2, 1><<<g2, 256>>>(a); });
    if (carve >= 0) {
    CK(cudaFuncSetAttribute(k_pk<3, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    CK(cudaFuncSetAttribute(k_pk<2, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    CK(cudaFuncSetAttribute(k_pk<4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    CK(cudaFuncSetAttribute(k_pk<2, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  }
  run("s1 packed TMA minb3", 22, [&] { k_pk<3, 1><<<g1, 256>>>(a); });
    run("s1 packed TMA minb4", 23, [&] { k_pk<4, 1><<<g1, 256>>>(a); });
    run("s2 packed TMA minb2", 24, [&] { k_pk<2, 2><<<g2, 256>>>(a); });
    run("s1 packed TMA minb2", 25, [&] { k_pk<2, 1><<<g1, 256>>>(a); });
    const unsigned g4 = (n5 / 4 + 255) / 256, g4b = (n5 / 4 + 127) / 128;
    run("s4 nt256 minb1", 15, [&] { k_s4<1, 256><<<g4, 256>>>(a); });
    run("s4 nt128 minb2", 16, [&] { k_s4<2, 128><<<g4b, 128>>>(a); });
    run("s4 nt128 minb3", 17, [&] { k_s4<3, 128><<<g4b, 128>>>(a); });
  }
  return 0;
}
