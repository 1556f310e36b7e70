// dhop_kernel.cuh -- the Wilson / domain-wall hopping term site kernel (hand-written for sm_100a).
//
// Replaces WilsonKernels::GenericDhopSite / GenericDhopSiteDag (+Int/Ext variants)
//   ref: Grid/qcd/action/fermion/implementation/WilsonKernelsImplementation.h:57-68 (leg), :112-163 (site),
//        :167-285 (interior / exterior), and the launch macros :415-433.
// Differences by design (not a port):
//   * one thread per (checkerboard 4D site, s); lanes of a half warp are the 16 fifth-dimension slices of one
//     4D site, so spinor loads are 256-byte contiguous float4 rows and the link load is a broadcast;
//   * neighbour addressing is arithmetic (no StencilEntry table, ref: Grid/stencil/Stencil.h:79-91,134-136);
//   * CTAs are rasterised in (y,z,t)-blocked order so that t/z neighbours are re-read from L2, not HBM;
//   * full-lattice fields are [even block][odd block], so Dhop(full) is the two checkerboard hops in one launch.
#pragma once
#include "internal.hpp"
#include "kernels_common.cuh"

namespace gb {

struct DhopArgs {
  const void *in[2];   // in[p]  = parity-p block of the input field (legs of output parity q read in[1-q])
  void *out[2];        // out[p] = parity-p block of the output field
  const void *U[2];    // doubled links of parity-p output sites: [site][8][LV] vecs
  // optional fused epilogue: out = a*out_hop + b*axpy_in  (used by DW / M composites), disabled when axpy[p]==nullptr
  const void *axpy[2];
  double axpy_a, axpy_b;
  // halo buffers for legs that leave the local volume (multi-GPU): halo[point] holds projected half spinors
  const void *halo[8];
  size_t halo_parity_stride[4]; // vecs between the parity-0 and parity-1 faces of a halo buffer
  // peer-to-peer halos: epoch flags written by the neighbours' pack kernels (nullptr: halos already complete)
  const unsigned long long *flags;
  unsigned long long epoch;
  int comm_dim_mask;   // bit mu set: dimension mu is decomposed over ranks
  // 12-real link storage (gb_op_set_link_reconstruct): U[] then holds two rows of the bare SU(3) matrices (Recon12<T>::LV vecs per link),
  // the third row is rebuilt in registers and the folded-in factor (-1/2, times the boundary phase on the global boundary) is applied
  // to the product: link_c = {re, im} per stencil point on the boundary, (-1/2, 0) elsewhere
  int recon12;
  int gL[4], origin[4];
  double bnd_re[8], bnd_im[8];
  int halo_lowp;       // halos travel one precision down (load_half_lowp): fp32 operators bf16, fp64 operators fp32
  int mode;            // 0 = all legs (single rank or serial comms), 1 = interior legs only, 2 = exterior legs only (accumulate)
  int Ls, Lx, Lxh, Ly, Lz, Lt;
  int By, Bz, Bt;      // rasterisation block extents (divide Ly, Lz, Lt)
  FastDiv dLs, dLxh, dBy, dBz, dBt, dNy, dNz;
  uint32_t n5cb;       // V4cb * Ls
  // optional sub-box (exterior pass over one surface slab): cb coordinates (xh,y,z,t) in [bo, bo+be)
  int box_on;
  int bo[4], be[4];
  FastDiv dbe0, dbe1, dbe2;
  int first_parity;    // output parity handled by blockIdx.y == 0
  int origin_parity;
  int leg_mask;        // bit (FWD ? mu : mu + 4) set: that leg contributes.  0xFF = the hopping term; one bit = DhopDir (generic kernel only)
};

// ------------------------------------------------------------------ register-resident site objects
template <class T> struct SpinorReg { T re[12], im[12]; };      // index = spin*3 + colour
template <class T> struct HalfReg { T re[6], im[6]; };          // index = hspin*3 + colour
template <class T> struct LinkReg { T re[9], im[9]; };          // index = row*3 + col

__device__ __forceinline__ void load_spinor(SpinorReg<float> &f, const float4 *__restrict__ p /* element (blk,0,lane) */) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float4 v = __ldg(p + (k << LOGW));
    f.re[2 * k] = v.x; f.im[2 * k] = v.y; f.re[2 * k + 1] = v.z; f.im[2 * k + 1] = v.w;
  }
}
__device__ __forceinline__ void load_spinor(SpinorReg<double> &f, const double2 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 12; k++) {
    double2 v = __ldg(p + (k << LOGW));
    f.re[k] = v.x; f.im[k] = v.y;
  }
}
__device__ __forceinline__ void store_spinor(const SpinorReg<float> &f, float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) p[k << LOGW] = make_float4(f.re[2 * k], f.im[2 * k], f.re[2 * k + 1], f.im[2 * k + 1]);
}
__device__ __forceinline__ void store_spinor(const SpinorReg<double> &f, double2 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 12; k++) p[k << LOGW] = make_double2(f.re[k], f.im[k]);
}
// Halo half-spinors are written by the neighbours' pack kernels WHILE this kernel may already be running (peer-to-peer path:
// the epoch flag is acquired on the device), and the receive buffers are reused every second epoch -- so they are read with
// plain coherent loads, never through the non-coherent (ld.global.nc / __ldg) path, which is only legal for data that is
// read-only for the whole kernel lifetime and is not ordered by the flag acquire.
__device__ __forceinline__ void load_half(HalfReg<float> &h, const float4 *p) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float4 v = p[k << LOGW];
    h.re[2 * k] = v.x; h.im[2 * k] = v.y; h.re[2 * k + 1] = v.z; h.im[2 * k + 1] = v.w;
  }
}
__device__ __forceinline__ void load_half(HalfReg<double> &h, const double2 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double2 v = p[k << LOGW];
    h.re[k] = v.x; h.im[k] = v.y;
  }
}
__device__ __forceinline__ void store_half(const HalfReg<float> &h, float4 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 3; k++) p[k << LOGW] = make_float4(h.re[2 * k], h.im[2 * k], h.re[2 * k + 1], h.im[2 * k + 1]);
}
__device__ __forceinline__ void store_half(const HalfReg<double> &h, double2 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) p[k << LOGW] = make_double2(h.re[k], h.im[k]);
}
// Compressed halos ("half-precision comms"; ref: FermionOperatorImpl.h:96-137 LowerPrecisionMapper / CoeffRealHalfComms,
// WilsonImpl.h:59-65 SiteHalfCommSpinor, WilsonCompressor.h:244-306, tests/Test_dwf_mixedcg_prec_halfcomms.cc:71 DomainWallFermionFH): the projected half spinor travels one precision down and is widened
// again by the consumer; arithmetic stays in the operator's precision.  Every 16-byte vec of the uncompressed layout becomes an 8-byte
// vec at the SAME vec index, so face indexing is unchanged and the bytes on the link halve.  fp64 operators send fp32; fp32 operators
// send bf16 rather than the reference's fp16: same 16 bits, but fp32's exponent range, so the ever smaller residual vectors of a
// restarted solve keep their relative precision instead of sinking into fp16's subnormals (the reference leaves its fp16 form disabled).
template <class T> struct LowpVec;
template <> struct LowpVec<float> { using type = uint2; };     // 4 x bf16
template <> struct LowpVec<double> { using type = float2; };   // one complex fp32
__device__ __forceinline__ uint32_t f32_to_bf16(float x) {     // round to nearest even; NaN stays NaN
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (u >> 16) | 0x40u;
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) { return f32_to_bf16(lo) | (f32_to_bf16(hi) << 16); }
__device__ __forceinline__ void load_half_lowp(HalfReg<float> &h, const uint2 *p) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint2 v = p[k << LOGW];
    h.re[2 * k] = __uint_as_float(v.x << 16); h.im[2 * k] = __uint_as_float(v.x & 0xFFFF0000u);
    h.re[2 * k + 1] = __uint_as_float(v.y << 16); h.im[2 * k + 1] = __uint_as_float(v.y & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void load_half_lowp(HalfReg<double> &h, const float2 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { const float2 v = p[k << LOGW]; h.re[k] = v.x; h.im[k] = v.y; }
}
__device__ __forceinline__ void store_half_lowp(const HalfReg<float> &h, uint2 *p) {
#pragma unroll
  for (int k = 0; k < 3; k++) p[k << LOGW] = make_uint2(bf16x2(h.re[2 * k], h.im[2 * k]), bf16x2(h.re[2 * k + 1], h.im[2 * k + 1]));
}
__device__ __forceinline__ void store_half_lowp(const HalfReg<double> &h, float2 *p) {
#pragma unroll
  for (int k = 0; k < 6; k++) p[k << LOGW] = make_float2((float)h.re[k], (float)h.im[k]);
}
// stored link: fp32 5 x float4 (18 reals + 2 pad), fp64 9 x double2
__device__ __forceinline__ void load_link(LinkReg<float> &u, const float4 *__restrict__ p) {
  float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3), v4 = __ldg(p + 4);
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w;
  u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w;
  u.re[6] = v3.x; u.im[6] = v3.y; u.re[7] = v3.z; u.im[7] = v3.w;
  u.re[8] = v4.x; u.im[8] = v4.y;
}
__device__ __forceinline__ void load_link(LinkReg<double> &u, const double2 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 9; k++) { double2 v = __ldg(p + k); u.re[k] = v.x; u.im[k] = v.y; }
}

// 12-real links: rows 0 and 1 stored (fp32 3 x float4, fp64 6 x double2), row 2 = conj(row0 x row1) for a special unitary matrix
template <class T> struct Recon12 { static constexpr int LV = sizeof(T) == 4 ? 3 : 6; };
template <class T> __device__ __forceinline__ void recon_row2(LinkReg<T> &u) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
    // (row0 x row1)_c = u0[c1] u1[c2] - u0[c2] u1[c1], conjugated
    const T re = (u.re[c1] * u.re[3 + c2] - u.im[c1] * u.im[3 + c2]) - (u.re[c2] * u.re[3 + c1] - u.im[c2] * u.im[3 + c1]);
    const T im = (u.re[c1] * u.im[3 + c2] + u.im[c1] * u.re[3 + c2]) - (u.re[c2] * u.im[3 + c1] + u.im[c2] * u.re[3 + c1]);
    u.re[6 + c] = re; u.im[6 + c] = -im;
  }
}
__device__ __forceinline__ void load_link12(LinkReg<float> &u, const float4 *__restrict__ p) {
  const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
  u.re[0] = v0.x; u.im[0] = v0.y; u.re[1] = v0.z; u.im[1] = v0.w;
  u.re[2] = v1.x; u.im[2] = v1.y; u.re[3] = v1.z; u.im[3] = v1.w;
  u.re[4] = v2.x; u.im[4] = v2.y; u.re[5] = v2.z; u.im[5] = v2.w;
  recon_row2(u);
}
__device__ __forceinline__ void load_link12(LinkReg<double> &u, const double2 *__restrict__ p) {
#pragma unroll
  for (int k = 0; k < 6; k++) { const double2 v = __ldg(p + k); u.re[k] = v.x; u.im[k] = v.y; }
  recon_row2(u);
}
template <class T> __device__ __forceinline__ void scale_half(HalfReg<T> &h, T cr, T ci) {
#pragma unroll
  for (int k = 0; k < 6; k++) { const T re = h.re[k], im = h.im[k]; h.re[k] = cr * re - ci * im; h.im[k] = cr * im + ci * re; }
}

// ------------------------------------------------------------------ spin projection: h = (1 + SIGN*gamma_MU) f, upper two components
// ref: Grid/qcd/spin/TwoSpinor.h:75-133
template <int MU, int SIGN, class T> __device__ __forceinline__ void sp_proj(HalfReg<T> &h, const SpinorReg<T> &f) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const T f0r = f.re[c], f0i = f.im[c], f1r = f.re[3 + c], f1i = f.im[3 + c];
    const T f2r = f.re[6 + c], f2i = f.im[6 + c], f3r = f.re[9 + c], f3i = f.im[9 + c];
    if (MU == 0) {          // h0 = f0 +- i f3 ; h1 = f1 +- i f2
      if (SIGN > 0) { h.re[c] = f0r - f3i; h.im[c] = f0i + f3r; h.re[3 + c] = f1r - f2i; h.im[3 + c] = f1i + f2r; }
      else          { h.re[c] = f0r + f3i; h.im[c] = f0i - f3r; h.re[3 + c] = f1r + f2i; h.im[3 + c] = f1i - f2r; }
    } else if (MU == 1) {   // h0 = f0 -+ f3 ; h1 = f1 +- f2
      if (SIGN > 0) { h.re[c] = f0r - f3r; h.im[c] = f0i - f3i; h.re[3 + c] = f1r + f2r; h.im[3 + c] = f1i + f2i; }
      else          { h.re[c] = f0r + f3r; h.im[c] = f0i + f3i; h.re[3 + c] = f1r - f2r; h.im[3 + c] = f1i - f2i; }
    } else if (MU == 2) {   // h0 = f0 +- i f2 ; h1 = f1 -+ i f3
      if (SIGN > 0) { h.re[c] = f0r - f2i; h.im[c] = f0i + f2r; h.re[3 + c] = f1r + f3i; h.im[3 + c] = f1i - f3r; }
      else          { h.re[c] = f0r + f2i; h.im[c] = f0i - f2r; h.re[3 + c] = f1r - f3i; h.im[3 + c] = f1i + f3r; }
    } else {                // h0 = f0 +- f2 ; h1 = f1 +- f3
      if (SIGN > 0) { h.re[c] = f0r + f2r; h.im[c] = f0i + f2i; h.re[3 + c] = f1r + f3r; h.im[3 + c] = f1i + f3i; }
      else          { h.re[c] = f0r - f2r; h.im[c] = f0i - f2i; h.re[3 + c] = f1r - f3r; h.im[3 + c] = f1i - f3i; }
    }
  }
}

// (U h)_row = sum_col U[row][col] h_col  for both half-spin components. ref: WilsonImpl.h:84-91
template <class T> __device__ __forceinline__ void mult_link(HalfReg<T> &o, const LinkReg<T> &u, const HalfReg<T> &h) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      T re = u.re[3 * r] * h.re[3 * s];
      T im = u.re[3 * r] * h.im[3 * s];
      re = fma(-u.im[3 * r], h.im[3 * s], re);
      im = fma(u.im[3 * r], h.re[3 * s], im);
#pragma unroll
      for (int c = 1; c < 3; c++) {
        re = fma(u.re[3 * r + c], h.re[3 * s + c], re);
        im = fma(u.re[3 * r + c], h.im[3 * s + c], im);
        re = fma(-u.im[3 * r + c], h.im[3 * s + c], re);
        im = fma(u.im[3 * r + c], h.re[3 * s + c], im);
      }
      o.re[3 * s + r] = re; o.im[3 * s + r] = im;
    }
}

// result += Recon_{MU,SIGN}(h).  ref: Grid/qcd/spin/TwoSpinor.h:193-354 (accumRecon*)
template <int MU, int SIGN, class T> __device__ __forceinline__ void accum_recon(SpinorReg<T> &f, const HalfReg<T> &h) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const T h0r = h.re[c], h0i = h.im[c], h1r = h.re[3 + c], h1i = h.im[3 + c];
    f.re[c] += h0r; f.im[c] += h0i; f.re[3 + c] += h1r; f.im[3 + c] += h1i;
    if (MU == 0) {        // Xp: f2 -= i h1, f3 -= i h0
      if (SIGN > 0) { f.re[6 + c] += h1i; f.im[6 + c] -= h1r; f.re[9 + c] += h0i; f.im[9 + c] -= h0r; }
      else          { f.re[6 + c] -= h1i; f.im[6 + c] += h1r; f.re[9 + c] -= h0i; f.im[9 + c] += h0r; }
    } else if (MU == 1) { // Yp: f2 += h1, f3 -= h0
      if (SIGN > 0) { f.re[6 + c] += h1r; f.im[6 + c] += h1i; f.re[9 + c] -= h0r; f.im[9 + c] -= h0i; }
      else          { f.re[6 + c] -= h1r; f.im[6 + c] -= h1i; f.re[9 + c] += h0r; f.im[9 + c] += h0i; }
    } else if (MU == 2) { // Zp: f2 -= i h0, f3 += i h1
      if (SIGN > 0) { f.re[6 + c] += h0i; f.im[6 + c] -= h0r; f.re[9 + c] -= h1i; f.im[9 + c] += h1r; }
      else          { f.re[6 + c] -= h0i; f.im[6 + c] += h0r; f.re[9 + c] += h1i; f.im[9 + c] -= h1r; }
    } else {              // Tp: f2 += h0, f3 += h1
      if (SIGN > 0) { f.re[6 + c] += h0r; f.im[6 + c] += h0i; f.re[9 + c] += h1r; f.im[9 + c] += h1i; }
      else          { f.re[6 + c] -= h0r; f.im[6 + c] -= h0i; f.re[9 + c] -= h1r; f.im[9 + c] -= h1i; }
    }
  }
}

} // namespace gb
