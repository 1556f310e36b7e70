// kernels_common.cuh -- small device helpers shared by the kernels
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace gb {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// uniform double in [0,1) from 53 random bits
__host__ __device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

// ---- 16-byte vector arithmetic (float4 = two complex fp32, double2 = one complex fp64)
__device__ __forceinline__ float4 vscale(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ double2 vscale(double a, double2 x) { return make_double2(a * x.x, a * x.y); }
__device__ __forceinline__ float4 vaxpy(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ double2 vaxpy(double a, double2 x, double2 y) { return make_double2(fma(a, x.x, y.x), fma(a, x.y, y.y)); }
__device__ __forceinline__ float4 vaxpby(float a, float4 x, float b, float4 y) {
  return make_float4(fmaf(a, x.x, b * y.x), fmaf(a, x.y, b * y.y), fmaf(a, x.z, b * y.z), fmaf(a, x.w, b * y.w));
}
__device__ __forceinline__ double2 vaxpby(double a, double2 x, double b, double2 y) {
  return make_double2(fma(a, x.x, b * y.x), fma(a, x.y, b * y.y));
}
__device__ __forceinline__ float4 vadd(float4 x, float4 y) { return make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w); }
__device__ __forceinline__ double2 vadd(double2 x, double2 y) { return make_double2(x.x + y.x, x.y + y.y); }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ double2 vzero(double2) { return make_double2(0., 0.); }
// site products in working precision, promoted to double for the lattice sum (ref: Lattice_reduction.h:276-283)
__device__ __forceinline__ double vnorm2(float4 x) { return (double)(x.x * x.x + x.y * x.y) + (double)(x.z * x.z + x.w * x.w); }
__device__ __forceinline__ double vnorm2(double2 x) { return x.x * x.x + x.y * x.y; }
__device__ __forceinline__ void vinner(float4 l, float4 r, double &re, double &im) { // conj(l)*r
  re += (double)(l.x * r.x + l.y * r.y) + (double)(l.z * r.z + l.w * r.w);
  im += (double)(l.x * r.y - l.y * r.x) + (double)(l.z * r.w - l.w * r.z);
}
__device__ __forceinline__ void vinner(double2 l, double2 r, double &re, double &im) {
  re += l.x * r.x + l.y * r.y;
  im += l.x * r.y - l.y * r.x;
}

} // namespace gb
