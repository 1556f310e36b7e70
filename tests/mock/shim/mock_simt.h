// mock_simt.h -- TEST-ONLY cooperative execution of one CUDA thread block on the CPU (tests/mock/README.md).
//
// Kernels whose threads talk to each other (shared memory + __syncthreads, warp shuffles, mbarriers fed by TMA bulk copies or
// cp.async) cannot be run as a plain loop over threads.  Here every CUDA thread of a block is a fibre (ucontext) on the calling
// OS thread; barriers and mbarrier waits hand control to the next fibre, asynchronous copies complete at once.  This checks index
// arithmetic, shared-memory layout and barrier protocol (a missing arrival or a wrong phase parity hangs and is reported), not
// memory ordering: fibres never run concurrently.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>

namespace gb_mock {

void coop_launch(unsigned gx, unsigned gy, unsigned nthreads, size_t smem_bytes, const std::function<void()> &body);
void count_coop_launch(const char *kernel);   // read back through gb_mock_coop_launches(kernel) (extern "C", for the tests)
unsigned char *dynamic_smem();
void sync_block();
int sync_block_or(int pred);
void sync_warp();
void yield();
uint64_t shfl_raw(uint64_t v, int src_lane_delta);
uint64_t shfl_idx(uint64_t v, int src_lane);          // value held by lane src_lane of the caller's warp   // value held by lane + delta of the caller's warp (own value if out of range)

// ---- mbarrier (transaction barrier in shared memory; state kept in the 8 bytes the kernel declares)
void mbar_init(uint64_t *bar, int count);
void mbar_expect_tx(uint64_t *bar, uint32_t bytes);               // one arrival + bytes expected
void mbar_complete_tx(uint64_t *bar, uint32_t bytes);
void mbar_arrive(uint64_t *bar);
void mbar_wait(uint64_t *bar, uint32_t parity);
inline void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) { std::memcpy(smem_dst, gsrc, bytes); mbar_complete_tx(bar, bytes); }
inline void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t) { bulk_g2s(smem_dst, gsrc, bytes, bar); }
inline void cp_async16(void *smem_dst, const void *gsrc) { std::memcpy(smem_dst, gsrc, 16); }
inline void cp_async_arrive(uint64_t *bar) { mbar_arrive(bar); }   // the copies of this thread have already landed

// ---- packed f32x2 arithmetic (PTX fma.rn.f32x2 and friends): lo = bits 0..31, hi = bits 32..63
typedef unsigned long long f2;
inline f2 pk(float lo, float hi) { uint32_t a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4); return (f2)a | ((f2)b << 32); }
inline void upk(f2 d, float &lo, float &hi) { uint32_t a = (uint32_t)d, b = (uint32_t)(d >> 32); std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4); }
inline f2 fma2(f2 a, f2 b, f2 c) { float al, ah, bl, bh, cl, ch; upk(a, al, ah); upk(b, bl, bh); upk(c, cl, ch); return pk(std::fmaf(al, bl, cl), std::fmaf(ah, bh, ch)); }
inline f2 mul2(f2 a, f2 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al * bl, ah * bh); }
inline f2 add2(f2 a, f2 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al + bl, ah + bh); }
inline f2 sub2(f2 a, f2 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al - bl, ah - bh); }

} // namespace gb_mock

inline void __syncthreads() { gb_mock::sync_block(); }
inline int __syncthreads_or(int p) { return gb_mock::sync_block_or(p); }
inline void __syncwarp(unsigned = 0xffffffffu) { gb_mock::sync_warp(); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
  raw = gb_mock::shfl_raw(raw, delta);
  T r; std::memcpy(&r, &raw, sizeof(T)); return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
  raw = gb_mock::shfl_idx(raw, src_lane);
  T r; std::memcpy(&r, &raw, sizeof(T)); return r;
}
inline void __threadfence_block() {}
