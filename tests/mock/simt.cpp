// simt.cpp -- TEST-ONLY: the fibre scheduler behind tests/mock/shim/mock_simt.h (one CUDA block at a time per OS thread).
#include "cuda_runtime.h"
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <ucontext.h>
#include <vector>

namespace gb_mock {

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;

struct Fibre {
  ucontext_t ctx;
  unsigned char *stack = nullptr;
  bool done = true;
};
struct Engine {
  ucontext_t main_ctx;
  std::vector<Fibre> fibres;
  std::vector<unsigned char> smem;
  std::vector<uint64_t> shfl_slot;
  const std::function<void()> *body = nullptr;
  unsigned nthreads = 0, current = 0, alive = 0;
  // block barrier
  unsigned bar_count = 0, bar_gen = 0; int bar_or = 0, bar_or_result = 0;
  // warp barriers
  std::vector<unsigned> warp_count, warp_gen, warp_alive;
  uint64_t progress = 0;
  bool active = false;
  std::exception_ptr error;
  ~Engine() { for (Fibre &f : fibres) std::free(f.stack); }
};
thread_local Engine E;

void trampoline() {
  try { (*E.body)(); } catch (...) { E.error = std::current_exception(); }
  Fibre &f = E.fibres[E.current];
  f.done = true;
  E.alive--; E.warp_alive[E.current / 32]--;
  E.progress++;
  swapcontext(&f.ctx, &E.main_ctx);
}
void release_block_if_complete() {
  if (E.alive && E.bar_count >= E.alive) { E.bar_count = 0; E.bar_gen++; E.bar_or_result = E.bar_or; E.bar_or = 0; E.progress++; }
}
void release_warp_if_complete(unsigned w) {
  if (E.warp_alive[w] && E.warp_count[w] >= E.warp_alive[w]) { E.warp_count[w] = 0; E.warp_gen[w]++; E.progress++; }
}
struct MBar { uint16_t count, pending; int32_t tx : 31; uint32_t phase : 1; };
static_assert(sizeof(MBar) == 8, "mbarrier state lives in the kernel's own 8 bytes");
void mbar_check(MBar *b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->count; E.progress++; }
}
} // namespace

static std::mutex g_count_mutex;
static std::map<std::string, long> g_coop_counts;
void count_coop_launch(const char *kernel) { std::lock_guard<std::mutex> l(g_count_mutex); g_coop_counts[kernel]++; }

void yield() {
  if (!E.active) return;   // not inside a cooperative launch: nothing to hand over to
  swapcontext(&E.fibres[E.current].ctx, &E.main_ctx);
}
unsigned char *dynamic_smem() { return E.smem.data(); }

void sync_block() { sync_block_or(0); }
int sync_block_or(int pred) {
  if (!E.active) return pred;
  const unsigned gen = E.bar_gen;
  E.bar_or |= pred != 0;
  E.bar_count++;
  release_block_if_complete();
  while (E.bar_gen == gen) yield();
  return E.bar_or_result;
}
void sync_warp() {
  if (!E.active) return;
  const unsigned w = E.current / 32, gen = E.warp_gen[w];
  E.warp_count[w]++;
  release_warp_if_complete(w);
  while (E.warp_gen[w] == gen) yield();
}
uint64_t shfl_raw(uint64_t v, int delta) {
  if (!E.active) return v;
  const unsigned me = E.current, lane = me % 32;
  E.shfl_slot[me] = v;
  sync_warp();
  const int src = (int)lane + delta;
  const unsigned from = me - lane + (unsigned)src;
  const uint64_t r = (src >= 0 && src < 32 && from < E.nthreads) ? E.shfl_slot[from] : v;
  sync_warp();
  return r;
}

uint64_t shfl_idx(uint64_t v, int src_lane) {
  if (!E.active) return v;
  return shfl_raw(v, src_lane - (int)(E.current % 32));
}

void mbar_init(uint64_t *bar, int count) { MBar *b = (MBar *)bar; b->count = (uint16_t)count; b->pending = (uint16_t)count; b->tx = 0; b->phase = 0; }
void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { MBar *b = (MBar *)bar; b->tx += (int32_t)bytes; b->pending--; mbar_check(b); }
void mbar_complete_tx(uint64_t *bar, uint32_t bytes) { MBar *b = (MBar *)bar; b->tx -= (int32_t)bytes; mbar_check(b); }
void mbar_arrive(uint64_t *bar) { MBar *b = (MBar *)bar; b->pending--; mbar_check(b); }
void mbar_wait(uint64_t *bar, uint32_t parity) {
  MBar *b = (MBar *)bar;
  while (b->phase == (parity & 1u)) {
    if (!E.active) throw std::runtime_error("mock: mbarrier wait outside a cooperative launch would never return");
    yield();
  }
}

void coop_launch(unsigned gx, unsigned gy, unsigned nthreads, size_t smem_bytes, const std::function<void()> &body) {
  if (E.active) throw std::runtime_error("mock: nested cooperative launch");
  if (E.fibres.size() < nthreads) E.fibres.resize(nthreads);
  for (unsigned i = 0; i < nthreads; i++) if (!E.fibres[i].stack) {
    E.fibres[i].stack = (unsigned char *)std::malloc(STACK_BYTES);
    if (!E.fibres[i].stack) throw std::bad_alloc();
  }
  const unsigned nwarps = (nthreads + 31) / 32;
  E.smem.assign(smem_bytes + 64, 0);
  E.shfl_slot.assign(nthreads, 0);
  E.body = &body; E.nthreads = nthreads; E.error = nullptr;
  t_gridDim = dim3(gx, gy, 1); t_blockDim = dim3(nthreads, 1, 1);
  struct Guard { ~Guard() { E.active = false; } } guard;
  for (unsigned by = 0; by < gy; by++) for (unsigned bx = 0; bx < gx; bx++) {
    E.alive = nthreads; E.bar_count = 0; E.bar_or = 0;
    E.warp_count.assign(nwarps, 0); E.warp_gen.assign(nwarps, 0); E.warp_alive.assign(nwarps, 0);
    for (unsigned i = 0; i < nthreads; i++) {
      Fibre &f = E.fibres[i];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, trampoline, 0);
      f.done = false;
      E.warp_alive[i / 32]++;
    }
    E.active = true;
    t_blockIdx = {bx, by, 0};
    auto last_progress = std::chrono::steady_clock::now();
    uint64_t seen = E.progress;
    while (E.alive) {
      for (unsigned i = 0; i < nthreads; i++) {
        Fibre &f = E.fibres[i];
        if (f.done) continue;
        E.current = i; t_threadIdx = {i, 0, 0};
        swapcontext(&E.main_ctx, &f.ctx);
        if (f.done) { release_block_if_complete(); release_warp_if_complete(i / 32); }
      }
      if (E.error) { E.active = false; std::rethrow_exception(E.error); }
      // a sweep in which nothing moved: either a wait on another rank's flag (fine, for a while) or a deadlock
      if (E.progress != seen) { seen = E.progress; last_progress = std::chrono::steady_clock::now(); }
      else if (std::chrono::steady_clock::now() - last_progress > std::chrono::seconds(60)) {
        std::fprintf(stderr, "mock: block (%u,%u) made no progress for 60 s: %u threads alive, %u at the block barrier -- deadlock in the kernel's barrier protocol\n",
                     bx, by, E.alive, E.bar_count);
        std::abort();
      }
    }
    E.active = false;
  }
}

} // namespace gb_mock

// how many cooperative launches this process has made of kernels whose launch expression (as written in the source, e.g.
// "dhop_fast_kernel<LS, 0, 2>") contains `kernel`: the tests use it to prove that the tuned kernels, not the generic fall-back,
// produced the numbers they compared
extern "C" long gb_mock_coop_launches(const char *kernel) {
  std::lock_guard<std::mutex> l(gb_mock::g_count_mutex);
  long n = 0;
  for (const auto &kv : gb_mock::g_coop_counts) if (kv.first.find(kernel) != std::string::npos) n += kv.second;
  return n;
}
