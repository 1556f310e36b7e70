// halo_p2p.cu -- halo exchange by direct peer stores over NVLink (no NCCL on the data path).
//
// Replaces the reference's intra-node path -- gather kernel, then cudaMemcpyAsync into the neighbour's cudaIpc-mapped
// receive buffer, then MPI_Barrier (ref: Grid/communicator/Communicator_mpi3.cc:437-465 ; SharedMemoryMPI.cc:598-680 ;
// Grid/stencil/Stencil.h:367-430,453-511) -- with ONE kernel that projects the boundary slices and writes the half
// spinors straight into the neighbour's receive buffer through its IPC mapping, then publishes an epoch flag there with
// system-scope release semantics.  The consumer (the exterior pass of the hopping kernel) acquires the flag on the
// device, so there is no host synchronisation, no separate copy and no SM-hungry collective kernel competing with the
// interior pass.  Receive buffers are double buffered by epoch: a rank can only start packing hop e after it finished
// hop e-1, which needed every neighbour's epoch e-1 flag, which those neighbours wrote after finishing hop e-2 -- so
// buffer (e mod 2) is never overwritten while its previous contents (epoch e-2) are still being read.
#include "dhop_kernel.cuh"
#include "comm.hpp"
#include "fermop.hpp"
#include <cstring>

namespace gb {

struct PackItem {
  const void *src;   // parity block being packed
  void *dst;         // peer-mapped destination (this epoch, this point, this parity slot)
  uint32_t nface;    // face sites (cb)
  int mu, fwd, ip;
  int zedge;         // t faces only: pack just the planes z = 0 and z = Lz-1 (the rest of the face is sent by the hop kernel itself)
};
struct PackSendArgs {
  PackItem item[16];
  int nitems;
  int Ls, Lx, Lxh, Ly, Lz, Lt, origin_parity;
  int lowp;                       // compressed halos (dhop_kernel.cuh store_half_lowp): half the bytes over the link
  unsigned int *counter;          // last-CTA detection
  unsigned long long *flag[8];    // peer-mapped flag slot per point (nullptr if unused)
  unsigned long long epoch;
};

template <class T, int DAG, int MU, int FWD>
__device__ __forceinline__ void pack_body(const PackSendArgs &a, const PackItem &it, uint32_t q) {
  using P = Prec<T>;
  using V = typename P::vec;
  const uint32_t fi = q / a.Ls, s = q - fi * a.Ls;
  int xh, y, z, t;
  const int slice = FWD ? 0 : (MU == 0 ? a.Lx : MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt) - 1;
  uint32_t r = fi;
  if (MU == 0) {
    int yhalf = r % (a.Ly >> 1); r /= (a.Ly >> 1); z = r % a.Lz; t = r / a.Lz;
    int ypar = (slice + it.ip + a.origin_parity + z + t) & 1;
    y = 2 * yhalf + ypar; xh = slice >> 1;
  } else if (MU == 1) { xh = r % a.Lxh; r /= a.Lxh; z = r % a.Lz; t = r / a.Lz; y = slice; }
  else if (MU == 2) { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; t = r / a.Ly; z = slice; }
  else { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; z = r / a.Ly; t = slice; if (it.zedge) z = z ? a.Lz - 1 : 0; }
  const uint32_t site = xh + a.Lxh * (y + a.Ly * (z + a.Lz * t));
  const uint32_t i = site * a.Ls + s;
  if (MU == 3 && it.zedge) q = (xh + a.Lxh * (y + a.Ly * z)) * a.Ls + s;    // position inside the whole face
  SpinorReg<T> f;
  load_spinor(f, (const V *)it.src + ((size_t)(i >> LOGW) * P::NV << LOGW) + (i & (W - 1)));
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfReg<T> h;
  sp_proj<MU, SIGN>(h, f);
  const size_t vi = ((size_t)(q >> LOGW) * (P::NV / 2) << LOGW) + (q & (W - 1));
  if (a.lowp) store_half_lowp(h, (typename LowpVec<T>::type *)it.dst + vi);
  else store_half(h, (V *)it.dst + vi);
}

template <class T, int DAG> __global__ void __launch_bounds__(256) pack_send_kernel(const PackSendArgs a) {
  const PackItem &it = a.item[blockIdx.y];
  // grid-stride: the launch may cover a face with fewer CTAs than it has 256-thread tiles (persistent form, GB_PACK_CTAS), so
  // that the pack kernel occupies a few SMs for the duration of the transfer instead of sweeping over all of them
  const uint32_t nq = it.nface * (uint32_t)a.Ls;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    switch (it.mu * 2 + it.fwd) {
    case 0: pack_body<T, DAG, 0, 0>(a, it, q); break;
    case 1: pack_body<T, DAG, 0, 1>(a, it, q); break;
    case 2: pack_body<T, DAG, 1, 0>(a, it, q); break;
    case 3: pack_body<T, DAG, 1, 1>(a, it, q); break;
    case 4: pack_body<T, DAG, 2, 0>(a, it, q); break;
    case 5: pack_body<T, DAG, 2, 1>(a, it, q); break;
    case 6: pack_body<T, DAG, 3, 0>(a, it, q); break;
    default: pack_body<T, DAG, 3, 1>(a, it, q); break;
    }
  }
  // publish: every CTA fences its peer stores, the last one to finish writes the epoch flags into the peers' memory
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int prev = atomicAdd(a.counter, 1u);
    if (prev == total - 1) {
      *a.counter = 0;
      __threadfence_system();
#pragma unroll
      for (int p = 0; p < 8; p++)
        if (a.flag[p]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[p]), "l"(a.epoch) : "memory");
    }
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// one-time: allocate the receive buffers, exchange IPC handles through the NCCL communicator, map the neighbours
bool p2p_setup(gb_fermop *op) {
  P2PState &S = op->p2p;
  if (S.tried) return S.ok;
  S.tried = true;
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  const int hv = nv_of(op->prec) / 2;
  // a local failure or an opt-out (GB_NO_P2P) is folded into `ok`: every rank still takes part in the AllGather and the vote
  // below, so that a rank that cannot map its neighbours makes ALL ranks fall back to NCCL instead of hanging them
  bool ok = getenv("GB_NO_P2P") == nullptr;
  size_t off = 0;
  for (int p = 0; p < 8; p++) { S.pt_off[p] = 0; S.peer_base[p] = nullptr; }
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    const size_t nface5 = (size_t)(g->V4cb / g->ldims[mu]) * op->Ls;
    const size_t blocks = (nface5 + W - 1) / W;
    op->halo_parity_stride[mu] = blocks * hv * W;
    const size_t bytes = align_up(2 * op->halo_parity_stride[mu] * 16, 256);
    S.pt_off[mu] = off; off += bytes;
    S.pt_off[mu + 4] = off; off += bytes;
  }
  S.epoch_stride = off;
  S.flags_off = 2 * off;
  S.recv_bytes = 2 * off + 2 * 8 * sizeof(unsigned long long);
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok) {
    GB_CUDA(cudaMalloc(&S.recv_base, S.recv_bytes));
    GB_CUDA(cudaMemset(S.recv_base, 0, S.recv_bytes));
    GB_CUDA(cudaMalloc(&S.d_counter, 2 * sizeof(unsigned int)));     // [0] pack kernel, [1] sender CTAs of the hop kernel
    GB_CUDA(cudaMemset(S.d_counter, 0, 2 * sizeof(unsigned int)));
    GB_CUDA(cudaDeviceSynchronize());
  }
  // does any halo leave this rank?  (GB_SELF_HALO routes undecomposed dimensions through the same path: the "peer" is this rank)
  bool any_remote = false;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1)
    if (g->nbr_rank[mu][0] != ctx->rank || g->nbr_rank[mu][1] != ctx->rank) any_remote = true;
  const int n = ctx->nranks;
  std::vector<cudaIpcMemHandle_t> all(n);
  if (n > 1) {
    // exchange handles (every rank of the communicator, whether or not its own halos are remote)
    if (ok && any_remote && cudaIpcGetMemHandle(&mine, S.recv_base) != cudaSuccess) { cudaGetLastError(); ok = false; }
    char *d_all = nullptr;
    GB_CUDA(cudaMalloc(&d_all, (size_t)n * sizeof(mine)));
    GB_CUDA(cudaMemcpy(d_all + (size_t)ctx->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NcclApi &N = nccl();
    GB_REQUIRE(N.AllGather != nullptr, "ncclAllGather missing");
    nccl_check(N.AllGather(d_all + (size_t)ctx->rank * sizeof(mine), d_all, sizeof(mine), ncclChar, ctx->nccl, ctx->stream), "ncclAllGather");
    GB_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)n * sizeof(mine), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(cudaStreamSynchronize(ctx->stream));
    GB_CUDA(cudaFree(d_all));
  }
  // map each distinct neighbour once
  std::vector<void *> mapped(n, nullptr);
  for (int mu = 0; mu < 4 && ok; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    // receiver of my data for ITS point mu (its forward leg) is my backward neighbour; for point mu+4 my forward neighbour
    const int dest[2] = {g->nbr_rank[mu][1], g->nbr_rank[mu][0]};
    for (int k = 0; k < 2; k++) {
      const int r = dest[k];
      if (r == ctx->rank) mapped[r] = S.recv_base;          // halo to self: plain device pointer
      if (!mapped[r]) {
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        mapped[r] = ptr;
        S.opened.push_back(ptr);
      }
      S.peer_base[k == 0 ? mu : mu + 4] = mapped[r];
    }
  }
  // every rank must agree, otherwise some would wait for flags that never come
  double v = ok ? 0.0 : 1.0;
  global_sum(ctx, &v, 1);
  S.ok = (v == 0.0);
  if (!S.ok) p2p_teardown(op);
  op->halo_ready = true;
  return S.ok;
}

void p2p_teardown(gb_fermop *op) {
  P2PState &S = op->p2p;
  for (void *p : S.opened) cudaIpcCloseMemHandle(p);
  S.opened.clear();
  if (S.recv_base) cudaFree(S.recv_base);
  if (S.d_counter) cudaFree(S.d_counter);
  S.recv_base = nullptr; S.d_counter = nullptr;
}

unsigned long long p2p_next_epoch(gb_fermop *op) { return ++op->p2p.epoch; }

// pack + send every face of this hop in one launch; returns the epoch the consumer must wait for
unsigned long long p2p_pack_send(gb_fermop *op, const void *const in[2], int parity_out_first, int nparity, int dag, cudaStream_t st) {
  GB_TRACE("Gather");   // pack (project) + peer stores: the reference's Gather + CommunicateBegin
  const unsigned long long epoch = p2p_next_epoch(op);
  p2p_send_only(op, epoch, in, parity_out_first, nparity, dag, st, 0);
  return epoch;
}
// hop_sends_t: the column-sweep hop that follows sends the t faces itself (dhop_col2.cuh, send_on) -- 2: all of them (its columns
// visit every plane), 1: all but the planes z = 0 and z = Lz-1 of a z-decomposed lattice (columns over planes 1 ... Lz-2), which this
// kernel then packs; the t flags are the hop's to set in both cases
void p2p_send_only(gb_fermop *op, unsigned long long epoch, const void *const in[2], int parity_out_first, int nparity, int dag, cudaStream_t st,
                   int hop_sends_t) {
  P2PState &S = op->p2p;
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  const size_t eoff = (size_t)(epoch & 1) * S.epoch_stride;
  PackSendArgs a;
  std::memset(&a, 0, sizeof(a));
  a.Ls = op->Ls; a.Lx = g->ldims[0]; a.Lxh = a.Lx / 2; a.Ly = g->ldims[1]; a.Lz = g->ldims[2]; a.Lt = g->ldims[3];
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  a.counter = S.d_counter;
  a.epoch = epoch;
  a.lowp = op->halo_lowp;
  uint32_t maxn = 0;
  const bool z_comm = (op->comm_dim_mask >> 2) & 1;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    const bool hop_face = hop_sends_t != 0 && mu == 3;
    if (hop_face && (!z_comm || hop_sends_t == 2)) continue;
    const uint32_t nface = hop_face ? (uint32_t)(2 * (g->ldims[0] / 2) * g->ldims[1]) : (uint32_t)(g->V4cb / g->ldims[mu]);
    maxn = std::max(maxn, nface * (uint32_t)op->Ls);
    for (int fwd = 0; fwd < 2; fwd++) {
      const int point = fwd ? mu : mu + 4;
      if (!hop_face) a.flag[point] = (unsigned long long *)((char *)S.peer_base[point] + S.flags_off) + (epoch & 1) * 8 + point;
      for (int j = 0; j < nparity; j++) {
        const int po = parity_out_first ^ j, ip = 1 - po;
        const int slot = nparity == 1 ? 0 : ip;
        PackItem &it = a.item[a.nitems++];
        it.src = in[ip];
        it.dst = (char *)S.peer_base[point] + eoff + S.pt_off[point] + (size_t)slot * op->halo_parity_stride[mu] * 16;
        it.nface = nface; it.mu = mu; it.fwd = fwd; it.ip = ip; it.zedge = hop_face ? 1 : 0;
      }
    }
  }
  // GB_PACK_CTAS=<n>: at most n CTAs per face item (grid-stride loop inside), 0 = one CTA per 256 face elements
  static const unsigned pack_ctas = getenv("GB_PACK_CTAS") ? (unsigned)atoi(getenv("GB_PACK_CTAS")) : 0u;
  if (a.nitems == 0) return;                     // every face of this hop is sent by the hop kernel
  unsigned gx = (maxn + 255) / 256;
  if (pack_ctas > 0 && gx > pack_ctas) gx = pack_ctas;
  dim3 grid(gx, a.nitems);
  if (op->prec == GB_F32) { if (dag) pack_send_kernel<float, 1><<<grid, 256, 0, st>>>(a); else pack_send_kernel<float, 0><<<grid, 256, 0, st>>>(a); }
  else { if (dag) pack_send_kernel<double, 1><<<grid, 256, 0, st>>>(a); else pack_send_kernel<double, 0><<<grid, 256, 0, st>>>(a); }
  count_launch(ctx);
  check_launch(ctx, "pack_send");
}

// sender-side pointers of this epoch for a hop that sends its own t faces: destination buffers and flags in the neighbours' memory
void p2p_fill_send_t(gb_fermop *op, unsigned long long epoch, void *dst[2], unsigned long long *flag[2], unsigned int **counter) {
  P2PState &S = op->p2p;
  const size_t eoff = (size_t)(epoch & 1) * S.epoch_stride;
  const int points[2] = {3, 7};   // plane 0 feeds the backward neighbour's forward t leg (point 3), plane Lt-1 the forward neighbour's point 7
  for (int k = 0; k < 2; k++) {
    dst[k] = (char *)S.peer_base[points[k]] + eoff + S.pt_off[points[k]];
    flag[k] = (unsigned long long *)((char *)S.peer_base[points[k]] + S.flags_off) + (epoch & 1) * 8 + points[k];
  }
  *counter = S.d_counter + 1;
}
// receive-side pointers of this epoch for the hop kernels
void p2p_fill_halo(gb_fermop *op, unsigned long long epoch, const void *halo[8], const unsigned long long **flags) {
  P2PState &S = op->p2p;
  const size_t eoff = (size_t)(epoch & 1) * S.epoch_stride;
  for (int p = 0; p < 8; p++) halo[p] = ((op->comm_dim_mask >> (p & 3)) & 1) ? (const char *)S.recv_base + eoff + S.pt_off[p] : nullptr;
  *flags = (const unsigned long long *)((const char *)S.recv_base + S.flags_off) + (epoch & 1) * 8;
}

} // namespace gb
