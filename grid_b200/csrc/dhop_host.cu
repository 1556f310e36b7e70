// dhop_host.cu -- Dhop on HOST-resident fields (gb_op_dhop_host), pipelined over t-slices.
//
// This is the call a reference-side binding whose Lattice objects live in host memory makes (ref: FermionOperator::Dhop(in,out,dag),
// Grid/qcd/action/fermion/FermionOperator.h:75-77, with in/out unvectorised by unvectorizeToLexOrdArray).  H2D of slice t+1, the hop
// of slice t and D2H of slice t-1 run concurrently (PCIe is full duplex), so a host-to-host Dhop costs one direction of PCIe traffic
// instead of import + hop + export back to back.
//
// Decomposed lattices (z and / or t split over ranks; each rank passes its LOCAL volume) are pipelined too: the sites the neighbours
// need -- t-slices 0 and Lt-1, and on z splits the planes z = 0 and z = Lz-1 of every other slice, fetched with one strided copy
// per face -- are imported first, the halo exchange of the whole hop (project + send + arrival) runs once on them, and the slices
// then stream through as on a single rank, every slab hop reading the receive buffers for the legs that leave the rank.
// (x / y splits would need a strided gather per lattice row: they keep the import + hop + export form.)
#include "internal.hpp"
#include "fermop.hpp"
#include "kernels_common.cuh"
#include <map>
#include <mutex>
#include <vector>

using namespace gb;

namespace {
struct SlabGeom {
  int L[4];      // local dims
  int Lxh, Ls, origin_parity;
  int64_t hblk;  // blocks of W sites per parity block
};
SlabGeom slab_geom_of(const gb_fermion *f) {
  SlabGeom G;
  for (int d = 0; d < 4; d++) G.L[d] = f->grid->ldims[d];
  G.Lxh = G.L[0] / 2; G.Ls = f->Ls;
  G.origin_parity = (f->grid->origin[0] + f->grid->origin[1] + f->grid->origin[2] + f->grid->origin[3]) & 1;
  G.hblk = f->hblk;
  return G;
}
// (parity block p, cb site index) -> local lexicographic 4D index.  ref: Cartesian_red_black.h:271-286
__device__ __forceinline__ int64_t slab_cb_to_lex(const SlabGeom &G, int p, int64_t site) {
  int xh = site % G.Lxh; site /= G.Lxh;
  int y = site % G.L[1]; site /= G.L[1];
  int z = site % G.L[2];
  int t = site / G.L[2];
  int x = 2 * xh + ((p + G.origin_parity + y + z + t) & 1);
  return x + (int64_t)G.L[0] * (y + (int64_t)G.L[1] * (z + (int64_t)G.L[2] * t));
}
} // namespace

// Layout change between a staging buffer in host order and a run of whole blocks of the device field, both parities.  One thread per
// device vec element.  blockIdx.y = j selects the j-th piece of a strided set (the planes of a z face, one per t-slice): piece j
// covers the device blocks [blk0 + j * blk_stride, + nblk) of each parity block, and the host-order element with lattice index h sits
// at slab[h - (host_elem0 + j * host_stride)].  DIR = 0: staging -> field, 1: field -> staging.
template <class TD, class TH, int DIR>
__global__ void fermion_slab_transfer_kernel(typename Prec<TD>::vec *dev, TH *slab, SlabGeom G, int64_t blk0, int64_t nblk, int64_t host_elem0,
                                             int64_t blk_stride, int64_t host_stride) {
  using P = Prec<TD>;
  const int64_t nelem = 2 * nblk * P::NV * W;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  blk0 += blockIdx.y * blk_stride;
  host_elem0 += blockIdx.y * host_stride;
  const int lane = e & (W - 1);
  int64_t r = e >> LOGW;
  const int k = r % P::NV;
  int64_t lb = r / P::NV;
  const int p = lb / nblk;
  lb -= (int64_t)p * nblk;
  const int64_t i5cb = (blk0 + lb) * W + lane;
  const int64_t site = i5cb / G.Ls;
  const int s = i5cb - site * G.Ls;
  const int64_t hidx = s + (int64_t)G.Ls * slab_cb_to_lex(G, p, site) - host_elem0;
  TH *h = slab + hidx * 24;
  const int64_t de = (((int64_t)p * G.hblk + blk0 + lb) * P::NV + k) * W + lane;
  if constexpr (sizeof(TD) == 4) {
    if (DIR == 0) dev[de] = make_float4((float)h[4 * k], (float)h[4 * k + 1], (float)h[4 * k + 2], (float)h[4 * k + 3]);
    else { float4 v = dev[de]; h[4 * k] = (TH)v.x; h[4 * k + 1] = (TH)v.y; h[4 * k + 2] = (TH)v.z; h[4 * k + 3] = (TH)v.w; }
  } else {
    if (DIR == 0) dev[de] = make_double2((double)h[2 * k], (double)h[2 * k + 1]);
    else { double2 v = dev[de]; h[2 * k] = (TH)v.x; h[2 * k + 1] = (TH)v.y; }
  }
}

namespace {
struct HostPipe {          // per-context scratch of the pipelined path (grow-only)
  // four streams: the two copy engines never wait for a layout kernel (those run on xin / xout between them and the compute stream)
  cudaStream_t h2d = nullptr, d2h = nullptr, xin = nullptr, xout = nullptr;
  static constexpr int NB = 3;                               // staging buffers per direction
  void *stage_in[NB] = {nullptr, nullptr, nullptr}, *stage_out[NB] = {nullptr, nullptr, nullptr};
  void *stage_face = nullptr;                                // z faces of a decomposed lattice: 2 x (Lt-2) planes in host order
  size_t stage_bytes = 0, face_bytes = 0;
  std::vector<cudaEvent_t> ev_in, ev_hop;
  cudaEvent_t ev_copied[NB] = {}, ev_xin_done[NB] = {}, ev_packed[NB] = {}, ev_out_copied[NB] = {}, ev_done = nullptr, ev_face_copied = nullptr,
              ev_faces = nullptr;
};
std::mutex pipes_mu;     // one pipe per context (one context per process in this library's usage: one process per GPU)
std::map<gb_context *, HostPipe> &pipes() { static std::map<gb_context *, HostPipe> m; return m; }
HostPipe &host_pipe(gb_context *ctx, size_t slab_bytes, size_t face_bytes, int nslab) {
  std::unique_lock<std::mutex> lk(pipes_mu);
  HostPipe &P = pipes()[ctx];
  lk.unlock();
  if (!P.h2d) {
    GB_CUDA(cudaStreamCreateWithFlags(&P.h2d, cudaStreamNonBlocking));
    GB_CUDA(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
    GB_CUDA(cudaStreamCreateWithFlags(&P.xin, cudaStreamNonBlocking));
    GB_CUDA(cudaStreamCreateWithFlags(&P.xout, cudaStreamNonBlocking));
    for (int i = 0; i < HostPipe::NB; i++)
      for (cudaEvent_t *e : {&P.ev_copied[i], &P.ev_xin_done[i], &P.ev_packed[i], &P.ev_out_copied[i]}) GB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (cudaEvent_t *e : {&P.ev_done, &P.ev_face_copied, &P.ev_faces}) GB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  if (P.stage_bytes < slab_bytes) {
    GB_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < HostPipe::NB; i++) {
      if (P.stage_in[i]) cudaFree(P.stage_in[i]);
      if (P.stage_out[i]) cudaFree(P.stage_out[i]);
      GB_CUDA(cudaMalloc(&P.stage_in[i], slab_bytes));
      GB_CUDA(cudaMalloc(&P.stage_out[i], slab_bytes));
    }
    P.stage_bytes = slab_bytes;
  }
  if (P.face_bytes < face_bytes) {
    GB_CUDA(cudaDeviceSynchronize());
    if (P.stage_face) cudaFree(P.stage_face);
    GB_CUDA(cudaMalloc(&P.stage_face, face_bytes));
    P.face_bytes = face_bytes;
  }
  while ((int)P.ev_in.size() < nslab) {
    cudaEvent_t a, b;
    GB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); GB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    P.ev_in.push_back(a); P.ev_hop.push_back(b);
  }
  return P;
}
template <int DIR>
void launch_transfer(gb_context *ctx, const gb_fermion *f, void *stage, int host_prec, int64_t blk0, int64_t nblk, int64_t host_elem0, int npieces,
                     int64_t blk_stride, int64_t host_stride, cudaStream_t st) {
  SlabGeom G = slab_geom_of(f);
  const int64_t nelem = 2 * nblk * nv_of(f->prec) * W;
  const dim3 grid((unsigned)((nelem + 255) / 256), (unsigned)npieces);
#define GB_L(TD, TH) fermion_slab_transfer_kernel<TD, TH, DIR><<<grid, 256, 0, st>>>((typename Prec<TD>::vec *)f->data, (TH *)stage, G, blk0, nblk, host_elem0, blk_stride, host_stride)
  if (f->prec == GB_F32 && host_prec == GB_F32) GB_L(float, float);
  else if (f->prec == GB_F32) GB_L(float, double);
  else if (host_prec == GB_F32) GB_L(double, float);
  else GB_L(double, double);
#undef GB_L
  count_launch(ctx);
}
} // namespace

// called by gb_context_destroy (context.cu) with the device idle: streams, events and staging buffers of that context's pipe
namespace gb {
void host_pipe_release(gb_context *ctx) {
  std::unique_lock<std::mutex> lk(pipes_mu);
  auto it = pipes().find(ctx);
  if (it == pipes().end()) return;
  HostPipe &P = it->second;
  for (int i = 0; i < HostPipe::NB; i++) {
    if (P.stage_in[i]) cudaFree(P.stage_in[i]);
    if (P.stage_out[i]) cudaFree(P.stage_out[i]);
    for (cudaEvent_t e : {P.ev_copied[i], P.ev_xin_done[i], P.ev_packed[i], P.ev_out_copied[i]}) if (e) cudaEventDestroy(e);
  }
  if (P.stage_face) cudaFree(P.stage_face);
  for (cudaEvent_t e : {P.ev_done, P.ev_face_copied, P.ev_faces}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : P.ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : P.ev_hop) cudaEventDestroy(e);
  for (cudaStream_t s : {P.h2d, P.d2h, P.xin, P.xout}) if (s) cudaStreamDestroy(s);
  pipes().erase(it);
}
} // namespace gb

extern "C" int gb_op_dhop_host(gb_fermop *op, const void *host_in, void *host_out, gb_precision host_prec, int dag) {
  GB_API_BEGIN
  GB_REQUIRE(op && host_in && host_out, "null argument");
  GB_TRACE("DhopHost");
  GB_REQUIRE(op->kind != GB_KIND_STAGGERED, "gb_op_dhop_host serves the Wilson-type operators");
  gb_context *ctx = op->ctx;
  gb_grid *g = op->grid;
  GB_CUDA(cudaSetDevice(ctx->device));
  gb_fermion *fin = op_tmp_full(op, 0), *fout = op_tmp_full(op, 1);
  const int Lz = g->ldims[2], Lt = g->ldims[3];
  const int64_t v3cb = g->V4cb / Lt;
  const bool decomposed = op->comm_dim_mask != 0, z_split = (op->comm_dim_mask >> 2) & 1;
  const int64_t plane_cb = v3cb / Lz;      // checkerboard sites of one (z, t) plane
  const bool pipe_decomposed = !(getenv("GB_HOST_PIPE_DECOMP") && atoi(getenv("GB_HOST_PIPE_DECOMP")) == 0);   // read per call (bench.py falls back on it)
  const bool pipelined = Lt >= 4 && (v3cb * op->Ls) % W == 0 &&
                         (!decomposed || (pipe_decomposed && !(op->comm_dim_mask & 3) && (!z_split || (plane_cb * op->Ls) % W == 0)));
  if (!pipelined) {   // x / y splits, odd shapes: import, hop, export
    int rc = gb_fermion_import(fin, host_in, host_prec); if (rc != GB_OK) return rc;
    op_apply(op, GB_OP_DHOP, fin, fout, dag);
    return gb_fermion_export(fout, host_out, host_prec);
  }
  // The unit of the pipeline is a z-chunk of a t-slice: (t, c) = planes [c zc, (c+1) zc) of slice t, contiguous in host memory and a
  // run of whole blocks on the device.  Its hop needs the chunks (t, c-1), (t, c), (t, c+1) (z legs, periodic) and (t-1, c), (t+1, c)
  // (t legs), so the first result leaves after 1 + 2/k slices have arrived instead of 3, and 2 + 1/k slices are left to export when
  // the H2D stream runs dry instead of 3.  k = GB_HOST_PIPE_ZCHUNKS, reduced until the chunks are whole blocks.  Default 1 (whole
  // slices): measured on the B200 at 32^4 x 16, k = 1 / 4 / 8 -> 37.5 / 42.3 / 47.7 ms per call -- with 4 or 8 times as many copies, layout
  // kernels and stream hand-overs (20 driver calls per unit) the host cannot enqueue the units as fast as PCIe moves them, which costs
  // more than the shorter ramps save (profiles/r2_host_dhop_zchunks.jsonl).
  int k = getenv("GB_HOST_PIPE_ZCHUNKS") ? atoi(getenv("GB_HOST_PIPE_ZCHUNKS")) : 1;
  if (k < 1) k = 1;
  while (k > 1 && (Lz % k != 0 || (plane_cb * op->Ls * (Lz / k)) % W != 0)) k--;
  const int zc = Lz / k;
  const size_t hsz = host_prec == GB_F32 ? 4 : 8;
  const size_t slab_bytes = (size_t)2 * v3cb * op->Ls * 24 * hsz;
  const size_t plane_bytes = slab_bytes / Lz, chunk_bytes = plane_bytes * zc;
  const int64_t chunk_blocks = plane_cb * op->Ls * zc / W;          // device blocks of a chunk per parity
  const int64_t chunk_elems = 2 * plane_cb * op->Ls * zc;           // host-order spinors of a chunk
  HostPipe &P = host_pipe(ctx, chunk_bytes, z_split ? 2 * (size_t)(Lt - 2) * plane_bytes : 0, Lt * k);
  const void *ib[2] = {fin->block(0), fin->block(1)};
  void *ob[2] = {fout->block(0), fout->block(1)};
  // everything previously queued on the compute stream must be done before the copy streams touch the temporaries
  GB_CUDA(cudaEventRecord(P.ev_done, ctx->stream));
  GB_CUDA(cudaStreamWaitEvent(P.h2d, P.ev_done, 0));
  GB_CUDA(cudaStreamWaitEvent(P.d2h, P.ev_done, 0));
  GB_CUDA(cudaStreamWaitEvent(P.xin, P.ev_done, 0));
  GB_CUDA(cudaStreamWaitEvent(P.xout, P.ev_done, 0));
  int nin = 0, nout = 0;
  std::vector<char> imported((size_t)Lt * k, 0);   // host-side record: a hop may only wait for events that were recorded in THIS call
  auto import_unit = [&](int t, int c) {
    const int u = t * k + c;
    const int b = nin++ % HostPipe::NB;
    GB_CUDA(cudaStreamWaitEvent(P.h2d, P.ev_xin_done[b], 0));          // the layout kernel that last read this staging buffer
    GB_CUDA(cudaMemcpyAsync(P.stage_in[b], (const char *)host_in + (size_t)u * chunk_bytes, chunk_bytes, cudaMemcpyHostToDevice, P.h2d));
    GB_CUDA(cudaEventRecord(P.ev_copied[b], P.h2d));
    GB_CUDA(cudaStreamWaitEvent(P.xin, P.ev_copied[b], 0));
    launch_transfer<0>(ctx, fin, P.stage_in[b], host_prec, (int64_t)u * chunk_blocks, chunk_blocks, (int64_t)u * chunk_elems, 1, 0, 0, P.xin);
    GB_CUDA(cudaEventRecord(P.ev_in[u], P.xin));
    GB_CUDA(cudaEventRecord(P.ev_xin_done[b], P.xin));
    imported[u] = 1;
  };
  auto import_slice = [&](int t) { for (int c = 0; c < k; c++) import_unit(t, c); };
  const void *halo[8];
  const void *const *halo_arg = nullptr;
  auto hop_and_export = [&](int t, int c) {
    const int u = t * k + c;
    const int tm = t == 0 ? Lt - 1 : t - 1, tp = t == Lt - 1 ? 0 : t + 1, cm = c == 0 ? k - 1 : c - 1, cp = c == k - 1 ? 0 : c + 1;
    // (on a decomposed lattice the legs that leave the rank read the halo instead: the periodic neighbour is a harmless extra wait)
    for (int need : {tm * k + c, tp * k + c, t * k + cm, u, t * k + cp}) {
      GB_REQUIRE(imported[need], "host pipeline: a unit is scheduled before its inputs");
      GB_CUDA(cudaStreamWaitEvent(ctx->stream, P.ev_in[need], 0));
    }
    dhop_tslab(op, ib, ob, dag, t, 1, ctx->stream, halo_arg, c * zc, zc);
    GB_CUDA(cudaEventRecord(P.ev_hop[u], ctx->stream));
    const int b = nout++ % HostPipe::NB;
    GB_CUDA(cudaStreamWaitEvent(P.xout, P.ev_hop[u], 0));
    GB_CUDA(cudaStreamWaitEvent(P.xout, P.ev_out_copied[b], 0));       // the D2H copy that last read this staging buffer
    launch_transfer<1>(ctx, fout, P.stage_out[b], host_prec, (int64_t)u * chunk_blocks, chunk_blocks, (int64_t)u * chunk_elems, 1, 0, 0, P.xout);
    GB_CUDA(cudaEventRecord(P.ev_packed[b], P.xout));
    GB_CUDA(cudaStreamWaitEvent(P.d2h, P.ev_packed[b], 0));
    GB_CUDA(cudaMemcpyAsync((char *)host_out + (size_t)u * chunk_bytes, P.stage_out[b], chunk_bytes, cudaMemcpyDeviceToHost, P.d2h));
    GB_CUDA(cudaEventRecord(P.ev_out_copied[b], P.d2h));
  };
  if (decomposed) {
    // the sites the neighbours need go in first: the t faces are the slices 0 and Lt-1; z faces: planes z = 0 and z = Lz-1 of the
    // slices 1 ... Lt-2, one strided copy per face (rows = planes, pitch = a slice), then one layout launch per face
    import_slice(Lt - 1); import_slice(0);
    if (z_split) {
      const int np = Lt - 2;
      const int64_t nblk_plane = plane_cb * op->Ls / W;
      const int64_t plane_elems = 2 * plane_cb * op->Ls;               // host-order spinors of one plane
      for (int f = 0; f < 2; f++) {
        const int z = f ? Lz - 1 : 0;
        char *stage = (char *)P.stage_face + (size_t)f * np * plane_bytes;
        GB_CUDA(cudaMemcpy2DAsync(stage, plane_bytes, (const char *)host_in + slab_bytes + (size_t)z * plane_bytes, slab_bytes, plane_bytes, (size_t)np,
                                  cudaMemcpyHostToDevice, P.h2d));
      }
      GB_CUDA(cudaEventRecord(P.ev_face_copied, P.h2d));
      GB_CUDA(cudaStreamWaitEvent(P.xin, P.ev_face_copied, 0));
      for (int f = 0; f < 2; f++) {
        const int z = f ? Lz - 1 : 0;
        char *stage = (char *)P.stage_face + (size_t)f * np * plane_bytes;
        // piece j = plane (z, t = j + 1): device blocks ((j+1) Lz + z) nblk_plane ..., host elements from (z + Lz (j+1)) plane_elems,
        // which sit at j * plane_elems in the staging buffer
        launch_transfer<0>(ctx, fin, stage, host_prec, ((int64_t)Lz + z) * nblk_plane, nblk_plane, ((int64_t)z + Lz) * plane_elems, np,
                           (int64_t)Lz * nblk_plane, (int64_t)(Lz - 1) * plane_elems, P.xin);
      }
    }
    GB_CUDA(cudaEventRecord(P.ev_faces, P.xin));                        // after both slices and both faces (xin is in order)
    GB_CUDA(cudaStreamWaitEvent(ctx->stream, P.ev_faces, 0));
    halo_exchange_only(op, fin, dag, halo);                            // project + send the faces, wait for the neighbours' (compute stream)
    halo_arg = halo;
    for (int c = 0; c < k; c++) { import_unit(1, c); hop_and_export(0, c); }
  } else {
    // slice 0 whole (its chunks are each other's z neighbours), then chunk by chunk its two t neighbours
    import_slice(0);
    for (int c = 0; c < k; c++) { import_unit(Lt - 1, c); import_unit(1, c); hop_and_export(0, c); }
  }
  for (int t = 1; t < Lt - 2; t++)
    for (int c = 0; c < k; c++) { import_unit(t + 1, c); hop_and_export(t, c); }
  for (int c = 0; c < k; c++) hop_and_export(Lt - 2, c);
  for (int c = 0; c < k; c++) hop_and_export(Lt - 1, c);
  GB_CUDA(cudaEventRecord(P.ev_done, P.d2h));
  GB_CUDA(cudaStreamWaitEvent(ctx->stream, P.ev_done, 0));
  GB_CUDA(cudaStreamSynchronize(P.d2h));
  check_launch(ctx, "dhop_host");
  GB_API_END
}
