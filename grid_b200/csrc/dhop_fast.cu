// dhop_fast.cu -- instantiations + launcher of the tuned fp32 hopping kernel (see dhop_fast.cuh)
#include "dhop_fast.cuh"
#include "dhop_col.cuh"
#include "dhop_col2.cuh"
#include <cstdlib>
#include "fermop.hpp"
#include <algorithm>

namespace gb {

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

template <int LS> static void launch_ls(const FastArgs &a, int nparity, int dag, int interior, cudaStream_t st) {
  dim3 grid((a.V4cb + FAST_NSITE - 1) / FAST_NSITE, nparity);
  if (interior == 2) grid = dim3(a.nhop_ctas_per_parity * nparity + a.npack_ctas, 1);
  const int threads = FAST_NSITE * LS;
  if (!dag) {
    if (interior == 2) dhop_fast_kernel<LS, 0, 2><<<grid, threads, 0, st>>>(a);
    else if (interior) dhop_fast_kernel<LS, 0, 1><<<grid, threads, 0, st>>>(a);
    else dhop_fast_kernel<LS, 0, 0><<<grid, threads, 0, st>>>(a);
  } else {
    if (interior == 2) dhop_fast_kernel<LS, 1, 2><<<grid, threads, 0, st>>>(a);
    else if (interior) dhop_fast_kernel<LS, 1, 1><<<grid, threads, 0, st>>>(a);
    else dhop_fast_kernel<LS, 1, 0><<<grid, threads, 0, st>>>(a);
  }
}

// ---- column-sweep kernel (dhop_col.cuh): single-rank hops and the interior pass of z/t-decomposed lattices
template <int LS> static bool launch_col_ls(const ColArgs &a, dim3 grid, int dag, int interior, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = col_smem_bytes<LS>();
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int threads = COL_NSITE * LS;
  if (a.N < 0) {   // marker set by the launcher: two t-slices per CTA
    ColArgs a2 = a; a2.N = -a.N;
    if constexpr (LS * COL_NSITE * 2 <= 1024 && col_smem_bytes<LS, 2>() <= 232448) {
      static bool attr2 = false;
      const size_t smem2 = col_smem_bytes<LS, 2>();
      if (!attr2) {
        GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 0, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        GB_CUDA(cudaFuncSetAttribute(dhop_col_kernel<LS, 1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        attr2 = true;
      }
      if (!dag) dhop_col_kernel<LS, 0, 0, 2><<<grid, 2 * threads, smem2, st>>>(a2); else dhop_col_kernel<LS, 1, 0, 2><<<grid, 2 * threads, smem2, st>>>(a2);
      return true;
    }
    return false;   // the two-slice variant does not exist for this Ls: nothing was launched, the caller takes another kernel
  }
  if (!dag) { if (interior) dhop_col_kernel<LS, 0, 1><<<grid, threads, smem, st>>>(a); else dhop_col_kernel<LS, 0, 0><<<grid, threads, smem, st>>>(a); }
  else { if (interior) dhop_col_kernel<LS, 1, 1><<<grid, threads, smem, st>>>(a); else dhop_col_kernel<LS, 1, 0><<<grid, threads, smem, st>>>(a); }
  return true;
}
static bool dhop_col_launch(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                            const void *const ax[2], double axa, double axb, int interior, cudaStream_t st) {
  static const bool disabled = getenv("GB_NO_COL") != nullptr;
  static const int env_n = getenv("GB_COL_N") ? atoi(getenv("GB_COL_N")) : 0;
  const gb_grid *g = op->grid;
  const int Ls = op->Ls;
  if (disabled || op->no_col || !(Ls == 8 || Ls == 12 || Ls == 16)) return false;
  // decomposed lattices: the interior pass runs beside the pack / exterior / copy kernels of the halo exchange, and two
  // column CTAs fill an SM's shared memory (207 KB) so those cannot co-reside: measured on 2 B200, t split, 32^4 x 16 per
  // GPU: 1.248 ms with this kernel as interior pass vs 1.174 ms with the micro-block kernel.  GB_COL_INTERIOR=1 forces it.
  static const bool col_interior = getenv("GB_COL_INTERIOR") != nullptr;
  if (interior > 1 || (interior && (!col_interior || (op->comm_dim_mask & 3)))) return false;
  const int Lxh = g->ldims[0] / 2, Ly = g->ldims[1], Lz = g->ldims[2], Lt = g->ldims[3];
  if (Lxh % 4 || Ly % 4) return false;
  int N = env_n > 0 ? env_n : (op->col_n > 0 ? op->col_n : 16);
  if (N > Lz) N = Lz;
  while (Lz % N) N--;
  ColArgs a;
  const size_t per_parity = (size_t)g->V4cb * 8 * 5;
  for (int p = 0; p < 2; p++) {
    a.in[p] = (const float4 *)in[p]; a.out[p] = (float4 *)out[p];
    a.U[p] = (const float4 *)op->Uds + p * per_parity;
    a.axpy[p] = ax ? (const float4 *)ax[p] : nullptr;
  }
  a.axpy_a = (float)axa; a.axpy_b = (float)axb;
  a.comm_dim_mask = interior ? op->comm_dim_mask : 0;
  // GB_COL_NT=2: two adjacent t-slices per CTA (only for the single-rank kernel and Ls = 8, 12, 16)
  static const int env_nt = getenv("GB_COL_NT") ? atoi(getenv("GB_COL_NT")) : 1;
  const int NT = (env_nt == 2 && !interior && Lt % 2 == 0) ? 2 : 1;
  a.Lxh = Lxh; a.Ly = Ly; a.Lz = Lz; a.Lt = Lt; a.N = NT == 2 ? -N : N;
  a.dLt = FastDiv(Lt / NT); a.dNxo = FastDiv(Lxh / 4); a.dNyo = FastDiv(Ly / 4);
  a.first_parity = parity_out_first;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  dim3 grid((unsigned)((Lt / NT) * (Lxh / 4) * (Ly / 4) * (Lz / N)), nparity);
  bool launched;
  switch (Ls) {
  case 8: launched = launch_col_ls<8>(a, grid, dag, interior, st); break;
  case 12: launched = launch_col_ls<12>(a, grid, dag, interior, st); break;
  default: launched = launch_col_ls<16>(a, grid, dag, interior, st); break;
  }
  if (!launched) return false;
  count_launch(op->ctx);
  check_launch(op->ctx, "dhop_col");
  return true;
}

// ---- second-generation column-sweep kernel (dhop_col2.cuh): TMA-filled ring; single rank (mode 0) and z/t-decomposed lattices
//      (mode 1: local z legs only -> planes [1, Lz-2] when z is split; off-node t legs from the receive buffers in the surface-t
//      CTAs, which come last in the grid).  The caller computes the z-surface planes with the micro-block kernel (interior 5 / 6).
template <class K> static void launch_col2_k(K dhop_col2_kernel_fn, const Col2Args &a, unsigned nblocks, int threads, size_t smem, int cluster, cudaStream_t st) {
#ifdef __CUDACC__   // (the CPU mock of tests/mock compiles this file with a host compiler: no cluster launches there)
  if (cluster > 1 && nblocks % cluster == 0) {
    // thread-block clusters along the fastest grid index (t): the hardware starts the CTAs of a cluster together, which keeps
    // t neighbours in lockstep along z (no distributed shared memory is used)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GB_CUDA(cudaLaunchKernelEx(&cfg, dhop_col2_kernel_fn, a));
    return;
  }
#endif
  dhop_col2_kernel_fn<<<nblocks, threads, smem, st>>>(a);
}
template <int LS> static void launch_col2_ls(const Col2Args &a, unsigned nblocks, int dag, int mode, int cluster, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = col2_smem_bytes<LS>();
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<LS, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<LS, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<LS, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<LS, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int threads = COL_NSITE * LS;
  if (!dag) { if (mode) launch_col2_k(dhop_col2_kernel<LS, 0, 1>, a, nblocks, threads, smem, cluster, st); else launch_col2_k(dhop_col2_kernel<LS, 0, 0>, a, nblocks, threads, smem, cluster, st); }
  else { if (mode) launch_col2_k(dhop_col2_kernel<LS, 1, 1>, a, nblocks, threads, smem, cluster, st); else launch_col2_k(dhop_col2_kernel<LS, 1, 0>, a, nblocks, threads, smem, cluster, st); }
}
// the two epilogue instantiations (Ls = 16, one for each direction of the Schur operator)
static void launch_col2_epi(const Col2Args &a, unsigned nblocks, int dag, int mode, int epi, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = col2_smem_bytes<16>();
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<16, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<16, 0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<16, 1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaFuncSetAttribute(dhop_col2_kernel<16, 1, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int threads = COL_NSITE * 16;
  if (epi == 1) { if (mode) dhop_col2_kernel<16, 0, 1, 1><<<nblocks, threads, smem, st>>>(a); else dhop_col2_kernel<16, 0, 0, 1><<<nblocks, threads, smem, st>>>(a); }
  else { if (mode) dhop_col2_kernel<16, 1, 1, 2><<<nblocks, threads, smem, st>>>(a); else dhop_col2_kernel<16, 1, 0, 2><<<nblocks, threads, smem, st>>>(a); }
}
bool dhop_col2_zplanes_inkernel() { return !(getenv("GB_COL2_ZPLANES") && atoi(getenv("GB_COL2_ZPLANES")) == 0); }
bool dhop_col2_applicable(const gb_fermop *op, int mode) {
  static const bool disabled = getenv("GB_NO_COL") != nullptr || (getenv("GB_COL2") && atoi(getenv("GB_COL2")) == 0);
  const gb_grid *g = op->grid;
  const int Ls = op->Ls;
  if (disabled || op->no_col || op->prec != GB_F32 || op->disable_fast || op->recon12 || !(Ls == 8 || Ls == 12 || Ls == 16)) return false;
  if ((g->ldims[0] / 2) % 4 || g->ldims[1] % 4) return false;
  if (mode == 0 && op->comm_dim_mask) return false;
  if (mode == 1 && (op->comm_dim_mask & 3)) return false;
  if (mode == 1 && op->halo_lowp) return false;   // compressed halos: generic surface kernel (dhop.cu)
  return true;
}
bool dhop_col2_launch(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                      const void *const ax[2], double axa, double axb, int mode, cudaStream_t st, const void *const halo[8],
                      const unsigned long long *flags, unsigned long long epoch, const Col2Send *snd) {
  static const int env_n = getenv("GB_COL_N") ? atoi(getenv("GB_COL_N")) : 0;
  // rasterisation: 1 (default) = t fastest, then the y blocks, then the x blocks: the y-edge rows a column shares with its y
  // neighbours (half a fetch per site) are then 32 CTAs apart in launch order instead of 128 and are re-hit in L2
  static const int env_raster = getenv("GB_COL_RASTER") ? atoi(getenv("GB_COL_RASTER")) : 1;
  const gb_grid *g = op->grid;
  const int Ls = op->Ls;
  if (!dhop_col2_applicable(op, mode)) return false;
  const int Lxh = g->ldims[0] / 2, Ly = g->ldims[1], Lz = g->ldims[2], Lt = g->ldims[3];
  if (mode == 1 && (halo == nullptr || flags == nullptr)) return false;
  const bool z_comm = mode == 1 && ((op->comm_dim_mask >> 2) & 1), t_comm = mode == 1 && ((op->comm_dim_mask >> 3) & 1);
  Col2Args a;
  const size_t per_parity = (size_t)g->V4cb * 8 * 5;
  for (int p = 0; p < 2; p++) {
    a.in[p] = (const float4 *)in[p]; a.out[p] = (float4 *)out[p];
    a.U[p] = (const float4 *)op->Uds + p * per_parity;
    a.axpy[p] = ax ? (const float4 *)ax[p] : nullptr;
  }
  a.axpy_a = (float)axa; a.axpy_b = (float)axb;
  a.Lxh = Lxh; a.Ly = Ly; a.Lz = Lz; a.Lt = Lt;
  // z-decomposed lattices: by default the columns sweep all planes and take the outward z legs of planes 0 and Lz-1 from the receive
  // buffers themselves (z_inkernel); GB_COL2_ZPLANES=0 restores the round-2a form (columns over planes 1 ... Lz-2, the two surface planes
  // by box launches of the micro-block kernel)
  const bool z_inkernel = z_comm && dhop_col2_zplanes_inkernel() && Lz >= 2;
  a.z_inkernel = z_inkernel ? 1 : 0;
  a.halo_zm = halo ? (const float4 *)halo[6] : nullptr; a.halo_zp = halo ? (const float4 *)halo[2] : nullptr;
  a.hstride_z = op->halo_parity_stride[2];
  if (z_comm && !z_inkernel) {        // planes 0 and Lz-1 have an off-node z leg: the caller's box launches take them
    a.z0 = 1; a.N = Lz - 2; a.nzc = 1;
    if (a.N <= 0) return true;        // nothing but surface planes
  } else {
    // default 16 planes per column: measured on the B200 at 32^4 x 16, 0.938 ms against 0.977 ms for whole-z columns (short
    // columns keep t neighbours in step along z, so their re-reads hit L2; the two extra planes per chunk cost less than that)
    int N = env_n > 0 ? env_n : (op->col_n > 0 ? op->col_n : 16);
    if (N > Lz) N = Lz;
    while (Lz % N) N--;
    a.z0 = 0; a.N = N; a.nzc = Lz / N;
  }
  const uint32_t cols_per_t = (uint32_t)(Lxh / 4) * (Ly / 4) * a.nzc;
  a.t_comm = t_comm ? 1 : 0;
  a.nt_surf = t_comm ? 2 : 0;
  a.nt_int = t_comm ? Lt - 2 : Lt; a.t_int0 = t_comm ? 1 : 0;
  a.n_int = (uint32_t)a.nt_int * cols_per_t; a.n_surf = (uint32_t)a.nt_surf * cols_per_t;
  a.dnt_int = FastDiv(std::max(1, a.nt_int)); a.dnt_surf = FastDiv(std::max(1, a.nt_surf));
  a.dNxo = FastDiv(Lxh / 4); a.dNyo = FastDiv(Ly / 4);
  a.raster = env_raster;
  static const int env_tb = getenv("GB_COL_TB") ? atoi(getenv("GB_COL_TB")) : 0;
  static const int env_l2pf = getenv("GB_COL_L2PF") ? atoi(getenv("GB_COL_L2PF")) : 0;
  a.tb = (env_tb > 0 && a.nt_int > env_tb && a.nt_int % env_tb == 0) ? env_tb : 0;
  a.l2pf = env_l2pf;
  a.dtb = FastDiv(std::max(1, a.tb)); a.dNzc = FastDiv(std::max(1, a.nzc));
  if (a.raster == 2 && t_comm) a.raster = 1;
  static const int env_sync = getenv("GB_COL2_SYNC") ? atoi(getenv("GB_COL2_SYNC")) : 0;
  a.cta_sync = env_sync;
  a.zero = 0;
  static const int env_hints = getenv("GB_COL_HINTS") ? atoi(getenv("GB_COL_HINTS")) : 0;
  a.hints = env_hints;
  a.nparity = nparity; a.first_parity = parity_out_first;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  a.halo_tm = halo ? (const float4 *)halo[7] : nullptr; a.halo_tp = halo ? (const float4 *)halo[3] : nullptr;
  a.hstride = op->halo_parity_stride[3];
  a.flags = flags; a.epoch = epoch;
  a.send_on = 0; a.send_dst[0] = a.send_dst[1] = nullptr; a.send_flag[0] = a.send_flag[1] = nullptr; a.send_counter = nullptr; a.n_senders = 0;
  a.send_pstride = op->halo_parity_stride[3];
  if (snd != nullptr) {
    GB_REQUIRE(t_comm && Lt >= 4, "hop-sent t faces need a t-decomposed lattice with Lt >= 4");
    a.send_on = 1;
    a.send_dst[0] = (float4 *)snd->dst[0]; a.send_dst[1] = (float4 *)snd->dst[1];
    a.send_flag[0] = snd->flag[0]; a.send_flag[1] = snd->flag[1];
    a.send_counter = snd->counter;
    a.n_senders = 2u * cols_per_t * (uint32_t)nparity;      // the columns of slices t = 1 and t = Lt-2
    a.tb = 0;
  }
  const unsigned nblocks = (unsigned)((a.n_int + a.n_surf) * (uint32_t)nparity);
  if (nblocks == 0) return true;
  // an s-space pass of the Schur CG parked on the operator for this hop (fermop.cu): EPI 1 rides on the plain hop, EPI 2 on the daggered one
  a.e_aux = nullptr; a.e_r = nullptr; a.e_c = a.e_d = nullptr; a.e_partials = nullptr;
  if (HopEpilogue *E = op->hop_epi) {
    const bool want = E->kind != 0 && !E->applied && getenv("GB_NO_HOP_EPI") == nullptr;
    if (want && Ls == 16 && nparity == 1 && !(z_comm && !z_inkernel) && ax == nullptr && dag == (E->kind == 2 ? 1 : 0) &&
        smat_tri_onesided(op, E->Maux, a.e_ad, a.e_ao, a.e_adir) && (E->kind == 1 || smat_tri_onesided(op, E->Mhop, a.e_hd, a.e_ho, a.e_hdir))) {
      a.e_aux = (const float4 *)E->aux->data;
      a.e_r = E->r ? (float4 *)E->r->data : nullptr;
      a.e_c = E->d_c; a.e_d = E->d_d;
      a.e_partials = smat_partials_ensure(op, nblocks);
      launch_col2_epi(a, nblocks, dag, mode, E->kind, st);
      count_launch(op->ctx);
      check_launch(op->ctx, "dhop_col2 + s-space epilogue");
      smat_reduce_partials(op, nblocks, E->d_out);
      E->applied = true;
      return true;
    }
  }
  static const int env_cluster = getenv("GB_COL_CLUSTER") ? atoi(getenv("GB_COL_CLUSTER")) : 0;
  switch (Ls) {
  case 8: launch_col2_ls<8>(a, nblocks, dag, mode, env_cluster, st); break;
  case 12: launch_col2_ls<12>(a, nblocks, dag, mode, env_cluster, st); break;
  default: launch_col2_ls<16>(a, nblocks, dag, mode, env_cluster, st); break;
  }
  count_launch(op->ctx);
  check_launch(op->ctx, "dhop_col2");
  return true;
}

// returns false when the configuration is not covered by the fast path (caller falls back to dhop_kernel)
bool dhop_fast_launch(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                      const void *const ax[2], double axa, double axb, int interior, cudaStream_t st, const void *const halo[8],
                      const unsigned long long *flags, unsigned long long epoch) {
  const gb_grid *g = op->grid;
  if (op->prec != GB_F32 || op->disable_fast || op->recon12) return false;   // (two-row links: generic kernel)
  if (interior == 0 && dhop_col2_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 0, st, nullptr, nullptr, 0)) return true;
  if (interior < 3 && dhop_col_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, interior, st)) return true;
  const int Ls = op->Ls;
  if (!(Ls == 8 || Ls == 12 || Ls == 16 || Ls == 24 || Ls == 32)) return false;
  if (g->V4cb % FAST_NSITE || ((size_t)(g->ldims[0] / 2) * g->ldims[1]) % FAST_NSITE) return false;
  FastArgs a;
  const size_t per_parity = (size_t)g->V4cb * 8 * 5;
  for (int p = 0; p < 2; p++) {
    a.in[p] = (const float4 *)in[p]; a.out[p] = (float4 *)out[p];
    a.U[p] = (const float4 *)op->Uds + p * per_parity;
    a.axpy[p] = ax ? (const float4 *)ax[p] : nullptr;
  }
  a.axpy_a = (float)axa; a.axpy_b = (float)axb;
  // interior == 3: only the sites whose every leg is local (z, t in [1, L-1) of the decomposed dimensions), all 8 legs;
  //                the surface sites are computed, also with all legs, by the caller's exterior pass -- no read-modify-write
  // interior == 4: "semi-fused" -- one launch does the local legs AND, in the CTAs that own surface sites (rasterised last),
  //                acquires the neighbours' epoch flags and adds the off-node legs from the receive buffers; the faces were
  //                projected and sent by the separate pack_send kernel that precedes it in the stream.  Unlike interior == 2
  //                (pack CTAs inside the same launch) no CTA ever waits for work of its own launch, so it cannot deadlock.
  // interior == 5 / 6: the semi-fused launch restricted to the plane z = 0 / z = Lz-1 (the z surface of a column-sweep hop)
  const int zplane = interior == 5 ? 0 : interior == 6 ? g->ldims[2] - 1 : -1;
  const bool no_pack = interior == 4 || zplane >= 0;
  if (no_pack) interior = 2;
  const bool inner_box = interior == 3;
  if (inner_box) {
    if (op->comm_dim_mask & 3) return false;               // x / y decomposition: caller uses interior + accumulate passes
    interior = 0;
  }
  a.comm_dim_mask = interior ? op->comm_dim_mask : 0;
  a.Lxh = g->ldims[0] / 2; a.Ly = g->ldims[1]; a.Lz = g->ldims[2]; a.Lt = g->ldims[3];
  a.ibx = gcd_int(a.Lxh, 4); a.iby = gcd_int(a.Ly, 4);
  a.zo = a.to = 0;
  int nz = a.Lz, nt = a.Lt;
  if (inner_box) {
    if ((op->comm_dim_mask >> 2) & 1) { a.zo = 1; nz = a.Lz - 2; }
    if ((op->comm_dim_mask >> 3) & 1) { a.to = 1; nt = a.Lt - 2; }
    if (nz <= 0 || nt <= 0) return true;                   // no interior sites at all
  }
  if (zplane >= 0) { a.zo = zplane; nz = 1; }
  int bz = op->Bz <= 0 ? nz : op->Bz;
  if (bz > nz) bz = nz;
  if (nz % bz) { int best = 1; for (int d = 2; d <= 12 && d <= nz; d++) if (nz % d == 0) best = d; bz = best == 1 ? nz : best; }
  a.Bz = bz;
  a.dibx = FastDiv(a.ibx); a.diby = FastDiv(a.iby); a.dNxo = FastDiv(a.Lxh / a.ibx); a.dNyo = FastDiv(a.Ly / a.iby);
  a.dBz = FastDiv(a.Bz); a.dLt = FastDiv(nt);
  a.V4cb = (uint32_t)((size_t)a.Lxh * a.Ly * nz * nt);
  a.first_parity = parity_out_first;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  for (int i = 0; i < 8; i++) a.halo[i] = halo ? (const float4 *)halo[i] : nullptr;
  for (int i = 0; i < 4; i++) a.hstride[i] = op->halo_parity_stride[i];
  a.flags = flags; a.epoch = epoch;
  a.rot_z = (zplane < 0) && ((op->comm_dim_mask >> 2) & 1); a.rot_t = (op->comm_dim_mask >> 3) & 1;
  if (interior == 2 && (halo == nullptr || flags == nullptr)) return false;
  a.npack_items = 0; a.npack_ctas = 0; a.pack_ratio = 4; a.pack_counter = nullptr;
  a.nhop_ctas_per_parity = (a.V4cb + FAST_NSITE - 1) / FAST_NSITE;
  for (int k = 0; k < 8; k++) a.peer_flag[k] = nullptr;
  if (interior == 2 && !no_pack) {
    // pack items of this hop: for every decomposed dimension, both faces, every input parity
    P2PState &S = op->p2p;
    const size_t eoff = (size_t)(epoch & 1) * S.epoch_stride;
    const uint32_t cta_threads = FAST_NSITE * Ls;
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
      const uint32_t nface = (uint32_t)(g->V4cb / g->ldims[mu]);
      for (int fwd = 0; fwd < 2; fwd++) {
        const int point = fwd ? mu : mu + 4;
        a.peer_flag[point] = (unsigned long long *)((char *)S.peer_base[point] + S.flags_off) + (epoch & 1) * 8 + point;
        for (int j = 0; j < nparity; j++) {
          const int po = parity_out_first ^ j, ip = 1 - po;
          const int slot = nparity == 1 ? 0 : ip;
          FastArgs::PackItemF &it = a.pack[a.npack_items++];
          it.src = (const float4 *)in[ip];
          it.dst = (float4 *)((char *)S.peer_base[point] + eoff + S.pt_off[point] + (size_t)slot * op->halo_parity_stride[mu] * 16);
          it.nface = nface; it.mu = mu; it.fwd = fwd; it.ip = ip;
          it.cta_start = a.npack_ctas;
          a.npack_ctas += (nface * (uint32_t)Ls + cta_threads - 1) / cta_threads;
        }
      }
    }
    a.pack_counter = S.d_counter;
    // pack CTAs sit at block indices 0, ratio, 2*ratio, ...: they must all exist inside the grid
    const uint32_t nhop_total = a.nhop_ctas_per_parity * (uint32_t)nparity;
    a.pack_ratio = std::max<uint32_t>(1u, std::min<uint32_t>(4u, 1u + nhop_total / std::max<uint32_t>(a.npack_ctas, 1u)));
  }
  switch (Ls) {
  case 8: launch_ls<8>(a, nparity, dag, interior, st); break;
  case 12: launch_ls<12>(a, nparity, dag, interior, st); break;
  case 16: launch_ls<16>(a, nparity, dag, interior, st); break;
  case 24: launch_ls<24>(a, nparity, dag, interior, st); break;
  default: launch_ls<32>(a, nparity, dag, interior, st); break;
  }
  count_launch(op->ctx);
  check_launch(op->ctx, "dhop_fast");
  return true;
}

} // namespace gb
