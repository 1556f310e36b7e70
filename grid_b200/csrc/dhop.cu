// dhop.cu -- hopping-term kernels, gauge DoubleStore, halo pack/exchange and their host orchestration.
//
// Host orchestration replaces WilsonFermion5D::DhopInternal{Serial,Overlapped}Comms
//   (ref: Grid/qcd/action/fermion/implementation/WilsonFermion5DImplementation.h:308-411) and
//   WilsonFermion::DhopInternal* (ref: .../WilsonFermionImplementation.h:394-498);
// the pack kernel replaces WilsonStencil::HaloGatherOpt + FaceGatherSimple::Gather_plane_simple with the
//   projecting WilsonCompressor (ref: WilsonCompressor.h:244-356,461-511 ; Grid/stencil/SimpleCompressor.h:22-37);
// the exchange replaces CartesianStencil::CommunicateBegin/Complete -> StencilSendToRecvFromBegin
//   (ref: Grid/stencil/Stencil.h:367-430 ; Grid/communicator/Communicator_mpi3.cc:390-461) with grouped
//   ncclSend/ncclRecv on a dedicated stream, overlapped with the interior kernel;
// double_store_kernel replaces WilsonImpl::DoubleStore + the -0.5 prefactor of ImportGauge
//   (ref: WilsonImpl.h:127-171 ; WilsonFermion5DImplementation.h:149-181).
#include "dhop_kernel.cuh"
#include "comm.hpp"
#include "fermop.hpp"
#include <algorithm>
#include <functional>

namespace gb {

// =====================================================================================================
// site kernel
// =====================================================================================================
template <class T> struct SiteCtx {
  int xh, y, z, t, pb, s;
  uint32_t site;
};

// face index of a site for a halo in dimension MU (cb index with dimension MU removed; for MU==0 the
// parity constraint halves y).  Must match pack_face_kernel.
template <int MU> __device__ __forceinline__ uint32_t face_index(const DhopArgs &a, int xh, int y, int z, int t) {
  if (MU == 0) return (uint32_t)(y >> 1) + (uint32_t)(a.Ly >> 1) * (z + a.Lz * t);
  if (MU == 1) return (uint32_t)xh + (uint32_t)a.Lxh * (z + a.Lz * t);
  if (MU == 2) return (uint32_t)xh + (uint32_t)a.Lxh * (y + a.Ly * t);
  return (uint32_t)xh + (uint32_t)a.Lxh * (y + a.Ly * z);
}

// MODE 0: every leg (halo legs read the receive buffers); 1: local legs only; 2: off-node legs only
template <class T, int DAG, int MODE, int MU, int FWD>
__device__ __forceinline__ void dhop_leg(const DhopArgs &a, const typename Prec<T>::vec *__restrict__ in,
                                         const typename Prec<T>::vec *__restrict__ Usite, const SiteCtx<T> &c, int ip,
                                         SpinorReg<T> &result, int &nleg) {
  using P = Prec<T>;
  // non-dag: forward legs (x+mu) carry (1-gamma), backward legs (1+gamma); dag swaps.
  // ref: WilsonKernelsImplementation.h:127-134 (dag) vs :154-161
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  if (!((a.leg_mask >> (FWD ? MU : MU + 4)) & 1)) return;   // single-leg applications (DhopDir, force terms)
  const int Lmu = MU == 0 ? a.Lx : MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt;
  int coord;   // local coordinate along MU of the output site
  if (MU == 0) coord = 2 * c.xh + c.pb; else coord = MU == 1 ? c.y : MU == 2 ? c.z : c.t;
  const bool at_edge = FWD ? (coord == Lmu - 1) : (coord == 0);
  const bool offnode = at_edge && ((a.comm_dim_mask >> MU) & 1);
  if (MODE == 1 && offnode) return;
  if (MODE == 2 && !offnode) return;
  HalfReg<T> chi, Uchi;
  if (MODE != 1 && offnode) {
    const uint32_t fi = face_index<MU>(a, c.xh, c.y, c.z, c.t);
    const uint32_t i = fi * a.Ls + c.s;
    const typename P::vec *slot = (const typename P::vec *)a.halo[FWD ? MU : MU + 4] + (size_t)ip * a.halo_parity_stride[MU];
    const size_t vi = ((size_t)(i >> LOGW) * (P::NV / 2) << LOGW) + (i & (W - 1));
    if (a.halo_lowp) load_half_lowp(chi, (const typename LowpVec<T>::type *)slot + vi);
    else load_half(chi, slot + vi);
  } else {
    uint32_t nsite;
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
    const uint32_t i = nsite * a.Ls + c.s;
    SpinorReg<T> f;
    load_spinor(f, in + ((size_t)(i >> LOGW) * P::NV << LOGW) + (i & (W - 1)));
    sp_proj<MU, SIGN>(chi, f);
  }
  LinkReg<T> u;
  if (a.recon12) {
    load_link12(u, Usite + (FWD ? MU : MU + 4) * Recon12<T>::LV);
    mult_link(Uchi, u, chi);
    // the factor the full store folds into the link: -1/2, times the boundary phase (its conjugate on the backward link) on the global boundary
    const int gc = coord + a.origin[MU];
    const bool bnd = FWD ? (gc == a.gL[MU] - 1) : (gc == 0);
    scale_half(Uchi, bnd ? (T)a.bnd_re[FWD ? MU : MU + 4] : (T)-0.5, bnd ? (T)a.bnd_im[FWD ? MU : MU + 4] : (T)0);
  } else {
    load_link(u, Usite + (FWD ? MU : MU + 4) * P::LV);
    mult_link(Uchi, u, chi);
  }
  accum_recon<MU, SIGN>(result, Uchi);
  nleg++;
}

template <class T, int DAG, int MODE>
__global__ void __launch_bounds__(256) dhop_kernel(const DhopArgs a) {
  using P = Prec<T>;
  using V = typename P::vec;
  if (MODE != 1 && a.flags != nullptr) {
    // acquire the neighbours' epoch flags (peer-written, system scope) before touching the receive buffers
    if (threadIdx.x < 8 && ((a.comm_dim_mask >> (threadIdx.x & 3)) & 1)) {
      unsigned long long v;
      do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.flags + threadIdx.x) : "memory"); } while (v < a.epoch);
    }
    __syncthreads();
  }
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n5cb) return;
  const int p = a.first_parity ^ (int)blockIdx.y; // output parity
  SiteCtx<T> c;
  uint32_t r, yl, zl, tl, yh, zh, th, s, xh;
  a.dLs.divmod(q, r, s);
  c.s = s;
  if (a.box_on) {
    uint32_t y, z, t;
    a.dbe0.divmod(r, r, xh); a.dbe1.divmod(r, r, y); a.dbe2.divmod(r, t, z);
    c.xh = a.bo[0] + xh; c.y = a.bo[1] + y; c.z = a.bo[2] + z; c.t = a.bo[3] + t;
  } else {
    a.dLxh.divmod(r, r, xh);
    a.dBy.divmod(r, r, yl);
    a.dBz.divmod(r, r, zl);
    a.dBt.divmod(r, r, tl);
    a.dNy.divmod(r, r, yh);
    a.dNz.divmod(r, th, zh);
    c.xh = xh;
    c.y = yh * a.By + yl; c.z = zh * a.Bz + zl; c.t = th * a.Bt + tl;
  }
  c.site = c.xh + a.Lxh * (c.y + a.Ly * (c.z + a.Lz * c.t));
  c.pb = (p + a.origin_parity + c.y + c.z + c.t) & 1;

  const V *__restrict__ in = (const V *)a.in[1 - p];
  const V *__restrict__ Usite = (const V *)a.U[p] + (size_t)c.site * 8 * (a.recon12 ? Recon12<T>::LV : P::LV);
  const int ip = 1 - p;

  SpinorReg<T> result;
#pragma unroll
  for (int k = 0; k < 12; k++) { result.re[k] = 0; result.im[k] = 0; }
  int nleg = 0;
  // same leg order as the reference site kernel: Xm,Ym,Zm,Tm then Xp,Yp,Zp,Tp
  dhop_leg<T, DAG, MODE, 0, 0>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 1, 0>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 2, 0>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 3, 0>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 0, 1>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 1, 1>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 2, 1>(a, in, Usite, c, ip, result, nleg);
  dhop_leg<T, DAG, MODE, 3, 1>(a, in, Usite, c, ip, result, nleg);

  const uint32_t i = c.site * a.Ls + c.s;
  const size_t off = ((size_t)(i >> LOGW) * P::NV << LOGW) + (i & (W - 1));
  V *outp = (V *)a.out[p] + off;
  if (MODE == 2) {
    if (nleg == 0) return;
    SpinorReg<T> prev;
    load_spinor(prev, outp);
    const T sa = a.axpy[p] ? (T)a.axpy_a : (T)1;
#pragma unroll
    for (int k = 0; k < 12; k++) { result.re[k] = fma(sa, result.re[k], prev.re[k]); result.im[k] = fma(sa, result.im[k], prev.im[k]); }
  } else if (a.axpy[p] != nullptr) {
    SpinorReg<T> ax;
    load_spinor(ax, (const V *)a.axpy[p] + off);
    const T sa = (T)a.axpy_a, sb = (T)a.axpy_b;
#pragma unroll
    for (int k = 0; k < 12; k++) { result.re[k] = fma(sa, result.re[k], sb * ax.re[k]); result.im[k] = fma(sa, result.im[k], sb * ax.im[k]); }
  }
  store_spinor(result, outp);
}

// =====================================================================================================
// halo pack: project the boundary slice with the receiving leg's projector and write it contiguously.
// One thread per (face site, s).  point = the stencil point on the RECEIVER that will consume the data:
//   point mu   (receiver's forward leg, x+mu)  <- sender's slice x_mu = 0,    projector of the forward leg
//   point mu+4 (receiver's backward leg, x-mu) <- sender's slice x_mu = L-1,  projector of the backward leg
// =====================================================================================================
struct PackArgs {
  const void *in;      // parity block being packed (the hop's input parity)
  void *buf;           // send buffer for this (point, parity)
  int Ls, Lx, Lxh, Ly, Lz, Lt;
  int ip;              // parity of the packed field
  int origin_parity;
  int lowp;            // compressed halos (store_half_lowp)
  uint32_t nface;      // face sites (cb) = V4cb / L_mu
};
template <class T, int DAG, int MU, int FWD>
__global__ void pack_face_kernel(const PackArgs a) {
  using P = Prec<T>;
  using V = typename P::vec;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.nface * a.Ls) return;
  const uint32_t fi = q / a.Ls, s = q - fi * a.Ls;
  int xh, y, z, t;
  const int slice = FWD ? 0 : (MU == 0 ? a.Lx : MU == 1 ? a.Ly : MU == 2 ? a.Lz : a.Lt) - 1;
  uint32_t r = fi;
  if (MU == 0) {
    int yhalf = r % (a.Ly >> 1); r /= (a.Ly >> 1); z = r % a.Lz; t = r / a.Lz;
    // x = slice has parity (ip + origin + y + z + t)&1 == slice&1  => fixes y parity
    int ypar = (slice + a.ip + a.origin_parity + z + t) & 1;
    y = 2 * yhalf + ypar; xh = slice >> 1;
  } else if (MU == 1) { xh = r % a.Lxh; r /= a.Lxh; z = r % a.Lz; t = r / a.Lz; y = slice; }
  else if (MU == 2) { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; t = r / a.Ly; z = slice; }
  else { xh = r % a.Lxh; r /= a.Lxh; y = r % a.Ly; z = r / a.Ly; t = slice; }
  const uint32_t site = xh + a.Lxh * (y + a.Ly * (z + a.Lz * t));
  const uint32_t i = site * a.Ls + s;
  SpinorReg<T> f;
  load_spinor(f, (const V *)a.in + ((size_t)(i >> LOGW) * P::NV << LOGW) + (i & (W - 1)));
  constexpr int SIGN = (FWD ? -1 : +1) * (DAG ? -1 : +1);
  HalfReg<T> h;
  sp_proj<MU, SIGN>(h, f);
  const size_t vi = ((size_t)(q >> LOGW) * (P::NV / 2) << LOGW) + (q & (W - 1));
  if (a.lowp) store_half_lowp(h, (typename LowpVec<T>::type *)a.buf + vi);
  else store_half(h, (V *)a.buf + vi);
}

// =====================================================================================================
// DoubleStore
// =====================================================================================================
struct DoubleStoreArgs {
  const void *Ulex;         // [V4][4][9] complex
  const void *Uhalo[4];     // backward-neighbour faces of U_mu (x_mu = L-1 slice of the -mu rank), or nullptr
  void *Uds[2];
  int L[4], gL[4], origin[4];
  int comm_dim_mask;
  double ph_re[4], ph_im[4];
  uint32_t V4cb;
};
template <class T> __global__ void double_store_kernel(const DoubleStoreArgs a) {
  using P = Prec<T>;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; // (parity, site, dir)
  if (e >= 2u * a.V4cb * 8) return;
  const int dir = e & 7;
  uint32_t r = e >> 3;
  const int p = r / a.V4cb;
  const uint32_t site = r - p * a.V4cb;
  const int Lxh = a.L[0] / 2;
  int x[4];
  uint32_t q = site;
  int xh = q % Lxh; q /= Lxh; x[1] = q % a.L[1]; q /= a.L[1]; x[2] = q % a.L[2]; x[3] = q / a.L[2];
  const int op = (a.origin[0] + a.origin[1] + a.origin[2] + a.origin[3]) & 1;
  x[0] = 2 * xh + ((p + op + x[1] + x[2] + x[3]) & 1);
  const int mu = dir & 3;
  const T *src;
  T m[18];
  const T pre = (T)-0.5;
  const int gx = x[mu] + a.origin[mu];
  if (dir < 4) {
    const size_t lex = x[0] + (size_t)a.L[0] * (x[1] + (size_t)a.L[1] * (x[2] + (size_t)a.L[2] * x[3]));
    src = (const T *)a.Ulex + (lex * 4 + mu) * 18;
    T pr = 1, pi = 0;
    if (gx == a.gL[mu] - 1) { pr = (T)a.ph_re[mu]; pi = (T)a.ph_im[mu]; }
    for (int k = 0; k < 9; k++) {
      T ur = src[2 * k], ui = src[2 * k + 1];
      m[2 * k] = pre * (pr * ur - pi * ui);
      m[2 * k + 1] = pre * (pr * ui + pi * ur);
    }
  } else {
    int y[4] = {x[0], x[1], x[2], x[3]};
    if (x[mu] == 0 && ((a.comm_dim_mask >> mu) & 1)) {
      // face index of the lexicographic face with dimension mu removed
      size_t fi = 0, st = 1;
      for (int d = 0; d < 4; d++) if (d != mu) { fi += st * x[d]; st *= a.L[d]; }
      src = (const T *)a.Uhalo[mu] + fi * 18;
    } else {
      y[mu] = (x[mu] + a.L[mu] - 1) % a.L[mu];
      const size_t lex = y[0] + (size_t)a.L[0] * (y[1] + (size_t)a.L[1] * (y[2] + (size_t)a.L[2] * y[3]));
      src = (const T *)a.Ulex + (lex * 4 + mu) * 18;
    }
    T pr = 1, pi = 0;
    if (gx == 0) { pr = (T)a.ph_re[mu]; pi = -(T)a.ph_im[mu]; } // conj(phase)
    for (int rr = 0; rr < 3; rr++) for (int cc = 0; cc < 3; cc++) {
      T ur = src[2 * (cc * 3 + rr)], ui = -src[2 * (cc * 3 + rr) + 1]; // adjoint
      m[2 * (rr * 3 + cc)] = pre * (pr * ur - pi * ui);
      m[2 * (rr * 3 + cc) + 1] = pre * (pr * ui + pi * ur);
    }
  }
  if constexpr (sizeof(T) == 4) {
    float4 *o = (float4 *)a.Uds[p] + ((size_t)site * 8 + dir) * P::LV;
    o[0] = make_float4(m[0], m[1], m[2], m[3]);
    o[1] = make_float4(m[4], m[5], m[6], m[7]);
    o[2] = make_float4(m[8], m[9], m[10], m[11]);
    o[3] = make_float4(m[12], m[13], m[14], m[15]);
    o[4] = make_float4(m[16], m[17], 0.f, 0.f);
  } else {
    double2 *o = (double2 *)a.Uds[p] + ((size_t)site * 8 + dir) * P::LV;
    for (int k = 0; k < 9; k++) o[k] = make_double2(m[2 * k], m[2 * k + 1]);
  }
}
// gather the x_mu = L-1 slice of U_mu (lexicographic face order) for the gauge halo
template <class T> __global__ void gauge_face_kernel(const T *Ulex, T *face, int4 L, int mu, uint32_t nface) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nface * 18) return;
  const uint32_t fi = e / 18, k = e - fi * 18;
  int Ld[4] = {L.x, L.y, L.z, L.w}, x[4];
  uint32_t r = fi;
  for (int d = 0; d < 4; d++) if (d != mu) { x[d] = r % Ld[d]; r /= Ld[d]; }
  x[mu] = Ld[mu] - 1;
  const size_t lex = x[0] + (size_t)Ld[0] * (x[1] + (size_t)Ld[1] * (x[2] + (size_t)Ld[2] * x[3]));
  face[e] = Ulex[(lex * 4 + mu) * 18 + k];
}

// =====================================================================================================
// host side
// =====================================================================================================
// One halo message: my `send` buffer goes to rank `to`, `recv` is filled by rank `from`.  A neighbour that is this rank itself
// (undecomposed dimension whose halos are forced on, GB_SELF_HALO) is a device-to-device copy, like the reference's
// comms-to-self path (ref: Grid/communicator/Communicator_none.cc SendToRecvFrom).
struct HaloMsg { const void *send; void *recv; size_t bytes; int to, from; };
static void halo_sendrecv(gb_context *ctx, const std::vector<HaloMsg> &msgs, cudaStream_t st) {
  bool remote = false;
  for (const HaloMsg &m : msgs) {
    if (m.to == ctx->rank && m.from == ctx->rank) GB_CUDA(cudaMemcpyAsync(m.recv, m.send, m.bytes, cudaMemcpyDeviceToDevice, st));
    else remote = true;
  }
  if (!remote) return;
  GB_REQUIRE(ctx->nccl != nullptr, "decomposed lattice: call gb_comm_init first");
  NcclApi &N = nccl();
  nccl_check(N.GroupStart(), "ncclGroupStart");
  for (const HaloMsg &m : msgs) if (!(m.to == ctx->rank && m.from == ctx->rank)) {
    nccl_check(N.Send(m.send, m.bytes, ncclChar, m.to, ctx->nccl, st), "ncclSend");
    nccl_check(N.Recv(m.recv, m.bytes, ncclChar, m.from, ctx->nccl, st), "ncclRecv");
  }
  nccl_check(N.GroupEnd(), "ncclGroupEnd");
}

static int pick_block(int L, int want) {
  if (want <= 0 || want >= L) return L;
  int b = want;
  while (L % b) b--;
  return b;
}

void op_import_gauge(gb_fermop *op, const gb_gauge *Umu) {
  GB_TRACE("ImportGauge");
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  GB_REQUIRE(Umu->grid == g, "gauge field lives on a different grid");
  GB_REQUIRE(Umu->prec == op->prec, "gauge precision must match the operator precision (use gb_gauge_create + import at that precision)");
  GB_CUDA(cudaSetDevice(ctx->device));
  const size_t vb = 16;
  const size_t per_parity = (size_t)g->V4cb * 8 * lv_of(op->prec) * vb;
  if (!op->Uds) {
    op->uds_bytes = 2 * per_parity;
    GB_CUDA(cudaMalloc(&op->Uds, op->uds_bytes));
  }
  DoubleStoreArgs a;
  a.Ulex = Umu->data;
  a.comm_dim_mask = op->comm_dim_mask;
  a.V4cb = (uint32_t)g->V4cb;
  for (int d = 0; d < 4; d++) {
    a.L[d] = g->ldims[d]; a.gL[d] = g->gdims[d]; a.origin[d] = g->origin[d];
    a.ph_re[d] = op->phases[2 * d]; a.ph_im[d] = op->phases[2 * d + 1];
    a.Uhalo[d] = nullptr;
  }
  a.Uds[0] = op->Uds; a.Uds[1] = (char *)op->Uds + per_parity;
  // gauge halo: my x_mu = L-1 slice of U_mu goes to the forward neighbour; I receive the backward neighbour's
  void *sendf[4] = {nullptr, nullptr, nullptr, nullptr}, *recvf[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t esz = op->prec == GB_F32 ? 4 : 8;
  if (op->comm_dim_mask) {
    int4 L4 = make_int4(g->ldims[0], g->ldims[1], g->ldims[2], g->ldims[3]);
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
      const uint32_t nface = (uint32_t)(g->V4 / g->ldims[mu]);
      GB_CUDA(cudaMalloc(&sendf[mu], (size_t)nface * 18 * esz));
      GB_CUDA(cudaMalloc(&recvf[mu], (size_t)nface * 18 * esz));
      const unsigned blocks = (nface * 18 + 255) / 256;
      if (op->prec == GB_F32) gauge_face_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float *)Umu->data, (float *)sendf[mu], L4, mu, nface);
      else gauge_face_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double *)Umu->data, (double *)sendf[mu], L4, mu, nface);
      count_launch(ctx);
    }
    std::vector<HaloMsg> msgs;
    for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1)
      msgs.push_back({sendf[mu], recvf[mu], (size_t)(g->V4 / g->ldims[mu]) * 18 * esz, g->nbr_rank[mu][0], g->nbr_rank[mu][1]});
    halo_sendrecv(ctx, msgs, ctx->stream);
    for (int mu = 0; mu < 4; mu++) a.Uhalo[mu] = recvf[mu];
  }
  const uint32_t n = 2u * a.V4cb * 8;
  if (op->prec == GB_F32) double_store_kernel<float><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a);
  else double_store_kernel<double><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a);
  count_launch(ctx);
  check_launch(ctx, "double_store");
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int mu = 0; mu < 4; mu++) { if (sendf[mu]) cudaFree(sendf[mu]); if (recvf[mu]) cudaFree(recvf[mu]); }
  if (op->recon12) op_build_recon12(op);   // a new gauge field: the two-row store follows (and is checked again)
}

// allocate halo send/recv buffers (both parities) on first use
static void ensure_halo(gb_fermop *op) {
  if (!op->comm_dim_mask || op->halo_send[0] || op->halo_send[1] || op->halo_send[2] || op->halo_send[3]) return;
  const gb_grid *g = op->grid;
  const int hv = nv_of(op->prec) / 2;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    const size_t nface5 = (size_t)(g->V4cb / g->ldims[mu]) * op->Ls;
    const size_t blocks = (nface5 + W - 1) / W;
    op->halo_parity_stride[mu] = blocks * hv * W; // vecs per parity
    const size_t bytes = 2 * op->halo_parity_stride[mu] * 16;
    for (int dir = 0; dir < 2; dir++) {
      GB_CUDA(cudaMalloc(&op->halo_send[mu + 4 * dir], bytes));
      GB_CUDA(cudaMalloc(&op->halo_recv[mu + 4 * dir], bytes));
    }
  }
  op->halo_ready = true;
}

template <class T, int DAG> static void launch_pack(gb_fermop *op, const void *in_block, int ip, int slot, cudaStream_t st) {
  // slot: 0 for cb operations / parity 0 of a full hop, 1 for the second parity of a full hop
  const gb_grid *g = op->grid;
  gb_context *ctx = op->ctx;
  PackArgs a;
  a.in = in_block; a.Ls = op->Ls;
  a.Lx = g->ldims[0]; a.Lxh = a.Lx / 2; a.Ly = g->ldims[1]; a.Lz = g->ldims[2]; a.Lt = g->ldims[3];
  a.ip = ip;
  a.lowp = op->halo_lowp;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    a.nface = (uint32_t)(g->V4cb / g->ldims[mu]);
    const uint32_t n = a.nface * op->Ls;
    const unsigned blocks = (n + 255) / 256;
    for (int fwd = 0; fwd < 2; fwd++) {
      const int point = fwd ? mu : mu + 4;
      a.buf = (char *)op->halo_send[point] + (size_t)slot * op->halo_parity_stride[mu] * 16;
#define PK(M, F) pack_face_kernel<T, DAG, M, F><<<blocks, 256, 0, st>>>(a)
      if (mu == 0) { if (fwd) PK(0, 1); else PK(0, 0); }
      else if (mu == 1) { if (fwd) PK(1, 1); else PK(1, 0); }
      else if (mu == 2) { if (fwd) PK(2, 1); else PK(2, 0); }
      else { if (fwd) PK(3, 1); else PK(3, 0); }
#undef PK
      count_launch(ctx);
    }
  }
  check_launch(ctx, "pack_face");
}

static void exchange_halos(gb_fermop *op, int nslots, cudaStream_t st) {
  const gb_grid *g = op->grid;
  gb_context *ctx = op->ctx;
  std::vector<HaloMsg> msgs;
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) {
    // compressed halos fill the first half of each parity slot: the message ends with the last slot's data
    const size_t bytes = ((size_t)nslots * 16 - (op->halo_lowp ? 8 : 0)) * op->halo_parity_stride[mu];
    // data for the receiver's forward leg (point mu) is my x=0 slice: it travels to my backward neighbour
    msgs.push_back({op->halo_send[mu], op->halo_recv[mu], bytes, g->nbr_rank[mu][1], g->nbr_rank[mu][0]});
    // data for the receiver's backward leg (point mu+4) is my x=L-1 slice: it travels forward
    msgs.push_back({op->halo_send[mu + 4], op->halo_recv[mu + 4], bytes, g->nbr_rank[mu][0], g->nbr_rank[mu][1]});
  }
  halo_sendrecv(ctx, msgs, st);
}

template <class T> static void launch_dhop_T(gb_fermop *op, DhopArgs &a, int nparity, int dag, int mode, cudaStream_t st) {
  dim3 grid((a.n5cb + 255) / 256, nparity);
#define DK(D, M) dhop_kernel<T, D, M><<<grid, 256, 0, st>>>(a)
  if (!dag) { if (mode == 0) DK(0, 0); else if (mode == 1) DK(0, 1); else DK(0, 2); }
  else { if (mode == 0) DK(1, 0); else if (mode == 1) DK(1, 1); else DK(1, 2); }
#undef DK
  count_launch(op->ctx);
  check_launch(op->ctx, "dhop");
}

// link pointers of the generic kernel: the full doubled store, or (gb_op_set_link_reconstruct 12) the two-row store plus what the
// kernel needs to put the folded-in factors back
static void fill_link_args(const gb_fermop *op, DhopArgs &a) {
  const gb_grid *g = op->grid;
  a.recon12 = op->recon12 && op->Uds12 != nullptr;
  const size_t per_parity = (size_t)g->V4cb * 8 * (a.recon12 ? (op->prec == GB_F32 ? 3 : 6) : lv_of(op->prec)) * 16;
  for (int p = 0; p < 2; p++) a.U[p] = (const char *)(a.recon12 ? op->Uds12 : op->Uds) + p * per_parity;
  for (int d = 0; d < 4; d++) {
    a.gL[d] = g->gdims[d]; a.origin[d] = g->origin[d];
    a.bnd_re[d] = -0.5 * op->phases[2 * d]; a.bnd_im[d] = -0.5 * op->phases[2 * d + 1];            // forward link on the global boundary
    a.bnd_re[d + 4] = -0.5 * op->phases[2 * d]; a.bnd_im[d + 4] = 0.5 * op->phases[2 * d + 1];    // backward link: conjugate phase
  }
}

// Two-row link store built from the full doubled store (so the neighbours' backward links are already in place): rows 0 and 1 of
// every link divided by its folded-in factor.  dev[] receives, per link, how far the third row rebuilt from the two stored ones is
// from the stored third row -- the links must be special unitary for the reconstruction to be legitimate.
struct Recon12Args {
  const void *Uds; void *Uds12; float *dev;
  int L[4], gL[4], origin[4];
  double bnd_re[8], bnd_im[8];
  uint32_t V4cb;
};
template <class T> __global__ void recon12_store_kernel(const Recon12Args a) {
  using P = Prec<T>;
  using V = typename P::vec;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; // (parity, site, point)
  if (e >= 2u * a.V4cb * 8) return;
  const int pt = e & 7, mu = pt & 3;
  uint32_t r = e >> 3;
  const int p = r / a.V4cb;
  const uint32_t site = r - p * a.V4cb;
  const int Lxh = a.L[0] / 2;
  int x[4];
  uint32_t q = site;
  const int xh = q % Lxh; q /= Lxh; x[1] = q % a.L[1]; q /= a.L[1]; x[2] = q % a.L[2]; x[3] = q / a.L[2];
  const int opar = (a.origin[0] + a.origin[1] + a.origin[2] + a.origin[3]) & 1;
  x[0] = 2 * xh + ((p + opar + x[1] + x[2] + x[3]) & 1);
  const int gx = x[mu] + a.origin[mu];
  const bool bnd = pt < 4 ? (gx == a.gL[mu] - 1) : (gx == 0);
  const double cr = bnd ? a.bnd_re[pt] : -0.5, ci = bnd ? a.bnd_im[pt] : 0.0;
  const double n = 1.0 / (cr * cr + ci * ci);
  const T ir = (T)(cr * n), ii = (T)(-ci * n);          // 1 / c
  LinkReg<T> u;
  load_link(u, (const V *)a.Uds + (size_t)p * a.V4cb * 8 * P::LV + ((size_t)site * 8 + pt) * P::LV);
#pragma unroll
  for (int k = 0; k < 9; k++) { const T re = u.re[k], im = u.im[k]; u.re[k] = ir * re - ii * im; u.im[k] = ir * im + ii * re; }
  LinkReg<T> w = u;
  recon_row2(w);
  T d = 0;
#pragma unroll
  for (int k = 6; k < 9; k++) d = fmax(d, fmax(fabs(w.re[k] - u.re[k]), fabs(w.im[k] - u.im[k])));
  a.dev[e] = (float)d;
  constexpr int LV12 = Recon12<T>::LV;
  V *o = (V *)a.Uds12 + (size_t)p * a.V4cb * 8 * LV12 + ((size_t)site * 8 + pt) * LV12;
  if constexpr (sizeof(T) == 4) {
    o[0] = make_float4(u.re[0], u.im[0], u.re[1], u.im[1]);
    o[1] = make_float4(u.re[2], u.im[2], u.re[3], u.im[3]);
    o[2] = make_float4(u.re[4], u.im[4], u.re[5], u.im[5]);
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = make_double2(u.re[k], u.im[k]);
  }
}
// (re)build the two-row store of an operator that asked for it; throws GB_ERR_INVALID (and leaves the full store in use) when the
// links are not special unitary to working precision
void op_build_recon12(gb_fermop *op) {
  gb_context *ctx = op->ctx;
  const gb_grid *g = op->grid;
  GB_REQUIRE(op->Uds != nullptr, "operator has no gauge field: call ImportGauge first");
  GB_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)2 * g->V4cb * 8 * (op->prec == GB_F32 ? 3 : 6) * 16;
  if (!op->Uds12) GB_CUDA(cudaMalloc(&op->Uds12, bytes));
  const uint32_t n = 2u * (uint32_t)g->V4cb * 8;
  float *dev = nullptr;
  GB_CUDA(cudaMalloc(&dev, (size_t)n * sizeof(float)));
  struct Free { float *p; ~Free() { cudaFree(p); } } free_dev{dev};   // also on the error paths below
  Recon12Args a;
  a.Uds = op->Uds; a.Uds12 = op->Uds12; a.dev = dev; a.V4cb = (uint32_t)g->V4cb;
  DhopArgs la;
  fill_link_args(op, la);
  for (int d = 0; d < 4; d++) { a.L[d] = g->ldims[d]; a.gL[d] = la.gL[d]; a.origin[d] = la.origin[d]; }
  for (int k = 0; k < 8; k++) { a.bnd_re[k] = la.bnd_re[k]; a.bnd_im[k] = la.bnd_im[k]; }
  if (op->prec == GB_F32) recon12_store_kernel<float><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a);
  else recon12_store_kernel<double><<<(n + 127) / 128, 128, 0, ctx->stream>>>(a);
  count_launch(ctx);
  check_launch(ctx, "recon12_store");
  std::vector<float> h(n);
  GB_CUDA(cudaMemcpyAsync(h.data(), dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  float worst = 0;
  for (float v : h) worst = v > worst || v != v ? v : worst;
  const float tol = op->prec == GB_F32 ? 2e-5f : 1e-12f;
  if (!(worst <= tol)) {
    cudaFree(op->Uds12); op->Uds12 = nullptr;
    throw Error(GB_ERR_INVALID, "12-real link reconstruction needs special unitary links: the third row rebuilt from the first two differs from the stored one by " +
                                    std::to_string(worst));
  }
}

// The one entry used by every operator: hop from the parity blocks in[] to out[].
//   parity_out_first: output parity of the first (or only) parity; nparity 1 (DhopEO/OE) or 2 (full Dhop)
//   axpy: optional out = a*hop + b*ax (ax blocks indexed by output parity)
void dhop_blocks(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                 const void *const ax[2], double axa, double axb) {
  gb_context *ctx = op->ctx;
  GB_TRACE(dag ? "DhopDag" : "Dhop");
  const gb_grid *g = op->grid;
  GB_CUDA(cudaSetDevice(ctx->device));
  DhopArgs a;
  fill_link_args(op, a);
  for (int p = 0; p < 2; p++) {
    a.in[p] = in[p]; a.out[p] = out[p];
    a.axpy[p] = ax ? ax[p] : nullptr;
  }
  a.axpy_a = axa; a.axpy_b = axb;
  a.comm_dim_mask = op->comm_dim_mask;
  a.halo_lowp = op->halo_lowp;
  a.Ls = op->Ls; a.Lx = g->ldims[0]; a.Lxh = a.Lx / 2; a.Ly = g->ldims[1]; a.Lz = g->ldims[2]; a.Lt = g->ldims[3];
  a.By = pick_block(a.Ly, op->By); a.Bz = pick_block(a.Lz, op->Bz); a.Bt = pick_block(a.Lt, op->Bt);
  a.dLs = FastDiv(a.Ls); a.dLxh = FastDiv(a.Lxh); a.dBy = FastDiv(a.By); a.dBz = FastDiv(a.Bz); a.dBt = FastDiv(a.Bt);
  a.dNy = FastDiv(a.Ly / a.By); a.dNz = FastDiv(a.Lz / a.Bz);
  a.n5cb = (uint32_t)(g->V4cb * op->Ls);
  a.first_parity = parity_out_first;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  for (int i = 0; i < 8; i++) a.halo[i] = nullptr;
  for (int i = 0; i < 4; i++) a.halo_parity_stride[i] = 0;
  a.mode = 0;
  a.box_on = 0;
  a.flags = nullptr; a.epoch = 0;
  a.leg_mask = op->leg_mask;

  auto run = [&](int mode, cudaStream_t st) {
    if (op->prec == GB_F32) launch_dhop_T<float>(op, a, nparity, dag, mode, st);
    else launch_dhop_T<double>(op, a, nparity, dag, mode, st);
  };
  // exterior pass: only the surface slabs, as disjoint boxes (ref: st.surface_list, Stencil.h:664-686)
  auto run_exterior = [&](cudaStream_t st, int ext_mode) {
    int lo[4] = {0, 0, 0, 0}, hi[4] = {a.Lxh, a.Ly, a.Lz, a.Lt};
    const uint32_t n5_full = a.n5cb;
    for (int d = 3; d >= 0; d--) {
      if (!((op->comm_dim_mask >> d) & 1)) continue;
      for (int side = 0; side < 2; side++) {
        int blo[4] = {lo[0], lo[1], lo[2], lo[3]}, bhi[4] = {hi[0], hi[1], hi[2], hi[3]};
        if (side == 0) bhi[d] = lo[d] + 1; else blo[d] = hi[d] - 1;
        if (bhi[d] - blo[d] <= 0 || (side == 1 && hi[d] - lo[d] == 1)) continue;
        a.box_on = 1;
        uint64_t vol = 1;
        for (int k = 0; k < 4; k++) { a.bo[k] = blo[k]; a.be[k] = bhi[k] - blo[k]; vol *= a.be[k]; }
        if (vol == 0) continue;
        a.dbe0 = FastDiv(a.be[0]); a.dbe1 = FastDiv(a.be[1]); a.dbe2 = FastDiv(a.be[2]);
        a.n5cb = (uint32_t)(vol * op->Ls);
        run(ext_mode, st);
      }
      lo[d] += 1; hi[d] -= 1;
      if (hi[d] <= lo[d]) break;
    }
    a.box_on = 0;
    a.n5cb = n5_full;
  };
  if (!op->comm_dim_mask) {
    if (!dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 0, ctx->stream)) run(0, ctx->stream);
    return;
  }
  // Overlapped hop of a z/t-decomposed lattice, "split volume" form: the interior box (every leg local) is computed by the
  // tuned kernel with all 8 legs while the faces travel, then the surface slabs are computed ONCE, also with all 8 legs
  // (local legs + halo legs), by the generic kernel.  Compared with the reference's interior + accumulate-exterior passes
  // (WilsonFermion5DImplementation.h:320-384) no site is visited twice and nothing is read-modify-written.
  // Opt-in (GB_SPLIT=1): measured on B200 at 32^4 x 16 per GPU it gains 3-4 % on a bare Dhop (2 GPUs 1.205 -> 1.169 ms,
  // 4 GPUs 1.317 -> 1.265 ms) but the Schur CG, whose hops alternate with s-space kernels, ran 7 % slower (1.747 -> 1.869 s),
  // so the default stays interior + accumulate.
  static const bool no_split = getenv("GB_SPLIT") == nullptr;
  auto hop_overlapped = [&](cudaStream_t st, const std::function<void()> &before_exterior) {
    bool split = !no_split && !(op->comm_dim_mask & 3) && op->prec == GB_F32 && !op->disable_fast &&
                 dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 3, st);
    if (!split) {
      if (!dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 1, st)) run(1, st);
    }
    before_exterior();
    run_exterior(st, split ? 0 : 2);
  };

  // ---- multi-GPU, peer-to-peer path: one pack+send kernel (stores into the neighbours' buffers over NVLink),
  //      interior kernel, exterior slabs that acquire the neighbours' epoch flags on the device
  if (p2p_setup(op)) {
    static const bool prof = getenv("GB_PROFILE") != nullptr;
    static cudaEvent_t qe[3];
    static double qacc[2];
    static int qn = 0;
    if (prof && qn == 0) for (auto &e : qe) GB_CUDA(cudaEventCreate(&e));
    if (prof) GB_CUDA(cudaEventRecord(qe[0], ctx->stream));
    // fully fused path: ONE launch projects + sends the faces (pack CTAs interleaved at the head of the grid), does the
    // local legs, and surface CTAs acquire the neighbours' flags and add the off-node legs
    // (the tuned multi-rank launches read and send uncompressed halos: an operator with compressed halos takes the pack kernel and the
    //  reference's interior + exterior form -- tuned kernel on the local legs, generic kernel on the surface slabs -- or the serial form)
    const bool lowp = op->halo_lowp != 0;
    const bool try_fused = op->overlap_comms && !op->no_fused && op->prec == GB_F32 && !op->disable_fast && !lowp;
    unsigned long long epoch = 0;
    bool hop_sends_t = false;
    // semi-fused: after the pack+send kernel the hop does the local legs and, in its last (surface) CTAs, acquires the
    // neighbours' flags and adds the halo legs -- no exterior pass, nothing read-modify-written (GB_SEMIFUSED=0 disables).
    // Default engine: the column-sweep kernel (dhop_col2.cuh) over the planes whose z legs are local, t-surface columns last,
    // plus the micro-block semi-fused kernel on the two z-surface planes when z is split (GB_COL2_DECOMP=0: micro-block
    // kernel over the whole volume, the round-1 form).
    static const bool semifused = !(getenv("GB_SEMIFUSED") && atoi(getenv("GB_SEMIFUSED")) == 0);
    static const bool col2_decomp = !(getenv("GB_COL2_DECOMP") && atoi(getenv("GB_COL2_DECOMP")) == 0);
    static const bool pack_on_comm_stream = getenv("GB_PACK_STREAM") && atoi(getenv("GB_PACK_STREAM")) != 0;
    bool pack_pending = false;
    if (try_fused) {
      epoch = p2p_next_epoch(op);
      p2p_fill_halo(op, epoch, a.halo, &a.flags);
      for (int i = 0; i < 4; i++) a.halo_parity_stride[i] = op->halo_parity_stride[i];
      const void *hb[8];
      for (int i = 0; i < 8; i++) hb[i] = a.halo[i];
      if (nparity == 1) {
        const int ip = 1 - parity_out_first;
        for (int i = 0; i < 8; i++) if (hb[i]) hb[i] = (const char *)hb[i] - (size_t)ip * op->halo_parity_stride[i & 3] * 16;
      }
      if (dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 2, ctx->stream, hb, a.flags, epoch)) return;
      p2p_send_only(op, epoch, in, parity_out_first, nparity, dag, ctx->stream); // configuration not covered: separate pack kernel
    } else if (pack_on_comm_stream && op->overlap_comms) {
      // pack + send on the high-priority comm stream: it only reads the hop's input, so the hop need not wait for it; the
      // compute stream waits for its completion at the END of this call (the input may be overwritten by the next kernel)
      GB_CUDA(cudaEventRecord(ctx->ev_comp, ctx->stream));
      GB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comp, 0));
      epoch = p2p_pack_send(op, in, parity_out_first, nparity, dag, ctx->comm_stream);
      GB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
      pack_pending = true;
    } else {
      // default: the column-sweep hop sends its own t faces (no re-read of them by the pack kernel), the pack kernel the rest
      hop_sends_t = !lowp && op->overlap_comms && semifused && col2_decomp && !op->no_semifused && ((op->comm_dim_mask >> 3) & 1) && g->ldims[3] >= 4 &&
                    dhop_col2_applicable(op, 1) && !(getenv("GB_HOP_SENDS_T") && atoi(getenv("GB_HOP_SENDS_T")) == 0);
      epoch = p2p_next_epoch(op);
      {
        GB_TRACE("Gather");
        const bool z_in = ((op->comm_dim_mask >> 2) & 1) && dhop_col2_zplanes_inkernel();
        p2p_send_only(op, epoch, in, parity_out_first, nparity, dag, ctx->stream, hop_sends_t ? (z_in ? 2 : 1) : 0);
      }
    }
    struct PackJoin {   // runs on every exit path below
      gb_context *ctx; bool on;
      ~PackJoin() { if (on) cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0); }
    } pack_join{ctx, pack_pending};
    if (prof) GB_CUDA(cudaEventRecord(qe[1], ctx->stream));
    struct ProfEnd {
      gb_context *ctx; bool on; cudaEvent_t *qe; double *qacc; int *qn;
      ~ProfEnd() {
        if (!on) return;
        cudaEventRecord(qe[2], ctx->stream); cudaEventSynchronize(qe[2]);
        float ms; cudaEventElapsedTime(&ms, qe[0], qe[1]); qacc[0] += ms; cudaEventElapsedTime(&ms, qe[1], qe[2]); qacc[1] += ms;
        if (++(*qn) % 50 == 0 && ctx->rank == 0) fprintf(stderr, "[gb profile p2p] calls %d: pack+send %.4f  hop kernels %.4f ms\n", *qn, qacc[0] / *qn, qacc[1] / *qn);
      }
    } prof_end{ctx, prof, qe, qacc, &qn};
    p2p_fill_halo(op, epoch, a.halo, &a.flags);
    a.epoch = epoch;
    for (int i = 0; i < 4; i++) a.halo_parity_stride[i] = op->halo_parity_stride[i];
    if (nparity == 1) {
      const int ip = 1 - parity_out_first;
      for (int i = 0; i < 8; i++) if (a.halo[i]) a.halo[i] = (const char *)a.halo[i] - (size_t)ip * op->halo_parity_stride[i & 3] * 16;
    }
    if (hop_sends_t) {
      Col2Send snd;
      p2p_fill_send_t(op, epoch, snd.dst, snd.flag, &snd.counter);
      GB_REQUIRE(dhop_col2_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 1, ctx->stream, a.halo, a.flags, epoch, &snd),
                 "column-sweep hop with hop-sent t faces");
      if (((op->comm_dim_mask >> 2) & 1) && !dhop_col2_zplanes_inkernel()) {
        const bool z0 = dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 5, ctx->stream, a.halo, a.flags, epoch);
        const bool z1 = dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 6, ctx->stream, a.halo, a.flags, epoch);
        GB_REQUIRE(z0 && z1, "z-surface planes of the column-sweep hop");
      }
      return;
    }
    if (!lowp && op->overlap_comms && semifused && !op->no_semifused && op->prec == GB_F32 && !op->disable_fast && !(op->comm_dim_mask & 3)) {
      if (col2_decomp && dhop_col2_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 1, ctx->stream, a.halo, a.flags, epoch)) {
        if (((op->comm_dim_mask >> 2) & 1) && !dhop_col2_zplanes_inkernel()) {
          const bool z0 = dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 5, ctx->stream, a.halo, a.flags, epoch);
          const bool z1 = dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 6, ctx->stream, a.halo, a.flags, epoch);
          GB_REQUIRE(z0 && z1, "z-surface planes of the column-sweep hop");
        }
        return;
      }
      if (dhop_fast_launch(op, in, out, parity_out_first, nparity, dag, ax, axa, axb, 4, ctx->stream, a.halo, a.flags, epoch)) return;
    }
    if (op->overlap_comms) hop_overlapped(ctx->stream, [] {});
    else run(0, ctx->stream);
    return;
  }
  // ---- multi-GPU, NCCL path: pack -> exchange (comm stream) || interior (compute stream) -> exterior
  ensure_halo(op);
  for (int i = 0; i < 8; i++) a.halo[i] = op->halo_recv[i];
  for (int i = 0; i < 4; i++) a.halo_parity_stride[i] = op->halo_parity_stride[i];
  // the halo "slot" of an input parity: cb ops use slot 0; the full hop stores input parity ip in slot ip
  // (dhop_leg indexes with ip * stride, so for cb ops we bias the base pointer instead)
  if (nparity == 1) {
    const int ip = 1 - parity_out_first;
    for (int i = 0; i < 8; i++) if (a.halo[i]) a.halo[i] = (const char *)a.halo[i] - (size_t)ip * op->halo_parity_stride[i & 3] * 16;
  }
  // make sure previous consumers of the send/recv buffers are done (StencilBarrier analogue):
  // everything is stream ordered on ctx->stream, and the comm stream waits on the pack event.
  static const bool profile = getenv("GB_PROFILE") != nullptr;
  static cudaEvent_t pe[6];
  static double pacc[5];
  static int pcount = 0;
  if (profile && pcount == 0) for (auto &e : pe) GB_CUDA(cudaEventCreate(&e));
  if (profile) GB_CUDA(cudaEventRecord(pe[0], ctx->stream));
  for (int j = 0; j < nparity; j++) {
    const int po = parity_out_first ^ j, ip = 1 - po;
    const int slot = nparity == 1 ? 0 : ip;
    if (op->prec == GB_F32) { if (dag) launch_pack<float, 1>(op, in[ip], ip, slot, ctx->stream); else launch_pack<float, 0>(op, in[ip], ip, slot, ctx->stream); }
    else { if (dag) launch_pack<double, 1>(op, in[ip], ip, slot, ctx->stream); else launch_pack<double, 0>(op, in[ip], ip, slot, ctx->stream); }
  }
  GB_CUDA(cudaEventRecord(ctx->ev_comp, ctx->stream));
  if (profile) GB_CUDA(cudaEventRecord(pe[1], ctx->stream));
  GB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comp, 0));
  if (profile) GB_CUDA(cudaEventRecord(pe[4], ctx->comm_stream));
  exchange_halos(op, nparity, ctx->comm_stream);
  GB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
  if (profile) GB_CUDA(cudaEventRecord(pe[5], ctx->comm_stream));
  if (op->overlap_comms) {
    hop_overlapped(ctx->stream, [&] {                      // interior while the faces travel, then the surface slabs
      if (profile) GB_CUDA(cudaEventRecord(pe[2], ctx->stream));
      GB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    });
  } else {
    if (profile) GB_CUDA(cudaEventRecord(pe[2], ctx->stream));
    GB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    run(0, ctx->stream);
  }
  if (profile) {
    GB_CUDA(cudaEventRecord(pe[3], ctx->stream));
    GB_CUDA(cudaEventSynchronize(pe[3]));
    GB_CUDA(cudaEventSynchronize(pe[5]));
    float ms;
    cudaEventElapsedTime(&ms, pe[0], pe[1]); pacc[0] += ms;
    cudaEventElapsedTime(&ms, pe[1], pe[2]); pacc[1] += ms;
    cudaEventElapsedTime(&ms, pe[2], pe[3]); pacc[2] += ms;
    cudaEventElapsedTime(&ms, pe[4], pe[5]); pacc[3] += ms;
    cudaEventElapsedTime(&ms, pe[0], pe[3]); pacc[4] += ms;
    if (++pcount % 50 == 0 && ctx->rank == 0)
      fprintf(stderr, "[gb profile] calls %d: pack %.4f interior %.4f wait+exterior %.4f | exchange(comm stream) %.4f | total %.4f ms\n", pcount,
              pacc[0] / pcount, pacc[1] / pcount, pacc[2] / pcount, pacc[3] / pcount, pacc[4] / pcount);
  }
}

// The face exchange of one full-lattice hop on its own (no hopping kernel): project + send every face, then wait until the
// neighbours' faces have arrived.  This is the library's Benchmark_comms (ref: benchmarks/Benchmark_comms.cc:105-162 times
// StencilSendToRecvFrom of L^3 Ls half-spinor packets in all split directions concurrently); here the projection is fused
// into the send, so the timed unit is pack + transfer + arrival.  Returns the bytes this rank sent.
__global__ void halo_wait_kernel(const unsigned long long *flags, unsigned long long epoch, int comm_dim_mask) {
  if (threadIdx.x < 8 && ((comm_dim_mask >> (threadIdx.x & 3)) & 1)) {
    unsigned long long v;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory"); } while (v < epoch);
  }
}
size_t halo_exchange_only(gb_fermop *op, const gb_fermion *in, int dag, const void **halo_out) {
  GB_TRACE("HaloExchange");
  GB_REQUIRE(op && in && op->kind != GB_KIND_STAGGERED, "halo exchange benchmark: Wilson-type operators");
  GB_REQUIRE(in->grid == op->grid && in->Ls == op->Ls && in->prec == op->prec && in->kind == GB_FULL, "field is not a conformable full-grid field");
  gb_context *ctx = op->ctx;
  if (!op->comm_dim_mask) return 0;
  GB_CUDA(cudaSetDevice(ctx->device));
  const void *ib[2] = {in->block(0), in->block(1)};
  size_t bytes = 0;
  if (p2p_setup(op)) {
    const unsigned long long epoch = p2p_pack_send(op, ib, 0, 2, dag, ctx->stream);
    const void *halo[8]; const unsigned long long *flags = nullptr;
    p2p_fill_halo(op, epoch, halo, &flags);
    halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(flags, epoch, op->comm_dim_mask);
    count_launch(ctx);
    check_launch(ctx, "halo_wait");
    if (halo_out) for (int i = 0; i < 8; i++) halo_out[i] = halo[i];
  } else {
    ensure_halo(op);
    for (int ip = 0; ip < 2; ip++) {
      if (op->prec == GB_F32) { if (dag) launch_pack<float, 1>(op, ib[ip], ip, ip, ctx->stream); else launch_pack<float, 0>(op, ib[ip], ip, ip, ctx->stream); }
      else { if (dag) launch_pack<double, 1>(op, ib[ip], ip, ip, ctx->stream); else launch_pack<double, 0>(op, ib[ip], ip, ip, ctx->stream); }
    }
    exchange_halos(op, 2, ctx->stream);
    if (halo_out) for (int i = 0; i < 8; i++) halo_out[i] = op->halo_recv[i];
  }
  for (int mu = 0; mu < 4; mu++) if ((op->comm_dim_mask >> mu) & 1) bytes += 2 * 2 * op->halo_parity_stride[mu] * 16;   // two directions, two parities
  return bytes;
}

// Hop restricted to the t-slices [t0, t0 + nt) of both output parities, all 8 legs.  halo == nullptr: single rank (periodic wrap
// inside the local volume).  halo != nullptr (decomposed lattice): the legs that leave the rank read the receive buffers of a halo
// exchange that has COMPLETED in front of this launch on `st` (halo_exchange_only: both input parities, slot = parity).
// Used by the host-pipelined Dhop (dhop_host.cu), where slices are computed as their neighbours arrive over PCIe.
void dhop_tslab(gb_fermop *op, const void *const in[2], void *const out[2], int dag, int t0, int nt, cudaStream_t st, const void *const *halo,
                int z0, int nz) {
  const gb_grid *g = op->grid;
  DhopArgs a;
  fill_link_args(op, a);
  for (int p = 0; p < 2; p++) { a.in[p] = in[p]; a.out[p] = out[p]; a.axpy[p] = nullptr; }
  a.axpy_a = 1; a.axpy_b = 0;
  a.comm_dim_mask = halo ? op->comm_dim_mask : 0;
  a.halo_lowp = op->halo_lowp;
  a.Ls = op->Ls; a.Lx = g->ldims[0]; a.Lxh = a.Lx / 2; a.Ly = g->ldims[1]; a.Lz = g->ldims[2]; a.Lt = g->ldims[3];
  a.By = a.Ly; a.Bz = a.Lz; a.Bt = a.Lt;
  a.dLs = FastDiv(a.Ls); a.dLxh = FastDiv(a.Lxh); a.dBy = FastDiv(a.By); a.dBz = FastDiv(a.Bz); a.dBt = FastDiv(a.Bt);
  a.dNy = FastDiv(1); a.dNz = FastDiv(1);
  a.first_parity = 0;
  a.origin_parity = (g->origin[0] + g->origin[1] + g->origin[2] + g->origin[3]) & 1;
  for (int i = 0; i < 8; i++) a.halo[i] = halo ? halo[i] : nullptr;
  for (int i = 0; i < 4; i++) a.halo_parity_stride[i] = halo ? op->halo_parity_stride[i] : 0;
  a.mode = 0; a.flags = nullptr; a.epoch = 0;
  a.leg_mask = halo ? op->leg_mask : 0xFF;
  a.box_on = 1;
  if (nz < 0) { z0 = 0; nz = a.Lz; }                       // the planes [z0, z0 + nz) of those slices (default: all)
  a.bo[0] = 0; a.bo[1] = 0; a.bo[2] = z0; a.bo[3] = t0;
  a.be[0] = a.Lxh; a.be[1] = a.Ly; a.be[2] = nz; a.be[3] = nt;
  a.dbe0 = FastDiv(a.be[0]); a.dbe1 = FastDiv(a.be[1]); a.dbe2 = FastDiv(a.be[2]);
  a.n5cb = (uint32_t)((size_t)a.Lxh * a.Ly * nz * nt * op->Ls);
  if (op->prec == GB_F32) launch_dhop_T<float>(op, a, 2, dag, 0, st);
  else launch_dhop_T<double>(op, a, 2, dag, 0, st);
}

} // namespace gb
