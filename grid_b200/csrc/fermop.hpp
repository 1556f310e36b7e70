// fermop.hpp -- the FermionOperator object behind the C ABI and the internal kernel entry points.
#pragma once
#include "internal.hpp"
#include <cstdlib>
#include <map>

enum gb_opkind { GB_KIND_WILSON = 0, GB_KIND_CAYLEY = 1, GB_KIND_STAGGERED = 2 };

namespace gb {
// dense s-space operator: [2 chiralities][Ls][Ls] doubles, row-major (smat.cu)
struct SMat { int Ls = 0; std::vector<double> a; };
}

namespace gb {
// peer-to-peer halo state (halo_p2p.cu)
struct P2PState {
  bool tried = false, ok = false;
  void *recv_base = nullptr;       // [2 epochs][regions per point] + flags [2][8]
  size_t recv_bytes = 0, pt_off[8] = {0, 0, 0, 0, 0, 0, 0, 0}, epoch_stride = 0, flags_off = 0;
  void *peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // mapped base of the consumer of point p
  std::vector<void *> opened;
  unsigned long long epoch = 0;
  unsigned int *d_counter = nullptr;
};
}

namespace gb {
// An s-space pass of the Schur CG folded into the hop that precedes it (dhop_col2.cuh, EPI 1 / 2).  The caller parks a request in
// gb_fermop::hop_epi around ONE checkerboard hop; the column-sweep launch honours it where it can (fp32, Ls = 16, no x / y / z
// decomposition, cyclic-bidiagonal operators) and sets `applied`, otherwise the hop runs plain and the caller does the pass itself.
struct HopEpilogue {
  int kind = 0;                                  // 1: out = T_aux(aux) - hop, |out|^2 -> d_out ; 2: r -= (c/d)(T_aux(aux) + T_hop(hop)), |r|^2 -> d_out
  const gb_fermion *aux = nullptr;
  gb_fermion *r = nullptr;
  const void *Maux = nullptr, *Mhop = nullptr;   // device handles of the operators (their host copies: gb_fermop::smat_host)
  const double *d_c = nullptr, *d_d = nullptr;
  double *d_out = nullptr;
  bool applied = false;
};
}

struct gb_fermop {
  gb_grid *grid = nullptr;
  gb_context *ctx = nullptr;
  int kind = 0, prec = 0, Ls = 1;
  double mass = 0, M5 = 0;
  double phases[8] = {1, 0, 1, 0, 1, 0, 1, 0};
  gb::CayleyCoeffs k;
  void *Uds12 = nullptr;       // gb_op_set_link_reconstruct(12): two rows of the bare SU(3) links, [2][V4cb][8][3 float4 | 6 double2] (generic kernel)
  int recon12 = 0;
  void *Uds = nullptr; // doubled links [2 parities][V4cb][8][LV] vecs, -1/2 and phases folded in
  size_t uds_bytes = 0;
  // rasterisation blocking of the hopping kernel (0 = whole extent)
  int By = 0, Bz = 0, Bt = 0;
  // multi-GPU halos (ref: CartesianStencil u_send_buf/u_recv_buf, Stencil.h:839-848)
  int comm_dim_mask = 0;
  bool overlap_comms = true;
  bool no_semifused = false;   // overlapped hops: interior + accumulate-exterior passes instead of the semi-fused launch
  bool disable_fast = false;   // force the generic kernel (tests compare the two)
  bool no_col = false;         // fp32: use the micro-block kernel instead of the column-sweep kernel (tests compare them)
  int col_n = 0;               // z-planes per column of the column-sweep kernel (0 = default 16)
  int halo_lowp = 0;           // gb_op_set_halo_compression: halos one precision down (fp32 -> bf16, fp64 -> fp32); generic multi-rank forms only
  int leg_mask = 0xFF;         // legs of the hopping term that contribute (0xFF always, except inside op_dhop_leg: DhopDir / force terms)
  // multi-GPU: the single-launch pack+hop+halo kernel is EXPERIMENTAL (opt-in with GB_FUSED=1).  It is parity-green on small
  // lattices but can deadlock at 32^4 per GPU: surface CTAs spinning on the neighbours' flags can fill every resident slot
  // before the last pack CTAs of the same launch are scheduled.  Default = pack+send kernel, interior pass, exterior slabs.
  bool no_fused = getenv("GB_FUSED") == nullptr;
  bool halo_ready = false;
  void *halo_send[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void *halo_recv[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t halo_parity_stride[4] = {0, 0, 0, 0};
  gb::P2PState p2p;
  // dense s-space matrices on the device (operator precision); null when Ls is outside the smat kernel's set
  bool use_smat = false;
  const void *sm_meooe5d = nullptr, *sm_meooedag5d = nullptr, *sm_mooee = nullptr, *sm_mooeedag = nullptr, *sm_mooeeinv = nullptr,
             *sm_mooeeinvdag = nullptr, *sm_m5unit = nullptr, *sm_m5unitdag = nullptr, *sm_B = nullptr, *sm_Bdag = nullptr, *sm_negAdag = nullptr;
  std::vector<void *> smat_allocs;
  std::map<const void *, gb::SMat> smat_host;   // host copies by device pointer (the streaming kernel of smat.cu takes its coefficients from them)
  gb::HopEpilogue *hop_epi = nullptr;   // request parked around one hop (see HopEpilogue)
  double *smat_partials = nullptr;   // per-CTA partial sums of the s-space passes that carry a reduction (smat.cu)
  size_t smat_partials_n = 0;
  // improved staggered (stag.cu): 16 scaled + phased links per site, per output parity, streamed layout
  void *stag_links = nullptr;
  size_t stag_parity_bytes = 0;
  double stag_c1 = 1, stag_c2 = 1, stag_u0 = 1;
  // temporaries (ref: FermionOperator::tmp(), and the stack Fields of SchurDiagMooeeOperator)
  gb_fermion *tmp_h[4] = {nullptr, nullptr, nullptr, nullptr};
  gb_fermion *tmp_f[2] = {nullptr, nullptr};
};

namespace gb {
void op_import_gauge(gb_fermop *op, const gb_gauge *Umu);
void dhop_blocks(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                 const void *const ax[2], double axa, double axb);

bool dhop_fast_launch(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                      const void *const ax[2], double axa, double axb, int interior, cudaStream_t st, const void *const halo[8] = nullptr,
                      const unsigned long long *flags = nullptr, unsigned long long epoch = 0);

// t faces sent by the hop itself (dhop_col2.cuh send_on): destinations and flags in the neighbours' memory, last-sender counter
struct Col2Send { void *dst[2] = {nullptr, nullptr}; unsigned long long *flag[2] = {nullptr, nullptr}; unsigned int *counter = nullptr; };
bool dhop_col2_applicable(const gb_fermop *op, int mode);   // would dhop_col2_launch take this operator in this mode?
bool dhop_col2_zplanes_inkernel();                          // z-decomposed lattices: the columns take the z-surface planes themselves
bool dhop_col2_launch(gb_fermop *op, const void *const in[2], void *const out[2], int parity_out_first, int nparity, int dag,
                      const void *const ax[2], double axa, double axb, int mode, cudaStream_t st, const void *const halo[8],
                      const unsigned long long *flags, unsigned long long epoch, const Col2Send *snd = nullptr);

bool p2p_setup(gb_fermop *op);
void p2p_teardown(gb_fermop *op);
unsigned long long p2p_next_epoch(gb_fermop *op);
void p2p_send_only(gb_fermop *op, unsigned long long epoch, const void *const in[2], int parity_out_first, int nparity, int dag, cudaStream_t st,
                   int hop_sends_t = 0);
void p2p_fill_send_t(gb_fermop *op, unsigned long long epoch, void *dst[2], unsigned long long *flag[2], unsigned int **counter);
unsigned long long p2p_pack_send(gb_fermop *op, const void *const in[2], int parity_out_first, int nparity, int dag, cudaStream_t st);
void p2p_fill_halo(gb_fermop *op, unsigned long long epoch, const void *halo[8], const unsigned long long **flags);

// 5D s-direction kernels (cayley.cu).  All operate on `nparity` parity blocks of nblk blocks each.
// chi = diag_s*phi_s + upper_s*P(-/+)psi_{s+1} + lower_s*P(+/-)psi_{s-1} [+ alpha*w]
void m5d_apply(gb_fermop *op, const gb_fermion *psi, const gb_fermion *phi, gb_fermion *chi, const std::vector<double> &lower,
               const std::vector<double> &diag, const std::vector<double> &upper, int dag, const gb_fermion *w, double alpha);
void mooee_inv_apply(gb_fermop *op, const gb_fermion *psi, gb_fermion *chi, int dag);

// dense s-space operators (smat.cu)
SMat smat_identity(int Ls);
SMat smat_m5d(int Ls, const std::vector<double> &lower, const std::vector<double> &diag, const std::vector<double> &upper, int dag);
SMat smat_mooee_inv(const CayleyCoeffs &k, int dag);
SMat smat_mul(const SMat &A, const SMat &B);
SMat smat_scale(const SMat &A, double f);
const void *smat_device(gb_fermop *op, const SMat &m);
bool smat_apply(gb_fermop *op, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, double alpha, const gb_fermion *z,
                gb_fermion *out);
// the same pass with the conjugate-gradient linear algebra folded in (device-resident scalars; see smat.cu)
struct SMatCG { const double *d_c = nullptr, *d_d = nullptr, *d_cp = nullptr; double *d_out = nullptr; gb_fermion *psi = nullptr, *p = nullptr; };
bool smat_apply_norm(gb_fermop *op, const void *dM, const gb_fermion *x, double alpha, const gb_fermion *z, gb_fermion *out, double *d_out);
bool smat_apply_rupd(gb_fermop *op, const void *dM, const gb_fermion *x, const void *dN, const gb_fermion *y, gb_fermion *r, const double *d_c,
                     const double *d_d, double *d_out);
bool smat_apply_cgupd(gb_fermop *op, const void *dM, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, gb_fermion *out, const double *d_c,
                      const double *d_d, const double *d_cp);
void device_global_sum(gb_context *ctx, double *d_vals, int n); // context.cu: in-stream all-reduce of device scalars
// d / o / dir of a cyclic operator that couples s only to s + dir[c] per chirality c (false if M is anything else); partial-sum buffer
bool smat_tri_onesided(const gb_fermop *op, const void *dM, float d[2][16], float o[2][16], int dir[2]);
double *smat_partials_ensure(gb_fermop *op, size_t n);
void smat_reduce_partials(gb_fermop *op, size_t n, double *d_out);
// Schur CG with the linear algebra folded into the s-space passes (fermop.cu): available for Cayley operators on the dense s-space path
bool cg_fused_available(const gb_fermop *op);
void cg_fused_first(gb_fermop *op, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, const double *d_c, const double *d_d, const double *d_cp);
void cg_fused_rest(gb_fermop *op, const gb_fermion *p, gb_fermion *r, const double *d_c, double *d_d, double *d_cp);

void host_pipe_release(gb_context *ctx);   // dhop_host.cu: scratch of the host-pipelined Dhop of a context that is being destroyed
void op_build_recon12(gb_fermop *op);   // dhop.cu: (re)build and check the two-row link store
size_t halo_exchange_only(gb_fermop *op, const gb_fermion *in, int dag, const void **halo_out = nullptr);   // halo_out[8]: the receive buffers, complete in stream order
void dhop_tslab(gb_fermop *op, const void *const in[2], void *const out[2], int dag, int t0, int nt, cudaStream_t st, const void *const *halo = nullptr,
                int z0 = 0, int nz = -1);
// improved staggered operator entry points (stag.cu)
void stag_op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag);
// one leg of the hopping term on full-grid fields: point 0..3 forward mu, 4..7 backward mu (force.cu)
void op_dhop_leg(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int point, int dag);
// composite operator pieces used by the solvers
void op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag);
gb_fermion *op_tmp_half(gb_fermop *op, int i);
gb_fermion *op_tmp_full(gb_fermop *op, int i);
} // namespace gb
