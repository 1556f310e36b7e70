// fermop.cu -- FermionOperator objects (WilsonFermion, DomainWallFermion, MobiusFermion) and the dispatch of
// their entry points onto the kernels.  Composition follows the reference:
//   WilsonFermion  ref: Grid/qcd/action/fermion/implementation/WilsonFermionImplementation.h:114-176,310-342
//   CayleyFermion5D ref: .../CayleyFermion5DImplementation.h:156-329
//   SchurDiagMooeeOperator ref: Grid/algorithms/LinearOperator.h:286-349
#include "fermop.hpp"
#include <cstring>

namespace gb {

gb_fermion *op_tmp_half(gb_fermop *op, int i) {
  GB_REQUIRE(i >= 0 && i < 4, "tmp index");
  if (!op->tmp_h[i]) {
    int rc = gb_fermion_create(op->grid, op->Ls, (gb_precision)op->prec, GB_HALF, &op->tmp_h[i]);
    if (rc != GB_OK) throw Error(rc, gb_last_error());
  }
  return op->tmp_h[i];
}
gb_fermion *op_tmp_full(gb_fermop *op, int i) {
  GB_REQUIRE(i >= 0 && i < 2, "tmp index");
  if (!op->tmp_f[i]) {
    int rc = gb_fermion_create(op->grid, op->Ls, (gb_precision)op->prec, GB_FULL, &op->tmp_f[i]);
    if (rc != GB_OK) throw Error(rc, gb_last_error());
  }
  return op->tmp_f[i];
}

static void check_field(const gb_fermop *op, const gb_fermion *f, int kind, const char *what) {
  GB_REQUIRE(f != nullptr, "null field");
  if (!(f->grid == op->grid && f->Ls == op->Ls && f->prec == op->prec && f->kind == kind))
    throw Error(GB_ERR_INVALID, std::string(what) + ": field is not conformable with the operator (grid / Ls / precision / full-vs-redblack)");
}

// ---- hopping term on fields
static void dhop_full(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dag, const gb_fermion *ax = nullptr, double axa = 1, double axb = 0) {
  const void *ib[2] = {in->block(0), in->block(1)};
  void *ob[2] = {out->block(0), out->block(1)};
  const void *ab[2] = {ax ? ax->block(0) : nullptr, ax ? ax->block(1) : nullptr};
  dhop_blocks(op, ib, ob, 0, 2, dag, ax ? ab : nullptr, axa, axb);
}
// in has parity in->cb, out gets the opposite parity
static void dhop_cb(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dag) {
  const int ip = in->cb, po = 1 - ip;
  const void *ib[2] = {nullptr, nullptr};
  void *ob[2] = {nullptr, nullptr};
  ib[ip] = in->data; ob[po] = out->data;
  dhop_blocks(op, ib, ob, po, 1, dag, nullptr, 1, 0);
  out->cb = po;
}

// ---- 5D coefficient sets (ref: CayleyFermion5DImplementation.h:156-271)
struct Tri { std::vector<double> lower, diag, upper; };
static Tri coef_meooe5d(const gb_fermop *op) {
  const auto &k = op->k; const int Ls = op->Ls;
  Tri t{k.cs, k.bs, k.cs};
  t.upper[Ls - 1] = -op->mass * t.upper[Ls - 1];
  t.lower[0] = -op->mass * t.lower[0];
  return t;
}
static Tri coef_meooedag5d(const gb_fermop *op) {
  const auto &k = op->k; const int Ls = op->Ls;
  Tri t{std::vector<double>(Ls), k.bs, std::vector<double>(Ls)};
  for (int s = 0; s < Ls; s++) {
    if (s == 0) { t.upper[s] = k.cs[s + 1]; t.lower[s] = -op->mass * k.cs[Ls - 1]; }
    else if (s == Ls - 1) { t.upper[s] = -op->mass * k.cs[0]; t.lower[s] = k.cs[s - 1]; }
    else { t.upper[s] = k.cs[s + 1]; t.lower[s] = k.cs[s - 1]; }
  }
  return t;
}
static Tri coef_mooee(const gb_fermop *op) {
  const auto &k = op->k; const int Ls = op->Ls;
  Tri t{std::vector<double>(Ls), k.bee, std::vector<double>(Ls)};
  for (int i = 0; i < Ls; i++) { t.upper[i] = -k.cee[i]; t.lower[i] = -k.cee[i]; }
  t.upper[Ls - 1] = -op->mass * t.upper[Ls - 1];
  t.lower[0] = -op->mass * t.lower[0];
  return t;
}
static Tri coef_mooeedag(const gb_fermop *op) {
  const auto &k = op->k; const int Ls = op->Ls;
  Tri t{std::vector<double>(Ls), k.bee, std::vector<double>(Ls)};
  for (int s = 0; s < Ls; s++) {
    if (s == 0) { t.upper[s] = -k.cee[s + 1]; t.lower[s] = op->mass * k.cee[Ls - 1]; }
    else if (s == Ls - 1) { t.upper[s] = op->mass * k.cee[0]; t.lower[s] = -k.cee[s - 1]; }
    else { t.upper[s] = -k.cee[s + 1]; t.lower[s] = -k.cee[s - 1]; }
  }
  return t;
}
static Tri coef_m5d_unit(const gb_fermop *op) { // M5D(psi,chi) of the unpreconditioned M, ref :156-163 ; dag :236-245
  const int Ls = op->Ls;
  Tri t{std::vector<double>(Ls, -1.0), std::vector<double>(Ls, 1.0), std::vector<double>(Ls, -1.0)};
  t.upper[Ls - 1] = op->mass; t.lower[0] = op->mass;
  return t;
}

static void scale_field(gb_fermion *out, double a, const gb_fermion *in) {
  int rc = gb_scale(out, a, in);
  if (rc != GB_OK) throw Error(rc, gb_last_error());
}

static void apply_mooee_like(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, const gb_fermion *w = nullptr, double alpha = 0) {
  // in/out may be half or full fields (Mooee is diagonal in 4D)
  if (op->kind == GB_KIND_WILSON) {
    const double d = 4.0 + op->mass;
    const double f = (which == GB_OP_MOOEE || which == GB_OP_MOOEE_DAG) ? d : 1.0 / d;
    if (w) { int rc = gb_axpby(out, f, alpha, in, w); if (rc != GB_OK) throw Error(rc, gb_last_error()); }
    else scale_field(out, f, in);
    return;
  }
  if (op->use_smat) {
    const void *m = which == GB_OP_MOOEE ? op->sm_mooee : which == GB_OP_MOOEE_DAG ? op->sm_mooeedag : which == GB_OP_MOOEE_INV ? op->sm_mooeeinv
                  : which == GB_OP_MOOEE_INV_DAG ? op->sm_mooeeinvdag : which == GB_OP_MEOOE5D ? op->sm_meooe5d : which == GB_OP_MEOOEDAG5D ? op->sm_meooedag5d : nullptr;
    GB_REQUIRE(m != nullptr, "bad opcode");
    if (smat_apply(op, m, in, nullptr, nullptr, alpha, w, out)) return;
  }
  switch (which) {
  case GB_OP_MOOEE: { Tri t = coef_mooee(op); m5d_apply(op, in, in, out, t.lower, t.diag, t.upper, 0, w, alpha); break; }
  case GB_OP_MOOEE_DAG: { Tri t = coef_mooeedag(op); m5d_apply(op, in, in, out, t.lower, t.diag, t.upper, 1, w, alpha); break; }
  case GB_OP_MOOEE_INV: mooee_inv_apply(op, in, out, 0); break;
  case GB_OP_MOOEE_INV_DAG: mooee_inv_apply(op, in, out, 1); break;
  case GB_OP_MEOOE5D: { Tri t = coef_meooe5d(op); m5d_apply(op, in, in, out, t.lower, t.diag, t.upper, 0, w, alpha); break; }
  case GB_OP_MEOOEDAG5D: { Tri t = coef_meooedag5d(op); m5d_apply(op, in, in, out, t.lower, t.diag, t.upper, 1, w, alpha); break; }
  default: GB_REQUIRE(false, "bad opcode");
  }
}

static void apply_meooe(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dagger) {
  if (op->kind == GB_KIND_WILSON) { dhop_cb(op, in, out, dagger); return; }
  gb_fermion *tmp = op_tmp_half(op, 0);
  if (!dagger) { // Meooe = Dhop_{eo|oe} o Meooe5D   ref :308-317
    apply_mooee_like(op, GB_OP_MEOOE5D, in, tmp);
    dhop_cb(op, tmp, out, 0);
  } else {       // MeooeDag = MeooeDag5D o Dhop^dag   ref :320-329
    dhop_cb(op, in, tmp, 1);
    apply_mooee_like(op, GB_OP_MEOOEDAG5D, tmp, out);
  }
}

// SchurDiagMooeeOperator::Mpc / MpcDag   ref: LinearOperator.h:330-348
static void apply_mpc(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dagger) {
  gb_fermion *t1 = op_tmp_half(op, 1), *t2 = op_tmp_half(op, 2);
  if (op->kind == GB_KIND_CAYLEY && op->use_smat) {
    // s-space operators folded on the host: one dense pass per hop instead of M5D + MooeeInv + M5D
    //   Mpc    = Mooee    - Dhop   [Meooe5D MooeeInv]      Dhop   Meooe5D
    //   MpcDag = MooeeDag - MeooeDag5D Dhop^dag [MooeeInvDag MeooeDag5D] Dhop^dag
    if (!dagger) {
      GB_REQUIRE(smat_apply(op, op->sm_meooe5d, in, nullptr, nullptr, 0, nullptr, t1), "smat");   // t1 = A in
      dhop_cb(op, t1, t2, 0);                                                                     // t2 = Dhop t1
      GB_REQUIRE(smat_apply(op, op->sm_B, t2, nullptr, nullptr, 0, nullptr, t1), "smat");         // t1 = A MooeeInv t2
      t1->cb = t2->cb;
      dhop_cb(op, t1, t2, 0);                                                                     // t2 = Dhop t1
      GB_REQUIRE(smat_apply(op, op->sm_mooee, in, nullptr, nullptr, -1.0, t2, out), "smat");      // out = Mooee in - t2
    } else {
      dhop_cb(op, in, t2, 1);                                                                     // t2 = Dhop^dag in
      GB_REQUIRE(smat_apply(op, op->sm_Bdag, t2, nullptr, nullptr, 0, nullptr, t1), "smat");      // t1 = MooeeInvDag MeooeDag5D t2
      t1->cb = t2->cb;
      dhop_cb(op, t1, t2, 1);                                                                     // t2 = Dhop^dag t1
      GB_REQUIRE(smat_apply(op, op->sm_mooeedag, in, op->sm_negAdag, t2, 0, nullptr, out), "smat"); // out = MooeeDag in - MeooeDag5D t2
    }
    out->cb = in->cb;
    return;
  }
  apply_meooe(op, in, t1, dagger);                                                       // tmp = Meooe in
  apply_mooee_like(op, dagger ? GB_OP_MOOEE_INV_DAG : GB_OP_MOOEE_INV, t1, t2);           // out' = MooeeInv tmp
  apply_meooe(op, t2, t1, dagger);                                                       // tmp = Meooe out'
  apply_mooee_like(op, dagger ? GB_OP_MOOEE_DAG : GB_OP_MOOEE, in, out, t1, -1.0);        // out = Mooee in - tmp (fused axpy)
  out->cb = in->cb;
}

// ---- Schur CG with its linear algebra folded into the s-space passes (solver.cu: cg_schur_device_scalars).
// One iteration of ConjugateGradient on MpcDagMpc (ref: ConjugateGradient.h:151-231, LinearOperator.h:291-348) is
//   t1 = Meooe5D p | hop | t1 = [Meooe5D MooeeInv] t2 | hop | w = Mooee p - t2 | hop^dag | t1 = [MooeeInvDag MeooeDag5D] t2 | hop^dag |
//   q = MooeeDag w - MeooeDag5D t2 ; d = <p, q> ; r -= (c/d) q ; cp = |r|^2 ; psi += (c/d) p ; p = r + (cp/c) p
// Here d = |w|^2 (= <p, MpcDag Mpc p> exactly; it falls out of the pass that writes w), q is never stored (the last s-space pass
// updates r and reduces |r|^2), and the update of psi and p rides on the first s-space pass of the NEXT iteration:
// 25 field passes per iteration instead of 30, 11 launches instead of 14.
bool cg_fused_available(const gb_fermop *op) { return op->kind == GB_KIND_CAYLEY && op->use_smat && op->Uds != nullptr; }
// t1 = Meooe5D p, after (if scalars are given) psi += (c/d) p, p = r + (cp/c) p
void cg_fused_first(gb_fermop *op, gb_fermion *psi, gb_fermion *p, const gb_fermion *r, const double *d_c, const double *d_d, const double *d_cp) {
  gb_fermion *t1 = op_tmp_half(op, 1);
  if (d_c == nullptr) GB_REQUIRE(smat_apply(op, op->sm_meooe5d, p, nullptr, nullptr, 0, nullptr, t1), "smat");
  else GB_REQUIRE(smat_apply_cgupd(op, op->sm_meooe5d, psi, p, r, t1, d_c, d_d, d_cp), "smat");
  t1->cb = p->cb;
}
// the rest of A p with t1 = Meooe5D p in place: *d_d = <p, A p> and r -= (c/d) A p, *d_cp = |r|^2 (both summed over ranks, on the device)
void cg_fused_rest(gb_fermop *op, const gb_fermion *p, gb_fermion *r, const double *d_c, double *d_d, double *d_cp) {
  gb_fermion *t1 = op_tmp_half(op, 1), *t2 = op_tmp_half(op, 2), *w = op_tmp_half(op, 3);
  dhop_cb(op, t1, t2, 0);
  GB_REQUIRE(smat_apply(op, op->sm_B, t2, nullptr, nullptr, 0, nullptr, t1), "smat");
  t1->cb = t2->cb;
  {   // w = Mooee p - Dhop t1 ; d = |w|^2 : as the hop's epilogue where the column kernel can, else hop + streaming pass
    HopEpilogue e; e.kind = 1; e.aux = p; e.Maux = op->sm_mooee; e.d_out = d_d;
    op->hop_epi = &e;
    try { dhop_cb(op, t1, w, 0); } catch (...) { op->hop_epi = nullptr; throw; }
    op->hop_epi = nullptr;
    if (!e.applied) GB_REQUIRE(smat_apply_norm(op, op->sm_mooee, p, -1.0, w, w, d_d), "smat");
  }
  device_global_sum(op->ctx, d_d, 1);
  w->cb = p->cb;
  dhop_cb(op, w, t2, 1);
  GB_REQUIRE(smat_apply(op, op->sm_Bdag, t2, nullptr, nullptr, 0, nullptr, t1), "smat");
  t1->cb = t2->cb;
  {   // r -= (c/d)(MooeeDag w - MeooeDag5D Dhop^dag t1) ; cp = |r|^2 : likewise
    HopEpilogue e; e.kind = 2; e.aux = w; e.r = r; e.Maux = op->sm_mooeedag; e.Mhop = op->sm_negAdag; e.d_c = d_c; e.d_d = d_d; e.d_out = d_cp;
    op->hop_epi = &e;
    try { dhop_cb(op, t1, t2, 1); } catch (...) { op->hop_epi = nullptr; throw; }
    op->hop_epi = nullptr;
    if (!e.applied) GB_REQUIRE(smat_apply_rupd(op, op->sm_mooeedag, w, op->sm_negAdag, t2, r, d_c, d_d, d_cp), "smat");
  }
  device_global_sum(op->ctx, d_cp, 1);
}

void op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag) {
  GB_REQUIRE(op && in && out, "null argument");
  GB_REQUIRE(in != out, "in and out must be distinct fields");
  if (op->kind == GB_KIND_STAGGERED) { stag_op_apply(op, which, in, out, dag ? 1 : 0); return; }
  GB_REQUIRE(op->Uds != nullptr, "operator has no gauge field: call ImportGauge first");
  dag = dag ? 1 : 0;
  switch (which) {
  case GB_OP_DHOP:
    check_field(op, in, GB_FULL, "Dhop"); check_field(op, out, GB_FULL, "Dhop");
    dhop_full(op, in, out, dag);
    break;
  case GB_OP_DHOP_OE: // ref: WilsonFermion5DImplementation.h:415-424 asserts in.Checkerboard()==Even
    check_field(op, in, GB_HALF, "DhopOE"); check_field(op, out, GB_HALF, "DhopOE");
    GB_REQUIRE(in->cb == GB_EVEN, "DhopOE needs an Even-checkerboard input");
    dhop_cb(op, in, out, dag);
    break;
  case GB_OP_DHOP_EO:
    check_field(op, in, GB_HALF, "DhopEO"); check_field(op, out, GB_HALF, "DhopEO");
    GB_REQUIRE(in->cb == GB_ODD, "DhopEO needs an Odd-checkerboard input");
    dhop_cb(op, in, out, dag);
    break;
  case GB_OP_DW: // DW = Dhop + (4 - M5)
    check_field(op, in, GB_FULL, "DW"); check_field(op, out, GB_FULL, "DW");
    dhop_full(op, in, out, dag, in, 1.0, 4.0 - op->M5);
    break;
  case GB_OP_DMINUS: case GB_OP_DMINUS_DAG: { // chi_s = psi_s - cs[s] DW(psi)_s   ref: CayleyFermion5DImplementation.h:132-153
    check_field(op, in, GB_FULL, "Dminus"); check_field(op, out, GB_FULL, "Dminus");
    if (op->kind == GB_KIND_WILSON) { scale_field(out, 1.0, in); break; }   // ref: FermionOperator.h:172-173 (chi = psi)
    gb_fermion *t = op_tmp_full(op, 0);
    dhop_full(op, in, t, which == GB_OP_DMINUS_DAG, in, 1.0, 4.0 - op->M5);   // t = DW psi
    std::vector<double> none(op->Ls, 0.0), mcs(op->Ls);
    for (int s = 0; s < op->Ls; s++) mcs[s] = -op->k.cs[s];
    m5d_apply(op, t, t, out, none, mcs, none, 0, in, 1.0);                    // out = -cs[s] t + psi
    break;
  }
  case GB_OP_M:
    check_field(op, in, GB_FULL, "M"); check_field(op, out, GB_FULL, "M");
    if (op->kind == GB_KIND_WILSON) { dhop_full(op, in, out, 0, in, 1.0, 4.0 + op->mass); break; }
    {
      gb_fermion *Din = op_tmp_full(op, 0);
      apply_mooee_like(op, GB_OP_MEOOE5D, in, Din);
      dhop_full(op, Din, out, 0, Din, 1.0, 4.0 - op->M5);         // chi = DW(Din)
      Tri t = coef_m5d_unit(op);
      m5d_apply(op, in, out, out, t.lower, t.diag, t.upper, 0, in, 1.0); // chi = chi + psi + (-1|m) P psi_{s+-1}
    }
    break;
  case GB_OP_MDAG:
    check_field(op, in, GB_FULL, "Mdag"); check_field(op, out, GB_FULL, "Mdag");
    if (op->kind == GB_KIND_WILSON) { dhop_full(op, in, out, 1, in, 1.0, 4.0 + op->mass); break; }
    {
      gb_fermion *Din = op_tmp_full(op, 0);
      dhop_full(op, in, Din, 1, in, 1.0, 4.0 - op->M5);           // Din = DW^dag psi
      apply_mooee_like(op, GB_OP_MEOOEDAG5D, Din, out);
      Tri t = coef_m5d_unit(op);
      m5d_apply(op, in, out, out, t.lower, t.diag, t.upper, 1, in, 1.0);
    }
    break;
  case GB_OP_MEOOE: case GB_OP_MEOOE_DAG:
    check_field(op, in, GB_HALF, "Meooe"); check_field(op, out, GB_HALF, "Meooe");
    apply_meooe(op, in, out, which == GB_OP_MEOOE_DAG);
    break;
  case GB_OP_MOOEE: case GB_OP_MOOEE_DAG: case GB_OP_MOOEE_INV: case GB_OP_MOOEE_INV_DAG: case GB_OP_MEOOE5D: case GB_OP_MEOOEDAG5D:
    check_field(op, in, in->kind, "Mooee"); check_field(op, out, in->kind, "Mooee");
    if ((which == GB_OP_MEOOE5D || which == GB_OP_MEOOEDAG5D) && op->kind == GB_KIND_WILSON) GB_REQUIRE(false, "Meooe5D is a 5D operator");
    apply_mooee_like(op, which, in, out);
    out->cb = in->cb;
    break;
  case GB_OP_MPC: case GB_OP_MPC_DAG:
    check_field(op, in, GB_HALF, "Mpc"); check_field(op, out, GB_HALF, "Mpc");
    apply_mpc(op, in, out, which == GB_OP_MPC_DAG);
    break;
  case GB_OP_HERMOP: {
    check_field(op, in, GB_HALF, "HermOp"); check_field(op, out, GB_HALF, "HermOp");
    gb_fermion *t3 = op_tmp_half(op, 3);
    apply_mpc(op, in, t3, 0);
    apply_mpc(op, t3, out, 1);
    break;
  }
  default:
    GB_REQUIRE(false, "unknown opcode");
  }
}

static gb_fermop *make_op(gb_grid *g, const gb_gauge *Umu, int kind, int Ls, double mass, double M5, double b, double c, const double *ph) {
  GB_REQUIRE(g && Umu, "null argument");
  GB_REQUIRE(Ls >= 1, "Ls must be >= 1");
  gb_fermop *op = new gb_fermop();
  op->grid = g; op->ctx = g->ctx; op->kind = kind; op->prec = Umu->prec; op->Ls = Ls; op->mass = mass; op->M5 = M5;
  if (ph) std::memcpy(op->phases, ph, sizeof(double) * 8);
  if (kind == GB_KIND_CAYLEY) { GB_REQUIRE(Ls >= 2, "Cayley operators need Ls >= 2"); op->k = cayley_coeffs(Ls, mass, M5, b, c); }
  // GB_SELF_HALO=<bitmask of dimensions>, read at creation: also routes those UNdecomposed dimensions through the halo path
  // (gauge faces, pack -> store into this rank's own receive buffers -> epoch flags -> interior / exterior or semi-fused hop), so
  // that one GPU exercises the whole multi-rank machinery (tests/test_gpu_self_halo.py)
  const int self_mask = getenv("GB_SELF_HALO") ? atoi(getenv("GB_SELF_HALO")) & 15 : 0;
  for (int d = 0; d < 4; d++) if (g->mpi[d] > 1 || ((self_mask >> d) & 1)) op->comm_dim_mask |= 1 << d;
  // default rasterisation: whole y, 8 z-planes at a time, all t ("L2 blocking")
  op->By = 0; op->Bz = 8; op->Bt = 0;
  if (kind == GB_KIND_CAYLEY && (Ls == 8 || Ls == 12 || Ls == 16)) {
    Tri a = coef_meooe5d(op), ad = coef_meooedag5d(op), c = coef_mooee(op), cd = coef_mooeedag(op);
    SMat A = smat_m5d(Ls, a.lower, a.diag, a.upper, 0), Ad = smat_m5d(Ls, ad.lower, ad.diag, ad.upper, 1);
    SMat C = smat_m5d(Ls, c.lower, c.diag, c.upper, 0), Cd = smat_m5d(Ls, cd.lower, cd.diag, cd.upper, 1);
    SMat Mi = smat_mooee_inv(op->k, 0), Mid = smat_mooee_inv(op->k, 1);
    op->sm_meooe5d = smat_device(op, A); op->sm_meooedag5d = smat_device(op, Ad);
    op->sm_mooee = smat_device(op, C); op->sm_mooeedag = smat_device(op, Cd);
    op->sm_mooeeinv = smat_device(op, Mi); op->sm_mooeeinvdag = smat_device(op, Mid);
    op->sm_B = smat_device(op, smat_mul(A, Mi));
    op->sm_Bdag = smat_device(op, smat_mul(Mid, Ad));
    op->sm_negAdag = smat_device(op, smat_scale(Ad, -1.0));
    op->use_smat = op->sm_B != nullptr && op->sm_Bdag != nullptr;   // the dense s-space path needs its matrices on the device
  }
  try { op_import_gauge(op, Umu); } catch (...) { delete op; throw; }
  return op;
}

} // namespace gb

using namespace gb;

extern "C" {
int gb_op_create_wilson(gb_grid *g, const gb_gauge *Umu, double mass, const double *ph, gb_fermop **out) {
  GB_API_BEGIN
  *out = make_op(g, Umu, GB_KIND_WILSON, 1, mass, 0, 1, 0, ph);
  GB_API_END
}
int gb_op_create_dwf(gb_grid *g, const gb_gauge *Umu, int Ls, double mass, double M5, const double *ph, gb_fermop **out) {
  GB_API_BEGIN
  *out = make_op(g, Umu, GB_KIND_CAYLEY, Ls, mass, M5, 1.0, 0.0, ph); // ref: DomainWallFermion.h:125-131 (b=1,c=0)
  GB_API_END
}
int gb_op_create_mobius(gb_grid *g, const gb_gauge *Umu, int Ls, double mass, double M5, double b, double c, const double *ph, gb_fermop **out) {
  GB_API_BEGIN
  *out = make_op(g, Umu, GB_KIND_CAYLEY, Ls, mass, M5, b, c, ph);     // ref: MobiusFermion.h:63-66
  GB_API_END
}
int gb_op_import_gauge(gb_fermop *op, const gb_gauge *Umu) {
  GB_API_BEGIN
  GB_REQUIRE(op && op->kind != GB_KIND_STAGGERED, "staggered operators take thin and fat links: gb_op_import_gauge_staggered");
  op_import_gauge(op, Umu);
  GB_API_END
}
int gb_op_destroy(gb_fermop *op) {
  if (!op) return GB_OK;
  cudaFree(op->Uds);
  cudaFree(op->Uds12);
  cudaFree(op->stag_links);
  for (int i = 0; i < 8; i++) { if (op->halo_send[i]) cudaFree(op->halo_send[i]); if (op->halo_recv[i]) cudaFree(op->halo_recv[i]); }
  p2p_teardown(op);
  for (void *p : op->smat_allocs) cudaFree(p);
  if (op->smat_partials) cudaFree(op->smat_partials);
  for (auto *f : op->tmp_h) gb_fermion_destroy(f);
  for (auto *f : op->tmp_f) gb_fermion_destroy(f);
  delete op;
  return GB_OK;
}
int gb_op_Ls(const gb_fermop *op) { return op->Ls; }
int gb_op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag) {
  GB_API_BEGIN
  op_apply(op, which, in, out, dag);
  GB_API_END
}
int gb_op_halo_exchange(gb_fermop *op, const gb_fermion *in, int dag, int64_t *bytes_sent) {
  GB_API_BEGIN
  const size_t b = halo_exchange_only(op, in, dag ? 1 : 0);
  if (bytes_sent) *bytes_sent = (int64_t)b;
  GB_API_END
}
int gb_op_set_tiling(gb_fermop *op, int by, int bz, int bt) {
  op->By = by; op->Bz = bz; op->Bt = bt;
  op->col_n = bz;   // column-sweep kernel: z-planes per column
  return GB_OK;
}
int gb_op_set_fast_kernel(gb_fermop *op, int enable) {
  op->disable_fast = enable == 0;
  op->no_col = enable == 2;
  op->use_smat = enable != 0 && op->sm_B != nullptr;
  return GB_OK;
}
int gb_op_set_link_reconstruct(gb_fermop *op, int nreal) {
  GB_API_BEGIN
  GB_REQUIRE(op != nullptr, "null operator");
  GB_REQUIRE(nreal == 18 || nreal == 12, "links are stored with 18 reals (full) or 12 (two rows, third row rebuilt)");
  GB_REQUIRE(op->kind != GB_KIND_STAGGERED, "link reconstruction serves the Wilson-type operators (fat staggered links are not unitary)");
  if (nreal == 18) { op->recon12 = 0; return GB_OK; }
  op->recon12 = 1;
  try { op_build_recon12(op); } catch (...) { op->recon12 = 0; throw; }
  GB_API_END
}
int gb_op_set_halo_compression(gb_fermop *op, int on) {
  GB_API_BEGIN
  GB_REQUIRE(op != nullptr, "null operator");
  GB_REQUIRE(op->kind != GB_KIND_STAGGERED, "compressed halos serve the Wilson-type operators (the reference's ...ImplFH / ...ImplDF: fp32 compute with half comms, fp64 compute with fp32 comms)");
  op->halo_lowp = on != 0;
  GB_API_END
}
int gb_op_set_overlap(gb_fermop *op, int overlap) {
  op->overlap_comms = overlap != 0;
  op->no_semifused = overlap == 2;
  return GB_OK;
}
}
