/* =====================================================================================
 * gridb200.h -- C ABI of the B200-native Dirac hopping term / Schur-CG library (libgridb200.so)
 *
 * This is the drop-in boundary for paboyle/Grid's data-parallel hot path.  Every entry point names
 * the reference interface it replaces ("ref:" paths are relative to the reference tree).  Signatures
 * carry only plain pointers, sizes and opaque handles: no C++ or torch types.
 *
 * Host-side data layouts at the boundary (identical to what Grid's unvectorizeToLexOrdArray /
 * vectorizeFromLexOrdArray, ref: Grid/lattice/Lattice_transfer.h:1123,1218, or a peekLocalSite loop,
 * ref: Grid/lattice/Lattice_peekpoke.h:159-226, produce on each rank for its LOCAL lattice):
 *   4D local lexicographic site index   i4 = x + Lx*(y + Ly*(z + Lz*t))       ref: Grid/util/Lexicographic.h:18-27
 *   5D                                  i5 = s + Ls*i4   (s fastest)          ref: Grid/qcd/utils/SpaceTimeGrid.cc:49-63
 *   fermion site object   psi[spin 4][colour 3] {re,im}                       (SpinColourVector)
 *   gauge site object     U[mu 4][row 3][col 3] {re,im}, (U chi)_row = sum_col U[row][col] chi_col
 *                                                                             ref: Grid/qcd/QCD.h:106
 *   half (red-black) fields: icb = s + Ls*((x>>1) + (Lx/2)*(y + Ly*(z + Lz*t))), parity (x+y+z+t)&1 of the
 *   GLOBAL coordinate, Even=0 / Odd=1, s ignored                             ref: Grid/cartesian/Cartesian_red_black.h:68-76,271-286
 * The device layout is private to the library.
 *
 * Error model: the reference only asserts/exits (ref: ConjugateGradient.h:97,225,254).  Here every call
 * returns GB_OK (0) or a negative gb_status; gb_last_error() gives the message.  The C++ wrappers in
 * gridb200.hpp assert on failure to mimic the reference.  There is NO CPU fallback: without a CUDA device
 * gb_context_create fails with GB_ERR_NO_DEVICE and no compute entry point can be reached.
 *
 * Threading: like the reference (one global computeStream, mutable statics; ref: Grid/threads/Accelerator.h:109-110)
 * a context must be driven by one host thread at a time.
 * ===================================================================================== */
#ifndef GRIDB200_H
#define GRIDB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gb_context gb_context; /* device + streams + communicator   (ref: Grid_init / acceleratorInit, Grid/util/Init.cc:300-560) */
typedef struct gb_grid gb_grid;       /* GridCartesian 4D + its red-black + 5D companions (ref: SpaceTimeGrid.cc:36-78) */
typedef struct gb_fermion gb_fermion; /* LatticeFermion{F,D} on a full or red-black, 4D or 5D grid (ref: Grid/lattice/Lattice_base.h) */
typedef struct gb_gauge gb_gauge;     /* LatticeGaugeField{F,D} (ref: Grid/qcd/QCD.h:106) */
typedef struct gb_fermop gb_fermop;   /* FermionOperator<Impl> (ref: Grid/qcd/action/fermion/FermionOperator.h:40-192) */

typedef enum { GB_OK = 0, GB_ERR_INVALID = -1, GB_ERR_CUDA = -2, GB_ERR_NO_DEVICE = -3, GB_ERR_NOT_CONVERGED = -4, GB_ERR_COMM = -5 } gb_status;
typedef enum { GB_F32 = 0, GB_F64 = 1 } gb_precision;
typedef enum { GB_EVEN = 0, GB_ODD = 1 } gb_parity;       /* ref: Cartesian_red_black.h:34-37 */
typedef enum { GB_FULL = 0, GB_HALF = 1 } gb_gridkind;    /* GridCartesian vs GridRedBlackCartesian */

/* Operator entry points of FermionOperator / SchurDiagMooeeOperator, selected by code so that one
 * C symbol serves them all.  ref: FermionOperator.h:63-78 ; LinearOperator.h:286-349 */
typedef enum {
  GB_OP_DHOP = 0,          /* Dhop(in,out,dag)    full grid          ref: WilsonFermion5DImplementation.h:437-445 */
  GB_OP_DHOP_OE = 1,       /* DhopOE(in,out,dag)  in Even -> out Odd ref: :415-424 */
  GB_OP_DHOP_EO = 2,       /* DhopEO(in,out,dag)  in Odd -> out Even ref: :426-435 */
  GB_OP_M = 3,             /* ref: CayleyFermion5DImplementation.h:274-286 ; WilsonFermionImplementation.h:114-119 */
  GB_OP_MDAG = 4,          /* ref: :289-304 */
  GB_OP_MEOOE = 5,         /* ref: :308-317 (dispatches on in.Checkerboard()) */
  GB_OP_MEOOE_DAG = 6,     /* ref: :320-329 */
  GB_OP_MOOEE = 7,         /* ref: :191-204 */
  GB_OP_MOOEE_DAG = 8,     /* ref: :206-233 */
  GB_OP_MOOEE_INV = 9,     /* ref: CayleyFermion5Dcache.h:117-172 */
  GB_OP_MOOEE_INV_DAG = 10,/* ref: CayleyFermion5Dcache.h:174-230 */
  GB_OP_MPC = 11,          /* SchurDiagMooeeOperator::Mpc      ref: LinearOperator.h:330-339 */
  GB_OP_MPC_DAG = 12,      /* SchurDiagMooeeOperator::MpcDag   ref: :340-348 */
  GB_OP_HERMOP = 13,       /* SchurOperatorBase::HermOp = MpcDagMpc  ref: :291-307 */
  GB_OP_DW = 14,           /* WilsonFermion5D::DW = Dhop + (4-M5)    ref: WilsonFermion5DImplementation.h:447-452 */
  GB_OP_MEOOE5D = 15,      /* ref: CayleyFermion5DImplementation.h:165-174 */
  GB_OP_MEOOEDAG5D = 16,   /* ref: :248-271 */
  GB_OP_DMINUS = 17,       /* CayleyFermion5D::Dminus: chi_s = psi_s - cs[s] DW psi_s (identity for 4D operators)  ref: :132-142 */
  GB_OP_DMINUS_DAG = 18    /* ref: :143-153 */
} gb_opcode;

/* ---------------------------------------------------------------- context / runtime
 * replaces Grid_init -> acceleratorInit (device by local rank, compute+copy streams)
 * ref: Grid/threads/Accelerator.cc:19-110 ; Grid/util/Init.cc:300-560 */
int gb_context_create(int device, gb_context **out);
int gb_context_destroy(gb_context *ctx);
const char *gb_last_error(void);
int gb_device_count(void);
int gb_synchronize(gb_context *ctx);                       /* ref: accelerator_barrier, Accelerator.h:233-245 */
/* CUDA-event stopwatch on the library's compute stream (GridStopWatch analogue, ref: Grid/perfmon/Timer.h:83) */
int gb_timer_start(gb_context *ctx);
int gb_timer_stop(gb_context *ctx, double *elapsed_ms);
/* number of kernels this library has launched on ctx since creation (evidence for "gpu_launches") */
int64_t gb_launch_count(gb_context *ctx);
/* write > L2-capacity bytes so that the next timed call starts with a cold L2 */
int gb_flush_l2(gb_context *ctx);

/* Multi-GPU: one process per GPU.  Replaces CartesianCommunicator (MPI_Cart_create + cudaIpc peer buffers,
 * ref: Grid/communicator/Communicator_mpi3.cc:226,390-461 ; SharedMemoryMPI.cc:598-680) with NCCL over NVLink.
 * The caller bootstraps: rank 0 calls gb_comm_unique_id, broadcasts the 128 bytes (MPI_Bcast / torch.distributed),
 * every rank calls gb_comm_init.  Rank -> processor coordinate is lexicographic with dimension 0 fastest
 * (ref: Communicator_base.h ShiftedRanks / Lexicographic::CoorFromIndex). */
#define GB_UNIQUE_ID_BYTES 128
int gb_comm_unique_id(void *id_out);
int gb_comm_init(gb_context *ctx, int rank, int nranks, const void *id);
int gb_comm_rank(gb_context *ctx, int *rank, int *nranks);
/* GlobalSum of host doubles, ref: Communicator_mpi3.cc:299-307 */
int gb_comm_global_sum(gb_context *ctx, double *vals, int n);
int gb_comm_barrier(gb_context *ctx);

/* ---------------------------------------------------------------- grids
 * gdims = global 4D lattice (--grid), mpi = processor grid (--mpi).  ref: SpaceTimeGrid::makeFourDimGrid,
 * makeFiveDimGrid, make*RedBlackGrid (Grid/qcd/utils/SpaceTimeGrid.cc:36-78).  The fifth dimension is never
 * decomposed (ref: WilsonFermion5DImplementation.h:80-88).  Local extents must be even. */
int gb_grid_create(gb_context *ctx, const int gdims[4], const int mpi[4], gb_grid **out);
/* host-only geometry of one rank: local extents, global origin, neighbour ranks nbr[2*mu + {0:forward,1:backward}]
 * (ref: CartesianCommunicator::ShiftedRanks, Communicator_base.h) -- needs no device */
int gb_geometry_query(const int gdims[4], const int mpi[4], int rank, int ldims[4], int origin[4], int nbr[8]);
int gb_grid_destroy(gb_grid *g);
int gb_grid_local_dims(const gb_grid *g, int ldims[4]);
int gb_grid_local_origin(const gb_grid *g, int origin[4]);   /* global coordinate of local site 0 */

/* ---------------------------------------------------------------- fermion fields
 * Ls = 1 for 4D fields.  kind GB_HALF = field on the red-black grid; its Checkerboard() is set by
 * gb_fermion_set_checkerboard or by the operation that fills it (ref: Lattice_base.h Checkerboard()). */
int gb_fermion_create(gb_grid *g, int Ls, gb_precision prec, gb_gridkind kind, gb_fermion **out);
/* LatticeStaggeredFermion{F,D}: ColourVector sites chi[colour 3] {re,im}, 4D only (ref: StaggeredImpl.h:60-75 SiteSpinor =
 * iScalar<iScalar<iVector<Simd,Nc>>>).  Every gb_fermion_* / BLAS / reduction / checkerboard entry point accepts it. */
int gb_staggered_fermion_create(gb_grid *g, gb_precision prec, gb_gridkind kind, gb_fermion **out);
int gb_fermion_destroy(gb_fermion *f);
int gb_fermion_checkerboard(const gb_fermion *f);
int gb_fermion_set_checkerboard_tag(gb_fermion *f, int cb);
int64_t gb_fermion_local_sites(const gb_fermion *f);          /* number of 5D sites held locally */
/* host lexicographic array (layout at top of file) -> device; host_prec may differ from the field's.
 * ref: vectorizeFromLexOrdArray, Lattice_transfer.h:1218-1260 */
int gb_fermion_import(gb_fermion *f, const void *host, gb_precision host_prec);
/* ref: unvectorizeToLexOrdArray, Lattice_transfer.h:1123-1166 */
int gb_fermion_export(const gb_fermion *f, void *host, gb_precision host_prec);
/* ref: pickCheckerboard / setCheckerboard, Lattice_transfer.h:50-86 */
int gb_pick_checkerboard(int cb, gb_fermion *half, const gb_fermion *full);
int gb_set_checkerboard(gb_fermion *full, const gb_fermion *half);
/* ref: precisionChange, Lattice_transfer.h:1461-1492 */
int gb_precision_change(gb_fermion *out, const gb_fermion *in);
/* synthetic sources generated on the device, decomposition independent (keyed by global site):
 * uniform [0,1) real & imaginary parts like Grid's random() (ref: Benchmark_dwf_fp32.cc:163) */
int gb_fermion_random(gb_fermion *f, uint64_t seed);

/* BLAS-1 / reductions.  ref: Grid/lattice/Lattice_arith.h:231-258 ; Lattice_reduction.h:256-372.
 * Site products in working precision, lattice sums in double, fixed reduction order (bit-reproducible),
 * then GlobalSum over ranks. */
int gb_zero(gb_fermion *z);
int gb_copy(gb_fermion *z, const gb_fermion *x);
int gb_scale(gb_fermion *z, double a, const gb_fermion *x);
int gb_axpy(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y);             /* z = a x + y */
int gb_axpby(gb_fermion *z, double a, double b, const gb_fermion *x, const gb_fermion *y);  /* z = a x + b y */
int gb_axpy_norm(gb_fermion *z, double a, const gb_fermion *x, const gb_fermion *y, double *norm2_z);
int gb_norm2(const gb_fermion *x, double *out);
int gb_inner_product(const gb_fermion *l, const gb_fermion *r, double out_re_im[2]);         /* sum conj(l) r */

/* ---------------------------------------------------------------- gauge fields */
int gb_gauge_create(gb_grid *g, gb_precision prec, gb_gauge **out);
int gb_gauge_destroy(gb_gauge *u);
int gb_gauge_import(gb_gauge *u, const void *host, gb_precision host_prec);  /* [V4 local][4][3][3] complex */
int gb_gauge_export(const gb_gauge *u, void *host, gb_precision host_prec);
/* random SU(3) links on the device: exp(Ta(gaussian)) restating SU<Nc>::HotConfiguration
 * (ref: Grid/qcd/utils/GaugeGroup.h:332-349); decomposition independent */
int gb_gauge_random(gb_gauge *u, uint64_t seed);
int gb_gauge_unit(gb_gauge *u);

/* NERSC gauge configurations.  ref: Grid/parallelIO/NerscIO.h:63-290, MetaData.h:143-215.
 * Reading validates like NerscIO::readConfiguration: checksum (sum of the payload's 32-bit words) exact, plaquette within 1e-5
 * and link trace within 1e-6 of the header, else GB_ERR_INVALID (the reference exits / asserts); two-row "4D_SU3_GAUGE" links
 * get their third row reconstructed; IEEE64BIG / IEEE32BIG / IEEE64 / IEEE32 payloads.  Writing is always IEEE64BIG (NerscIO.h:262),
 * two_row = 1 drops the third row.  The *_host entry points work on a GLOBAL lexicographic [V][4][3][3] complex double array and
 * need no device; gb_gauge_read_nersc has every rank read and validate the file and import its local block. */
typedef struct {
  int dimension[4];
  double link_trace, plaquette;      /* header values */
  uint32_t checksum;
  char data_type[64], floating_point[32], ensemble_id[64], ensemble_label[64];
  int sequence_number;
  int64_t data_start;                /* byte offset of the payload */
  double computed_link_trace, computed_plaquette;   /* recomputed from the payload (0 for a header-only query) */
  uint32_t computed_checksum;
} gb_nersc_header;
int gb_nersc_read_host(const char *path, double *U_out /* NULL: header only */, gb_nersc_header *hdr);
int gb_nersc_write_host(const char *path, const double *U, const int dims[4], int two_row, const char *ens_label, const char *ens_id, int sequence_number);
int gb_gauge_read_nersc(gb_gauge *Umu, const char *path, gb_nersc_header *hdr);
int gb_gauge_write_nersc(const gb_gauge *Umu, const char *path, int two_row, const char *ens_label, const char *ens_id, int sequence_number);

/* ---------------------------------------------------------------- fermion operators
 * ctor analogues: WilsonFermion(Umu,Grid,RBGrid,mass) ref: WilsonFermion.h:139-142
 *                 DomainWallFermion(Umu,FGrid,FrbGrid,UGrid,UrbGrid,mass,M5) ref: DomainWallFermion.h:108-134
 *                 MobiusFermion(...,mass,M5,b,c) ref: MobiusFermion.h:45-71
 * boundary_phases: 4 complex numbers {re,im} (ImplParams::boundary_phases, ref: WilsonImpl.h:143-170) or NULL = periodic.
 * The operator owns its double-stored links (ref: WilsonFermion5D.h:188-207); ImportGauge copies. */
int gb_op_create_wilson(gb_grid *g, const gb_gauge *Umu, double mass, const double *boundary_phases, gb_fermop **out);
int gb_op_create_dwf(gb_grid *g, const gb_gauge *Umu, int Ls, double mass, double M5, const double *boundary_phases, gb_fermop **out);
int gb_op_create_mobius(gb_grid *g, const gb_gauge *Umu, int Ls, double mass, double M5, double b, double c, const double *boundary_phases, gb_fermop **out);
int gb_op_import_gauge(gb_fermop *op, const gb_gauge *Umu);    /* ref: WilsonFermion5DImplementation.h:149-181 */
/* ImprovedStaggeredFermion(Uthin,Ufat,Fgrid,Hgrid,mass,c1,c2,u0)  ref: ImprovedStaggeredFermion.h:115-121 ;
 * ImportGauge(Uthin,Ufat): staggered phases, one-link (fat) and Naik three-link (thin) double store, c1/u0 and c2/u0^3
 * ref: StaggeredImpl.h:105-162, ImprovedStaggeredFermionImplementation.h:137-167.
 * gb_op_apply on it serves Dhop/DhopOE/DhopEO (dag = overall minus sign, ref: StaggeredKernelsImplementation.h:117-119),
 * M/Mdag, Meooe(+Dag), Mooee(+Dag) = mass, MooeeInv(+Dag) = 1/mass, and GB_OP_MPC = GB_OP_MPC_DAG = GB_OP_HERMOP =
 * SchurStaggeredOperator::Mpc = mass^2 - Meooe Meooe (ref: LinearOperator.h:543-584); gb_cg_schur runs CG on that.
 * Decomposed lattices: the Naik term reaches three sites, so every split dimension carries three-deep halos of the input
 * field (packed and exchanged per hop) and of U_mu for the double store (ref: displacements +-1, +-3,
 * instantiation/ImprovedStaggeredFermionInstantiation.cc:33-34 ; Stencil.h:709 needs local extents > 3).  gb_op_set_overlap:
 * 1 (default) interior sites while the faces travel, then the exterior sites; 0 exchange, then one launch.
 * Environment GB_STAG_SELF_HALO=<bitmask of dimensions>, read at creation: also routes those UNdecomposed dimensions through the
 * pack / exchange (with the rank itself) / halo-lookup path -- a test knob that lets one GPU exercise the decomposed code. */
int gb_op_create_staggered(gb_grid *g, const gb_gauge *Uthin, const gb_gauge *Ufat, double mass, double c1, double c2, double u0, gb_fermop **out);
int gb_op_import_gauge_staggered(gb_fermop *op, const gb_gauge *Uthin, const gb_gauge *Ufat);
int gb_op_destroy(gb_fermop *op);
int gb_op_Ls(const gb_fermop *op);
/* which: gb_opcode.  dag only matters for GB_OP_DHOP*, GB_OP_DW.  Checkerboard asserts as in the reference
 * (DhopOE needs in.cb==Even ...).  in and out must be distinct fields of the operator's precision. */
int gb_op_apply(gb_fermop *op, int which, const gb_fermion *in, gb_fermion *out, int dag);
/* Dhop(in,out,dag) on HOST-resident full-lattice fields (layout at top of file), for callers whose Lattice objects live in
 * host memory (the reference's CPU build; ref: FermionOperator.h:72 Dhop + Lattice_transfer.h:1123,1218 for the layout).
 * Pipelined over t-slices -- H2D of slice t+1, the hop of slice t and D2H of slice t-1 overlap on separate streams, so the call
 * costs one direction of PCIe traffic.  On z / t decomposed lattices (each rank passes its LOCAL volume; collective over the
 * ranks like Dhop itself) the sites the neighbours need go in first (t-slices 0 and Lt-1, z planes 0 and Lz-1 of every slice by
 * one strided copy per face), the halo exchange runs once, and the slices stream through with every slab hop reading the receive
 * buffers (GB_HOST_PIPE_DECOMP=0: import, hop, export; also the form of x / y splits).  Pinned host memory recommended. */
int gb_op_dhop_host(gb_fermop *op, const void *host_in, void *host_out, gb_precision host_prec, int dag);
/* The face exchange of one full-lattice Dhop on its own: project + send every face of both parities, then wait for the
 * neighbours' faces (no hopping kernel).  The halo microbenchmark of SURVEY 8(d)/(e): time N calls with gb_timer_start/stop and
 * divide bytes_sent (this rank, all split directions, both senses) by the time.  ref: benchmarks/Benchmark_comms.cc:105-162,
 * Grid/stencil/Stencil.h:367-430.  bytes_sent = 0 on an undecomposed lattice. */
int gb_op_halo_exchange(gb_fermop *op, const gb_fermion *in, int dag, int64_t *bytes_sent);
/* tuning knob of the hopping kernel's CTA rasterisation (z/t blocking for L2 reuse); 0 = default */
int gb_op_set_tiling(gb_fermop *op, int block_y, int block_z, int block_t);
/* 1 (default): the hop overlaps the face exchange (ref: --comms-overlap, WilsonFermion5DImplementation.h:320-384); on z/t
 *    decomposed fp32 lattices with peer access that is ONE launch after the pack+send kernel: local legs everywhere, and the
 *    CTAs owning surface sites (rasterised last) acquire the neighbours' flags and add the halo legs ("semi-fused");
 * 2: overlapped in the reference's form: interior kernel, then an accumulate pass over the surface slabs;
 * 0: exchange then compute (ref: DhopInternalSerialComms, :388-411) */
int gb_op_set_overlap(gb_fermop *op, int overlap);
/* Link storage of the hopping term: 18 (default) = the full doubled 3x3 store with -1/2 and the boundary phases folded in;
 * 12 = two rows of the bare SU(3) matrix per link, the third row rebuilt in registers as conj(row0 x row1) and the folded-in factor
 * applied to the product (the north star's "optional 12-real SU(3) link reconstruction"; the reference always stores 18 reals,
 * WilsonImpl.h:127-171 DoubleStore).  Trades a third of the link bytes for ~60 flops per link; it takes the generic kernel (any Ls,
 * both precisions, every multi-rank form), so it pays where links dominate the traffic -- 4D Wilson, small Ls -- not at Ls = 16 where the
 * tuned kernels already reuse each link across the fifth dimension.  Needs a gauge field (call after creation / ImportGauge; a later
 * ImportGauge rebuilds it) whose links are special unitary to working precision: otherwise GB_ERR_INVALID and the full store stays. */
int gb_op_set_link_reconstruct(gb_fermop *op, int nreal);
/* Compressed halos ("half-precision comms"): 1 = the projected half spinors of every face travel one precision below the
 * operator's -- an fp32 operator sends bf16, an fp64 operator sends fp32 -- and are widened by the consuming leg; the arithmetic stays
 * in the operator's precision and sites without an off-rank leg are bit-identical.  0 (default) = uncompressed.
 * ref: FermionOperatorImpl.h:96-137 (LowerPrecisionMapper: vComplexF -> vComplexH, vComplexD -> vComplexF; CoeffRealHalfComms),
 * WilsonImpl.h:59-65 (SiteHalfCommSpinor), WilsonCompressor.h:244-306 (compressor on _HCspinor; Decompress is "out = in" today),
 * tests/Test_dwf_mixedcg_prec_halfcomms.cc:71-96 (DomainWallFermionFH as the inner operator of the mixed / reliable-update CG; the
 * program is compiled out at :33-34).
 * bf16 instead of the reference's fp16: the same 16 bits with fp32's exponent range, so the shrinking residual vectors of a restarted
 * solve keep their relative precision (2^-9 per component) instead of underflowing.  A compressed-halo operator takes the pack kernel
 * plus the interior + exterior (or serial) hop; the fused multi-rank launches stay with uncompressed halos.  Wilson-type operators. */
int gb_op_set_halo_compression(gb_fermop *op, int on);
/* fp32 operators: 1 (default) = column-sweep kernel (shared-memory z-column reuse) where it applies, else the micro-block
 * FFMA2 + TMA kernel, else the generic kernel; 2 = skip the column-sweep kernel; 0 = always the generic kernel */
int gb_op_set_fast_kernel(gb_fermop *op, int enable);

/* Physical 4D <-> 5D field maps of the 5D operators.  4D fields are gb_fermion with Ls = 1 on the same grid.
 * For WilsonFermion / ImprovedStaggeredFermion every map is a copy (ref: FermionOperator.h:172-191).
 *   ImportPhysicalFermionSource:   5D = Dminus [ P+ in4d at s=0 ; P- in4d at s=Ls-1 ]   ref: CayleyFermion5DImplementation.h:115-130
 *   ImportUnphysicalFermion:       the same without Dminus                              ref: :100-113
 *   ExportPhysicalFermionSolution: 4D = P- sol5d[s=0] + P+ sol5d[s=Ls-1]                ref: :58-69
 *   ExportPhysicalFermionSource:   4D = P+ src5d[s=0] + P- src5d[s=Ls-1]                ref: :88-99 */
int gb_op_import_physical_fermion_source(gb_fermop *op, const gb_fermion *in4d, gb_fermion *out5d);
int gb_op_import_unphysical_fermion(gb_fermop *op, const gb_fermion *in4d, gb_fermion *out5d);
int gb_op_export_physical_fermion_solution(gb_fermop *op, const gb_fermion *sol5d, gb_fermion *out4d);
int gb_op_export_physical_fermion_source(gb_fermop *op, const gb_fermion *src5d, gb_fermion *out4d);

/* Single hop legs and force terms on full-grid fields (Wilson-type operators).
 *   DhopDir(in, out, dir, disp): the leg of the hopping term that reads x + disp * dir, dir = 0..3 (x,y,z,t), disp = +-1, with the
 *     hopping term's own -1/2, boundary phases and halo exchange: summed over the eight legs it IS Dhop(in, out, DaggerNo).
 *     ref: WilsonFermion5DImplementation.h:183-200 (there dir5 = dir + 1), WilsonFermionImplementation.h:344-360
 *   DhopDeriv(mat, A, B, dag): mat_mu(x) = sum_s trace_spin [ Btilde_mu(x,s) A(x,s)^dagger ], Btilde_mu = forward leg mu of
 *     Dhop^(dag) applied to B; mat is a LatticeGaugeField [V4][4][3][3] (every entry overwritten).
 *     ref: WilsonFermion5DImplementation.h:212-275, WilsonImpl.h:173-238 (InsertForce4D/5D)
 *   MDeriv(mat, U, V, dag): Cayley operators apply Meooe5D to V (dag: to U) first; others = DhopDeriv.
 *     ref: CayleyFermion5DImplementation.h:347-360 ; driver tests/forces/Test_dwf_force.cc:71-75 */
int gb_op_dhop_dir(gb_fermop *op, const gb_fermion *in, gb_fermion *out, int dir, int disp);
int gb_op_dhop_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *A, const gb_fermion *B, int dag);
int gb_op_mderiv(gb_fermop *op, gb_gauge *mat, const gb_fermion *U, const gb_fermion *V, int dag);
/* Even-odd force terms.  MeoDeriv (U Even, V Odd) / MoeDeriv (U Odd, V Even), selected by U's checkerboard: writes the sites of
 * U's parity of the full-lattice mat and leaves the others alone (the reference assembles ForceE / ForceO with setCheckerboard).
 *   ref: CayleyFermion5DImplementation.h:361-390, WilsonFermion5DImplementation.h:277-305
 * SchurDifferentiableOperator::MpcDeriv (dagger = 0) / MpcDagDeriv (1): U, V on the Odd checkerboard, Force on the full grid
 *   ref: Grid/qcd/action/pseudofermion/EvenOddSchurDifferentiable.h:52-137 (what TwoFlavourEvenOddPseudoFermionAction::deriv calls) */
int gb_op_meooe_deriv(gb_fermop *op, gb_gauge *mat, const gb_fermion *U, const gb_fermion *V, int dag);
int gb_op_mpc_deriv(gb_fermop *op, gb_gauge *Force, const gb_fermion *U, const gb_fermion *V, int dagger);

/* ---------------------------------------------------------------- solvers
 * ConjugateGradient on SchurDiagMooeeOperator(op).HermOp, fused device path.
 * ref: Grid/algorithms/iterative/ConjugateGradient.h:68-257.  sol is the initial guess on entry.
 * iters_out = IterationsToComplete, true_resid_out = TrueResidual.  Returns GB_ERR_NOT_CONVERGED if
 * MaxIterations is reached (the wrapper asserts when ErrorOnNoConverge). */
int gb_cg_schur(gb_fermop *op, const gb_fermion *src, gb_fermion *sol, double tol, int maxit, int *iters_out, double *true_resid_out);
/* Generic LinearOperatorBase path: HermOp supplied by the caller (any user-written Schur operator). */
typedef int (*gb_hermop_fn)(void *user, const gb_fermion *in, gb_fermion *out);
int gb_cg(gb_context *ctx, gb_hermop_fn hermop, void *user, const gb_fermion *src, gb_fermion *sol, double tol, int maxit,
          int *iters_out, double *true_resid_out);
/* MixedPrecisionConjugateGradient: ref: Grid/algorithms/iterative/ConjugateGradientMixedPrec.h:71-167
 * iters_out[3] = {TotalInnerIterations, TotalOuterIterations, TotalFinalStepIterations} */
int gb_mixed_cg_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d, gb_fermion *sol_d, double tol, int max_inner,
                      int max_outer, int iters_out[3], double *true_resid_out);
/* the same with the class's public tuning members InnerTolerance (<= 0: = Tolerance) and OuterLoopNormMult (default 100)  ref: :41-45,:64-65 */
int gb_mixed_cg_schur_ex(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d, gb_fermion *sol_d, double tol, double inner_tol,
                         double outer_loop_norm_mult, int max_inner, int max_outer, int iters_out[3], double *true_resid_out);


/* MixedPrecisionConjugateGradientBatched(tol, maxinnerit, maxouterit, maxpatchit, sp_grid, Linop_f, Linop_d)(srcs, sols) on the Schur
 * operators of op_f (fp32) and op_d (fp64): the defect-correction loop over a batch of right-hand sides with ONE restart schedule -- all
 * residuals recomputed in double precision each outer iteration, the common inner tolerance loosened by the largest residual / target of
 * the batch, one fp32 CG per right-hand side, then a double-precision patch-up CG per right-hand side.  sols are the initial guesses.
 * iters_out[1 + 2 nbatch] = {restarts, inner iterations per right-hand side..., patch-up iterations per right-hand side...};
 * true_resid_out[nbatch].   ref: Grid/algorithms/iterative/ConjugateGradientMixedPrecBatched.h:36-213 */
int gb_mixed_cg_batched_schur(gb_fermop *op_f, gb_fermop *op_d, int nbatch, const gb_fermion *const *srcs_d, gb_fermion *const *sols_d,
                              double tol, int max_inner, int max_outer, int max_patchup, int update_residual, int *iters_out,
                              double *true_resid_out);

/* ConjugateGradientReliableUpdate(tol, maxit, Delta, sp_grid, Linop_f, Linop_d)(src, psi) on the Schur operators of op_f (fp32)
 * and op_d (fp64); psi is the initial guess.  iters_out[3] = {IterationsToComplete, ReliableUpdatesPerformed, IterationsToCleanup}.
 * ref: Grid/algorithms/iterative/ConjugateGradientReliableUpdate.h:36-270 ; driver tests/solver/Test_dwf_relupcg_prec.cc:88-104 */
int gb_relup_cg_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d, gb_fermion *sol_d, double tol, int maxit, double delta,
                      int iters_out[3], double *true_resid_out);

/* ConjugateGradientMultiShift on SchurDiagMooeeOperator(op).HermOp (staggered: SchurStaggeredOperator), SURVEY 8 row f3:
 * (HermOp + poles[s]) results[s] = src for s < nshift from one Krylov space; zero guess; poles[0] must be the lightest;
 * shift s stops when its residual estimate drops below tolerances[s] |src|.
 * ref: Grid/algorithms/iterative/ConjugateGradientMultiShift.h:84-343 ; shifts = MultiShiftFunction{poles, tolerances},
 * Grid/algorithms/approx/MultiShiftFunction.h:34-44.  iters_out[nshift + 1] = {IterationsToCompleteShift..., IterationsToComplete},
 * true_resid_out[nshift] = TrueResidualShift.  Returns GB_ERR_NOT_CONVERGED when MaxIterations is reached (the reference
 * only logs that case, :336-338; the results hold the iterate reached). */
int gb_cg_multishift_schur(gb_fermop *op, const gb_fermion *src, int nshift, const double *poles, const double *tolerances, int maxit,
                           gb_fermion *const *results, int *iters_out, double *true_resid_out);
/* ConjugateGradientMultiShiftMixedPrec(maxit, shifts, sp_grid, Linop_f, ReliableUpdateFreq)(Linop_d, src, results): the same
 * recurrences with fp64 vectors and ONE fp32 operator application per iteration; every relup_freq iterations the residual is
 * replaced by the true fp64 residual of the primary shift; shifts that miss their tolerance are cleaned up with
 * MixedPrecisionConjugateGradient on HermOp + pole.  ref: Grid/algorithms/iterative/ConjugateGradientMultiShiftMixedPrec.h:36-410 ;
 * driver tests/solver/Test_dwf_multishift_mixedprec.cc:113-128.  Outputs as gb_cg_multishift_schur. */
int gb_cg_multishift_mixed_schur(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src_d, int nshift, const double *poles,
                                 const double *tolerances, int maxit, int relup_freq, gb_fermion *const *results, int *iters_out,
                                 double *true_resid_out);

/* SchurRedBlackDiagMooeeSolve (Wilson-type operators) / SchurRedBlackStaggeredSolve (staggered): the full-lattice solve
 * M sol = src through the even-odd Schur decomposition.  ref: Grid/algorithms/iterative/SchurRedBlack.h:238-290,294-349,385-430
 *   RedBlackSource:   src_e = src|Even ; src_o = MpcDag (src|Odd - Meooe MooeeInv src_e)   (staggered: Mooee in place of MpcDag)
 *   RedBlackSolution: sol = [ MooeeInv (src_e - Meooe sol_o) | sol_o ]
 *   gb_schur_solve:   RedBlackSource, ConjugateGradient on the Odd checkerboard (ZeroGuesser, or the Odd part of sol when
 *                     use_sol_as_guess != 0, ref: :253-257), RedBlackSolution.  resid_out = {CG TrueResidual,
 *                     |M sol - src| / |src| (the "true unprec resid" the reference logs, ref: :277-285)}
 *   gb_schur_solve_mixed: the same with MixedPrecisionConjugateGradient as the red-black solver (fp64 fields and op_d outside,
 *                     fp32 op_f inside); iters_out as gb_mixed_cg_schur */
int gb_schur_redblack_source(gb_fermop *op, const gb_fermion *src, gb_fermion *src_e, gb_fermion *src_o);
int gb_schur_redblack_solution(gb_fermop *op, const gb_fermion *sol_o, const gb_fermion *src_e, gb_fermion *sol);
int gb_schur_solve(gb_fermop *op, const gb_fermion *src, gb_fermion *sol, double tol, int maxit, int use_sol_as_guess, int *iters_out,
                   double resid_out[2]);
int gb_schur_solve_mixed(gb_fermop *op_f, gb_fermop *op_d, const gb_fermion *src, gb_fermion *sol, double tol, int max_inner, int max_outer,
                         int iters_out[3], double resid_out[2]);

#ifdef __cplusplus
}
#endif
#endif /* GRIDB200_H */
