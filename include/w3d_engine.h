/*
 * w3d_engine.h -- C ABI of the B200 (sm_100a) element-block engine.
 *
 * This is the drop-in boundary for WARP3D's element-block hot path.  The reference has no FFI at
 * this level (no iso_c_binding anywhere in src/); the seam is the body of the two OpenMP block
 * loops and the block part of the n+1 -> n rollover.  Each entry point below cites the reference
 * code it replaces (paths relative to the reference tree's src/).  The Fortran side binds these
 * with ISO_C_BINDING (see INTEGRATION.md for the interface block and the patched call sites).
 *
 * Conventions
 *   - plain C: int32_t, int64_t, float, double and pointers to them; no C++/torch types.
 *   - arrays are in the reference's own (Fortran, column-major) layouts; index tables hold the
 *     reference's 1-based indices unchanged.
 *   - the caller owns every host array; the engine copies what it needs during the call.
 *   - return value: 0 = ok; >0 = a condition on which the reference would call die_abort
 *     (W3D_ERR_*).  A requested adaptive step cut is data (material_cut_step), not an error.
 *   - one handle per GPU / MPI rank; calls on one handle are not re-entrant.
 *   - there is no CPU fallback: every entry point fails with W3D_ERR_CUDA when no sm_100
 *     device / driver is usable.
 */
#ifndef W3D_ENGINE_H
#define W3D_ENGINE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W3D_ABI_VERSION 3

/* error codes (return values) */
#define W3D_OK             0
#define W3D_ERR_CUDA       1  /* CUDA runtime/driver failure or no device                       */
#define W3D_ERR_CONFIG     2  /* inconsistent tables in w3d_engine_config                       */
#define W3D_ERR_UNSUPPORTED 3 /* element type / material model / option outside the engine scope */
#define W3D_ERR_ARG        4  /* bad argument (null pointer, block index out of range, ...)      */
#define W3D_ERR_MM01_LAMBDA 5 /* mm01_general: negative plastic multiplier (mm01.f:878-881)      */
#define W3D_ERR_DETF       6  /* det F <= 0 and the step is not adaptive (rstgp1.f:136-145)      */
#define W3D_ERR_COMM       9  /* NCCL missing / failed, or a collective call on a handle without
                                 a communicator                                                 */

/* which = selector for w3d_push_state / w3d_pull_state.  Layout per block as in
 * mod_eleblocks.f:52-131 (blk is 1-based like the reference's block numbers):
 *   HIST_N/HIST_N1  history_blocks / history1_blocks (hist_size, ngp, span)
 *   EPS_N/EPS_N1    eps_n_blocks / eps_n1_blocks     (6, ngp, span)
 *   URCS_N/URCS_N1  urcs_n_blocks / urcs_n1_blocks   (9, ngp, span)
 *   ROT_N/ROT_N1    rot_n_blocks / rot_n1_blocks     (9, ngp, span)
 *   CEP             cep_blocks(blk)%vector           (21, span, ngp)   [rstgp1.f:5366]
 *   ESTIFF          estiff_blocks(blk)%ptr           (utsz, span)      [tanstf.f:422]
 *   ELEM_VOL        element_vol_blocks(blk)%ptr      (span)
 */
enum {
  W3D_HIST_N = 0, W3D_HIST_N1 = 1, W3D_EPS_N = 2, W3D_EPS_N1 = 3, W3D_URCS_N = 4,
  W3D_URCS_N1 = 5, W3D_ROT_N = 6, W3D_ROT_N1 = 7, W3D_CEP = 8, W3D_ESTIFF = 9,
  W3D_ELEM_VOL = 10
};

/* Model description handed over once, after the reference has read its input and run the
 * *_init routines (reopen.f:1015-1408, setup.f:268-276, inelbk.f:282-345).                      */
typedef struct w3d_engine_config {
  int32_t abi_version;       /* W3D_ABI_VERSION                                                  */
  int32_t mxvl, mxelpr, mxmtpr;          /* param_def:20-60                                      */
  int32_t nelblk, noelem, nonode, nodof; /* mod_main.f global_data                               */
  int32_t myid, numprocs;    /* MPI rank / domain count; the engine owns blocks whose
                                elblks(2,blk) == myid (drive_eps_sig_internal_forces.f:166)      */
  const int32_t *elblks;     /* (0:3, nelblk): span, felem, domain, -   (inelbk.f:47-52)         */
  const void    *props;      /* (mxelpr, noelem) 4-byte words, props/iprops/lprops aliases
                                (mod_main.f:41-46); rows used: 1,2,4,5,6,7,8,14,15,17,18,19,23,
                                25,38 (drive_eps_sig_internal_forces.f:448-461, rknstr.f:1036-1059)*/
  const float   *matprp;     /* (mxmtpr, nmat) real*4 material table, may be NULL                */
  int32_t        nmat;
  const double  *c;          /* (nodof) nodal coordinates (mod_main.f:48)                        */
  const int32_t *edest;      /* per-block (totdof, span) tables concatenated in block order
                                (edest_blocks, reopen.f:1315-1408)                               */
  const int32_t *cdest;      /* same for cdest_blocks (reopen.f:1200-1293)                       */
  const int32_t *history_blk_list; /* (nelblk) history words per Gauss point                     */
  double  eps_bbar;          /* global_data eps_bbar                                             */
  int32_t qbar_flag;         /* global_data qbar_flag ("consistent q")                           */
  double  beta_fact;         /* global_data beta_fact                                            */
  int32_t adaptive_flag;     /* global_data adaptive_flag                                        */
  int32_t device;            /* CUDA device ordinal, -1 = current device                         */
  /* ---- crystal plasticity (mm10) input; NULL / 0 when no block uses material model 10.
   * crystal_types holds the c_array(cnum) fields the path reads (mod_crystals.f:148-236) as
   * records of W3D_CRYSTAL_TYPE_DOUBLES doubles: elast_stiff(6,6) row-major, bi(48,3), ni(48,3)
   * row-major [system][xyz], then harden_n, tau_y, tau_v, voche_m, theta_o, k_o, b, iD_v,
   * eps_dot_o_y, atol, atol1, rtol, rtol1, then (as doubles) nslip, num_hard, h_type, tang_calc,
   * miter, solver, alter_mode.  elem_cnum / elem_angles give, per element and crystal, the
   * crystal number (1-based) and the Kocks angles exactly as setup_mm10_rknstr resolves them
   * from imatprp(104:107), crystal_input and angle_input (rknstr.f:1690-1720).                   */
  const double  *crystal_types;
  int32_t        n_crystal_types;
  int32_t        cp_ncrystals;     /* crystals per Gauss point, imatprp(101,matnum)              */
  const int32_t *elem_cnum;        /* (cp_ncrystals, noelem)                                     */
  const double  *elem_angles;      /* (3, cp_ncrystals, noelem)                                  */
  int32_t        cp_angle_type;    /* 1 = degrees, 2 = radians (imatprp(103)); Kocks convention  */
  int32_t        asymmetric_assembly; /* main_data asymmetric_assembly (mm10_a.f:1165)           */
} w3d_engine_config;
#define W3D_CRYSTAL_TYPE_DOUBLES 344

/* scalars the stress/ifv pass returns (drive_eps_sig_internal_forces.f:210-226, 316-319) */
typedef struct w3d_scalars {
  double  internal_energy;   /* sum of block_energies                                            */
  double  plastic_work;      /* sum of block_plastic_work                                        */
  double  sum_ifv;           /* sum |f_e| over all element force terms (addifv.f:104-110)        */
  int64_t num_term_ifv;      /* number of terms in sum_ifv                                       */
  int32_t material_cut_step; /* 1 = some Gauss point asked for an adaptive step cut              */
  int32_t reserved;
} w3d_scalars;

typedef struct w3d_engine w3d_engine;   /* opaque */

/* Build the device-side model: block tables, index maps, property vectors, zeroed state with
 * R = I (history_cep_init / rotation_init, reopen.f:1015-1176).
 * Built: hex8 (l3disop, 2x2x2), geonl and small strain, B-bar on/off; material models 1 (mm01
 * bilinear), 3 (mm03 `mises`: no voids, no nucleation, no segmental curves), 5 (mm05, both
 * options), 10 (mm10, Voce hardening, <= 48 slip systems).  Anything else in the tables returns
 * W3D_ERR_UNSUPPORTED here, for the whole model -- the caller keeps its own path for that model. */
int32_t w3d_engine_create(const w3d_engine_config *cfg, w3d_engine **handle);
int32_t w3d_engine_destroy(w3d_engine *handle);
/* last error text for this handle (or for a failed create when handle == NULL) */
const char *w3d_engine_last_error(const w3d_engine *handle);

/* global_data dt (time increment of the load step); read by rate-dependent models (mm10).       */
int32_t w3d_set_time_step(w3d_engine *h, double dt);

/* Replaces drive_eps_sig_internal_forces.f:115-349 (allocate_ifv, the OMP block loop over
 * do_nleps_block -> dupstr / dupstr_blocked / rknstr / rstgp1|rstgp2 / mm01+cnst1 / rplstr /
 * rknifv / gpifv1, the scalar reductions and the addifv scatter).  u, du, ifv are (nodof) host
 * vectors; ifv is overwritten with this engine's (this rank's) contribution, already scaled by
 * beta_fact, and zeroed when material_cut_step is set (:253-263).                               */
int32_t w3d_eps_sig_ifv(w3d_engine *h, int32_t step, int32_t iter, int32_t extrapolated_du,
                        const double *u, const double *du, double *ifv, w3d_scalars *out);

/* Replaces tanstf.f:56-95 (estiff_allocate + the OMP block loop over do_nlek_block -> dptstf /
 * dptstf_blocks / rktstf / gptns1 / bdbtgen / ctran1 / kgstiff / rktstf_do_transpose).  u, du
 * may be NULL to reuse the vectors of the last w3d_eps_sig_ifv call.  estiff, when not NULL,
 * receives every owned block's (utsz, span) matrix, blocks concatenated in block order
 * (offsets: w3d_block_offset); NULL keeps the result on the device (w3d_pull_state).           */
int32_t w3d_tanstf(w3d_engine *h, int32_t step, int32_t iter, const double *u, const double *du,
                   double *estiff);

/* Replaces the block part of update (update.f:75-93): history, eps, urcs  n+1 -> n.             */
int32_t w3d_update(w3d_engine *h);
/* adaptive_save (mnralg.f:1258-1310): rot n+1 -> n.                                             */
int32_t w3d_adaptive_save(w3d_engine *h);
/* adaptive_reset (mnralg.f:1335-1401): history, rot, eps, urcs  n -> n+1.                       */
int32_t w3d_adaptive_reset(w3d_engine *h);

/* Move one block of blocked state between the reference's host arrays and the device
 * (restart store.f/reopen.f, output ou*.f, J-integral di*.f read these).  blk is 1-based.      */
int32_t w3d_push_state(w3d_engine *h, int32_t which, int32_t blk, const double *src);
int32_t w3d_pull_state(w3d_engine *h, int32_t which, int32_t blk, double *dst);
/* offset in doubles of block blk (1-based; blk = nelblk+1 gives the total) inside an array that
 * concatenates all blocks of kind `which` in block order.                                       */
int64_t w3d_block_offset(const w3d_engine *h, int32_t which, int32_t blk);

/* ---- device-resident variants (no host<->device copies inside) --------------------------------
 * The engine owns device copies of u, du and ifv.  A caller that already lives on the GPU (the
 * NCCL layer, a device-side residual/assembly) fills u/du through these pointers, runs the
 * resident passes and reads ifv / scalars back when it needs them.                              */
int32_t w3d_device_vectors(w3d_engine *h, double **d_u, double **d_du, double **d_ifv);
/* host -> device copy of u and du (nodof each) / device -> host copy of ifv, on the engine's
 * stream, returning after completion.  Together with the resident passes these let a caller put
 * its own exchange step (NCCL sum of ifv) between the kernel and the copy back.                 */
int32_t w3d_upload_displacements(w3d_engine *h, const double *u, const double *du);
int32_t w3d_download_ifv(w3d_engine *h, double *ifv);
int32_t w3d_download_estiff(w3d_engine *h, double *estiff);
int32_t w3d_eps_sig_ifv_resident(w3d_engine *h, int32_t step, int32_t iter,
                                 int32_t extrapolated_du);
int32_t w3d_tanstf_resident(w3d_engine *h, int32_t step, int32_t iter);
/* scalars of the last resident stress pass (synchronises the engine's stream) */
int32_t w3d_last_scalars(w3d_engine *h, w3d_scalars *out);
/* the CUDA stream (cudaStream_t as void*) every engine kernel is launched on */
void   *w3d_engine_stream(w3d_engine *h);
int32_t w3d_engine_synchronize(w3d_engine *h);
/* number of kernels the engine has launched since create (bench.py: gpu_launches) */
int64_t w3d_engine_launch_count(const w3d_engine *h);

/* ---- the exchange step between domains (one engine = one GPU = one MPI rank) -------------------
 * Replaces what the reference's MPI build does at the end of drive_eps_sig_internal_forces:
 * wmpi_reduce_vec(ifv, nodof) (drive_eps_sig_internal_forces.f:338-340), the reductions of
 * internal_energy / plastic_work / sum_ifv / num_term_ifv and the logical OR of
 * material_cut_step (:248-250, mpi_code_real.f wmpi_redlog).  Transport is NCCL over NVLink,
 * bound at run time (dlopen of libnccl.so.2; W3D_NCCL_LIB overrides the name).
 *
 * Bootstrap as an MPI code does it: rank 0 calls w3d_comm_unique_id, the 128 bytes go to every
 * rank by the caller's own means (MPI_Bcast in the reference), then every rank calls
 * w3d_comm_init on its handle (created with cfg.myid = rank, cfg.numprocs = nranks).  From then
 * on w3d_eps_sig_ifv / w3d_eps_sig_ifv_resident are COLLECTIVE calls: they end with the exchange
 * of the internal forces and one 5-double all-reduce of the scalars and the cut flag, both on the
 * engine's stream, and w3d_scalars holds the sums over all ranks on every rank.
 *
 * w3d_comm_set_options
 *   exchange  1 (default): neighbour exchange.  Every engine knows from the replicated block
 *             tables which dofs its own blocks share with each other rank (w3d_comm_plan).  The
 *             partial sums on those dofs are packed, sent / received pairwise (grouped
 *             ncclSend/ncclRecv) and added: every rank ends with the COMPLETE internal force on
 *             every dof its own blocks touch, other entries keep this rank's partial (zero).
 *             The full vector is the concatenation of what the ranks hold on the dofs they own
 *             (w3d_comm_owned: the lowest rank touching a dof owns it).
 *             0: ncclAllReduce of the whole vector -- the reference's result on every rank.
 *   bcast_inputs  1: u and du given to rank 0 are broadcast over NVLink (ncclBroadcast) and the
 *             pointers of the other ranks are ignored (may be NULL): one host->device copy per
 *             iteration instead of one per GPU.  0 (default): every rank uploads its own copy,
 *             as the reference's ranks hold them (replicated).                                    */
int32_t w3d_comm_unique_id(uint8_t id[128]);
int32_t w3d_comm_init(w3d_engine *h, const uint8_t id[128], int32_t rank, int32_t nranks);
int32_t w3d_comm_set_options(w3d_engine *h, int32_t exchange, int32_t bcast_inputs);
/* Host-only planning step (no GPU needed), the same routine the engine uses at create: for the
 * domain `rank` of cfg, counts[r] = number of dofs shared with rank r (numprocs entries) and, when
 * dofs != NULL, the 1-based dof lists concatenated in rank order, each ascending.  Both sides of
 * a pair compute the same list.                                                                  */
int32_t w3d_comm_plan(const w3d_engine_config *cfg, int32_t rank, int64_t *counts, int32_t *dofs);
/* owned(nodof) = 1 where this rank owns the dof (lowest rank whose blocks touch it), else 0      */
int32_t w3d_comm_owned(const w3d_engine_config *cfg, int32_t rank, uint8_t *owned);
/* device -> host copy of ifv(first .. first+n-1), first 1-based: a rank returns only what it owns */
int32_t w3d_download_ifv_range(w3d_engine *h, int32_t first, int32_t n, double *dst);

/* ---- assembled equilibrium equations (SURVEY 8(f) NEXT-1), symmetric case -------------------
 * Replaces assemble_code.f: dof_map (:561-600), count_profile_symmetric (:20-250),
 * build_col_sparse_symm (:260-500) and assem_by_row / assem_a_row (:659-935) as
 * drive_assemble_solve.f calls them, so that the 2 400 B per element of estiff_blocks never
 * cross to the host.  Format = the reference's upper-triangle "VSS" vectors
 * (assemble_code.f:603-620): k_ptrs(neqns) terms right of the diagonal per row, k_indexes(ncoeff)
 * their equation numbers (1-based, ascending in a row), k_coeffs(ncoeff), k_diag(neqns).
 *
 * w3d_assemble_setup: once per constraint set.  cstmap(nodof): the reference's cstmap, 0 = dof
 * without an absolute constraint; edest: the packed per-block (totdof, span) tables of the
 * configuration.  Returns the number of equations of the structure and the number of
 * off-diagonal terms in the rows this domain holds.
 *
 * Several domains (numprocs > 1): every engine holds the rows of the nodes its own elements touch
 * -- a contiguous range of equations, w3d_assemble_rows -- in the layout of the WHOLE structure
 * (a row lists every column the structure has there, whichever domain's elements contribute),
 * with the values summed over its own elements.  Ranges of neighbouring domains overlap on the
 * interface nodes; there the rows have identical structure and the caller adds the values
 * (what wmpi_combine_stf + the root assembly do in the reference, mpi_code_real.f:2551).
 * One domain: the range is all equations and nothing below changes meaning.                      */
int32_t w3d_assemble_setup(w3d_engine *h, const int32_t *cstmap, const int32_t *edest,
                           int32_t *neqns_out, int64_t *ncoeff_out);
/* the rows of this domain: equations first_eqn .. first_eqn + n_eqns - 1 (1-based)              */
int32_t w3d_assemble_rows(w3d_engine *h, int32_t *first_eqn, int32_t *n_eqns);
/* the integer structure: dof_eqn_map(nodof), eqn_node_map(neqns) for the whole structure;
 * k_ptrs(n_eqns), k_indexes(ncoeff) for this domain's rows; any pointer may be NULL              */
int32_t w3d_assemble_structure(w3d_engine *h, int32_t *dof_eqn_map, int32_t *eqn_node_map,
                               int32_t *k_ptrs, int32_t *k_indexes);
/* assem_by_row on the element stiffnesses of the last w3d_tanstf / w3d_tanstf_resident (call
 * w3d_tanstf with estiff == NULL to keep them on the device).  Sums are taken in the reference's
 * order (elements on the node ascending, element columns ascending): bit-reproducible.  k_diag
 * (n_eqns) / k_coeffs (ncoeff): host buffers for this domain's rows, or both NULL to leave the
 * result in device memory.                                                                       */
int32_t w3d_assemble(w3d_engine *h, double *k_diag, double *k_coeffs);
/* Replaces convert_vss_csr (drive_assemble_solve.f:1501-1590) as drive_assemble_solve.f:757 calls
 * it for hypre and for asymmetric assembly: the full, structurally symmetric CSR pattern of the
 * structure (per row: lower terms, diagonal, upper terms, columns ascending; k_ptrs(neqns+1)
 * 1-based offsets, nnz = neqns + 2 ncoeff) -- bit-exact integers -- and, in k_coeffs(nnz), the
 * values of the last w3d_assemble in that pattern (the symmetric stiffness in full storage).  Any
 * pointer may be NULL.  One domain.  Not built: assem_by_row_asymmetric's VALUES for an
 * unsymmetric element stiffness (needs the full 24 x 24 element matrices, which k_tanstf does not
 * form).                                                                                          */
int32_t w3d_assemble_csr(w3d_engine *h, int64_t *nnz, int32_t *k_ptrs, int32_t *k_indexes,
                         double *k_coeffs);

/* ---- structure residual and its norms (SURVEY 8(f) NEXT-2) ------------------------------------
 * Replaces upres_iter_0 (upres.f:13-77), upres (upres.f:121-226) and the residual loop of cvtest
 * (cvtest.f:62-78) as mnralg.f calls them after drive_eps_sig_internal_forces: the internal
 * forces and du are already on the device, so only res (the next right-hand side) and these
 * scalars come back.  All vectors are (nodof) in structure-dof order.  NULL = the vector is
 * identically zero (statics: no mdiag/velocity/accel; no contact; d_lagrange_forces NULL = the
 * reference's have_mpc_equations false).                                                          */
typedef struct w3d_residual_loads {
  const int32_t *cstmap;            /* != 0: dof has an absolute constraint                        */
  const double  *pbar;              /* applied nodal forces at n+1 as mnralg.f holds them          */
  const double  *load;              /* `load` of upres: user nodal loads at n+1 (mgload measure)   */
  const double  *mdiag, *velocity, *accel;
  const double  *contact_force;
  const double  *d_lagrange_forces;
} w3d_residual_loads;

typedef struct w3d_residual_norms {
  double  norm_res;        /* sqrt(sum res^2): norm_load of upres_iter_0, `sum` of cvtest.f:75      */
  double  mgload;          /* iter >= 1 only (upres.f:219); 0 at iter 0                            */
  double  sum_loads;       /* iter >= 1 only                                                       */
  int64_t num_term_loads;  /* iter >= 1 only                                                       */
  double  res_max;         /* res(res_max_dof), signed (cvtest.f:77)                               */
  int32_t res_max_dof;     /* 1-based dof of the largest |res|, the lowest such dof                */
} w3d_residual_norms;

/* once per load step (or whenever one of the vectors changes); copies them to the device          */
int32_t w3d_residual_set_loads(w3d_engine *h, const w3d_residual_loads *loads);
/* iter == 0: upres_iter_0, iter >= 1: upres; then the cvtest residual loop.  Uses the internal
 * forces of the last stress pass and the du it was given.  res: host (nodof) or NULL to leave it
 * on the device (w3d_residual_device).  res(i) follows the reference's expression and operation
 * order; the three sums are fixed-order tree sums (run-to-run reproducible, equal to the
 * reference's sequential sums to rounding); num_term_loads and res_max_dof are exact.
 * Several domains (w3d_comm_init): collective.  After the neighbour exchange a rank's internal
 * forces are complete only on the dofs its own blocks touch, so every rank evaluates the dofs it
 * OWNS (w3d_comm_owned: the ranks' owned sets tile the vector), res is zero elsewhere -- the sum
 * of the ranks' res vectors is the structure residual -- and the norms come back summed over the
 * ranks (one small all-reduce), the same on every rank.  After the all-reduce mode every rank
 * holds the whole force vector and evaluates all of it.                                           */
int32_t w3d_residual(w3d_engine *h, int32_t iter, double dt, double nbeta, double *res,
                     w3d_residual_norms *out);
int32_t w3d_residual_device(w3d_engine *h, double **d_res);

#ifdef __cplusplus
}
#endif
#endif /* W3D_ENGINE_H */
