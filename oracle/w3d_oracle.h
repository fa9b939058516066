/*
 * w3d_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A scalar, double-precision restatement of WARP3D's element-block hot path
 *   drive_eps_sig_internal_forces / do_nleps_block / rknstr / rstgp1 / rstgp2 / mm01 / cnst1 /
 *   rknifv / gpifv1 / addifv / tanstf / do_nlek_block / rktstf / gptns1 / kgstiff / update
 * for hex8 (l3disop, type 2) blocks.  Every function in w3d_oracle.cpp cites the reference
 * file:line it follows (paths relative to the reference tree's src/).
 *
 * PARITY STATUS.  The reference ships no function-level golden vectors for this path and cannot
 * be compiled in this environment (no Fortran compiler, needs ifort+MKL).
 *   - mm10 (crystal plasticity) together with the hex8 small-strain kinematics, B-bar, internal
 *     force, tangent and the Newton structure around them: PINNED end to end against numbers the
 *     Fortran produced -- the 79-point stretch/relaxation curve published with
 *     example_problems_cp/simple_voce is reproduced to 1.1e-5 at worst, 1e-7 at 80 h
 *     (tests/test_reference_known_answers.py); and the known answer of
 *     verification/test82/voche_model (bcc, 12 systems, Voce exponent 5, cube with a void under
 *     biaxial face pressures): sigma_xx of element 64 = 476.664670 against 476.645594, 4.0e-5
 *     (the reference stops that analysis at 1e-4 of the force norm and accepts 1e-4).
 *   - mm01 (bilinear Mises) with the finite-strain kinematics (polar decomposition, un-rotated
 *     rates), B-bar, internal force and tangent: PINNED end to end by the reference's known answer
 *     for verification/test24 (hollow sphere sector, 21 elastic-plastic load steps, the inner
 *     radius grows by a quarter, skew constraints): sigma_xx of element 1 = -205.046618 here against
 *     -205.047360 printed by the Fortran, 3.6e-6 (same test file; the reference's own tolerance for
 *     this number is 1e-4 and it stops its Newton loop at 1e-5).
 *   - mm05 (cyclic plasticity), both options, with the same finite-strain chain: PINNED end to end
 *     by the reference's known answers for verification/test60 a and b (notched bar, 405 elements,
 *     420 load steps of cyclic end displacement): total axial reaction 0.0210364867 against
 *     0.021036833 (nonlinear_hardening, 1.6e-5) and 0.0232109041 against 0.023211111
 *     (generalized_plasticity, 8.9e-6).
 *   - the assembly (w3d_oracle_assemble.cpp: dof_map, count_profile_symmetric,
 *     build_col_sparse_symm, assem_by_row) and the structure residual with its norms
 *     (w3d_oracle_residual.cpp: upres_iter_0, upres, the cvtest residual loop): literal sequential
 *     restatements, checked against independent scipy / numpy formulations of the same
 *     definitions (tests/test_assembly.py, tests/test_residual.py); the reference holds no
 *     vectors for them: "parity unpinned".
 *   Nothing is pinned at function granularity or at 1e-10: there the status stays "parity
 *   unpinned" (literal transcription + the physics invariants in tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library.  The CUDA engine package must never include or link it.
 *
 * Array conventions: all arrays are in the reference's (Fortran, column-major) layouts, indices
 * inside index tables are 1-based exactly as the reference stores them.
 */
#ifndef W3D_ORACLE_H
#define W3D_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W3DO_MXVL   128   /* param_def: mxvl   */
#define W3DO_MXELPR 43    /* param_def: mxelpr */
#define W3DO_MXMTPR 300   /* param_def: mxmtpr */

/* Global (model-level) description.  Mirrors the module variables the reference hot path reads:
 * global_data (mod_main.f:28-130), elem_block_data (mod_eleblocks.f:49-131). */
typedef struct w3do_model {
  int32_t nelblk, noelem, nonode, nodof;
  const int32_t *elblks;   /* (0:3, nelblk): span, felem (1-based), domain, unused  (inelbk.f:47-52) */
  const void    *props;    /* (mxelpr, noelem) 4-byte words: real/int/logical aliases (mod_main.f:41-46) */
  const float   *matprp;   /* (mxmtpr, nmat) real*4 material table (may be NULL for mm01 w/o fgm)      */
  int32_t        nmat;
  const double  *c;        /* (nodof) nodal coordinates, 3 per node                                    */
  const int32_t *edest;    /* concatenated per-block (totdof, span) 1-based dof indices (reopen.f:1315) */
  const int32_t *cdest;    /* concatenated per-block (totdof, span) 1-based coord indices (reopen.f:1200)*/
  const int32_t *hist_size;/* history_blk_list(nelblk)                                                  */
  double  eps_bbar;        /* global_data eps_bbar                                                      */
  int32_t qbar_flag;       /* global_data qbar_flag ("consistent q")                                    */
  double  dt, total_model_time, beta_fact;
  int32_t extrapolated_du; /* main_data extrapolated_du                                                 */
  int32_t adaptive_flag;   /* global_data adaptive_flag                                                 */
  int32_t fpconstant_data; /* parity-quirk switch #2: DATA literals evaluated in double      */
  /* crystal-plasticity tables (mm10) -- NULL unless a block uses material type 10 */
  const void *cp;          /* struct w3do_cp*, see w3d_oracle_cp.h                                       */
} w3do_model;

/* crystal-plasticity data (mm10): per (element, crystal) property records produced by
 * w3do_mm10_setup_props == setup_mm10_rknstr (rknstr.f:1568-2027).  One record is
 * W3DO_MM10_PROPS_DOUBLES doubles (w3do_mm10::Props<12>); all elements carry `ncrystals`.        */
typedef struct w3do_cp {
  int32_t ncrystals, nslip, num_hard, asymmetric_assembly;
  const double *props;     /* (noelem, ncrystals, W3DO_MM10_PROPS_DOUBLES) */
} w3do_cp;

/* Blocked state, concatenated over blocks in block order; per-block pieces are in the reference
 * layouts: history (H,ngp,span), eps (6,ngp,span), urcs (9,ngp,span), rot (9,ngp,span),
 * cep (21,span,ngp) [note span before gp], estiff (utsz,span), element_vol (span).
 * (mod_eleblocks.f:52-131, rstgp1.f:5366, tanstf.f:422)                                              */
typedef struct w3do_state {
  double *history, *history1;
  double *eps_n, *eps_n1;
  double *urcs_n, *urcs_n1;
  double *rot_n, *rot_n1;
  double *cep;
  double *estiff;
  double *element_vol;
} w3do_state;

/* scalars returned by the stress/ifv driver (drive_eps_sig_internal_forces.f:210-226,316-319) */
typedef struct w3do_scalars {
  double  internal_energy, plastic_work, sum_ifv;
  int64_t num_term_ifv;
  int32_t material_cut_step;
} w3do_scalars;

/* drive_eps_sig_internal_forces (drive_eps_sig_internal_forces.f:20).  nthreads<=1 => serial and
 * deterministic scatter in block order; >1 => OpenMP over blocks with atomic scatter like the
 * reference.  Returns 0, or >0 on a fatal condition the reference would die_abort on.           */
int32_t w3do_drive_eps_sig_internal_forces(const w3do_model *m, w3do_state *s, int32_t step,
                                           int32_t iter, const double *u, const double *du,
                                           double *ifv, w3do_scalars *out, int32_t nthreads);

/* tanstf (tanstf.f:17): fills s->estiff for all blocks. */
int32_t w3do_tanstf(const w3do_model *m, w3do_state *s, int32_t step, int32_t iter,
                    const double *u, const double *du, int32_t nthreads);

/* block part of update (update.f:75-93): history, eps, urcs n1 -> n */
void w3do_update(const w3do_model *m, w3do_state *s);
/* adaptive_save (mnralg.f:1305): rot n1 -> n ; adaptive_reset: n -> n1 for all four */
void w3do_adaptive_save(const w3do_model *m, w3do_state *s);
void w3do_adaptive_reset(const w3do_model *m, w3do_state *s);

/* offsets (in doubles) of block blk (0-based) inside the concatenated arrays; which:
 * 0 history, 1 eps, 2 urcs, 3 rot, 4 cep, 5 estiff, 6 element_vol, 7 edest/cdest (in int32s)    */
int64_t w3do_block_offset(const w3do_model *m, int32_t blk, int32_t which);

/* mm10: number of doubles in one property record; fill one record from a crystal-type table
 * entry (w3do_mm10::CrystalType laid out as doubles, see oracle_py.py) and Kocks angles.          */
int32_t w3do_mm10_props_doubles(void);
int32_t w3do_mm10_props_doubles_for(int32_t nslip);   /* 12-system record, or 48-system */
int32_t w3do_mm10_crystal_type_doubles(void);
void w3do_mm10_setup_props(const double *crystal_type, const double *angles, int32_t angle_type,
                           double *props_out);
int32_t w3do_mm10_hist_size(int32_t nslip, int32_t num_hard, int32_t ncrystals);

/* small pieces exposed for unit tests */
void w3do_quad2(int32_t gpn, double *xi, double *eta, double *zeta, double *w);
void w3do_deriv2(double xi, double eta, double zeta, double *nxi, double *neta, double *nzeta);
void w3do_cp_icp(int32_t n, int32_t *cp, int32_t *dcp, int32_t *icp /* (utsz,2) */);

#ifdef __cplusplus
}
#endif
#endif
