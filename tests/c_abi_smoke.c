/* c_abi_smoke.c -- plain-C driver of the engine's C ABI (include/w3d_engine.h), compiled with gcc by
 * tests/test_c_abi.py.  It makes exactly the calls, in the order and with the array layouts, that
 * the ISO_C_BINDING interface in fortran/mod_w3d_engine.f makes from WARP3D:
 *   w3d_engine_create -> w3d_eps_sig_ifv (iteration 0, 1) -> w3d_tanstf (element blocks out) ->
 *   w3d_tanstf (kept on the device) + w3d_assemble_setup / _structure / w3d_assemble / _csr ->
 *   w3d_residual_set_loads / w3d_residual -> w3d_update / w3d_adaptive_save -> w3d_engine_destroy
 * on a structured box of hex8 elements whose tables (elblks, props words, matprp, coordinates,
 * edest/cdest in the all-u, all-v, all-w order) are filled here from scratch in C, i.e. without
 * the Python host layer.  Prints one line of check sums; the GPU test compares them with the same
 * mesh driven through ctypes.  Without a usable device create must fail loudly with W3D_ERR_CUDA
 * (exit code 77), never fall back.
 * usage: c_abi_smoke nx ny nz block_size */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "w3d_engine.h"

#define MXELPR 43
#define MXMTPR 300

static int32_t fbits(float x) { int32_t i; memcpy(&i, &x, 4); return i; }

int main(int argc, char **argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 4, ny = argc > 2 ? atoi(argv[2]) : 3, nz = argc > 3 ? atoi(argv[3]) : 2;
  const int bsz = argc > 4 ? atoi(argv[4]) : 7;
  const int nnx = nx + 1, nny = ny + 1, nnz = nz + 1;
  const int nonode = nnx * nny * nnz, noelem = nx * ny * nz, nodof = 3 * nonode;
  /* WARP3D hex8 node order in (xi, eta, zeta): derivs.f:283-343 */
  static const int sg[8][3] = {{0, 0, 1}, {0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 0, 0}, {1, 1, 0}, {1, 1, 1}};
  double *c = calloc(nodof, sizeof(double)), *u = calloc(nodof, sizeof(double)), *du = calloc(nodof, sizeof(double));
  double *ifv = calloc(nodof, sizeof(double));
  for (int k = 0; k < nnz; ++k)
    for (int j = 0; j < nny; ++j)
      for (int i = 0; i < nnx; ++i) {
        const int n = (k * nny + j) * nnx + i;
        c[3 * n] = i; c[3 * n + 1] = j; c[3 * n + 2] = k;
        /* a stretch in x with Poisson contraction, enough to yield */
        du[3 * n] = 3.0e-3 * i; du[3 * n + 1] = -0.9e-3 * j; du[3 * n + 2] = -0.9e-3 * k;
      }
  /* material table and element property words (elprp.f:715-935, l3disop) */
  float *matprp = calloc(MXMTPR, sizeof(float));
  const float e = 200.0e9f, nu = 0.3f, tan_e = 2.0e9f, yld = 400.0e6f, beta = 1.0f;
  matprp[0] = e; matprp[1] = nu; matprp[2] = beta; matprp[3] = tan_e; matprp[4] = yld; matprp[8] = 1.0f;
  int32_t *props = calloc((size_t)noelem * MXELPR, sizeof(int32_t));
  int32_t *edest = calloc((size_t)noelem * 24, sizeof(int32_t));
  const float hprime = (e * tan_e) / (e - tan_e);
  for (int el = 0; el < noelem; ++el) {
    int32_t *p = props + (size_t)el * MXELPR;
    p[0] = 2; p[1] = 8; p[2] = 24 * 65536 + 6; p[3] = 3; p[4] = 1; p[5] = 8; p[10] = 6;
    p[6] = fbits(e); p[7] = fbits(nu); p[13] = fbits(beta); p[14] = fbits(hprime); p[16] = 0;
    p[17] = -1; p[18] = -1;               /* geonl, bbar: ifort .true. */
    p[22] = fbits(yld); p[24] = 1; p[37] = 1; p[41] = -1;
    const int ei = el % nx, ej = (el / nx) % ny, ek = el / (nx * ny);
    for (int a = 0; a < 8; ++a) {
      const int n = ((ek + sg[a][2]) * nny + (ej + sg[a][1])) * nnx + (ei + sg[a][0]);
      for (int d = 0; d < 3; ++d) edest[(size_t)el * 24 + d * 8 + a] = 3 * n + d + 1;   /* all u, all v, all w */
    }
  }
  /* automatic blocking (inelbk.f:282-345): runs of bsz consecutive elements */
  const int nelblk = (noelem + bsz - 1) / bsz;
  int32_t *elblks = calloc((size_t)nelblk * 4, sizeof(int32_t)), *hbl = calloc(nelblk, sizeof(int32_t));
  for (int b = 0; b < nelblk; ++b) {
    elblks[4 * b] = (b + 1) * bsz <= noelem ? bsz : noelem - b * bsz;
    elblks[4 * b + 1] = b * bsz + 1;
    hbl[b] = 11;                           /* mm01 history words */
  }
  w3d_engine_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = W3D_ABI_VERSION;
  cfg.mxvl = 128; cfg.mxelpr = MXELPR; cfg.mxmtpr = MXMTPR;
  cfg.nelblk = nelblk; cfg.noelem = noelem; cfg.nonode = nonode; cfg.nodof = nodof;
  cfg.myid = 0; cfg.numprocs = 1;
  cfg.elblks = elblks; cfg.props = props; cfg.matprp = matprp; cfg.nmat = 1; cfg.c = c;
  cfg.edest = edest; cfg.cdest = edest; cfg.history_blk_list = hbl;
  cfg.eps_bbar = 0.0; cfg.qbar_flag = 1; cfg.beta_fact = 1.0; cfg.adaptive_flag = 1; cfg.device = -1;
  w3d_engine *h = NULL;
  int32_t rc = w3d_engine_create(&cfg, &h);
  if (rc == W3D_ERR_CUDA) {
    printf("c_abi_smoke: no usable device: %s\n", w3d_engine_last_error(NULL));
    return 77;
  }
#define CHECK(call)                                                                         \
  do {                                                                                      \
    rc = (call);                                                                            \
    if (rc) { printf("c_abi_smoke: %s -> %d: %s\n", #call, rc, w3d_engine_last_error(h)); return 1; } \
  } while (0)
  if (rc) { printf("c_abi_smoke: create -> %d: %s\n", rc, w3d_engine_last_error(NULL)); return 1; }
  w3d_scalars sc;
  CHECK(w3d_eps_sig_ifv(h, 1, 0, 0, u, du, ifv, &sc));
  CHECK(w3d_eps_sig_ifv(h, 1, 1, 0, u, du, ifv, &sc));
  double sum_abs_ifv = 0.0;
  for (int i = 0; i < nodof; ++i) sum_abs_ifv += fabs(ifv[i]);
  /* element blocks back in the reference's (300, span) layout */
  const int64_t nk = w3d_block_offset(h, W3D_ESTIFF, nelblk + 1);
  double *ek = calloc((size_t)nk, sizeof(double));
  CHECK(w3d_tanstf(h, 1, 1, NULL, NULL, ek));
  double sum_abs_ek = 0.0;
  for (int64_t i = 0; i < nk; ++i) sum_abs_ek += fabs(ek[i]);
  /* assembled equations: x = 0 face clamped */
  int32_t *cst = calloc(nodof, sizeof(int32_t));
  for (int n = 0; n < nonode; ++n)
    if (c[3 * n] == 0.0) cst[3 * n] = cst[3 * n + 1] = cst[3 * n + 2] = 1;
  int32_t neqns = 0; int64_t ncoeff = 0, nnz_csr = 0;
  CHECK(w3d_tanstf(h, 1, 1, NULL, NULL, NULL));
  CHECK(w3d_assemble_setup(h, cst, edest, &neqns, &ncoeff));
  int32_t *kptrs = calloc(neqns, sizeof(int32_t)), *kidx = calloc(ncoeff ? ncoeff : 1, sizeof(int32_t));
  CHECK(w3d_assemble_structure(h, NULL, NULL, kptrs, kidx));
  double *kd = calloc(neqns, sizeof(double)), *kc = calloc(ncoeff ? ncoeff : 1, sizeof(double));
  CHECK(w3d_assemble(h, kd, kc));
  CHECK(w3d_assemble_csr(h, &nnz_csr, NULL, NULL, NULL));
  double sum_kd = 0.0, sum_abs_kc = 0.0;
  for (int i = 0; i < neqns; ++i) sum_kd += kd[i];
  for (int64_t i = 0; i < ncoeff; ++i) sum_abs_kc += fabs(kc[i]);
  /* residual against zero applied loads: the norm of the free internal forces */
  double *zero = calloc(nodof, sizeof(double)), *res = calloc(nodof, sizeof(double));
  w3d_residual_loads ld;
  memset(&ld, 0, sizeof ld);
  ld.cstmap = cst; ld.pbar = zero; ld.load = zero;
  w3d_residual_norms nrm;
  CHECK(w3d_residual_set_loads(h, &ld));
  CHECK(w3d_residual(h, 1, 1.0, 0.25, res, &nrm));
  CHECK(w3d_update(h));
  CHECK(w3d_adaptive_save(h));
  printf("c_abi_smoke: noelem %d nelblk %d num_term_ifv %lld cut %d sum_abs_ifv %.15e plastic_work %.15e "
         "sum_abs_estiff %.15e neqns %d ncoeff %lld nnz %lld sum_kdiag %.15e sum_abs_kcoeffs %.15e norm_res %.15e "
         "launches %lld\n",
         noelem, nelblk, (long long)sc.num_term_ifv, sc.material_cut_step, sum_abs_ifv, sc.plastic_work, sum_abs_ek,
         neqns, (long long)ncoeff, (long long)nnz_csr, sum_kd, sum_abs_kc, nrm.norm_res, (long long)w3d_engine_launch_count(h));
  CHECK(w3d_engine_destroy(h));
  return 0;
}
