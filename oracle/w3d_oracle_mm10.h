// w3d_oracle_mm10.h -- ORACLE (test infrastructure): statement-for-statement CPU restatement of
// the crystal-plasticity material update mm10 for one Gauss point.  Independent of the product:
// nothing of the engine package is included here, and the engine package includes nothing of oracle/.
// The engine's own core (csrc/w3d_mm10_core.h and the lattice-frame kernel of the engine package) re-orders
// the arithmetic for the GPU; this file keeps the reference's routine structure, one C++ function
// per Fortran routine, each slip system evaluated where the Fortran evaluates it (including the
// repeated pow() calls), DGESV/DPOSV as LU with partial pivoting in DGETF2 order.
//
// Follows reference src/: mm10_a.f (mm10 :29, mm10_a_do_crystal :204, mm10_a_store_crystal :290,
// mm10_init_cc_hist0 :480, mm10_tangent :664, mm10_setup :852, mm10_store_cryhist :998,
// mm10_solve_crystal :1102, mm10_update_euler_angles :1195, mm10_rt2rve :1431, mm10_rt2rvw :1492,
// mm10_setup_np1 :1689, mm10_init_voche :2071, mm10_setup_voche :2086, mm10_copy_cc_hist :2493,
// mm10_solve_strup :2657, mm10_solve :2901 (predict :3032, update :3224), mm10_update_rotation
// :3375, mm10_expmw3x3 :3447, mm10_output :3493), mm10_b.f (mm10_formR2 :63, mm10_formJ11 :177,
// mm10_formJ12 :281, mm10_formJ21 :377, mm10_formJ22 :439, mm10_formR1 :1073, mm10_form_wbarp
// :1239, mm10_form_wp :1312, mm10_symsw :1513, mm10_iw :1613, mm10_slipinc / mm10_rs :1735-1790,
// voche routines :1813-2074), mm10_d.f :25-330 (history layout), rknstr.f :1568-2027
// (setup_mm10_rknstr).
// Scope: hardening model 1 (Voce/"voche"), default Newton solver (props%solver), tang_calc 0,
// any slip-system table up to NS systems, any number of crystals per point.
// Parity status: pinned end to end by the reference's published simple_voce curve (48 systems,
// 25 crystals) and verification/test82/voche_model (tests/test_reference_known_answers.py); at
// function level 1e-10 is unpinned (no Fortran compiler here).
#pragma once
#include <math.h>

#define W3D_HD   /* host only */

// optional operation counters (build with -DW3D_MM10_STATS; bench.py uses them to state the
// algorithmic work of a launch): [0] state evaluations (one mm10_formR1 = one pass over the slip
// systems), [1] Jacobians (mm10_formJ11 / mm10_formJ), [2] dense solves, [3] crystal updates that
// ran the solver, [4] sub-increment solves
#if defined(W3D_MM10_STATS)
extern "C" long long w3d_mm10_stats[8];
#define W3D_STAT(k) __atomic_fetch_add(&w3d_mm10_stats[k], 1LL, __ATOMIC_RELAXED)
#else
#define W3D_STAT(k) ((void)0)
#endif

namespace w3do_mm10 {

// strided view of one Gauss point's history vector (stride = span on the CPU block layout,
// = number of Gauss-point slots in the engine's structure-of-arrays layout)
struct Hist {
  double *p;
  long long stride;
  W3D_HD double &operator[](int k) const { return p[(long long)k * stride]; }
};
struct CHist {
  const double *p;
  long long stride;
  W3D_HD double operator[](int k) const { return p[(long long)k * stride]; }
};

// history layout of mm10_set_history_locs (mm10_d.f:153-330), 0-based offsets
struct Layout {
  int nslip, num_hard, len_slip, len_tt, len_u, len_rate;
  int common_size, crystal_size;
  // common part: [D 36][gradFe 27][R 9][work 3][slip history len_slip]
  // per crystal: stress 6, euler 3, Rp 9, D 6, eps 6, slip_incs, tau_tilde, u, tt_rate, ep 6, ed 6
  W3D_HD static Layout make(int nslip, int num_hard) {
    Layout L;
    const bool use_max = (num_hard == 48) || (nslip == 48);   // max_uhard = max_slip_sys = 48
    L.nslip = nslip; L.num_hard = num_hard;
    L.len_slip = use_max ? 48 : nslip;
    L.len_tt = use_max ? 48 : num_hard;
    L.len_u = use_max ? 48 : 15;
    L.len_rate = use_max ? 48 : num_hard;
    L.common_size = 36 + 27 + 9 + 3 + L.len_slip;
    L.crystal_size = 6 + 3 + 9 + 6 + 6 + L.len_slip + L.len_tt + L.len_u + L.len_rate + 6 + 6;
    return L;
  }
  W3D_HD int size(int ncrystals) const { return common_size + ncrystals * crystal_size; }
  W3D_HD int o_D() const { return 0; }
  W3D_HD int o_gradfe() const { return 36; }
  W3D_HD int o_R() const { return 63; }
  W3D_HD int o_work() const { return 72; }
  W3D_HD int o_sliphist() const { return 75; }
  W3D_HD int c0(int c) const { return common_size + c * crystal_size; }
  W3D_HD int c_stress(int c) const { return c0(c); }
  W3D_HD int c_euler(int c) const { return c0(c) + 6; }
  W3D_HD int c_Rp(int c) const { return c0(c) + 9; }
  W3D_HD int c_D(int c) const { return c0(c) + 18; }
  W3D_HD int c_eps(int c) const { return c0(c) + 24; }
  W3D_HD int c_slip(int c) const { return c0(c) + 30; }
  W3D_HD int c_tt(int c) const { return c_slip(c) + len_slip; }
  W3D_HD int c_u(int c) const { return c_tt(c) + len_tt; }
  W3D_HD int c_rate(int c) const { return c_u(c) + len_u; }
  W3D_HD int c_ep(int c) const { return c_rate(c) + len_rate; }
  W3D_HD int c_ed(int c) const { return c_ep(c) + 6; }
};

// per (element, crystal) properties = local_work%c_props(i,c) after setup_mm10_rknstr, packed
// as doubles so the same record is an HBM array on the device.  Matrices are row-major [r][c].
template <int NS>
struct Props {
  double rate_n, tau_y, tau_v, voche_m, theta_0, k_0, burgers, iD_v, eps_dot_0_y;
  double atol, atol1, rtol, rtol1;
  double init_angles[3];
  double g[3][3];            // rotation_g
  double stiffness[6][6];    // init_elast_stiff rotated to the reference frame
  double ms[NS][6], qs[NS][3], ns[NS][3];
  double nslip_d, miter_d, alter_mode_d, angle_type_d;   // small integers kept as doubles
  W3D_HD int nslip() const { return (int)nslip_d; }
  W3D_HD int miter() const { return (int)miter_d; }
};

// ---- small dense helpers ---------------------------------------------------------------------
W3D_HD inline void rt2rve(const double (&rt)[3][3], double (&rv)[6][6]) {   // mm10_a.f:1431
  const double two = 2.0;
  rv[0][0] = rt[0][0] * rt[0][0]; rv[0][1] = rt[0][1] * rt[0][1]; rv[0][2] = rt[0][2] * rt[0][2];
  rv[0][3] = two * rt[0][0] * rt[0][1]; rv[0][4] = two * rt[0][2] * rt[0][1]; rv[0][5] = two * rt[0][0] * rt[0][2];
  rv[1][0] = rt[1][0] * rt[1][0]; rv[1][1] = rt[1][1] * rt[1][1]; rv[1][2] = rt[1][2] * rt[1][2];
  rv[1][3] = two * rt[1][0] * rt[1][1]; rv[1][4] = two * rt[1][2] * rt[1][1]; rv[1][5] = two * rt[1][0] * rt[1][2];
  rv[2][0] = rt[2][0] * rt[2][0]; rv[2][1] = rt[2][1] * rt[2][1]; rv[2][2] = rt[2][2] * rt[2][2];
  rv[2][3] = two * rt[2][0] * rt[2][1]; rv[2][4] = two * rt[2][2] * rt[2][1]; rv[2][5] = two * rt[2][0] * rt[2][2];
  rv[3][0] = rt[0][0] * rt[1][0]; rv[3][1] = rt[0][1] * rt[1][1]; rv[3][2] = rt[0][2] * rt[1][2];
  rv[3][3] = rt[0][0] * rt[1][1] + rt[1][0] * rt[0][1];
  rv[3][4] = rt[0][1] * rt[1][2] + rt[0][2] * rt[1][1];
  rv[3][5] = rt[0][0] * rt[1][2] + rt[0][2] * rt[1][0];
  rv[4][0] = rt[1][0] * rt[2][0]; rv[4][1] = rt[2][1] * rt[1][1]; rv[4][2] = rt[1][2] * rt[2][2];
  rv[4][3] = rt[1][0] * rt[2][1] + rt[1][1] * rt[2][0];
  rv[4][4] = rt[1][1] * rt[2][2] + rt[2][1] * rt[1][2];
  rv[4][5] = rt[1][0] * rt[2][2] + rt[1][2] * rt[2][0];
  rv[5][0] = rt[0][0] * rt[2][0]; rv[5][1] = rt[0][1] * rt[2][1]; rv[5][2] = rt[0][2] * rt[2][2];
  rv[5][3] = rt[0][0] * rt[2][1] + rt[0][1] * rt[2][0];
  rv[5][4] = rt[0][1] * rt[2][2] + rt[0][2] * rt[2][1];
  rv[5][5] = rt[0][0] * rt[2][2] + rt[2][0] * rt[0][2];
}
W3D_HD inline void rt2rvw(const double (&rt)[3][3], double (&rv)[3][3]) {   // mm10_a.f:1492
  rv[0][0] = rt[1][1] * rt[2][2] - rt[1][2] * rt[2][1];
  rv[0][1] = rt[1][0] * rt[2][2] - rt[1][2] * rt[2][0];
  rv[0][2] = rt[1][0] * rt[2][1] - rt[1][1] * rt[2][0];
  rv[1][0] = rt[0][1] * rt[2][2] - rt[0][2] * rt[2][1];
  rv[1][1] = rt[0][0] * rt[2][2] - rt[0][2] * rt[2][0];
  rv[1][2] = rt[0][0] * rt[2][1] - rt[0][1] * rt[2][0];
  rv[2][0] = rt[0][1] * rt[1][2] - rt[0][2] * rt[1][1];
  rv[2][1] = rt[0][0] * rt[1][2] - rt[0][2] * rt[1][0];
  rv[2][2] = rt[0][0] * rt[1][1] - rt[0][1] * rt[1][0];
}
W3D_HD inline void mat33(double (&a)[3][3], const double (&b)[3][3], const double (&c)[3][3]) {  // a = b c
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j];
}
W3D_HD inline void mat33_bt(double (&a)[3][3], const double (&b)[3][3], const double (&c)[3][3]) {  // a = b c^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = b[i][0] * c[j][0] + b[i][1] * c[j][1] + b[i][2] * c[j][2];
}
W3D_HD inline void mv6(double (&a)[6], const double (&b)[6][6], const double *c) {   // mult_type_2
  for (int i = 0; i < 6; ++i) a[i] = b[i][0] * c[0];
  for (int j = 1; j < 6; ++j)
    for (int i = 0; i < 6; ++i) a[i] = a[i] + b[i][j] * c[j];
}
W3D_HD inline void mv3(double *a, const double (&b)[3][3], const double *c) {       // mult_type_3
  for (int i = 0; i < 3; ++i) a[i] = b[i][0] * c[0] + b[i][1] * c[1] + b[i][2] * c[2];
}
W3D_HD inline void symsw(const double *s, const double *w, double (&sw)[6]) {      // mm10_b.f:1513
  const double half = 0.5;
  sw[0] = s[3] * w[2] - s[5] * w[1];
  sw[1] = s[3] * w[2] - s[4] * w[0];
  sw[2] = s[5] * w[1] + s[4] * w[0];
  sw[3] = half * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]);
  sw[4] = half * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]);
  sw[5] = half * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4]);
}
// dense solve A x = b (n <= 7) by LU with partial pivoting in DGETF2/DGETRS order (column
// pivot = first largest |a|, row interchanges applied to the whole matrix); A is destroyed,
// b (nrhs columns, b[r][k]) is overwritten with the solution.  Replaces DGESV.
template <int N, int NRHS>
W3D_HD inline int lu_solve(double (&A)[N][N], double (&B)[N][NRHS], int n, int nrhs) {
  int info = 0;
  for (int j = 0; j < n; ++j) {
    int p = j;
    double amax = fabs(A[j][j]);
    for (int i = j + 1; i < n; ++i)
      if (fabs(A[i][j]) > amax) { amax = fabs(A[i][j]); p = i; }
    if (A[p][j] != 0.0) {
      if (p != j) {
        for (int k = 0; k < n; ++k) { const double t = A[j][k]; A[j][k] = A[p][k]; A[p][k] = t; }
        for (int k = 0; k < nrhs; ++k) { const double t = B[j][k]; B[j][k] = B[p][k]; B[p][k] = t; }
      }
      const double r = 1.0 / A[j][j];
      for (int i = j + 1; i < n; ++i) A[i][j] *= r;
    } else if (info == 0) {
      info = j + 1;
    }
    for (int i = j + 1; i < n; ++i)
      for (int k = j + 1; k < n; ++k) A[i][k] -= A[i][j] * A[j][k];
  }
  for (int k = 0; k < nrhs; ++k) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) B[i][k] -= A[i][j] * B[j][k];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) B[i][k] -= A[i][j] * B[j][k];
      B[i][k] /= A[i][i];
    }
  }
  return info;
}

// ---- state of one crystal --------------------------------------------------------------------
template <int NS>
struct Work {
  // quantities at n read from history
  double n_stress[6], n_Rp[3][3], n_R[3][3], n_D[6], n_tt, n_tt_rate, n_gradfe[27];
  // set up for n+1 (mm10_setup): current Schmid tensors and geometric hardening
  double R[3][3], ms[NS][6], qs[NS][3], qc[NS][3], tau_l[NS], mu_harden;
  // the (sub-)increment being solved: curr%D, curr%tinc, curr%dg
  double D[6], tinc, dg, tt_rate;
};

template <int NS>
W3D_HD inline double rs_of(const Work<NS> &w, const double *stress, int i) {   // mm10_rs
  return stress[0] * w.ms[i][0] + stress[1] * w.ms[i][1] + stress[2] * w.ms[i][2] +
         stress[3] * w.ms[i][3] + stress[4] * w.ms[i][4] + stress[5] * w.ms[i][5];
}
template <int NS>
W3D_HD inline double slipinc(const Props<NS> &P, const Work<NS> &w, const double *stress, double tt,
                             int i) {                                          // mm10_slipinc
  const double rs = rs_of(w, stress, i);
  return w.dg / tt * pow(fabs(rs / tt), P.rate_n - 1.0) * rs;
}
template <int NS>
W3D_HD inline double effective_dg(const double *D) {                            // mm10_setup :880
  const double t1 = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
  const double t2 = D[3] * D[3] + D[4] * D[4] + D[5] * D[5];
  return sqrt((2.0 / 3.0) * (t1 + 0.5 * t2));
}

// mm10_form_wp (Voce): w = sum (slipinc + rs tinc iD_v) qc
template <int NS>
W3D_HD inline void form_w(const Props<NS> &P, const Work<NS> &w, const double *stress, double tt,
                          const double (*q)[3], double (&out)[3]) {
  out[0] = out[1] = out[2] = 0.0;
  const int ns = P.nslip();
  for (int i = 0; i < ns; ++i) {
    const double si = slipinc(P, w, stress, tt, i);
    for (int k = 0; k < 3; ++k) out[k] = out[k] + si * q[i][k];
    const double rs = rs_of(w, stress, i);
    for (int k = 0; k < 3; ++k) out[k] = out[k] + rs * w.tinc * P.iD_v * q[i][k];
  }
}
// mm10_formR1 (mm10_b.f:1073): R1 = stress - n%stress - C (D - dbarp) + 2 sym(T W)
template <int NS>
W3D_HD inline void formR1(const Props<NS> &P, const Work<NS> &w, const double *stress, double tt,
                          double *R1) {
  W3D_STAT(0);
  double dbarp[6] = {0, 0, 0, 0, 0, 0};
  const int ns = P.nslip();
  for (int i = 0; i < ns; ++i) {
    const double si = slipinc(P, w, stress, tt, i);
    const double rs = rs_of(w, stress, i);
    const double f = rs * w.tinc * P.iD_v + si;
    for (int k = 0; k < 6; ++k) dbarp[k] = dbarp[k] + f * w.ms[i][k];
  }
  double wp[3], symTW[6], v1[6], v2[6];
  form_w(P, w, stress, tt, w.qc, wp);
  symsw(stress, wp, symTW);
  for (int k = 0; k < 6; ++k) v1[k] = w.D[k] - dbarp[k];
  mv6(v2, P.stiffness, v1);
  for (int k = 0; k < 6; ++k) R1[k] = stress[k] - w.n_stress[k] - v2[k] + 2.0 * symTW[k];
}
// mm10_h_voche (mm10_b.f:1896) and mm10_formR2 (:63); also records curr%tt_rate
template <int NS>
W3D_HD inline double formR2(const Props<NS> &P, Work<NS> &w, const double *stress, double tt) {
  double h = 0.0;
  const int ns = P.nslip();
  for (int i = 0; i < ns; ++i) {
    const double si = slipinc(P, w, stress, tt, i);
    const double h_term = 1.0 - (tt - P.tau_y) / P.tau_v + w.tau_l[i] / (tt - P.tau_y);
    h = h + pow(fabs(h_term), P.voche_m) * copysign(1.0, h_term) * fabs(si);
  }
  h = w.n_tt + P.theta_0 * h;
  w.tt_rate = (h - w.n_tt) / w.tinc;
  return tt - h;
}
template <int NS>
W3D_HD inline void formR(const Props<NS> &P, Work<NS> &w, const double *x, double *R) {
  formR1(P, w, x, x[6], R);
  R[6] = formR2(P, w, x, x[6]);
}
// mm10_formJ11 (mm10_b.f:177) for Voce: sum_i dgdt_i (C ms_i + 2 symtq_i) (x) ms_i + Iw + I
template <int NS>
W3D_HD inline void formJ11(const Props<NS> &P, const Work<NS> &w, const double *stress, double tt,
                           double (&J)[7][7]) {
  W3D_STAT(1);
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) J[r][c] = 0.0;
  const int ns = P.nslip();
  for (int i = 0; i < ns; ++i) {
    double sq[6], wv[6];
    symsw(stress, w.qc[i], sq);                                   // mm10_symSWmat column i
    const double rs = rs_of(w, stress, i);
    double dgdt = pow(fabs(rs), P.rate_n - 1.0);                  // mm10_dgdt_voche
    dgdt = w.dg * P.rate_n / pow(tt, P.rate_n) * dgdt;
    dgdt = dgdt + w.tinc * P.iD_v;
    for (int r = 0; r < 6; ++r) wv[r] = 2.0 * sq[r];              // mult_type_4
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) wv[r] = wv[r] + P.stiffness[r][c] * w.ms[i][c];
    for (int c = 0; c < 6; ++c)                                   // DGER
      for (int r = 0; r < 6; ++r) J[r][c] = J[r][c] + dgdt * wv[r] * w.ms[i][c];
  }
  double wp[3];
  form_w(P, w, stress, tt, w.qc, wp);
  // mm10_IW (mm10_b.f:1613)
  J[0][3] += 2.0 * wp[2]; J[0][5] += -2.0 * wp[1];
  J[1][3] += 2.0 * wp[2]; J[1][4] += -2.0 * wp[0];
  J[2][4] += 2.0 * wp[0]; J[2][5] += 2.0 * wp[1];
  J[3][0] += wp[2]; J[3][1] += -wp[2]; J[3][4] += -wp[1]; J[3][5] += wp[0];
  J[4][1] += wp[0]; J[4][2] += -wp[0]; J[4][3] += wp[1]; J[4][5] += wp[2];
  J[5][0] += wp[1]; J[5][2] += -wp[1]; J[5][3] += wp[0]; J[5][4] += -wp[2];
  for (int r = 0; r < 6; ++r) J[r][r] = J[r][r] + 1.0;
}
// full 7x7 Jacobian (mm10_formJ, mm10_b.f:909) for Voce with one hardening variable
template <int NS>
W3D_HD inline void formJ(const Props<NS> &P, const Work<NS> &w, const double *x, double (&J)[7][7]) {
  const double tt = x[6];
  formJ11(P, w, x, tt, J);
  const int ns = P.nslip();
  for (int r = 0; r < 6; ++r) { J[r][6] = 0.0; J[6][r] = 0.0; }
  double et[6] = {0, 0, 0, 0, 0, 0};
  double etau = 0.0;
  for (int i = 0; i < ns; ++i) {
    // J12 (mm10_formJ12 + mm10_dgdh_voche)
    double sq[6], tm[6];
    symsw(x, w.qc[i], sq);
    for (int r = 0; r < 6; ++r) tm[r] = 2.0 * sq[r];
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) tm[r] = tm[r] + P.stiffness[r][c] * w.ms[i][c];
    const double dgam = slipinc(P, w, x, tt, i);
    const double dgdh = -P.rate_n / tt * dgam;
    for (int r = 0; r < 6; ++r) J[r][6] = J[r][6] + 1.0 * tm[r] * dgdh;
    // J21 = -estress (mm10_estress_voche)
    const double rs = rs_of(w, x, i);
    const double h_term = 1.0 - (tt - P.tau_y) / P.tau_v + w.tau_l[i] / (tt - P.tau_y);
    const double hp = pow(fabs(h_term), P.voche_m);
    const double f = hp * copysign(1.0, h_term) * pow(fabs(rs), P.rate_n - 2.0) * rs;
    for (int k = 0; k < 6; ++k) et[k] = et[k] + f * w.ms[i][k];
    // J22 (mm10_ehard_voche)
    etau = etau + (P.voche_m * (-1.0 / P.tau_v - w.tau_l[i] / ((tt - P.tau_y) * (tt - P.tau_y))) *
                       fabs(dgam) / fabs(h_term) -
                   fabs(dgam) * P.rate_n / tt * copysign(1.0, h_term) * copysign(1.0, dgam)) *
                      hp;
  }
  const double sc = P.theta_0 * w.dg * P.rate_n / pow(tt, P.rate_n);
  for (int k = 0; k < 6; ++k) J[6][k] = -(sc * et[k]);
  etau = P.theta_0 * etau;
  J[6][6] = 1.0 - etau;
}

template <int NS>
W3D_HD inline bool any_nan(const double *x, int n) {
  for (int k = 0; k < n; ++k)
    if (x[k] != x[k]) return true;
  return false;
}

// mm10_solve (mm10_a.f:2901): stress prediction (:3032) then material update (:3224).
// x = [stress(6), tt]; returns fail.  J holds the last Jacobian formed by the update loop.
template <int NS>
W3D_HD inline bool solve(const Props<NS> &P, Work<NS> &w, double *stress, double &tt, double dtinc,
                         double (&J)[7][7]) {
  const double c = 1.0e-4, red = 0.5;
  const int mls = 10, mmin = 1;
  const int miter = P.miter();
  bool fail = false;
  double x[7];
  for (int k = 0; k < 6; ++k) x[k] = stress[k];
  x[6] = tt;
  double inR1;
  {  // ---- mm10_solve_predict
    int iter = 0;
    double x1[6], d1[6], d2[6], R1[6];
    for (int k = 0; k < 6; ++k) x1[k] = x[k];
    for (int k = 0; k < 6; ++k) d1[k] = w.D[k];
    double dtrace = (d1[0] + d1[1] + d1[2]) / 3.0;
    d1[0] -= dtrace; d1[1] -= dtrace; d1[2] -= dtrace;
    double t1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
    double t2 = d1[3] * d1[3] + d1[4] * d1[4] + d1[5] * d1[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d1[k] = 0.0; }
    else { const double s = sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d1[k] = d1[k] / s; }
    for (int k = 0; k < 6; ++k) d2[k] = w.n_D[k];
    dtrace = (d2[0] + d2[1] + d2[2]) / 3.0;
    d2[0] -= dtrace; d2[1] -= dtrace; d2[2] -= dtrace;
    t1 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
    t2 = d2[3] * d2[3] + d2[4] * d2[4] + d2[5] * d2[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d2[k] = 0.0; }
    else { const double s = sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d2[k] = d2[k] / s; }
    t1 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
    t2 = d1[3] * d2[3] + d1[4] * d2[4] + d1[5] * d2[5];
    const double cos_ang = fmax(t1 + t2, 0.0);
    const double x2 = x[6] + cos_ang * w.n_tt_rate * dtinc;
    formR1(P, w, x1, x2, R1);
    double nR1 = sqrt(R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] +
                      R1[4] * R1[4] + R1[5] * R1[5]);
    inR1 = nR1;
    bool pfail = false;
    while ((nR1 > P.atol1) && (nR1 / inR1 > P.rtol1)) {
      formJ11(P, w, x1, x2, J);
      double A[6][6], dx[6][1], wv[6];
      for (int r = 0; r < 6; ++r) {
        dx[r][0] = R1[r];
        for (int cc = 0; cc < 6; ++cc) A[r][cc] = -J[r][cc];
      }
      W3D_STAT(2);
      lu_solve<6, 1>(A, dx, 6, 1);
      double alpha = 1.0;
      double dotR1 = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] +
                     R1[4] * R1[4] + R1[5] * R1[5];
      const double ls1 = 0.5 * dotR1;
      for (int k = 0; k < 6; ++k)   // trans(J11) R1
        wv[k] = J[0][k] * R1[0] + J[1][k] * R1[1] + J[2][k] * R1[2] + J[3][k] * R1[3] +
                J[4][k] * R1[4] + J[5][k] * R1[5];
      const double ls2 = c * (dx[0][0] * wv[0] + dx[1][0] * wv[1] + dx[2][0] * wv[2] +
                              dx[3][0] * wv[3] + dx[4][0] * wv[4] + dx[5][0] * wv[5]);
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xnew[6];
        for (int k = 0; k < 6; ++k) xnew[k] = x1[k] + alpha * dx[k][0];
        formR1(P, w, xnew, x2, R1);
        dotR1 = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] +
                R1[5] * R1[5];
        nR1 = sqrt(dotR1);
        const double nRs = 0.5 * dotR1;
        if ((nRs <= nlsx) || (ls > mls)) {
          for (int k = 0; k < 6; ++k) x1[k] = xnew[k];
          break;
        }
        alpha = red * alpha;
        ls = ls + 1;
      }
      iter = iter + 1;
      if ((iter > miter) || any_nan<NS>(x1, 6)) { pfail = true; break; }
    }
    if (pfail) {
      fail = true;   // the reference returns from the predictor only; the update still runs on
                     // the un-predicted x and fail stays set (mm10_a.f:3187-3203)
    } else {
      for (int k = 0; k < 6; ++k) x[k] = x1[k];
      x[6] = x2;
    }
  }
  {  // ---- mm10_solve_update
    int iter = 0;
    double R[7];
    formR(P, w, x, R);
    double dotR = 0.0;
    for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
    double nR = sqrt(dotR);
    double inR = nR;
    if (inR == 0.0) inR = inR1;
    while (((nR > P.atol) && (nR / inR > P.rtol)) || (iter < mmin)) {
      formJ(P, w, x, J);
      double A[7][7], dx[7][1], wv[7];
      for (int r = 0; r < 7; ++r) {
        dx[r][0] = R[r];
        for (int cc = 0; cc < 7; ++cc) A[r][cc] = -J[r][cc];
      }
      W3D_STAT(2);
      lu_solve<7, 1>(A, dx, 7, 1);
      double alpha = 1.0;
      dotR = 0.0;
      for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
      const double ls1 = 0.5 * dotR;
      for (int k = 0; k < 7; ++k) {   // trans(J) R
        double s = 0.0;
        for (int r = 0; r < 7; ++r) s = s + J[r][k] * R[r];
        wv[k] = s;
      }
      double ls2 = 0.0;
      for (int k = 0; k < 7; ++k) ls2 = ls2 + dx[k][0] * wv[k];
      ls2 = c * ls2;
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xnew[7];
        for (int k = 0; k < 7; ++k) xnew[k] = x[k] + alpha * dx[k][0];
        formR(P, w, xnew, R);
        dotR = 0.0;
        for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
        nR = sqrt(dotR);
        const double nRs = 0.5 * dotR;
        if ((nRs <= nlsx) || (ls > mls)) {
          for (int k = 0; k < 7; ++k) x[k] = xnew[k];
          break;
        }
        alpha = red * alpha;
        ls = ls + 1;
      }
      iter = iter + 1;
      if ((iter > miter) || any_nan<NS>(x, 7)) { fail = true; return fail; }
    }
  }
  for (int k = 0; k < 6; ++k) stress[k] = x[k];
  tt = x[6];
  return fail;
}

// result of one crystal update (what mm10_a_do_crystal accumulates / mm10_store_cryhist stores)
template <int NS>
struct CrystalOut {
  double stress[6], tangent[6][6], Rp[3][3], euler[3], eps[6], slip_incs[NS], tt, tt_rate, u[15],
      ep[6], ed[6], work_inc, p_work_inc, p_strain_inc, p_strain_ten[6];
};

// mm10_setup (mm10_a.f:852) for the Voce model
template <int NS>
W3D_HD inline void setup(const Props<NS> &P, Work<NS> &w) {
  double tr[3][3], RE[6][6], RW[3][3], RWC[3][3], wk[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tr[i][j] = w.n_Rp[j][i];
  rt2rve(tr, RE);
  rt2rvw(tr, RW);
  mat33(wk, w.R, tr);
  rt2rvw(wk, RWC);
  const int ns = P.nslip();
  for (int i = 0; i < ns; ++i) {
    double a6[6];
    mv6(a6, RE, P.ms[i]);
    for (int k = 0; k < 6; ++k) w.ms[i][k] = a6[k];
    mv3(w.qs[i], RW, P.qs[i]);
    mv3(w.qc[i], RWC, P.qs[i]);
  }
  w.mu_harden = P.stiffness[5][5];   // mm10_setup_voche
  // geometric hardening tau_l from the lattice curvature (Acharya), :925-975
  double curv[3][3][3];
  const double *g = w.n_gradfe;     // gradFeinv(i,j,k) column-major: i + 3 j + 9 k
#define GF(i, j, k) g[(i) + 3 * (j) + 9 * (k)]
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) curv[i][j][k] = GF(i, j, k) - GF(i, k, j);
#undef GF
  const double alpha = 1.0 / 3.0;
  const double cst = P.k_0 * P.burgers * alpha * alpha * w.mu_harden * w.mu_harden / 2.0 / P.theta_0;
  for (int t = 0; t < ns; ++t) {
    double wv[3], cn[3], tm[3];
    mv3(wv, tr, P.ns[t]);
    mv3(cn, w.n_R, wv);
    for (int i = 0; i < 3; ++i)
      tm[i] = 0.0 + 0.5 * (-curv[i][0][1] * cn[2] + curv[i][0][2] * cn[1] + curv[i][1][0] * cn[2] -
                           curv[i][1][2] * cn[0] - curv[i][2][0] * cn[1] + curv[i][2][1] * cn[0]);
    w.tau_l[t] = cst * sqrt(tm[0] * tm[0] + tm[1] * tm[1] + tm[2] * tm[2]);
  }
}

// One crystal: mm10_a_do_crystal -> mm10_solve_crystal and everything below it.
// Returns true when the crystal asks for a global step cut (solver failed after 4 halvings).
template <int NS>
W3D_HD inline bool crystal_update(const Props<NS> &P, Work<NS> &w, const double *uddt, double dt,
                                  bool iter_0_extrapolate_off, bool asymmetric,
                                  CrystalOut<NS> &o) {
  const int ns = P.nslip();
  // mm10_setup_np1 defaults for everything mm10_solve_crystal may leave untouched
  for (int k = 0; k < 6; ++k) { o.stress[k] = 0.0; o.eps[k] = 0.0; o.ep[k] = 0.0; o.ed[k] = 0.0; o.p_strain_ten[k] = 0.0; }
  for (int i = 0; i < 3; ++i) { o.euler[i] = 0.0; for (int j = 0; j < 3; ++j) o.Rp[i][j] = 0.0; }
  for (int i = 0; i < NS; ++i) o.slip_incs[i] = 0.0;
  for (int i = 0; i < 15; ++i) o.u[i] = 0.0;
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) o.tangent[r][c] = 0.0;
  o.tt = 0.0; o.tt_rate = 0.0; o.work_inc = 0.0; o.p_work_inc = 0.0; o.p_strain_inc = 0.0;

  // ---- mm10_solve_strup
  for (int k = 0; k < 6; ++k) w.D[k] = uddt[k];
  w.tinc = dt;
  w.dg = effective_dg<NS>(w.D);
  if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;
  setup(P, w);
  double stress[6], ostress[6];
  for (int k = 0; k < 6; ++k) { stress[k] = w.n_stress[k]; ostress[k] = stress[k]; }
  double tt = w.n_tt, ott = tt;
  const double temp1 = uddt[0] * uddt[0] + uddt[1] * uddt[1] + uddt[2] * uddt[2] + uddt[3] * uddt[3] +
                       uddt[4] * uddt[4] + uddt[5] * uddt[5];
  const double temp2 = stress[0] * stress[0] + stress[1] * stress[1] + stress[2] * stress[2] +
                       stress[3] * stress[3] + stress[4] * stress[4] + stress[5] * stress[5];
  const bool no_load = (temp2 == 0.0) && (temp1 == 0.0);
  double J[7][7];
  if (iter_0_extrapolate_off || no_load) {
    // mm10_solve_strup_a: elastic stiffness as tangent; stress corrected for creep only
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) o.tangent[r][c] = P.stiffness[r][c];
    if (!no_load) {
      double R1[6];
      formR1(P, w, stress, tt, R1);
      for (int k = 0; k < 6; ++k) stress[k] = stress[k] - R1[k];
    }
    for (int k = 0; k < 6; ++k) o.stress[k] = stress[k];
    o.tt = tt;
    o.tt_rate = 0.0;
    return false;
  }
  W3D_STAT(3);
  {  // mm10_solve_strup_iterate: adaptive sub-incrementation, at most 4 halvings
    double frac = 0.0, step = 1.0;
    int cuts = 0;
    const int mcuts = 4;
    bool fail = false;
    while (frac < 1.0) {
      for (int k = 0; k < 6; ++k) w.D[k] = uddt[k] * (step + frac);
      w.tinc = dt * (step + frac);
      w.dg = effective_dg<NS>(w.D);
      if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;
      tt = w.n_tt;   // "some subroutines modify n%tau_tilde in setup": tt restarts from n
      W3D_STAT(4);
      fail = solve(P, w, stress, tt, dt * step, J);
      if (fail) {
        for (int k = 0; k < 6; ++k) stress[k] = ostress[k];
        tt = ott;
        step = step * 0.5;
        cuts = cuts + 1;
        if (cuts > mcuts) break;
        fail = false;
      } else {
        for (int k = 0; k < 6; ++k) ostress[k] = stress[k];
        ott = tt;
        frac = frac + step;
      }
    }
    if (fail || tt != tt || any_nan<NS>(stress, 6)) return true;
  }
  for (int k = 0; k < 6; ++k) o.stress[k] = stress[k];
  o.tt = tt;
  o.tt_rate = w.tt_rate;
  // from here on the full-step quantities of np1 apply (np1%D, np1%tinc, np1%dg)
  for (int k = 0; k < 6; ++k) w.D[k] = uddt[k];
  w.tinc = dt;
  w.dg = effective_dg<NS>(w.D);
  if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;

  // ---- mm10_tangent (mm10_a.f:664): T = JJ^-1 (C - JA - JB); for Voce ed = 0 and dgdd = 0 so
  //      JA = JB = 0 and JJ = J11 - J12 J22^-1 J21 with the Jacobian of the last solver pass
  {
    double JJ[6][6], JR[6][6];
    double b21[6];
    for (int k = 0; k < 6; ++k) b21[k] = J[6][k] / J[6][6];   // DGESV on the 1x1 J22
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        JJ[r][c] = J[r][c] + (-1.0) * J[r][6] * b21[c];
        JR[r][c] = P.stiffness[r][c] - 0.0 - 0.0 * J[r][6];
      }
    lu_solve<6, 6>(JJ, JR, 6, 6);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) o.tangent[r][c] = JR[r][c];
    if (!asymmetric) {   // mm10_a_make_symm_1
      double wk[6][6];
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) wk[r][c] = (o.tangent[r][c] + o.tangent[c][r]) * 0.5;
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) o.tangent[r][c] = wk[r][c];
    }
  }
  // ---- mm10_update_rotation: Rp_{n+1} = exp(wbarp) Rp_n
  {
    double wb[3], W[3][3], A[3][3];
    form_w(P, w, stress, tt, w.qs, wb);
    W[0][0] = 0.0; W[0][1] = wb[2]; W[0][2] = wb[1];
    W[1][0] = -wb[2]; W[1][1] = 0.0; W[1][2] = wb[0];
    W[2][0] = -wb[1]; W[2][1] = -wb[0]; W[2][2] = 0.0;
    const double alpha = sqrt(W[1][2] * W[1][2] + W[0][2] * W[0][2] + W[0][1] * W[0][1]);
    if (alpha < 1.0e-16) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = 0.0;
    } else {
      double WW[3][3];
      mat33(WW, W, W);
      const double ca = (1.0 - cos(alpha)) / (alpha * alpha), sa = sin(alpha) / alpha;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = ca * WW[i][j] + sa * W[i][j];
    }
    A[0][0] += 1.0; A[1][1] += 1.0; A[2][2] += 1.0;
    mat33(o.Rp, A, w.n_Rp);
  }
  // ---- mm10_output (mm10_a.f:3493)
  {
    const double pi = 3.141592653589793, one_eighty = 180.0;
    for (int i = 0; i < ns; ++i) o.slip_incs[i] = slipinc(P, w, stress, tt, i);
    {  // mm10_update_euler_angles
      double w1[3][3], full[3][3];
      mat33_bt(w1, o.Rp, w.R);
      mat33(full, P.g, w1);
      double psi = atan2(full[2][1], full[2][0]);
      if (psi < 0.0) psi = psi + 2.0 * pi;
      double phi = atan2(full[1][2], full[0][2]);
      if (phi < 0.0) phi = phi + 2.0 * pi;
      if (full[2][2] > 1.0) full[2][2] = 1.0;
      const double theta = acos(full[2][2]);
      if ((int)P.angle_type_d == 1) {
        o.euler[0] = one_eighty / pi * psi; o.euler[1] = one_eighty / pi * theta; o.euler[2] = one_eighty / pi * phi;
      } else {
        o.euler[0] = psi; o.euler[1] = theta; o.euler[2] = phi;
      }
    }
    double ed[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ns; ++i) {
      const double dif = rs_of(w, stress, i) * w.tinc * P.iD_v;
      for (int k = 0; k < 6; ++k) ed[k] = ed[k] + dif * w.ms[i][k];
    }
    for (int k = 0; k < 6; ++k) o.ed[k] = ed[k] / w.tinc;
    double t1 = ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2];
    double t2 = ed[3] * ed[3] + ed[4] * ed[4] + ed[5] * ed[5];
    double ed_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ed_dot = ed_dot / w.tinc;
    o.u[14] = ed_dot;
    o.work_inc = stress[0] * w.D[0] + stress[1] * w.D[1] + stress[2] * w.D[2] + stress[3] * w.D[3] +
                 stress[4] * w.D[4] + stress[5] * w.D[5];
    // lattice strain: C^-1 stress (DPOSV -> LU here), rotated with R
    double S[6][6], ee0[6][1], erot[6][6], ee[6];
    for (int r = 0; r < 6; ++r) { ee0[r][0] = stress[r]; for (int c = 0; c < 6; ++c) S[r][c] = P.stiffness[r][c]; }
    lu_solve<6, 1>(S, ee0, 6, 1);
    rt2rve(w.R, erot);
    { double v[6]; for (int k = 0; k < 6; ++k) v[k] = ee0[k][0]; mv6(ee, erot, v); }
    double dbarp[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ns; ++i)
      for (int k = 0; k < 6; ++k) dbarp[k] = dbarp[k] + o.slip_incs[i] * w.ms[i][k];
    double wp[3], ewwe[6], ep[6];
    form_w(P, w, stress, tt, w.qc, wp);
    symsw(ee, wp, ewwe);
    for (int k = 0; k < 6; ++k) { ep[k] = dbarp[k] + ewwe[k]; o.ep[k] = ep[k] / w.tinc; }
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    double ep_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ep_dot = ep_dot / w.tinc;
    o.u[10] = ep_dot;
    for (int k = 0; k < 6; ++k) { ep[k] = ep[k] + ed[k]; o.p_strain_ten[k] = ep[k]; }
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    o.p_strain_inc = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    o.p_work_inc = stress[0] * ep[0] + stress[1] * ep[1] + stress[2] * ep[2] + stress[3] * ep[3] +
                   stress[4] * ep[4] + stress[5] * ep[5];
    for (int k = 0; k < 6; ++k) o.eps[k] = ee[k];
    const double ec_dot = o.p_strain_inc / w.tinc;
    double n_eff;
    if (ec_dot > 0.0) {
      for (int k = 0; k < 6; ++k) ep[k] = ep[k] / w.tinc;
      n_eff = 0.0;
      for (int i = 0; i < ns; ++i) {
        const double rs = rs_of(w, stress, i);
        double dgdt = pow(fabs(rs), P.rate_n - 1.0);
        dgdt = w.dg * P.rate_n / pow(tt, P.rate_n) * dgdt;
        dgdt = dgdt + w.tinc * P.iD_v;
        t1 = w.ms[i][0] * ep[0] + w.ms[i][1] * ep[1] + w.ms[i][2] * ep[2];
        t2 = w.ms[i][3] * ep[3] + w.ms[i][4] * ep[4] + w.ms[i][5] * ep[5];
        const double ec_slip = t1 + 0.5 * t2;
        t1 = rs * dgdt / ec_dot;
        t2 = ec_slip / ec_dot / w.tinc;
        n_eff = n_eff + (2.0 / 3.0) * t1 * t2;
      }
    } else {
      n_eff = 1.0e10;
    }
    o.u[11] = n_eff;
    double s_trace = (stress[0] + stress[1] + stress[2]) / 3.0;
    double se[6];
    for (int k = 0; k < 6; ++k) se[k] = stress[k];
    se[0] -= s_trace; se[1] -= s_trace; se[2] -= s_trace;
    t1 = se[0] * se[0] + se[1] * se[1] + se[2] * se[2];
    t2 = se[3] * se[3] + se[4] * se[4] + se[5] * se[5];
    s_trace = sqrt(1.5 * (t1 + 2.0 * t2));
    o.u[12] = s_trace;
    double B_eff;
    if (ec_dot < 1.0e-100) B_eff = 0.0;
    else if (n_eff > 100.0) B_eff = -1.0;
    else B_eff = ec_dot / pow(s_trace, n_eff);
    o.u[13] = B_eff;
    for (int i = 0; i < ns; ++i)
      o.slip_incs[i] = o.slip_incs[i] + rs_of(w, stress, i) * w.tinc * P.iD_v;
    double maxslip = 0.0;
    int sysID = 0;
    for (int i = 0; i < ns; ++i) {
      const double cur = fabs(o.slip_incs[i]);
      if (cur > maxslip) { maxslip = cur; sysID = i + 1; }
    }
    o.u[5] = maxslip / w.tinc;
    o.u[6] = (double)sysID;
    int numAct = 0;
    maxslip = 0.1 * maxslip;
    for (int i = 0; i < ns; ++i)
      if (fabs(o.slip_incs[i]) >= maxslip) numAct = numAct + 1;
    o.u[7] = (double)numAct;
  }
  return false;
}

// mm10 (mm10_a.f:29): all crystals of one Gauss point.  hn / h1: history at n / n+1;
// props: ncrystals consecutive Props records of this element; Rur: rot_blk_n1 (R of F = R U at
// n+1, identity for small strain), column-major 9; uddt: un-rotated strain increment;
// cgn / cgn1: urcs at n / n+1 (9).  Returns true when a crystal requests a step cut.
template <int NS>
W3D_HD inline bool gauss_point_update(int step, int ncrystals, const Props<NS> *props, const Layout &L,
                                      const CHist &hn, const Hist &h1, const double *Rur,
                                      const double *uddt, double dt, bool iter_0_extrapolate_off,
                                      bool asymmetric, const double *cgn, double *cgn1) {
  const int ns = L.nslip;
  // At step 1 the reference initialises its block-local copy of history at n
  // (mm10_init_general_hist / _uout_hist / _slip_hist, mm10_init_cc_hist0) and never writes that
  // copy back; the same values are substituted here on the fly, history at n stays untouched.
  const bool init = (step == 1);
  double sig_avg[6] = {0, 0, 0, 0, 0, 0}, tang_avg[6][6], p_strain_ten[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) tang_avg[r][c] = 0.0;
  double t_work_inc = 0.0, p_work_inc = 0.0, p_strain_inc = 0.0;
  // the slip-history sum is accumulated directly into history at n+1 (zeroed first)
  for (int k = 0; k < L.len_slip; ++k) h1[L.o_sliphist() + k] = 0.0;

  Work<NS> w;
  CrystalOut<NS> o;
  for (int c = 0; c < ncrystals; ++c) {
    const Props<NS> &P = props[c];
    // mm10_copy_cc_hist
    if (init) {
      for (int k = 0; k < 27; ++k) w.n_gradfe[k] = 0.0;
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { w.n_R[i][j] = (i == j) ? 1.0 : 0.0; w.n_Rp[i][j] = (i == j) ? 1.0 : 0.0; }
      for (int k = 0; k < 6; ++k) { w.n_stress[k] = cgn[k]; w.n_D[k] = 0.0; }
      w.n_tt = P.tau_y + 1.0e-5;   // mm10_init_voche
      w.n_tt_rate = 0.0;
    } else {
      for (int k = 0; k < 27; ++k) w.n_gradfe[k] = hn[L.o_gradfe() + k];
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.n_R[i][j] = hn[L.o_R() + j * 3 + i];
      for (int k = 0; k < 6; ++k) w.n_stress[k] = hn[L.c_stress(c) + k];
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.n_Rp[i][j] = hn[L.c_Rp(c) + j * 3 + i];
      for (int k = 0; k < 6; ++k) w.n_D[k] = hn[L.c_D(c) + k];
      w.n_tt = hn[L.c_tt(c)];
      w.n_tt_rate = hn[L.c_rate(c)];
    }
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.R[i][j] = Rur[j * 3 + i];
    w.tt_rate = 0.0;

    const bool cut = crystal_update(P, w, uddt, dt, iter_0_extrapolate_off, asymmetric, o);
    if (cut) return true;

    for (int k = 0; k < 6; ++k) sig_avg[k] = sig_avg[k] + o.stress[k];
    for (int r = 0; r < 6; ++r) for (int cc = 0; cc < 6; ++cc) tang_avg[r][cc] = tang_avg[r][cc] + o.tangent[r][cc];
    for (int k = 0; k < ns; ++k) h1[L.o_sliphist() + k] = h1[L.o_sliphist() + k] + o.slip_incs[k];
    t_work_inc = t_work_inc + o.work_inc;
    p_work_inc = p_work_inc + o.p_work_inc;
    p_strain_inc = p_strain_inc + o.p_strain_inc;
    for (int k = 0; k < 6; ++k) p_strain_ten[k] = p_strain_ten[k] + o.p_strain_ten[k];
    // mm10_store_cryhist
    for (int k = 0; k < 6; ++k) h1[L.c_stress(c) + k] = o.stress[k];
    for (int k = 0; k < 3; ++k) h1[L.c_euler(c) + k] = o.euler[k];
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) h1[L.c_Rp(c) + j * 3 + i] = o.Rp[i][j];
    for (int k = 0; k < 6; ++k) h1[L.c_D(c) + k] = uddt[k];
    for (int k = 0; k < 6; ++k) h1[L.c_eps(c) + k] = o.eps[k];
    for (int k = 0; k < L.len_slip; ++k) h1[L.c_slip(c) + k] = (k < NS) ? o.slip_incs[k] : 0.0;
    for (int k = 0; k < L.len_tt; ++k) h1[L.c_tt(c) + k] = 0.0;
    h1[L.c_tt(c)] = o.tt;
    for (int k = 0; k < L.len_u; ++k) h1[L.c_u(c) + k] = (k < 15) ? o.u[k] : 0.0;
    for (int k = 0; k < L.len_rate; ++k) h1[L.c_rate(c) + k] = 0.0;
    h1[L.c_rate(c)] = o.tt_rate;
    for (int k = 0; k < 6; ++k) h1[L.c_ep(c) + k] = o.ep[k];
    for (int k = 0; k < 6; ++k) h1[L.c_ed(c) + k] = o.ed[k];
  }
  // mm10_a_crystal_avgs + mm10_a_store_crystal
  const double rn = (double)ncrystals;
  for (int k = 0; k < 6; ++k) cgn1[k] = sig_avg[k] / rn;
  for (int cc = 0; cc < 6; ++cc)
    for (int r = 0; r < 6; ++r) h1[L.o_D() + cc * 6 + r] = tang_avg[r][cc] / rn;   // column-major
  for (int k = 0; k < 9; ++k) h1[L.o_R() + k] = Rur[k];
  for (int k = 0; k < L.len_slip; ++k)
    h1[L.o_sliphist() + k] = (init ? 0.0 : hn[L.o_sliphist() + k]) + h1[L.o_sliphist() + k] / rn;
  t_work_inc = t_work_inc / rn; p_work_inc = p_work_inc / rn; p_strain_inc = p_strain_inc / rn;
  cgn1[6] = cgn[6] + t_work_inc;
  cgn1[7] = cgn[7] + p_work_inc;
  cgn1[8] = cgn[8] + p_strain_inc;
  h1[L.o_work() + 0] = (init ? 0.0 : hn[L.o_work() + 0]) + t_work_inc;
  h1[L.o_work() + 1] = (init ? 0.0 : hn[L.o_work() + 1]) + p_work_inc;
  h1[L.o_work() + 2] = (init ? 0.0 : hn[L.o_work() + 2]) + p_strain_inc;
  return false;
}

// mm10_calc_grads (mm10_a.f:1579) for the 2x2x2 rule: least-squares linear fit of
// Rp^T R (geonl) / Rp (small strain) over the 8 Gauss points; the DGELS solution of the fixed
// 8x4 system [xi eta zeta 1] has the closed form  grad_k = sum_i x_ik f_i / (8 l^2).
// rt[gp][9]: the fitted tensors column-major; jacinv: inverse of the reference's local_work%jac
// (the last Jacobian formed, undeformed configuration, Gauss point 8); out: grads(3,3,3) col-major.
W3D_HD inline void calc_grads(const double (*rt)[9], bool geonl, const double (&jacinv)[3][3],
                              double *out) {
  const double l = 0.5773502691896257645;
  const int sx[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  const int sy[8] = {-1, 1, -1, 1, -1, 1, -1, 1};
  const int sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) {
      double g0 = 0.0, g1 = 0.0, g2 = 0.0;
      for (int i = 0; i < 8; ++i) {
        const double f = rt[i][b * 3 + a];
        g0 += sx[i] * f; g1 += sy[i] * f; g2 += sz[i] * f;
      }
      const double s = 1.0 / (8.0 * l);
      g0 *= s; g1 *= s; g2 *= s;
      double r0 = g0, r1 = g1, r2 = g2;
      if (geonl) {
        r0 = jacinv[0][0] * g0 + jacinv[0][1] * g1 + jacinv[0][2] * g2;
        r1 = jacinv[1][0] * g0 + jacinv[1][1] * g1 + jacinv[1][2] * g2;
        r2 = jacinv[2][0] * g0 + jacinv[2][1] * g1 + jacinv[2][2] * g2;
      }
      out[a + 3 * b + 9 * 0] = r0;
      out[a + 3 * b + 9 * 1] = r1;
      out[a + 3 * b + 9 * 2] = r2;
    }
}

// ---- host-side property setup (setup_mm10_rknstr, rknstr.f:1568-2027) -------------------------
// crystal-type record = the c_array(cnum) fields the path reads (mod_crystals.f:148-236)
struct CrystalType {
  double elast_stiff[6][6];        // row-major
  double bi[48][3], ni[48][3];     // slip directions / normals in crystal axes
  double harden_n, tau_y, tau_v, voche_m, theta_o, k_o, b, iD_v, eps_dot_o_y;
  double atol, atol1, rtol, rtol1;
  int nslip, num_hard, h_type, tang_calc, miter, solver, alter_mode;
};

// mm10_rotation_matrix (mm10_a.f:1310), Kocks convention
inline void rotation_matrix(const double *angles, int angle_type /*1 degrees, 2 radians*/,
                            double (&r)[3][3]) {
  const double pi = 3.141592653589793, one_eighty = 180.0;
  double a = angles[0], b = angles[1], c = angles[2];
  if (angle_type == 1) { a = a * pi / one_eighty; b = b * pi / one_eighty; c = c * pi / one_eighty; }
  const double psi = a, theta = b, phi = c;
  r[0][0] = -sin(psi) * sin(phi) - cos(psi) * cos(phi) * cos(theta);
  r[0][1] = cos(psi) * sin(phi) - sin(psi) * cos(phi) * cos(theta);
  r[0][2] = cos(phi) * sin(theta);
  r[1][0] = sin(psi) * cos(phi) - cos(psi) * sin(phi) * cos(theta);
  r[1][1] = -cos(psi) * cos(phi) - sin(psi) * sin(phi) * cos(theta);
  r[1][2] = sin(phi) * sin(theta);
  r[2][0] = cos(psi) * sin(theta);
  r[2][1] = sin(psi) * sin(theta);
  r[2][2] = cos(theta);
}

template <int NS>
inline void setup_props(const CrystalType &ct, const double *angles, int angle_type, Props<NS> &P) {
  P.rate_n = ct.harden_n; P.tau_y = ct.tau_y; P.tau_v = ct.tau_v; P.voche_m = ct.voche_m;
  P.theta_0 = ct.theta_o; P.k_0 = ct.k_o; P.burgers = ct.b; P.iD_v = ct.iD_v;
  P.eps_dot_0_y = ct.eps_dot_o_y;
  P.atol = ct.atol; P.atol1 = ct.atol1; P.rtol = ct.rtol; P.rtol1 = ct.rtol1;
  for (int k = 0; k < 3; ++k) P.init_angles[k] = angles[k];
  P.nslip_d = ct.nslip; P.miter_d = ct.miter; P.alter_mode_d = ct.alter_mode; P.angle_type_d = angle_type;
  rotation_matrix(angles, angle_type, P.g);
  double tr[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tr[i][j] = P.g[j][i];
  for (int s = 0; s < NS; ++s) {
    for (int k = 0; k < 6; ++k) P.ms[s][k] = 0.0;
    for (int k = 0; k < 3; ++k) { P.qs[s][k] = 0.0; P.ns[s][k] = 0.0; }
  }
  for (int s = 0; s < ct.nslip && s < NS; ++s) {
    double bs[3], nn[3], A[3][3];
    mv3(bs, tr, ct.bi[s]);
    mv3(nn, tr, ct.ni[s]);
    for (int k = 0; k < 3; ++k) P.ns[s][k] = nn[k];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = bs[i] * nn[j];
    double As[3][3], Aa[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { As[i][j] = 0.5 * (A[i][j] + A[j][i]); Aa[i][j] = 0.5 * (A[i][j] - A[j][i]); }
    P.ms[s][0] = As[0][0]; P.ms[s][1] = As[1][1]; P.ms[s][2] = As[2][2];       // mm10_ET2EV
    P.ms[s][3] = 2.0 * As[0][1]; P.ms[s][4] = 2.0 * As[1][2]; P.ms[s][5] = 2.0 * As[0][2];
    P.qs[s][0] = Aa[1][2]; P.qs[s][1] = Aa[0][2]; P.qs[s][2] = Aa[0][1];       // mm10_WT2WV
  }
  // rotate the stiffness: C = Rstiff C0 Rstiff^T with Rstiff = RT2RVE(g^T)
  double Rs[6][6], t66[6][6];
  rt2rve(tr, Rs);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += ct.elast_stiff[i][k] * Rs[j][k];
    t66[i][j] = s;
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += Rs[i][k] * t66[k][j];
    P.stiffness[i][j] = s;
  }
}

}  // namespace w3do_mm10
