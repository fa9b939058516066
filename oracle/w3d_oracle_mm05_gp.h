/* w3d_oracle_mm05_gp.h -- CPU ORACLE (test infrastructure only; never part of the product path).
 *
 * The generalized_plasticity option of the reference's cyclic plasticity model mm05, restated from
 * src/mm05.f a second time in the Fortran's own BLOCK form (see w3d_oracle_mm05_fa.h for why and
 * how: written in round 2 directly from the Fortran, without consulting the per-point
 * transcription w3d_oracle_mm05.h that the engine's device core descends from).
 *
 *   mm05_gp                    src/mm05.f:1882-2050
 *   mm05_gp_step1_set          :2059-2120
 *   mm05_gp_init               :2128-2365
 *   mm05_gp_compute            :2372-2640
 *   mm05_gp_nsteps             :2651-2850
 *   mm05_gp_lambda_solve       :2859-2995
 *   mm05_gp_elastic_fraction   :3002-3420  (Ridder's method; golden-section search for the minimum
 *                                           of the yield function when it is positive at both ends)
 *   mm05_gp_felas_eta          :3429-3487
 *   cnst5_gp                   :3584-3870
 *   mm05_final, mm05_plastic_work: shared with the other option (w3d_oracle_mm05_fa.h)
 *
 * Property vectors as setup_mm05_rknstr (src/rknstr.f:1234-1317) and drive_05_update_c
 * (src/rstgp1.f:2627-2700) fill them, isothermal: values at n and n+1 are the same numbers but the
 * arithmetic that mixes them (lk = hk_n1 / hk_n, mk = g_j1 / g_j ...) is kept.  Bare REAL*4
 * literals are evaluated in double (-fpconstant, W3DO5F_K).  Parity status: pinned end to end by
 * verification/test60 b; function level at 1e-10 unpinned (no Fortran compiler here).
 */
#ifndef W3D_ORACLE_MM05_GP_H
#define W3D_ORACLE_MM05_GP_H
#include "w3d_oracle_mm05_fa.h"

namespace w3do_mm05_gp {

using w3do_mm05_fa::Block;
using w3do_mm05_fa::fsign;
using w3do_mm05_fa::state_word;
enum { OK = 0, CUT = 1, ABORT_LAMBDA_COMPLEX = 31, ABORT_LAMBDA_NEGATIVE = 32, ABORT_LAMBDA_QUADRATIC = 33,
       ABORT_ETA_ONE = 34, ABORT_ETA_BRACKET = 35, ABORT_RIDDER = 36, ABORT_GOLDEN = 37 };

// the additional vectors of the generalized_plasticity option (values at n+1 and at n)
struct GpVectors {
  double *e_np1, *e_n, *nu_np1, *nu_n, *sigyld_np1, *sigyld_n, *h_np1, *h_n, *beta_np1, *beta_n, *delta_np1,
      *delta_n, *tau;
};

// mm05_gp_step1_set
inline void step1_set(Block &B, const GpVectors &V) {
  const double root2third = W3DO5F_K(0.81649658), zero = W3DO5F_K(0.0);
  const double dword = state_word(3);
  for (int i = 0; i < B.span; ++i) {
    B.stress_n[i][6] = zero;
    B.stress_n[i][7] = zero;
    B.stress_n[i][8] = zero;
    B.history_n[i][0] = zero;
    B.history_n[i][1] = V.sigyld_n[i] * root2third;
    B.history_n[i][2] = zero;
    B.history_n[i][3] = dword;
    for (int k = 4; k <= 13; ++k) B.history_n[i][k] = zero;
  }
}

// mm05_gp_init
inline void init(Block &B, const GpVectors &V, int iter, int *iostat, bool *prior_linear, double *g_vec_np1,
                 double *g_vec_n, double (*alpha_n)[6], int *instat, bool *yield, double *trace_eps_np1) {
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), three = W3DO5F_K(3.0),
               twothird = W3DO5F_K(0.66666666), yld_tol = W3DO5F_K(0.0000001), root2third = W3DO5F_K(0.8164965809);
  (void)zero;
  for (int i = 0; i < B.span; ++i) {
    int32_t iw[2];
    std::memcpy(iw, &B.history_n[i][3], 8);
    iostat[i] = iw[0];
    prior_linear[i] = iostat[i] == 3;
    const double *deps = B.deps[i];
    const double trace_deps = deps[0] + deps[1] + deps[2];
    const double deps_mean = trace_deps / three;
    const double de1 = deps[0] - deps_mean, de2 = deps[1] - deps_mean, de3 = deps[2] - deps_mean;
    const double de4 = deps[3], de5 = deps[4], de6 = deps[5];
    const double e_n = V.e_n[i], nu_n = V.nu_n[i];
    const double g_n = e_n / (two * (one + nu_n));
    const double *sn = B.stress_n[i];
    const double een1 = (sn[0] - nu_n * (sn[1] + sn[2])) / e_n;
    const double een2 = (sn[1] - nu_n * (sn[0] + sn[2])) / e_n;
    const double een3 = (sn[2] - nu_n * (sn[0] + sn[1])) / e_n;
    const double een4 = sn[3] / g_n, een5 = sn[4] / g_n, een6 = sn[5] / g_n;
    g_vec_n[i] = g_n;
    trace_eps_np1[i] = een1 + een2 + een3 + trace_deps;
    const double eps_mean = (een1 + een2 + een3) / three;
    const double e1 = (een1 - eps_mean) + de1, e2 = (een2 - eps_mean) + de2, e3 = (een3 - eps_mean) + de3;
    const double e4 = een4 + de4, e5 = een5 + de5, e6 = een6 + de6;
    const double e_n1 = V.e_np1[i], nu_n1 = V.nu_np1[i];
    const double g_n1 = e_n1 / (two * (one + nu_n1));
    B.trial[i][0] = two * g_n1 * e1;
    B.trial[i][1] = two * g_n1 * e2;
    B.trial[i][2] = two * g_n1 * e3;
    B.trial[i][3] = g_n1 * e4;
    B.trial[i][4] = g_n1 * e5;
    B.trial[i][5] = g_n1 * e6;
    g_vec_np1[i] = g_n1;
    for (int k = 0; k < 6; ++k) alpha_n[i][k] = B.history_n[i][5 + k];
    const double hk_n = (one - V.tau[i]) * twothird * V.h_n[i];
    const double hk_n1 = (one - V.tau[i]) * twothird * V.h_np1[i];
    double lk = one;
    if (std::fabs(hk_n) > yld_tol) lk = hk_n1 / hk_n;
    double bar_alpha_n[6], rtse[6], rsen[6];
    for (int k = 0; k < 6; ++k) bar_alpha_n[k] = lk * alpha_n[i][k];
    for (int k = 0; k < 6; ++k) rtse[k] = B.trial[i][k] - bar_alpha_n[k];
    const double norm_rtse = std::sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                                       two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
    const double stress_n_mean = (sn[0] + sn[1] + sn[2]) / three;
    rsen[0] = sn[0] - stress_n_mean - alpha_n[i][0];
    rsen[1] = sn[1] - stress_n_mean - alpha_n[i][1];
    rsen[2] = sn[2] - stress_n_mean - alpha_n[i][2];
    rsen[3] = sn[3] - alpha_n[i][3];
    rsen[4] = sn[4] - alpha_n[i][4];
    rsen[5] = sn[5] - alpha_n[i][5];
    const double norm_rsen = std::sqrt(rsen[0] * rsen[0] + rsen[1] * rsen[1] + rsen[2] * rsen[2] +
                                       two * (rsen[3] * rsen[3] + rsen[4] * rsen[4] + rsen[5] * rsen[5]));
    const double k_n = B.history_n[i][1];
    const double hi_n1 = V.tau[i] * (twothird * V.h_np1[i]);
    const double eps_n = B.history_n[i][2];
    const double kb_n1 = root2third * V.sigyld_np1[i] + hi_n1 * eps_n;
    const double f_n = norm_rsen - k_n;
    const double f_n1 = norm_rtse - kb_n1;
    const double F_trial = f_n1 - f_n;
    instat[i] = 1;
    yield[i] = true;
    if ((f_n1 < yld_tol * k_n) || (F_trial < yld_tol * k_n)) {
      yield[i] = false;
      instat[i] = 3;
    }
    if (iter > 0) continue;
    if (prior_linear[i]) {
      yield[i] = false;
      instat[i] = 3;
      continue;
    }
  }
}

// mm05_gp_felas_eta
inline void felas_eta(double eta, double &felas, double g_n, double g_n1, double Hi_n, double Hi_n1, double Hk_n,
                      double Hk_n1, double k0_n, double k0_n1, double t1, double t2, double t3, double t4, double t5,
                      double t6, double eps_n) {
  const double one = W3DO5F_K(1.0), four = W3DO5F_K(4.0), tol = W3DO5F_K(0.00000001);
  const double g_b = eta * g_n1 + (one - eta) * g_n;
  const double Hi_b = eta * Hi_n1 + (one - eta) * Hi_n;
  const double Hk_b = eta * Hk_n1 + (one - eta) * Hk_n;
  const double k0_b = eta * k0_n1 + (one - eta) * k0_n;
  const double c1 = g_b / g_n;
  double c2 = one;
  if (std::fabs(Hk_n) > tol) c2 = Hk_b / Hk_n;
  const double a1 = (c1 * c1) * t1 - 2 * c1 * c2 * t2 + (c2 * c2) * t3;
  const double a2 = four * g_b * eta * (c1 * t4 - c2 * t5);
  const double a3 = four * (g_b * g_b) * (eta * eta) * t6;
  const double a4 = k0_b + Hi_b * eps_n;
  felas = std::sqrt(a1 + a2 + a3) - a4;
}

// mm05_gp_elastic_fraction.  rc: OK, CUT or ABORT_*
inline int elastic_fraction(double *stress_n, double *alpha_n, double *devdeps, double &eta, double k_n, double g_n,
                            double g_n1, double &g_s, double Hi_n, double Hi_n1, double &Hi_s, double Hk_n,
                            double Hk_n1, double &Hk_s, double beta_n, double beta_n1, double &beta_s, double delta_n,
                            double delta_n1, double &delta_s, double k0_n, double k0_n1, double &k0_s, double eps_n,
                            bool adaptive_possible) {
  const double zero = W3DO5F_K(0.0), half = W3DO5F_K(0.5), one = W3DO5F_K(1.0), two = W3DO5F_K(2.0),
               tol = W3DO5F_K(1.0e-10), stol = W3DO5F_K(1.0e-10), eta_comp = W3DO5F_K(1.0e-10),
               unused = -W3DO5F_K(1.11e30), ptol = W3DO5F_K(0.0001), c = W3DO5F_K(0.6180339887);
  const double t1 = stress_n[0] * stress_n[0] + stress_n[1] * stress_n[1] + stress_n[2] * stress_n[2] +
                    two * (stress_n[3] * stress_n[3] + stress_n[4] * stress_n[4] + stress_n[5] * stress_n[5]);
  const double t2 = stress_n[0] * alpha_n[0] + stress_n[1] * alpha_n[1] + stress_n[2] * alpha_n[2] +
                    two * (stress_n[3] * alpha_n[3] + stress_n[4] * alpha_n[4] + stress_n[5] * alpha_n[5]);
  const double t3 = alpha_n[0] * alpha_n[0] + alpha_n[1] * alpha_n[1] + alpha_n[2] * alpha_n[2] +
                    two * (alpha_n[3] * alpha_n[3] + alpha_n[4] * alpha_n[4] + alpha_n[5] * alpha_n[5]);
  const double t4 = devdeps[0] * stress_n[0] + devdeps[1] * stress_n[1] + devdeps[2] * stress_n[2] +
                    devdeps[3] * stress_n[3] + devdeps[4] * stress_n[4] + devdeps[5] * stress_n[5];
  const double t5 = devdeps[0] * alpha_n[0] + devdeps[1] * alpha_n[1] + devdeps[2] * alpha_n[2] +
                    devdeps[3] * alpha_n[3] + devdeps[4] * alpha_n[4] + devdeps[5] * alpha_n[5];
  const double t6 = devdeps[0] * devdeps[0] + devdeps[1] * devdeps[1] + devdeps[2] * devdeps[2] +
                    half * (devdeps[3] * devdeps[3] + devdeps[4] * devdeps[4] + devdeps[5] * devdeps[5]);
  auto f = [&](double x, double &fx) {
    felas_eta(x, fx, g_n, g_n1, Hi_n, Hi_n1, Hk_n, Hk_n1, k0_n, k0_n1, t1, t2, t3, t4, t5, t6, eps_n);
  };
  double eta_low = zero, eta_high = one, f_low, f_high;
  f(eta_low, f_low);
  f(eta_high, f_high);
  bool found = false;
  int branch = 0;                                   // 100 = Ridder, 200 = golden section
  if (f_low < zero && f_high > zero) branch = 100;
  if (branch == 0 && f_low > zero && f_high > zero) {
    const double eta_pt = eta_low + ptol;
    double f_pt;
    f(eta_pt, f_pt);
    const double df = (f_pt - f_low) / ptol;
    if (df > zero) {
      eta = eta_low;
      found = true;
    } else {
      branch = 200;
    }
  }
  if (branch == 0 && !found) {
    if (std::fabs(f_low) <= stol * k_n) {
      eta = eta_low;
      found = true;
    } else if (std::fabs(f_high) <= stol * k_n) {
      return ABORT_ETA_ONE;
    } else {
      return ABORT_ETA_BRACKET;
    }
  }
  if (branch == 100) {
    double eta_ridder = unused;
    int loop;
    for (loop = 1; loop <= 50; ++loop) {
      const double eta_mid = half * (eta_low + eta_high);
      double f_mid;
      f(eta_mid, f_mid);
      const double sridder = std::sqrt(f_mid * f_mid - f_low * f_high);
      if (sridder == zero) return adaptive_possible ? CUT : ABORT_RIDDER;
      const double eta_new = eta_mid + (eta_mid - eta_low) * (fsign(one, f_low - f_high) * f_mid / sridder);
      if (std::fabs(eta_new - eta_ridder) <= eta_comp) {
        eta = eta_new;
        found = true;
        break;
      }
      eta_ridder = eta_new;
      double f_new;
      f(eta_ridder, f_new);
      if (std::fabs(f_new) <= stol * k_n) {
        eta = eta_new;
        found = true;
        break;
      }
      if (fsign(f_mid, f_new) != f_mid) {
        eta_low = eta_mid;
        f_low = f_mid;
        eta_high = eta_ridder;
        f_high = f_new;
      } else if (fsign(f_low, f_new) != f_low) {
        eta_high = eta_ridder;
        f_high = f_new;
      } else if (fsign(f_high, f_new) != f_high) {
        eta_low = eta_ridder;
        f_low = f_new;
      } else {
        return CUT;                                 // cut_step_size_now = .true. whatever adaptive_possible says
      }
      if (std::fabs(eta_high - eta_low) <= eta_comp) {
        eta = (eta_high + eta_low) * half;
        found = true;
        break;
      }
    }
    if (!found) return adaptive_possible ? CUT : ABORT_RIDDER;
  }
  if (branch == 200) {
    double a = eta_low, b = eta_high;
    double x1 = c * a + (one - c) * b, fx1;
    f(x1, fx1);
    double x2 = (one - c) * a + c * b, fx2;
    f(x2, fx2);
    int loop;
    for (loop = 1; loop <= 10000; ++loop) {
      if (fx1 < fx2) {
        b = x2;
        x2 = x1;
        fx2 = fx1;
        x1 = c * a + (one - c) * b;
        f(x1, fx1);
      } else {
        a = x1;
        x1 = x2;
        fx1 = fx2;
        x2 = (one - c) * a + c * b;
        f(x2, fx2);
      }
      if (std::fabs(b - a) < eta_comp) {
        eta = half * (a + b);
        found = true;
        break;
      }
    }
    if (!found) return adaptive_possible ? CUT : ABORT_GOLDEN;
  }
  // 1000
  g_s = eta * g_n1 + (one - eta) * g_n;
  Hi_s = eta * Hi_n1 + (one - eta) * Hi_n;
  Hk_s = eta * Hk_n1 + (one - eta) * Hk_n;
  k0_s = eta * k0_n1 + (one - eta) * k0_n;
  beta_s = eta * beta_n1 + (one - eta) * beta_n;
  delta_s = eta * delta_n1 + (one - eta) * delta_n;
  stress_n[0] = g_s / g_n * stress_n[0] + eta * 2 * g_s * devdeps[0];
  stress_n[1] = g_s / g_n * stress_n[1] + eta * 2 * g_s * devdeps[1];
  stress_n[2] = g_s / g_n * stress_n[2] + eta * 2 * g_s * devdeps[2];
  stress_n[3] = g_s / g_n * stress_n[3] + eta * g_s * devdeps[3];
  stress_n[4] = g_s / g_n * stress_n[4] + eta * g_s * devdeps[4];
  stress_n[5] = g_s / g_n * stress_n[5] + eta * g_s * devdeps[5];
  double lk = one;
  if (std::fabs(Hk_n) > tol) lk = Hk_s / Hk_n;
  for (int k = 0; k < 6; ++k) alpha_n[k] = lk * alpha_n[k];
  for (int k = 0; k < 6; ++k) devdeps[k] = (1 - eta) * devdeps[k];
  return OK;
}

// mm05_gp_lambda_solve
inline int lambda_solve(double norm_equiv_t, double norm_equiv_j, double k_j, double g_n1, double delta_n1,
                        double beta_n1, double Hk_n1, double Hi_n1, double dHi, double k0_n1, bool adaptive_possible,
                        double eps_j, double &lambda) {
  const double zero = W3DO5F_K(0.0), half = W3DO5F_K(0.5), two = W3DO5F_K(2.0);
  const double H_n1 = Hi_n1 + Hk_n1;
  const double f_j = norm_equiv_j - k_j;
  const double kb_j1 = k0_n1 + Hi_n1 * eps_j;
  const double g1 = half * (two * g_n1 + H_n1);
  const double a1 = norm_equiv_t - kb_j1;
  const double a2 = a1 - f_j;
  const double a3 = delta_n1 - (two * g_n1 + dHi);
  const double a4 = beta_n1 * (delta_n1 + H_n1);
  const double a = two * g1 * a3;
  const double b = two * g1 * a2 + a4 - a1 * a3;
  const double c = -a1 * a2;
  double temp = b * b - 4 * a * c;
  if (temp < zero) return adaptive_possible ? CUT : ABORT_LAMBDA_COMPLEX;
  const double l1 = half / a * (-b - std::sqrt(temp));
  const double l2 = half / a * (-b + std::sqrt(temp));
  if (l1 < zero && l2 < zero) return adaptive_possible ? CUT : ABORT_LAMBDA_NEGATIVE;
  if (l1 > zero) {
    lambda = l1;
    if (lambda > l2 && l2 > zero) lambda = l2;
  } else {
    lambda = l2;
  }
  temp = a * (lambda * lambda) + b * lambda + c;
  temp = std::fabs(temp);
  if (temp > W3DO5F_K(0.000001)) return adaptive_possible ? CUT : ABORT_LAMBDA_QUADRATIC;
  return OK;
}

// mm05_gp_nsteps
inline int nsteps(int nsubinc, const double *devdeps, const double *dev_stress_s, const double *alpha_s, double eps_s,
                  double g_s, double Hi_s, double Hk_s, double beta_s, double delta_s, double k0_s, double g_n1,
                  double Hi_n1, double Hk_n1, double beta_n1, double delta_n1, double k0_n1, double *stress_j,
                  double *alpha_j, double &eps_j, double &lambda, double &dHi, double *sig_t, double &norm_equiv_j,
                  bool adaptive_possible) {
  const double one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), tol = W3DO5F_K(0.00000001);
  const double dnsubinc = (double)nsubinc;
  const double d = one / dnsubinc;
  const double dHk = d * (Hk_n1 - Hk_s);
  dHi = d * (Hi_n1 - Hi_s);
  const double dg = d * (g_n1 - g_s);
  const double dk0 = d * (k0_n1 - k0_s);
  const double dbeta = d * (beta_n1 - beta_s);
  const double ddel = d * (delta_n1 - delta_s);
  double Hk_j = Hk_s, Hi_j = Hi_s, g_j = g_s, k0_j = k0_s, beta_j = beta_s, delta_j = delta_s;
  for (int k = 0; k < 6; ++k) { stress_j[k] = dev_stress_s[k]; alpha_j[k] = alpha_s[k]; }
  eps_j = eps_s;
  for (int j = 1; j <= nsubinc; ++j) {
    const double Hk_j1 = Hk_j + dHk, Hi_j1 = Hi_j + dHi, g_j1 = g_j + dg, k0_j1 = k0_j + dk0,
                 beta_j1 = beta_j + dbeta, delta_j1 = delta_j + ddel;
    const double mk = g_j1 / g_j;
    double lk = one;
    if (std::fabs(Hk_j) > tol) lk = Hk_j1 / Hk_j;
    const double k_j = k0_j + Hi_j * eps_j;
    sig_t[0] = mk * stress_j[0] + two * g_j1 * devdeps[0] * d;
    sig_t[1] = mk * stress_j[1] + two * g_j1 * devdeps[1] * d;
    sig_t[2] = mk * stress_j[2] + two * g_j1 * devdeps[2] * d;
    sig_t[3] = mk * stress_j[3] + g_j1 * devdeps[3] * d;
    sig_t[4] = mk * stress_j[4] + g_j1 * devdeps[4] * d;
    sig_t[5] = mk * stress_j[5] + g_j1 * devdeps[5] * d;
    double alp_t[6], equiv_t[6];
    for (int k = 0; k < 6; ++k) alp_t[k] = lk * alpha_j[k];
    for (int k = 0; k < 6; ++k) equiv_t[k] = sig_t[k] - alp_t[k];
    const double norm_equiv_t = std::sqrt(equiv_t[0] * equiv_t[0] + equiv_t[1] * equiv_t[1] + equiv_t[2] * equiv_t[2] +
                                          two * (equiv_t[3] * equiv_t[3] + equiv_t[4] * equiv_t[4] + equiv_t[5] * equiv_t[5]));
    norm_equiv_j = std::sqrt((stress_j[0] - alpha_j[0]) * (stress_j[0] - alpha_j[0]) +
                             (stress_j[1] - alpha_j[1]) * (stress_j[1] - alpha_j[1]) +
                             (stress_j[2] - alpha_j[2]) * (stress_j[2] - alpha_j[2]) +
                             two * ((stress_j[3] - alpha_j[3]) * (stress_j[3] - alpha_j[3])) +
                             two * ((stress_j[4] - alpha_j[4]) * (stress_j[4] - alpha_j[4])) +
                             two * ((stress_j[5] - alpha_j[5]) * (stress_j[5] - alpha_j[5])));
    const int rc = lambda_solve(norm_equiv_t, norm_equiv_j, k_j, g_j1, delta_j1, beta_j1, Hk_j1, Hi_j1, dHi, k0_j1,
                                adaptive_possible, eps_j, lambda);
    if (rc != OK) return rc;
    const double eps_j1 = eps_j + lambda;
    double stress_j1[6], alpha_j1[6];
    for (int k = 0; k < 6; ++k) stress_j1[k] = sig_t[k] - 2 * g_j1 * lambda * equiv_t[k] / norm_equiv_t;
    for (int k = 0; k < 6; ++k) alpha_j1[k] = alp_t[k] + Hk_j1 * lambda * equiv_t[k] / norm_equiv_t;
    for (int k = 0; k < 6; ++k) { alpha_j[k] = alpha_j1[k]; stress_j[k] = stress_j1[k]; }
    eps_j = eps_j1;
    Hk_j = Hk_j1; Hi_j = Hi_j1; g_j = g_j1; k0_j = k0_j1; beta_j = beta_j1; delta_j = delta_j1;
  }
  return OK;
}

// mm05_gp_compute
inline int compute(Block &B, const GpVectors &V, const bool *yield, const double *g_vec_np1, const double *g_vec_n,
                   bool adaptive_possible) {
  const double one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), three = W3DO5F_K(3.0), twothird = W3DO5F_K(0.66666666),
               root2third = W3DO5F_K(0.81649658);
  const int mxsubinc = 10;
  for (int i = 0; i < B.span; ++i) {
    if (!yield[i]) continue;
    const double g_n = g_vec_n[i], g_n1 = g_vec_np1[i];
    const double Hi_n = V.tau[i] * twothird * V.h_n[i], Hi_n1 = V.tau[i] * twothird * V.h_np1[i];
    const double Hk_n = (one - V.tau[i]) * twothird * V.h_n[i], Hk_n1 = (one - V.tau[i]) * twothird * V.h_np1[i];
    const double beta_n = root2third * V.beta_n[i], beta_n1 = root2third * V.beta_np1[i];
    const double delta_n = twothird * V.delta_n[i], delta_n1 = twothird * V.delta_np1[i];
    const double k0_n = root2third * V.sigyld_n[i], k0_n1 = root2third * V.sigyld_np1[i];
    const double sig_tol = one * B.mm_props[i][5];
    const double k_n = B.history_n[i][1];
    const double eps_n = B.history_n[i][2];
    const double rel_tol = sig_tol * k_n;
    const double *sn = B.stress_n[i];
    const double mean_stress = (sn[0] + sn[1] + sn[2]) / three;
    double dev_stress_n[6] = {sn[0] - mean_stress, sn[1] - mean_stress, sn[2] - mean_stress, sn[3], sn[4], sn[5]};
    double alpha_n[6];
    for (int k = 0; k < 6; ++k) alpha_n[k] = B.history_n[i][5 + k];
    const double *dp = B.deps[i];
    const double mean_deps = (dp[0] + dp[1] + dp[2]) / three;
    double devdeps[6] = {dp[0] - mean_deps, dp[1] - mean_deps, dp[2] - mean_deps, dp[3], dp[4], dp[5]};
    double eta, g_s, Hi_s, Hk_s, beta_s, delta_s, k0_s;
    int rc = elastic_fraction(dev_stress_n, alpha_n, devdeps, eta, k_n, g_n, g_n1, g_s, Hi_n, Hi_n1, Hi_s, Hk_n, Hk_n1,
                              Hk_s, beta_n, beta_n1, beta_s, delta_n, delta_n1, delta_s, k0_n, k0_n1, k0_s, eps_n,
                              adaptive_possible);
    if (rc != OK) return rc;
    double stress_1step[6], alpha_1step[6], eps_1step, lambda_1step, dHis, sig_t[6], norm_equiv_sm1;
    double stress_2step[6], alpha_2step[6], eps_2step, lambda_2step;
    double stress_nstep[6], alpha_nstep[6], eps_nstep, lambda_nstep;
    rc = nsteps(1, devdeps, dev_stress_n, alpha_n, eps_n, g_s, Hi_s, Hk_s, beta_s, delta_s, k0_s, g_n1, Hi_n1, Hk_n1,
                beta_n1, delta_n1, k0_n1, stress_1step, alpha_1step, eps_1step, lambda_1step, dHis, sig_t,
                norm_equiv_sm1, adaptive_possible);
    if (rc != OK) return rc;
    rc = nsteps(2, devdeps, dev_stress_n, alpha_n, eps_n, g_s, Hi_s, Hk_s, beta_s, delta_s, k0_s, g_n1, Hi_n1, Hk_n1,
                beta_n1, delta_n1, k0_n1, stress_2step, alpha_2step, eps_2step, lambda_2step, dHis, sig_t,
                norm_equiv_sm1, adaptive_possible);
    if (rc != OK) return rc;
    double error = 0.0;
    for (int k = 0; k < 6; ++k) {
      const double diff = std::fabs(stress_2step[k] - stress_1step[k]);
      if (k == 0 || diff > error) error = diff;
    }
    int nsubinc = (int)std::ceil(W3DO5F_K(1.1) * error / rel_tol);
    if (nsubinc < two) nsubinc = (int)two;
    if (nsubinc > mxsubinc) nsubinc = mxsubinc;
    B.history_np1[i][11] = nsubinc * one;
    double dev_stress_np1[6], alpha_np1[6], eps_np1;
    if (nsubinc < three) {
      for (int k = 0; k < 6; ++k) dev_stress_np1[k] = two * stress_2step[k] - stress_1step[k];
      for (int k = 0; k < 6; ++k) alpha_np1[k] = two * alpha_2step[k] - alpha_1step[k];
      eps_np1 = two * eps_2step - eps_1step;
      lambda_nstep = lambda_2step;
    } else {
      rc = nsteps(nsubinc, devdeps, dev_stress_n, alpha_n, eps_n, g_s, Hi_s, Hk_s, beta_s, delta_s, k0_s, g_n1, Hi_n1,
                  Hk_n1, beta_n1, delta_n1, k0_n1, stress_nstep, alpha_nstep, eps_nstep, lambda_nstep, dHis, sig_t,
                  norm_equiv_sm1, adaptive_possible);
      if (rc != OK) return rc;
      const double a = (one * nsubinc) / (nsubinc - two);
      const double b = (one * two) / (nsubinc - two);
      for (int k = 0; k < 6; ++k) dev_stress_np1[k] = a * stress_nstep[k] - b * stress_2step[k];
      for (int k = 0; k < 6; ++k) alpha_np1[k] = a * alpha_nstep[k] - b * alpha_2step[k];
      eps_np1 = a * eps_nstep - b * eps_2step;
    }
    for (int k = 0; k < 6; ++k) B.stress_np1[i][k] = dev_stress_np1[k];
    B.history_np1[i][0] = lambda_nstep;
    B.history_np1[i][1] = k0_n1 + Hi_n1 * eps_np1;
    B.history_np1[i][2] = eps_np1;
    B.history_np1[i][4] = norm_equiv_sm1;
    for (int k = 0; k < 6; ++k) B.history_np1[i][5 + k] = alpha_np1[k];
    B.history_np1[i][12] = k_n;
    B.history_np1[i][13] = dHis;
    for (int k = 0; k < 6; ++k) B.trial[i][k] = sig_t[k];
  }
  return OK;
}

// mm05_gp.  rc: OK, CUT or ABORT_*
inline int mm05_gp(int step, int iter, Block &B, const GpVectors &V, bool adaptive_possible) {
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0), htol = W3DO5F_K(0.000001), two = W3DO5F_K(2.0),
               twothird = W3DO5F_K(0.66666666), root2third = W3DO5F_K(0.81649658);
  const int span = B.span;
  if (span > 256) return ABORT_GOLDEN;
  bool yield[256], prior_linear[256];
  int iostat[256], instat[256];
  double g_vec_n[256], g_vec_np1[256], alpha_n[256][6], trace_eps_np1[256];
  if (step == 1) step1_set(B, V);
  init(B, V, iter, iostat, prior_linear, g_vec_np1, g_vec_n, alpha_n, instat, yield, trace_eps_np1);
  const int rc = compute(B, V, yield, g_vec_np1, g_vec_n, adaptive_possible);
  if (rc != OK) return rc;
  for (int i = 0; i < span; ++i) {
    if (yield[i]) continue;
    const double eps_n = B.history_n[i][2];
    const double hi_n = twothird * V.tau[i] * V.h_n[i];
    const double hi_n1 = twothird * V.tau[i] * V.h_np1[i];
    const double hk_n = twothird * (one - V.tau[i]) * V.h_n[i];
    const double hk_n1 = twothird * (one - V.tau[i]) * V.h_np1[i];
    double lk = one;
    if (std::fabs(hk_n) > htol) lk = hk_n1 / hk_n;
    B.history_np1[i][0] = zero;
    B.history_np1[i][1] = root2third * V.sigyld_np1[i] + hi_n1 * eps_n;
    B.history_np1[i][2] = B.history_n[i][2];
    for (int k = 5; k <= 10; ++k) B.history_np1[i][k] = lk * B.history_n[i][k];
    B.history_np1[i][11] = zero;
    B.history_np1[i][12] = root2third * V.sigyld_n[i] + hi_n * eps_n;
    B.history_np1[i][13] = hi_n1 - hi_n;
    for (int k = 0; k < 6; ++k) B.stress_np1[i][k] = B.trial[i][k];
    double rse[6];
    for (int k = 0; k < 6; ++k) rse[k] = B.stress_np1[i][k] - B.history_np1[i][5 + k];
    B.history_np1[i][4] = std::sqrt(rse[0] * rse[0] + rse[1] * rse[1] + rse[2] * rse[2] +
                                    two * (rse[3] * rse[3] + rse[4] * rse[4] + rse[5] * rse[5]));
  }
  // mm05_final / mm05_plastic_work with the n+1 moduli
  Block F = B;
  F.e_vec = V.e_np1;
  F.nu_vec = V.nu_np1;
  w3do_mm05_fa::final_(F, g_vec_np1, instat, trace_eps_np1);
  if (iter > 0) w3do_mm05_fa::plastic_work(F, yield, g_vec_np1);
  return OK;
}

// cnst5_gp
inline void cnst5_gp(const Block &B, const GpVectors &V, double (*dmat)[6][6]) {
  const double zero = 0.0, one = 1.0, two = 2.0, twothird = 0.66666666, onethird = 0.33333333, half = 0.5,
               rt2third = 0.81649658, three = 3.0;
  for (int i = 0; i < B.span; ++i) {
    int32_t iw[2];
    std::memcpy(iw, &B.history_np1[i][3], 8);
    const bool yield = iw[0] == 1;
    double(*D)[6] = dmat[i];
    if (!yield) {
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) D[r][c] = zero;
      const double c1 = V.e_np1[i] / ((one + V.nu_np1[i]) * (one - two * V.nu_np1[i]));
      const double c2 = (one - V.nu_np1[i]) * c1;
      const double c3 = ((one - two * V.nu_np1[i]) / two) * c1;
      const double c4 = V.nu_np1[i] * c1;
      D[0][0] = c2; D[1][1] = c2; D[2][2] = c2;
      D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
      D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
      continue;
    }
    const double g = V.e_np1[i] / (two * (one + V.nu_np1[i]));
    const double bulk = V.e_np1[i] / (three * (one - two * V.nu_np1[i]));
    const double del_n1 = twothird * V.delta_np1[i];
    const double beta_n1 = rt2third * V.beta_np1[i];
    const double H_n1 = twothird * V.h_np1[i];
    const double Hk_n1 = (one - V.tau[i]) * H_n1;
    const double G1 = g + half * H_n1;
    const double lambda = B.history_np1[i][0];
    const double k_n1 = B.history_np1[i][1];
    const double dHi = B.history_np1[i][13];
    const double norm_equiv_n = B.history_np1[i][4];
    const double k_n = B.history_np1[i][12];
    const double *s1 = B.stress_np1[i];
    const double mean_sig = (s1[0] + s1[1] + s1[2]) / three;
    double equiv[6];
    equiv[0] = s1[0] - B.history_np1[i][5] - mean_sig;
    equiv[1] = s1[1] - B.history_np1[i][6] - mean_sig;
    equiv[2] = s1[2] - B.history_np1[i][7] - mean_sig;
    equiv[3] = s1[3] - B.history_np1[i][8];
    equiv[4] = s1[4] - B.history_np1[i][9];
    equiv[5] = s1[5] - B.history_np1[i][10];
    const double norm_equiv_np1 = std::sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                            two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
    const double f_n = norm_equiv_n - k_n;
    const double f_n1 = norm_equiv_np1 - k_n1;
    const double BB1 = H_n1 + del_n1 - dHi;
    const double BB2 = (del_n1 + H_n1) * beta_n1;
    const double BB3 = two * f_n1 - f_n + BB1 * lambda;
    const double BB4 = BB2 - BB1 * f_n1;
    const double A_GP = two * g * BB3 / (two * G1 * BB3 + BB4);
    const double C_GP = two * g * lambda / (norm_equiv_np1 + (two * g + Hk_n1) * lambda);
    const double c1 = two * g * (1 - C_GP);
    const double c2 = two * g * (C_GP - A_GP);
    double n[6];
    for (int q = 0; q < 6; ++q) n[q] = equiv[q] / norm_equiv_np1;
    D[0][0] = bulk + twothird * c1 + c2 * n[0] * n[0];
    D[1][1] = bulk + twothird * c1 + c2 * n[1] * n[1];
    D[2][2] = bulk + twothird * c1 + c2 * n[2] * n[2];
    D[3][3] = half * c1 + c2 * n[3] * n[3];
    D[4][4] = half * c1 + c2 * n[4] * n[4];
    D[5][5] = half * c1 + c2 * n[5] * n[5];
    D[0][1] = (bulk - onethird * c1 + c2 * n[0] * n[1]);
    D[0][2] = (bulk - onethird * c1 + c2 * n[0] * n[2]);
    D[0][3] = (c2 * n[0] * n[3]);
    D[0][4] = (c2 * n[0] * n[4]);
    D[0][5] = (c2 * n[0] * n[5]);
    D[1][2] = bulk - onethird * c1 + c2 * n[1] * n[2];
    D[1][3] = c2 * n[1] * n[3];
    D[1][4] = c2 * n[1] * n[4];
    D[1][5] = c2 * n[1] * n[5];
    D[2][3] = c2 * n[2] * n[3];
    D[2][4] = c2 * n[2] * n[4];
    D[2][5] = c2 * n[2] * n[5];
    D[3][4] = c2 * n[3] * n[4];
    D[3][5] = c2 * n[3] * n[5];
    D[4][5] = c2 * n[4] * n[5];
    for (int r = 0; r < 6; ++r)
      for (int cc = 0; cc < r; ++cc) D[r][cc] = D[cc][r];
  }
}

}  // namespace w3do_mm05_gp
#endif
