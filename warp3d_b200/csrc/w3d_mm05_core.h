// w3d_mm05_core.h -- cyclic plasticity model mm05, "nonlinear_hardening" (Frederick-Armstrong)
// and "generalized_plasticity" options, for ONE Gauss point: the engine's device core.  The
// parity tests check it against an independent per-point CPU restatement of mm05.f kept with the
// test infrastructure; no file is shared with it (the scalar Newton solve for the plastic
// multiplier and the Richardson sub-increment count are tolerance / threshold driven, so both
// must make the same decisions -- asserted by the 1e-10 parity and integer-state tests).
//
// Follows reference src/mm05.f: mm05 :9, mm05_fa :288, mm05_fa_step1_set :424, mm05_fa_init :490,
// mm05_fa_compute :648, mm05_fa_nsteps :883, mm05_fa_lambda_solve :981, mm05_find_pos_lambda
// :1120, mm05_compute_residual :1269, mm05_update_lambda :1323, mm05_fa_elastic_fraction :1367,
// cnst5 :1566, mm05_final :3879, mm05_plastic_work :3949; properties src/rknstr.f:1234-1317,
// src/inmat.f:2003-2040.  The generalized_plasticity option (mm05_gp :1882 ... cnst5_gp :3584)
// follows the nonlinear_hardening one below.
//
// Parity quirks kept: the 8-digit DATA constants 0.81649658 / 0.66666666 / 0.33333333, the typo
// 0.8164956809 (for sqrt(2/3) = 0.8164965809) in the local solve, tol = 1e-6 if iter > 1 else
// 1e-3, max 15 Newton passes, integer state in the low word of history(4).  The reference is
// built with -fpconstant, so DATA literals without a d-exponent are taken as evaluated in double
// (define W3D_MM05_DATA_SINGLE to evaluate them in single instead).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define W3D5_HD __host__ __device__
#else
#define W3D5_HD
#endif

namespace w3d_mm05 {

#ifdef W3D_MM05_DATA_SINGLE
#define W3D5_K(x) ((double)(x##f))
#else
#define W3D5_K(x) (x)
#endif

W3D5_HD inline double state_dword(int st) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(0, st);
#else
  int iw[2] = {st, 0};
  double d;
  memcpy(&d, iw, 8);
  return d;
#endif
}
W3D5_HD inline int dword_state(double d) {
#if defined(__CUDA_ARCH__)
  return __double2loint(d);
#else
  int iw[2];
  memcpy(iw, &d, 8);
  return iw[0];
#endif
}

struct Props {
  double e, nu, sigyld;   // props(7), props(8), props(23) (real*4 widened)
  double mm[10];          // matprp(55:64): q_u, b, h, option, gamma, sig_tol, -
};

enum { OK = 0, CUT = 1, FATAL_LAMBDA = 7, FATAL_FRACTION = 8 };

// mm05_compute_residual (mm05.f:1269)
W3D5_HD inline void compute_residual(double lambda, double aps_n, const double *sig_t,
                                     const double *alpha_n, double gamma, double sigy, double H_bar,
                                     double Q_bar, double b, double g, double &k_np1, double &H_np1,
                                     double &lg, double &res, double *equiv, double &aps_np1,
                                     double &dk_np1, double &dH_np1) {
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0), two = W3D5_K(2.0), root2third = W3D5_K(0.8164956809);
  aps_np1 = aps_n + lambda;
  k_np1 = root2third * (sigy + Q_bar * (one - exp(-b * aps_np1)));
  dk_np1 = root2third * b * Q_bar * (exp(-b * aps_np1));
  H_np1 = H_bar;
  dH_np1 = zero;
  lg = one + lambda * gamma;
  for (int k = 0; k < 6; ++k) equiv[k] = lg * sig_t[k] - alpha_n[k];
  const double norm_equiv = sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                 two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
  res = lg * (k_np1 + two * g * lambda) + lambda * H_np1 - norm_equiv;
}
// mm05_update_lambda (mm05.f:1323)
W3D5_HD inline void update_lambda(double &lambda, double res, const double *equiv, double gamma,
                                  double k_np1, double g, double H_np1, double dk_np1, double dH_np1,
                                  const double *sig_t) {
  const double one = W3D5_K(1.0), two = W3D5_K(2.0);
  const double norm_equiv = sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                 two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
  const double eq_dot_sigt = equiv[0] * sig_t[0] + equiv[1] * sig_t[1] + equiv[2] * sig_t[2] +
                             two * equiv[3] * sig_t[3] + two * equiv[4] * sig_t[4] + two * equiv[5] * sig_t[5];
  const double dres = gamma * (k_np1 + two * g * lambda) + (one + lambda * gamma) * (dk_np1 + two * g) +
                      H_np1 + lambda * dH_np1 - gamma * eq_dot_sigt / norm_equiv;
  lambda = lambda - res / dres;
}
W3D5_HD inline double sgn1(double x) { return copysign(1.0, x); }

// mm05_find_pos_lambda (mm05.f:1120): bracket + bisection with Newton polishing.
// returns OK, CUT or FATAL_LAMBDA
W3D5_HD inline int find_pos_lambda(double sigy, const double *sig_t, double *equiv, double aps_n,
                                   double gamma, double &lambda, double &k_np1, double Q, double b,
                                   double H_bar, double &lg, double g, const double *alpha_n,
                                   bool adaptive_possible, double &H_np1, double tol, double &aps_np1) {
  const double zero = W3D5_K(0.0);
  const int sc_max = 4, bs_max = 10, nr_max = 10;
  double dk_np1, dH_np1, res_i, res_f, res;
  double lambda_i = zero;
  int sc_count = 0, bs_count = 0, nr_count = 0;
  compute_residual(lambda_i, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_i,
                   equiv, aps_np1, dk_np1, dH_np1);
  double lambda_f = -1 * lambda;
  compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_f,
                   equiv, aps_np1, dk_np1, dH_np1);
  while (sgn1(res_i) * sgn1(res_f) > zero) {
    sc_count = sc_count + 1;
    if (sc_count > sc_max) return adaptive_possible ? CUT : FATAL_LAMBDA;
    lambda_f = 2 * lambda_f;
    compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_f,
                     equiv, aps_np1, dk_np1, dH_np1);
  }
  while (bs_count < bs_max) {
    bs_count = bs_count + 1;
    lambda = lambda_i + 0.5 * (lambda_f - lambda_i);
    while (nr_count < nr_max) {
      compute_residual(lambda, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res,
                       equiv, aps_np1, dk_np1, dH_np1);
      if (fabs(res) < tol * sigy) return OK;
      nr_count = nr_count + 1;
      update_lambda(lambda, res, equiv, gamma, k_np1, g, H_np1, dk_np1, dH_np1, sig_t);
      if (lambda < lambda_i || lambda > lambda_f) break;
    }
    lambda = lambda_i + 0.5 * (lambda_f - lambda_i);
    compute_residual(lambda, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res,
                     equiv, aps_np1, dk_np1, dH_np1);
    if ((sgn1(res_i) * sgn1(res)) > zero) {
      lambda_i = lambda;
      compute_residual(lambda_i, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg,
                       res_i, equiv, aps_np1, dk_np1, dH_np1);
    } else {
      lambda_f = lambda;
      compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg,
                       res_f, equiv, aps_np1, dk_np1, dH_np1);
    }
    nr_count = 0;
  }
  return adaptive_possible ? CUT : FATAL_LAMBDA;
}

// mm05_fa_lambda_solve (mm05.f:981): scalar Newton for the plastic multiplier of one sub-step;
// aps_n is advanced to aps_np1 on success
W3D5_HD inline int lambda_solve(double sig_y, const double *alpha_s, const double *sig_t, double *equiv,
                                double &aps_n, double gamma, double &lambda, double &k_np1,
                                double Q_bar, double b, double H_bar, double &lg, double g,
                                bool adaptive_possible, double &H_np1, int iter) {
  const int max_iter = 15;
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0), two = W3D5_K(2.0), root2third = W3D5_K(0.8164956809);
  const double tol = (iter > 1) ? W3D5_K(0.000001) : W3D5_K(0.001);
  lambda = zero;
  for (int j = 1; j <= max_iter; ++j) {
    double aps_np1 = aps_n + lambda;
    k_np1 = root2third * (sig_y + Q_bar * (one - exp(-b * aps_np1)));
    const double dk_np1 = root2third * b * Q_bar * (exp(-b * aps_np1));
    H_np1 = H_bar;
    const double dH_np1 = zero;
    lg = one + lambda * gamma;
    for (int k = 0; k < 6; ++k) equiv[k] = lg * sig_t[k] - alpha_s[k];
    const double norm_equiv = sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                   two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
    const double res = lg * (k_np1 + two * g * lambda) + lambda * H_np1 - norm_equiv;
    if (fabs(res) < tol * sig_y) {
      if (lambda < zero) {
        const int rc = find_pos_lambda(sig_y, sig_t, equiv, aps_n, gamma, lambda, k_np1, Q_bar, b, H_bar, lg, g,
                                       alpha_s, adaptive_possible, H_np1, tol, aps_np1);
        if (rc != OK) return rc;
      }
      aps_n = aps_np1;
      return OK;
    }
    const double eq_dot_sigt = equiv[0] * sig_t[0] + equiv[1] * sig_t[1] + equiv[2] * sig_t[2] +
                               two * equiv[3] * sig_t[3] + two * equiv[4] * sig_t[4] + two * equiv[5] * sig_t[5];
    const double dres = gamma * (k_np1 + two * g * lambda) + lg * (dk_np1 + two * g) + H_np1 +
                        lambda * dH_np1 - gamma * eq_dot_sigt / norm_equiv;
    lambda = lambda - res / dres;
  }
  return adaptive_possible ? CUT : FATAL_LAMBDA;
}

// mm05_fa_nsteps (mm05.f:883): nsubinc equal sub-increments of the deviatoric strain increment
W3D5_HD inline int nsteps(int nsubinc, const double *devdeps, const double *dev_stress_n,
                          const double *alpha_n, double aps_n, double g, double Q_bar, double b,
                          double H_bar, double gamma, double sigyld, double *stress_j, double *alpha_j,
                          double &aps_j, double &lambda, double *sig_t, bool adaptive_possible, int iter) {
  const double two = W3D5_K(2.0), one = W3D5_K(1.0);
  for (int k = 0; k < 6; ++k) { stress_j[k] = dev_stress_n[k]; alpha_j[k] = alpha_n[k]; }
  aps_j = aps_n;
  const double d = one / nsubinc;
  for (int s = 1; s <= nsubinc; ++s) {
    sig_t[0] = stress_j[0] + two * g * devdeps[0] * d;
    sig_t[1] = stress_j[1] + two * g * devdeps[1] * d;
    sig_t[2] = stress_j[2] + two * g * devdeps[2] * d;
    sig_t[3] = stress_j[3] + g * devdeps[3] * d;
    sig_t[4] = stress_j[4] + g * devdeps[4] * d;
    sig_t[5] = stress_j[5] + g * devdeps[5] * d;
    double equiv[6], k_np1, lg, H_np1;
    const int rc = lambda_solve(sigyld, alpha_j, sig_t, equiv, aps_j, gamma, lambda, k_np1, Q_bar, b, H_bar,
                                lg, g, adaptive_possible, H_np1, iter);
    if (rc != OK) return rc;
    const double kappa = two * g * lambda / k_np1;
    const double beta = lambda * H_np1 / k_np1;
    const double ro = lg * (one + kappa) + beta;
    for (int k = 0; k < 6; ++k) stress_j[k] = sig_t[k] - kappa / ro * equiv[k];
    for (int k = 0; k < 6; ++k) alpha_j[k] = (alpha_j[k] + beta / ro * equiv[k]) / lg;
  }
  return OK;
}

// mm05_fa_elastic_fraction (mm05.f:1367): part of the increment that is still elastic
W3D5_HD inline int elastic_fraction(double g, double *stress_n, const double *alpha_n, double *devdeps,
                                    double &eta, double k_n, bool prior_linear) {
  const double zero = W3D5_K(0.0), two = W3D5_K(2.0), tol = W3D5_K(0.00001);
  eta = zero;
  const double a = two * g * ((stress_n[0] - alpha_n[0]) * devdeps[0] + (stress_n[1] - alpha_n[1]) * devdeps[1] +
                              (stress_n[2] - alpha_n[2]) * devdeps[2] + (stress_n[3] - alpha_n[3]) * devdeps[3] +
                              (stress_n[4] - alpha_n[4]) * devdeps[4] + (stress_n[5] - alpha_n[5]) * devdeps[5]);
  const double b = g * g * (4 * (devdeps[0] * devdeps[0] + devdeps[1] * devdeps[1] + devdeps[2] * devdeps[2]) +
                            2 * (devdeps[3] * devdeps[3] + devdeps[4] * devdeps[4] + devdeps[5] * devdeps[5]));
  double c = 0;
  if (prior_linear) {
    c = ((stress_n[0] - alpha_n[0]) * (stress_n[0] - alpha_n[0]) + (stress_n[1] - alpha_n[1]) * (stress_n[1] - alpha_n[1]) +
         (stress_n[2] - alpha_n[2]) * (stress_n[2] - alpha_n[2]) +
         two * ((stress_n[3] - alpha_n[3]) * (stress_n[3] - alpha_n[3]) + (stress_n[4] - alpha_n[4]) * (stress_n[4] - alpha_n[4]) +
                (stress_n[5] - alpha_n[5]) * (stress_n[5] - alpha_n[5])));
    const double temp = k_n - sqrt(c);
    c = k_n * k_n - c;
    if (c < zero) {
      if (temp < tol * k_n) c = zero;
      else return FATAL_FRACTION;
    }
  }
  if (c == zero) {
    if (a > zero) eta = zero;
    else eta = -two * a / b;
  } else {
    eta = -a / b + sqrt((a / b) * (a / b) + c / b);
    if (eta < zero) return FATAL_FRACTION;
  }
  stress_n[0] = stress_n[0] + eta * 2 * g * devdeps[0];
  stress_n[1] = stress_n[1] + eta * 2 * g * devdeps[1];
  stress_n[2] = stress_n[2] + eta * 2 * g * devdeps[2];
  stress_n[3] = stress_n[3] + eta * g * devdeps[3];
  stress_n[4] = stress_n[4] + eta * g * devdeps[4];
  stress_n[5] = stress_n[5] + eta * g * devdeps[5];
  for (int k = 0; k < 6; ++k) devdeps[k] = (1 - eta) * devdeps[k];
  return OK;
}

// mm05_final (mm05.f:3879) and mm05_plastic_work (:3949): shared by both options
W3D5_HD inline void final_and_work(const Props &P, const double *cgn, double *cgn1, const double *deps,
                                   double *h1, double trace_eps_np1, double shear_mod, int instat,
                                   bool yield, int iter) {
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0), two = W3D5_K(2.0), three = W3D5_K(3.0);
  const double e = P.e, nu = P.nu;
  // ---- mm05_final
  {
    const double half = W3D5_K(0.5);
    const double sig_mean_np1 = trace_eps_np1 * (three * e * nu / ((one + nu) * (one - two * nu)) + two * shear_mod) / three;
    cgn1[0] = cgn1[0] + sig_mean_np1; cgn1[1] = cgn1[1] + sig_mean_np1; cgn1[2] = cgn1[2] + sig_mean_np1;
    cgn1[6] = cgn[6] + half * (deps[0] * (cgn[0] + cgn1[0]) + deps[1] * (cgn[1] + cgn1[1]) + deps[2] * (cgn[2] + cgn1[2]) +
                               deps[3] * (cgn[3] + cgn1[3]) + deps[4] * (cgn[4] + cgn1[4]) + deps[5] * (cgn[5] + cgn1[5]));
    h1[3] = state_dword(instat);
  }
  // ---- mm05_plastic_work (only iter > 0; otherwise stress_np1(8:9) keep the local zero init)
  if (iter > 0) {
    cgn1[7] = cgn[7];
    cgn1[8] = cgn[8];
    if (yield) {
      const double half = 0.5, root2 = sqrt(2.0);
      double dsig[6], dp[6];
      for (int k = 0; k < 6; ++k) dsig[k] = cgn1[k] - cgn[k];
      dp[0] = deps[0] - (dsig[0] - nu * (dsig[1] + dsig[2])) / e;
      dp[1] = deps[1] - (dsig[1] - nu * (dsig[0] + dsig[2])) / e;
      dp[2] = deps[2] - (dsig[2] - nu * (dsig[1] + dsig[0])) / e;
      dp[3] = deps[3] - dsig[3] / shear_mod;
      dp[4] = deps[4] - dsig[4] / shear_mod;
      dp[5] = deps[5] - dsig[5] / shear_mod;
      cgn1[7] = cgn[7] + half * (dp[0] * (cgn1[0] + cgn[0]) + dp[1] * (cgn1[1] + cgn[1]) + dp[2] * (cgn1[2] + cgn[2]) +
                                 dp[3] * (cgn1[3] + cgn[3]) + dp[4] * (cgn1[4] + cgn[4]) + dp[5] * (cgn1[5] + cgn[5]));
      const double f1 = (dp[0] - dp[1]) * (dp[0] - dp[1]) + (dp[0] - dp[2]) * (dp[0] - dp[2]) + (dp[2] - dp[1]) * (dp[2] - dp[1]);
      const double f2 = dp[3] * dp[3] + dp[4] * dp[4] + dp[5] * dp[5];
      const double bar = (root2 / 3.0) * sqrt(f1 + (3.0 / 2.0) * f2);
      cgn1[8] = cgn[8] + bar;
    }
  } else {
    cgn1[7] = zero;
    cgn1[8] = zero;
  }
}

// mm05 -> mm05_fa (mm05.f:288) for one Gauss point.
// cgn (stress at n, 9) may be modified at step 1 (entries 7..9 zeroed) exactly like the
// reference's block-local copy; h / h1: history at n / n+1 (14); trial: trial elastic deviator
// handed to cnst5 (rtse).  rc: OK, CUT (adaptive step cut requested) or a fatal code.
W3D5_HD inline int fa_update(int step, int iter, const Props &P, double *cgn, double *cgn1,
                             const double *deps, double *h, double *h1, double *trial,
                             bool adaptive_possible) {
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0), two = W3D5_K(2.0), three = W3D5_K(3.0);
  const double root2third = W3D5_K(0.81649658);
  if (step == 1) {   // mm05_fa_step1_set
    cgn[6] = zero; cgn[7] = zero; cgn[8] = zero;
    h[0] = zero; h[1] = P.sigyld * root2third; h[2] = zero; h[3] = state_dword(3); h[4] = P.mm[2];
    for (int k = 5; k < 12; ++k) h[k] = zero;
  }
  // ---- mm05_fa_init
  const int iostat = dword_state(h[3]);
  const bool prior_linear = iostat == 3;
  const double trace_deps = deps[0] + deps[1] + deps[2];
  const double deps_mean = trace_deps / three;
  const double de1 = deps[0] - deps_mean, de2 = deps[1] - deps_mean, de3 = deps[2] - deps_mean;
  const double e = P.e, nu = P.nu;
  const double g = e / two / (one + nu);
  const double een1 = (cgn[0] - nu * (cgn[1] + cgn[2])) / e;
  const double een2 = (cgn[1] - nu * (cgn[0] + cgn[2])) / e;
  const double een3 = (cgn[2] - nu * (cgn[0] + cgn[1])) / e;
  const double een4 = cgn[3] / g, een5 = cgn[4] / g, een6 = cgn[5] / g;
  const double trace_eps_np1 = een1 + een2 + een3 + trace_deps;
  const double eps_mean = (een1 + een2 + een3) / three;
  const double e1 = (een1 - eps_mean) + de1, e2 = (een2 - eps_mean) + de2, e3 = (een3 - eps_mean) + de3;
  const double e4 = een4 + deps[3], e5 = een5 + deps[4], e6 = een6 + deps[5];
  const double shear_mod = g;
  trial[0] = two * shear_mod * e1; trial[1] = two * shear_mod * e2; trial[2] = two * shear_mod * e3;
  trial[3] = shear_mod * e4; trial[4] = shear_mod * e5; trial[5] = shear_mod * e6;
  double alpha_n[6];
  for (int k = 0; k < 6; ++k) alpha_n[k] = h[5 + k];
  const double k_n = h[1];
  double rtse[6];
  for (int k = 0; k < 6; ++k) rtse[k] = trial[k] - alpha_n[k];
  const double mrts = sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                           two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
  const double yf = mrts - k_n;
  const double yld_tol = W3D5_K(0.00001);
  int instat = 3;
  bool yield = false;
  if (yf >= yld_tol * k_n) { yield = true; instat = 1; }
  if (!(iter > 0)) {
    if (prior_linear) { yield = false; instat = 3; }
  }
  // ---- mm05_fa_compute
  if (yield) {
    const double aps_n = h[2];
    const double Q_bar = P.mm[0], b = P.mm[1], H_bar = P.mm[2], gamma = P.mm[4], sig_tol = P.mm[5];
    const double rel_tol = sig_tol * P.sigyld;
    const double mean_stress = (cgn[0] + cgn[1] + cgn[2]) / three;
    double dev_stress_n[6] = {cgn[0] - mean_stress, cgn[1] - mean_stress, cgn[2] - mean_stress, cgn[3], cgn[4], cgn[5]};
    const double mean_deps = (deps[0] + deps[1] + deps[2]) / three;
    double devdeps[6] = {deps[0] - mean_deps, deps[1] - mean_deps, deps[2] - mean_deps, deps[3], deps[4], deps[5]};
    double eta;
    int rc = elastic_fraction(g, dev_stress_n, alpha_n, devdeps, eta, k_n, prior_linear);
    if (rc != OK) return rc;
    double s1[6], a1[6], aps1, lam1, s2[6], a2[6], aps2, lam2, sig_t[6];
    rc = nsteps(1, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, P.sigyld, s1, a1, aps1, lam1,
                sig_t, adaptive_possible, iter);
    if (rc != OK) return rc;
    rc = nsteps(2, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, P.sigyld, s2, a2, aps2, lam2,
                sig_t, adaptive_possible, iter);
    if (rc != OK) return rc;
    double error = 0.0;
    for (int k = 0; k < 6; ++k) { const double dk = fabs(s2[k] - s1[k]); if (dk > error) error = dk; }
    int nsubinc = (int)ceil(W3D5_K(1.1) * error / rel_tol);
    if (nsubinc < 2) nsubinc = 2;
    h1[11] = nsubinc * one;
    double dev_np1[6], alpha_np1[6], aps_np1, lambda_n;
    if (nsubinc < 3) {
      for (int k = 0; k < 6; ++k) { dev_np1[k] = 2 * s2[k] - s1[k]; alpha_np1[k] = 2 * a2[k] - a1[k]; }
      aps_np1 = 2 * aps2 - aps1;
      lambda_n = lam2;
    } else {
      double sn[6], an[6], apsn;
      rc = nsteps(nsubinc, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, P.sigyld, sn, an, apsn,
                  lambda_n, sig_t, adaptive_possible, iter);
      if (rc != OK) return rc;
      const double a = (one * nsubinc) / (nsubinc - 2);
      const double bb = (one * 2) / (nsubinc - 2);
      for (int k = 0; k < 6; ++k) { dev_np1[k] = a * sn[k] - bb * s2[k]; alpha_np1[k] = a * an[k] - bb * a2[k]; }
      aps_np1 = a * apsn - bb * aps2;
    }
    for (int k = 0; k < 6; ++k) cgn1[k] = dev_np1[k];
    h1[0] = lambda_n;
    h1[1] = root2third * (P.sigyld + Q_bar * (one - exp(-b * aps_np1)));
    h1[2] = aps_np1;
    h1[4] = H_bar;
    for (int k = 0; k < 6; ++k) h1[5 + k] = alpha_np1[k];
    for (int k = 0; k < 6; ++k) trial[k] = sig_t[k];
  } else {
    h1[0] = zero; h1[1] = h[1]; h1[2] = h[2];
    for (int k = 4; k < 11; ++k) h1[k] = h[k];
    h1[11] = zero;
    for (int k = 0; k < 6; ++k) cgn1[k] = trial[k];
  }
  final_and_work(P, cgn, cgn1, deps, h1, trace_eps_np1, shear_mod, instat, yield, iter);
  return OK;
}

// cnst5 (mm05.f:1566), FA branch.  D: full symmetric 6x6 [r][c].
W3D5_HD inline void cnst5_fa(const Props &P, const double *sig_trial, const double *h1, const double *stress_np1,
                             double (&D)[6][6]) {
  const double zero = 0.0, one = 1.0, two = 2.0, twothird = 0.66666666, onethird = 0.33333333, half = 0.5,
               rt2third = 0.81649658, three = 3.0, four = 4.0;
  const double e = P.e, nu = P.nu;
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) D[r][c] = zero;
  if (dword_state(h1[3]) != 1) {
    const double c1 = e / ((one + nu) * (one - two * nu));
    const double c2 = (one - nu) * c1, c3 = ((one - two * nu) / two) * c1, c4 = nu * c1;
    D[0][0] = c2; D[1][1] = c2; D[2][2] = c2; D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
    D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
    return;
  }
  const double g = e / (two * (1 + nu));
  const double bulk = e / (three * (one - two * nu));
  const double Q_bar = P.mm[0], b0 = P.mm[1], gamma = P.mm[4];
  const double lambda = h1[0], k = h1[1], aps = h1[2], H = h1[4];
  const double dk = rt2third * b0 * Q_bar * (exp(-b0 * aps));
  const double dH = zero;
  const double lg = one + lambda * gamma;
  const double kappa = two * g * lambda / k;
  const double beta = lambda * H / k;
  const double ro = lg * (one + kappa) + beta;
  const double mean_sig = (stress_np1[0] + stress_np1[1] + stress_np1[2]) / three;
  double equiv[6], n[6];
  equiv[0] = stress_np1[0] - mean_sig - h1[5];
  equiv[1] = stress_np1[1] - mean_sig - h1[6];
  equiv[2] = stress_np1[2] - mean_sig - h1[7];
  equiv[3] = stress_np1[3] - h1[8];
  equiv[4] = stress_np1[4] - h1[9];
  equiv[5] = stress_np1[5] - h1[10];
  const double norm_equiv = sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                 two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
  for (int q = 0; q < 6; ++q) n[q] = equiv[q] / norm_equiv;
  const double dkappa = two * g * (k - lambda * dk) / (k * k);
  const double dbeta = H / k + (lambda / (k * k)) * (dH * k - H * dk);
  const double temp = n[0] * sig_trial[0] + n[1] * sig_trial[1] + n[2] * sig_trial[2] + two * n[3] * sig_trial[3] +
                      two * n[4] * sig_trial[4] + two * n[5] * sig_trial[5];
  const double zeta = (two * g * lg) / (dk + two * g + H + lambda * dH + gamma * (k + four * g * lambda + dk * lambda - temp));
  const double a = two * g / ro * (lg + beta);
  const double b = k * zeta / ro * (gamma * (kappa * (one + kappa) - lambda * dkappa) + dbeta * kappa - beta * dkappa - dkappa);
  const double c = gamma * zeta * kappa / ro;
  const double mean_n = onethird * (n[0] + n[1] + n[2]);
  double t[6], m[6];   // t_i = b n_i - c sig_trial_i ;  m_i = n_i - mean_n (normal) or n_i (shear)
  for (int q = 0; q < 6; ++q) { t[q] = b * n[q] - c * sig_trial[q]; m[q] = (q < 3) ? (n[q] - mean_n) : n[q]; }
  D[0][0] = (bulk + twothird * a + t[0] * m[0]);
  D[1][1] = (bulk + twothird * a + t[1] * m[1]);
  D[2][2] = (bulk + twothird * a + t[2] * m[2]);
  D[3][3] = (half * a + t[3] * n[3]);
  D[4][4] = (half * a + t[4] * n[4]);
  D[5][5] = (half * a + t[5] * n[5]);
  D[0][1] = (bulk - onethird * a + half * (t[0] * m[1] + t[1] * m[0]));
  D[0][2] = (bulk - onethird * a + half * (t[0] * m[2] + t[2] * m[0]));
  D[1][2] = (bulk - onethird * a + half * (t[1] * m[2] + t[2] * m[1]));
  for (int r = 0; r < 3; ++r)
    for (int s = 3; s < 6; ++s) D[r][s] = half * (t[r] * n[s] + t[s] * m[r]);
  D[3][4] = half * (t[3] * n[4] + t[4] * n[3]);
  D[3][5] = half * (t[3] * n[5] + t[5] * n[3]);
  D[4][5] = half * (t[4] * n[5] + t[5] * n[4]);
  for (int r = 0; r < 6; ++r)
    for (int s = r + 1; s < 6; ++s) D[s][r] = D[r][s];
}

// =============================================================================================
// generalized_plasticity option (mm05_gp mm05.f:1882, mm05_gp_step1_set :2059, mm05_gp_init :2128,
// mm05_gp_compute :2372, mm05_gp_nsteps :2651, mm05_gp_lambda_solve :2859,
// mm05_gp_elastic_fraction :3002, mm05_gp_felas_eta :3429, cnst5_gp :3584).
// Properties (inmat.f:2017-2036, rknstr.f:1290-1297): matprp(55) gp_h_u, (56) gp_tau,
// (57) gp_beta_u, (58) = -1, (59) gp_delta_u, (60) sig_tol; sigma_0 = props(23).
// The reference carries every property at n and at n+1 (temperature dependence through
// segmental curves); this path is isothermal, so both sets are equal here -- the formulas keep
// both so that the ratios (lk, mk, c1, c2) are evaluated as the reference evaluates them.
// History (14): lambda, k, eps_p, state, |sigma - alpha| of the last sub-increment, alpha(6),
// number of sub-increments, k_n, dHi.
// =============================================================================================
struct GpSet {   // 3-D values at one temperature (mm05_gp_compute :2432-2443)
  double g, Hi, Hk, beta, delta, k0;
};

// mm05_gp_felas_eta (mm05.f:3429): yield function after an elastic fraction eta of the step
W3D5_HD inline double gp_felas_eta(double eta, const GpSet &n, const GpSet &n1, const double *t, double eps_n) {
  const double one = W3D5_K(1.0), four = W3D5_K(4.0), tol = W3D5_K(0.00000001);
  const double g_b = eta * n1.g + (one - eta) * n.g;
  const double Hi_b = eta * n1.Hi + (one - eta) * n.Hi;
  const double Hk_b = eta * n1.Hk + (one - eta) * n.Hk;
  const double k0_b = eta * n1.k0 + (one - eta) * n.k0;
  const double c1 = g_b / n.g;
  double c2 = one;
  if (fabs(n.Hk) > tol) c2 = Hk_b / n.Hk;
  const double a1 = (c1 * c1) * t[0] - 2 * c1 * c2 * t[1] + (c2 * c2) * t[2];
  const double a2 = four * g_b * eta * (c1 * t[3] - c2 * t[4]);
  const double a3 = four * (g_b * g_b) * (eta * eta) * t[5];
  const double a4 = k0_b + Hi_b * eps_n;
  return sqrt(a1 + a2 + a3) - a4;
}

// mm05_gp_elastic_fraction (mm05.f:3002): Ridder's method for the elastic fraction eta of the
// step (golden-section search for the minimum of the yield function when the state at n is
// already outside the surface); updates stress_n, alpha_n, devdeps and returns the properties
// at eta in s.
W3D5_HD inline int gp_elastic_fraction(double *stress_n, double *alpha_n, double *devdeps, double &eta,
                                       double k_n, const GpSet &n, const GpSet &n1, GpSet &s, double eps_n,
                                       bool adaptive_possible) {
  const double zero = W3D5_K(0.0), half = W3D5_K(0.5), one = W3D5_K(1.0), two = W3D5_K(2.0),
               tol = W3D5_K(1.0e-10), stol = W3D5_K(1.0e-10), eta_comp = W3D5_K(1.0e-10),
               unused = W3D5_K(-1.11e30), ptol = W3D5_K(0.0001), c = W3D5_K(0.6180339887);
  double t[6];
  t[0] = stress_n[0] * stress_n[0] + stress_n[1] * stress_n[1] + stress_n[2] * stress_n[2] +
         two * (stress_n[3] * stress_n[3] + stress_n[4] * stress_n[4] + stress_n[5] * stress_n[5]);
  t[1] = stress_n[0] * alpha_n[0] + stress_n[1] * alpha_n[1] + stress_n[2] * alpha_n[2] +
         two * (stress_n[3] * alpha_n[3] + stress_n[4] * alpha_n[4] + stress_n[5] * alpha_n[5]);
  t[2] = alpha_n[0] * alpha_n[0] + alpha_n[1] * alpha_n[1] + alpha_n[2] * alpha_n[2] +
         two * (alpha_n[3] * alpha_n[3] + alpha_n[4] * alpha_n[4] + alpha_n[5] * alpha_n[5]);
  t[3] = devdeps[0] * stress_n[0] + devdeps[1] * stress_n[1] + devdeps[2] * stress_n[2] +
         devdeps[3] * stress_n[3] + devdeps[4] * stress_n[4] + devdeps[5] * stress_n[5];
  t[4] = devdeps[0] * alpha_n[0] + devdeps[1] * alpha_n[1] + devdeps[2] * alpha_n[2] +
         devdeps[3] * alpha_n[3] + devdeps[4] * alpha_n[4] + devdeps[5] * alpha_n[5];
  t[5] = devdeps[0] * devdeps[0] + devdeps[1] * devdeps[1] + devdeps[2] * devdeps[2] +
         half * (devdeps[3] * devdeps[3] + devdeps[4] * devdeps[4] + devdeps[5] * devdeps[5]);

  double eta_low = zero, eta_high = one;
  double f_low = gp_felas_eta(eta_low, n, n1, t, eps_n);
  double f_high = gp_felas_eta(eta_high, n, n1, t, eps_n);
  bool found = false;
  if (f_low < zero && f_high > zero) {
    // ---- 2. at most 50 passes of regula falsi with Ridder's update
    double eta_ridder = unused;
    for (int loop = 1; loop <= 50 && !found; ++loop) {
      const double eta_mid = half * (eta_low + eta_high);
      const double f_mid = gp_felas_eta(eta_mid, n, n1, t, eps_n);
      const double sridder = sqrt(f_mid * f_mid - f_low * f_high);
      if (sridder == zero) return adaptive_possible ? CUT : FATAL_FRACTION;
      const double eta_new = eta_mid + (eta_mid - eta_low) * (copysign(one, f_low - f_high) * f_mid / sridder);
      if (fabs(eta_new - eta_ridder) <= eta_comp) { eta = eta_new; found = true; break; }
      eta_ridder = eta_new;
      const double f_new = gp_felas_eta(eta_ridder, n, n1, t, eps_n);
      if (fabs(f_new) <= stol * k_n) { eta = eta_new; found = true; break; }
      if (copysign(fabs(f_mid), f_new) != f_mid) {
        eta_low = eta_mid; f_low = f_mid; eta_high = eta_ridder; f_high = f_new;
      } else if (copysign(fabs(f_low), f_new) != f_low) {
        eta_high = eta_ridder; f_high = f_new;
      } else if (copysign(fabs(f_high), f_new) != f_high) {
        eta_low = eta_ridder; f_low = f_new;
      } else {
        return CUT;   // the reference sets cut_step_size_now unconditionally here (:3260)
      }
      if (fabs(eta_high - eta_low) <= eta_comp) { eta = (eta_high + eta_low) * half; found = true; break; }
    }
    if (!found) return adaptive_possible ? CUT : FATAL_FRACTION;
  } else if (f_low > zero && f_high > zero) {
    // state at n already outside the surface: secant of the yield function near eta = 0
    const double eta_pt = eta_low + ptol;
    const double f_pt = gp_felas_eta(eta_pt, n, n1, t, eps_n);
    const double df = (f_pt - f_low) / ptol;
    if (df > zero) {
      eta = eta_low;
    } else {
      // ---- 3. golden-section search for the minimum of the yield function
      double a = eta_low, b = eta_high;
      double x1 = c * a + (one - c) * b;
      double fx1 = gp_felas_eta(x1, n, n1, t, eps_n);
      double x2 = (one - c) * a + c * b;
      double fx2 = gp_felas_eta(x2, n, n1, t, eps_n);
      for (int loop = 1; loop <= 10000 && !found; ++loop) {
        if (fx1 < fx2) {
          b = x2; x2 = x1; fx2 = fx1;
          x1 = c * a + (one - c) * b;
          fx1 = gp_felas_eta(x1, n, n1, t, eps_n);
        } else {
          a = x1; x1 = x2; fx1 = fx2;
          x2 = (one - c) * a + c * b;
          fx2 = gp_felas_eta(x2, n, n1, t, eps_n);
        }
        if (fabs(b - a) < eta_comp) { eta = half * (a + b); found = true; }
      }
      if (!found) return adaptive_possible ? CUT : FATAL_FRACTION;
    }
  } else {
    // remaining, improperly bracketed cases (:3150-3170)
    if (fabs(f_low) <= stol * k_n) eta = eta_low;
    else return FATAL_FRACTION;
  }
  // ---- 4. properties, stress and back stress at eta
  s.g = eta * n1.g + (one - eta) * n.g;
  s.Hi = eta * n1.Hi + (one - eta) * n.Hi;
  s.Hk = eta * n1.Hk + (one - eta) * n.Hk;
  s.k0 = eta * n1.k0 + (one - eta) * n.k0;
  s.beta = eta * n1.beta + (one - eta) * n.beta;
  s.delta = eta * n1.delta + (one - eta) * n.delta;
  stress_n[0] = s.g / n.g * stress_n[0] + eta * 2 * s.g * devdeps[0];
  stress_n[1] = s.g / n.g * stress_n[1] + eta * 2 * s.g * devdeps[1];
  stress_n[2] = s.g / n.g * stress_n[2] + eta * 2 * s.g * devdeps[2];
  stress_n[3] = s.g / n.g * stress_n[3] + eta * s.g * devdeps[3];
  stress_n[4] = s.g / n.g * stress_n[4] + eta * s.g * devdeps[4];
  stress_n[5] = s.g / n.g * stress_n[5] + eta * s.g * devdeps[5];
  double lk = one;
  if (fabs(n.Hk) > tol) lk = s.Hk / n.Hk;
  for (int k = 0; k < 6; ++k) alpha_n[k] = lk * alpha_n[k];
  for (int k = 0; k < 6; ++k) devdeps[k] = (1 - eta) * devdeps[k];
  return OK;
}

// mm05_gp_lambda_solve (mm05.f:2859): consistency parameter = smallest positive root of a quadratic
W3D5_HD inline int gp_lambda_solve(double norm_equiv_t, double norm_equiv_j, double k_j, double g_n1,
                                   double delta_n1, double beta_n1, double Hk_n1, double Hi_n1, double dHi,
                                   double k0_n1, bool adaptive_possible, double eps_j, double &lambda) {
  const double zero = W3D5_K(0.0), half = W3D5_K(0.5), two = W3D5_K(2.0);
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
  if (temp < zero) return adaptive_possible ? CUT : FATAL_LAMBDA;
  const double l1 = half / a * (-b - sqrt(temp));
  const double l2 = half / a * (-b + sqrt(temp));
  if (l1 < zero && l2 < zero) return adaptive_possible ? CUT : FATAL_LAMBDA;
  if (l1 > zero) {
    lambda = l1;
    if (lambda > l2 && l2 > zero) lambda = l2;
  } else {
    lambda = l2;
  }
  temp = a * (lambda * lambda) + b * lambda + c;
  temp = fabs(temp);
  if (temp > W3D5_K(0.000001)) return adaptive_possible ? CUT : FATAL_LAMBDA;
  return OK;
}

// mm05_gp_nsteps (mm05.f:2651): nsubinc equal sub-increments from the state (dev_stress_s,
// alpha_s, eps_s) with properties s to the end of the step with properties n1
W3D5_HD inline int gp_nsteps(int nsubinc, const double *devdeps, const double *dev_stress_s,
                             const double *alpha_s, double eps_s, const GpSet &s, const GpSet &n1,
                             double *stress_j, double *alpha_j, double &eps_j, double &lambda, double &dHi,
                             double *sig_t, double &norm_equiv_j, bool adaptive_possible) {
  const double one = W3D5_K(1.0), two = W3D5_K(2.0), tol = W3D5_K(0.00000001);
  const double dnsubinc = (double)nsubinc;
  const double d = one / dnsubinc;
  const double dHk = d * (n1.Hk - s.Hk);
  dHi = d * (n1.Hi - s.Hi);
  const double dg = d * (n1.g - s.g);
  const double dk0 = d * (n1.k0 - s.k0);
  const double dbeta = d * (n1.beta - s.beta);
  const double ddel = d * (n1.delta - s.delta);
  double Hk_j = s.Hk, Hi_j = s.Hi, g_j = s.g, k0_j = s.k0, beta_j = s.beta, delta_j = s.delta;
  for (int k = 0; k < 6; ++k) { stress_j[k] = dev_stress_s[k]; alpha_j[k] = alpha_s[k]; }
  eps_j = eps_s;
  for (int j = 1; j <= nsubinc; ++j) {
    const double Hk_j1 = Hk_j + dHk, Hi_j1 = Hi_j + dHi, g_j1 = g_j + dg, k0_j1 = k0_j + dk0,
                 beta_j1 = beta_j + dbeta, delta_j1 = delta_j + ddel;
    const double mk = g_j1 / g_j;
    double lk = one;
    if (fabs(Hk_j) > tol) lk = Hk_j1 / Hk_j;
    const double k_j = k0_j + Hi_j * eps_j;
    sig_t[0] = mk * stress_j[0] + two * g_j1 * devdeps[0] * d;
    sig_t[1] = mk * stress_j[1] + two * g_j1 * devdeps[1] * d;
    sig_t[2] = mk * stress_j[2] + two * g_j1 * devdeps[2] * d;
    sig_t[3] = mk * stress_j[3] + g_j1 * devdeps[3] * d;
    sig_t[4] = mk * stress_j[4] + g_j1 * devdeps[4] * d;
    sig_t[5] = mk * stress_j[5] + g_j1 * devdeps[5] * d;
    double alp_t[6], equiv_t[6];
    for (int k = 0; k < 6; ++k) { alp_t[k] = lk * alpha_j[k]; equiv_t[k] = sig_t[k] - alp_t[k]; }
    const double norm_equiv_t = sqrt(equiv_t[0] * equiv_t[0] + equiv_t[1] * equiv_t[1] + equiv_t[2] * equiv_t[2] +
                                     two * (equiv_t[3] * equiv_t[3] + equiv_t[4] * equiv_t[4] + equiv_t[5] * equiv_t[5]));
    norm_equiv_j = sqrt((stress_j[0] - alpha_j[0]) * (stress_j[0] - alpha_j[0]) +
                        (stress_j[1] - alpha_j[1]) * (stress_j[1] - alpha_j[1]) +
                        (stress_j[2] - alpha_j[2]) * (stress_j[2] - alpha_j[2]) +
                        two * (stress_j[3] - alpha_j[3]) * (stress_j[3] - alpha_j[3]) +
                        two * (stress_j[4] - alpha_j[4]) * (stress_j[4] - alpha_j[4]) +
                        two * (stress_j[5] - alpha_j[5]) * (stress_j[5] - alpha_j[5]));
    const int rc = gp_lambda_solve(norm_equiv_t, norm_equiv_j, k_j, g_j1, delta_j1, beta_j1, Hk_j1, Hi_j1, dHi,
                                   k0_j1, adaptive_possible, eps_j, lambda);
    if (rc != OK) return rc;
    const double eps_j1 = eps_j + lambda;
    for (int k = 0; k < 6; ++k) {
      const double sj1 = sig_t[k] - 2 * g_j1 * lambda * equiv_t[k] / norm_equiv_t;
      const double aj1 = alp_t[k] + Hk_j1 * lambda * equiv_t[k] / norm_equiv_t;
      stress_j[k] = sj1;
      alpha_j[k] = aj1;
    }
    eps_j = eps_j1;
    Hk_j = Hk_j1; Hi_j = Hi_j1; g_j = g_j1; k0_j = k0_j1; beta_j = beta_j1; delta_j = delta_j1;
  }
  return OK;
}

// mm05 -> mm05_gp (mm05.f:1882) for one Gauss point; arguments as fa_update.
W3D5_HD inline int gp_update(int step, int iter, const Props &P, double *cgn, double *cgn1,
                             const double *deps, double *h, double *h1, double *trial,
                             bool adaptive_possible) {
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0), two = W3D5_K(2.0), three = W3D5_K(3.0);
  const double twothird = W3D5_K(0.66666666), root2third = W3D5_K(0.81649658);
  // isothermal: the property set at n equals the one at n+1
  const double h_gp_n = P.mm[0], h_gp_np1 = P.mm[0], tau = P.mm[1];
  const double beta_gp_n = P.mm[2], beta_gp_np1 = P.mm[2], delta_gp_n = P.mm[4], delta_gp_np1 = P.mm[4];
  const double sigyld_n = P.sigyld, sigyld_np1 = P.sigyld;
  if (step == 1) {   // mm05_gp_step1_set
    cgn[6] = zero; cgn[7] = zero; cgn[8] = zero;
    h[0] = zero; h[1] = sigyld_n * root2third; h[2] = zero; h[3] = state_dword(3);
    for (int k = 4; k < 14; ++k) h[k] = zero;
  }
  // ---- mm05_gp_init
  const double yld_tol = W3D5_K(0.0000001), root2third_i = W3D5_K(0.8164965809);
  const int iostat = dword_state(h[3]);
  const bool prior_linear = iostat == 3;
  const double trace_deps = deps[0] + deps[1] + deps[2];
  const double deps_mean = trace_deps / three;
  const double de1 = deps[0] - deps_mean, de2 = deps[1] - deps_mean, de3 = deps[2] - deps_mean;
  const double e_n = P.e, nu_n = P.nu;
  const double g_n = e_n / (two * (one + nu_n));
  const double een1 = (cgn[0] - nu_n * (cgn[1] + cgn[2])) / e_n;
  const double een2 = (cgn[1] - nu_n * (cgn[0] + cgn[2])) / e_n;
  const double een3 = (cgn[2] - nu_n * (cgn[0] + cgn[1])) / e_n;
  const double een4 = cgn[3] / g_n, een5 = cgn[4] / g_n, een6 = cgn[5] / g_n;
  const double trace_eps_np1 = een1 + een2 + een3 + trace_deps;
  const double eps_mean = (een1 + een2 + een3) / three;
  const double e1 = (een1 - eps_mean) + de1, e2 = (een2 - eps_mean) + de2, e3 = (een3 - eps_mean) + de3;
  const double e4 = een4 + deps[3], e5 = een5 + deps[4], e6 = een6 + deps[5];
  const double e_n1 = P.e, nu_n1 = P.nu;
  const double g_n1 = e_n1 / (two * (one + nu_n1));
  trial[0] = two * g_n1 * e1; trial[1] = two * g_n1 * e2; trial[2] = two * g_n1 * e3;
  trial[3] = g_n1 * e4; trial[4] = g_n1 * e5; trial[5] = g_n1 * e6;
  double alpha_n[6];
  for (int k = 0; k < 6; ++k) alpha_n[k] = h[5 + k];
  bool yield;
  int instat;
  {
    const double hk_n = (one - tau) * twothird * h_gp_n;
    const double hk_n1 = (one - tau) * twothird * h_gp_np1;
    double lk = one;
    if (fabs(hk_n) > yld_tol) lk = hk_n1 / hk_n;
    double rtse[6], rsen[6];
    for (int k = 0; k < 6; ++k) rtse[k] = trial[k] - lk * alpha_n[k];
    const double norm_rtse = sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                                  two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
    const double stress_n_mean = (cgn[0] + cgn[1] + cgn[2]) / three;
    rsen[0] = cgn[0] - stress_n_mean - alpha_n[0];
    rsen[1] = cgn[1] - stress_n_mean - alpha_n[1];
    rsen[2] = cgn[2] - stress_n_mean - alpha_n[2];
    rsen[3] = cgn[3] - alpha_n[3];
    rsen[4] = cgn[4] - alpha_n[4];
    rsen[5] = cgn[5] - alpha_n[5];
    const double norm_rsen = sqrt(rsen[0] * rsen[0] + rsen[1] * rsen[1] + rsen[2] * rsen[2] +
                                  two * (rsen[3] * rsen[3] + rsen[4] * rsen[4] + rsen[5] * rsen[5]));
    const double k_n = h[1];
    const double hi_n1 = tau * (twothird * h_gp_np1);
    const double eps_n = h[2];
    const double kb_n1 = root2third_i * sigyld_np1 + hi_n1 * eps_n;
    const double f_n = norm_rsen - k_n;
    const double f_n1 = norm_rtse - kb_n1;
    const double F_trial = f_n1 - f_n;
    instat = 1;
    yield = true;
    if ((f_n1 < yld_tol * k_n) || (F_trial < yld_tol * k_n)) { yield = false; instat = 3; }
    if (!(iter > 0)) {
      if (prior_linear) { yield = false; instat = 3; }
    }
  }
  if (yield) {
    // ---- mm05_gp_compute
    GpSet n, n1, s;
    n.g = g_n; n1.g = g_n1;
    n.Hi = tau * twothird * h_gp_n;            n1.Hi = tau * twothird * h_gp_np1;
    n.Hk = (one - tau) * twothird * h_gp_n;    n1.Hk = (one - tau) * twothird * h_gp_np1;
    n.beta = root2third * beta_gp_n;           n1.beta = root2third * beta_gp_np1;
    n.delta = twothird * delta_gp_n;           n1.delta = twothird * delta_gp_np1;
    n.k0 = root2third * sigyld_n;              n1.k0 = root2third * sigyld_np1;
    const double sig_tol = one * P.mm[5];
    const double k_n = h[1], eps_n = h[2];
    const double rel_tol = sig_tol * k_n;
    const double mean_stress = (cgn[0] + cgn[1] + cgn[2]) / three;
    double dev_stress_n[6] = {cgn[0] - mean_stress, cgn[1] - mean_stress, cgn[2] - mean_stress, cgn[3], cgn[4], cgn[5]};
    const double mean_deps = (deps[0] + deps[1] + deps[2]) / three;
    double devdeps[6] = {deps[0] - mean_deps, deps[1] - mean_deps, deps[2] - mean_deps, deps[3], deps[4], deps[5]};
    double eta;
    int rc = gp_elastic_fraction(dev_stress_n, alpha_n, devdeps, eta, k_n, n, n1, s, eps_n, adaptive_possible);
    if (rc != OK) return rc;
    double s1[6], a1[6], eps1, lam1, s2[6], a2[6], eps2, lam2, sig_t[6], dHis, norm_equiv_sm1;
    rc = gp_nsteps(1, devdeps, dev_stress_n, alpha_n, eps_n, s, n1, s1, a1, eps1, lam1, dHis, sig_t,
                   norm_equiv_sm1, adaptive_possible);
    if (rc != OK) return rc;
    rc = gp_nsteps(2, devdeps, dev_stress_n, alpha_n, eps_n, s, n1, s2, a2, eps2, lam2, dHis, sig_t,
                   norm_equiv_sm1, adaptive_possible);
    if (rc != OK) return rc;
    double error = 0.0;
    for (int k = 0; k < 6; ++k) { const double dk = fabs(s2[k] - s1[k]); if (dk > error) error = dk; }
    int nsubinc = (int)ceil(W3D5_K(1.1) * error / rel_tol);
    if (nsubinc < 2) nsubinc = 2;
    const int mxsubinc = 10;
    if (nsubinc > mxsubinc) nsubinc = mxsubinc;
    h1[11] = nsubinc * one;
    double dev_np1[6], alpha_np1[6], eps_np1, lambda_nstep;
    if (nsubinc < 3) {
      for (int k = 0; k < 6; ++k) { dev_np1[k] = two * s2[k] - s1[k]; alpha_np1[k] = two * a2[k] - a1[k]; }
      eps_np1 = two * eps2 - eps1;
      lambda_nstep = lam2;
    } else {
      double sn[6], an[6], epsn;
      rc = gp_nsteps(nsubinc, devdeps, dev_stress_n, alpha_n, eps_n, s, n1, sn, an, epsn, lambda_nstep, dHis, sig_t,
                     norm_equiv_sm1, adaptive_possible);
      if (rc != OK) return rc;
      const double a = (one * nsubinc) / (nsubinc - two);
      const double b = (one * two) / (nsubinc - two);
      for (int k = 0; k < 6; ++k) { dev_np1[k] = a * sn[k] - b * s2[k]; alpha_np1[k] = a * an[k] - b * a2[k]; }
      eps_np1 = a * epsn - b * eps2;
    }
    for (int k = 0; k < 6; ++k) cgn1[k] = dev_np1[k];
    h1[0] = lambda_nstep;
    h1[1] = n1.k0 + n1.Hi * eps_np1;
    h1[2] = eps_np1;
    h1[4] = norm_equiv_sm1;
    for (int k = 0; k < 6; ++k) h1[5 + k] = alpha_np1[k];
    h1[12] = k_n;
    h1[13] = dHis;
    for (int k = 0; k < 6; ++k) trial[k] = sig_t[k];
  } else {
    // ---- elastic points (mm05_gp :1985-2016)
    const double htol = W3D5_K(0.000001);
    const double eps_n = h[2];
    const double hi_n = twothird * tau * h_gp_n, hi_n1 = twothird * tau * h_gp_np1;
    const double hk_n = twothird * (one - tau) * h_gp_n, hk_n1 = twothird * (one - tau) * h_gp_np1;
    double lk = one;
    if (fabs(hk_n) > htol) lk = hk_n1 / hk_n;
    h1[0] = zero;
    h1[1] = root2third * sigyld_np1 + hi_n1 * eps_n;
    h1[2] = h[2];
    for (int k = 0; k < 6; ++k) h1[5 + k] = lk * h[5 + k];
    h1[11] = zero;
    h1[12] = root2third * sigyld_n + hi_n * eps_n;
    h1[13] = hi_n1 - hi_n;
    for (int k = 0; k < 6; ++k) cgn1[k] = trial[k];
    double rse[6];
    for (int k = 0; k < 6; ++k) rse[k] = cgn1[k] - h1[5 + k];
    h1[4] = sqrt(rse[0] * rse[0] + rse[1] * rse[1] + rse[2] * rse[2] +
                 two * (rse[3] * rse[3] + rse[4] * rse[4] + rse[5] * rse[5]));
  }
  final_and_work(P, cgn, cgn1, deps, h1, trace_eps_np1, g_n1, instat, yield, iter);
  return OK;
}

// cnst5_gp (mm05.f:3584).  D: full symmetric 6x6 [r][c]; its constants carry d0 exponents.
W3D5_HD inline void cnst5_gp(const Props &P, const double *h1, const double *stress_np1, double (&D)[6][6]) {
  const double zero = 0.0, one = 1.0, two = 2.0, twothird = 0.66666666, onethird = 0.33333333, half = 0.5,
               rt2third = 0.81649658, three = 3.0;
  const double e = P.e, nu = P.nu;
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) D[r][c] = zero;
  if (dword_state(h1[3]) != 1) {
    const double c1 = e / ((one + nu) * (one - two * nu));
    const double c2 = (one - nu) * c1, c3 = ((one - two * nu) / two) * c1, c4 = nu * c1;
    D[0][0] = c2; D[1][1] = c2; D[2][2] = c2; D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
    D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
    return;
  }
  const double g = e / (two * (one + nu));
  const double bulk = e / (three * (one - two * nu));
  const double del_n1 = twothird * P.mm[4];
  const double beta_n1 = rt2third * P.mm[2];
  const double H_n1 = twothird * P.mm[0];
  const double Hk_n1 = (one - P.mm[1]) * H_n1;
  const double G1 = g + half * H_n1;
  const double lambda = h1[0], k_n1 = h1[1], dHi = h1[13];
  const double norm_equiv_n = h1[4], k_n = h1[12];
  const double mean_sig = (stress_np1[0] + stress_np1[1] + stress_np1[2]) / three;
  double equiv[6], n[6];
  equiv[0] = stress_np1[0] - h1[5] - mean_sig;
  equiv[1] = stress_np1[1] - h1[6] - mean_sig;
  equiv[2] = stress_np1[2] - h1[7] - mean_sig;
  equiv[3] = stress_np1[3] - h1[8];
  equiv[4] = stress_np1[4] - h1[9];
  equiv[5] = stress_np1[5] - h1[10];
  const double norm_equiv_np1 = sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
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
  for (int q = 0; q < 6; ++q) n[q] = equiv[q] / norm_equiv_np1;
  D[0][0] = bulk + twothird * c1 + c2 * n[0] * n[0];
  D[1][1] = bulk + twothird * c1 + c2 * n[1] * n[1];
  D[2][2] = bulk + twothird * c1 + c2 * n[2] * n[2];
  D[3][3] = half * c1 + c2 * n[3] * n[3];
  D[4][4] = half * c1 + c2 * n[4] * n[4];
  D[5][5] = half * c1 + c2 * n[5] * n[5];
  D[0][1] = (bulk - onethird * c1 + c2 * n[0] * n[1]);
  D[0][2] = (bulk - onethird * c1 + c2 * n[0] * n[2]);
  D[1][2] = bulk - onethird * c1 + c2 * n[1] * n[2];
  for (int r = 0; r < 3; ++r)
    for (int s = 3; s < 6; ++s) D[r][s] = c2 * n[r] * n[s];
  D[3][4] = c2 * n[3] * n[4];
  D[3][5] = c2 * n[3] * n[5];
  D[4][5] = c2 * n[4] * n[5];
  for (int r = 0; r < 6; ++r)
    for (int s = r + 1; s < 6; ++s) D[s][r] = D[r][s];
}

// option dispatch of mm05 (mm05.f:232-254) and cnst5 (:1566): matprp(58) < 0 selects the
// generalized_plasticity option, > 0 the nonlinear_hardening one
W3D5_HD inline bool is_gp(const Props &P) { return P.mm[3] < 0.0; }
W3D5_HD inline int update(int step, int iter, const Props &P, double *cgn, double *cgn1, const double *deps,
                          double *h, double *h1, double *trial, bool adaptive_possible) {
  return is_gp(P) ? gp_update(step, iter, P, cgn, cgn1, deps, h, h1, trial, adaptive_possible)
                  : fa_update(step, iter, P, cgn, cgn1, deps, h, h1, trial, adaptive_possible);
}
W3D5_HD inline void cnst5(const Props &P, const double *sig_trial, const double *h1, const double *stress_np1,
                          double (&D)[6][6]) {
  if (is_gp(P)) cnst5_gp(P, h1, stress_np1, D);
  else cnst5_fa(P, sig_trial, h1, stress_np1, D);
}

}  // namespace w3d_mm05
