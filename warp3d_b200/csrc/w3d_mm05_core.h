// w3d_mm05_core.h -- cyclic plasticity model mm05, "nonlinear_hardening" (Frederick-Armstrong)
// option, for ONE Gauss point.  Single source compiled by nvcc (engine) and g++ (CPU check), like
// w3d_mm10_core.h and for the same reason: the scalar Newton solve for the plastic multiplier and
// the Richardson sub-increment count are tolerance / threshold driven.
//
// Follows reference src/mm05.f: mm05 :9, mm05_fa :288, mm05_fa_step1_set :424, mm05_fa_init :490,
// mm05_fa_compute :648, mm05_fa_nsteps :883, mm05_fa_lambda_solve :981, mm05_find_pos_lambda
// :1120, mm05_compute_residual :1269, mm05_update_lambda :1323, mm05_fa_elastic_fraction :1367,
// cnst5 :1566, mm05_final :3879, mm05_plastic_work :3949; properties src/rknstr.f:1234-1317,
// src/inmat.f:2003-2040.  The generalized_plasticity option (mm05_gp :1882, cnst5_gp :3584) is
// not built yet.
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
  const double zero = W3D5_K(0.0), one = W3D5_K(1.0);
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

}  // namespace w3d_mm05
