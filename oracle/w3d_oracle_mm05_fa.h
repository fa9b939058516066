/* w3d_oracle_mm05_fa.h -- CPU ORACLE (test infrastructure only; never part of the product path).
 *
 * The nonlinear_hardening (Frederick-Armstrong) option of the reference's cyclic plasticity model
 * mm05, restated from src/mm05.f a second time, independently of w3d_oracle_mm05.h and of the
 * engine's device core (which descend from one earlier per-point transcription): written in
 * round 2 directly from the Fortran, in the Fortran's own BLOCK form -- every routine loops over
 * the `span` elements of a block at one Gauss point, arrays are (element, component) -- without
 * consulting the per-point version.  tests/test_mm05_fa_second_restatement.py holds the two
 * restatements against each other; the oracle driver uses this one (and w3d_oracle_mm05_gp.h for
 * the other option), so the engine's 1e-10 parity for mm05 is against an independently written
 * checker.
 *
 *   mm05_fa                   src/mm05.f:288-422
 *   mm05_fa_step1_set         :424-480
 *   mm05_fa_init              :490-640
 *   mm05_fa_compute           :648-880
 *   mm05_fa_nsteps            :883-975
 *   mm05_fa_lambda_solve      :981-1110
 *   mm05_find_pos_lambda      :1120-1260
 *   mm05_compute_residual     :1269-1315
 *   mm05_update_lambda        :1323-1360
 *   mm05_fa_elastic_fraction  :1367-1480
 *   cnst5 (FA branch)         :1566-1875
 *   mm05_final                :3879-3940
 *   mm05_plastic_work         :3949-4020
 *
 * Bare REAL*4 literals of DATA statements and assignments are evaluated in double (the reference's
 * ifort -fpconstant build); -DW3D_MM05_DATA_SINGLE evaluates them in single.
 * killed_status is all false (no element extinction on this path).  Parity status: pinned end to
 * end by verification/test60 a (tests/test_reference_known_answers.py); function level at 1e-10
 * unpinned (no Fortran compiler here).
 */
#ifndef W3D_ORACLE_MM05_FA_H
#define W3D_ORACLE_MM05_FA_H
#include <cmath>
#include <cstdint>
#include <cstring>

namespace w3do_mm05_fa {

#ifdef W3D_MM05_DATA_SINGLE
#define W3DO5F_K(x) ((double)(x##f))
#else
#define W3DO5F_K(x) (x)
#endif

constexpr int NH = 14;                       // history words per point (mm05_set_sizes)
enum { OK = 0, CUT = 1, ABORT_C_NEGATIVE = 21, ABORT_ETA_NEGATIVE = 22, ABORT_LAMBDA = 23, ABORT_POS_LAMBDA = 24 };

// the reference's arrays at one Gauss point of a block, (element, component)
struct Block {
  int span;
  double (*mm_props)[10];                    // mm_props(i,1:10) = matprp(55:64)
  double *e_vec, *nu_vec, *sigyld_vec;
  double (*trial)[6];                        // trial_elas_stress_np1
  double (*stress_n)[9], (*stress_np1)[9];
  double (*deps)[6];
  double (*history_n)[NH], (*history_np1)[NH];
};

inline double state_word(int32_t st) {       // iword(1) = st ; iword(2) = 0 ; dword
  int32_t iw[2] = {st, 0};
  double d;
  std::memcpy(&d, iw, 8);
  return d;
}

// mm05_fa_step1_set
inline void step1_set(Block &B) {
  const double root2third = W3DO5F_K(0.81649658), zero = W3DO5F_K(0.0);
  const double dword = state_word(3);
  for (int i = 0; i < B.span; ++i) {
    B.stress_n[i][6] = zero;
    B.stress_n[i][7] = zero;
    B.stress_n[i][8] = zero;
    B.history_n[i][0] = zero;
    B.history_n[i][1] = B.sigyld_vec[i] * root2third;
    B.history_n[i][2] = zero;
    B.history_n[i][3] = dword;
    B.history_n[i][4] = B.mm_props[i][2];
    for (int k = 5; k <= 11; ++k) B.history_n[i][k] = zero;
  }
}

// mm05_fa_init
inline void init(Block &B, int iter, int *iostat, bool *prior_linear, double *shear_mod_vec, double (*alpha_n)[6],
                 int *instat, bool *yield, double *trace_eps_np1, double *yf_vec) {
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), three = W3DO5F_K(3.0),
               yld_tol = W3DO5F_K(0.00001);
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
    const double e = B.e_vec[i], nu = B.nu_vec[i];
    const double g = e / two / (one + nu);
    const double *sn = B.stress_n[i];
    const double een1 = (sn[0] - nu * (sn[1] + sn[2])) / e;
    const double een2 = (sn[1] - nu * (sn[0] + sn[2])) / e;
    const double een3 = (sn[2] - nu * (sn[0] + sn[1])) / e;
    const double een4 = sn[3] / g, een5 = sn[4] / g, een6 = sn[5] / g;
    trace_eps_np1[i] = een1 + een2 + een3 + trace_deps;
    const double eps_mean = (een1 + een2 + een3) / three;
    const double e1 = (een1 - eps_mean) + de1, e2 = (een2 - eps_mean) + de2, e3 = (een3 - eps_mean) + de3;
    const double e4 = een4 + de4, e5 = een5 + de5, e6 = een6 + de6;
    shear_mod_vec[i] = g;
    B.trial[i][0] = two * shear_mod_vec[i] * e1;
    B.trial[i][1] = two * shear_mod_vec[i] * e2;
    B.trial[i][2] = two * shear_mod_vec[i] * e3;
    B.trial[i][3] = shear_mod_vec[i] * e4;
    B.trial[i][4] = shear_mod_vec[i] * e5;
    B.trial[i][5] = shear_mod_vec[i] * e6;
    for (int k = 0; k < 6; ++k) alpha_n[i][k] = B.history_n[i][5 + k];
    const double k_n = B.history_n[i][1];
    double rtse[6];
    for (int k = 0; k < 6; ++k) rtse[k] = B.trial[i][k] - alpha_n[i][k];
    const double mrts = std::sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                                  two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
    yf_vec[i] = mrts - k_n;
    instat[i] = 3;
    yield[i] = false;
    if (yf_vec[i] >= yld_tol * k_n) {
      yield[i] = true;
      instat[i] = 1;
    }
    if (iter > 0) continue;
    if (prior_linear[i]) {
      yield[i] = false;
      instat[i] = 3;
      continue;
    }
  }
}

// mm05_compute_residual
inline void compute_residual(double lambda, double aps_n, const double *sig_t, const double *alpha_n, double gamma,
                             double sigy, double H_bar, double Q_bar, double b, double g, double &k_np1,
                             double &H_np1, double &lg, double &res, double *equiv, double &aps_np1,
                             double &dk_np1, double &dH_np1) {
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), root2third = W3DO5F_K(0.8164956809);
  aps_np1 = aps_n + lambda;
  k_np1 = root2third * (sigy + Q_bar * (one - std::exp(-b * aps_np1)));
  dk_np1 = root2third * b * Q_bar * (std::exp(-b * aps_np1));
  H_np1 = H_bar;
  dH_np1 = zero;
  lg = one + lambda * gamma;
  for (int k = 0; k < 6; ++k) equiv[k] = lg * sig_t[k] - alpha_n[k];
  const double norm_equiv = std::sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                      two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
  res = lg * (k_np1 + two * g * lambda) + lambda * H_np1 - norm_equiv;
}

// mm05_update_lambda
inline void update_lambda(double &lambda, double res, const double *equiv, double gamma, double k_np1, double g,
                          double H_np1, double dk_np1, double dH_np1, const double *sig_t) {
  const double one = W3DO5F_K(1.0), two = W3DO5F_K(2.0);
  const double norm_equiv = std::sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                      two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
  const double eq_dot_sigt = equiv[0] * sig_t[0] + equiv[1] * sig_t[1] + equiv[2] * sig_t[2] +
                             two * equiv[3] * sig_t[3] + two * equiv[4] * sig_t[4] + two * equiv[5] * sig_t[5];
  const double dres = gamma * (k_np1 + two * g * lambda) + (one + lambda * gamma) * (dk_np1 + two * g) + H_np1 +
                      lambda * dH_np1 - gamma * eq_dot_sigt / norm_equiv;
  lambda = lambda - res / dres;
}

inline double fsign(double a, double b) { return std::copysign(std::fabs(a), b); }   // Fortran SIGN

// mm05_find_pos_lambda.  rc: OK, CUT or ABORT_POS_LAMBDA
inline int find_pos_lambda(double sigy, const double *sig_t, double *equiv, double aps_n, double gamma,
                           double &lambda, double &k_np1, double Q, double b, double H_bar, double &lg, double g,
                           const double *alpha_n, bool adaptive_possible, double &H_np1, double tol,
                           double &aps_np1) {
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0);
  const int sc_max = 4, bs_max = 10, nr_max = 10;
  double dk_np1, dH_np1, res_i, res_f, lambda_i, lambda_f, res;
  lambda_i = zero;
  int sc_count = 0, bs_count = 0, nr_count = 0;
  compute_residual(lambda_i, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_i, equiv,
                   aps_np1, dk_np1, dH_np1);
  lambda_f = -1 * lambda;
  compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_f, equiv,
                   aps_np1, dk_np1, dH_np1);
  while (fsign(one, res_i) * fsign(one, res_f) > zero) {
    sc_count = sc_count + 1;
    if (sc_count > sc_max) return adaptive_possible ? CUT : ABORT_POS_LAMBDA;
    lambda_f = 2 * lambda_f;
    compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_f, equiv,
                     aps_np1, dk_np1, dH_np1);
  }
  while (bs_count < bs_max) {
    bs_count = bs_count + 1;
    lambda = lambda_i + 0.5 * (lambda_f - lambda_i);
    while (nr_count < nr_max) {
      compute_residual(lambda, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res, equiv,
                       aps_np1, dk_np1, dH_np1);
      if (std::fabs(res) < tol * sigy) return OK;
      nr_count = nr_count + 1;
      update_lambda(lambda, res, equiv, gamma, k_np1, g, H_np1, dk_np1, dH_np1, sig_t);
      if (lambda < lambda_i || lambda > lambda_f) break;
    }
    lambda = lambda_i + 0.5 * (lambda_f - lambda_i);
    compute_residual(lambda, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res, equiv,
                     aps_np1, dk_np1, dH_np1);
    if ((fsign(one, res_i) * fsign(one, res)) > zero) {
      lambda_i = lambda;
      compute_residual(lambda_i, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_i, equiv,
                       aps_np1, dk_np1, dH_np1);
    } else {
      lambda_f = lambda;
      compute_residual(lambda_f, aps_n, sig_t, alpha_n, gamma, sigy, H_bar, Q, b, g, k_np1, H_np1, lg, res_f, equiv,
                       aps_np1, dk_np1, dH_np1);
    }
    nr_count = 0;
  }
  return adaptive_possible ? CUT : ABORT_POS_LAMBDA;
}

// mm05_fa_lambda_solve.  rc: OK, CUT, ABORT_*
inline int lambda_solve(double sig_y, const double *alpha_s, const double *sig_t, double *equiv, double &aps_n,
                        double gamma, double &lambda, double &k_np1, double Q_bar, double b, double H_bar, double &lg,
                        double g, bool adaptive_possible, double &H_np1, int iter) {
  const int max_iter = 15;
  const double zero = W3DO5F_K(0.0), one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), root2third = W3DO5F_K(0.8164956809);
  double tol;
  if (iter > 1) tol = W3DO5F_K(0.000001);
  else tol = W3DO5F_K(0.001);
  lambda = zero;
  for (int j = 1; j <= max_iter; ++j) {
    double aps_np1 = aps_n + lambda;
    k_np1 = root2third * (sig_y + Q_bar * (one - std::exp(-b * aps_np1)));
    const double dk_np1 = root2third * b * Q_bar * (std::exp(-b * aps_np1));
    H_np1 = H_bar;
    const double dH_np1 = zero;
    lg = one + lambda * gamma;
    equiv[0] = lg * sig_t[0] - alpha_s[0];
    equiv[1] = lg * sig_t[1] - alpha_s[1];
    equiv[2] = lg * sig_t[2] - alpha_s[2];
    equiv[3] = lg * sig_t[3] - alpha_s[3];
    equiv[4] = lg * sig_t[4] - alpha_s[4];
    equiv[5] = lg * sig_t[5] - alpha_s[5];
    const double norm_equiv = std::sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                        two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
    const double res = lg * (k_np1 + two * g * lambda) + lambda * H_np1 - norm_equiv;
    if (std::fabs(res) < tol * sig_y) {
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
    const double dres = gamma * (k_np1 + two * g * lambda) + lg * (dk_np1 + two * g) + H_np1 + lambda * dH_np1 -
                        gamma * eq_dot_sigt / norm_equiv;
    lambda = lambda - res / dres;
  }
  return adaptive_possible ? CUT : ABORT_LAMBDA;
}

// mm05_fa_nsteps
inline int nsteps(int nsubinc, const double *devdeps, const double *dev_stress_n, const double *alpha_n, double aps_n,
                  double g, double Q_bar, double b, double H_bar, double gamma, double sigyld, double *stress_j,
                  double *alpha_j, double &aps_j, double &lambda, double *sig_t, bool adaptive_possible, int iter) {
  const double two = W3DO5F_K(2.0), one = W3DO5F_K(1.0);
  double equiv[6], k_np1 = 0.0, H_np1 = 0.0, lg = 0.0;
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
    const int rc = lambda_solve(sigyld, alpha_j, sig_t, equiv, aps_j, gamma, lambda, k_np1, Q_bar, b, H_bar, lg, g,
                                adaptive_possible, H_np1, iter);
    if (rc != OK) return rc;      // the reference carries on with stale values; the step is cut or the job dies
    const double kappa = two * g * lambda / k_np1;
    const double beta = lambda * H_np1 / k_np1;
    const double ro = lg * (one + kappa) + beta;
    for (int k = 0; k < 6; ++k) stress_j[k] = sig_t[k] - kappa / ro * equiv[k];
    for (int k = 0; k < 6; ++k) alpha_j[k] = (alpha_j[k] + beta / ro * equiv[k]) / lg;
  }
  return OK;
}

// mm05_fa_elastic_fraction
inline int elastic_fraction(double g, double *stress_n, const double *alpha_n, double *devdeps, double &eta,
                            double k_n, bool prior_linear) {
  const double zero = W3DO5F_K(0.0), two = W3DO5F_K(2.0), tol = W3DO5F_K(0.00001);
  eta = zero;
  const double a = two * g *
                   ((stress_n[0] - alpha_n[0]) * devdeps[0] + (stress_n[1] - alpha_n[1]) * devdeps[1] +
                    (stress_n[2] - alpha_n[2]) * devdeps[2] + (stress_n[3] - alpha_n[3]) * devdeps[3] +
                    (stress_n[4] - alpha_n[4]) * devdeps[4] + (stress_n[5] - alpha_n[5]) * devdeps[5]);
  const double b = g * g *
                   (4 * (devdeps[0] * devdeps[0] + devdeps[1] * devdeps[1] + devdeps[2] * devdeps[2]) +
                    2 * (devdeps[3] * devdeps[3] + devdeps[4] * devdeps[4] + devdeps[5] * devdeps[5]));
  double c = 0;
  if (prior_linear) {
    c = ((stress_n[0] - alpha_n[0]) * (stress_n[0] - alpha_n[0]) + (stress_n[1] - alpha_n[1]) * (stress_n[1] - alpha_n[1]) +
         (stress_n[2] - alpha_n[2]) * (stress_n[2] - alpha_n[2]) +
         two * ((stress_n[3] - alpha_n[3]) * (stress_n[3] - alpha_n[3]) +
                (stress_n[4] - alpha_n[4]) * (stress_n[4] - alpha_n[4]) +
                (stress_n[5] - alpha_n[5]) * (stress_n[5] - alpha_n[5])));
    const double temp = k_n - std::sqrt(c);
    c = k_n * k_n - c;
    if (c < zero) {
      if (temp < tol * k_n) c = zero;
      else return ABORT_C_NEGATIVE;
    }
  }
  if (c == zero) {
    if (a > zero) eta = zero;
    else eta = -two * a / b;
  } else {
    eta = -a / b + std::sqrt((a / b) * (a / b) + c / b);
    if (eta < zero) return ABORT_ETA_NEGATIVE;
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

// mm05_fa_compute.  rc: OK, CUT (cut_step_size_now) or an ABORT_* code (die_abort)
inline int compute(Block &B, int iter, const bool *yield, const double *shear_mod_vec, const bool *prior_linear,
                   bool adaptive_possible) {
  const double two = W3DO5F_K(2.0), one = W3DO5F_K(1.0), three = W3DO5F_K(3.0), root2third = W3DO5F_K(0.81649658);
  for (int i = 0; i < B.span; ++i) {
    if (!yield[i]) continue;
    const double g = shear_mod_vec[i];
    const double aps_n = B.history_n[i][2];
    const double k_n = B.history_n[i][1];
    const double Q_bar = B.mm_props[i][0], b = B.mm_props[i][1], H_bar = B.mm_props[i][2], gamma = B.mm_props[i][4],
                 sig_tol = B.mm_props[i][5];
    const double rel_tol = sig_tol * B.sigyld_vec[i];
    const double *sn = B.stress_n[i];
    const double mean_stress = (sn[0] + sn[1] + sn[2]) / three;
    double dev_stress_n[6] = {sn[0] - mean_stress, sn[1] - mean_stress, sn[2] - mean_stress, sn[3], sn[4], sn[5]};
    double alpha_n[6];
    for (int k = 0; k < 6; ++k) alpha_n[k] = B.history_n[i][5 + k];
    const double *dp = B.deps[i];
    const double mean_deps = (dp[0] + dp[1] + dp[2]) / three;
    double devdeps[6] = {dp[0] - mean_deps, dp[1] - mean_deps, dp[2] - mean_deps, dp[3], dp[4], dp[5]};
    double eta;
    int rc = elastic_fraction(g, dev_stress_n, alpha_n, devdeps, eta, k_n, prior_linear[i]);
    if (rc != OK) return rc;
    double stress_1step[6], alpha_1step[6], aps_1step, lambda_1step, sig_t[6];
    double stress_2step[6], alpha_2step[6], aps_2step, lambda_2step;
    double stress_nstep[6], alpha_nstep[6], aps_nstep, lambda_nstep;
    rc = nsteps(1, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, B.sigyld_vec[i], stress_1step,
                alpha_1step, aps_1step, lambda_1step, sig_t, adaptive_possible, iter);
    if (rc != OK) return rc;
    rc = nsteps(2, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, B.sigyld_vec[i], stress_2step,
                alpha_2step, aps_2step, lambda_2step, sig_t, adaptive_possible, iter);
    if (rc != OK) return rc;
    double error = 0.0;
    for (int k = 0; k < 6; ++k) {
      const double diff = std::fabs(stress_2step[k] - stress_1step[k]);
      if (k == 0 || diff > error) error = diff;
    }
    int nsubinc = (int)std::ceil(W3DO5F_K(1.1) * error / rel_tol);
    if (nsubinc < two) nsubinc = (int)two;
    B.history_np1[i][11] = nsubinc * one;
    double dev_stress_np1[6], alpha_np1[6], aps_np1;
    if (nsubinc < three) {
      for (int k = 0; k < 6; ++k) dev_stress_np1[k] = 2 * stress_2step[k] - stress_1step[k];
      for (int k = 0; k < 6; ++k) alpha_np1[k] = 2 * alpha_2step[k] - alpha_1step[k];
      aps_np1 = 2 * aps_2step - aps_1step;
      lambda_nstep = lambda_2step;
    } else {
      rc = nsteps(nsubinc, devdeps, dev_stress_n, alpha_n, aps_n, g, Q_bar, b, H_bar, gamma, B.sigyld_vec[i],
                  stress_nstep, alpha_nstep, aps_nstep, lambda_nstep, sig_t, adaptive_possible, iter);
      if (rc != OK) return rc;
      const double a = (one * nsubinc) / (nsubinc - 2);
      const double bb = (one * 2) / (nsubinc - 2);
      for (int k = 0; k < 6; ++k) dev_stress_np1[k] = a * stress_nstep[k] - bb * stress_2step[k];
      for (int k = 0; k < 6; ++k) alpha_np1[k] = a * alpha_nstep[k] - bb * alpha_2step[k];
      aps_np1 = a * aps_nstep - bb * aps_2step;
    }
    for (int k = 0; k < 6; ++k) B.stress_np1[i][k] = dev_stress_np1[k];
    B.history_np1[i][0] = lambda_nstep;
    B.history_np1[i][1] = root2third * (B.sigyld_vec[i] + Q_bar * (one - std::exp(-b * aps_np1)));
    B.history_np1[i][2] = aps_np1;
    B.history_np1[i][4] = H_bar;
    for (int k = 0; k < 6; ++k) B.history_np1[i][5 + k] = alpha_np1[k];
    for (int k = 0; k < 6; ++k) B.trial[i][k] = sig_t[k];
  }
  return OK;
}

// mm05_final
inline void final_(Block &B, const double *shear_mod_vec, const int *instat, const double *trace_eps_np1) {
  const double one = W3DO5F_K(1.0), two = W3DO5F_K(2.0), three = W3DO5F_K(3.0), half = W3DO5F_K(0.5);
  for (int i = 0; i < B.span; ++i) {
    const double e = B.e_vec[i], nu = B.nu_vec[i];
    const double sig_mean_np1 =
        trace_eps_np1[i] * (three * e * nu / ((one + nu) * (one - two * nu)) + two * shear_mod_vec[i]) / three;
    double *s1 = B.stress_np1[i];
    const double *sn = B.stress_n[i], *dp = B.deps[i];
    s1[0] = s1[0] + sig_mean_np1;
    s1[1] = s1[1] + sig_mean_np1;
    s1[2] = s1[2] + sig_mean_np1;
    s1[6] = sn[6] + half * (dp[0] * (sn[0] + s1[0]) + dp[1] * (sn[1] + s1[1]) + dp[2] * (sn[2] + s1[2]) +
                            dp[3] * (sn[3] + s1[3]) + dp[4] * (sn[4] + s1[4]) + dp[5] * (sn[5] + s1[5]));
    // iword(1) = instat(i); history_np1(i,4) = dword.  dword is a local of mm05_final that is never
    // assigned as a whole: its high word is whatever the stack held; the state lives in the low word
    // only, and like the rest of this oracle the undefined half is taken as zero
    B.history_np1[i][3] = state_word(instat[i]);
  }
}

// mm05_plastic_work
inline void plastic_work(Block &B, const bool *yield, const double *shear_mod_vec) {
  const double zero = 0.0, half = 0.5, root2 = std::sqrt(2.0), three = 3.0, two = 2.0;
  for (int i = 0; i < B.span; ++i) {
    double *s1 = B.stress_np1[i];
    const double *sn = B.stress_n[i], *dp = B.deps[i];
    s1[7] = sn[7];
    s1[8] = sn[8];
    if (yield[i]) {
      double deps_plas_bar = zero;
      double dsig[6], deps_plas[6];
      for (int k = 0; k < 6; ++k) dsig[k] = s1[k] - sn[k];
      deps_plas[0] = dp[0] - (dsig[0] - B.nu_vec[i] * (dsig[1] + dsig[2])) / B.e_vec[i];
      deps_plas[1] = dp[1] - (dsig[1] - B.nu_vec[i] * (dsig[0] + dsig[2])) / B.e_vec[i];
      deps_plas[2] = dp[2] - (dsig[2] - B.nu_vec[i] * (dsig[1] + dsig[0])) / B.e_vec[i];
      deps_plas[3] = dp[3] - dsig[3] / shear_mod_vec[i];
      deps_plas[4] = dp[4] - dsig[4] / shear_mod_vec[i];
      deps_plas[5] = dp[5] - dsig[5] / shear_mod_vec[i];
      s1[7] = sn[7] + half * (deps_plas[0] * (s1[0] + sn[0]) + deps_plas[1] * (s1[1] + sn[1]) +
                              deps_plas[2] * (s1[2] + sn[2]) + deps_plas[3] * (s1[3] + sn[3]) +
                              deps_plas[4] * (s1[4] + sn[4]) + deps_plas[5] * (s1[5] + sn[5]));
      const double factor1 = (deps_plas[0] - deps_plas[1]) * (deps_plas[0] - deps_plas[1]) +
                             (deps_plas[0] - deps_plas[2]) * (deps_plas[0] - deps_plas[2]) +
                             (deps_plas[2] - deps_plas[1]) * (deps_plas[2] - deps_plas[1]);
      const double factor2 = deps_plas[3] * deps_plas[3] + deps_plas[4] * deps_plas[4] + deps_plas[5] * deps_plas[5];
      deps_plas_bar = (root2 / three) * std::sqrt(factor1 + (three / two) * factor2);
      s1[8] = sn[8] + deps_plas_bar;
    }
  }
}

// mm05_fa.  rc: OK, CUT or ABORT_*
inline int mm05_fa(int step, int iter, Block &B, bool adaptive_possible) {
  const double zero = 0.0;
  const int span = B.span;
  bool yield[256], prior_linear[256];
  int iostat[256], instat[256];
  double shear_mod_vec[256], alpha_n[256][6], trace_eps_np1[256], yf_vec[256];
  if (span > 256) return ABORT_LAMBDA;
  if (step == 1) step1_set(B);
  init(B, iter, iostat, prior_linear, shear_mod_vec, alpha_n, instat, yield, trace_eps_np1, yf_vec);
  const int rc = compute(B, iter, yield, shear_mod_vec, prior_linear, adaptive_possible);
  if (rc != OK) return rc;                   // cut_step_size_now: return at once; die_abort: the job ends
  for (int i = 0; i < span; ++i) {
    if (yield[i]) continue;
    B.history_np1[i][0] = zero;
    B.history_np1[i][1] = B.history_n[i][1];
    B.history_np1[i][2] = B.history_n[i][2];
    for (int k = 4; k <= 10; ++k) B.history_np1[i][k] = B.history_n[i][k];
    B.history_np1[i][11] = zero;
    for (int k = 0; k < 6; ++k) B.stress_np1[i][k] = B.trial[i][k];
  }
  final_(B, shear_mod_vec, instat, trace_eps_np1);
  if (iter > 0) plastic_work(B, yield, shear_mod_vec);
  return OK;
}

// cnst5, nonlinear_hardening branch: dmat(i, r, c) for the block
inline void cnst5_fa(const Block &B, double (*dmat)[6][6]) {
  const double zero = 0.0, one = 1.0, two = 2.0, twothird = 0.66666666, onethird = 0.33333333, half = 0.5,
               rt2third = 0.81649658, three = 3.0, four = 4.0;
  for (int i = 0; i < B.span; ++i) {
    int32_t iw[2];
    std::memcpy(iw, &B.history_np1[i][3], 8);
    const bool yield = iw[0] == 1;
    double(*D)[6] = dmat[i];
    if (!yield) {
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) D[r][c] = zero;
      const double c1 = B.e_vec[i] / ((one + B.nu_vec[i]) * (one - two * B.nu_vec[i]));
      const double c2 = (one - B.nu_vec[i]) * c1;
      const double c3 = ((one - two * B.nu_vec[i]) / two) * c1;
      const double c4 = B.nu_vec[i] * c1;
      D[0][0] = c2; D[1][1] = c2; D[2][2] = c2;
      D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
      D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
      continue;
    }
    const double g = B.e_vec[i] / (two * (1 + B.nu_vec[i]));
    const double bulk = B.e_vec[i] / (three * (one - two * B.nu_vec[i]));
    const double Q_bar = B.mm_props[i][0];
    double b = B.mm_props[i][1];
    const double gamma = B.mm_props[i][4];
    const double lambda = B.history_np1[i][0];
    const double k = B.history_np1[i][1];
    const double aps = B.history_np1[i][2];
    const double H = B.history_np1[i][4];
    const double dk = rt2third * b * Q_bar * (std::exp(-b * aps));
    const double dH = zero;
    const double lg = one + lambda * gamma;
    const double kappa = two * g * lambda / k;
    const double beta = lambda * H / k;
    const double ro = lg * (one + kappa) + beta;
    const double *s1 = B.stress_np1[i];
    const double mean_sig = (s1[0] + s1[1] + s1[2]) / three;
    double equiv[6];
    equiv[0] = s1[0] - mean_sig - B.history_np1[i][5];
    equiv[1] = s1[1] - mean_sig - B.history_np1[i][6];
    equiv[2] = s1[2] - mean_sig - B.history_np1[i][7];
    equiv[3] = s1[3] - B.history_np1[i][8];
    equiv[4] = s1[4] - B.history_np1[i][9];
    equiv[5] = s1[5] - B.history_np1[i][10];
    const double norm_equiv = std::sqrt(equiv[0] * equiv[0] + equiv[1] * equiv[1] + equiv[2] * equiv[2] +
                                        two * (equiv[3] * equiv[3] + equiv[4] * equiv[4] + equiv[5] * equiv[5]));
    double n[6];
    for (int q = 0; q < 6; ++q) n[q] = equiv[q] / norm_equiv;
    const double dkappa = two * g * (k - lambda * dk) / (k * k);
    const double dbeta = H / k + (lambda / (k * k)) * (dH * k - H * dk);
    const double *st = B.trial[i];
    const double temp = n[0] * st[0] + n[1] * st[1] + n[2] * st[2] + two * n[3] * st[3] + two * n[4] * st[4] +
                        two * n[5] * st[5];
    const double zeta = (two * g * lg) / (dk + two * g + H + lambda * dH + gamma * (k + four * g * lambda + dk * lambda - temp));
    const double a = two * g / ro * (lg + beta);
    b = k * zeta / ro * (gamma * (kappa * (one + kappa) - lambda * dkappa) + dbeta * kappa - beta * dkappa - dkappa);
    const double c = gamma * zeta * kappa / ro;
    const double mean_n = onethird * (n[0] + n[1] + n[2]);
    double w[6], m[6];                       // (b n - c sig_trial) and the deviatoric direction terms
    for (int q = 0; q < 6; ++q) w[q] = (b * n[q] - c * st[q]);
    m[0] = (n[0] - mean_n); m[1] = (n[1] - mean_n); m[2] = (n[2] - mean_n);
    m[3] = n[3]; m[4] = n[4]; m[5] = n[5];
    D[0][0] = (bulk + twothird * a + w[0] * m[0]);
    D[1][1] = (bulk + twothird * a + w[1] * m[1]);
    D[2][2] = (bulk + twothird * a + w[2] * m[2]);
    D[3][3] = (half * a + w[3] * n[3]);
    D[4][4] = (half * a + w[4] * n[4]);
    D[5][5] = (half * a + w[5] * n[5]);
    D[0][1] = (bulk - onethird * a + half * (w[0] * m[1] + w[1] * m[0]));
    D[0][2] = (bulk - onethird * a + half * (w[0] * m[2] + w[2] * m[0]));
    D[0][3] = half * (w[0] * n[3] + w[3] * m[0]);
    D[0][4] = half * (w[0] * n[4] + w[4] * m[0]);
    D[0][5] = half * (w[0] * n[5] + w[5] * m[0]);
    D[1][2] = (bulk - onethird * a + half * (w[1] * m[2] + w[2] * m[1]));
    D[1][3] = half * (w[1] * n[3] + w[3] * m[1]);
    D[1][4] = half * (w[1] * n[4] + w[4] * m[1]);
    D[1][5] = half * (w[1] * n[5] + w[5] * m[1]);
    D[2][3] = half * (w[2] * n[3] + w[3] * m[2]);
    D[2][4] = half * (w[2] * n[4] + w[4] * m[2]);
    D[2][5] = half * (w[2] * n[5] + w[5] * m[2]);
    D[3][4] = half * (w[3] * n[4] + w[4] * n[3]);
    D[3][5] = half * (w[3] * n[5] + w[5] * n[3]);
    D[4][5] = half * (w[4] * n[5] + w[5] * n[4]);
    for (int r = 0; r < 6; ++r)
      for (int cc = 0; cc < r; ++cc) D[r][cc] = D[cc][r];
  }
}

}  // namespace w3do_mm05_fa
#endif
