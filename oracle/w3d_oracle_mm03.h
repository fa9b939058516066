/* w3d_oracle_mm03.h -- CPU ORACLE (test infrastructure only; never part of the product path).
 *
 * Statement-for-statement restatement of the reference's material model 3 for a MISES material
 * (no voids: f0 = 0, nucleation off) with constant or linear + power-law hardening, optionally
 * power-law rate sensitive; written from the Fortran, independent of the engine's device code:
 *   mm03p   src/mm03.f:10-361    block pre-processor (trial state, yield check, step-1 history)
 *   mm03    src/mm03.f:2063-2604 per-point update, elastic/unloading and plastic branches
 *   mm03q   src/mm03.f:2820-3100 scalar return by Ridder's method
 *   mm03qq  src/mm03.f:3113-3160 residual of the scalar equation
 *   mm03is  src/mm03.f:963-1077  linear + power-law uniaxial curve with cubic transition
 *   cnst3   src/mm03.f:3252-3790 consistent tangent (the Gurson expressions evaluated at f = 0)
 * NOT restated (the product refuses the same inputs): Gurson branch (mm03s and friends),
 * segmental curves (mm03sc), temperature dependence, fgm properties, killed elements.
 *
 * PARITY STATUS: **parity unpinned** against Fortran output.  The reference ships no golden vectors
 * for mm03, and every known answer of its verification suite that uses a `mises` material on hex8
 * (test18, test57, test88, test89, testH/I: verification/run_tests.py) is a J-integral or a crack-
 * growth result, i.e. needs the domain integral / element extinction that stay with the reference;
 * the Fortran cannot be compiled here.  What stands in: line-by-line transcription with citations,
 * the physical checks of tests/test_mm03_core.py (yield surface, uniaxial power-law curve, finite-
 * difference tangent, patch test) and the function-level dump replay of fortran/w3d_dump_hook.f +
 * tools/compare_fortran_dump.py for whoever has ifort.
 *
 * Literal oddities kept on purpose:
 *  - mm03's `third` is the bare literal 0.3333333333 of a DATA statement (:2130-2131): ten digits
 *    in double under -fpconstant (the default here, W3DO3_K), 0.3333333432674408 if the compiler
 *    evaluated it in single -- the mean stress removed in the radial return (:3064-3074) is off by
 *    1e-10 or 3e-8 relative accordingly; mm03p uses 1.d0/3.d0; cnst3 uses 0.333333333d0 and
 *    0.6666666667d0.
 *  - mm03q's tolerances are bare literals 0.000001 (:2843), same switch.
 *  - the state flag (+1 / -1) lives in the LOW 32-bit word of history word 6 through an
 *    EQUIVALENCE (:316-317, :2124): the double is a denormal.  The high word is 0 from the step-1
 *    set-up; where the Fortran leaves it undefined (all-null block, :66-67) this file writes 0.
 *  - a null point (sum |deps| <= 2e-13) that is already plastic skips mm03 (rstgp1.f:1937): its
 *    stress_n1(7:9) stay zero and history1(10:15) keep whatever the local block held -- the
 *    reference allocates that block uninitialised (drive_eps_sig_internal_forces.f:1145); here,
 *    as in the rest of this oracle, it starts from zero.
 */
#ifndef W3D_ORACLE_MM03_H
#define W3D_ORACLE_MM03_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace w3do_mm03 {

// The reference is built with ifort -fpconstant (src/Makefile.linux_Intel.inc:101): a bare REAL*4
// literal that initialises a double is evaluated in double (same convention as the
// mm05 restatement).  -DW3D_MM03_DATA_SINGLE evaluates them in single instead (0.3333333333 ->
// 0.3333333432674408, 0.000001 -> 9.99999997e-07).
#ifdef W3D_MM03_DATA_SINGLE
#define W3DO3_K(x) ((double)(x##f))
#else
#define W3DO3_K(x) (x)
#endif

constexpr int HIST = 15;       // mm03_set_sizes (mm03.f:1623-1640)

inline double state_word(int32_t state) {          // iword(1) = state; iword(2) = 0
  const int32_t iw[2] = {state, 0};
  double d;
  std::memcpy(&d, iw, 8);
  return d;
}
inline int32_t word_state(double dword) {          // state = iword(1)
  int32_t iw[2];
  std::memcpy(iw, &dword, 8);
  return iw[0];
}

// mm03is (mm03.f:963-1077)
inline double mm03is(double eps_pls, double sigma_o, double e, double power, double &hprime) {
  const double zero = 0.0, one = 1.0, two = 2.0, three = 3.0, beta = 0.95, alpha = 1.1, toler = 0.000001;
  const int max_itr = 20;
  if (eps_pls <= zero) {
    hprime = e * (0.9999 * e) / (e - 0.9999 * e);
    return sigma_o;
  }
  const double sig_o = sigma_o / beta;
  const double eps_o = sig_o / e;
  const double poweri = one / power;
  const double eps_a = beta * eps_o;
  const double eps_b = alpha * eps_o;
  const double h = eps_b - eps_a;
  const double h2 = h * h;
  const double h3 = h2 * h;
  const double sig_a = beta * sig_o;
  const double sig_b = sig_o * std::pow(alpha, poweri);
  const double epower = e / power;
  const double pm1 = poweri - one;
  const double et_b = epower * std::pow(alpha, pm1);
  int iterno = 1;
  double strain = eps_pls + eps_o;
  double stress = zero;
  double stress_new, fprime;
  for (;;) {
    if (strain >= eps_b) {
      stress_new = sig_o * std::pow(strain / eps_o, poweri);
      fprime = epower * std::pow(strain / eps_o, pm1);
    } else {
      const double ta = strain - eps_a;
      const double tb = strain - eps_b;
      const double n1 = -tb * tb * (-h - two * ta) / h3;
      const double n2 = ta * tb * tb / h2;
      const double n3 = ta * ta * (h - two * tb) / h3;
      const double n4 = ta * ta * tb / h2;
      stress_new = n1 * sig_a + n3 * sig_b + n2 * e + n4 * et_b;
      const double pn1 = two * (tb) * (three * strain - two * eps_a - eps_b + h) / h3;
      const double pn3 = two * (ta) * (-three * strain + eps_a + two * eps_b + h) / h3;
      const double pn2 = -tb * (-three * strain + two * eps_a + eps_b) / h2;
      const double pn4 = -ta * (-three * strain + eps_a + two * eps_b) / h2;
      fprime = pn1 * sig_a + pn3 * sig_b + pn2 * e + pn4 * et_b;
    }
    const double resid = strain - stress_new / e - eps_pls;
    const double deps = -resid / (one - fprime / e);
    const double strain_new = strain + deps;
    const bool converge = std::fabs(strain_new - strain) < toler * strain_new &&
                          std::fabs(stress_new - stress) < toler * stress_new;
    if (converge) break;
    iterno = iterno + 1;
    if (iterno > max_itr) {
      hprime = e * e;
      return sig_o;
    }
    strain = strain_new;
    stress = stress_new;
  }
  const double et = fprime;
  hprime = (e * et) / (e - et);
  return stress_new;
}

// the per-element property vectors of setup_mm03_rknstr (rknstr.f:1142-1215) at one element
struct Props {
  double e, nu, sigyld /* already reduced by trans_factor when n_power > 0 */, h_fixed, n_power,
      m_power, eps_ref, f0, q1, q2, q3;
  bool nucleation;
};

// mm03p_a (mm03.f:304-361): history at n on step 1
inline void mm03p_a(const Props &P, const double *stress_n, double *h /*15 words, 0-based*/) {
  const double zero = 0.0, one = 1.0, two = 2.0;
  double hp;
  if (P.n_power > zero) hp = P.e * (0.9999 * P.e) / (P.e - 0.9999 * P.e);
  else hp = P.h_fixed;
  h[3] = hp;
  h[0] = zero;
  h[1] = P.sigyld;
  h[2] = zero;
  h[4] = P.f0;
  h[5] = state_word(-1);
  h[6] = zero; h[7] = zero; h[8] = zero;
  for (int k = 9; k < 15; ++k) h[k] = zero;
  const double e_n = P.e;                // ym_n: no temperature dependence here
  if (e_n <= zero) return;
  const double nu_n = P.nu;
  const double g_n = e_n / two / (one + nu_n);
  h[9] = (stress_n[0] - nu_n * (stress_n[1] + stress_n[2])) / e_n;
  h[10] = (stress_n[1] - nu_n * (stress_n[0] + stress_n[2])) / e_n;
  h[11] = (stress_n[2] - nu_n * (stress_n[0] + stress_n[1])) / e_n;
  h[12] = stress_n[3] / g_n;
  h[13] = stress_n[4] / g_n;
  h[14] = stress_n[5] / g_n;
}

// one element's row of the block arrays mm03p / mm03 / cnst3 exchange
struct Point {
  double deps[6];
  double stress_n[9], stress_n1[9], stress_n1_elas[6];
  double history[15], history1[15];
  double p_trial, q_trial, yld_func;
  bool null_point, nonlinear_flag;
};

// mm03p (mm03.f:10-300) over the `span` rows of a block.  Returns process_block.
inline bool mm03p(int step, int span, Point *pt, const Props *P) {
  const double root22 = std::sqrt(2.0) / 2.0, third = 1.0 / 3.0, zero = 0.0, one = 1.0, two = 2.0,
               six = 6.0, toler = 1.0e-10, half = 0.5, chk_yld_eps = 0.002;
  if (step == 1)
    for (int i = 0; i < span; ++i) mm03p_a(P[i], pt[i].stress_n, pt[i].history);
  bool process_block = false;
  for (int i = 0; i < span; ++i) {
    const double *d = pt[i].deps;
    pt[i].null_point = std::fabs(d[0]) + std::fabs(d[1]) + std::fabs(d[2]) + std::fabs(d[3]) +
                       std::fabs(d[4]) + std::fabs(d[5]) <= toler * chk_yld_eps;
    process_block = process_block || !pt[i].null_point;
    pt[i].stress_n1[6] = zero; pt[i].stress_n1[7] = zero; pt[i].stress_n1[8] = zero;
  }
  if (!process_block) {                                   // :63-79 all points null
    for (int i = 0; i < span; ++i) {
      for (int k = 0; k < 9; ++k) pt[i].stress_n1[k] = pt[i].stress_n[k];
      for (int k = 0; k < 6; ++k) pt[i].stress_n1_elas[k] = zero;
      for (int k = 0; k < 5; ++k) pt[i].history1[k] = pt[i].history[k];
      pt[i].history1[5] = state_word(-1);
      pt[i].history1[6] = zero; pt[i].history1[7] = zero; pt[i].history1[8] = zero;
      for (int k = 9; k < 15; ++k) pt[i].history1[k] = pt[i].history[k];
    }
    return false;
  }
  for (int i = 0; i < span; ++i) {                        // 1. trial elastic stress (:86-108)
    Point &p = pt[i];
    double trial_eps[6];
    for (int k = 0; k < 6; ++k) trial_eps[k] = p.history[9 + k] + p.deps[k];
    const double e = P[i].e, nu = P[i].nu;
    const double c = e / ((one + nu) * (one - two * nu));
    const double a = c * (one - nu);
    const double b = c * nu;
    const double shear_mod = e / (two * (one + nu));
    p.stress_n1[0] = a * trial_eps[0] + b * (trial_eps[1] + trial_eps[2]);
    p.stress_n1[1] = a * trial_eps[1] + b * (trial_eps[0] + trial_eps[2]);
    p.stress_n1[2] = a * trial_eps[2] + b * (trial_eps[0] + trial_eps[1]);
    p.stress_n1[3] = shear_mod * trial_eps[3];
    p.stress_n1[4] = shear_mod * trial_eps[4];
    p.stress_n1[5] = shear_mod * trial_eps[5];
    for (int k = 0; k < 6; ++k) p.stress_n1_elas[k] = p.stress_n1[k];
  }
  for (int i = 0; i < span; ++i) {                        // 2. p, q of the trial state (:120-131)
    Point &p = pt[i];
    const double *s = p.stress_n1;
    p.p_trial = -(s[0] + s[1] + s[2]) * third;
    const double a = s[0] - s[2];
    const double b = s[0] - s[1];
    const double c = s[1] - s[2];
    const double d = s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
    p.q_trial = root22 * std::sqrt(a * a + b * b + c * c + six * d);
  }
  for (int i = 0; i < span; ++i) {                        // 3.-5. (:147-196), mises branch
    Point &p = pt[i];
    const double flow_stress = p.history[1];
    p.yld_func = p.q_trial - flow_stress;
    p.nonlinear_flag = (p.history[0] > zero) || (p.yld_func > zero);
  }
  for (int i = 0; i < span; ++i) {                        // 6. still-linear points (:211-228)
    Point &p = pt[i];
    if (p.nonlinear_flag) continue;
    p.stress_n1[6] = p.stress_n[6] +
                     half * (p.deps[0] * (p.stress_n1[0] + p.stress_n[0]) +
                             p.deps[1] * (p.stress_n1[1] + p.stress_n[1]) +
                             p.deps[2] * (p.stress_n1[2] + p.stress_n[2]) +
                             p.deps[3] * (p.stress_n1[3] + p.stress_n[3]) +
                             p.deps[4] * (p.stress_n1[4] + p.stress_n[4]) +
                             p.deps[5] * (p.stress_n1[5] + p.stress_n[5]));
    p.stress_n1[7] = p.stress_n[7];
    p.stress_n1[8] = p.stress_n[8];
    for (int k = 0; k < 6; ++k) p.history1[9 + k] = p.history[9 + k] + p.deps[k];
  }
  for (int i = 0; i < span; ++i)                          // 7. (:243-246)
    for (int j = 0; j < 9; ++j) pt[i].history1[j] = pt[i].history[j];
  return true;
}

// the host-associated locals of mm03 that mm03q / mm03qq share (mm03.f:2099-2135)
struct Shared {
  double e, nu, eps_ref, sigma_o, m_power, n_power, sbar, ebarp, ebarp_new, p_new, q_new, shear_mod,
      h, q_trial, deq, dep, h_new, f_new, time_incr, newyld, threeg, mpoweri, h_constant, sbar_new,
      eps_o, h_inviscid, inviscid_stress, hprime_init, sig_cur_min_val;
  double zero = 0.0, half = 0.5, one = 1.0, two = 2.0, three = 3.0,
         third = W3DO3_K(0.3333333333), six = 6.0;       // data statement of bare literals (:2130)
  bool rate_depen, consth, power_law;
  int vm_converge;

  // mm03qq (mm03.f:3113-3160); caseh only matters for segmental curves
  void mm03qq(double deplas, double &sig_bar, double &resid_now) {
    if (consth) {
      inviscid_stress = sigma_o + h_constant * (ebarp + deplas);
      h_inviscid = h_constant;
    } else {       // power_law
      inviscid_stress = mm03is(ebarp + deplas, sigma_o, e, n_power, h_inviscid);
    }
    if (rate_depen) {
      const double factor = eps_ref * deplas / time_incr + one;
      sig_bar = inviscid_stress * std::pow(factor, mpoweri);
      resid_now = q_trial - threeg * deplas - sig_bar;
    } else {
      sig_bar = inviscid_stress;
      resid_now = q_trial - threeg * deplas - sig_bar;
    }
  }

  // mm03q (mm03.f:2820-3100)
  void mm03q(double *newsig) {
    const double toler = W3DO3_K(0.000001), sig_tol = W3DO3_K(0.000001), unused = -W3DO3_K(1.11e30);
    double newyld_high, newyld_mid, deplas_low, deplas_high, resid_high, fl, fh, deplas = 0.0,
           deplas_riddr, deplas_mid, resid_mid, sridder, deplas_new, resid_new, resid, resid_low;
    int iconverge = 0;
    vm_converge = 0;
    const double eps_compare = toler * std::max(ebarp, eps_o);
    deplas_low = zero;
    resid_low = q_trial - sbar;
    if (hprime_init >= zero) deplas_high = (q_trial - sig_cur_min_val) / threeg;
    else deplas_high = (q_trial - sig_cur_min_val) / (threeg + hprime_init);
    mm03qq(deplas_high, newyld_high, resid_high);
    fl = resid_low;
    fh = resid_high;
    bool bracketed = (fl > zero && fh < zero) || (fl < zero && fh > zero);
    if (!bracketed) {
      if (std::fabs(fl) <= sig_tol * sbar) {
        newyld = sbar;
        deplas = zero;
        iconverge = 1;
      } else if (std::fabs(fh) <= sig_tol * sbar) {
        deplas = deplas_high;
        newyld = newyld_high;
        iconverge = 2;
      } else {
        vm_converge = 1;
        return;
      }
    } else {
      deplas_riddr = unused;
      int loop;
      for (loop = 1; loop <= 20; ++loop) {
        deplas_mid = half * (deplas_low + deplas_high);
        mm03qq(deplas_mid, newyld_mid, resid_mid);
        sridder = std::sqrt(resid_mid * resid_mid - resid_low * resid_high);
        if (sridder == zero) {
          vm_converge = 1;
          return;
        }
        deplas_new = deplas_mid + (deplas_mid - deplas_low) *
                                      (std::copysign(one, resid_low - resid_high) * resid_mid / sridder);
        if (std::fabs(deplas_new - deplas_riddr) <= eps_compare) {
          iconverge = 3;
          deplas = deplas_new;
          break;
        }
        deplas_riddr = deplas_new;
        mm03qq(deplas_riddr, newyld, resid_new);
        if (std::fabs(resid_new) <= sig_tol * newyld) {
          iconverge = 4;
          deplas = deplas_new;
          break;
        }
        if (std::copysign(std::fabs(resid_mid), resid_new) != resid_mid) {
          deplas_low = deplas_mid;
          resid_low = resid_mid;
          deplas_high = deplas_riddr;
          resid_high = resid_new;
        } else if (std::copysign(std::fabs(resid_low), resid_new) != resid_low) {
          deplas_high = deplas_riddr;
          resid_high = resid_new;
        } else if (std::copysign(std::fabs(resid_high), resid_new) != resid_high) {
          deplas_low = deplas_riddr;
          resid_low = resid_new;
        } else {
          vm_converge = 1;
          return;
        }
        if (std::fabs(deplas_high - deplas_low) <= eps_compare) {
          iconverge = 5;
          deplas = (deplas_high + deplas_low) * half;
          break;
        }
      }
      if (loop > 20) {
        vm_converge = 2;
        return;
      }
    }
    // 1000: (:3031-3076)
    if (iconverge == 3 || iconverge == 5) mm03qq(deplas, newyld, resid);
    if (deplas <= zero) {
      if (std::fabs(deplas) > W3DO3_K(0.001) * eps_o) vm_converge = 3;
    }
    h = h_inviscid;
    if (rate_depen) {
      const double term0 = std::pow(newyld / inviscid_stress, one - m_power);
      const double term1 = term0 * eps_ref * inviscid_stress / m_power / time_incr;
      const double term2 = h_inviscid * newyld / inviscid_stress;
      h = term1 + term2;
    }
    h_new = h;
    ebarp_new = ebarp + deplas;
    sbar_new = newyld;
    f_new = zero;
    q_new = sbar_new;
    p_new = -(newsig[0] + newsig[1] + newsig[2]) * third;
    dep = zero;
    deq = deplas;
    const double scale = threeg * deplas / q_trial;
    const double sm = -p_new;
    newsig[0] = newsig[0] - scale * (newsig[0] - sm);
    newsig[1] = newsig[1] - scale * (newsig[1] - sm);
    newsig[2] = newsig[2] - scale * (newsig[2] - sm);
    newsig[3] = newsig[3] - scale * newsig[3];
    newsig[4] = newsig[4] - scale * newsig[4];
    newsig[5] = newsig[5] - scale * newsig[5];
  }
};

// mm03 (mm03.f:2063-2560), mises path.  Returns 0, or 1 = cut the step / abort (vm_converge 1, 2).
inline int mm03(const Props &P, double dtime, Point &p) {
  Shared S;
  S.e = P.e; S.nu = P.nu; S.eps_ref = P.eps_ref; S.sigma_o = P.sigyld; S.m_power = P.m_power;
  S.n_power = P.n_power;
  S.time_incr = dtime;
  S.q_trial = p.q_trial;
  S.power_law = true;                                   // rknstr.f:819 (not segmental)
  S.sig_cur_min_val = S.sigma_o;
  S.consth = S.n_power == S.zero;
  S.eps_o = S.sigma_o / S.e;
  S.shear_mod = S.half * S.e / (S.one + S.nu);
  S.threeg = S.three * S.shear_mod;
  S.hprime_init = S.zero;
  S.rate_depen = S.time_incr > S.zero && S.m_power > S.zero;
  S.h_constant = 0.0;
  if (S.n_power == S.zero) S.h_constant = P.h_fixed;
  S.mpoweri = 0.0;
  if (S.rate_depen) S.mpoweri = S.one / S.m_power;
  const double *stress_n = p.stress_n;
  double *stress_n1 = p.stress_n1;
  S.ebarp = p.history[0];
  S.sbar = p.history[1];
  S.h = p.history[3];
  const double dword = p.history[5];
  int32_t state = word_state(dword);
  double senerg = stress_n[6];
  const double f1 = p.yld_func;
  bool plastic_loading;
  if (state == -1) plastic_loading = f1 > S.zero;
  else plastic_loading = f1 >= S.zero;
  if (!plastic_loading) {                               // :2262-2296
    const double dsenrg = p.deps[0] * (stress_n[0] + stress_n1[0]) + p.deps[1] * (stress_n[1] + stress_n1[1]) +
                          p.deps[2] * (stress_n[2] + stress_n1[2]) + p.deps[3] * (stress_n[3] + stress_n1[3]) +
                          p.deps[4] * (stress_n[4] + stress_n1[4]) + p.deps[5] * (stress_n[5] + stress_n1[5]);
    senerg = senerg + dsenrg * S.half;
    stress_n1[6] = senerg;
    stress_n1[7] = stress_n[7];
    stress_n1[8] = stress_n[8];
    state = -1;
    int32_t iw[2];
    std::memcpy(iw, &dword, 8);
    iw[0] = state;
    std::memcpy(&p.history1[5], iw, 8);
    p.history1[6] = S.zero; p.history1[7] = S.zero; p.history1[8] = S.zero;
    for (int k = 0; k < 6; ++k) p.history1[9 + k] = p.history[9 + k] + p.deps[k];
    return 0;
  }
  S.mm03q(stress_n1);                                   // :2424
  if (S.vm_converge == 1 || S.vm_converge == 2) return 1;
  {
    const double dsenrg = p.deps[0] * (stress_n[0] + stress_n1[0]) + p.deps[1] * (stress_n[1] + stress_n1[1]) +
                          p.deps[2] * (stress_n[2] + stress_n1[2]) + p.deps[3] * (stress_n[3] + stress_n1[3]) +
                          p.deps[4] * (stress_n[4] + stress_n1[4]) + p.deps[5] * (stress_n[5] + stress_n1[5]);
    senerg = senerg + dsenrg * S.half;
  }
  stress_n1[6] = senerg;                                // :2458-2500
  state = 1;
  p.history1[0] = S.ebarp_new;
  p.history1[1] = S.sbar_new;
  p.history1[2] = S.p_new;
  p.history1[3] = S.h_new;
  p.history1[4] = S.f_new;
  {
    int32_t iw[2];
    std::memcpy(iw, &dword, 8);
    iw[0] = state;
    std::memcpy(&p.history1[5], iw, 8);
  }
  p.history1[6] = S.dep;
  p.history1[7] = S.deq;
  p.history1[8] = S.q_new;
  double dsig[6], deps_plas[6];
  for (int k = 0; k < 6; ++k) dsig[k] = stress_n1[k] - stress_n[k];
  deps_plas[0] = p.deps[0] - (dsig[0] - S.nu * (dsig[1] + dsig[2])) / S.e;
  deps_plas[1] = p.deps[1] - (dsig[1] - S.nu * (dsig[0] + dsig[2])) / S.e;
  deps_plas[2] = p.deps[2] - (dsig[2] - S.nu * (dsig[0] + dsig[1])) / S.e;
  deps_plas[3] = p.deps[3] - dsig[3] / S.shear_mod;
  deps_plas[4] = p.deps[4] - dsig[4] / S.shear_mod;
  deps_plas[5] = p.deps[5] - dsig[5] / S.shear_mod;
  for (int k = 0; k < 6; ++k) p.history1[9 + k] = p.history[9 + k] + p.deps[k] - deps_plas[k];
  stress_n1[7] = stress_n[7] + S.half * (S.sbar_new + S.sbar) * (S.ebarp_new - S.ebarp);
  stress_n1[8] = stress_n[8] + (S.ebarp_new - S.ebarp);
  return 0;
}

// cnst3 (mm03.f:3252-3790) at one point; cep(6,6) 0-based [row][col]
inline void cnst3(int iter, const Props &P, const double *stress_trial, const double *history,
                  const double *history1, double cep[6][6]) {
  const double twothd = 0.6666666667, third = 0.333333333, half = 0.5, root32 = 1.224744871, zero = 0.0,
               one = 1.0, two = 2.0, three = 3.0;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) cep[r][c] = zero;
  const int32_t state = word_state(history1[5]);
  const double e = P.e, nu = P.nu;
  if (state == -1) {
    const double c1 = (e / ((one + nu) * (one - two * nu)));
    const double c2 = (one - nu) * c1;
    const double c3 = ((one - two * nu) / two) * c1;
    const double c4 = nu * c1;
    cep[0][0] = c2; cep[1][1] = c2; cep[2][2] = c2; cep[3][3] = c3; cep[4][4] = c3; cep[5][5] = c3;
    cep[0][1] = c4; cep[0][2] = c4; cep[1][0] = c4; cep[2][0] = c4; cep[1][2] = c4; cep[2][1] = c4;
    return;
  }
  if (state != 1) return;
  const double ebarp = history1[0], sbar = history1[1], pn1 = history1[2], hprime = history1[3], f = history1[4];
  double dep = history1[6], deq = history1[7];
  const double qn1 = history1[8], ebarpn = history[0];
  const double shear_mod = half * e / (one + nu);
  const double bulk_mod = e * third / (one - two * nu);
  if (iter <= 1) { dep = zero; deq = zero; }
  const double twog = two * shear_mod, threeg = three * shear_mod;
  const double debarp = ebarp - ebarpn;
  const double sbar2 = sbar * sbar, sbar3 = sbar * sbar * sbar;
  const double p = pn1, q = qn1;
  const double sm = (stress_trial[0] + stress_trial[1] + stress_trial[2]) * third;
  const double sx = stress_trial[0] - sm, sy = stress_trial[1] - sm, sz = stress_trial[2] - sm;
  const double qe = root32 * std::sqrt(sx * sx + sy * sy + sz * sz +
                                       two * (stress_trial[3] * stress_trial[3] + stress_trial[4] * stress_trial[4] +
                                              stress_trial[5] * stress_trial[5]));
  const double sl = three / two / qe;
  double n[6];
  n[0] = sx * sl; n[1] = sy * sl; n[2] = sz * sl;
  n[3] = stress_trial[3] * sl; n[4] = stress_trial[4] * sl; n[5] = stress_trial[5] * sl;
  const double a1 = (deq * q - dep * p) / ((one - f) * (one - f)) / sbar;
  const double a2 = (deq * q - dep * p) * hprime / (f - one) / sbar2;
  const double a3 = one / sbar / (f - one);
  const double a4 = p + dep * bulk_mod;
  const double a5 = threeg * deq - q;
  const double anuc = zero, anuc_prime = zero;          // nucleation off
  const double b1 = (one - f) / (one + dep);
  const double b2 = (anuc + debarp * anuc_prime) / (one + dep);
  const double d1 = (a2 - one) / (a1 * b2 + a2 - one);
  const double d2 = -b2 / (a1 * b2 + a2 - one);
  const double d3 = one / (one - a2 - a1 * b2);
  const double d4 = a1 / (one - a2 - a1 * b2);
  const double h1 = d1 * b1 + d2 * a3 * a4;
  const double h2 = d2 * a3 * a5;
  const double h3 = a3 * dep * d2;
  const double h4 = a3 * deq * d2;
  const double h5 = a3 * a4 * d3 + b1 * d4;
  const double h6 = a3 * a5 * d3;
  const double h7 = a3 * dep * d3;
  const double h8 = a3 * deq * d3;
  const double beta = -(one + half) * P.q2;
  double ch = std::cosh(beta * p / sbar);
  double sh = std::sinh(beta * p / sbar);
  const double pgp = two * beta * f * P.q1 * sh / sbar;
  const double pgq = two * q / sbar2;
  const double pgsbar = (two / sbar3) * (-q * q - beta * f * p * P.q1 * sbar * sh);
  const double pgf = two * (-f * P.q3 + P.q1 * ch);
  const double cap_a11 = h1 * pgf + bulk_mod * pgp + hprime * h5 * pgsbar;
  const double cap_a12 = h2 * pgf - threeg * pgq + hprime * h6 * pgsbar;
  const double cap_b1 = (h3 * pgf + pgp + hprime * h7 * pgsbar) * bulk_mod;
  const double cap_c1 = (h4 * pgf - pgq + hprime * h8 * pgsbar) * twog;
  const double c5 = two / sbar2;
  const double c6 = two * two * q * hprime / sbar3;
  const double e10 = c6 * h7;
  const double e11 = c5 + c6 * h8;
  const double e12 = c6 * h5;
  const double e13 = threeg * c5 + c6 * h6;
  const double m1 = two * P.q1;
  const double m2 = -three * P.q2 * half;
  sh = std::sinh(m2 * p / sbar);
  ch = std::cosh(m2 * p / sbar);
  const double c7 = f * m1 * m2 * m2 * ch / sbar2;
  const double c8 = m1 * m2 * sh / sbar;
  const double c9 = -f * hprime * (m1 * m2 * m2 * p * ch / sbar3 + m1 * m2 * sh / sbar2);
  const double d10 = -c7 - c8 * h3 - c9 * h7;
  const double d11 = -c9 * h8 - c8 * h4;
  const double d12 = c8 * h1 + c9 * h5 + c7 * bulk_mod;
  const double d13 = c8 * h2 + c9 * h6;
  const double cap_a21 = deq * d12 - dep * e12 + pgq;
  const double cap_a22 = deq * d13 - dep * e13 + pgp;
  const double cap_b2 = -(deq * d10 + dep * e10) * bulk_mod;
  const double cap_c2 = -(deq * d11 + dep * e11) * twog;
  const double denom = cap_a11 * cap_a22 - cap_a12 * cap_a21;
  const double h10 = cap_a22 / denom;
  const double h11 = cap_a12 / denom;
  const double h12 = cap_a11 / denom;
  const double h13 = cap_a21 / denom;
  const double mpi = h10 * cap_b1 - h11 * cap_b2;
  const double mpn = h10 * cap_c1 - h11 * cap_c2;
  const double mqi = h12 * cap_b2 - h13 * cap_b1;
  const double mqn = h12 * cap_c2 - h13 * cap_c1;
  double con_1 = twog * (one - threeg * deq / qe);
  const double con_2 = bulk_mod * (one - mpi);
  const double con_3 = half * (twog * mqi + bulk_mod * mpn);
  cep[0][0] = con_1 * twothd + con_2 - two * con_3 * n[0];
  cep[1][0] = -con_1 * third + con_2 - con_3 * (n[1] + n[0]);
  cep[2][0] = -con_1 * third + con_2 - con_3 * (n[2] + n[0]);
  cep[0][1] = -con_1 * third + con_2 - con_3 * (n[0] + n[1]);
  cep[1][1] = con_1 * twothd + con_2 - two * con_3 * n[1];
  cep[2][1] = -con_1 * third + con_2 - con_3 * (n[1] + n[2]);
  cep[0][2] = -con_1 * third + con_2 - con_3 * (n[0] + n[2]);
  cep[1][2] = -con_1 * third + con_2 - con_3 * (n[1] + n[2]);
  cep[2][2] = con_1 * twothd + con_2 - two * con_3 * n[2];
  cep[3][3] = con_1 * half; cep[4][4] = con_1 * half; cep[5][5] = con_1 * half;
  for (int r = 3; r < 6; ++r)
    for (int c = 0; c < 3; ++c) {
      cep[r][c] = -con_3 * n[r];
      cep[c][r] = -con_3 * n[r];
    }
  con_1 = twog * (twog * deq / qe - mqn);
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < 6; ++r) cep[r][c] = (cep[r][c] + con_1 * n[r] * n[c]);
}

}  // namespace w3do_mm03
#endif
