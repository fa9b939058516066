// w3d_mm03_core.h -- material model 3 (mises flow theory, constant or linear + power-law hardening,
// optional power-law rate sensitivity) for ONE Gauss point: the engine's device core, also
// compiled for the host by tests/native/mm03_point_check.cpp.  The parity tests hold it against
// an independent CPU restatement of mm03.f kept with the test infrastructure (no shared file).
//
// Follows reference src/mm03.f: mm03p :10 (trial state, yield check, step-1 history :304), mm03
// :2063 (elastic / unloading :2262, plastic :2424-2500), mm03q :2820 (Ridder's method for the
// plastic strain increment), mm03qq :3113, mm03is :963 (uniaxial curve, Newton on the strain),
// cnst3 :3252 (consistent tangent); driver src/rstgp1.f:1418-1990; properties
// src/rknstr.f:1142-1215.  The Gurson branch (f0 > 0 or nucleation), segmental curves,
// temperatures and fgm properties are refused by w3d_engine_create.
//
// Kept on purpose (the reference's numbers are the bar):
//  * mm03's mean-stress factor is the bare literal 0.3333333333 (mm03.f:2131), mm03q's tolerances
//    and its -1.11e30 start value are bare literals too (:2843): evaluated in double, as the
//    reference's -fpconstant build does (W3D_MM03_DATA_SINGLE: in single); cnst3's are
//    0.333333333d0 / 0.6666666667d0 / 1.224744871d0 (:3291-3292);
//  * history word 6 carries the state (+1 plastic, -1 elastic) in its low 32 bits (:2124);
//  * a block whose points are ALL null at this Gauss point (sum |deps| <= 2e-13) copies the
//    state at n and marks every point elastic (:63-79): `block_null`, decided by the caller over
//    the element block; a null point in a live block that is already plastic is skipped by the
//    driver (rstgp1.f:1937): energy words zero, elastic strain words written as zero (the
//    reference leaves them undefined).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define W3D3_HD __host__ __device__
#else
#define W3D3_HD
#endif

namespace w3d_mm03 {

#ifdef W3D_MM03_DATA_SINGLE
#define W3D3_K(x) ((double)(x##f))
#else
#define W3D3_K(x) (x)
#endif

constexpr int HIST = 15;
enum { OK = 0, NO_CONVERGENCE = 1 };

struct Props {
  double e, nu;        // props(7), props(8)
  double sigyld;       // props(23), times 0.95 when n_power > 0 (rstgp1.f:1898-1906)
  double h_fixed;      // props(15)
  double n_power, m_power, eps_ref;   // props(21), props(20), props(22)
  double q1, q2, q3;   // props(27:29): zero for a `mises` card, kept in the tangent expressions
};

W3D3_HD inline double flag_word(int st) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(0, st);
#else
  int iw[2] = {st, 0};
  double d;
  memcpy(&d, iw, 8);
  return d;
#endif
}
W3D3_HD inline int word_flag(double d) {
#if defined(__CUDA_ARCH__)
  return __double2loint(d);
#else
  int iw[2];
  memcpy(iw, &d, 8);
  return iw[0];
#endif
}
W3D3_HD inline double with_flag(double d, int st) {      // iword(1) = st, iword(2) untouched
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(d), st);
#else
  int iw[2];
  memcpy(iw, &d, 8);
  iw[0] = st;
  memcpy(&d, iw, 8);
  return d;
#endif
}

// uniaxial stress and plastic modulus at a total plastic strain (mm03is)
W3D3_HD inline double power_law_stress(double eps_pls, double sigma_o, double e, double power, double &hprime) {
  const double beta = 0.95, alpha = 1.1, toler = 0.000001;
  if (eps_pls <= 0.0) {
    hprime = e * (0.9999 * e) / (e - 0.9999 * e);
    return sigma_o;
  }
  const double sig_o = sigma_o / beta, eps_o = sig_o / e, poweri = 1.0 / power;
  const double eps_a = beta * eps_o, eps_b = alpha * eps_o;
  const double h = eps_b - eps_a, h2 = h * h, h3 = h2 * h;
  const double sig_a = beta * sig_o, sig_b = sig_o * pow(alpha, poweri);
  const double epower = e / power, pm1 = poweri - 1.0;
  const double et_b = epower * pow(alpha, pm1);
  double strain = eps_pls + eps_o, stress = 0.0, stress_new = 0.0, fprime = 0.0;
  for (int it = 1;; ++it) {
    if (strain >= eps_b) {
      const double r = strain / eps_o;
      stress_new = sig_o * pow(r, poweri);
      fprime = epower * pow(r, pm1);
    } else {      // cubic (Hermite) transition between the linear and the power-law parts
      const double ta = strain - eps_a, tb = strain - eps_b;
      const double n1 = -tb * tb * (-h - 2.0 * ta) / h3;
      const double n2 = ta * tb * tb / h2;
      const double n3 = ta * ta * (h - 2.0 * tb) / h3;
      const double n4 = ta * ta * tb / h2;
      stress_new = n1 * sig_a + n3 * sig_b + n2 * e + n4 * et_b;
      const double pn1 = 2.0 * (tb) * (3.0 * strain - 2.0 * eps_a - eps_b + h) / h3;
      const double pn3 = 2.0 * (ta) * (-3.0 * strain + eps_a + 2.0 * eps_b + h) / h3;
      const double pn2 = -tb * (-3.0 * strain + 2.0 * eps_a + eps_b) / h2;
      const double pn4 = -ta * (-3.0 * strain + eps_a + 2.0 * eps_b) / h2;
      fprime = pn1 * sig_a + pn3 * sig_b + pn2 * e + pn4 * et_b;
    }
    const double resid = strain - stress_new / e - eps_pls;
    const double strain_new = strain + (-resid / (1.0 - fprime / e));
    if (fabs(strain_new - strain) < toler * strain_new && fabs(stress_new - stress) < toler * stress_new) break;
    if (it + 1 > 20) {
      hprime = e * e;
      return sig_o;
    }
    strain = strain_new;
    stress = stress_new;
  }
  hprime = (e * fprime) / (e - fprime);
  return stress_new;
}

// the scalar equation of the radial return: q_trial - 3G d - sigma_bar(ebarp + d, d / dt) = 0
struct Return {
  double e, sigma_o, n_power, m_power, eps_ref, dt, ebarp, q_trial, threeg, mpoweri, h_constant;
  bool consth, rate;
  double inviscid, h_inviscid;      // of the last evaluation (mm03qq leaves them in the host's locals)
  W3D3_HD double resid(double d, double &sig_bar) {
    if (consth) {
      inviscid = sigma_o + h_constant * (ebarp + d);
      h_inviscid = h_constant;
    } else {
      inviscid = power_law_stress(ebarp + d, sigma_o, e, n_power, h_inviscid);
    }
    sig_bar = inviscid;
    if (rate) sig_bar = inviscid * pow(eps_ref * d / dt + 1.0, mpoweri);
    return q_trial - threeg * d - sig_bar;
  }
};

W3D3_HD inline void elastic_D(double e, double nu, double D[6][6]) {
  const double c1 = (e / ((1.0 + nu) * (1.0 - 2.0 * nu)));
  const double c2 = (1.0 - nu) * c1, c3 = ((1.0 - 2.0 * nu) / 2.0) * c1, c4 = nu * c1;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) D[r][c] = 0.0;
  D[0][0] = c2; D[1][1] = c2; D[2][2] = c2; D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
  D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
}

// consistent tangent at a plastic point (cnst3 with f = 0, nucleation off): lower triangle only
W3D3_HD inline void plastic_tangent(int iter, const Props &P, const double *trial, double ebarp_n,
                                    const double *h1, double D[6][6]) {
  const double twothd = 0.6666666667, third = 0.333333333, root32 = 1.224744871;
  const double e = P.e, nu = P.nu;
  const double ebarp = h1[0], sbar = h1[1], p = h1[2], hprime = h1[3], f = h1[4], q = h1[8];
  double dep = h1[6], deq = h1[7];
  if (iter <= 1) { dep = 0.0; deq = 0.0; }      // continuum tangent at the contact point
  const double shear_mod = 0.5 * e / (1.0 + nu), bulk_mod = e * third / (1.0 - 2.0 * nu);
  const double twog = 2.0 * shear_mod, threeg = 3.0 * shear_mod;
  const double debarp = ebarp - ebarp_n;
  const double sbar2 = sbar * sbar, sbar3 = sbar * sbar * sbar;
  const double sm = (trial[0] + trial[1] + trial[2]) * third;
  const double sx = trial[0] - sm, sy = trial[1] - sm, sz = trial[2] - sm;
  const double qe = root32 * sqrt(sx * sx + sy * sy + sz * sz +
                                  2.0 * (trial[3] * trial[3] + trial[4] * trial[4] + trial[5] * trial[5]));
  const double sl = 3.0 / 2.0 / qe;
  const double n[6] = {sx * sl, sy * sl, sz * sl, trial[3] * sl, trial[4] * sl, trial[5] * sl};
  const double a1 = (deq * q - dep * p) / ((1.0 - f) * (1.0 - f)) / sbar;
  const double a2 = (deq * q - dep * p) * hprime / (f - 1.0) / sbar2;
  const double a3 = 1.0 / sbar / (f - 1.0);
  const double a4 = p + dep * bulk_mod;
  const double a5 = threeg * deq - q;
  const double b1 = (1.0 - f) / (1.0 + dep);
  const double b2 = (0.0 + debarp * 0.0) / (1.0 + dep);
  const double d1 = (a2 - 1.0) / (a1 * b2 + a2 - 1.0);
  const double d2 = -b2 / (a1 * b2 + a2 - 1.0);
  const double d3 = 1.0 / (1.0 - a2 - a1 * b2);
  const double d4 = a1 / (1.0 - a2 - a1 * b2);
  const double h1_ = d1 * b1 + d2 * a3 * a4, h2 = d2 * a3 * a5, h3 = a3 * dep * d2, h4 = a3 * deq * d2;
  const double h5 = a3 * a4 * d3 + b1 * d4, h6 = a3 * a5 * d3, h7 = a3 * dep * d3, h8 = a3 * deq * d3;
  const double beta = -(1.0 + 0.5) * P.q2;
  double ch = cosh(beta * p / sbar), sh = sinh(beta * p / sbar);
  const double pgp = 2.0 * beta * f * P.q1 * sh / sbar;
  const double pgq = 2.0 * q / sbar2;
  const double pgsbar = (2.0 / sbar3) * (-q * q - beta * f * p * P.q1 * sbar * sh);
  const double pgf = 2.0 * (-f * P.q3 + P.q1 * ch);
  const double A11 = h1_ * pgf + bulk_mod * pgp + hprime * h5 * pgsbar;
  const double A12 = h2 * pgf - threeg * pgq + hprime * h6 * pgsbar;
  const double B1 = (h3 * pgf + pgp + hprime * h7 * pgsbar) * bulk_mod;
  const double C1 = (h4 * pgf - pgq + hprime * h8 * pgsbar) * twog;
  const double c5 = 2.0 / sbar2, c6 = 2.0 * 2.0 * q * hprime / sbar3;
  const double e10 = c6 * h7, e11 = c5 + c6 * h8, e12 = c6 * h5, e13 = threeg * c5 + c6 * h6;
  const double m1 = 2.0 * P.q1, m2 = -3.0 * P.q2 * 0.5;
  sh = sinh(m2 * p / sbar);
  ch = cosh(m2 * p / sbar);
  const double c7 = f * m1 * m2 * m2 * ch / sbar2;
  const double c8 = m1 * m2 * sh / sbar;
  const double c9 = -f * hprime * (m1 * m2 * m2 * p * ch / sbar3 + m1 * m2 * sh / sbar2);
  const double d10 = -c7 - c8 * h3 - c9 * h7, d11 = -c9 * h8 - c8 * h4;
  const double d12 = c8 * h1_ + c9 * h5 + c7 * bulk_mod, d13 = c8 * h2 + c9 * h6;
  const double A21 = deq * d12 - dep * e12 + pgq;
  const double A22 = deq * d13 - dep * e13 + pgp;
  const double B2 = -(deq * d10 + dep * e10) * bulk_mod;
  const double C2 = -(deq * d11 + dep * e11) * twog;
  const double denom = A11 * A22 - A12 * A21;
  const double h10 = A22 / denom, h11 = A12 / denom, h12 = A11 / denom, h13 = A21 / denom;
  const double mpi = h10 * B1 - h11 * B2, mpn = h10 * C1 - h11 * C2;
  const double mqi = h12 * B2 - h13 * B1, mqn = h12 * C2 - h13 * C1;
  const double k1 = twog * (1.0 - threeg * deq / qe);
  const double k2 = bulk_mod * (1.0 - mpi);
  const double k3 = 0.5 * (twog * mqi + bulk_mod * mpn);
  const double k4 = twog * (twog * deq / qe - mqn);
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) D[r][c] = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c <= r; ++c)
      D[r][c] = (r == c) ? k1 * twothd + k2 - 2.0 * k3 * n[r] : -k1 * third + k2 - k3 * (n[r] + n[c]);
  // the reference writes n(2)+n(1) for (2,1) but n(1)+n(2)-ordered sums elsewhere: addition is
  // commutative, so the lower triangle above is bit-identical to both
  for (int r = 3; r < 6; ++r) {
    D[r][r] = k1 * 0.5;
    for (int c = 0; c < 3; ++c) D[r][c] = -k3 * n[r];
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) D[r][c] = (D[r][c] + k4 * n[r] * n[c]);
  for (int r = 0; r < 6; ++r)
    for (int c = r + 1; c < 6; ++c) D[r][c] = D[c][r];
}

// One Gauss point, nonlinear update (iter >= 1 or extrapolated du).  hist_n: the 15 words at n
// (ignored on step 1, where the step-1 set-up of mm03p_a applies).  Outputs: cgn1 (9), h1 (15),
// D.  `keep_h1_strain` = false when the caller must leave nothing defined ... (always written).
W3D3_HD inline int update(int step, int iter, const Props &P, double dt, bool block_null, const double *cgn,
                          const double *uddt, const double *hist_n, double *cgn1, double *h1, double D[6][6]) {
  double h[HIST];
  const double e = P.e, nu = P.nu;
  if (step == 1) {      // mm03p_a
    h[0] = 0.0; h[1] = P.sigyld; h[2] = 0.0;
    h[3] = (P.n_power > 0.0) ? e * (0.9999 * e) / (e - 0.9999 * e) : P.h_fixed;
    h[4] = 0.0;                      // f0
    h[5] = flag_word(-1);
    h[6] = 0.0; h[7] = 0.0; h[8] = 0.0;
    for (int k = 9; k < HIST; ++k) h[k] = 0.0;
    if (e > 0.0) {                   // elastic strains of user initial stresses
      const double g_n = e / 2.0 / (1.0 + nu);
      h[9] = (cgn[0] - nu * (cgn[1] + cgn[2])) / e;
      h[10] = (cgn[1] - nu * (cgn[0] + cgn[2])) / e;
      h[11] = (cgn[2] - nu * (cgn[0] + cgn[1])) / e;
      h[12] = cgn[3] / g_n; h[13] = cgn[4] / g_n; h[14] = cgn[5] / g_n;
    }
  } else {
    for (int k = 0; k < HIST; ++k) h[k] = hist_n[k];
  }
  if (block_null) {                  // mm03.f:63-79
    for (int k = 0; k < 9; ++k) cgn1[k] = cgn[k];
    for (int k = 0; k < 5; ++k) h1[k] = h[k];
    h1[5] = flag_word(-1);
    h1[6] = 0.0; h1[7] = 0.0; h1[8] = 0.0;
    for (int k = 9; k < HIST; ++k) h1[k] = h[k];
    elastic_D(e, nu, D);
    return OK;
  }
  const bool null_point =
      fabs(uddt[0]) + fabs(uddt[1]) + fabs(uddt[2]) + fabs(uddt[3]) + fabs(uddt[4]) + fabs(uddt[5]) <= 1.0e-10 * 0.002;
  // trial elastic stress from the elastic strain at n plus the increment
  double trial[6];
  {
    double te[6];
    for (int k = 0; k < 6; ++k) te[k] = h[9 + k] + uddt[k];
    const double c = e / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double a = c * (1.0 - nu), b = c * nu, g = e / (2.0 * (1.0 + nu));
    trial[0] = a * te[0] + b * (te[1] + te[2]);
    trial[1] = a * te[1] + b * (te[0] + te[2]);
    trial[2] = a * te[2] + b * (te[0] + te[1]);
    trial[3] = g * te[3]; trial[4] = g * te[4]; trial[5] = g * te[5];
  }
  double q_trial;
  {
    const double a = trial[0] - trial[2], b = trial[0] - trial[1], c = trial[1] - trial[2];
    const double d = trial[3] * trial[3] + trial[4] * trial[4] + trial[5] * trial[5];
    q_trial = (sqrt(2.0) / 2.0) * sqrt(a * a + b * b + c * c + 6.0 * d);
  }
  const double yld = q_trial - h[1];
  const bool nonlinear = (h[0] > 0.0) || (yld > 0.0);
  for (int k = 0; k < 6; ++k) cgn1[k] = trial[k];
  cgn1[6] = 0.0; cgn1[7] = 0.0; cgn1[8] = 0.0;
  for (int k = 0; k < 9; ++k) h1[k] = h[k];
  const int state_n = word_flag(h[5]);
  const bool loading = nonlinear && ((state_n == -1) ? (yld > 0.0) : (yld >= 0.0));
  if (nonlinear && null_point) {
    // skipped by the driver loop (rstgp1.f:1937): see the header
    for (int k = 9; k < HIST; ++k) h1[k] = 0.0;
  } else if (!loading) {
    // still linear (mm03p step 6) or elastic unloading (mm03 :2262-2296): same energy update
    const double ds = uddt[0] * (cgn[0] + cgn1[0]) + uddt[1] * (cgn[1] + cgn1[1]) + uddt[2] * (cgn[2] + cgn1[2]) +
                      uddt[3] * (cgn[3] + cgn1[3]) + uddt[4] * (cgn[4] + cgn1[4]) + uddt[5] * (cgn[5] + cgn1[5]);
    if (nonlinear) {
      cgn1[6] = cgn[6] + ds * 0.5;
      h1[5] = with_flag(h[5], -1);
      h1[6] = 0.0; h1[7] = 0.0; h1[8] = 0.0;
    } else {
      // mm03p writes deps*(stress_n1 + stress_n) and half * ( sum )
      const double ds2 = uddt[0] * (cgn1[0] + cgn[0]) + uddt[1] * (cgn1[1] + cgn[1]) + uddt[2] * (cgn1[2] + cgn[2]) +
                         uddt[3] * (cgn1[3] + cgn[3]) + uddt[4] * (cgn1[4] + cgn[4]) + uddt[5] * (cgn1[5] + cgn[5]);
      cgn1[6] = cgn[6] + 0.5 * ds2;
    }
    cgn1[7] = cgn[7];
    cgn1[8] = cgn[8];
    for (int k = 0; k < 6; ++k) h1[9 + k] = h[9 + k] + uddt[k];
  } else {
    // ---- plastic loading: Ridder's method on the plastic strain increment (mm03q)
    const double third_r4 = W3D3_K(0.3333333333);
    const double toler = W3D3_K(0.000001), sig_tol = W3D3_K(0.000001);
    const double shear_mod = 0.5 * e / (1.0 + nu);
    Return R;
    R.e = e; R.sigma_o = P.sigyld; R.n_power = P.n_power; R.m_power = P.m_power; R.eps_ref = P.eps_ref;
    R.dt = dt; R.ebarp = h[0]; R.q_trial = q_trial; R.threeg = 3.0 * shear_mod;
    R.consth = P.n_power == 0.0;
    R.h_constant = R.consth ? P.h_fixed : 0.0;
    R.rate = dt > 0.0 && P.m_power > 0.0;
    R.mpoweri = R.rate ? 1.0 / P.m_power : 0.0;
    R.inviscid = 0.0; R.h_inviscid = 0.0;
    const double sbar = h[1], ebarp = h[0], eps_o = P.sigyld / e;
    const double eps_compare = toler * fmax(ebarp, eps_o);
    double lo = 0.0, f_lo = q_trial - sbar;
    double hi = (q_trial - P.sigyld) / R.threeg;      // hprime_init = 0, sig_cur_min_val = sigma_o
    double yld_hi, f_hi = R.resid(hi, yld_hi);
    double d = 0.0, newyld = 0.0;
    bool refresh = false;                              // converged on the bracket: re-evaluate at d
    if (!((f_lo > 0.0 && f_hi < 0.0) || (f_lo < 0.0 && f_hi > 0.0))) {
      if (fabs(f_lo) <= sig_tol * sbar) { newyld = sbar; d = 0.0; }
      else if (fabs(f_hi) <= sig_tol * sbar) { d = hi; newyld = yld_hi; }
      else return NO_CONVERGENCE;
    } else {
      double last = -W3D3_K(1.11e30);
      bool done = false;
      for (int loop = 1; loop <= 20 && !done; ++loop) {
        const double mid = 0.5 * (lo + hi);
        double yld_mid;
        const double f_mid = R.resid(mid, yld_mid);
        const double s = sqrt(f_mid * f_mid - f_lo * f_hi);
        if (s == 0.0) return NO_CONVERGENCE;
        const double nw = mid + (mid - lo) * (copysign(1.0, f_lo - f_hi) * f_mid / s);
        if (fabs(nw - last) <= eps_compare) { d = nw; refresh = true; done = true; break; }
        last = nw;
        const double f_new = R.resid(last, newyld);
        if (fabs(f_new) <= sig_tol * newyld) { d = nw; done = true; break; }
        if (copysign(fabs(f_mid), f_new) != f_mid) { lo = mid; f_lo = f_mid; hi = last; f_hi = f_new; }
        else if (copysign(fabs(f_lo), f_new) != f_lo) { hi = last; f_hi = f_new; }
        else if (copysign(fabs(f_hi), f_new) != f_hi) { lo = last; f_lo = f_new; }
        else return NO_CONVERGENCE;
        if (fabs(hi - lo) <= eps_compare) { d = (hi + lo) * 0.5; refresh = true; done = true; break; }
      }
      if (!done) return NO_CONVERGENCE;
    }
    if (refresh) { double r_ = R.resid(d, newyld); (void)r_; }
    double hnew = R.h_inviscid;
    if (R.rate) {
      const double t0 = pow(newyld / R.inviscid, 1.0 - P.m_power);
      const double t1 = t0 * P.eps_ref * R.inviscid / P.m_power / dt;
      const double t2 = R.h_inviscid * newyld / R.inviscid;
      hnew = t1 + t2;
    }
    const double ebarp_new = ebarp + d;
    const double p_new = -(cgn1[0] + cgn1[1] + cgn1[2]) * third_r4;
    const double scale = R.threeg * d / q_trial, sm = -p_new;
    cgn1[0] = cgn1[0] - scale * (cgn1[0] - sm);
    cgn1[1] = cgn1[1] - scale * (cgn1[1] - sm);
    cgn1[2] = cgn1[2] - scale * (cgn1[2] - sm);
    cgn1[3] = cgn1[3] - scale * cgn1[3];
    cgn1[4] = cgn1[4] - scale * cgn1[4];
    cgn1[5] = cgn1[5] - scale * cgn1[5];
    const double ds = uddt[0] * (cgn[0] + cgn1[0]) + uddt[1] * (cgn[1] + cgn1[1]) + uddt[2] * (cgn[2] + cgn1[2]) +
                      uddt[3] * (cgn[3] + cgn1[3]) + uddt[4] * (cgn[4] + cgn1[4]) + uddt[5] * (cgn[5] + cgn1[5]);
    cgn1[6] = cgn[6] + ds * 0.5;
    h1[0] = ebarp_new; h1[1] = newyld; h1[2] = p_new; h1[3] = hnew; h1[4] = 0.0;
    h1[5] = with_flag(h[5], 1);
    h1[6] = 0.0; h1[7] = d; h1[8] = newyld;
    double dsig[6], dpl[6];
    for (int k = 0; k < 6; ++k) dsig[k] = cgn1[k] - cgn[k];
    dpl[0] = uddt[0] - (dsig[0] - nu * (dsig[1] + dsig[2])) / e;
    dpl[1] = uddt[1] - (dsig[1] - nu * (dsig[0] + dsig[2])) / e;
    dpl[2] = uddt[2] - (dsig[2] - nu * (dsig[0] + dsig[1])) / e;
    dpl[3] = uddt[3] - dsig[3] / shear_mod;
    dpl[4] = uddt[4] - dsig[4] / shear_mod;
    dpl[5] = uddt[5] - dsig[5] / shear_mod;
    for (int k = 0; k < 6; ++k) h1[9 + k] = h[9 + k] + uddt[k] - dpl[k];
    cgn1[7] = cgn[7] + 0.5 * (newyld + sbar) * (ebarp_new - ebarp);
    cgn1[8] = cgn[8] + (ebarp_new - ebarp);
  }
  // ---- tangent (cnst3) from the state flag now in h1
  const int st1 = word_flag(h1[5]);
  if (st1 == -1) elastic_D(e, nu, D);
  else if (st1 == 1) plastic_tangent(iter, P, trial, h[0], h1, D);
  else
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) D[r][c] = 0.0;
  return OK;
}

}  // namespace w3d_mm03
