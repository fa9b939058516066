// w3d_mm01.cuh -- bilinear Mises plasticity (mm01) and its consistent tangent (cnst1) for one
// Gauss point, register resident.  Behaviour follows the reference's src/mm01.f:
//   mm01 :24-276 (driver + mm01_set_history :222-276), mm01_init :308-520, mm01_general :776-940,
//   mm01_sig_final :600-660, mm01_wrapup :524-598, cnst1 :1196-1348
// including the parity quirks listed in DESIGN.md (truncated constants, integer state stored in
// the low word of history(4), wbar reading history(1,11) of the block's first element).
#pragma once
#include "w3d_device.cuh"

namespace w3d {

struct Mm01Props {
  double e, nu, beta, hprime, yld;
  bool lnelas;
};

// history layout (mm01.f:222-276): 1 lambda*dt, 2 k, 3 ebarp, 4 state, 5 H', 6..11 backstress
// cgn/cgn1: unrotated Cauchy stress (1..6), 7 energy density, 8 plastic work density,
// 9 accumulated plastic strain.  hist_n_first11 = history(1,11) of the block's first element.
// Returns 0, or W3D_ERR_MM01_LAMBDA (5) when the reference would die_abort.
__device__ __forceinline__ int mm01_update(int step, const Mm01Props &pr, double (&cgn)[9],
                                           double (&cgn1)[9], const double (&deps)[6],
                                           double (&h)[11], double (&h1)[11], double (&rtse)[6],
                                           double hist_n_first11) {
  int fatal = 0;
  const double ym_n1 = pr.e, nu_n1 = pr.nu, ym_n = pr.e, nu_n = pr.nu, beta = pr.beta,
               hprime_n1 = pr.hprime, yld_n1 = pr.yld;
  const double zero = 0.0, one = 1.0, two = 2.0, three = 3.0;
  if (step == 1) {  // mm01_set_history
    const double root3 = sqrt(3.0);
    const double kn = yld_n1 / root3;
    h[0] = zero; h[1] = kn; h[2] = zero; h[3] = state_to_dword(3); h[4] = hprime_n1;
#pragma unroll
    for (int k = 5; k < 11; ++k) h[k] = zero;
#pragma unroll
    for (int k = 0; k < 11; ++k) h1[k] = h[k];
    cgn[7] = zero; cgn[8] = zero;
    hist_n_first11 = zero;
  }
  // ---- mm01_init
  double alpha_n[6], devstr_n1[6];
  double mrts, kbar, lk, shear_mod_n1, eps_vol_n1;
  bool yield = false;
  int instat = 3;
  {
    const double htol = 0.000001, root3 = sqrt(3.0), root2 = sqrt(2.0), yld_tol = 0.0000001;
    const double deps_vol = deps[0] + deps[1] + deps[2];
    const double eps_mean = deps_vol * (one / three);
    const double de1 = deps[0] - eps_mean, de2 = deps[1] - eps_mean, de3 = deps[2] - eps_mean;
    const double de4 = deps[3], de5 = deps[4], de6 = deps[5];
    // one reciprocal each for 1/E and 1/G instead of six IEEE divisions (<= 1 ulp difference)
    const double iym = one / ym_n;
    const double ig = (two * (one + nu_n)) * iym;
    const double een1 = (cgn[0] - nu_n * (cgn[1] + cgn[2])) * iym;
    const double een2 = (cgn[1] - nu_n * (cgn[0] + cgn[2])) * iym;
    const double een3 = (cgn[2] - nu_n * (cgn[0] + cgn[1])) * iym;
    const double een4 = cgn[3] * ig, een5 = cgn[4] * ig, een6 = cgn[5] * ig;
    eps_vol_n1 = een1 + een2 + een3 + deps_vol;
    const double eps_mean_n = (een1 + een2 + een3) * (one / three);
    const double e1 = (een1 - eps_mean_n) + de1;
    const double e2 = (een2 - eps_mean_n) + de2;
    const double e3 = (een3 - eps_mean_n) + de3;
    const double e4 = een4 + de4, e5 = een5 + de5, e6 = een6 + de6;
    shear_mod_n1 = ym_n1 / (two * (one + nu_n1));
    double dse[6];
    dse[0] = two * shear_mod_n1 * e1;
    dse[1] = two * shear_mod_n1 * e2;
    dse[2] = two * shear_mod_n1 * e3;
    dse[3] = shear_mod_n1 * e4;
    dse[4] = shear_mod_n1 * e5;
    dse[5] = shear_mod_n1 * e6;
#pragma unroll
    for (int k = 0; k < 6; ++k) alpha_n[k] = h[5 + k];
    const double hbari_np1 = beta * hprime_n1;
    const double hbark_np1 = (one - beta) * hprime_n1;
    const double hbark_n = (one - beta) * h[4];
    kbar = (yld_n1 + hbari_np1 * h[2]) / root3;
    lk = one;
    const bool hk = fabs(hbark_n) > htol;            // operands kept off the division's slow path
    const double lkq = xdiv(hbark_np1, hk ? hbark_n : one);
    if (hk) lk = lkq;
#pragma unroll
    for (int k = 0; k < 6; ++k) rtse[k] = dse[k] - alpha_n[k] * lk;
    mrts = sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
    const double yfn = mrts - root2 * kbar;
#pragma unroll
    for (int k = 0; k < 6; ++k) rtse[k] = dse[k] - alpha_n[k];
    mrts = sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                two * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
    if (!pr.lnelas && yfn >= yld_tol * root2 * kbar) {
      yield = true;
      instat = 1;
    }
  }
  if (yield) {
    // ---- mm01_general (constants are the reference's truncated literals)
    const double root2 = 1.414213562373095, twthrd = 0.666666666666667, four = 4.0,
                 root23 = 0.816496580927, root2o3 = 0.4714045207910;
    const double hbari_np1 = beta * hprime_n1;
    const double hbark_np1 = (one - beta) * hprime_n1;
    const double a = (two * shear_mod_n1 + twthrd * hbark_np1) / mrts;
    double b = one - lk;
    const double d = root2o3 * hbari_np1;
    const double vbar = mrts * mrts;
    const double wbar = h[5] * h[5] + h[6] * h[6] + h[7] * h[7] +
                        two * (h[8] * h[8] + h[9] * h[9] + hist_n_first11 * hist_n_first11);
    const double vwbar = rtse[0] * h[5] + rtse[1] * h[6] + rtse[2] * h[7] +
                         two * (rtse[3] * h[8] + rtse[4] * h[9] + rtse[5] * h[10]);
    const double t1 = a * a * vbar - two * d * d;
    const double t2 = -four * d * kbar - two * a * vbar - two * a * b * vwbar;
    const double t3 = b * b * wbar + two * b * vwbar + vbar - two * kbar * kbar;
    const double discr = t2 * t2 - four * t1 * t3;
    const double sq = sqrt(discr);
    const double qroot1 = (-t2 + sq) / (two * t1);
    const double qroot2 = (-t2 - sq) / (two * t1);
    double lambda_deltat = fmin(qroot1, qroot2);
    // mm01.f:864-881: a negative discriminant leaves the roots undefined in the reference and a
    // negative multiplier is a die_abort; both are reported as fatal here.
    if (discr < zero || !(lambda_deltat >= zero)) {
      fatal = 5;
      lambda_deltat = zero;
    }
    h1[0] = lambda_deltat;
    h1[1] = kbar + (root2 / three) * hbari_np1 * lambda_deltat;
    h1[2] = h[2] + lambda_deltat * root23;
    h1[4] = hprime_n1;
    const double const1 = xdiv(twthrd * hbark_np1 * lambda_deltat, mrts);
    const double const2 = one - lambda_deltat * (twthrd * hbark_np1 + two * shear_mod_n1) / mrts;
    b = one - lk;
#pragma unroll
    for (int k = 0; k < 6; ++k) h1[5 + k] = lk * h[5 + k] + const1 * rtse[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) devstr_n1[k] = h1[5 + k] + b * h[5 + k] + const2 * rtse[k];
  } else {
    // elastic points (mm01.f:173-182)
    h1[0] = zero;
    h1[1] = pr.lnelas ? zero : kbar;
    h1[2] = h[2];
    h1[4] = hprime_n1;
#pragma unroll
    for (int k = 0; k < 6; ++k) h1[5 + k] = alpha_n[k] * lk;
#pragma unroll
    for (int k = 0; k < 6; ++k) devstr_n1[k] = rtse[k] + alpha_n[k];
  }
  // ---- mm01_sig_final
  {
    const double half = 0.5;
    const double sig_mean_np1 =
        eps_vol_n1 *
        (three * ym_n1 * nu_n1 / ((one + nu_n1) * (one - two * nu_n1)) + two * shear_mod_n1) /
        three;
    cgn1[0] = devstr_n1[0] + sig_mean_np1;
    cgn1[1] = devstr_n1[1] + sig_mean_np1;
    cgn1[2] = devstr_n1[2] + sig_mean_np1;
    cgn1[3] = devstr_n1[3];
    cgn1[4] = devstr_n1[4];
    cgn1[5] = devstr_n1[5];
    cgn1[6] = cgn[6] + half * (deps[0] * (cgn1[0] + cgn[0]) + deps[1] * (cgn1[1] + cgn[1]) +
                               deps[2] * (cgn1[2] + cgn[2]) + deps[3] * (cgn1[3] + cgn[3]) +
                               deps[4] * (cgn1[4] + cgn[4]) + deps[5] * (cgn1[5] + cgn[5]));
    h1[3] = state_to_dword(instat);
  }
  // ---- mm01_wrapup
  cgn1[7] = cgn[7];
  cgn1[8] = cgn[8];
  if (yield) {
    const double half = 0.5, six = 6.0, root2 = sqrt(2.0);
    const double sbar_n =
        sqrt((cgn[0] - cgn[1]) * (cgn[0] - cgn[1]) + (cgn[1] - cgn[2]) * (cgn[1] - cgn[2]) +
             (cgn[0] - cgn[2]) * (cgn[0] - cgn[2]) +
             six * (cgn[3] * cgn[3] + cgn[4] * cgn[4] + cgn[5] * cgn[5])) /
        root2;
    const double sbar_n1 =
        sqrt((cgn1[0] - cgn1[1]) * (cgn1[0] - cgn1[1]) + (cgn1[1] - cgn1[2]) * (cgn1[1] - cgn1[2]) +
             (cgn1[0] - cgn1[2]) * (cgn1[0] - cgn1[2]) +
             six * (cgn1[3] * cgn1[3] + cgn1[4] * cgn1[4] + cgn1[5] * cgn1[5])) /
        root2;
    const double depbar = h1[2] - h[2];
    cgn1[7] = cgn[7] + half * (sbar_n1 + sbar_n) * depbar;
    cgn1[8] = cgn[8] + depbar;
  }
  return fatal;
}

// linear-elastic [D] of drive_01_update_a (rstgp1.f:942-975) / cnst1's elastic branch, as the 21
// lower-triangle terms stored by rstgp1_store_cep (rstgp1.f:5360-5397):
// (1,1),(2,1),(2,2),(3,1),(3,2),(3,3),(4,1)..(4,4),(5,1)..(5,5),(6,1)..(6,6)
__device__ __forceinline__ void elastic_cep21(double e, double nu, double (&c)[21]) {
  const double one = 1.0, two = 2.0;
  const double c1 = (e / ((one + nu) * (one - two * nu)));
  const double c2 = (one - nu) * c1;
  const double c3 = ((one - two * nu) / two) * c1;
  const double c4 = nu * c1;
#pragma unroll
  for (int k = 0; k < 21; ++k) c[k] = 0.0;
  c[0] = c2; c[1] = c4; c[2] = c2; c[3] = c4; c[4] = c4; c[5] = c2;
  c[9] = c3; c[14] = c3; c[20] = c3;
}

// cnst1 (mm01.f:1196-1348): consistent tangent at n+1 from the trial elastic deviator rtse and
// the updated history (k, H', lambda*dt, state).  Output: 21 lower-triangle terms.
__device__ __forceinline__ void cnst1_cep21(const Mm01Props &pr, const double (&rtsg)[6],
                                            const double (&h1)[11], double (&c)[21]) {
  const double one = 1.0, two = 2.0, three = 3.0, root2 = 1.414213562;  // 10-digit literal
  const double e = pr.e, nu = pr.nu;
  if (dword_to_state(h1[3]) != 1) {
    elastic_cep21(e, nu, c);
    return;
  }
  const double kn1 = h1[1], hprime = h1[4], ldt = h1[0], beta = pr.beta;
  const double g = e / (two * (one + nu));
  const double l = (e * nu) / ((one + nu) * (one - two * nu));
  const double k = (three * l + two * g) / three;
  const double r1 = rtsg[0], r2 = rtsg[1], r3 = rtsg[2], r4 = rtsg[3], r5 = rtsg[4], r6 = rtsg[5];
  const double mrtsq = r1 * r1 + r2 * r2 + r3 * r3 + two * (r4 * r4 + r5 * r5 + r6 * r6);
  const double bb = (root2 * kn1 + (two / three) * (one - beta) * hprime * ldt) / sqrt(mrtsq);
  const double gamma = one / (one + hprime / (three * g));
  const double gambar = gamma - one + bb;
  const double gbar = g * bb;
  const double albar = k - two * gbar / three;
  const double thbar = two * g * gambar;
  const double t = thbar / mrtsq;            // one division instead of 21
  const double dg = albar + two * gbar;
  c[0] = dg - t * (r1 * r1);      // 1,1
  c[1] = albar - t * (r1 * r2);   // 2,1
  c[2] = dg - t * (r2 * r2);      // 2,2
  c[3] = albar - t * (r1 * r3);   // 3,1
  c[4] = albar - t * (r3 * r2);   // 3,2
  c[5] = dg - t * (r3 * r3);      // 3,3
  c[6] = -(t * (r1 * r4));        // 4,1
  c[7] = -(t * (r2 * r4));        // 4,2
  c[8] = -(t * (r3 * r4));        // 4,3
  c[9] = gbar - t * (r4 * r4);    // 4,4
  c[10] = -(t * (r1 * r5));       // 5,1
  c[11] = -(t * (r2 * r5));       // 5,2
  c[12] = -(t * (r3 * r5));       // 5,3
  c[13] = -(t * (r4 * r5));       // 5,4
  c[14] = gbar - t * (r5 * r5);   // 5,5
  c[15] = -(t * (r1 * r6));       // 6,1
  c[16] = -(t * (r2 * r6));       // 6,2
  c[17] = -(t * (r3 * r6));       // 6,3
  c[18] = -(t * (r4 * r6));       // 6,4
  c[19] = -(t * (r5 * r6));       // 6,5
  c[20] = gbar - t * (r6 * r6);   // 6,6
}

}  // namespace w3d
