/*
 * w3d_oracle_mm01.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): bilinear Mises model mm01 + cnst1.
 * Follows reference src/mm01.f line by line, including its parity quirks (SURVEY 8a register):
 *   - integer state kept in the low 32 bits of history(.,4) through an EQUIVALENCE (mm01.f:248-251)
 *   - truncated constants root2, twthrd, root23, root2o3 in mm01_general (mm01.f:803-807)
 *   - wbar reads history(1,11) (first element of the block), not (i,11) (mm01.f:843-845)
 *   - lambda_deltat = min(qroot1,qroot2) re-evaluated after the discr<0 branch (mm01.f:864-877)
 *   - cnst1 uses root2 = 1.414213562d0 (10 digits) (mm01.f:1220-1221)
 * PARITY STATUS: see w3d_oracle.h (mm01, mm05 and mm10 pinned end to end by reference known
 * answers at 1e-5; unpinned at function granularity).
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include "w3d_oracle_internal.h"

namespace w3do {

// history(span,*) : leading dimension is span (drive_eps_sig_internal_forces.f:1145)
#define H(h, i, k) h[((k)-1) * span + (i)]

static inline double state_to_dword(int32_t st) {
  // equivalence( dword, iword ); iword(1) = st ; iword(2) = 0   (little-endian low word)
  int32_t iw[2] = {st, 0};
  double d;
  std::memcpy(&d, iw, 8);
  return d;
}
static inline int32_t dword_to_state(double d) {
  int32_t iw[2];
  std::memcpy(iw, &d, 8);
  return iw[0];
}

// mm01.f:24-276 (mm01) with mm01_set_history (222-276), mm01_init (308-520),
// mm01_general (776-940), mm01_sig_final (600-660), mm01_wrapup (524-598)
void mm01(int span, int step, const double *ym_n1, const double *nu_n1, const double *beta,
          const double *hprime_n1, const bool *lnelas, const double *yld_n1,
          double *cgn /*(mxvl,*)*/, double *cgn1 /*(mxvl,*)*/, const double *deps /*(mxvl,*)*/,
          double *history /*(span,*)*/, double *history1 /*(span,*)*/, double *rtse /*(mxvl,6)*/,
          const double *dtemps, const double *ym_n, const double *nu_nn, int *fatal) {
  double yfn[mxvl], mrts[mxvl], alpha_n[mxvl * 6], devstr_n1[mxvl * nstr], shear_mod_n1[mxvl],
      lk[mxvl], kbar[mxvl], eps_vol_n1[mxvl];
  int instat[mxvl];
  bool yield[mxvl];
  (void)dtemps;
  *fatal = 0;

  // ---- mm01_set_history (mm01.f:222-276), only during load step 1
  if (step == 1) {
    const double root3 = std::sqrt(3.0), zero = 0.0;
    const double dword = state_to_dword(3);
    for (int i = 0; i < span; ++i) {
      const double kn = yld_n1[i] / root3;
      H(history, i, 1) = zero;
      H(history, i, 2) = kn;
      H(history, i, 3) = zero;
      H(history, i, 4) = dword;
      H(history, i, 5) = hprime_n1[i];
      for (int k = 6; k <= 11; ++k) H(history, i, k) = zero;
      H(history1, i, 1) = zero;
      H(history1, i, 2) = kn;
      H(history1, i, 3) = zero;
      H(history1, i, 4) = dword;
      H(history1, i, 5) = hprime_n1[i];
      for (int k = 6; k <= 11; ++k) H(history1, i, k) = zero;
      V2(cgn, i, 8) = zero;
      V2(cgn, i, 9) = zero;
    }
  }

  // ---- mm01_init (mm01.f:308-520)
  {
    const double one = 1.0, two = 2.0, three = 3.0, htol = 0.000001, root3 = std::sqrt(3.0),
                 root2 = std::sqrt(2.0), yld_tol = 0.0000001;
    for (int i = 0; i < span; ++i) {
      const double deps_vol = V2(deps, i, 1) + V2(deps, i, 2) + V2(deps, i, 3);
      const double eps_mean = deps_vol / three;
      const double de1 = V2(deps, i, 1) - eps_mean;
      const double de2 = V2(deps, i, 2) - eps_mean;
      const double de3 = V2(deps, i, 3) - eps_mean;
      const double de4 = V2(deps, i, 4);
      const double de5 = V2(deps, i, 5);
      const double de6 = V2(deps, i, 6);
      const double e_n = ym_n[i];
      const double nu_n = nu_nn[i];
      const double g_n = e_n / two / (one + nu_n);
      const double een1 = (V2(cgn, i, 1) - nu_n * (V2(cgn, i, 2) + V2(cgn, i, 3))) / e_n;
      const double een2 = (V2(cgn, i, 2) - nu_n * (V2(cgn, i, 1) + V2(cgn, i, 3))) / e_n;
      const double een3 = (V2(cgn, i, 3) - nu_n * (V2(cgn, i, 1) + V2(cgn, i, 2))) / e_n;
      const double een4 = V2(cgn, i, 4) / g_n;
      const double een5 = V2(cgn, i, 5) / g_n;
      const double een6 = V2(cgn, i, 6) / g_n;
      eps_vol_n1[i] = een1 + een2 + een3 + deps_vol;
      const double eps_mean_n = (een1 + een2 + een3) / three;
      const double e1 = (een1 - eps_mean_n) + de1;
      const double e2 = (een2 - eps_mean_n) + de2;
      const double e3 = (een3 - eps_mean_n) + de3;
      const double e4 = een4 + de4;
      const double e5 = een5 + de5;
      const double e6 = een6 + de6;
      shear_mod_n1[i] = ym_n1[i] / (two * (one + nu_n1[i]));
      double dse[7];
      dse[1] = two * shear_mod_n1[i] * e1;
      dse[2] = two * shear_mod_n1[i] * e2;
      dse[3] = two * shear_mod_n1[i] * e3;
      dse[4] = shear_mod_n1[i] * e4;
      dse[5] = shear_mod_n1[i] * e5;
      dse[6] = shear_mod_n1[i] * e6;
      for (int k = 1; k <= 6; ++k) V2(alpha_n, i, k) = H(history, i, 5 + k);
      const double hbari_np1 = beta[i] * hprime_n1[i];
      const double hbark_np1 = (one - beta[i]) * hprime_n1[i];
      const double hbark_n = (one - beta[i]) * H(history, i, 5);
      kbar[i] = (yld_n1[i] + hbari_np1 * H(history, i, 3)) / root3;
      lk[i] = one;
      if (std::fabs(hbark_n) > htol) lk[i] = hbark_np1 / hbark_n;
      for (int k = 1; k <= 6; ++k) V2(rtse, i, k) = dse[k] - V2(alpha_n, i, k) * lk[i];
      mrts[i] = std::sqrt(V2(rtse, i, 1) * V2(rtse, i, 1) + V2(rtse, i, 2) * V2(rtse, i, 2) +
                          V2(rtse, i, 3) * V2(rtse, i, 3) +
                          two * (V2(rtse, i, 4) * V2(rtse, i, 4) + V2(rtse, i, 5) * V2(rtse, i, 5) +
                                 V2(rtse, i, 6) * V2(rtse, i, 6)));
      yfn[i] = mrts[i] - root2 * kbar[i];
      for (int k = 1; k <= 6; ++k) V2(rtse, i, k) = dse[k] - V2(alpha_n, i, k);
      mrts[i] = std::sqrt(V2(rtse, i, 1) * V2(rtse, i, 1) + V2(rtse, i, 2) * V2(rtse, i, 2) +
                          V2(rtse, i, 3) * V2(rtse, i, 3) +
                          two * (V2(rtse, i, 4) * V2(rtse, i, 4) + V2(rtse, i, 5) * V2(rtse, i, 5) +
                                 V2(rtse, i, 6) * V2(rtse, i, 6)));
      instat[i] = 3;
      yield[i] = false;
      if (lnelas[i]) continue;
      if (yfn[i] >= yld_tol * root2 * kbar[i]) {
        yield[i] = true;
        instat[i] = 1;
      }
    }
  }

  // ---- mm01_general (mm01.f:776-940)
  {
    const double root2 = 1.414213562373095, twthrd = 0.666666666666667, one = 1.0, two = 2.0,
                 three = 3.0, root23 = 0.816496580927, four = 4.0, zero = 0.0,
                 root2o3 = 0.4714045207910;
    for (int i = 0; i < span; ++i) {
      if (!yield[i]) continue;
      const double hbari_np1 = beta[i] * hprime_n1[i];
      const double hbark_np1 = (one - beta[i]) * hprime_n1[i];
      const double a = (two * shear_mod_n1[i] + twthrd * hbark_np1) / mrts[i];
      double b = one - lk[i];
      const double d = root2o3 * hbari_np1;
      const double vbar = mrts[i] * mrts[i];
      // quirk: history(1,11) -- first element of the block (mm01.f:843-845)
      const double wbar = H(history, i, 6) * H(history, i, 6) + H(history, i, 7) * H(history, i, 7) +
                          H(history, i, 8) * H(history, i, 8) +
                          two * (H(history, i, 9) * H(history, i, 9) +
                                 H(history, i, 10) * H(history, i, 10) +
                                 H(history, 0, 11) * H(history, 0, 11));
      const double vwbar = V2(rtse, i, 1) * H(history, i, 6) + V2(rtse, i, 2) * H(history, i, 7) +
                           V2(rtse, i, 3) * H(history, i, 8) +
                           two * (V2(rtse, i, 4) * H(history, i, 9) + V2(rtse, i, 5) * H(history, i, 10) +
                                  V2(rtse, i, 6) * H(history, i, 11));
      const double t1 = a * a * vbar - two * d * d;
      const double t2 = -four * d * kbar[i] - two * a * vbar - two * a * b * vwbar;
      const double t3 = b * b * wbar + two * b * vwbar + vbar - two * kbar[i] * kbar[i];
      const double discr = t2 * t2 - four * t1 * t3;
      // quirk (mm01.f:864-877): qroot1/qroot2 are used after the if-block even when discr < 0
      // (undefined in the reference: uninitialised locals).  We evaluate them the same way in
      // both branches, which yields NaN for discr < 0 -> flagged fatal below.  Unreachable for
      // isothermal loading (lk == 1 => b == 0 => discr > 0).
      const double qroot1 = (-t2 + std::sqrt(discr)) / (two * t1);
      const double qroot2 = (-t2 - std::sqrt(discr)) / (two * t1);
      double lambda_deltat = std::fmin(qroot1, qroot2);
      if (discr < zero) lambda_deltat = std::nan("");
      if (!(lambda_deltat >= zero)) {  // mm01.f:878-881 die_abort (and the undefined case)
        *fatal = 1;
        lambda_deltat = zero;
      }
      H(history1, i, 1) = lambda_deltat;
      H(history1, i, 2) = kbar[i] + (root2 / three) * hbari_np1 * lambda_deltat;
      H(history1, i, 3) = H(history, i, 3) + lambda_deltat * root23;
      H(history1, i, 5) = hprime_n1[i];
      const double const1 = twthrd * hbark_np1 * lambda_deltat / mrts[i];
      const double const2 =
          one - lambda_deltat * (twthrd * hbark_np1 + two * shear_mod_n1[i]) / mrts[i];
      b = one - lk[i];
      for (int k = 1; k <= 6; ++k)
        H(history1, i, 5 + k) = lk[i] * H(history, i, 5 + k) + const1 * V2(rtse, i, k);
      for (int k = 1; k <= 6; ++k)
        V2(devstr_n1, i, k) = H(history1, i, 5 + k) + b * H(history, i, 5 + k) + const2 * V2(rtse, i, k);
    }
  }

  // ---- elastic points (mm01.f:173-182)
  for (int i = 0; i < span; ++i) {
    if (yield[i]) continue;
    H(history1, i, 1) = 0.0;
    H(history1, i, 2) = kbar[i];
    if (lnelas[i]) H(history1, i, 2) = 0.0;
    H(history1, i, 3) = H(history, i, 3);
    H(history1, i, 5) = hprime_n1[i];
    for (int k = 1; k <= 6; ++k) H(history1, i, 5 + k) = V2(alpha_n, i, k) * lk[i];
    for (int k = 1; k <= 6; ++k) V2(devstr_n1, i, k) = V2(rtse, i, k) + V2(alpha_n, i, k);
  }

  // ---- mm01_sig_final (mm01.f:600-660)
  {
    const double one = 1.0, two = 2.0, three = 3.0, half = 0.5;
    for (int i = 0; i < span; ++i) {
      const double sig_mean_np1 =
          eps_vol_n1[i] *
          (three * ym_n1[i] * nu_n1[i] / ((one + nu_n1[i]) * (one - two * nu_n1[i])) +
           two * shear_mod_n1[i]) /
          three;
      V2(cgn1, i, 1) = V2(devstr_n1, i, 1) + sig_mean_np1;
      V2(cgn1, i, 2) = V2(devstr_n1, i, 2) + sig_mean_np1;
      V2(cgn1, i, 3) = V2(devstr_n1, i, 3) + sig_mean_np1;
      V2(cgn1, i, 4) = V2(devstr_n1, i, 4);
      V2(cgn1, i, 5) = V2(devstr_n1, i, 5);
      V2(cgn1, i, 6) = V2(devstr_n1, i, 6);
      V2(cgn1, i, 7) =
          V2(cgn, i, 7) + half * (V2(deps, i, 1) * (V2(cgn1, i, 1) + V2(cgn, i, 1)) +
                                  V2(deps, i, 2) * (V2(cgn1, i, 2) + V2(cgn, i, 2)) +
                                  V2(deps, i, 3) * (V2(cgn1, i, 3) + V2(cgn, i, 3)) +
                                  V2(deps, i, 4) * (V2(cgn1, i, 4) + V2(cgn, i, 4)) +
                                  V2(deps, i, 5) * (V2(cgn1, i, 5) + V2(cgn, i, 5)) +
                                  V2(deps, i, 6) * (V2(cgn1, i, 6) + V2(cgn, i, 6)));
      H(history1, i, 4) = state_to_dword(instat[i]);
    }
  }

  // ---- mm01_wrapup (mm01.f:524-598); eps_elas is not an output of this path
  {
    const double half = 0.5, six = 6.0, root2 = std::sqrt(2.0);
    for (int i = 0; i < span; ++i) {
      V2(cgn1, i, 8) = V2(cgn, i, 8);
      V2(cgn1, i, 9) = V2(cgn, i, 9);
      if (!yield[i]) continue;
      double s[7];
      for (int k = 1; k <= 6; ++k) s[k] = V2(cgn, i, k);
      const double sbar_n =
          std::sqrt((s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[3]) * (s[2] - s[3]) +
                    (s[1] - s[3]) * (s[1] - s[3]) + six * (s[4] * s[4] + s[5] * s[5] + s[6] * s[6])) /
          root2;
      for (int k = 1; k <= 6; ++k) s[k] = V2(cgn1, i, k);
      const double sbar_n1 =
          std::sqrt((s[1] - s[2]) * (s[1] - s[2]) + (s[2] - s[3]) * (s[2] - s[3]) +
                    (s[1] - s[3]) * (s[1] - s[3]) + six * (s[4] * s[4] + s[5] * s[5] + s[6] * s[6])) /
          root2;
      const double ebarp_n = H(history, i, 3);
      const double ebarp_n1 = H(history1, i, 3);
      const double depbar = ebarp_n1 - ebarp_n;
      V2(cgn1, i, 8) = V2(cgn, i, 8) + half * (sbar_n1 + sbar_n) * depbar;
      V2(cgn1, i, 9) = V2(cgn, i, 9) + depbar;
    }
  }
}

// mm01.f:1196-1348 (cnst1): consistent tangent.  kn1, hprime, ldt, dstates are columns of the
// history1 array (stride 1 over i): history1(1,2), (1,5), (1,1), (1,4)  (rstgp1.f:851-856)
void cnst1(int span, double *cep /*(mxvl,6,6)*/, const double *rtsg /*(mxvl,6)*/, const double *nu,
           const double *e, const double *kn1, const double *hprime, const double *beta,
           const double *ldt, const double *dstates) {
  const double zero = 0.0, one = 1.0, two = 2.0, three = 3.0, root2 = 1.414213562;
#define CEP(i, r, c) Q6(cep, i, r, c)
  bool yield[mxvl];
  for (int i = 0; i < span; ++i) yield[i] = dword_to_state(dstates[i]) == 1;
  for (int i = 0; i < span; ++i) {
    if (yield[i]) continue;
    for (int r = 1; r <= 6; ++r)
      for (int c = 1; c <= 6; ++c)
        if ((r <= 3) != (c <= 3) || (r > 3 && c > 3 && r != c)) CEP(i, r, c) = zero;
    const double c1 = (e[i] / ((one + nu[i]) * (one - two * nu[i])));
    const double c2 = (one - nu[i]) * c1;
    const double c3 = ((one - two * nu[i]) / two) * c1;
    const double c4 = nu[i] * c1;
    CEP(i, 1, 1) = c2; CEP(i, 2, 2) = c2; CEP(i, 3, 3) = c2;
    CEP(i, 4, 4) = c3; CEP(i, 5, 5) = c3; CEP(i, 6, 6) = c3;
    CEP(i, 1, 2) = c4; CEP(i, 1, 3) = c4; CEP(i, 2, 1) = c4;
    CEP(i, 3, 1) = c4; CEP(i, 2, 3) = c4; CEP(i, 3, 2) = c4;
  }
  for (int i = 0; i < span; ++i) {
    if (!yield[i]) continue;
    const double g = e[i] / (two * (one + nu[i]));
    const double l = (e[i] * nu[i]) / ((one + nu[i]) * (one - two * nu[i]));
    const double k = (three * l + two * g) / three;
    const double r1 = V2(rtsg, i, 1), r2 = V2(rtsg, i, 2), r3 = V2(rtsg, i, 3), r4 = V2(rtsg, i, 4),
                 r5 = V2(rtsg, i, 5), r6 = V2(rtsg, i, 6);
    const double mrtsq = r1 * r1 + r2 * r2 + r3 * r3 + two * (r4 * r4 + r5 * r5 + r6 * r6);
    const double bb = (root2 * kn1[i] + (two / three) * (one - beta[i]) * hprime[i] * ldt[i]) /
                      std::sqrt(mrtsq);
    const double gamma = one / (one + hprime[i] / (three * g));
    const double gambar = gamma - one + bb;
    const double gbar = g * bb;
    const double albar = k - two * gbar / three;
    const double thbar = two * g * gambar;
    CEP(i, 1, 1) = (albar + two * gbar - thbar * (r1 * r1) / mrtsq);
    CEP(i, 2, 2) = (albar + two * gbar - thbar * (r2 * r2) / mrtsq);
    CEP(i, 3, 3) = (albar + two * gbar - thbar * (r3 * r3) / mrtsq);
    CEP(i, 4, 4) = (gbar - thbar * (r4 * r4) / mrtsq);
    CEP(i, 5, 5) = (gbar - thbar * (r5 * r5) / mrtsq);
    CEP(i, 6, 6) = (gbar - thbar * (r6 * r6) / mrtsq);
    CEP(i, 2, 1) = (albar - thbar * r1 * r2 / mrtsq);
    CEP(i, 3, 1) = (albar - thbar * r1 * r3 / mrtsq);
    CEP(i, 4, 1) = -(thbar * r1 * r4 / mrtsq);
    CEP(i, 5, 1) = -(thbar * r1 * r5 / mrtsq);
    CEP(i, 6, 1) = -(thbar * r1 * r6 / mrtsq);
    CEP(i, 3, 2) = (albar - thbar * r3 * r2 / mrtsq);
    CEP(i, 4, 2) = -(thbar * r2 * r4 / mrtsq);
    CEP(i, 5, 2) = -(thbar * r2 * r5 / mrtsq);
    CEP(i, 6, 2) = -(thbar * r2 * r6 / mrtsq);
    CEP(i, 4, 3) = -(thbar * r3 * r4 / mrtsq);
    CEP(i, 5, 3) = -(thbar * r3 * r5 / mrtsq);
    CEP(i, 6, 3) = -(thbar * r3 * r6 / mrtsq);
    CEP(i, 5, 4) = -(thbar * r4 * r5 / mrtsq);
    CEP(i, 6, 4) = -(thbar * r4 * r6 / mrtsq);
    CEP(i, 6, 5) = -(thbar * r5 * r6 / mrtsq);
    for (int r = 1; r <= 6; ++r)
      for (int c = r + 1; c <= 6; ++c) CEP(i, r, c) = CEP(i, c, r);
  }
#undef CEP
}

}  // namespace w3do
