/*
 * w3d_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).  See w3d_oracle.h for the rules.
 *
 * Literal scalar restatement of the WARP3D element-block hot path for hex8 blocks.  The code keeps
 * the reference's loop structure (outer loop Gauss points, inner loops i = 1..span over the block,
 * block-local arrays with leading dimension mxvl) and its evaluation order; compile with
 * -ffp-contract=off so no FMA contraction changes the arithmetic.
 *
 * PARITY STATUS: no function-level golden vectors exist in the reference and it cannot be built
 * here; pinned end to end by two of its known answers (see w3d_oracle.h).  Citations are
 * reference src/ file:line.
 */
#include "w3d_oracle.h"
#include "w3d_oracle_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace w3do {

// ---------------------------------------------------------------------------------------------
// model helpers
// ---------------------------------------------------------------------------------------------
BlockInfo block_info(const w3do_model *m, int blk /*0-based*/) {
  BlockInfo b;
  b.blk = blk;
  b.span = m->elblks[4 * blk + 0];
  b.felem = m->elblks[4 * blk + 1];
  const int fe = b.felem;
  b.elem_type = iprop(m, 1, fe);
  b.nnode = iprop(m, 2, fe);
  b.ndof = iprop(m, 4, fe);
  b.int_order = iprop(m, 5, fe);
  b.ngp = iprop(m, 6, fe);
  b.geonl = lprop(m, 18, fe);
  b.bbar = lprop(m, 19, fe);
  b.mat_type = iprop(m, 25, fe);
  b.matnum = iprop(m, 38, fe);
  b.totdof = b.nnode * b.ndof;
  b.hist_size = m->hist_size[blk];
  b.utsz = (b.totdof * b.totdof - b.totdof) / 2 + b.totdof;
  return b;
}

int64_t block_offset(const w3do_model *m, int blk, int which) {
  int64_t off = 0;
  for (int k = 0; k < blk; ++k) {
    BlockInfo b = block_info(m, k);
    int64_t per = 0;
    switch (which) {
      case 0: per = (int64_t)b.hist_size * b.ngp * b.span; break;
      case 1: per = (int64_t)nstr * b.ngp * b.span; break;
      case 2: per = (int64_t)nstrs * b.ngp * b.span; break;
      case 3: per = (int64_t)9 * b.ngp * b.span; break;
      case 4: per = (int64_t)21 * b.span * b.ngp; break;
      case 5: per = (int64_t)b.utsz * b.span; break;
      case 6: per = b.span; break;
      case 7: per = (int64_t)b.totdof * b.span; break;
      default: per = 0;
    }
    off += per;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// getgpts.f:642-718 (quad2) -- 2x2x2 rule, WARP3D point ordering, weight 1
// ---------------------------------------------------------------------------------------------
void quad2(int gpn, double &xi, double &eta, double &zeta, double &w) {
  const double root33 = 0.5773502691896257645;
  const double l1 = root33, w1 = 1.0;
  switch (gpn) {
    case 1: xi = -l1; eta = -l1; zeta = -l1; break;
    case 2: xi = -l1; eta = l1;  zeta = -l1; break;
    case 3: xi = l1;  eta = -l1; zeta = -l1; break;
    case 4: xi = l1;  eta = l1;  zeta = -l1; break;
    case 5: xi = -l1; eta = -l1; zeta = l1;  break;
    case 6: xi = -l1; eta = l1;  zeta = l1;  break;
    case 7: xi = l1;  eta = -l1; zeta = l1;  break;
    case 8: xi = l1;  eta = l1;  zeta = l1;  break;
    default: xi = eta = zeta = 0.0;
  }
  w = w1;
}

// ---------------------------------------------------------------------------------------------
// derivs.f:283-343 (deriv2) -- hex8 shape function derivatives, WARP3D node ordering
// ---------------------------------------------------------------------------------------------
void deriv2(double xi, double eta, double zeta, double *nxi, double *neta, double *nzeta) {
  const double one = 1.0, one25 = 0.125;
  const double xp = one + xi, ep = one + eta, zp = one + zeta;
  const double xm = one - xi, em = one - eta, zm = one - zeta;
  nxi[0] = -em * zp * one25; neta[0] = -xm * zp * one25; nzeta[0] = xm * em * one25;
  nxi[1] = -em * zm * one25; neta[1] = -xm * zm * one25; nzeta[1] = -xm * em * one25;
  nxi[2] = -ep * zm * one25; neta[2] = xm * zm * one25;  nzeta[2] = -xm * ep * one25;
  nxi[3] = -ep * zp * one25; neta[3] = xm * zp * one25;  nzeta[3] = xm * ep * one25;
  nxi[4] = em * zp * one25;  neta[4] = -xp * zp * one25; nzeta[4] = xp * em * one25;
  nxi[5] = em * zm * one25;  neta[5] = -xp * zm * one25; nzeta[5] = -xp * em * one25;
  nxi[6] = ep * zm * one25;  neta[6] = xp * zm * one25;  nzeta[6] = -xp * ep * one25;
  nxi[7] = ep * zp * one25;  neta[7] = xp * zp * one25;  nzeta[7] = xp * ep * one25;
}

// ---------------------------------------------------------------------------------------------
// setup.f:268-276, 407-416 -- cp, dcp, icp (upper-triangle packing)
// ---------------------------------------------------------------------------------------------
void make_cp_icp(int n, int *cp, int *dcp, int *icp /* (utsz,2) column-major */) {
  const int utsz = (n * n + n) / 2;
  for (int j = 1; j <= n; ++j) {
    int stpos = 0;  // clmloc(j)
    for (int i = 1; i <= j - 1; ++i) stpos = stpos + j - i;
    // note: clmloc(col) = sum_{i=1}^{col-1} (col - i) = col(col-1)/2
    cp[j - 1] = stpos;
    if (dcp) dcp[j - 1] = cp[j - 1] + j;
    for (int i = 1; i <= j; ++i) {
      const int pos = cp[j - 1] + i;
      icp[(pos - 1)] = i;
      icp[utsz + (pos - 1)] = j;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Contraction convention of the noise-amplifying chain (coordinates -> Jacobian -> Gamma -> theta
// -> F -> C -> Cardano -> U^-1 -> R).  The reference's closed-form polar decomposition turns a
// 1-ulp change of C into ~1e-9 of R for small strains, so "the reference's result" is defined
// only up to the compiler's choice of fused multiply-adds (ifort -O3 contracts; which products it
// fuses is not visible from the source).  The restatement therefore FIXES one pattern -- the one
// an optimising compiler produces when it contracts left to right:
//     a*b + c*d + e*f  ->  fma(e, f, fma(c, d, a*b)),    a*b - c*d  ->  fma(a, b, -(c*d)),
//     s = s + a*b      ->  fma(a, b, s)
// -- and writes it out explicitly (the file is compiled with -ffp-contract=off, so nothing else
// is fused).  The engine evaluates the same chain with the same fused operations in the same
// order, which makes it bit-identical up to libm's atan2 / sincos.
// ---------------------------------------------------------------------------------------------
static inline double fm(double a, double b, double c) { return std::fma(a, b, c); }

// ---------------------------------------------------------------------------------------------
// gtmat1.f:300-507 (jacob1), 3-D solid branch
// ---------------------------------------------------------------------------------------------
void jacob1(int span, double *jac /*(mxvl,3,3)*/, double *dj, double *gama /*(mxvl,3,3)*/,
            const double *nxi, const double *neta, const double *nzeta, const double *ce,
            int nnode) {
#define JAC(i, r, c) jac[(((c)-1) * 3 + ((r)-1)) * mxvl + (i)]
#define CE(i, j) ce[((j)-1) * mxvl + (i)]
  for (int k = 0; k < 9 * mxvl; ++k) { jac[k] = 0.0; gama[k] = 0.0; }
  for (int j = 1; j <= nnode; ++j) {
    for (int i = 0; i < span; ++i) {
      JAC(i, 1, 1) = fm(nxi[j - 1], CE(i, j), JAC(i, 1, 1));
      JAC(i, 1, 2) = fm(nxi[j - 1], CE(i, nnode + j), JAC(i, 1, 2));
      JAC(i, 1, 3) = fm(nxi[j - 1], CE(i, 2 * nnode + j), JAC(i, 1, 3));
      JAC(i, 2, 1) = fm(neta[j - 1], CE(i, j), JAC(i, 2, 1));
      JAC(i, 2, 2) = fm(neta[j - 1], CE(i, nnode + j), JAC(i, 2, 2));
      JAC(i, 2, 3) = fm(neta[j - 1], CE(i, 2 * nnode + j), JAC(i, 2, 3));
      JAC(i, 3, 1) = fm(nzeta[j - 1], CE(i, j), JAC(i, 3, 1));
      JAC(i, 3, 2) = fm(nzeta[j - 1], CE(i, nnode + j), JAC(i, 3, 2));
      JAC(i, 3, 3) = fm(nzeta[j - 1], CE(i, 2 * nnode + j), JAC(i, 3, 3));
    }
  }
  double j1[mxvl], j2[mxvl], j3[mxvl];
  for (int i = 0; i < span; ++i) {
    j1[i] = fm(JAC(i, 2, 2), JAC(i, 3, 3), -(JAC(i, 2, 3) * JAC(i, 3, 2)));
    j2[i] = fm(JAC(i, 2, 1), JAC(i, 3, 3), -(JAC(i, 2, 3) * JAC(i, 3, 1)));
    j3[i] = fm(JAC(i, 2, 1), JAC(i, 3, 2), -(JAC(i, 2, 2) * JAC(i, 3, 1)));
    dj[i] = fm(JAC(i, 1, 3), j3[i], fm(-JAC(i, 1, 2), j2[i], JAC(i, 1, 1) * j1[i]));
  }
  // gtmat1.f:445-459: non-positive det only produces warnings; no state change.
  for (int i = 0; i < span; ++i) {
    // division by the common determinant as ONE reciprocal and nine products -- what the
    // reference's build does with these statements (ifort -O3 runs with -no-prec-div); part of the
    // contraction convention above
    const double rdj = 1.0 / dj[i];
    GAM(i, 1, 1) = j1[i] * rdj;
    GAM(i, 2, 1) = -j2[i] * rdj;
    GAM(i, 3, 1) = j3[i] * rdj;
    GAM(i, 1, 2) = fm(JAC(i, 3, 2), JAC(i, 1, 3), -(JAC(i, 1, 2) * JAC(i, 3, 3))) * rdj;
    GAM(i, 2, 2) = fm(JAC(i, 1, 1), JAC(i, 3, 3), -(JAC(i, 3, 1) * JAC(i, 1, 3))) * rdj;
    GAM(i, 3, 2) = fm(JAC(i, 1, 2), JAC(i, 3, 1), -(JAC(i, 1, 1) * JAC(i, 3, 2))) * rdj;
    GAM(i, 1, 3) = fm(JAC(i, 1, 2), JAC(i, 2, 3), -(JAC(i, 1, 3) * JAC(i, 2, 2))) * rdj;
    GAM(i, 2, 3) = fm(JAC(i, 1, 3), JAC(i, 2, 1), -(JAC(i, 1, 1) * JAC(i, 2, 3))) * rdj;
    GAM(i, 3, 3) = fm(JAC(i, 1, 1), JAC(i, 2, 2), -(JAC(i, 1, 2) * JAC(i, 2, 1))) * rdj;
  }
#undef JAC
#undef CE
}

// ---------------------------------------------------------------------------------------------
// vol_terms.f:14-108, vol_avg.f:13-68 -- mean-dilatation B-bar volume terms (hex8)
// ---------------------------------------------------------------------------------------------
void vol_terms(const double *in_jacob /*(mxvl,3,3)*/, const double *dj, double *vol /*(mxvl,8,3)*/,
               const double *nxi, const double *neta, const double *nzeta, double *volume,
               int span) {
#define IJ(i, r, c) in_jacob[(((c)-1) * 3 + ((r)-1)) * mxvl + (i)]
  for (int i = 0; i < span; ++i) {
    double a, b, c, d;
    d = dj[i];
    for (int k = 1; k <= 3; ++k) {
      a = IJ(i, k, 1); b = IJ(i, k, 2); c = IJ(i, k, 3);
      for (int j = 1; j <= 8; ++j)
        VOL(i, j, k) = VOL(i, j, k) + (a * nxi[j - 1] + b * neta[j - 1] + c * nzeta[j - 1]) * d;
    }
    volume[i] = volume[i] + d;
  }
#undef IJ
}
void vol_avg(double *vol, const double *volume, int span) {
  for (int i = 0; i < span; ++i) {
    const double voli = 1.0 / volume[i];
    for (int k = 1; k <= 3; ++k)
      for (int j = 1; j <= 8; ++j) VOL(i, j, k) = VOL(i, j, k) * voli;
  }
}

// ---------------------------------------------------------------------------------------------
// gtmat1.f:522-584 (tcomp1): displacement gradients theta(mxvl,9)
// ---------------------------------------------------------------------------------------------
void tcomp1(int span, double *theta /*(mxvl,9)*/, const double *nxi, const double *neta,
            const double *nzeta, const double *gama, const double *ue /*(mxvl,*)*/, int nnode) {
#define TH(i, k) theta[((k)-1) * mxvl + (i)]
#define UE(i, j) ue[((j)-1) * mxvl + (i)]
  std::vector<double> thtemp((size_t)mxvl * mxndel * 3);
#define THT(i, j, k) thtemp[(((k)-1) * mxndel + ((j)-1)) * mxvl + (i)]
  for (int k = 0; k < 9 * mxvl; ++k) theta[k] = 0.0;
  for (int j = 1; j <= nnode; ++j)
    for (int i = 0; i < span; ++i) {
      THT(i, j, 1) = fm(GAM(i, 1, 3), nzeta[j - 1], fm(GAM(i, 1, 2), neta[j - 1], GAM(i, 1, 1) * nxi[j - 1]));
      THT(i, j, 2) = fm(GAM(i, 2, 3), nzeta[j - 1], fm(GAM(i, 2, 2), neta[j - 1], GAM(i, 2, 1) * nxi[j - 1]));
      THT(i, j, 3) = fm(GAM(i, 3, 3), nzeta[j - 1], fm(GAM(i, 3, 2), neta[j - 1], GAM(i, 3, 1) * nxi[j - 1]));
    }
  const int tpos1 = nnode, tpos2 = 2 * nnode;
  for (int j = 1; j <= nnode; ++j)
    for (int i = 0; i < span; ++i) {
      TH(i, 1) = fm(THT(i, j, 1), UE(i, j), TH(i, 1));
      TH(i, 2) = fm(THT(i, j, 1), UE(i, tpos1 + j), TH(i, 2));
      TH(i, 3) = fm(THT(i, j, 1), UE(i, tpos2 + j), TH(i, 3));
      TH(i, 4) = fm(THT(i, j, 2), UE(i, j), TH(i, 4));
      TH(i, 5) = fm(THT(i, j, 2), UE(i, tpos1 + j), TH(i, 5));
      TH(i, 6) = fm(THT(i, j, 2), UE(i, tpos2 + j), TH(i, 6));
      TH(i, 7) = fm(THT(i, j, 3), UE(i, j), TH(i, 7));
      TH(i, 8) = fm(THT(i, j, 3), UE(i, tpos1 + j), TH(i, 8));
      TH(i, 9) = fm(THT(i, j, 3), UE(i, tpos2 + j), TH(i, 9));
    }
#undef THT
#undef UE
}

// ---------------------------------------------------------------------------------------------
// gtmat1.f:601-674 (fcomp1): F = I + theta, det F; error=1 if det F <= 1e-20
// ---------------------------------------------------------------------------------------------
void fcomp1(int span, double *f /*(mxvl,3,3)*/, double *df, const double *theta, int &error) {
  const double one = 1.0, zero_check = 1.0e-20;
  for (int i = 0; i < span; ++i) {
    F3(f, i, 1, 1) = TH(i, 1) + one;
    F3(f, i, 1, 2) = TH(i, 4);
    F3(f, i, 1, 3) = TH(i, 7);
    F3(f, i, 2, 1) = TH(i, 2);
    F3(f, i, 2, 2) = TH(i, 5) + one;
    F3(f, i, 2, 3) = TH(i, 8);
    F3(f, i, 3, 1) = TH(i, 3);
    F3(f, i, 3, 2) = TH(i, 6);
    F3(f, i, 3, 3) = TH(i, 9) + one;
  }
  for (int i = 0; i < span; ++i) {
    const double f1 = fm(F3(f, i, 2, 2), F3(f, i, 3, 3), -(F3(f, i, 2, 3) * F3(f, i, 3, 2)));
    const double f2 = fm(F3(f, i, 2, 1), F3(f, i, 3, 3), -(F3(f, i, 2, 3) * F3(f, i, 3, 1)));
    const double f3 = fm(F3(f, i, 2, 1), F3(f, i, 3, 2), -(F3(f, i, 2, 2) * F3(f, i, 3, 1)));
    df[i] = fm(F3(f, i, 1, 3), f3, fm(-F3(f, i, 1, 2), f2, F3(f, i, 1, 1) * f1));
  }
  error = 0;
  for (int i = 0; i < span; ++i)
    if (df[i] <= zero_check) error = 1;
}
#undef TH

// ---------------------------------------------------------------------------------------------
// gtmat1.f:1134-1210 (evcmp1_new): Cardano eigenvalues of the metric tensor, ascending
// ---------------------------------------------------------------------------------------------
static void evcmp1_new(int span, const double *c /*(mxvl,6)*/, double *lamda /*(mxvl,3)*/) {
  const double two = 2.0, three = 3.0, onept5 = 1.5, thirteenpt5 = 13.5, twentyseven = 27.0,
               quarter = 0.25, sixpt75 = 6.75, third = 1.0 / 3.0, iroot3 = 1.0 / std::sqrt(3.0);
#define C6(i, k) c[((k)-1) * mxvl + (i)]
  for (int bel = 0; bel < span; ++bel) {
    const double m11 = C6(bel, 1), m12 = C6(bel, 2), m13 = C6(bel, 4), m22 = C6(bel, 3),
                 m23 = C6(bel, 5), m33 = C6(bel, 6);
    const double de = m12 * m23;
    const double dd = m12 * m12;
    const double ee = m23 * m23;
    const double ff = m13 * m13;
    const double m = m11 + m22 + m33;
    const double c1 = fm(m22, m33, fm(m11, m33, m11 * m22)) - (dd + ee + ff);
    const double c0 = fm(-(two * m13), de, fm(-(m11 * m22), m33, fm(m22, ff, fm(m11, ee, m33 * dd))));
    const double p = fm(m, m, -(three * c1));
    const double q = fm(m, fm(-onept5, c1, p), -(thirteenpt5 * c0));
    const double sqrtp = std::sqrt(std::fabs(p));
    double phi = twentyseven * fm(c0, fm(sixpt75, c0, q), quarter * c1 * c1 * (p - c1));
    phi = third * std::atan2(std::sqrt(std::fabs(phi)), q);
    const double cphi = sqrtp * std::cos(phi);
    const double sphi = iroot3 * sqrtp * std::sin(phi);
    double e2 = third * (m - cphi);
    double e3 = e2 + sphi;
    double e1 = e2 + cphi;
    e2 = e2 - sphi;
    if (e2 < e1) { double s = e1; e1 = e2; e2 = s; }
    if (e3 < e1) { double s = e1; e1 = e3; e3 = s; }
    if (e3 < e2) { double s = e2; e2 = e3; e3 = s; }
    lamda[0 * mxvl + bel] = e1;
    lamda[1 * mxvl + bel] = e2;
    lamda[2 * mxvl + bel] = e3;
  }
}

// ---------------------------------------------------------------------------------------------
// gtmat1.f:1031-1119 (ivcmp1) + 961-1014 (irscp1) + 691-723 (rtcmp1): R = F U^-1
// ---------------------------------------------------------------------------------------------
void rtcmp1(int span, const double *f /*(mxvl,3,3)*/, double *r /*(mxvl,3,3)*/) {
  std::vector<double> cv((size_t)mxvl * 6), ccv((size_t)mxvl * 6), ev((size_t)mxvl * 3),
      uiv((size_t)mxvl * 6);
  double *c = cv.data(), *cc = ccv.data(), *ui = uiv.data();
#define CC6(i, k) cc[((k)-1) * mxvl + (i)]
#define CW(i, k) c[((k)-1) * mxvl + (i)]
#define UI(i, k) ui[((k)-1) * mxvl + (i)]
  double iu[mxvl], iiu[mxvl], iiiu[mxvl];
  const double one = 1.0, two = 2.0;
  // ivcmp1
  for (int i = 0; i < span; ++i) {
#define DOT3(a1, b1, a2, b2, a3, b3) fm(a3, b3, fm(a2, b2, (a1) * (b1)))
    CW(i, 1) = DOT3(F3(f, i, 1, 1), F3(f, i, 1, 1), F3(f, i, 2, 1), F3(f, i, 2, 1), F3(f, i, 3, 1), F3(f, i, 3, 1));
    CW(i, 2) = DOT3(F3(f, i, 1, 1), F3(f, i, 1, 2), F3(f, i, 2, 1), F3(f, i, 2, 2), F3(f, i, 3, 1), F3(f, i, 3, 2));
    CW(i, 3) = DOT3(F3(f, i, 1, 2), F3(f, i, 1, 2), F3(f, i, 2, 2), F3(f, i, 2, 2), F3(f, i, 3, 2), F3(f, i, 3, 2));
    CW(i, 4) = DOT3(F3(f, i, 1, 1), F3(f, i, 1, 3), F3(f, i, 2, 1), F3(f, i, 2, 3), F3(f, i, 3, 1), F3(f, i, 3, 3));
    CW(i, 5) = DOT3(F3(f, i, 1, 2), F3(f, i, 1, 3), F3(f, i, 2, 2), F3(f, i, 2, 3), F3(f, i, 3, 2), F3(f, i, 3, 3));
    CW(i, 6) = DOT3(F3(f, i, 1, 3), F3(f, i, 1, 3), F3(f, i, 2, 3), F3(f, i, 2, 3), F3(f, i, 3, 3), F3(f, i, 3, 3));
  }
  for (int i = 0; i < span; ++i) {
    CC6(i, 1) = DOT3(CW(i, 1), CW(i, 1), CW(i, 2), CW(i, 2), CW(i, 4), CW(i, 4));
    CC6(i, 2) = DOT3(CW(i, 1), CW(i, 2), CW(i, 2), CW(i, 3), CW(i, 4), CW(i, 5));
    CC6(i, 3) = DOT3(CW(i, 2), CW(i, 2), CW(i, 3), CW(i, 3), CW(i, 5), CW(i, 5));
    CC6(i, 4) = DOT3(CW(i, 1), CW(i, 4), CW(i, 2), CW(i, 5), CW(i, 4), CW(i, 6));
    CC6(i, 5) = DOT3(CW(i, 2), CW(i, 4), CW(i, 3), CW(i, 5), CW(i, 5), CW(i, 6));
    CC6(i, 6) = DOT3(CW(i, 4), CW(i, 4), CW(i, 5), CW(i, 5), CW(i, 6), CW(i, 6));
  }
  evcmp1_new(span, c, ev.data());
  for (int i = 0; i < span; ++i) {
    ev[0 * mxvl + i] = std::sqrt(ev[0 * mxvl + i]);
    ev[1 * mxvl + i] = std::sqrt(ev[1 * mxvl + i]);
    ev[2 * mxvl + i] = std::sqrt(ev[2 * mxvl + i]);
  }
  for (int i = 0; i < span; ++i) {
    const double e1 = ev[0 * mxvl + i], e2 = ev[1 * mxvl + i], e3 = ev[2 * mxvl + i];
    iu[i] = e1 + e2 + e3;
    iiu[i] = DOT3(e1, e2, e2, e3, e1, e3);
    iiiu[i] = e1 * e2 * e3;
  }
  // irscp1
  for (int i = 0; i < span; ++i) {
    const double a2 = one / (iiiu[i] * fm(iu[i], iiu[i], -iiiu[i]));
    const double b2 = fm(iu[i] * iiu[i], iiu[i], -(iiiu[i] * fm(iu[i], iu[i], iiu[i])));
    const double c2 = fm(-iu[i], fm(iu[i], iu[i], -(two * iiu[i])), -iiiu[i]);
    const double d2 = iu[i];
    UI(i, 1) = a2 * fm(d2, CC6(i, 1), fm(c2, CW(i, 1), b2));
    UI(i, 2) = a2 * fm(d2, CC6(i, 2), c2 * CW(i, 2));
    UI(i, 3) = a2 * fm(d2, CC6(i, 3), fm(c2, CW(i, 3), b2));
    UI(i, 4) = a2 * fm(d2, CC6(i, 4), c2 * CW(i, 4));
    UI(i, 5) = a2 * fm(d2, CC6(i, 5), c2 * CW(i, 5));
    UI(i, 6) = a2 * fm(d2, CC6(i, 6), fm(c2, CW(i, 6), b2));
  }
  // rtcmp1
  for (int i = 0; i < span; ++i) {
    F3(r, i, 1, 1) = DOT3(F3(f, i, 1, 1), UI(i, 1), F3(f, i, 1, 2), UI(i, 2), F3(f, i, 1, 3), UI(i, 4));
    F3(r, i, 1, 2) = DOT3(F3(f, i, 1, 1), UI(i, 2), F3(f, i, 1, 2), UI(i, 3), F3(f, i, 1, 3), UI(i, 5));
    F3(r, i, 1, 3) = DOT3(F3(f, i, 1, 1), UI(i, 4), F3(f, i, 1, 2), UI(i, 5), F3(f, i, 1, 3), UI(i, 6));
    F3(r, i, 2, 1) = DOT3(F3(f, i, 2, 1), UI(i, 1), F3(f, i, 2, 2), UI(i, 2), F3(f, i, 2, 3), UI(i, 4));
    F3(r, i, 2, 2) = DOT3(F3(f, i, 2, 1), UI(i, 2), F3(f, i, 2, 2), UI(i, 3), F3(f, i, 2, 3), UI(i, 5));
    F3(r, i, 2, 3) = DOT3(F3(f, i, 2, 1), UI(i, 4), F3(f, i, 2, 2), UI(i, 5), F3(f, i, 2, 3), UI(i, 6));
    F3(r, i, 3, 1) = DOT3(F3(f, i, 3, 1), UI(i, 1), F3(f, i, 3, 2), UI(i, 2), F3(f, i, 3, 3), UI(i, 4));
    F3(r, i, 3, 2) = DOT3(F3(f, i, 3, 1), UI(i, 2), F3(f, i, 3, 2), UI(i, 3), F3(f, i, 3, 3), UI(i, 5));
    F3(r, i, 3, 3) = DOT3(F3(f, i, 3, 1), UI(i, 4), F3(f, i, 3, 2), UI(i, 5), F3(f, i, 3, 3), UI(i, 6));
#undef DOT3
  }
#undef CC6
#undef CW
#undef UI
#undef C6
}

// ---------------------------------------------------------------------------------------------
// gtmat1.f:740-943 (getrm1), opt 1 (d = q D) and opt 2 (T = q t); opt 3 = opt 2 on R^T
// ---------------------------------------------------------------------------------------------
void getrm1(int span, double *q /*(mxvl,6,6)*/, const double *rin /*(mxvl,3,3)*/, int opt) {
  const double two = 2.0;
  std::vector<double> rbar;
  const double *r = rin;
  if (opt == 3) {
    rbar.resize((size_t)mxvl * 9);
    for (int i = 0; i < span; ++i)
      for (int a = 1; a <= 3; ++a)
        for (int b = 1; b <= 3; ++b) F3(rbar, i, a, b) = F3(rin, i, b, a);
    r = rbar.data();
    opt = 2;
  }
#define R(i, a, b) F3(r, i, a, b)
  if (opt == 1) {
    for (int i = 0; i < span; ++i) {
      Q6(q, i, 1, 1) = R(i, 1, 1) * R(i, 1, 1);
      Q6(q, i, 1, 2) = R(i, 2, 1) * R(i, 2, 1);
      Q6(q, i, 1, 3) = R(i, 3, 1) * R(i, 3, 1);
      Q6(q, i, 1, 4) = R(i, 1, 1) * R(i, 2, 1);
      Q6(q, i, 1, 5) = R(i, 3, 1) * R(i, 2, 1);
      Q6(q, i, 1, 6) = R(i, 1, 1) * R(i, 3, 1);
      Q6(q, i, 2, 1) = R(i, 1, 2) * R(i, 1, 2);
      Q6(q, i, 2, 2) = R(i, 2, 2) * R(i, 2, 2);
      Q6(q, i, 2, 3) = R(i, 3, 2) * R(i, 3, 2);
      Q6(q, i, 2, 4) = R(i, 1, 2) * R(i, 2, 2);
      Q6(q, i, 2, 5) = R(i, 3, 2) * R(i, 2, 2);
      Q6(q, i, 2, 6) = R(i, 1, 2) * R(i, 3, 2);
      Q6(q, i, 3, 1) = R(i, 1, 3) * R(i, 1, 3);
      Q6(q, i, 3, 2) = R(i, 2, 3) * R(i, 2, 3);
      Q6(q, i, 3, 3) = R(i, 3, 3) * R(i, 3, 3);
      Q6(q, i, 3, 4) = R(i, 1, 3) * R(i, 2, 3);
      Q6(q, i, 3, 5) = R(i, 3, 3) * R(i, 2, 3);
      Q6(q, i, 3, 6) = R(i, 1, 3) * R(i, 3, 3);
      Q6(q, i, 4, 1) = two * R(i, 1, 1) * R(i, 1, 2);
      Q6(q, i, 4, 2) = two * R(i, 2, 1) * R(i, 2, 2);
      Q6(q, i, 4, 3) = two * R(i, 3, 1) * R(i, 3, 2);
      Q6(q, i, 4, 4) = R(i, 1, 1) * R(i, 2, 2) + R(i, 1, 2) * R(i, 2, 1);
      Q6(q, i, 4, 5) = R(i, 2, 1) * R(i, 3, 2) + R(i, 3, 1) * R(i, 2, 2);
      Q6(q, i, 4, 6) = R(i, 1, 1) * R(i, 3, 2) + R(i, 3, 1) * R(i, 1, 2);
      Q6(q, i, 5, 1) = two * R(i, 1, 2) * R(i, 1, 3);
      Q6(q, i, 5, 2) = two * R(i, 2, 3) * R(i, 2, 2);
      Q6(q, i, 5, 3) = two * R(i, 3, 2) * R(i, 3, 3);
      Q6(q, i, 5, 4) = R(i, 1, 2) * R(i, 2, 3) + R(i, 2, 2) * R(i, 1, 3);
      Q6(q, i, 5, 5) = R(i, 2, 2) * R(i, 3, 3) + R(i, 2, 3) * R(i, 3, 2);
      Q6(q, i, 5, 6) = R(i, 1, 2) * R(i, 3, 3) + R(i, 3, 2) * R(i, 1, 3);
      Q6(q, i, 6, 1) = two * R(i, 1, 1) * R(i, 1, 3);
      Q6(q, i, 6, 2) = two * R(i, 2, 1) * R(i, 2, 3);
      Q6(q, i, 6, 3) = two * R(i, 3, 1) * R(i, 3, 3);
      Q6(q, i, 6, 4) = R(i, 1, 1) * R(i, 2, 3) + R(i, 2, 1) * R(i, 1, 3);
      Q6(q, i, 6, 5) = R(i, 2, 1) * R(i, 3, 3) + R(i, 3, 1) * R(i, 2, 3);
      Q6(q, i, 6, 6) = R(i, 1, 1) * R(i, 3, 3) + R(i, 1, 3) * R(i, 3, 1);
    }
    return;
  }
  // opt 2
  for (int i = 0; i < span; ++i) {
    Q6(q, i, 1, 1) = R(i, 1, 1) * R(i, 1, 1);
    Q6(q, i, 1, 2) = R(i, 1, 2) * R(i, 1, 2);
    Q6(q, i, 1, 3) = R(i, 1, 3) * R(i, 1, 3);
    Q6(q, i, 1, 4) = two * R(i, 1, 1) * R(i, 1, 2);
    Q6(q, i, 1, 5) = two * R(i, 1, 3) * R(i, 1, 2);
    Q6(q, i, 1, 6) = two * R(i, 1, 1) * R(i, 1, 3);
    Q6(q, i, 2, 1) = R(i, 2, 1) * R(i, 2, 1);
    Q6(q, i, 2, 2) = R(i, 2, 2) * R(i, 2, 2);
    Q6(q, i, 2, 3) = R(i, 2, 3) * R(i, 2, 3);
    Q6(q, i, 2, 4) = two * R(i, 2, 1) * R(i, 2, 2);
    Q6(q, i, 2, 5) = two * R(i, 2, 3) * R(i, 2, 2);
    Q6(q, i, 2, 6) = two * R(i, 2, 1) * R(i, 2, 3);
    Q6(q, i, 3, 1) = R(i, 3, 1) * R(i, 3, 1);
    Q6(q, i, 3, 2) = R(i, 3, 2) * R(i, 3, 2);
    Q6(q, i, 3, 3) = R(i, 3, 3) * R(i, 3, 3);
    Q6(q, i, 3, 4) = two * R(i, 3, 1) * R(i, 3, 2);
    Q6(q, i, 3, 5) = two * R(i, 3, 3) * R(i, 3, 2);
    Q6(q, i, 3, 6) = two * R(i, 3, 1) * R(i, 3, 3);
    Q6(q, i, 4, 1) = R(i, 1, 1) * R(i, 2, 1);
    Q6(q, i, 4, 2) = R(i, 1, 2) * R(i, 2, 2);
    Q6(q, i, 4, 3) = R(i, 1, 3) * R(i, 2, 3);
    Q6(q, i, 4, 4) = R(i, 1, 1) * R(i, 2, 2) + R(i, 2, 1) * R(i, 1, 2);
    Q6(q, i, 4, 5) = R(i, 1, 2) * R(i, 2, 3) + R(i, 1, 3) * R(i, 2, 2);
    Q6(q, i, 4, 6) = R(i, 1, 1) * R(i, 2, 3) + R(i, 1, 3) * R(i, 2, 1);
    Q6(q, i, 5, 1) = R(i, 2, 1) * R(i, 3, 1);
    Q6(q, i, 5, 2) = R(i, 3, 2) * R(i, 2, 2);
    Q6(q, i, 5, 3) = R(i, 2, 3) * R(i, 3, 3);
    Q6(q, i, 5, 4) = R(i, 2, 1) * R(i, 3, 2) + R(i, 2, 2) * R(i, 3, 1);
    Q6(q, i, 5, 5) = R(i, 2, 2) * R(i, 3, 3) + R(i, 3, 2) * R(i, 2, 3);
    Q6(q, i, 5, 6) = R(i, 2, 1) * R(i, 3, 3) + R(i, 2, 3) * R(i, 3, 1);
    Q6(q, i, 6, 1) = R(i, 1, 1) * R(i, 3, 1);
    Q6(q, i, 6, 2) = R(i, 1, 2) * R(i, 3, 2);
    Q6(q, i, 6, 3) = R(i, 1, 3) * R(i, 3, 3);
    Q6(q, i, 6, 4) = R(i, 1, 1) * R(i, 3, 2) + R(i, 1, 2) * R(i, 3, 1);
    Q6(q, i, 6, 5) = R(i, 1, 2) * R(i, 3, 3) + R(i, 1, 3) * R(i, 3, 2);
    Q6(q, i, 6, 6) = R(i, 1, 1) * R(i, 3, 3) + R(i, 3, 1) * R(i, 1, 3);
  }
#undef R
}

// ---------------------------------------------------------------------------------------------
// qmply1.f:15-49 : m2 = q * m1
// ---------------------------------------------------------------------------------------------
void qmply1(int span, const double *q, const double *m1 /*(mxvl,*)*/, double *m2 /*(mxvl,*)*/) {
#define M1(i, k) m1[((k)-1) * mxvl + (i)]
#define M2(i, k) m2[((k)-1) * mxvl + (i)]
  for (int j = 1; j <= 6; ++j)
    for (int i = 0; i < span; ++i)
      M2(i, j) = Q6(q, i, j, 1) * M1(i, 1) + Q6(q, i, j, 2) * M1(i, 2) + Q6(q, i, j, 3) * M1(i, 3) +
                 Q6(q, i, j, 4) * M1(i, 4) + Q6(q, i, j, 5) * M1(i, 5) + Q6(q, i, j, 6) * M1(i, 6);
#undef M1
#undef M2
}

// ---------------------------------------------------------------------------------------------
// blcmp1.f:15-81 : linear strain-displacement matrix b(mxvl,mxedof,nstr)
// ---------------------------------------------------------------------------------------------
void blcmp1(int span, double *b, const double *gama, const double *nxi, const double *neta,
            const double *nzeta, int nnode) {
  std::vector<double> btemp((size_t)mxvl * mxndel * 3);
#define BT(i, j, k) btemp[(((k)-1) * mxndel + ((j)-1)) * mxvl + (i)]
  for (int j = 1; j <= nnode; ++j)
    for (int i = 0; i < span; ++i) {
      BT(i, j, 1) = GAM(i, 1, 1) * nxi[j - 1] + GAM(i, 1, 2) * neta[j - 1] + GAM(i, 1, 3) * nzeta[j - 1];
      BT(i, j, 2) = GAM(i, 2, 1) * nxi[j - 1] + GAM(i, 2, 2) * neta[j - 1] + GAM(i, 2, 3) * nzeta[j - 1];
      BT(i, j, 3) = GAM(i, 3, 1) * nxi[j - 1] + GAM(i, 3, 2) * neta[j - 1] + GAM(i, 3, 3) * nzeta[j - 1];
    }
  const int bpos1 = nnode, bpos2 = 2 * nnode;
  for (int j = 1; j <= nnode; ++j)
    for (int i = 0; i < span; ++i) {
      BM(b, i, j, 1) = BT(i, j, 1);
      BM(b, i, j, 4) = BT(i, j, 2);
      BM(b, i, j, 6) = BT(i, j, 3);
      BM(b, i, bpos1 + j, 2) = BT(i, j, 2);
      BM(b, i, bpos1 + j, 4) = BT(i, j, 1);
      BM(b, i, bpos1 + j, 5) = BT(i, j, 3);
      BM(b, i, bpos2 + j, 3) = BT(i, j, 3);
      BM(b, i, bpos2 + j, 5) = BT(i, j, 2);
      BM(b, i, bpos2 + j, 6) = BT(i, j, 1);
    }
#undef BT
}

// ---------------------------------------------------------------------------------------------
// bmod.f:16-188 : mean-dilatation B-bar modification (hex8), statement order preserved
// ---------------------------------------------------------------------------------------------
void bmod(double *b, const double *vol, int span, double eps_stab) {
  const double one = 1.0, two = 2.0, third = 1.0 / 3.0;
  const double alpha = (two + eps_stab) * third;
  const double beta = (one - eps_stab) * third;
  for (int i = 0; i < span; ++i) {
    double dummy[9];
    for (int j = 1; j <= 8; ++j) dummy[j] = BM(b, i, j, 1);
    for (int j = 1; j <= 8; ++j) BM(b, i, j, 1) = alpha * dummy[j] + beta * VOL(i, j, 1);
    for (int j = 1; j <= 8; ++j) BM(b, i, j, 2) = beta * (VOL(i, j, 1) - dummy[j]);
  }
  for (int i = 0; i < span; ++i) {
    for (int j = 1; j <= 8; ++j) BM(b, i, 8 + j, 1) = beta * (VOL(i, j, 2) - BM(b, i, 8 + j, 2));
    for (int j = 1; j <= 8; ++j) BM(b, i, 16 + j, 1) = beta * (VOL(i, j, 3) - BM(b, i, 16 + j, 3));
    for (int j = 1; j <= 8; ++j) BM(b, i, 8 + j, 2) = alpha * BM(b, i, 8 + j, 2) + beta * VOL(i, j, 2);
    for (int j = 1; j <= 8; ++j) BM(b, i, 16 + j, 2) = BM(b, i, 16 + j, 1);
    for (int j = 1; j <= 8; ++j) BM(b, i, j, 3) = BM(b, i, j, 2);
    for (int j = 1; j <= 8; ++j) BM(b, i, 8 + j, 3) = BM(b, i, 8 + j, 1);
    for (int j = 1; j <= 8; ++j) BM(b, i, 16 + j, 3) = alpha * BM(b, i, 16 + j, 3) + beta * VOL(i, j, 3);
  }
}

// ---------------------------------------------------------------------------------------------
// gtlsn1.f:25-82 : deps = B * due, exploiting sparsity exactly as the reference does
// ---------------------------------------------------------------------------------------------
void gtlsn1(int span, int nnode, const double *due, double *deps /*(mxvl,nstr)*/,
            const double *gama, const double *nxi, const double *neta, const double *nzeta,
            const double *vol_block, bool bbar, double eps_bbar, double *b) {
#define DUE(i, j) due[((j)-1) * mxvl + (i)]
#define DEPS(i, k) deps[((k)-1) * mxvl + (i)]
  blcmp1(span, b, gama, nxi, neta, nzeta, nnode);
  if (bbar) bmod(b, vol_block, span, eps_bbar);
  for (int k = 0; k < nstr * mxvl; ++k) deps[k] = 0.0;
  const int bpos1 = nnode, bpos2 = 2 * nnode;
  for (int j = 1; j <= nnode; ++j)
    for (int i = 0; i < span; ++i) {
      DEPS(i, 1) = DEPS(i, 1) + BM(b, i, j, 1) * DUE(i, j) + BM(b, i, bpos1 + j, 1) * DUE(i, bpos1 + j) +
                   BM(b, i, bpos2 + j, 1) * DUE(i, bpos2 + j);
      DEPS(i, 2) = DEPS(i, 2) + BM(b, i, j, 2) * DUE(i, j) + BM(b, i, bpos1 + j, 2) * DUE(i, bpos1 + j) +
                   BM(b, i, bpos2 + j, 2) * DUE(i, bpos2 + j);
      DEPS(i, 3) = DEPS(i, 3) + BM(b, i, j, 3) * DUE(i, j) + BM(b, i, bpos1 + j, 3) * DUE(i, bpos1 + j) +
                   BM(b, i, bpos2 + j, 3) * DUE(i, bpos2 + j);
      DEPS(i, 4) = DEPS(i, 4) + BM(b, i, j, 4) * DUE(i, j) + BM(b, i, bpos1 + j, 4) * DUE(i, bpos1 + j);
      DEPS(i, 5) = DEPS(i, 5) + BM(b, i, bpos1 + j, 5) * DUE(i, bpos1 + j) +
                   BM(b, i, bpos2 + j, 5) * DUE(i, bpos2 + j);
      DEPS(i, 6) = DEPS(i, 6) + BM(b, i, j, 6) * DUE(i, j) + BM(b, i, bpos2 + j, 6) * DUE(i, bpos2 + j);
    }
#undef DUE
#undef DEPS
}

}  // namespace w3do
