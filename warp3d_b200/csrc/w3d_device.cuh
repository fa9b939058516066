// w3d_device.cuh -- device-side building blocks of the B200 element-block engine (sm_100a, FP64).
//
// Work decomposition (hex8, 2x2x2): one thread per (element, Gauss point); the 8 Gauss points of
// an element sit in 8 consecutive lanes, a warp owns 4 whole elements, so every per-element
// exchange (B-bar means, B^T sigma sums) is a sub-warp shuffle.  State lives in HBM as
// structure-of-arrays  a[component][element_slot*8 + gp]  so each warp access is one 256 B line.
//
// Numerical contract (see DESIGN.md "parity"):  the chain  coordinates/displacements -> Gamma_0 ->
// F -> C -> Cardano eigenvalues -> U^-1 -> R  is evaluated with explicitly rounded operations in
// the reference's statement order (gtmat1.f:300-507, 522-674, 1031-1210), because the
// reference's closed-form polar decomposition amplifies a 1-ulp change of C into ~1e-9 of R for
// strains below ~5e-3.  Products and sums are fused in ONE fixed pattern -- left-to-right
// contraction, a*b + c*d + e*f -> fma(e, f, fma(c, d, a*b)), a*b - c*d -> fma(a, b, -(c*d)),
// s + a*b -> fma(a, b, s) -- written out with xf() below; the CPU restatement the parity tests
// use states the same pattern (an optimising compiler such as the reference's own contracts
// these expressions too; which products it fuses is not visible from the source, so the pattern
// is a convention either way).  ~450 fewer FP64 instructions per Gauss point than the un-fused
// chain of round 1.  Everything downstream is well conditioned and uses FMA freely.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace w3d {

constexpr int EPC = 16;          // elements per CTA
constexpr int TPB = EPC * 8;     // threads per CTA (one per Gauss point)
constexpr int EPROP = 16;        // floats per element in the compact property table:
                                 // e, nu, beta, hprime, yld, lnelas, then matprp(55:64) (mm05)

// one CTA = up to EPC consecutive elements of one reference block
struct CtaDesc {
  int slot0;      // first element slot (segment-relative) handled by this CTA
  int nvalid;     // number of real elements (1..EPC)
  int blk_slot0;  // slot of the first element of the owning block (mm01.f:843-845 quirk)
  int epos0;      // position of the CTA's first element in the concatenated, unpadded element
                  // order of all owned blocks (estiff / pull_state addressing)
  int blk_span;   // span of the owning block (mm10 small-strain curvature fit, mm10_a.f:1622)
};

// L2 prefetch of the line holding p (no register, no shared memory, no wait)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// un-contracted, round-to-nearest operations (never fused by ptxas)
__device__ __forceinline__ double xm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xa(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xs(double a, double b) { return __dsub_rn(a, b); }
// the fused operation of the contraction convention (never split or re-associated by ptxas)
__device__ __forceinline__ double xf(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double xdot3(double a1, double b1, double a2, double b2, double a3, double b3) {
  return xf(a3, b3, xf(a2, b2, xm(a1, b1)));
}
// a / b, correctly rounded like __ddiv_rn, without the division's ~65-instruction subroutine for
// a == 0.  That subroutine runs whenever the quotient is zero or subnormal, and zero numerators
// are the rule, not the exception, on this path: the off-diagonal cofactors of an axis-aligned
// element, every back-stress term of an isotropically hardening material.  +-0 / b = +-0 * (1 / b)
// in every IEEE case (b = 0 and NaN give NaN, b = inf gives the signed zero) except |b| subnormal.
__device__ __forceinline__ double xdiv(double a, double b) {
  const bool z = (a == 0.0);
  const double q = __ddiv_rn(z ? 1.0 : a, b);
  return z ? __dmul_rn(a, q) : q;
}

// ------------------------------------------------------------------------------------------------
// Jacobian at one Gauss point.  jacob1 (gtmat1.f:300-507), 3-D branch.  X,Y,Z: the element's 8
// nodal coordinates; nx,ne,nz: parent-space shape-function derivatives at this point.
// G[r][c] = gama(r,c) = inverse Jacobian, dj = det J.
// EXACT = reference statement order without contraction.
// ------------------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void jacobian(const double (&nx)[8], const double (&ne)[8],
                                         const double (&nz)[8], const double (&X)[8],
                                         const double (&Y)[8], const double (&Z)[8],
                                         double (&G)[3][3], double &dj) {
  double J[3][3];
  if (EXACT) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      J[0][0] = xf(nx[j], X[j], J[0][0]);
      J[0][1] = xf(nx[j], Y[j], J[0][1]);
      J[0][2] = xf(nx[j], Z[j], J[0][2]);
      J[1][0] = xf(ne[j], X[j], J[1][0]);
      J[1][1] = xf(ne[j], Y[j], J[1][1]);
      J[1][2] = xf(ne[j], Z[j], J[1][2]);
      J[2][0] = xf(nz[j], X[j], J[2][0]);
      J[2][1] = xf(nz[j], Y[j], J[2][1]);
      J[2][2] = xf(nz[j], Z[j], J[2][2]);
    }
    const double j1 = xf(J[1][1], J[2][2], -xm(J[1][2], J[2][1]));
    const double j2 = xf(J[1][0], J[2][2], -xm(J[1][2], J[2][0]));
    const double j3 = xf(J[1][0], J[2][1], -xm(J[1][1], J[2][0]));
    dj = xf(J[0][2], j3, xf(-J[0][1], j2, xm(J[0][0], j1)));
    const double rdj = __ddiv_rn(1.0, dj);      // one reciprocal, nine products (convention)
    G[0][0] = xm(j1, rdj);
    G[1][0] = xm(-j2, rdj);
    G[2][0] = xm(j3, rdj);
    G[0][1] = xm(xf(J[2][1], J[0][2], -xm(J[0][1], J[2][2])), rdj);
    G[1][1] = xm(xf(J[0][0], J[2][2], -xm(J[2][0], J[0][2])), rdj);
    G[2][1] = xm(xf(J[0][1], J[2][0], -xm(J[0][0], J[2][1])), rdj);
    G[0][2] = xm(xf(J[0][1], J[1][2], -xm(J[0][2], J[1][1])), rdj);
    G[1][2] = xm(xf(J[0][2], J[1][0], -xm(J[0][0], J[1][2])), rdj);
    G[2][2] = xm(xf(J[0][0], J[1][1], -xm(J[0][1], J[1][0])), rdj);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double (&P)[8] = (c == 0) ? X : (c == 1) ? Y : Z;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0 = fma(nx[j], P[j], a0);
        a1 = fma(ne[j], P[j], a1);
        a2 = fma(nz[j], P[j], a2);
      }
      J[0][c] = a0; J[1][c] = a1; J[2][c] = a2;
    }
    const double j1 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double j2 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
    const double j3 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    dj = J[0][0] * j1 - J[0][1] * j2 + J[0][2] * j3;
    const double r = 1.0 / dj;
    G[0][0] = j1 * r;
    G[1][0] = -j2 * r;
    G[2][0] = j3 * r;
    G[0][1] = (J[2][1] * J[0][2] - J[0][1] * J[2][2]) * r;
    G[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * r;
    G[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    G[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    G[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    G[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  }
}

// ------------------------------------------------------------------------------------------------
// R = F U^-1 by the reference's closed form: ivcmp1 (gtmat1.f:1031-1119), evcmp1_new (:1134-1210,
// Cardano), irscp1 (:961-1014), rtcmp1 (:691-723).  Un-contracted, statement order preserved.
// F[r][c], R[r][c].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void polar_rotation(const double (&F)[3][3], double (&R)[3][3]) {
  const double two = 2.0, three = 3.0, onept5 = 1.5, thirteenpt5 = 13.5, twentyseven = 27.0,
               quarter = 0.25, sixpt75 = 6.75, one = 1.0;
  const double third = 1.0 / 3.0;                          // constants module: one/three
  const double iroot3 = 0x1.279a74590331dp-1;              // one/sqrt(three) as IEEE evaluates it
  // metric tensor C = F^T F (symmetric storage 11,12,22,13,23,33) and C*C
  const double c1 = xdot3(F[0][0], F[0][0], F[1][0], F[1][0], F[2][0], F[2][0]);
  const double c2 = xdot3(F[0][0], F[0][1], F[1][0], F[1][1], F[2][0], F[2][1]);
  const double c3 = xdot3(F[0][1], F[0][1], F[1][1], F[1][1], F[2][1], F[2][1]);
  const double c4 = xdot3(F[0][0], F[0][2], F[1][0], F[1][2], F[2][0], F[2][2]);
  const double c5 = xdot3(F[0][1], F[0][2], F[1][1], F[1][2], F[2][1], F[2][2]);
  const double c6 = xdot3(F[0][2], F[0][2], F[1][2], F[1][2], F[2][2], F[2][2]);
  const double cc1 = xdot3(c1, c1, c2, c2, c4, c4);
  const double cc2 = xdot3(c1, c2, c2, c3, c4, c5);
  const double cc3 = xdot3(c2, c2, c3, c3, c5, c5);
  const double cc4 = xdot3(c1, c4, c2, c5, c4, c6);
  const double cc5 = xdot3(c2, c4, c3, c5, c5, c6);
  const double cc6 = xdot3(c4, c4, c5, c5, c6, c6);
  // Cardano eigenvalues of C
  const double m11 = c1, m12 = c2, m13 = c4, m22 = c3, m23 = c5, m33 = c6;
  const double de = xm(m12, m23);
  const double dd = xm(m12, m12);
  const double ee = xm(m23, m23);
  const double ff = xm(m13, m13);
  const double m = xa(xa(m11, m22), m33);
  const double cc1e = xs(xf(m22, m33, xf(m11, m33, xm(m11, m22))), xa(xa(dd, ee), ff));
  const double c0 = xf(-xm(two, m13), de, xf(-xm(m11, m22), m33, xf(m22, ff, xf(m11, ee, xm(m33, dd)))));
  const double p = xf(m, m, -xm(three, cc1e));
  const double q = xf(m, xf(-onept5, cc1e, p), -xm(thirteenpt5, c0));
  const double sqrtp = sqrt(fabs(p));
  double phi = xm(twentyseven, xf(c0, xf(sixpt75, c0, q), xm(xm(xm(quarter, cc1e), cc1e), xs(p, cc1e))));
  // the argument is a difference of nearly equal numbers and comes out exactly zero at a good
  // share of the points: atan2(+0, q) is 0 or pi by the sign of q -- the same bits without the
  // slow paths of sqrt(0) and of the zero quotient inside atan2
  if (phi == 0.0 && q == q)
    phi = (q > 0.0 || (q == 0.0 && !signbit(q))) ? 0.0 : xm(third, 3.141592653589793116);
  else
    phi = xm(third, atan2(sqrt(fabs(phi)), q));
  double sn, cs;
  sincos(phi, &sn, &cs);
  const double cphi = xm(sqrtp, cs);
  const double sphi = xm(xm(iroot3, sqrtp), sn);
  double e2 = xm(third, xs(m, cphi));
  double e3 = xa(e2, sphi);
  double e1 = xa(e2, cphi);
  e2 = xs(e2, sphi);
  if (e2 < e1) { const double s = e1; e1 = e2; e2 = s; }
  if (e3 < e1) { const double s = e1; e1 = e3; e3 = s; }
  if (e3 < e2) { const double s = e2; e2 = e3; e3 = s; }
  e1 = sqrt(e1); e2 = sqrt(e2); e3 = sqrt(e3);
  const double iu = xa(xa(e1, e2), e3);
  const double iiu = xdot3(e1, e2, e2, e3, e1, e3);
  const double iiiu = xm(xm(e1, e2), e3);
  // irscp1
  const double a2 = __ddiv_rn(one, xm(iiiu, xf(iu, iiu, -iiiu)));
  const double b2 = xf(xm(iu, iiu), iiu, -xm(iiiu, xf(iu, iu, iiu)));
  const double c2m = xf(-iu, xf(iu, iu, -xm(two, iiu)), -iiiu);
  const double d2 = iu;
  const double u1 = xm(a2, xf(d2, cc1, xf(c2m, c1, b2)));
  const double u2 = xm(a2, xf(d2, cc2, xm(c2m, c2)));
  const double u3 = xm(a2, xf(d2, cc3, xf(c2m, c3, b2)));
  const double u4 = xm(a2, xf(d2, cc4, xm(c2m, c4)));
  const double u5 = xm(a2, xf(d2, cc5, xm(c2m, c5)));
  const double u6 = xm(a2, xf(d2, cc6, xf(c2m, c6, b2)));
  // rtcmp1: R = F * Ui
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    R[r][0] = xdot3(F[r][0], u1, F[r][1], u2, F[r][2], u4);
    R[r][1] = xdot3(F[r][0], u2, F[r][1], u3, F[r][2], u5);
    R[r][2] = xdot3(F[r][0], u4, F[r][1], u5, F[r][2], u6);
  }
}

// getrm1 (gtmat1.f:740-943).  q[r][c] 6x6 from R[a][b].
// opt 1: unrotated rate of deformation  d = q D ;  opt 2: Cauchy stress  T = q t.
__device__ __forceinline__ void getrm1_opt1(const double (&R)[3][3], double (&q)[6][6]) {
  const double two = 2.0;
  q[0][0] = R[0][0] * R[0][0]; q[0][1] = R[1][0] * R[1][0]; q[0][2] = R[2][0] * R[2][0];
  q[0][3] = R[0][0] * R[1][0]; q[0][4] = R[2][0] * R[1][0]; q[0][5] = R[0][0] * R[2][0];
  q[1][0] = R[0][1] * R[0][1]; q[1][1] = R[1][1] * R[1][1]; q[1][2] = R[2][1] * R[2][1];
  q[1][3] = R[0][1] * R[1][1]; q[1][4] = R[2][1] * R[1][1]; q[1][5] = R[0][1] * R[2][1];
  q[2][0] = R[0][2] * R[0][2]; q[2][1] = R[1][2] * R[1][2]; q[2][2] = R[2][2] * R[2][2];
  q[2][3] = R[0][2] * R[1][2]; q[2][4] = R[2][2] * R[1][2]; q[2][5] = R[0][2] * R[2][2];
  q[3][0] = two * R[0][0] * R[0][1]; q[3][1] = two * R[1][0] * R[1][1];
  q[3][2] = two * R[2][0] * R[2][1];
  q[3][3] = R[0][0] * R[1][1] + R[0][1] * R[1][0];
  q[3][4] = R[1][0] * R[2][1] + R[2][0] * R[1][1];
  q[3][5] = R[0][0] * R[2][1] + R[2][0] * R[0][1];
  q[4][0] = two * R[0][1] * R[0][2]; q[4][1] = two * R[1][2] * R[1][1];
  q[4][2] = two * R[2][1] * R[2][2];
  q[4][3] = R[0][1] * R[1][2] + R[1][1] * R[0][2];
  q[4][4] = R[1][1] * R[2][2] + R[1][2] * R[2][1];
  q[4][5] = R[0][1] * R[2][2] + R[2][1] * R[0][2];
  q[5][0] = two * R[0][0] * R[0][2]; q[5][1] = two * R[1][0] * R[1][2];
  q[5][2] = two * R[2][0] * R[2][2];
  q[5][3] = R[0][0] * R[1][2] + R[1][0] * R[0][2];
  q[5][4] = R[1][0] * R[2][2] + R[2][0] * R[1][2];
  q[5][5] = R[0][0] * R[2][2] + R[0][2] * R[2][0];
}
__device__ __forceinline__ void getrm1_opt2(const double (&R)[3][3], double (&q)[6][6]) {
  const double two = 2.0;
  q[0][0] = R[0][0] * R[0][0]; q[0][1] = R[0][1] * R[0][1]; q[0][2] = R[0][2] * R[0][2];
  q[0][3] = two * R[0][0] * R[0][1]; q[0][4] = two * R[0][2] * R[0][1];
  q[0][5] = two * R[0][0] * R[0][2];
  q[1][0] = R[1][0] * R[1][0]; q[1][1] = R[1][1] * R[1][1]; q[1][2] = R[1][2] * R[1][2];
  q[1][3] = two * R[1][0] * R[1][1]; q[1][4] = two * R[1][2] * R[1][1];
  q[1][5] = two * R[1][0] * R[1][2];
  q[2][0] = R[2][0] * R[2][0]; q[2][1] = R[2][1] * R[2][1]; q[2][2] = R[2][2] * R[2][2];
  q[2][3] = two * R[2][0] * R[2][1]; q[2][4] = two * R[2][2] * R[2][1];
  q[2][5] = two * R[2][0] * R[2][2];
  q[3][0] = R[0][0] * R[1][0]; q[3][1] = R[0][1] * R[1][1]; q[3][2] = R[0][2] * R[1][2];
  q[3][3] = R[0][0] * R[1][1] + R[1][0] * R[0][1];
  q[3][4] = R[0][1] * R[1][2] + R[0][2] * R[1][1];
  q[3][5] = R[0][0] * R[1][2] + R[0][2] * R[1][0];
  q[4][0] = R[1][0] * R[2][0]; q[4][1] = R[2][1] * R[1][1]; q[4][2] = R[1][2] * R[2][2];
  q[4][3] = R[1][0] * R[2][1] + R[1][1] * R[2][0];
  q[4][4] = R[1][1] * R[2][2] + R[2][1] * R[1][2];
  q[4][5] = R[1][0] * R[2][2] + R[1][2] * R[2][0];
  q[5][0] = R[0][0] * R[2][0]; q[5][1] = R[0][1] * R[2][1]; q[5][2] = R[0][2] * R[2][2];
  q[5][3] = R[0][0] * R[2][1] + R[0][1] * R[2][0];
  q[5][4] = R[0][1] * R[2][2] + R[0][2] * R[2][1];
  q[5][5] = R[0][0] * R[2][2] + R[2][0] * R[0][2];
}

// qmply1 (qmply1.f:15-49): m2 = q m1
__device__ __forceinline__ void qmply(const double (&q)[6][6], const double (&a)[6],
                                      double (&b)[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j)
    b[j] = q[j][0] * a[0] + q[j][1] * a[1] + q[j][2] * a[2] + q[j][3] * a[3] + q[j][4] * a[4] +
           q[j][5] * a[5];
}

// integer state kept in the low word of a history double (mm01.f:248-251 EQUIVALENCE)
__device__ __forceinline__ double state_to_dword(int st) { return __hiloint2double(0, st); }
__device__ __forceinline__ int dword_to_state(double d) { return __double2loint(d); }

// sum over the 8 Gauss-point lanes of an element (all 8 lanes get the result)
__device__ __forceinline__ double sum8(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace w3d
