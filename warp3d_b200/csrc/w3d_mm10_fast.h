// w3d_mm10_fast.h -- crystal-plasticity material update (mm10, Voce hardening) for ONE Gauss point,
// the engine's fast path: one crystal type with at most 12 slip systems (fcc, bcc12, ...), any
// number of crystals per point.  Same algorithm as the reference (mm10_a.f: mm10_solve_strup :2657,
// mm10_solve :2901 with its predictor :3032 and update :3224 and their line searches, mm10_tangent
// :664, mm10_output :3493, mm10_update_rotation :3375; mm10_b.f: formR1 :1073, formR2 :63, formJ
// :909, voche routines :1813-2074), organised for a GPU thread:
//
//  * Lattice-frame slip sums.  The rotated Schmid data of a crystal are linear images of the
//    crystal-frame tables,  ms_i = B6 mc_i,  qc_i = B3 qcr_i  with  B6 = RE E6  (RE = rt2rve(Rp_n^T),
//    E6 = the element's crystal->reference map of strain-type vectors) and  B3 = RQ W3.  All sums
//    over slip systems are therefore taken with the CONSTANT tables mc, qcr (operands straight
//    from __constant__ memory, no per-element Schmid records, no per-thread arrays):
//        rs_i = (B6^T T) . mc_i,   dbarp = B6 sum f_i mc_i,   M = B6 (sum g_i mc_i x mc_i) B6^T, ...
//    The same numbers as the reference's per-system evaluation up to re-association (rounding).
//  * One inlined copy of the residual and of the Jacobian: predictor, update, their line searches
//    and the sub-increment restarts run through ONE flat loop (the slip-system loops inside it
//    stay rolled), so the loop body fits the instruction cache.  The stress predictor is the 7x7
//    update with the hardening equation switched off (row / column 7 = identity), which is the
//    same arithmetic as the reference's 6x6 system.
//  * Small per-thread state: the 7x7 system lives in registers (compile-time indices), B6, C B6,
//    B3 and the few vectors that must survive the Jacobian in a per-thread shared-memory column;
//    the stack serves only the out-of-line pivoting LU (lu7_slow), which the usual path never calls.
//  * The tangent  T = JJ^-1 C,  JJ = J11 - J12 J22^-1 J21 (mm10_tangent) is the leading block of
//    J^-1 [C; 0] with J the Jacobian of the last solver pass, no Schur complement: every full
//    Newton pass back-substitutes the six columns with the LU factors it holds and stores them,
//    the last pass's columns are the tangent (crystal_update_fast).  C^-1 for the lattice strain
//    comes from the host.
//
// Compiled by nvcc for the device and by g++ for the host-side point checks under tests/ (which
// compare it with the independent statement-for-statement restatement kept with the tests).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define W3DF_HD __host__ __device__
#else
#define W3DF_HD
#endif
#if defined(__CUDACC__)
#define W3DF_NOINLINE __noinline__
#else
#define W3DF_NOINLINE __attribute__((noinline))
#endif
#if defined(__CUDA_ARCH__)
#define W3DF_UNROLL _Pragma("unroll")
#define W3DF_ROLLED _Pragma("unroll 1")
#else
#define W3DF_UNROLL
#define W3DF_ROLLED
#endif
// slip systems handled per trip of the (rolled) slip-system loops: the loops over the systems
// stay rolled so that residual + Jacobian + LU -- the solver loop -- fit the instruction cache;
// the systems of one trip are independent work for the FP64 pipe.  Measured on the 1 M mesh
// (stress pass): 1 per trip 30.1 ms, 2: 30.5, 4: 33.5 (one system's 6x6 and 3x6 updates already
// fill the pipe; more per trip only lengthens the code)
#ifndef W3DF_CHUNK
#define W3DF_CHUNK 1
#endif

namespace w3d_mm10f {

constexpr int NSF = 12;   // slip-system capacity of the fast path

// one crystal type = the c_array(cnum) fields the path reads, with the slip tables in CRYSTAL axes
struct CrystalConsts {
  double mc[NSF][6];    // ET2EV(sym(b_i n_i^T)): 11 22 33 2*12 2*23 2*13
  double qcr[NSF][3];   // WT2WV(skew(b_i n_i^T)): 23, 13, 12
  double ncr[NSF][3];   // slip-plane normals (geometric hardening)
  double rate_n, tau_y, tau_v, voche_m, theta_0, k_0, burgers, iD_v, eps_dot_0_y;
  double atol, atol1, rtol, rtol1;
  int nslip, miter, alter_mode, angle_type;
  int rate_ie, voche_ie;   // whole-number exponents rate_n - 1 and voche_m (binary powering), or -1
};

// per (element, crystal) record, 128 doubles
struct FProps {
  double g[3][3];      // rotation_g (Kocks angles)
  double E6[6][6];     // strain-type vector map crystal -> reference axes: ms0_i = E6 mc_i
  double W3[3][3];     // spin-vector map: qs0_i = W3 qcr_i
  double C[6][6];      // elastic stiffness in reference axes
  double Cinv[6][6];   // its inverse (lattice strain, mm10_output)
  double pad[2];
};
static_assert(sizeof(FProps) == 128 * sizeof(double), "FProps layout");

#if defined(__CUDACC__)
static __constant__ CrystalConsts c_mm10f;
#ifdef W3D_MM10F_HIST
static __device__ int *w3df_hist;    // debug: histogram of solver-loop trips per crystal update
#endif
#endif
#if defined(__CUDA_ARCH__)
#define W3DF_CT c_mm10f
#else
static const CrystalConsts *g_host_ct = nullptr;   // host-side checks point this at their table
#define W3DF_CT (*g_host_ct)
#endif

// history layout of mm10_set_history_locs (mm10_d.f:153-330) for nslip <= 12, num_hard = 1
struct LayoutF {
  int nslip, common_size, crystal_size;
  W3DF_HD static LayoutF make(int nslip) {
    LayoutF L;
    L.nslip = nslip;
    L.common_size = 36 + 27 + 9 + 3 + nslip;
    L.crystal_size = 6 + 3 + 9 + 6 + 6 + nslip + 1 + 15 + 1 + 6 + 6;
    return L;
  }
  W3DF_HD int size(int nc) const { return common_size + nc * crystal_size; }
  W3DF_HD int o_D() const { return 0; }
  W3DF_HD int o_gradfe() const { return 36; }
  W3DF_HD int o_R() const { return 63; }
  W3DF_HD int o_work() const { return 72; }
  W3DF_HD int o_sliphist() const { return 75; }
  W3DF_HD int c0(int c) const { return common_size + c * crystal_size; }
  W3DF_HD int c_stress(int c) const { return c0(c); }
  W3DF_HD int c_euler(int c) const { return c0(c) + 6; }
  W3DF_HD int c_Rp(int c) const { return c0(c) + 9; }
  W3DF_HD int c_D(int c) const { return c0(c) + 18; }
  W3DF_HD int c_eps(int c) const { return c0(c) + 24; }
  W3DF_HD int c_slip(int c) const { return c0(c) + 30; }
  W3DF_HD int c_tt(int c) const { return c_slip(c) + nslip; }
  W3DF_HD int c_u(int c) const { return c_tt(c) + 1; }
  W3DF_HD int c_rate(int c) const { return c_u(c) + 15; }
  W3DF_HD int c_ep(int c) const { return c_rate(c) + 1; }
  W3DF_HD int c_ed(int c) const { return c_ep(c) + 6; }
};

// strided views of one Gauss point's history vector
struct HistF {
  double *p;
  long long stride;
  W3DF_HD double &operator[](int k) const { return p[(long long)k * stride]; }
};
struct CHistF {
  const double *p;
  long long stride;
  W3DF_HD double operator[](int k) const { return p[(long long)k * stride]; }
};

// per-thread workspace column (shared memory on the device): element k at ws[k * wss]
// (volatile: every read is a load at its point of use -- otherwise the compiler hoists the 81
// words of B6 / C B6 / B3 into registers across the solver loop and spills them to the stack)
struct WorkF {
  volatile double *ws;
  int wss;
  W3DF_HD volatile double &operator()(int k) const { return ws[k * wss]; }
};
// O_SC: loop-carried scalars of the solver that are touched once per sub-increment or once per
// trip (sub-increment fraction and step, hardening state at the sub-increment start, first
// residual norm): in the column they cost a shared-memory access instead of a register or -- what
// ptxas made of them -- a spill slot served from L2
enum { O_B6 = 0, O_CB = 36, O_B3 = 72, O_K0 = 81, O_OST = 87, O_XP = 93, O_RV = 100, O_SC = 107, O_TL = 111 };
// Inputs of one Gauss point, re-read from memory where they are used (set-up, sub-increment
// start, output) instead of being carried in registers through the solver loop: rot_blk_n1
// (column-major 9), the un-rotated strain increment (6), urcs at n (9) -- strided like history
struct PointInF {
  const double *rur, *ud, *cgn;
  long long stride;
  W3DF_HD double R(int i, int j) const { return rur[(long long)(j * 3 + i) * stride]; }
  W3DF_HD double uddt(int k) const { return ud[(long long)k * stride]; }
  W3DF_HD double urcs_n(int k) const { return cgn[(long long)k * stride]; }
};
constexpr int WORK_DOUBLES = 111;          // + NSF with geometric hardening (tau_l)
// Geometric hardening: tau_l of the 12 systems takes the 7 words O_XP (free: the previous iterate
// was only kept for a separate tangent trip), the words of the 4 cold scalars (which go back to
// registers then) and one more -- 112 words per thread, so that two CTAs still fit an SM (with 12
// words of its own: 123, one CTA).
template <bool TAUL>
W3DF_HD inline int tl_word(int i) {
  return i < 7 ? O_XP + i : (i < 11 ? O_SC + (i - 7) : O_TL);
}
constexpr int work_doubles(bool taul) { return taul ? WORK_DOUBLES + 1 : WORK_DOUBLES; }
// a cold scalar of the solver loop: a word of the column, or a plain variable
template <bool IN_COLUMN>
struct ColdScalar {
  volatile double *p;
  double v;
  W3DF_HD explicit ColdScalar(volatile double *q) : p(q), v(0.0) {}
  W3DF_HD operator double() const { return IN_COLUMN ? *p : v; }
  W3DF_HD ColdScalar &operator=(double x) {
    if (IN_COLUMN) *p = x;
    else v = x;
    return *this;
  }
};

// ---- small dense helpers (mm10_a.f:1431 rt2rve, :1492 rt2rvw, mm10_b.f:1513 symsw) -------------
W3DF_HD inline void rt2rve(const double (&rt)[3][3], double (&rv)[6][6]) {
  const double two = 2.0;
  rv[0][0] = rt[0][0] * rt[0][0]; rv[0][1] = rt[0][1] * rt[0][1]; rv[0][2] = rt[0][2] * rt[0][2];
  rv[0][3] = two * rt[0][0] * rt[0][1]; rv[0][4] = two * rt[0][2] * rt[0][1]; rv[0][5] = two * rt[0][0] * rt[0][2];
  rv[1][0] = rt[1][0] * rt[1][0]; rv[1][1] = rt[1][1] * rt[1][1]; rv[1][2] = rt[1][2] * rt[1][2];
  rv[1][3] = two * rt[1][0] * rt[1][1]; rv[1][4] = two * rt[1][2] * rt[1][1]; rv[1][5] = two * rt[1][0] * rt[1][2];
  rv[2][0] = rt[2][0] * rt[2][0]; rv[2][1] = rt[2][1] * rt[2][1]; rv[2][2] = rt[2][2] * rt[2][2];
  rv[2][3] = two * rt[2][0] * rt[2][1]; rv[2][4] = two * rt[2][2] * rt[2][1]; rv[2][5] = two * rt[2][0] * rt[2][2];
  rv[3][0] = rt[0][0] * rt[1][0]; rv[3][1] = rt[0][1] * rt[1][1]; rv[3][2] = rt[0][2] * rt[1][2];
  rv[3][3] = rt[0][0] * rt[1][1] + rt[1][0] * rt[0][1];
  rv[3][4] = rt[0][1] * rt[1][2] + rt[0][2] * rt[1][1];
  rv[3][5] = rt[0][0] * rt[1][2] + rt[0][2] * rt[1][0];
  rv[4][0] = rt[1][0] * rt[2][0]; rv[4][1] = rt[2][1] * rt[1][1]; rv[4][2] = rt[1][2] * rt[2][2];
  rv[4][3] = rt[1][0] * rt[2][1] + rt[1][1] * rt[2][0];
  rv[4][4] = rt[1][1] * rt[2][2] + rt[2][1] * rt[1][2];
  rv[4][5] = rt[1][0] * rt[2][2] + rt[1][2] * rt[2][0];
  rv[5][0] = rt[0][0] * rt[2][0]; rv[5][1] = rt[0][1] * rt[2][1]; rv[5][2] = rt[0][2] * rt[2][2];
  rv[5][3] = rt[0][0] * rt[2][1] + rt[0][1] * rt[2][0];
  rv[5][4] = rt[0][1] * rt[2][2] + rt[0][2] * rt[2][1];
  rv[5][5] = rt[0][0] * rt[2][2] + rt[2][0] * rt[0][2];
}
W3DF_HD inline void rt2rvw(const double (&rt)[3][3], double (&rv)[3][3]) {
  rv[0][0] = rt[1][1] * rt[2][2] - rt[1][2] * rt[2][1];
  rv[0][1] = rt[1][0] * rt[2][2] - rt[1][2] * rt[2][0];
  rv[0][2] = rt[1][0] * rt[2][1] - rt[1][1] * rt[2][0];
  rv[1][0] = rt[0][1] * rt[2][2] - rt[0][2] * rt[2][1];
  rv[1][1] = rt[0][0] * rt[2][2] - rt[0][2] * rt[2][0];
  rv[1][2] = rt[0][0] * rt[2][1] - rt[0][1] * rt[2][0];
  rv[2][0] = rt[0][1] * rt[1][2] - rt[0][2] * rt[1][1];
  rv[2][1] = rt[0][0] * rt[1][2] - rt[0][2] * rt[1][0];
  rv[2][2] = rt[0][0] * rt[1][1] - rt[0][1] * rt[1][0];
}
W3DF_HD inline void mat33(double (&a)[3][3], const double (&b)[3][3], const double (&c)[3][3]) {
  W3DF_UNROLL
  for (int i = 0; i < 3; ++i)
    W3DF_UNROLL
    for (int j = 0; j < 3; ++j) a[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j];
}
W3DF_HD inline void symsw(const double *s, const double *w, double (&sw)[6]) {
  const double half = 0.5;
  sw[0] = s[3] * w[2] - s[5] * w[1];
  sw[1] = s[3] * w[2] - s[4] * w[0];
  sw[2] = s[5] * w[1] + s[4] * w[0];
  sw[3] = half * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]);
  sw[4] = half * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]);
  sw[5] = half * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4]);
}
W3DF_HD inline double sgn1(double x) { return copysign(1.0, x); }

// x^e, x >= 0, for the CH values of one trip at once (instruction-level parallelism): binary powering for the
// whole-number exponents of the reference's decks, pow() otherwise
// pow() for exponents that are not whole numbers: out of line (its ~290 inlined instructions per
// call site would otherwise sit in the middle of the solver loop's code)
static W3DF_HD W3DF_NOINLINE double pow_general(double x, double e) { return pow(x, e); }
// lock-step policy of the solver loop.  On the device all threads of a CTA walk the loop together
// (CtaSync in w3d_cp.cu: __syncthreads_or): the 8 warps then share one instruction stream instead
// of thrashing the instruction cache with 8 (measured: 12 stall cycles per issue without it), and
// since a CTA owns its SM until its slowest warp is done, waiting costs nothing.  Host: no-op.
struct NoSync {
  static W3DF_HD bool any(bool p) { return p; }
};
constexpr int CH = W3DF_CHUNK;
static_assert(NSF % CH == 0, "chunk must divide the slip-system capacity");
// square-and-multiply with a compile-time exponent (straight-line code)
template <int E>
W3DF_HD inline double pow_const(double x) {
  if (E == 0) return 1.0;
  if (E == 1) return x;
  const double h = pow_const<E / 2>(x);
  return (E & 1) ? h * h * x : h * h;
}
W3DF_HD inline void pow4(int ie, double e, double (&b)[CH], double (&r)[CH]) {
  if (ie < 0) {
    W3DF_UNROLL
    for (int q = 0; q < CH; ++q) r[q] = pow_general(b[q], e);
    return;
  }
  // the rate exponents of the reference's decks (harden_n - 1 = 19 in verification/test82 and
  // test74, 4 .. 49 elsewhere) as straight-line chains: the rolled loop below spends more
  // instructions on its control than on multiplying (measured: 20 % of all instructions executed)
#define W3DF_POW_CASE(E)                                   \
  case E:                                                  \
    W3DF_UNROLL                                            \
    for (int q = 0; q < CH; ++q) r[q] = pow_const<E>(b[q]); \
    return;
  switch (ie) {
    W3DF_POW_CASE(0) W3DF_POW_CASE(1) W3DF_POW_CASE(2) W3DF_POW_CASE(3) W3DF_POW_CASE(4)
    W3DF_POW_CASE(5) W3DF_POW_CASE(7) W3DF_POW_CASE(9) W3DF_POW_CASE(11) W3DF_POW_CASE(14)
    W3DF_POW_CASE(19) W3DF_POW_CASE(24) W3DF_POW_CASE(29) W3DF_POW_CASE(49)
    default: break;
  }
#undef W3DF_POW_CASE
  W3DF_UNROLL
  for (int q = 0; q < CH; ++q) r[q] = 1.0;
  W3DF_ROLLED
  for (int k = ie; k != 0; k >>= 1) {
    if (k & 1) {
      W3DF_UNROLL
      for (int q = 0; q < CH; ++q) r[q] = r[q] * b[q];
    }
    W3DF_UNROLL
    for (int q = 0; q < CH; ++q) b[q] = b[q] * b[q];
  }
}
// RATE_E >= 0: the rate exponent is known when the kernel is compiled (no dispatch per call)
template <int RATE_E>
W3DF_HD inline void pow_rate(double (&b)[CH], double (&r)[CH]) {
  if (RATE_E >= 0) {
    W3DF_UNROLL
    for (int q = 0; q < CH; ++q) r[q] = pow_const<(RATE_E >= 0 ? RATE_E : 0)>(b[q]);
  } else {
    pow4(W3DF_CT.rate_ie, W3DF_CT.rate_n - 1.0, b, r);
  }
}
W3DF_HD inline double pow1(int ie, double e, double x) {
  if (ie < 0) return pow_general(x, e);
  if (ie < 16) {
    // the loop below for exponents under 16 without its branches: the same products in the same
    // order, the multiplications by 1 left out by selects
    const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
    double r = (ie & 1) ? x : 1.0;
    r = (ie & 2) ? r * x2 : r;
    r = (ie & 4) ? r * x4 : r;
    r = (ie & 8) ? r * x8 : r;
    return r;
  }
  double r = 1.0, b = x;
  for (int k = ie; k != 0; k >>= 1) {
    if (k & 1) r = r * b;
    b = b * b;
  }
  return r;
}

W3DF_HD inline double effective_dg(const double *D) {     // mm10_setup :880
  const double t1 = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
  const double t2 = D[3] * D[3] + D[4] * D[4] + D[5] * D[5];
  return sqrt((2.0 / 3.0) * (t1 + 0.5 * t2));
}
W3DF_HD inline void unit_dev(const double *D, double (&d)[6]) {   // mm10_solve_predict :3070-3110
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) d[k] = D[k];
  const double dtrace = (d[0] + d[1] + d[2]) / 3.0;
  d[0] -= dtrace; d[1] -= dtrace; d[2] -= dtrace;
  const double t1 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const double t2 = d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
  if (t1 + t2 == 0.0) {
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) d[k] = 0.0;
  } else {
    const double sn = sqrt(t1 + t2);
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) d[k] = d[k] / sn;
  }
}

// In-place LU of the 7x7 matrix in DGETF2 order for the usual case that no row interchange is
// needed (every diagonal entry is already the first largest of its column: the Newton matrices
// here are I + small terms).  All indices are compile-time: A stays in registers.  Returns 7 when
// done, else the first column j that needs an interchange -- columns < j are eliminated exactly
// as the pivoting routine would have done, so lu7_slow can take over from there.
// The reciprocal of each pivot replaces the pivot itself (the diagonal of U is only ever used as a
// divisor): no second array of seven doubles to keep alive next to the 49 of the matrix.
W3DF_HD inline int lu7_nopivot(double (&A)[7][7]) {
  int jstop = 7;
  W3DF_UNROLL
  for (int j = 0; j < 7; ++j) {
    if (jstop == 7) {
      const double amax = fabs(A[j][j]);
      bool ok = (A[j][j] != 0.0);
      W3DF_UNROLL
      for (int i = j + 1; i < 7; ++i) ok = ok && !(fabs(A[i][j]) > amax);
      if (!ok) {
        jstop = j;
      } else {
        const double r = 1.0 / A[j][j];
        A[j][j] = r;
        W3DF_UNROLL
        for (int i = j + 1; i < 7; ++i) {
          A[i][j] *= r;
          W3DF_UNROLL
          for (int k = j + 1; k < 7; ++k) A[i][k] -= A[i][j] * A[j][k];
        }
      }
    }
  }
  return jstop;
}
// forward / back substitution with the factors of lu7_nopivot
W3DF_HD inline void lu7_apply(const double (&A)[7][7], double (&b)[7]) {
  W3DF_UNROLL
  for (int i = 0; i < 7; ++i)
    W3DF_UNROLL
    for (int j = 0; j < i; ++j) b[i] -= A[i][j] * b[j];
  W3DF_UNROLL
  for (int i = 6; i >= 0; --i) {
    W3DF_UNROLL
    for (int j = i + 1; j < 7; ++j) b[i] -= A[i][j] * b[j];
    b[i] *= A[i][i];          // the pivot's reciprocal (lu7_nopivot)
  }
}
// The rare case: finish the factorisation from column j0 with partial pivoting (DGETF2: pivot =
// first largest |a| of the column, rows interchanged across the whole matrix) and solve for nrhs
// right-hand sides B[r * nrhs + k].  Plain loops on arrays in memory; kept out of line so that the
// usual path keeps its matrix in registers.
static W3DF_HD W3DF_NOINLINE void lu7_slow(double *A /*[49] row-major*/, int j0, double *B, int nrhs) {
  int piv[7];
  double rinv[7];
  W3DF_ROLLED
  for (int j = 0; j < 7; ++j) {
    piv[j] = j;
    if (j < j0) { rinv[j] = A[j * 7 + j]; continue; }   // columns done by lu7_nopivot hold 1 / pivot
    int p = j;
    double amax = fabs(A[j * 7 + j]);
    for (int i = j + 1; i < 7; ++i)
      if (fabs(A[i * 7 + j]) > amax) { amax = fabs(A[i * 7 + j]); p = i; }
    piv[j] = p;
    rinv[j] = 0.0;
    if (A[p * 7 + j] != 0.0) {
      if (p != j)
        for (int k = 0; k < 7; ++k) { const double t = A[j * 7 + k]; A[j * 7 + k] = A[p * 7 + k]; A[p * 7 + k] = t; }
      const double r = 1.0 / A[j * 7 + j];
      rinv[j] = r;
      for (int i = j + 1; i < 7; ++i) A[i * 7 + j] *= r;
    }
    for (int i = j + 1; i < 7; ++i)
      for (int k = j + 1; k < 7; ++k) A[i * 7 + k] -= A[i * 7 + j] * A[j * 7 + k];
  }
  W3DF_ROLLED
  for (int k = 0; k < nrhs; ++k) {
    W3DF_ROLLED
    for (int j = 0; j < 7; ++j)
      if (piv[j] != j) { const double t = B[j * nrhs + k]; B[j * nrhs + k] = B[piv[j] * nrhs + k]; B[piv[j] * nrhs + k] = t; }
    for (int i = 0; i < 7; ++i)
      for (int j = 0; j < i; ++j) B[i * nrhs + k] -= A[i * 7 + j] * B[j * nrhs + k];
    for (int i = 6; i >= 0; --i) {
      for (int j = i + 1; j < 7; ++j) B[i * nrhs + k] -= A[i * 7 + j] * B[j * nrhs + k];
      B[i * nrhs + k] *= rinv[i];
    }
  }
}

// sums over the crystals of one Gauss point (mm10_a_do_crystal accumulates, mm10_a_crystal_avgs
// divides); tangent and slip history are accumulated in history at n+1 directly
struct PointSumsF {
  double sig[6], t_work, p_work, p_strain;
};

// The state of the Voce hardening law shared by the residual and the Jacobian at one tt
struct HardF {
  double a, inv_d;      // h_term_i = a + tau_l_i * inv_d   (mm10_h_voche, mm10_b.f:1896)
};
W3DF_HD inline HardF hard_term(double tt) {
  HardF h;
  const double d = tt - W3DF_CT.tau_y;
  h.inv_d = 1.0 / d;
  h.a = 1.0 - d / W3DF_CT.tau_v;
  return h;
}

// ------------------------------------------------------------------------------------------------
// One crystal: mm10_a_do_crystal -> mm10_solve_crystal and everything below it, then
// mm10_store_cryhist.  Returns true when the crystal asks for a global step cut.
// n_* : state at n of this crystal; R: rot_blk_n1 (3x3); W: per-thread workspace column.
// ------------------------------------------------------------------------------------------------
// The tangent is stored ahead: every full Newton pass also solves for the tangent columns with
// the factors it has in registers and stores them (unsymmetrised) in history at n+1; the pass that
// turns out to be the last one has then left exactly what mm10_tangent forms from the Jacobian of
// the last solver pass -- no trip that re-forms and re-factors that Jacobian.
// SINGLE (one crystal per point): the columns go straight to the tangent's place and are
// symmetrised there; otherwise the tangents of the crystals are summed, and the columns of crystal
// c wait in words of its own history at n+1 that are only written after the solve (wait_word).
// word k = 0..35 of the waiting tangent columns of crystal c: Euler angles + Rp (12 consecutive
// words), eps (6), the user words (15), ep (3 of 6) -- none of them written before mm10_output
W3DF_HD inline int wait_word(const LayoutF &L, int c, int k) {
  return k < 12 ? L.c_euler(c) + k : (k < 18 ? L.c_eps(c) + (k - 12) : (k < 33 ? L.c_u(c) + (k - 18) : L.c_ep(c) + (k - 33)));
}
template <bool TAUL, class SYNC, bool SINGLE, int RATE_E>
W3DF_HD inline bool crystal_update_fast(const FProps &P, const WorkF &W, const PointInF &in,
                                        double n_tt, double n_tt_rate, double dt,
                                        bool iter_0_extrapolate_off, bool asymmetric, bool init,
                                        const LayoutF &L, int c, const CHistF &hn, const HistF &h1,
                                        PointSumsF &acc, unsigned &trips) {
  // state at n of this crystal, read where needed (mm10_copy_cc_hist / mm10_init_cc_hist0)
  auto n_stress = [&](int k) { return init ? in.urcs_n(k) : hn[L.c_stress(c) + k]; };
  auto n_D = [&](int k) { return init ? 0.0 : hn[L.c_D(c) + k]; };
  auto load_tr2 = [&](double (&t)[3][3]) {     // Rp_n^T
    W3DF_UNROLL
    for (int j = 0; j < 3; ++j)
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i) t[j][i] = init ? ((i == j) ? 1.0 : 0.0) : hn[L.c_Rp(c) + j * 3 + i];
  };
  auto load_R = [&](double (&t)[3][3]) {
    W3DF_UNROLL
    for (int j = 0; j < 3; ++j)
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i) t[i][j] = in.R(i, j);
  };
  const int ns = W3DF_CT.nslip;
  const double cvr = W3DF_CT.iD_v;
  // ---- mm10_setup: B6 = RE E6, CB = C B6, B3 = RQ W3 into the workspace
  {
    double tr2[3][3], RE[6][6];
    load_tr2(tr2);
    rt2rve(tr2, RE);
    W3DF_UNROLL
    for (int r = 0; r < 6; ++r)
      W3DF_UNROLL
      for (int k = 0; k < 6; ++k) {
        double v = RE[r][0] * P.E6[0][k];
        W3DF_UNROLL
        for (int q = 1; q < 6; ++q) v = v + RE[r][q] * P.E6[q][k];
        W(O_B6 + r * 6 + k) = v;
      }
  }
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) {           // CB column by column (keeps 6 values of B6 live)
    double bk[6];
    W3DF_UNROLL
    for (int q = 0; q < 6; ++q) bk[q] = W(O_B6 + q * 6 + k);
    W3DF_UNROLL
    for (int r = 0; r < 6; ++r) {
      double v = P.C[r][0] * bk[0];
      W3DF_UNROLL
      for (int q = 1; q < 6; ++q) v = v + P.C[r][q] * bk[q];
      W(O_CB + r * 6 + k) = v;
    }
  }
  {
    double tr2[3][3], R[3][3], wk[3][3], RQ[3][3];
    load_tr2(tr2);
    load_R(R);
    mat33(wk, R, tr2);
    rt2rvw(wk, RQ);
    W3DF_UNROLL
    for (int a = 0; a < 3; ++a)
      W3DF_UNROLL
      for (int b = 0; b < 3; ++b)
        W(O_B3 + a * 3 + b) = RQ[a][0] * P.W3[0][b] + RQ[a][1] * P.W3[1][b] + RQ[a][2] * P.W3[2][b];
  }
  if (TAUL) {
    // geometric hardening tau_l from the lattice curvature at n (Acharya), mm10_a.f:925-975
    const double mu_harden = P.C[5][5];
    const double alpha = 1.0 / 3.0;
    const double cst = W3DF_CT.k_0 * W3DF_CT.burgers * alpha * alpha * mu_harden * mu_harden / 2.0 / W3DF_CT.theta_0;
    double X3[3][3];   // n_R tr2 g^T : crystal-axes normal -> current configuration at n
    {
      double tr2[3][3], n_R[3][3], t1[3][3], gt[3][3];
      load_tr2(tr2);
      W3DF_UNROLL
      for (int j = 0; j < 3; ++j)
        W3DF_UNROLL
        for (int i = 0; i < 3; ++i) {
          n_R[i][j] = init ? ((i == j) ? 1.0 : 0.0) : hn[L.o_R() + j * 3 + i];
          gt[i][j] = P.g[j][i];
        }
      mat33(t1, tr2, gt);
      mat33(X3, n_R, t1);
    }
    double curv[3][3][3];
#define W3DF_GF(i, j, k) (init ? 0.0 : hn[L.o_gradfe() + (i) + 3 * (j) + 9 * (k)])
    W3DF_UNROLL
    for (int i = 0; i < 3; ++i)
      W3DF_UNROLL
      for (int j = 0; j < 3; ++j)
        W3DF_UNROLL
        for (int k = 0; k < 3; ++k) curv[i][j][k] = W3DF_GF(i, j, k) - W3DF_GF(i, k, j);
#undef W3DF_GF
    W3DF_UNROLL
    for (int t = 0; t < NSF; ++t) {
      double cn[3], tm[3];
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i)
        cn[i] = X3[i][0] * W3DF_CT.ncr[t][0] + X3[i][1] * W3DF_CT.ncr[t][1] + X3[i][2] * W3DF_CT.ncr[t][2];
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i)
        tm[i] = 0.0 + 0.5 * (-curv[i][0][1] * cn[2] + curv[i][0][2] * cn[1] + curv[i][1][0] * cn[2] -
                             curv[i][1][2] * cn[0] - curv[i][2][0] * cn[1] + curv[i][2][1] * cn[0]);
      W(tl_word<TAUL>(t)) = cst * sqrt(tm[0] * tm[0] + tm[1] * tm[1] + tm[2] * tm[2]);
    }
  }

  // ---- mm10_solve_strup
  double x[7];      // accepted state: stress (6), tau_tilde
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) x[k] = n_stress(k);
  x[6] = n_tt;
  bool no_load;
  {
    double temp1 = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) {
      const double u = in.uddt(k);
      temp1 = temp1 + u * u;
      h1[L.c_D(c) + k] = u;       // per-crystal history entries every path stores the same way
    }
    const double temp2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3] + x[4] * x[4] + x[5] * x[5];
    no_load = (temp2 == 0.0) && (temp1 == 0.0);
  }
  h1[L.c_tt(c)] = 0.0;
  h1[L.c_rate(c)] = 0.0;

  const bool elastic = iter_0_extrapolate_off || no_load;
  bool done = false, cut = false;
  const bool idle = elastic && no_load;
  if (idle) {
    // mm10_solve_strup_a without load: stress and tangent as they are
    W3DF_UNROLL
    for (int cc = 0; cc < 6; ++cc)
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        const double t = P.C[r][cc];
        h1[L.o_D() + cc * 6 + r] = (c == 0) ? t : h1[L.o_D() + cc * 6 + r] + t;
      }
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) { h1[L.c_stress(c) + k] = x[k]; acc.sig[k] = acc.sig[k] + x[k]; }
    h1[L.c_tt(c)] = x[6];
    W3DF_UNROLL
    for (int k = 0; k < 3; ++k) h1[L.c_euler(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 9; ++k) h1[L.c_Rp(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) { h1[L.c_eps(c) + k] = 0.0; h1[L.c_ep(c) + k] = 0.0; h1[L.c_ed(c) + k] = 0.0; }
    for (int k = 0; k < ns; ++k) h1[L.c_slip(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 15; ++k) h1[L.c_u(c) + k] = 0.0;
    done = true;       // nothing to solve; the thread still walks the loop's barriers
  }

  // ---- the flat solver loop.  Stages of one evaluation of the residual at xt:
  enum { ST_ELASTIC = 0, ST_INIT_PRED = 1, ST_LS = 3 };
  const double c_ls = 1.0e-4, red = 0.5;
  const int mls = 10, mmin = 1, mcuts = 4;
  const int miter = W3DF_CT.miter;
  // carried across the trips: x and the step dx; the trial point x + alpha dx is re-formed at the
  // top of a trip (seven doubles fewer to keep live through the Jacobian)
  double dx[7], Rv[7];
  double tinc = dt, dg, tt_rate = 0.0;
  constexpr bool SC_IN_COLUMN = !TAUL;
  ColdScalar<SC_IN_COLUMN> frac(&W(O_SC + 0)), stepf(&W(O_SC + 1)), ott(&W(O_SC + 2)), inR1(&W(O_SC + 3));
  frac = 0.0; stepf = 1.0; ott = x[6]; inR1 = 0.0;
  // loop control packed into one register (each of these was a register or a spill slot of its own)
  // (iter <= miter + 1 <= 251: mm10_fast_eligible; ls <= mls + 1; trips saturates)
  struct { unsigned cuts : 4, stage : 4, iter : 8, ls : 4, trips : 8, full : 1, failed : 1, begin_sub : 1; } ct;
  ct.cuts = 0; ct.stage = 0; ct.iter = 0; ct.ls = 0; ct.trips = 0;
  ct.full = 0; ct.failed = 0;
  double inR = 0.0, ls1 = 0.0, ls2 = 0.0, alpha = 1.0, nR = 0.0;
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) W(O_OST + k) = x[k];
  W3DF_UNROLL
  for (int k = 0; k < 7; ++k) { dx[k] = 0.0; Rv[k] = 0.0; }

#ifdef W3D_MM10F_HIST
  int w3df_trips = 0;
#endif
  double r2c = 0.0, ttrc = 0.0;   // hardening residual and rate of the state last evaluated
  ct.begin_sub = 1;
  for (;;) {
    if (!SYNC::any(!done)) break;
    if (done) continue;
    if (ct.trips < 255u) ct.trips = ct.trips + 1;
#ifdef W3D_MM10F_HIST
    ++w3df_trips;
#endif
    if (ct.begin_sub) {
      // (sub-)increment set-up: curr%D, curr%tinc, curr%dg; k0 = n%stress + C D; predictor start
      ct.begin_sub = false;
      const double fr = elastic ? 1.0 : (stepf + frac);
      double D[6];
      W3DF_UNROLL
      for (int k = 0; k < 6; ++k) D[k] = in.uddt(k) * fr;
      tinc = dt * fr;
      dg = effective_dg(D);
      if (W3DF_CT.alter_mode != 0) dg = W3DF_CT.eps_dot_0_y * tinc;
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        double v = n_stress(r);
        W3DF_UNROLL
        for (int q = 0; q < 6; ++q) v = v + P.C[r][q] * D[q];
        W(O_K0 + r) = v;
      }
      x[6] = n_tt;   // "some subroutines modify n%tau_tilde in setup": tt restarts from n
      // trial point of this trip: x itself, hardening entry moved by the predictor below
      alpha = 1.0;
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) dx[k] = 0.0;
      if (elastic) {
        ct.stage = ST_ELASTIC;
      } else {
        double d1[6], d2[6], nd[6];
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) nd[k] = n_D(k);
        unit_dev(D, d1);
        unit_dev(nd, d2);
        const double t1 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
        const double t2 = d1[3] * d2[3] + d1[4] * d2[4] + d1[5] * d2[5];
        const double cos_ang = fmax(t1 + t2, 0.0);
        dx[6] = cos_ang * n_tt_rate * (dt * stepf);
        ct.stage = ST_INIT_PRED;
      }
      ct.full = false;
      ct.iter = 0;
    }

    bool newton = false;
    {
      // ================= residual at xt (mm10_formR1, mm10_formR2) =================
      double dot;
      double xt[7];
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) xt[k] = x[k] + alpha * dx[k];
      {
        double sp[6];
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) {
          double v = W(O_B6 + k) * xt[0];
          W3DF_UNROLL
          for (int r = 1; r < 6; ++r) v = v + W(O_B6 + r * 6 + k) * xt[r];
          sp[k] = v;
        }
        const double tt = xt[6];
        const double itt = 1.0 / tt, dgt = dg / tt, cv = tinc * cvr;
        const HardF ht = hard_term(tt);
        // (the hardening sum is formed in the predictor passes too: when the predictor converges,
        //  the update's first residual is this same evaluation plus its 7th entry -- no second pass)
        double hp0 = 0.0, sg0 = 1.0;
        if (!TAUL) {
          hp0 = pow1(W3DF_CT.voche_ie, W3DF_CT.voche_m, fabs(ht.a));
          sg0 = sgn1(ht.a);
        }
        double a6[6] = {0, 0, 0, 0, 0, 0}, a3[3] = {0, 0, 0}, hsum = 0.0;
        W3DF_ROLLED
        for (int i0 = 0; i0 < NSF; i0 += CH) {
          double rs[CH], b[CH], pw[CH];
          W3DF_UNROLL
          for (int q = 0; q < CH; ++q) {
            const int i = i0 + q;
            rs[q] = sp[0] * W3DF_CT.mc[i][0] + sp[1] * W3DF_CT.mc[i][1] + sp[2] * W3DF_CT.mc[i][2] +
                    sp[3] * W3DF_CT.mc[i][3] + sp[4] * W3DF_CT.mc[i][4] + sp[5] * W3DF_CT.mc[i][5];
            b[q] = fabs(rs[q] * itt);
          }
          pow_rate<RATE_E>(b, pw);
          W3DF_UNROLL
          for (int q = 0; q < CH; ++q) {
            const int i = i0 + q;
            const double si = dgt * pw[q] * rs[q];
            const double f = rs[q] * cv + si;
            W3DF_UNROLL
            for (int k = 0; k < 6; ++k) a6[k] = a6[k] + f * W3DF_CT.mc[i][k];
            W3DF_UNROLL
            for (int k = 0; k < 3; ++k) a3[k] = a3[k] + f * W3DF_CT.qcr[i][k];
            if (TAUL) {
              const double h_term = ht.a + W(tl_word<TAUL>(i)) * ht.inv_d;
              hsum = hsum + pow1(W3DF_CT.voche_ie, W3DF_CT.voche_m, fabs(h_term)) * sgn1(h_term) * fabs(si);
            } else {
              hsum = hsum + fabs(si);
            }
          }
        }
        if (!TAUL) hsum = hp0 * sg0 * hsum;
        double wp[3], symTW[6];
        W3DF_UNROLL
        for (int a = 0; a < 3; ++a)
          wp[a] = W(O_B3 + a * 3) * a3[0] + W(O_B3 + a * 3 + 1) * a3[1] + W(O_B3 + a * 3 + 2) * a3[2];
        symsw(xt, wp, symTW);
        dot = 0.0;
        W3DF_UNROLL
        for (int r = 0; r < 6; ++r) {
          double v = W(O_CB + r * 6) * a6[0];
          W3DF_UNROLL
          for (int q = 1; q < 6; ++q) v = v + W(O_CB + r * 6 + q) * a6[q];
          Rv[r] = xt[r] - W(O_K0 + r) + v + 2.0 * symTW[r];
          dot = dot + Rv[r] * Rv[r];
        }
        {
          const double h = n_tt + W3DF_CT.theta_0 * hsum;
          ttrc = (h - n_tt) / tinc;
          r2c = tt - h;
        }
        if (ct.full) {
          tt_rate = ttrc;
          Rv[6] = r2c;
          dot = dot + Rv[6] * Rv[6];
        } else {
          Rv[6] = 0.0;
        }
      }
      if (ct.stage == ST_ELASTIC) {
        // mm10_solve_strup_a: elastic stiffness as tangent; stress corrected for creep only
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) x[k] = x[k] - Rv[k];
        done = true;
        continue;
      }
      nR = sqrt(dot);
      bool check = false;
      if (ct.stage == ST_INIT_PRED) {
        inR1 = nR;
        W3DF_UNROLL
        for (int k = 0; k < 7; ++k) x[k] = xt[k];     // x(7) = predicted tau_tilde x2
        check = true;
      } else {   // ST_LS
        const double nlsx = ls1 + ls2 * alpha;
        const double nRs = 0.5 * dot;
        if ((nRs <= nlsx) || (ct.ls > mls)) {
          W3DF_UNROLL
          for (int k = 0; k < 7; ++k) x[k] = xt[k];
          ct.iter = ct.iter + 1;
          bool bad = ct.iter > miter;
          W3DF_UNROLL
          for (int k = 0; k < 7; ++k) bad = bad || (x[k] != x[k]);
          if (bad) ct.failed = true;
          else check = true;
        } else {
          alpha = red * alpha;
          ct.ls = ct.ls + 1;
        }
      }
      if (check) {
        const bool cont = ct.full ? (((nR > W3DF_CT.atol) && (nR / inR > W3DF_CT.rtol)) || (ct.iter < mmin))
                               : ((nR > W3DF_CT.atol1) && (nR / inR1 > W3DF_CT.rtol1));
        if (cont) {
          newton = true;
        } else if (!ct.full) {
          // predictor converged: material update from the predicted state (mm10_solve_update).
          // Its first residual is the one just formed plus the hardening entry; the loop test
          // holds for iter = 0 < mmin, so the first update pass follows in this same trip
          ct.full = true;
          ct.iter = 0;
          tt_rate = ttrc;
          Rv[6] = r2c;
          dot = dot + Rv[6] * Rv[6];
          nR = sqrt(dot);
          inR = nR;
          if (inR == 0.0) inR = inR1;
          newton = true;
        } else {
          // this (sub-)increment is solved
          W3DF_UNROLL
          for (int k = 0; k < 6; ++k) W(O_OST + k) = x[k];
          ott = x[6];
          frac = frac + stepf;
          if (frac < 1.0) {
            ct.begin_sub = true;
          } else {
            done = true;      // the last full Newton pass has stored the tangent columns
          }
        }
      }
      if (ct.failed) {
        // mm10_solve failed (predictor or update): back to the last good state, halve the step
        ct.failed = false;
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) x[k] = W(O_OST + k);
        x[6] = ott;
        stepf = stepf * 0.5;
        ct.cuts = ct.cuts + 1;
        if (ct.cuts > mcuts) { cut = true; done = true; continue; }
        ct.begin_sub = true;
      }
    }
    if (!newton) continue;

    // ================= Jacobian at x (mm10_formJ) and the Newton step =================
    // (the residual is parked in the workspace while the Jacobian is accumulated: registers)
    W3DF_UNROLL
    for (int k = 0; k < 7; ++k) W(O_RV + k) = Rv[k];
    double (&xs)[7] = x;
    double A[7][7];
    {
      double sp[6];
      W3DF_UNROLL
      for (int k = 0; k < 6; ++k) {
        double v = W(O_B6 + k) * xs[0];
        W3DF_UNROLL
        for (int r = 1; r < 6; ++r) v = v + W(O_B6 + r * 6 + k) * xs[r];
        sp[k] = v;
      }
      const double tt = xs[6];
      const double itt = 1.0 / tt, dgt = dg / tt, cv = tinc * cvr;
      const double dgn = dg * W3DF_CT.rate_n / tt, nott = W3DF_CT.rate_n / tt;
      const HardF ht = hard_term(tt);
      const double mitv = -1.0 / W3DF_CT.tau_v;
      double hp0 = 0.0, sg0 = 1.0, rx0 = 0.0;
      if (!TAUL && ct.full) {
        hp0 = pow1(W3DF_CT.voche_ie, W3DF_CT.voche_m, fabs(ht.a));
        sg0 = sgn1(ht.a);
        rx0 = 1.0 / fabs(ht.a);
      }
      double M[6][6], Q[3][6], a6[6], q3[3], et[6], wpl[3], etau = 0.0;
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        W3DF_UNROLL
        for (int cc = r; cc < 6; ++cc) M[r][cc] = 0.0;
        a6[r] = 0.0; et[r] = 0.0;
      }
      W3DF_UNROLL
      for (int a = 0; a < 3; ++a) {
        W3DF_UNROLL
        for (int cc = 0; cc < 6; ++cc) Q[a][cc] = 0.0;
        q3[a] = 0.0; wpl[a] = 0.0;
      }
      W3DF_ROLLED
      for (int i0 = 0; i0 < NSF; i0 += CH) {
        double rs[CH], b[CH], pw[CH];
        W3DF_UNROLL
        for (int q = 0; q < CH; ++q) {
          const int i = i0 + q;
          rs[q] = sp[0] * W3DF_CT.mc[i][0] + sp[1] * W3DF_CT.mc[i][1] + sp[2] * W3DF_CT.mc[i][2] +
                  sp[3] * W3DF_CT.mc[i][3] + sp[4] * W3DF_CT.mc[i][4] + sp[5] * W3DF_CT.mc[i][5];
          b[q] = fabs(rs[q] * itt);
        }
        pow_rate<RATE_E>(b, pw);
        W3DF_UNROLL
        for (int q = 0; q < CH; ++q) {
          const int i = i0 + q;
          const double si = dgt * pw[q] * rs[q];
          const double g = dgn * pw[q] + cv;                    // mm10_dgdt_voche + viscous term
          double gm[6];
          W3DF_UNROLL
          for (int k = 0; k < 6; ++k) gm[k] = g * W3DF_CT.mc[i][k];
          W3DF_UNROLL
          for (int r = 0; r < 6; ++r)
            W3DF_UNROLL
            for (int cc = r; cc < 6; ++cc) M[r][cc] = M[r][cc] + gm[r] * W3DF_CT.mc[i][cc];
          W3DF_UNROLL
          for (int a = 0; a < 3; ++a)
            W3DF_UNROLL
            for (int cc = 0; cc < 6; ++cc) Q[a][cc] = Q[a][cc] + W3DF_CT.qcr[i][a] * gm[cc];
          const double fw = si + rs[q] * cv;                    // mm10_form_wp for mm10_IW
          W3DF_UNROLL
          for (int a = 0; a < 3; ++a) wpl[a] = wpl[a] + fw * W3DF_CT.qcr[i][a];
          if (ct.full) {
            const double dgdh = -nott * si;                     // mm10_dgdh_voche
            W3DF_UNROLL
            for (int k = 0; k < 6; ++k) a6[k] = a6[k] + dgdh * W3DF_CT.mc[i][k];
            W3DF_UNROLL
            for (int a = 0; a < 3; ++a) q3[a] = q3[a] + dgdh * W3DF_CT.qcr[i][a];
            double hp = hp0, sg = sg0, rx = rx0, tl = 0.0;
            if (TAUL) {
              tl = W(tl_word<TAUL>(i));
              const double h_term = ht.a + tl * ht.inv_d;
              hp = pow1(W3DF_CT.voche_ie, W3DF_CT.voche_m, fabs(h_term));
              sg = sgn1(h_term);
              rx = 1.0 / fabs(h_term);
            }
            const double f = hp * sg * pw[q] * sgn1(rs[q]);       // J21 = -estress (mm10_estress_voche)
            W3DF_UNROLL
            for (int k = 0; k < 6; ++k) et[k] = et[k] + f * W3DF_CT.mc[i][k];
            etau = etau + (W3DF_CT.voche_m * (mitv - tl * ht.inv_d * ht.inv_d) * fabs(si) * rx -
                           fabs(si) * nott * sg * sgn1(si)) * hp;  // J22 (mm10_ehard_voche)
          }
        }
      }
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r)
        W3DF_UNROLL
        for (int cc = 0; cc < r; ++cc) M[r][cc] = M[cc][r];
      // P = CB M' + 2 SW(T) (B3 Q')      (lattice columns)
      double Pm[6][6];
      W3DF_UNROLL
      for (int cc = 0; cc < 6; ++cc) {
        double scol[3], sq[6];
        W3DF_UNROLL
        for (int a = 0; a < 3; ++a)
          scol[a] = W(O_B3 + a * 3) * Q[0][cc] + W(O_B3 + a * 3 + 1) * Q[1][cc] + W(O_B3 + a * 3 + 2) * Q[2][cc];
        symsw(xs, scol, sq);
        W3DF_UNROLL
        for (int r = 0; r < 6; ++r) Pm[r][cc] = 2.0 * sq[r];
      }
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        double cb[6];
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) cb[k] = W(O_CB + r * 6 + k);
        W3DF_UNROLL
        for (int cc = 0; cc < 6; ++cc) {
          double v = Pm[r][cc];
          W3DF_UNROLL
          for (int k = 0; k < 6; ++k) v = v + cb[k] * M[k][cc];
          Pm[r][cc] = v;
        }
        if (ct.full) {                        // J12 = CB a6' + 2 symsw(T, B3 q3'), first part
          double v = cb[0] * a6[0];
          W3DF_UNROLL
          for (int k = 1; k < 6; ++k) v = v + cb[k] * a6[k];
          A[r][6] = v;
        }
      }
      // J11 = P B6^T + IW + I ; A = -J
      W3DF_UNROLL
      for (int cc = 0; cc < 6; ++cc) {
        double bc[6];
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) bc[k] = W(O_B6 + cc * 6 + k);
        W3DF_UNROLL
        for (int r = 0; r < 6; ++r) {
          double v = Pm[r][0] * bc[0];
          W3DF_UNROLL
          for (int k = 1; k < 6; ++k) v = v + Pm[r][k] * bc[k];
          A[r][cc] = v;
        }
        if (ct.full) {                        // J21 = -sc B6 et'
          double v = bc[0] * et[0];
          W3DF_UNROLL
          for (int k = 1; k < 6; ++k) v = v + bc[k] * et[k];
          A[6][cc] = v;
        }
      }
      {
        double wp[3];
        W3DF_UNROLL
        for (int a = 0; a < 3; ++a)
          wp[a] = W(O_B3 + a * 3) * wpl[0] + W(O_B3 + a * 3 + 1) * wpl[1] + W(O_B3 + a * 3 + 2) * wpl[2];
        // mm10_IW (mm10_b.f:1613)
        A[0][3] += 2.0 * wp[2]; A[0][5] += -2.0 * wp[1];
        A[1][3] += 2.0 * wp[2]; A[1][4] += -2.0 * wp[0];
        A[2][4] += 2.0 * wp[0]; A[2][5] += 2.0 * wp[1];
        A[3][0] += wp[2]; A[3][1] += -wp[2]; A[3][4] += -wp[1]; A[3][5] += wp[0];
        A[4][1] += wp[0]; A[4][2] += -wp[0]; A[4][3] += wp[1]; A[4][5] += wp[2];
        A[5][0] += wp[1]; A[5][2] += -wp[1]; A[5][3] += wp[0]; A[5][4] += -wp[2];
      }
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) A[r][r] = A[r][r] + 1.0;
      if (ct.full) {
        double q3r[3], sq[6];
        W3DF_UNROLL
        for (int a = 0; a < 3; ++a)
          q3r[a] = W(O_B3 + a * 3) * q3[0] + W(O_B3 + a * 3 + 1) * q3[1] + W(O_B3 + a * 3 + 2) * q3[2];
        symsw(xs, q3r, sq);
        const double sc = W3DF_CT.theta_0 * dg * nott;
        W3DF_UNROLL
        for (int r = 0; r < 6; ++r) {
          A[r][6] = 2.0 * sq[r] + A[r][6];
          A[6][r] = -(sc * A[6][r]);
        }
        A[6][6] = 1.0 - W3DF_CT.theta_0 * etau;
      } else {
        W3DF_UNROLL
        for (int r = 0; r < 6; ++r) { A[r][6] = 0.0; A[6][r] = 0.0; }
        A[6][6] = 1.0;
      }
    }
    // here A holds J.  Line-search terms need trans(J) R before the factorisation destroys J.
    double jtr[7];
    {
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) Rv[k] = W(O_RV + k);
      double dotR = 0.0;
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) dotR = dotR + Rv[k] * Rv[k];
      ls1 = 0.5 * dotR;
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) {   // trans(J) R, folded into ls2 after the solve
        double sj = 0.0;
        W3DF_UNROLL
        for (int r = 0; r < 7; ++r) sj = sj + A[r][k] * Rv[r];
        jtr[k] = sj;
      }
    }
    W3DF_UNROLL
    for (int r = 0; r < 7; ++r)
      W3DF_UNROLL
      for (int cc = 0; cc < 7; ++cc) A[r][cc] = -A[r][cc];
    const int jstop = lu7_nopivot(A);
    W3DF_UNROLL
    for (int k = 0; k < 7; ++k) dx[k] = Rv[k];
    auto tcol = [&](int k) { return SINGLE ? L.o_D() + k : wait_word(L, c, k); };
    if (ct.full) {
      // tangent columns of this pass: leading block of J^-1 [C; 0] (mm10_tangent)
      if (jstop == 7) {
        W3DF_ROLLED      // (unrolled: 31.0 vs 30.5 ms)
        for (int cc = 0; cc < 6; ++cc) {
          double b[7];
          W3DF_UNROLL
          for (int r = 0; r < 6; ++r) b[r] = -P.C[r][cc];
          b[6] = 0.0;
          lu7_apply(A, b);
          W3DF_UNROLL
          for (int r = 0; r < 6; ++r) h1[tcol(cc * 6 + r)] = b[r];
        }
        lu7_apply(A, dx);
      } else {
        double Am[49], Bm[49];
        W3DF_UNROLL
        for (int r = 0; r < 7; ++r)
          W3DF_UNROLL
          for (int cc = 0; cc < 7; ++cc) Am[r * 7 + cc] = A[r][cc];
        for (int r = 0; r < 7; ++r) {
          for (int cc = 0; cc < 6; ++cc) Bm[r * 7 + cc] = (r < 6) ? -P.C[r][cc] : 0.0;
          Bm[r * 7 + 6] = dx[r];
        }
        lu7_slow(Am, jstop, Bm, 7);
        for (int r = 0; r < 6; ++r)
          for (int cc = 0; cc < 6; ++cc) h1[tcol(cc * 6 + r)] = Bm[r * 7 + cc];
        for (int k = 0; k < 7; ++k) dx[k] = Bm[k * 7 + 6];
      }
    } else if (jstop == 7) {
      lu7_apply(A, dx);
    } else {
      double Am[49], Bm[7];
      W3DF_UNROLL
      for (int r = 0; r < 7; ++r)
        W3DF_UNROLL
        for (int cc = 0; cc < 7; ++cc) Am[r * 7 + cc] = A[r][cc];
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) Bm[k] = dx[k];
      lu7_slow(Am, jstop, Bm, 1);
      W3DF_UNROLL
      for (int k = 0; k < 7; ++k) dx[k] = Bm[k];
    }
    ls2 = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 7; ++k) ls2 = ls2 + dx[k] * jtr[k];
    ls2 = c_ls * ls2;
    alpha = 1.0;
    ct.ls = 0;
    ct.stage = ST_LS;
  }

#if defined(W3D_MM10F_HIST) && defined(__CUDA_ARCH__)
  atomicAdd(&w3df_hist[w3df_trips < 61 ? w3df_trips : 61], 1);
#endif
  trips = trips + ct.trips;      // solver-loop trips of this point (the next pass is ordered by them)
  if (idle) return false;
  if (cut) return true;
  if (!elastic) {     // the converged state as kept in the workspace
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) x[k] = W(O_OST + k);
    x[6] = ott;
  }
  const double tt = x[6];
  if (elastic) {
    W3DF_UNROLL
    for (int cc = 0; cc < 6; ++cc)
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        const double t = P.C[r][cc];
        h1[L.o_D() + cc * 6 + r] = (c == 0) ? t : h1[L.o_D() + cc * 6 + r] + t;
      }
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) { h1[L.c_stress(c) + k] = x[k]; acc.sig[k] = acc.sig[k] + x[k]; }
    h1[L.c_tt(c)] = tt;
    W3DF_UNROLL
    for (int k = 0; k < 3; ++k) h1[L.c_euler(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 9; ++k) h1[L.c_Rp(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) { h1[L.c_eps(c) + k] = 0.0; h1[L.c_ep(c) + k] = 0.0; h1[L.c_ed(c) + k] = 0.0; }
    for (int k = 0; k < ns; ++k) h1[L.c_slip(c) + k] = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 15; ++k) h1[L.c_u(c) + k] = 0.0;
    return false;
  }
  {
    bool bad = tt != tt;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) bad = bad || (x[k] != x[k]);
    if (bad) return true;
  }
  if (SINGLE && !asymmetric) {
    // mm10_a_make_symm_1 on the columns the last full Newton pass stored
    W3DF_UNROLL
    for (int cc = 1; cc < 6; ++cc)
      W3DF_UNROLL
      for (int r = 0; r < cc; ++r) {
        const double t = (h1[L.o_D() + cc * 6 + r] + h1[L.o_D() + r * 6 + cc]) * 0.5;
        h1[L.o_D() + cc * 6 + r] = t;
        h1[L.o_D() + r * 6 + cc] = t;
      }
  }
  if (!SINGLE) {
    // the columns of this crystal from where they waited, symmetrised unless asymmetric assembly,
    // into the sum over the crystals (before the output below overwrites those words)
    auto wcol = [&](int k) { return wait_word(L, c, k); };
    W3DF_UNROLL
    for (int cc = 0; cc < 6; ++cc)
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        const double jr = h1[wcol(cc * 6 + r)], jt = h1[wcol(r * 6 + cc)];
        const double t = asymmetric ? jr : (jr + jt) * 0.5;
        h1[L.o_D() + cc * 6 + r] = (c == 0) ? t : h1[L.o_D() + cc * 6 + r] + t;
      }
  }
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) { h1[L.c_stress(c) + k] = x[k]; acc.sig[k] = acc.sig[k] + x[k]; }
  h1[L.c_tt(c)] = tt;
  h1[L.c_rate(c)] = tt_rate;

  // ---- mm10_output (mm10_a.f:3493) with the full-step quantities of np1, then the plastic
  //      rotation update and the Euler angles
  tinc = dt;
  {
    double D[6];
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) D[k] = in.uddt(k);
    dg = effective_dg(D);
  }
  if (W3DF_CT.alter_mode != 0) dg = W3DF_CT.eps_dot_0_y * tinc;
  double wb[3];
  {
    const double cv = tinc * cvr, itinc = 1.0 / tinc;
    double sp[6];
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) {
      double v = W(O_B6 + k) * x[0];
      W3DF_UNROLL
      for (int r = 1; r < 6; ++r) v = v + W(O_B6 + r * 6 + k) * x[r];
      sp[k] = v;
    }
    const double itt = 1.0 / tt, dgt = dg / tt;
    // lattice-frame sums: viscous part, plastic part, spin part
    double edl[6] = {0, 0, 0, 0, 0, 0}, dbl[6] = {0, 0, 0, 0, 0, 0}, wl[3] = {0, 0, 0};
    double maxslip = 0.0;
    int sysID = 0;
    W3DF_ROLLED
    for (int i0 = 0; i0 < NSF; i0 += CH) {
      double rs[CH], b[CH], pw[CH];
      W3DF_UNROLL
      for (int q = 0; q < CH; ++q) {
        const int i = i0 + q;
        rs[q] = sp[0] * W3DF_CT.mc[i][0] + sp[1] * W3DF_CT.mc[i][1] + sp[2] * W3DF_CT.mc[i][2] +
                sp[3] * W3DF_CT.mc[i][3] + sp[4] * W3DF_CT.mc[i][4] + sp[5] * W3DF_CT.mc[i][5];
        b[q] = fabs(rs[q] * itt);
      }
      pow_rate<RATE_E>(b, pw);
      W3DF_UNROLL
      for (int q = 0; q < CH; ++q) {
        const int i = i0 + q;
        const double si = dgt * pw[q] * rs[q];
        const double dif = rs[q] * cv;
        const double f = si + dif;
        W3DF_UNROLL
        for (int k = 0; k < 6; ++k) { edl[k] = edl[k] + dif * W3DF_CT.mc[i][k]; dbl[k] = dbl[k] + si * W3DF_CT.mc[i][k]; }
        W3DF_UNROLL
        for (int k = 0; k < 3; ++k) wl[k] = wl[k] + f * W3DF_CT.qcr[i][k];
        // total slip increments (plastic + viscous), the most active system
        if (i < ns) {
          h1[L.c_slip(c) + i] = f;
          h1[L.o_sliphist() + i] = h1[L.o_sliphist() + i] + f;
          const double cur = fabs(f);
          if (cur > maxslip) { maxslip = cur; sysID = i + 1; }
        }
      }
    }
    double ed[6], dbarp[6], wp[3];
    W3DF_UNROLL
    for (int r = 0; r < 6; ++r) {
      double v1 = W(O_B6 + r * 6) * edl[0], v2 = W(O_B6 + r * 6) * dbl[0];
      W3DF_UNROLL
      for (int k = 1; k < 6; ++k) { v1 = v1 + W(O_B6 + r * 6 + k) * edl[k]; v2 = v2 + W(O_B6 + r * 6 + k) * dbl[k]; }
      ed[r] = v1; dbarp[r] = v2;
    }
    W3DF_UNROLL
    for (int a = 0; a < 3; ++a)
      wp[a] = W(O_B3 + a * 3) * wl[0] + W(O_B3 + a * 3 + 1) * wl[1] + W(O_B3 + a * 3 + 2) * wl[2];
    {
      // mm10_form_wbarp: qs_i = RW qs0_i = (rt2rvw(Rp_n^T) W3) qcr_i
      double tr2[3][3], RW[3][3];
      load_tr2(tr2);
      rt2rvw(tr2, RW);
      W3DF_UNROLL
      for (int a = 0; a < 3; ++a) {
        double v = 0.0;
        W3DF_UNROLL
        for (int bq = 0; bq < 3; ++bq)
          v = v + (RW[a][0] * P.W3[0][bq] + RW[a][1] * P.W3[1][bq] + RW[a][2] * P.W3[2][bq]) * wl[bq];
        wb[a] = v;
      }
    }
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) h1[L.c_ed(c) + k] = ed[k] * itinc;
    double t1 = ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2];
    double t2 = ed[3] * ed[3] + ed[4] * ed[4] + ed[5] * ed[5];
    double ed_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ed_dot = ed_dot * itinc;
    double work_inc = 0.0;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) work_inc = work_inc + x[k] * in.uddt(k);
    // lattice strain: C^-1 stress, rotated with R
    double ee[6];
    {
      double ee0[6], erot[6][6], R[3][3];
      load_R(R);
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        double v = P.Cinv[r][0] * x[0];
        W3DF_UNROLL
        for (int k = 1; k < 6; ++k) v = v + P.Cinv[r][k] * x[k];
        ee0[r] = v;
      }
      rt2rve(R, erot);
      W3DF_UNROLL
      for (int r = 0; r < 6; ++r) {
        double v = erot[r][0] * ee0[0];
        W3DF_UNROLL
        for (int k = 1; k < 6; ++k) v = v + erot[r][k] * ee0[k];
        ee[r] = v;
      }
    }
    double ewwe[6], ep[6];
    symsw(ee, wp, ewwe);
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) { ep[k] = dbarp[k] + ewwe[k]; h1[L.c_ep(c) + k] = ep[k] * itinc; }
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    double ep_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ep_dot = ep_dot * itinc;
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) ep[k] = ep[k] + ed[k];
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    const double p_strain_inc = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    const double p_work_inc = x[0] * ep[0] + x[1] * ep[1] + x[2] * ep[2] + x[3] * ep[3] + x[4] * ep[4] + x[5] * ep[5];
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) h1[L.c_eps(c) + k] = ee[k];
    const double ec_dot = p_strain_inc * itinc;
    double n_eff;
    if (ec_dot > 0.0) {
      // ec_slip_i = ms_i . (w o ep/tinc) = mc_i . (B6^T (w o ep/tinc)),  w = (1,1,1,.5,.5,.5)
      double epw[6], epl[6];
      W3DF_UNROLL
      for (int k = 0; k < 6; ++k) epw[k] = ep[k] * itinc * (k < 3 ? 1.0 : 0.5);
      W3DF_UNROLL
      for (int k = 0; k < 6; ++k) {
        double v = W(O_B6 + k) * epw[0];
        W3DF_UNROLL
        for (int r = 1; r < 6; ++r) v = v + W(O_B6 + r * 6 + k) * epw[r];
        epl[k] = v;
      }
      const double iec = 1.0 / ec_dot;
      const double dgn = dg * W3DF_CT.rate_n / tt;
      n_eff = 0.0;
      W3DF_ROLLED
      for (int i0 = 0; i0 < NSF; i0 += CH) {
        double rs[CH], b[CH], pw[CH];
        W3DF_UNROLL
        for (int q = 0; q < CH; ++q) {
          const int i = i0 + q;
          rs[q] = sp[0] * W3DF_CT.mc[i][0] + sp[1] * W3DF_CT.mc[i][1] + sp[2] * W3DF_CT.mc[i][2] +
                  sp[3] * W3DF_CT.mc[i][3] + sp[4] * W3DF_CT.mc[i][4] + sp[5] * W3DF_CT.mc[i][5];
          b[q] = fabs(rs[q] * itt);
        }
        pow_rate<RATE_E>(b, pw);
        W3DF_UNROLL
        for (int q = 0; q < CH; ++q) {
          const int i = i0 + q;
          const double dgdt = dgn * pw[q] + cv;
          const double ec_slip = epl[0] * W3DF_CT.mc[i][0] + epl[1] * W3DF_CT.mc[i][1] + epl[2] * W3DF_CT.mc[i][2] +
                                 epl[3] * W3DF_CT.mc[i][3] + epl[4] * W3DF_CT.mc[i][4] + epl[5] * W3DF_CT.mc[i][5];
          const double u1 = rs[q] * dgdt * iec;
          const double u2 = ec_slip * iec * itinc;
          n_eff = n_eff + (2.0 / 3.0) * u1 * u2;
        }
      }
    } else {
      n_eff = 1.0e10;
    }
    double s_trace = (x[0] + x[1] + x[2]) / 3.0;
    double se[6];
    W3DF_UNROLL
    for (int k = 0; k < 6; ++k) se[k] = x[k];
    se[0] -= s_trace; se[1] -= s_trace; se[2] -= s_trace;
    t1 = se[0] * se[0] + se[1] * se[1] + se[2] * se[2];
    t2 = se[3] * se[3] + se[4] * se[4] + se[5] * se[5];
    s_trace = sqrt(1.5 * (t1 + 2.0 * t2));
    double B_eff;
    if (ec_dot < 1.0e-100) B_eff = 0.0;
    else if (n_eff > 100.0) B_eff = -1.0;
    else B_eff = ec_dot / pow(s_trace, n_eff);
    int numAct = 0;
    const double thresh = 0.1 * maxslip;
    for (int i = 0; i < ns; ++i)
      if (fabs(h1[L.c_slip(c) + i]) >= thresh) numAct = numAct + 1;
    W3DF_UNROLL
    for (int k = 0; k < 15; ++k) h1[L.c_u(c) + k] = 0.0;
    h1[L.c_u(c) + 5] = maxslip * itinc;
    h1[L.c_u(c) + 6] = (double)sysID;
    h1[L.c_u(c) + 7] = (double)numAct;
    h1[L.c_u(c) + 10] = ep_dot;
    h1[L.c_u(c) + 11] = n_eff;
    h1[L.c_u(c) + 12] = s_trace;
    h1[L.c_u(c) + 13] = B_eff;
    h1[L.c_u(c) + 14] = ed_dot;
    acc.t_work = acc.t_work + work_inc;
    acc.p_work = acc.p_work + p_work_inc;
    acc.p_strain = acc.p_strain + p_strain_inc;
  }
  // ---- mm10_update_rotation: Rp_{n+1} = exp(wbarp) Rp_n, then mm10_update_euler_angles
  {
    double Rp[3][3], Wm[3][3], Am[3][3], n_Rp[3][3], R[3][3];
    {
      double tr2[3][3];
      load_tr2(tr2);
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i)
        W3DF_UNROLL
        for (int j = 0; j < 3; ++j) n_Rp[i][j] = tr2[j][i];
    }
    load_R(R);
    Wm[0][0] = 0.0; Wm[0][1] = wb[2]; Wm[0][2] = wb[1];
    Wm[1][0] = -wb[2]; Wm[1][1] = 0.0; Wm[1][2] = wb[0];
    Wm[2][0] = -wb[1]; Wm[2][1] = -wb[0]; Wm[2][2] = 0.0;
    const double al = sqrt(Wm[1][2] * Wm[1][2] + Wm[0][2] * Wm[0][2] + Wm[0][1] * Wm[0][1]);
    if (al < 1.0e-16) {
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i)
        W3DF_UNROLL
        for (int j = 0; j < 3; ++j) Am[i][j] = 0.0;
    } else {
      double WW[3][3];
      mat33(WW, Wm, Wm);
      const double ca = (1.0 - cos(al)) / (al * al), sa = sin(al) / al;
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i)
        W3DF_UNROLL
        for (int j = 0; j < 3; ++j) Am[i][j] = ca * WW[i][j] + sa * Wm[i][j];
    }
    Am[0][0] += 1.0; Am[1][1] += 1.0; Am[2][2] += 1.0;
    mat33(Rp, Am, n_Rp);
    W3DF_UNROLL
    for (int j = 0; j < 3; ++j)
      W3DF_UNROLL
      for (int i = 0; i < 3; ++i) h1[L.c_Rp(c) + j * 3 + i] = Rp[i][j];
    const double pi = 3.141592653589793, one_eighty = 180.0;
    double w1[3][3], fullm[3][3];
    W3DF_UNROLL
    for (int i = 0; i < 3; ++i)
      W3DF_UNROLL
      for (int j = 0; j < 3; ++j) w1[i][j] = Rp[i][0] * R[j][0] + Rp[i][1] * R[j][1] + Rp[i][2] * R[j][2];
    mat33(fullm, P.g, w1);
    double psi = atan2(fullm[2][1], fullm[2][0]);
    if (psi < 0.0) psi = psi + 2.0 * pi;
    double phi = atan2(fullm[1][2], fullm[0][2]);
    if (phi < 0.0) phi = phi + 2.0 * pi;
    if (fullm[2][2] > 1.0) fullm[2][2] = 1.0;
    const double theta = acos(fullm[2][2]);
    if (W3DF_CT.angle_type == 1) {
      h1[L.c_euler(c) + 0] = one_eighty / pi * psi;
      h1[L.c_euler(c) + 1] = one_eighty / pi * theta;
      h1[L.c_euler(c) + 2] = one_eighty / pi * phi;
    } else {
      h1[L.c_euler(c) + 0] = psi; h1[L.c_euler(c) + 1] = theta; h1[L.c_euler(c) + 2] = phi;
    }
  }
  return false;
}

// mm10 (mm10_a.f:29): all crystals of one Gauss point.  hn / h1: history at n / n+1; props:
// ncrystals consecutive FProps records of this element; Rur: rot_blk_n1 column-major 9; uddt:
// un-rotated strain increment; cgn / cgn1: urcs at n / n+1 (9).  true = step cut requested.
template <bool TAUL, class SYNC, bool SINGLE = false, int RATE_E = -1>
W3DF_HD inline bool gauss_point_update_fast(int step, int ncrystals, const FProps *props, const LayoutF &L,
                                            const CHistF &hn, const HistF &h1, const WorkF &W,
                                            const PointInF &in, double dt, bool iter_0_extrapolate_off,
                                            bool asymmetric, double *cgn1, unsigned *trips_out = nullptr) {
  // At step 1 the reference initialises its block-local copy of history at n and never writes
  // that copy back; the same values are substituted on the fly (mm10_init_cc_hist0 & co.)
  const bool init = (step == 1);
  const int ns = L.nslip;
  PointSumsF acc;
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) acc.sig[k] = 0.0;
  acc.t_work = 0.0; acc.p_work = 0.0; acc.p_strain = 0.0;
  for (int k = 0; k < ns; ++k) h1[L.o_sliphist() + k] = 0.0;
  bool cut = false;
  unsigned trips = 0;
  for (int c = 0; c < ncrystals; ++c) {
    if (cut) {                      // a crystal asked for a step cut: only keep the CTA's loop barriers company
      while (SYNC::any(false)) {}
      continue;
    }
    const double n_tt = init ? W3DF_CT.tau_y + 1.0e-5 : hn[L.c_tt(c)];      // mm10_init_voche
    const double n_tt_rate = init ? 0.0 : hn[L.c_rate(c)];
    cut = crystal_update_fast<TAUL, SYNC, SINGLE, RATE_E>(props[c], W, in, n_tt, n_tt_rate, dt, iter_0_extrapolate_off,
                                          asymmetric, init, L, c, hn, h1, acc, trips);
  }
  if (trips_out) *trips_out = trips;
  if (cut) return true;
  // mm10_a_crystal_avgs + mm10_a_store_crystal
  const double irn = 1.0 / (double)ncrystals;
  W3DF_UNROLL
  for (int k = 0; k < 6; ++k) cgn1[k] = acc.sig[k] * irn;
  if (ncrystals > 1)
    for (int k = 0; k < 36; ++k) h1[L.o_D() + k] = h1[L.o_D() + k] * irn;
  W3DF_UNROLL
  for (int k = 0; k < 9; ++k) h1[L.o_R() + k] = in.rur[(long long)k * in.stride];
  for (int k = 0; k < ns; ++k)
    h1[L.o_sliphist() + k] = (init ? 0.0 : hn[L.o_sliphist() + k]) + h1[L.o_sliphist() + k] * irn;
  const double t_work_inc = acc.t_work * irn, p_work_inc = acc.p_work * irn, p_strain_inc = acc.p_strain * irn;
  cgn1[6] = in.urcs_n(6) + t_work_inc;
  cgn1[7] = in.urcs_n(7) + p_work_inc;
  cgn1[8] = in.urcs_n(8) + p_strain_inc;
  h1[L.o_work() + 0] = (init ? 0.0 : hn[L.o_work() + 0]) + t_work_inc;
  h1[L.o_work() + 1] = (init ? 0.0 : hn[L.o_work() + 1]) + p_work_inc;
  h1[L.o_work() + 2] = (init ? 0.0 : hn[L.o_work() + 2]) + p_strain_inc;
  return false;
}

// ---- host-side set-up (setup_mm10_rknstr, rknstr.f:1568-2027, re-expressed for the fast path) ----
// crystal-type record = the c_array(cnum) fields the path reads (include/w3d_engine.h:
// W3D_CRYSTAL_TYPE_DOUBLES): elast_stiff(6,6) row-major, bi(48,3), ni(48,3), 13 reals, 7 integers
inline void setup_consts(const double *ct, CrystalConsts &K) {
  const double *bi = ct + 36, *ni = ct + 36 + 144, *sc = ct + 36 + 288, *in = ct + 36 + 288 + 13;
  K.rate_n = sc[0]; K.tau_y = sc[1]; K.tau_v = sc[2]; K.voche_m = sc[3]; K.theta_0 = sc[4]; K.k_0 = sc[5];
  K.burgers = sc[6]; K.iD_v = sc[7]; K.eps_dot_0_y = sc[8];
  K.atol = sc[9]; K.atol1 = sc[10]; K.rtol = sc[11]; K.rtol1 = sc[12];
  K.nslip = (int)in[0]; K.miter = (int)in[4]; K.alter_mode = (int)in[6]; K.angle_type = 1;
  auto whole = [](double e) { return (e >= 0.0 && e <= 1024.0 && e == (double)(int)e) ? (int)e : -1; };
  K.rate_ie = whole(K.rate_n - 1.0);
  K.voche_ie = whole(K.voche_m);
  for (int s = 0; s < NSF; ++s) {
    for (int k = 0; k < 6; ++k) K.mc[s][k] = 0.0;
    for (int k = 0; k < 3; ++k) { K.qcr[s][k] = 0.0; K.ncr[s][k] = 0.0; }
  }
  for (int s = 0; s < K.nslip && s < NSF; ++s) {
    const double *b = bi + 3 * s, *n = ni + 3 * s;
    double A[3][3], As[3][3], Aa[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = b[i] * n[j];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { As[i][j] = 0.5 * (A[i][j] + A[j][i]); Aa[i][j] = 0.5 * (A[i][j] - A[j][i]); }
    K.mc[s][0] = As[0][0]; K.mc[s][1] = As[1][1]; K.mc[s][2] = As[2][2];          // mm10_ET2EV
    K.mc[s][3] = 2.0 * As[0][1]; K.mc[s][4] = 2.0 * As[1][2]; K.mc[s][5] = 2.0 * As[0][2];
    K.qcr[s][0] = Aa[1][2]; K.qcr[s][1] = Aa[0][2]; K.qcr[s][2] = Aa[0][1];        // mm10_WT2WV
    for (int k = 0; k < 3; ++k) K.ncr[s][k] = n[k];
  }
}
// mm10_rotation_matrix (mm10_a.f:1310), Kocks convention
inline void rotation_matrix(const double *angles, int angle_type /*1 degrees, 2 radians*/, double (&r)[3][3]) {
  const double pi = 3.141592653589793, one_eighty = 180.0;
  double a = angles[0], b = angles[1], c = angles[2];
  if (angle_type == 1) { a = a * pi / one_eighty; b = b * pi / one_eighty; c = c * pi / one_eighty; }
  const double psi = a, theta = b, phi = c;
  r[0][0] = -sin(psi) * sin(phi) - cos(psi) * cos(phi) * cos(theta);
  r[0][1] = cos(psi) * sin(phi) - sin(psi) * cos(phi) * cos(theta);
  r[0][2] = cos(phi) * sin(theta);
  r[1][0] = sin(psi) * cos(phi) - cos(psi) * sin(phi) * cos(theta);
  r[1][1] = -cos(psi) * cos(phi) - sin(psi) * sin(phi) * cos(theta);
  r[1][2] = sin(phi) * sin(theta);
  r[2][0] = cos(psi) * sin(theta);
  r[2][1] = sin(psi) * sin(theta);
  r[2][2] = cos(theta);
}
// one (element, crystal) record.  Returns false when the rotated stiffness is singular.
inline bool setup_fprops(const double *ct, const double *angles, int angle_type, FProps &P) {
  rotation_matrix(angles, angle_type, P.g);
  double tr[3][3], Rs[6][6];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tr[i][j] = P.g[j][i];
  rt2rve(tr, Rs);
  rt2rvw(tr, P.W3);
  // strain-type vectors transform with Dm T_sigma Dm^-1, Dm = diag(1,1,1,2,2,2)
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) P.E6[r][c] = Rs[r][c] * (r >= 3 ? 2.0 : 1.0) / (c >= 3 ? 2.0 : 1.0);
  double t66[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += ct[i * 6 + k] * Rs[j][k];
    t66[i][j] = s;
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += Rs[i][k] * t66[k][j];
    P.C[i][j] = s;
  }
  // Cinv by Gauss-Jordan with partial pivoting
  double a[6][12];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { a[i][j] = P.C[i][j]; a[i][6 + j] = (i == j) ? 1.0 : 0.0; }
  for (int j = 0; j < 6; ++j) {
    int p = j;
    for (int i = j + 1; i < 6; ++i) if (fabs(a[i][j]) > fabs(a[p][j])) p = i;
    if (a[p][j] == 0.0) return false;
    if (p != j) for (int k = 0; k < 12; ++k) { const double t = a[j][k]; a[j][k] = a[p][k]; a[p][k] = t; }
    const double r = 1.0 / a[j][j];
    for (int k = 0; k < 12; ++k) a[j][k] *= r;
    for (int i = 0; i < 6; ++i) if (i != j) { const double f = a[i][j]; for (int k = 0; k < 12; ++k) a[i][k] -= f * a[j][k]; }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) P.Cinv[i][j] = a[i][6 + j];
  P.pad[0] = P.pad[1] = 0.0;
  return true;
}

}  // namespace w3d_mm10f
