// w3d_mm10_core.h -- crystal-plasticity material update (mm10) for ONE Gauss point: the engine's
// general core (any slip table up to NS systems, several crystal types), device code for k_mm10<NS>
// and host code for the property set-up.  The arithmetic is re-ordered for the GPU (one power per
// slip system per state, sums over slip systems taken before the 6x6 products, register-resident
// LU); the parity tests check it against an INDEPENDENT statement-for-statement CPU restatement
// of the Fortran that lives with the test infrastructure -- no file is shared with it.
//
// Follows reference src/: mm10_a.f (mm10 :29, mm10_a_do_crystal :204, mm10_a_store_crystal :290,
// mm10_init_cc_hist0 :480, mm10_tangent :664, mm10_setup :852, mm10_store_cryhist :998,
// mm10_solve_crystal :1102, mm10_update_euler_angles :1195, mm10_rt2rve :1431, mm10_rt2rvw :1492,
// mm10_setup_np1 :1689, mm10_init_voche :2071, mm10_setup_voche :2086, mm10_copy_cc_hist :2493,
// mm10_solve_strup :2657, mm10_solve :2901 (predict :3032, update :3224), mm10_update_rotation
// :3375, mm10_expmw3x3 :3447, mm10_output :3493), mm10_b.f (mm10_formR2 :63, mm10_formJ11 :177,
// mm10_formJ12 :281, mm10_formJ21 :377, mm10_formJ22 :439, mm10_formR1 :1073, mm10_form_wbarp
// :1239, mm10_form_wp :1312, mm10_symsw :1513, mm10_iw :1613, voche routines :1813-2074),
// mm10_d.f :25-330 (history layout), rknstr.f :1568-2027 (setup_mm10_rknstr).
// Scope: hardening model 1 (Voce/"voche"), default Newton solver (props%solver), tang_calc 0,
// any slip-system table up to NS systems, any number of crystals per point.
#pragma once
#include <math.h>

// Performance notes (round 1 measurements, 64k-element check): a thread-per-point transcription
// with every array on the thread's stack is bound by local-memory traffic (6 KB per thread, more
// than L2 holds for the resident threads: ncu shows ~60 GB of DRAM traffic per launch), not by
// FP64 issue.  Hence: per-system Schmid data behind a strided pointer (shared memory on the
// device), results stored straight to history instead of a staging record, sums over slip
// systems taken before the 6x6 products.  __noinline__ residual/Jacobian functions were tried and
// are slower (call + stack traffic).
#if defined(__CUDACC__)
#define W3D_HD __host__ __device__
#define W3D_ROLLED _Pragma("unroll 1")   // keep a loop rolled on the device: code size, see notes
#else
#define W3D_HD
#define W3D_ROLLED
#endif

// optional host-side operation counters (oracle built with -DW3D_MM10_STATS; bench.py uses them
// to state the algorithmic work of a launch): [0] slip evaluations (one pow per system each),
// [1] Jacobians, [2] 7x7 / 6x6 solves, [3] crystal updates that ran the solver, [4] sub-increment solves,
// [5] solves that needed a row interchange (general LU instead of the register-resident one)
#if defined(W3D_MM10_STATS) && !defined(__CUDA_ARCH__)
extern "C" long long w3d_mm10_stats[8];
#define W3D_STAT(k) __atomic_fetch_add(&w3d_mm10_stats[k], 1LL, __ATOMIC_RELAXED)
#else
#define W3D_STAT(k) ((void)0)
#endif

namespace w3d_mm10 {

// strided view of one Gauss point's history vector (stride = span on the CPU block layout,
// = number of Gauss-point slots in the engine's structure-of-arrays layout)
struct Hist {
  double *p;
  long long stride;
  W3D_HD double &operator[](int k) const { return p[(long long)k * stride]; }
};
struct CHist {
  const double *p;
  long long stride;
  W3D_HD double operator[](int k) const { return p[(long long)k * stride]; }
};

// history layout of mm10_set_history_locs (mm10_d.f:153-330), 0-based offsets
struct Layout {
  int nslip, num_hard, len_slip, len_tt, len_u, len_rate;
  int common_size, crystal_size;
  // common part: [D 36][gradFe 27][R 9][work 3][slip history len_slip]
  // per crystal: stress 6, euler 3, Rp 9, D 6, eps 6, slip_incs, tau_tilde, u, tt_rate, ep 6, ed 6
  W3D_HD static Layout make(int nslip, int num_hard) {
    Layout L;
    const bool use_max = (num_hard == 48) || (nslip == 48);   // max_uhard = max_slip_sys = 48
    L.nslip = nslip; L.num_hard = num_hard;
    L.len_slip = use_max ? 48 : nslip;
    L.len_tt = use_max ? 48 : num_hard;
    L.len_u = use_max ? 48 : 15;
    L.len_rate = use_max ? 48 : num_hard;
    L.common_size = 36 + 27 + 9 + 3 + L.len_slip;
    L.crystal_size = 6 + 3 + 9 + 6 + 6 + L.len_slip + L.len_tt + L.len_u + L.len_rate + 6 + 6;
    return L;
  }
  W3D_HD int size(int ncrystals) const { return common_size + ncrystals * crystal_size; }
  W3D_HD int o_D() const { return 0; }
  W3D_HD int o_gradfe() const { return 36; }
  W3D_HD int o_R() const { return 63; }
  W3D_HD int o_work() const { return 72; }
  W3D_HD int o_sliphist() const { return 75; }
  W3D_HD int c0(int c) const { return common_size + c * crystal_size; }
  W3D_HD int c_stress(int c) const { return c0(c); }
  W3D_HD int c_euler(int c) const { return c0(c) + 6; }
  W3D_HD int c_Rp(int c) const { return c0(c) + 9; }
  W3D_HD int c_D(int c) const { return c0(c) + 18; }
  W3D_HD int c_eps(int c) const { return c0(c) + 24; }
  W3D_HD int c_slip(int c) const { return c0(c) + 30; }
  W3D_HD int c_tt(int c) const { return c_slip(c) + len_slip; }
  W3D_HD int c_u(int c) const { return c_tt(c) + len_tt; }
  W3D_HD int c_rate(int c) const { return c_u(c) + len_u; }
  W3D_HD int c_ep(int c) const { return c_rate(c) + len_rate; }
  W3D_HD int c_ed(int c) const { return c_ep(c) + 6; }
};

// per (element, crystal) properties = local_work%c_props(i,c) after setup_mm10_rknstr, packed
// as doubles so the same record is an HBM array on the device.  Matrices are row-major [r][c].
template <int NS>
struct Props {
  double rate_n, tau_y, tau_v, voche_m, theta_0, k_0, burgers, iD_v, eps_dot_0_y;
  double atol, atol1, rtol, rtol1;
  double init_angles[3];
  double g[3][3];            // rotation_g
  double stiffness[6][6];    // init_elast_stiff rotated to the reference frame
  double ms[NS][6], qs[NS][3], ns[NS][3];
  double nslip_d, miter_d, alter_mode_d, angle_type_d;   // small integers kept as doubles
  W3D_HD int nslip() const { return (int)nslip_d; }
  W3D_HD int miter() const { return (int)miter_d; }
};

// ---- small dense helpers ---------------------------------------------------------------------
W3D_HD inline void rt2rve(const double (&rt)[3][3], double (&rv)[6][6]) {   // mm10_a.f:1431
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
W3D_HD inline void rt2rvw(const double (&rt)[3][3], double (&rv)[3][3]) {   // mm10_a.f:1492
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
W3D_HD inline void mat33(double (&a)[3][3], const double (&b)[3][3], const double (&c)[3][3]) {  // a = b c
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j];
}
W3D_HD inline void mat33_bt(double (&a)[3][3], const double (&b)[3][3], const double (&c)[3][3]) {  // a = b c^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = b[i][0] * c[j][0] + b[i][1] * c[j][1] + b[i][2] * c[j][2];
}
W3D_HD inline void mv6(double (&a)[6], const double (&b)[6][6], const double *c) {   // mult_type_2
  for (int i = 0; i < 6; ++i) a[i] = b[i][0] * c[0];
  for (int j = 1; j < 6; ++j)
    for (int i = 0; i < 6; ++i) a[i] = a[i] + b[i][j] * c[j];
}
W3D_HD inline void mv3(double *a, const double (&b)[3][3], const double *c) {       // mult_type_3
  for (int i = 0; i < 3; ++i) a[i] = b[i][0] * c[0] + b[i][1] * c[1] + b[i][2] * c[2];
}
W3D_HD inline void symsw(const double *s, const double *w, double (&sw)[6]) {      // mm10_b.f:1513
  const double half = 0.5;
  sw[0] = s[3] * w[2] - s[5] * w[1];
  sw[1] = s[3] * w[2] - s[4] * w[0];
  sw[2] = s[5] * w[1] + s[4] * w[0];
  sw[3] = half * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]);
  sw[4] = half * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]);
  sw[5] = half * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4]);
}
// dense solve A x = b (n <= 7) by LU with partial pivoting in DGETF2/DGETRS order (column
// pivot = first largest |a|, row interchanges applied to the whole matrix); A is destroyed,
// b (nrhs columns, b[r][k]) is overwritten with the solution.  Replaces DGESV.
template <int N, int NRHS>
W3D_HD inline int lu_solve(double (&A)[N][N], double (&B)[N][NRHS], int n, int nrhs) {
  int info = 0;
  double rinv[N];
  W3D_ROLLED
  for (int j = 0; j < n; ++j) {
    rinv[j] = 0.0;
    int p = j;
    double amax = fabs(A[j][j]);
    for (int i = j + 1; i < n; ++i)
      if (fabs(A[i][j]) > amax) { amax = fabs(A[i][j]); p = i; }
    if (A[p][j] != 0.0) {
      if (p != j) {
        for (int k = 0; k < n; ++k) { const double t = A[j][k]; A[j][k] = A[p][k]; A[p][k] = t; }
        for (int k = 0; k < nrhs; ++k) { const double t = B[j][k]; B[j][k] = B[p][k]; B[p][k] = t; }
      }
      const double r = 1.0 / A[j][j];
      rinv[j] = r;
      for (int i = j + 1; i < n; ++i) A[i][j] *= r;
    } else if (info == 0) {
      info = j + 1;
    }
    for (int i = j + 1; i < n; ++i)
      for (int k = j + 1; k < n; ++k) A[i][k] -= A[i][j] * A[j][k];
  }
  for (int k = 0; k < nrhs; ++k) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) B[i][k] -= A[i][j] * B[j][k];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) B[i][k] -= A[i][j] * B[j][k];
      B[i][k] *= rinv[i];
    }
  }
  return info;
}
// The same elimination when no row interchange is needed (every diagonal entry is already the
// first largest of its column, the usual case for the Newton matrices here: I + small terms).
// All indices are compile-time, so A and B live in registers and the code is a straight run of
// multiply-adds: operation for operation what lu_solve does when its pivot search returns p = j.
// Returns false -- with A and B destroyed -- as soon as a column would need an interchange; the
// caller then runs lu_solve on a fresh copy.
#if defined(__CUDA_ARCH__)
#define W3D_UNROLL _Pragma("unroll")
#else
#define W3D_UNROLL
#endif
template <int N>
W3D_HD inline bool lu_factor_nopivot(double (&A)[N][N], double (&rinv)[N]) {
  W3D_UNROLL
  for (int j = 0; j < N; ++j) {
    const double amax = fabs(A[j][j]);
    bool ok = (A[j][j] != 0.0);
    W3D_UNROLL
    for (int i = j + 1; i < N; ++i) ok = ok && !(fabs(A[i][j]) > amax);
    if (!ok) return false;
    const double r = 1.0 / A[j][j];
    rinv[j] = r;
    W3D_UNROLL
    for (int i = j + 1; i < N; ++i) {
      A[i][j] *= r;
      W3D_UNROLL
      for (int k = j + 1; k < N; ++k) A[i][k] -= A[i][j] * A[j][k];
    }
  }
  return true;
}
// forward and back substitution with the factors of lu_factor_nopivot, one right-hand side
template <int N>
W3D_HD inline void lu_apply(const double (&A)[N][N], const double (&rinv)[N], double (&b)[N]) {
  W3D_UNROLL
  for (int i = 0; i < N; ++i) {
    W3D_UNROLL
    for (int j = 0; j < i; ++j) b[i] -= A[i][j] * b[j];
  }
  W3D_UNROLL
  for (int i = N - 1; i >= 0; --i) {
    W3D_UNROLL
    for (int j = i + 1; j < N; ++j) b[i] -= A[i][j] * b[j];
    b[i] *= rinv[i];
  }
}
template <int N, int NRHS>
W3D_HD inline bool lu_solve_nopivot(double (&A)[N][N], double (&B)[N][NRHS]) {
  double rinv[N];
  if (!lu_factor_nopivot<N>(A, rinv)) return false;
  W3D_UNROLL
  for (int k = 0; k < NRHS; ++k) {
    double b[N];
    W3D_UNROLL
    for (int i = 0; i < N; ++i) b[i] = B[i][k];
    lu_apply<N>(A, rinv, b);
    W3D_UNROLL
    for (int i = 0; i < N; ++i) B[i][k] = b[i];
  }
  return true;
}

// ---- state of one crystal --------------------------------------------------------------------
// Schmid data of the current crystal in the unrotated configuration: ms_i = RE ms0_i (6) and the
// spin vector qc_i = RQ qs0_i (3), with RE / RQ the 6x6 / 3x3 operators of the rotations
// Rp_n^T and R Rp_n^T (mm10_setup) and ms0 / qs0 the element's reference-frame tables (Props).
// The CPU build computes them once per crystal into a plain array (sl, stride 1), as the
// reference does.  The device build recomputes them from RE / RQ wherever a loop over the slip
// systems needs them (45 multiply-adds per system, bit-identical values): 864 bytes per thread of
// storage cap the resident warps, the FP64 pipe has the room (see the notes at the top).
template <int NS>
struct Work {
  // quantities at n read from history
  double n_stress[6], n_Rp[3][3], n_D[6], n_tt, n_tt_rate;
  // set up for n+1 (mm10_setup)
  double R[3][3], RE[6][6], RQ[3][3], tau_l[NS], mu_harden;
  // the (sub-)increment being solved: curr%D, curr%tinc, curr%dg
  double D[6], tinc, dg, tt_rate;
  double *sl;
  int sls;
  W3D_HD double &ms(int i, int k) const { return sl[(6 * i + k) * sls]; }
  W3D_HD double &q(int i, int k) const { return sl[(6 * NS + 3 * i + k) * sls]; }
};
template <int NS>
W3D_HD constexpr int slip_store_doubles() { return 9 * NS; }

template <int NS>
W3D_HD inline void slip_system(const Props<NS> &P, const Work<NS> &w, int i, double (&m)[6],
                               double (&q)[3]) {
#if defined(__CUDA_ARCH__) && !defined(W3D_MM10_SLIP_STORE)
  mv6(m, w.RE, P.ms[i]);
  mv3(q, w.RQ, P.qs[i]);
#else
  for (int k = 0; k < 6; ++k) m[k] = w.ms(i, k);
  for (int k = 0; k < 3; ++k) q[k] = w.q(i, k);
#endif
}
template <int NS>
W3D_HD inline double effective_dg(const double *D) {                            // mm10_setup :880
  const double t1 = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
  const double t2 = D[3] * D[3] + D[4] * D[4] + D[5] * D[5];
  return sqrt((2.0 / 3.0) * (t1 + 0.5 * t2));
}

// x^e for x >= 0.  The exponents on this path are material constants that are whole numbers in
// every deck of the reference (rate sensitivity n - 1 = 19 in test82/voche_model, 9..49 in the
// example problems; Voce exponent m = 1): those are evaluated by binary powering (a handful of
// multiplies, within a few ulp of pow(), like two libm's differ) -- pow() itself is ~290 SASS
// instructions per call on the device and was a third of all instructions executed.  Anything
// else goes to pow().
struct Power {
  double e;
  int ie;      // >= 0: e is this whole number; -1: general exponent
  W3D_HD static Power make(double e) {
    Power p;
    p.e = e;
    p.ie = (e >= 0.0 && e <= 1024.0 && e == (double)(int)e) ? (int)e : -1;
    return p;
  }
  W3D_HD double operator()(double x) const {
    if (ie < 0) return pow(x, e);
    double r = 1.0, b = x;
    for (int k = ie; k != 0; k >>= 1) {
      if (k & 1) r = r * b;
      b = b * b;
    }
    return r;
  }
};

// Voce hardening term of one slip system (mm10_h_voche, mm10_b.f:1896):
//   h_term = 1 - (tt - tau_y)/tau_v + tau_l/(tt - tau_y),
// kept with |h_term|^m and 1/|h_term| in a one-entry memo: without geometric hardening (k_0 = 0)
// tau_l = 0 and h_term is the same for every slip system.  Divisions by tt - tau_y and by
// |h_term| are multiplications by reciprocals taken once per state: the operands are often zero
// or tiny (tau_l = 0, slip increments of inactive systems ~1e-40), which sends every device
// division down its ~65-instruction slow path.
struct PowMemo {
  double x, y, rx;   // h_term, |h_term|^m, 1/|h_term|
  bool have;
};
struct HardTerm {
  double a, inv_d;   // h_term = a + tau_l * inv_d
};
template <int NS>
W3D_HD inline HardTerm hard_term(const Props<NS> &P, double tt) {
  HardTerm h;
  const double d = tt - P.tau_y;
  h.inv_d = 1.0 / d;
  h.a = 1.0 - d / P.tau_v;
  return h;
}
W3D_HD inline void hterm_eval(PowMemo &m, double h_term, double voche_m) {
  if (!(m.have && m.x == h_term)) {
    m.x = h_term;
    m.y = Power::make(voche_m)(fabs(h_term));
    m.rx = 1.0 / fabs(h_term);
    m.have = true;
  }
}

// The reference re-evaluates the resolved shear stress and the slip increment of every system
// in each of mm10_form_dbarp, mm10_form_wp, mm10_h_voche, mm10_dgdh_voche, ... (3-4 times per
// residual, each with a pow()).  They are pure functions of (stress, tt), so one evaluation per
// system per state is what both builds of this core do.
//
// All powers of one state come from ONE pow() per slip system, the slip increment's own
//   pw_i = |rs_i/tt|^(n-1):
//   dg n / tt^n * |rs_i|^(n-1)             = dg n pw_i / tt              (mm10_dgdt_voche),
//   dg n / tt^n * |rs_i|^(n-2) * rs_i      = dg n pw_i sign(rs_i) / tt   (mm10_estress_voche),
// which differs from the reference's separate pow() calls by an ulp or two (as two libm's
// differ) and takes the Jacobian from 49 pow() per state to none: it is always formed at the
// state whose residual was evaluated last, so it reuses that evaluation.
template <int NS>
struct Slips {
  double rs[NS], si[NS], pw[NS];
  PowMemo pm;                       // |h_term|^m of this state, shared by formR2 and formJ
};
// One pass over the slip systems for the state (stress, tt): resolved shear stresses and slip
// increments into s (mm10_rs, mm10_slipinc), and mm10_formR1 (mm10_b.f:1073):
//   R1 = stress - n%stress - C (D - dbarp) + 2 sym(T W),
// dbarp = sum (slipinc + rs tinc iD_v) ms (mm10_form_dbarp), W from sum (...) qc (mm10_form_wp)
template <int NS>
W3D_HD inline void formR1(const Props<NS> &P, const Work<NS> &w, const double *stress, double tt,
                          Slips<NS> &s, double *R1) {
  W3D_STAT(0);
  // ms_i = RE ms0_i and qc_i = RQ qs0_i enter only through stress . ms_i and the two sums, so the
  // rotations are applied to the stress and to the sums instead of to every system:
  //   rs_i = (RE^T stress) . ms0_i,   dbarp = RE sum f_i ms0_i,   wp = RQ sum f_i qs0_i
  double sr[6], a6[6] = {0, 0, 0, 0, 0, 0}, a3[3] = {0, 0, 0};
  for (int k = 0; k < 6; ++k)
    sr[k] = w.RE[0][k] * stress[0] + w.RE[1][k] * stress[1] + w.RE[2][k] * stress[2] +
            w.RE[3][k] * stress[3] + w.RE[4][k] * stress[4] + w.RE[5][k] * stress[5];
  const int ns = P.nslip();
  const double cv = w.tinc * P.iD_v, dgt = w.dg / tt, itt = 1.0 / tt;
  const Power rate_pow = Power::make(P.rate_n - 1.0);
  for (int i = 0; i < ns; ++i) {
    const double rs = sr[0] * P.ms[i][0] + sr[1] * P.ms[i][1] + sr[2] * P.ms[i][2] +
                      sr[3] * P.ms[i][3] + sr[4] * P.ms[i][4] + sr[5] * P.ms[i][5];
    const double pw = rate_pow(fabs(rs * itt));
    const double si = dgt * pw * rs;
    s.rs[i] = rs; s.pw[i] = pw; s.si[i] = si;
    const double f = rs * cv + si;
    for (int k = 0; k < 6; ++k) a6[k] = a6[k] + f * P.ms[i][k];
    for (int k = 0; k < 3; ++k) a3[k] = a3[k] + f * P.qs[i][k];
  }
  s.pm.have = false;
  double dbarp[6], wp[3];
  mv6(dbarp, w.RE, a6);
  mv3(wp, w.RQ, a3);
  double symTW[6], v1[6], v2[6];
  symsw(stress, wp, symTW);
  for (int k = 0; k < 6; ++k) v1[k] = w.D[k] - dbarp[k];
  mv6(v2, P.stiffness, v1);
  for (int k = 0; k < 6; ++k) R1[k] = stress[k] - w.n_stress[k] - v2[k] + 2.0 * symTW[k];
}
// mm10_h_voche (mm10_b.f:1896) and mm10_formR2 (:63); also records curr%tt_rate
template <int NS>
W3D_HD inline double formR2(const Props<NS> &P, Work<NS> &w, Slips<NS> &s, double tt) {
  double h = 0.0;
  const int ns = P.nslip();
  const HardTerm ht = hard_term(P, tt);
  for (int i = 0; i < ns; ++i) {
    const double h_term = ht.a + w.tau_l[i] * ht.inv_d;
    hterm_eval(s.pm, h_term, P.voche_m);
    h = h + s.pm.y * copysign(1.0, h_term) * fabs(s.si[i]);
  }
  h = w.n_tt + P.theta_0 * h;
  w.tt_rate = (h - w.n_tt) / w.tinc;
  return tt - h;
}
template <int NS>
W3D_HD inline void formR(const Props<NS> &P, Work<NS> &w, const double *x, double *R,
                         Slips<NS> &s) {
  formR1(P, w, x, x[6], s, R);
  R[6] = formR2(P, w, s, x[6]);
}

// Jacobian of the local system (mm10_formJ, mm10_b.f:909; J11 :177, J12 :281, J21 :377, J22 :439)
// for Voce with one hardening variable, at the state s was evaluated for.  The reference sums,
// per slip system, dgdt_i (C ms_i + 2 symsw(T, qc_i)) (x) ms_i and dgdh_i (C ms_i + 2 symsw(T,
// qc_i)); C and symsw(T, .) are linear, so the sums over systems are taken first:
//   M  = sum dgdt_i ms_i (x) ms_i   (symmetric),   Q  = sum dgdt_i qc_i (x) ms_i,
//   a6 = sum dgdh_i ms_i,                           q3 = sum dgdh_i qc_i,
//   J11 = C M + 2 symsw(T, Q) + Iw + I,             J12 = C a6 + 2 symsw(T, q3)
// -- the same numbers to rounding with about 40 % of the multiply-adds.  full = false forms J11
// only (stress predictor).
template <int NS>
W3D_HD inline void formJ(const Props<NS> &P, const Work<NS> &w, Slips<NS> &s, const double *stress,
                         double tt, bool full, double (&J)[7][7]) {
  const int ns = P.nslip();
  const double dgn = w.dg * P.rate_n / tt, cv = w.tinc * P.iD_v, nott = P.rate_n / tt;
  W3D_STAT(1);
  const HardTerm ht = hard_term(P, tt);
  const double mitv = -1.0 / P.tau_v;
  double M[6][6], Q[3][6], a6[6], q3[3], et[6], wp[3];
  for (int r = 0; r < 6; ++r) {
    for (int c = r; c < 6; ++c) M[r][c] = 0.0;
    a6[r] = 0.0; et[r] = 0.0;
  }
  for (int a = 0; a < 3; ++a) {
    for (int c = 0; c < 6; ++c) Q[a][c] = 0.0;
    q3[a] = 0.0; wp[a] = 0.0;
  }
  double etau = 0.0;
  for (int i = 0; i < ns; ++i) {
    double m[6], gm[6], qc[3];
    slip_system(P, w, i, m, qc);
    const double g = dgn * s.pw[i] + cv;                          // mm10_dgdt_voche + viscous term
    for (int k = 0; k < 6; ++k) gm[k] = g * m[k];
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c) M[r][c] = M[r][c] + gm[r] * m[c];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 6; ++c) Q[a][c] = Q[a][c] + qc[a] * gm[c];
    const double fw = s.si[i] + s.rs[i] * cv;                     // mm10_form_wp for mm10_IW
    for (int a = 0; a < 3; ++a) wp[a] = wp[a] + fw * qc[a];
    if (full) {
      const double dgam = s.si[i];
      const double dgdh = -nott * dgam;                           // mm10_dgdh_voche
      for (int k = 0; k < 6; ++k) a6[k] = a6[k] + dgdh * m[k];
      for (int a = 0; a < 3; ++a) q3[a] = q3[a] + dgdh * qc[a];
      // J21 = -estress (mm10_estress_voche)
      const double h_term = ht.a + w.tau_l[i] * ht.inv_d;
      hterm_eval(s.pm, h_term, P.voche_m);
      const double hp = s.pm.y;
      const double f = hp * copysign(1.0, h_term) * s.pw[i] * copysign(1.0, s.rs[i]);
      for (int k = 0; k < 6; ++k) et[k] = et[k] + f * m[k];
      // J22 (mm10_ehard_voche)
      etau = etau + (P.voche_m * (mitv - w.tau_l[i] * ht.inv_d * ht.inv_d) * fabs(dgam) * s.pm.rx -
                     fabs(dgam) * nott * copysign(1.0, h_term) * copysign(1.0, dgam)) *
                        hp;
    }
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < r; ++c) M[r][c] = M[c][r];
  for (int c = 0; c < 6; ++c) {
    double sq[6];
    const double qcol[3] = {Q[0][c], Q[1][c], Q[2][c]};
    symsw(stress, qcol, sq);
    for (int r = 0; r < 6; ++r) {
      double v = 2.0 * sq[r];
      for (int k = 0; k < 6; ++k) v = v + P.stiffness[r][k] * M[k][c];
      J[r][c] = v;
    }
  }
  // mm10_IW (mm10_b.f:1613)
  J[0][3] += 2.0 * wp[2]; J[0][5] += -2.0 * wp[1];
  J[1][3] += 2.0 * wp[2]; J[1][4] += -2.0 * wp[0];
  J[2][4] += 2.0 * wp[0]; J[2][5] += 2.0 * wp[1];
  J[3][0] += wp[2]; J[3][1] += -wp[2]; J[3][4] += -wp[1]; J[3][5] += wp[0];
  J[4][1] += wp[0]; J[4][2] += -wp[0]; J[4][3] += wp[1]; J[4][5] += wp[2];
  J[5][0] += wp[1]; J[5][2] += -wp[1]; J[5][3] += wp[0]; J[5][4] += -wp[2];
  for (int r = 0; r < 6; ++r) J[r][r] = J[r][r] + 1.0;
  if (!full) return;
  {
    double sq[6];
    symsw(stress, q3, sq);
    for (int r = 0; r < 6; ++r) {
      double v = 2.0 * sq[r];
      for (int k = 0; k < 6; ++k) v = v + P.stiffness[r][k] * a6[k];
      J[r][6] = v;
    }
  }
  const double sc = P.theta_0 * w.dg * nott;
  for (int k = 0; k < 6; ++k) J[6][k] = -(sc * et[k]);
  J[6][6] = 1.0 - P.theta_0 * etau;
}

template <int NS>
W3D_HD inline bool any_nan(const double *x, int n) {
  for (int k = 0; k < n; ++k)
    if (x[k] != x[k]) return true;
  return false;
}

// mm10_solve (mm10_a.f:2901): stress prediction (:3032) then material update (:3224).
// x = [stress(6), tt]; returns fail.  J holds the last Jacobian formed by the update loop.
template <int NS>
W3D_HD inline bool solve(const Props<NS> &P, Work<NS> &w, double *stress, double &tt, double dtinc,
                         double (&J)[7][7], Slips<NS> &s) {
  const double c = 1.0e-4, red = 0.5;
  const int mls = 10, mmin = 1;
  const int miter = P.miter();
  bool fail = false;
  double x[7];
  for (int k = 0; k < 6; ++k) x[k] = stress[k];
  x[6] = tt;
  double inR1;
  // s: slip evaluation of the state whose residual was formed last; on a successful return that
  // is the converged (stress, tt)
  {  // ---- mm10_solve_predict
    int iter = 0;
    double x1[6], d1[6], d2[6], R1[6];
    for (int k = 0; k < 6; ++k) x1[k] = x[k];
    for (int k = 0; k < 6; ++k) d1[k] = w.D[k];
    double dtrace = (d1[0] + d1[1] + d1[2]) / 3.0;
    d1[0] -= dtrace; d1[1] -= dtrace; d1[2] -= dtrace;
    double t1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
    double t2 = d1[3] * d1[3] + d1[4] * d1[4] + d1[5] * d1[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d1[k] = 0.0; }
    else { const double sn = sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d1[k] = d1[k] / sn; }
    for (int k = 0; k < 6; ++k) d2[k] = w.n_D[k];
    dtrace = (d2[0] + d2[1] + d2[2]) / 3.0;
    d2[0] -= dtrace; d2[1] -= dtrace; d2[2] -= dtrace;
    t1 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
    t2 = d2[3] * d2[3] + d2[4] * d2[4] + d2[5] * d2[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d2[k] = 0.0; }
    else { const double sn = sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d2[k] = d2[k] / sn; }
    t1 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
    t2 = d1[3] * d2[3] + d1[4] * d2[4] + d1[5] * d2[5];
    const double cos_ang = fmax(t1 + t2, 0.0);
    const double x2 = x[6] + cos_ang * w.n_tt_rate * dtinc;
    formR1(P, w, x1, x2, s, R1);
    double nR1 = sqrt(R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] +
                      R1[4] * R1[4] + R1[5] * R1[5]);
    inR1 = nR1;
    bool pfail = false;
    while ((nR1 > P.atol1) && (nR1 / inR1 > P.rtol1)) {
      formJ(P, w, s, x1, x2, false, J);               // s: evaluated with R1 at (x1, x2)
      double A[6][6], dx[6][1], wv[6];
      for (int r = 0; r < 6; ++r) {
        dx[r][0] = R1[r];
        for (int cc = 0; cc < 6; ++cc) A[r][cc] = -J[r][cc];
      }
      double dotR1 = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] +
                     R1[4] * R1[4] + R1[5] * R1[5];
      const double ls1 = 0.5 * dotR1;
      for (int k = 0; k < 6; ++k)   // trans(J11) R1
        wv[k] = J[0][k] * R1[0] + J[1][k] * R1[1] + J[2][k] * R1[2] + J[3][k] * R1[3] +
                J[4][k] * R1[4] + J[5][k] * R1[5];
      W3D_STAT(2);
      if (!lu_solve_nopivot<6, 1>(A, dx)) {           // an interchange is needed: general routine
        W3D_STAT(5);
        double Ap[6][6];
        for (int r = 0; r < 6; ++r) {
          dx[r][0] = R1[r];
          for (int cc = 0; cc < 6; ++cc) Ap[r][cc] = -J[r][cc];
        }
        lu_solve<6, 1>(Ap, dx, 6, 1);
      }
      double alpha = 1.0;
      const double ls2 = c * (dx[0][0] * wv[0] + dx[1][0] * wv[1] + dx[2][0] * wv[2] +
                              dx[3][0] * wv[3] + dx[4][0] * wv[4] + dx[5][0] * wv[5]);
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xnew[6];
        for (int k = 0; k < 6; ++k) xnew[k] = x1[k] + alpha * dx[k][0];
        formR1(P, w, xnew, x2, s, R1);
        dotR1 = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] +
                R1[5] * R1[5];
        nR1 = sqrt(dotR1);
        const double nRs = 0.5 * dotR1;
        if ((nRs <= nlsx) || (ls > mls)) {
          for (int k = 0; k < 6; ++k) x1[k] = xnew[k];
          break;
        }
        alpha = red * alpha;
        ls = ls + 1;
      }
      iter = iter + 1;
      if ((iter > miter) || any_nan<NS>(x1, 6)) { pfail = true; break; }
    }
    if (pfail) {
      fail = true;   // the reference returns from the predictor only; the update still runs on
                     // the un-predicted x and fail stays set (mm10_a.f:3187-3203)
    } else {
      for (int k = 0; k < 6; ++k) x[k] = x1[k];
      x[6] = x2;
    }
  }
  {  // ---- mm10_solve_update
    int iter = 0;
    double R[7];
    formR(P, w, x, R, s);
    double dotR = 0.0;
    for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
    double nR = sqrt(dotR);
    double inR = nR;
    if (inR == 0.0) inR = inR1;
    while (((nR > P.atol) && (nR / inR > P.rtol)) || (iter < mmin)) {
      formJ(P, w, s, x, x[6], true, J);               // s: evaluated with R at x
      double A[7][7], dx[7][1], wv[7];
      for (int r = 0; r < 7; ++r) {
        dx[r][0] = R[r];
        for (int cc = 0; cc < 7; ++cc) A[r][cc] = -J[r][cc];
      }
      dotR = 0.0;
      for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
      const double ls1 = 0.5 * dotR;
      for (int k = 0; k < 7; ++k) {   // trans(J) R
        double sj = 0.0;
        for (int r = 0; r < 7; ++r) sj = sj + J[r][k] * R[r];
        wv[k] = sj;
      }
      W3D_STAT(2);
      if (!lu_solve_nopivot<7, 1>(A, dx)) {           // an interchange is needed: general routine
        W3D_STAT(5);
        double Ap[7][7];
        for (int r = 0; r < 7; ++r) {
          dx[r][0] = R[r];
          for (int cc = 0; cc < 7; ++cc) Ap[r][cc] = -J[r][cc];
        }
        lu_solve<7, 1>(Ap, dx, 7, 1);
      }
      double alpha = 1.0;
      double ls2 = 0.0;
      for (int k = 0; k < 7; ++k) ls2 = ls2 + dx[k][0] * wv[k];
      ls2 = c * ls2;
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xnew[7];
        for (int k = 0; k < 7; ++k) xnew[k] = x[k] + alpha * dx[k][0];
        formR(P, w, xnew, R, s);
        dotR = 0.0;
        for (int k = 0; k < 7; ++k) dotR = dotR + R[k] * R[k];
        nR = sqrt(dotR);
        const double nRs = 0.5 * dotR;
        if ((nRs <= nlsx) || (ls > mls)) {
          for (int k = 0; k < 7; ++k) x[k] = xnew[k];
          break;
        }
        alpha = red * alpha;
        ls = ls + 1;
      }
      iter = iter + 1;
      if ((iter > miter) || any_nan<NS>(x, 7)) { fail = true; return fail; }
    }
  }
  for (int k = 0; k < 6; ++k) stress[k] = x[k];
  tt = x[6];
  return fail;
}

// mm10_setup (mm10_a.f:852) for the Voce model.  n_R (R at n) and gradFe at n feed only the
// geometric-hardening term; they are read straight from history at n (identity / zero at step 1).
template <int NS>
W3D_HD inline void setup(const Props<NS> &P, Work<NS> &w, bool init, const CHist &hn, const Layout &L) {
  double tr[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tr[i][j] = w.n_Rp[j][i];
  const int ns = P.nslip();
  {
    double wk[3][3];
    rt2rve(tr, w.RE);
    mat33(wk, w.R, tr);
    rt2rvw(wk, w.RQ);
#if !defined(__CUDA_ARCH__) || defined(W3D_MM10_SLIP_STORE)
    for (int i = 0; i < ns; ++i) {
      double a6[6], a3[3];
      mv6(a6, w.RE, P.ms[i]);
      for (int k = 0; k < 6; ++k) w.ms(i, k) = a6[k];
      mv3(a3, w.RQ, P.qs[i]);
      for (int k = 0; k < 3; ++k) w.q(i, k) = a3[k];
    }
#endif
  }
  w.mu_harden = P.stiffness[5][5];   // mm10_setup_voche
  // geometric hardening tau_l from the lattice curvature (Acharya), :925-975
  const double alpha = 1.0 / 3.0;
  const double cst = P.k_0 * P.burgers * alpha * alpha * w.mu_harden * w.mu_harden / 2.0 / P.theta_0;
  if (cst == 0.0) {                  // k_0 = 0: no geometric hardening
    for (int t = 0; t < ns; ++t) w.tau_l[t] = 0.0;
    return;
  }
  double curv[3][3][3], n_R[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) n_R[i][j] = init ? ((i == j) ? 1.0 : 0.0) : hn[L.o_R() + j * 3 + i];
  // gradFeinv(i,j,k) column-major: i + 3 j + 9 k
#define GF(i, j, k) (init ? 0.0 : hn[L.o_gradfe() + (i) + 3 * (j) + 9 * (k)])
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) curv[i][j][k] = GF(i, j, k) - GF(i, k, j);
#undef GF
  for (int t = 0; t < ns; ++t) {
    double wv[3], cn[3], tm[3];
    mv3(wv, tr, P.ns[t]);
    mv3(cn, n_R, wv);
    for (int i = 0; i < 3; ++i)
      tm[i] = 0.0 + 0.5 * (-curv[i][0][1] * cn[2] + curv[i][0][2] * cn[1] + curv[i][1][0] * cn[2] -
                           curv[i][1][2] * cn[0] - curv[i][2][0] * cn[1] + curv[i][2][1] * cn[0]);
    w.tau_l[t] = cst * sqrt(tm[0] * tm[0] + tm[1] * tm[1] + tm[2] * tm[2]);
  }
}

// sums over the crystals of one Gauss point (mm10_a_do_crystal accumulates, mm10_a_crystal_avgs
// divides); the tangent and the slip history are accumulated in history at n+1 directly
struct PointSums {
  double sig[6], t_work, p_work, p_strain;
};

// mm10_store_cryhist for a crystal whose update took the elastic short cut (mm10_solve_strup_a):
// everything mm10_setup_np1 zeroes stays zero
template <int NS>
W3D_HD inline void store_crystal_defaults(const Layout &L, int c, const Hist &h1) {
  for (int k = 0; k < 3; ++k) h1[L.c_euler(c) + k] = 0.0;
  for (int k = 0; k < 9; ++k) h1[L.c_Rp(c) + k] = 0.0;
  for (int k = 0; k < 6; ++k) h1[L.c_eps(c) + k] = 0.0;
  for (int k = 0; k < L.len_slip; ++k) h1[L.c_slip(c) + k] = 0.0;
  for (int k = 0; k < L.len_u; ++k) h1[L.c_u(c) + k] = 0.0;
  for (int k = 0; k < 6; ++k) h1[L.c_ep(c) + k] = 0.0;
  for (int k = 0; k < 6; ++k) h1[L.c_ed(c) + k] = 0.0;
}

// One crystal: mm10_a_do_crystal -> mm10_solve_crystal and everything below it, then
// mm10_store_cryhist: the crystal's record goes straight to history at n+1, its contribution to
// the point averages into acc / history (tangent: column-major at o_D, stored for crystal 0 and
// added for the others).  Returns true when the crystal asks for a global step cut (solver failed
// after 4 halvings).
template <int NS>
W3D_HD inline bool crystal_update(const Props<NS> &P, Work<NS> &w, const double *uddt, double dt,
                                  bool iter_0_extrapolate_off, bool asymmetric, bool init,
                                  const Layout &L, int c, const CHist &hn, const Hist &h1,
                                  PointSums &acc) {
  const int ns = P.nslip();
  // ---- mm10_solve_strup
  for (int k = 0; k < 6; ++k) w.D[k] = uddt[k];
  w.tinc = dt;
  w.dg = effective_dg<NS>(w.D);
  if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;
  setup(P, w, init, hn, L);
  double stress[6];
  for (int k = 0; k < 6; ++k) stress[k] = w.n_stress[k];
  double tt = w.n_tt;
  const double temp1 = uddt[0] * uddt[0] + uddt[1] * uddt[1] + uddt[2] * uddt[2] + uddt[3] * uddt[3] +
                       uddt[4] * uddt[4] + uddt[5] * uddt[5];
  const double temp2 = stress[0] * stress[0] + stress[1] * stress[1] + stress[2] * stress[2] +
                       stress[3] * stress[3] + stress[4] * stress[4] + stress[5] * stress[5];
  const bool no_load = (temp2 == 0.0) && (temp1 == 0.0);
  // per-crystal history entries every path stores the same way
  for (int k = 0; k < 6; ++k) h1[L.c_D(c) + k] = uddt[k];
  for (int k = 0; k < L.len_tt; ++k) h1[L.c_tt(c) + k] = 0.0;
  for (int k = 0; k < L.len_rate; ++k) h1[L.c_rate(c) + k] = 0.0;
  Slips<NS> sf;   // slips of the converged state; the last solve runs with the full-step D, tinc, dg
  if (iter_0_extrapolate_off || no_load) {
    // mm10_solve_strup_a: elastic stiffness as tangent; stress corrected for creep only
    for (int cc = 0; cc < 6; ++cc)
      for (int r = 0; r < 6; ++r) {
        const double t = P.stiffness[r][cc];
        h1[L.o_D() + cc * 6 + r] = (c == 0) ? t : h1[L.o_D() + cc * 6 + r] + t;
      }
    if (!no_load) {
      double R1[6];
      formR1(P, w, stress, tt, sf, R1);
      for (int k = 0; k < 6; ++k) stress[k] = stress[k] - R1[k];
    }
    for (int k = 0; k < 6; ++k) { h1[L.c_stress(c) + k] = stress[k]; acc.sig[k] = acc.sig[k] + stress[k]; }
    h1[L.c_tt(c)] = tt;
    h1[L.c_rate(c)] = 0.0;
    store_crystal_defaults<NS>(L, c, h1);
    return false;
  }
  double J[7][7];
  W3D_STAT(3);
  {  // mm10_solve_strup_iterate: adaptive sub-incrementation, at most 4 halvings
    double ostress[6];
    for (int k = 0; k < 6; ++k) ostress[k] = stress[k];
    double ott = tt;
    double frac = 0.0, step = 1.0;
    int cuts = 0;
    const int mcuts = 4;
    bool fail = false;
    while (frac < 1.0) {
      for (int k = 0; k < 6; ++k) w.D[k] = uddt[k] * (step + frac);
      w.tinc = dt * (step + frac);
      w.dg = effective_dg<NS>(w.D);
      if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;
      tt = w.n_tt;   // "some subroutines modify n%tau_tilde in setup": tt restarts from n
      W3D_STAT(4);
      fail = solve(P, w, stress, tt, dt * step, J, sf);
      if (fail) {
        for (int k = 0; k < 6; ++k) stress[k] = ostress[k];
        tt = ott;
        step = step * 0.5;
        cuts = cuts + 1;
        if (cuts > mcuts) break;
        fail = false;
      } else {
        for (int k = 0; k < 6; ++k) ostress[k] = stress[k];
        ott = tt;
        frac = frac + step;
      }
    }
    if (fail || tt != tt || any_nan<NS>(stress, 6)) return true;
  }
  for (int k = 0; k < 6; ++k) { h1[L.c_stress(c) + k] = stress[k]; acc.sig[k] = acc.sig[k] + stress[k]; }
  h1[L.c_tt(c)] = tt;
  h1[L.c_rate(c)] = w.tt_rate;
  // from here on the full-step quantities of np1 apply (np1%D, np1%tinc, np1%dg)
  for (int k = 0; k < 6; ++k) w.D[k] = uddt[k];
  w.tinc = dt;
  w.dg = effective_dg<NS>(w.D);
  if (P.alter_mode_d != 0.0) w.dg = P.eps_dot_0_y * w.tinc;

  // ---- mm10_tangent (mm10_a.f:664): T = JJ^-1 (C - JA - JB); for Voce ed = 0 and dgdd = 0 so
  //      JA = JB = 0 and JJ = J11 - J12 J22^-1 J21 with the Jacobian of the last solver pass
  {
    double JJ[6][6], JR[6][6];
    double b21[6];
    const double r66 = 1.0 / J[6][6];
    for (int k = 0; k < 6; ++k) b21[k] = J[6][k] * r66;       // DGESV on the 1x1 J22
    for (int r = 0; r < 6; ++r)
      for (int cc = 0; cc < 6; ++cc) JJ[r][cc] = J[r][cc] + (-1.0) * J[r][6] * b21[cc];
    double rinv[6];
    if (lu_factor_nopivot<6>(JJ, rinv)) {             // one column of C at a time
      for (int cc = 0; cc < 6; ++cc) {
        double b[6];
        for (int r = 0; r < 6; ++r) b[r] = P.stiffness[r][cc] - 0.0 - 0.0 * J[r][6];
        lu_apply<6>(JJ, rinv, b);
        for (int r = 0; r < 6; ++r) JR[r][cc] = b[r];
      }
    } else {                                          // an interchange is needed: general routine
      W3D_STAT(5);
      double Jp[6][6];
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) {
          Jp[r][cc] = J[r][cc] + (-1.0) * J[r][6] * b21[cc];
          JR[r][cc] = P.stiffness[r][cc] - 0.0 - 0.0 * J[r][6];
        }
      lu_solve<6, 6>(Jp, JR, 6, 6);
    }
    for (int cc = 0; cc < 6; ++cc)
      for (int r = 0; r < 6; ++r) {
        // mm10_a_make_symm_1 unless asymmetric assembly
        const double t = asymmetric ? JR[r][cc] : (JR[r][cc] + JR[cc][r]) * 0.5;
        h1[L.o_D() + cc * 6 + r] = (c == 0) ? t : h1[L.o_D() + cc * 6 + r] + t;
      }
  }
  // ---- mm10_output (mm10_a.f:3493); the plastic rotation update (mm10_update_rotation) and the
  //      Euler angles that need it come last
  double wb[3];
  {
    // one pass over the slip systems for mm10_form_dbarp (plastic and viscous parts separately),
    // mm10_form_wp (qc) and mm10_form_wbarp (qs = RW qs0, for the rotation update below)
    double ed[6] = {0, 0, 0, 0, 0, 0}, dbarp[6] = {0, 0, 0, 0, 0, 0}, wp[3] = {0, 0, 0};
    const double cv = w.tinc * P.iD_v;
    {
      double tr[3][3], RW[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tr[i][j] = w.n_Rp[j][i];
      rt2rvw(tr, RW);
      wb[0] = wb[1] = wb[2] = 0.0;
      for (int i = 0; i < ns; ++i) {
        double m[6], q[3], qs[3];
        slip_system(P, w, i, m, q);
        mv3(qs, RW, P.qs[i]);
        const double dif = sf.rs[i] * cv;
        const double f = sf.si[i] + dif;
        for (int k = 0; k < 6; ++k) ed[k] = ed[k] + dif * m[k];
        for (int k = 0; k < 6; ++k) dbarp[k] = dbarp[k] + sf.si[i] * m[k];
        for (int k = 0; k < 3; ++k) wp[k] = wp[k] + f * q[k];
        for (int k = 0; k < 3; ++k) wb[k] = wb[k] + f * qs[k];
      }
    }
    // (quotients by tinc, ec_dot and the crystal count are products with reciprocals: many of
    //  the numerators are zero, see HardTerm)
    const double itinc = 1.0 / w.tinc;
    for (int k = 0; k < 6; ++k) h1[L.c_ed(c) + k] = ed[k] * itinc;
    double t1 = ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2];
    double t2 = ed[3] * ed[3] + ed[4] * ed[4] + ed[5] * ed[5];
    double ed_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ed_dot = ed_dot * itinc;
    const double work_inc = stress[0] * w.D[0] + stress[1] * w.D[1] + stress[2] * w.D[2] +
                            stress[3] * w.D[3] + stress[4] * w.D[4] + stress[5] * w.D[5];
    // lattice strain: C^-1 stress (DPOSV -> LU here), rotated with R
    double ee[6];
    {
      double S[6][6], ee0[6][1], erot[6][6];
      for (int r = 0; r < 6; ++r) { ee0[r][0] = stress[r]; for (int cc = 0; cc < 6; ++cc) S[r][cc] = P.stiffness[r][cc]; }
      if (!lu_solve_nopivot<6, 1>(S, ee0)) {
        W3D_STAT(5);
        double Sp[6][6];
        for (int r = 0; r < 6; ++r) { ee0[r][0] = stress[r]; for (int cc = 0; cc < 6; ++cc) Sp[r][cc] = P.stiffness[r][cc]; }
        lu_solve<6, 1>(Sp, ee0, 6, 1);
      }
      rt2rve(w.R, erot);
      double v[6];
      for (int k = 0; k < 6; ++k) v[k] = ee0[k][0];
      mv6(ee, erot, v);
    }
    double ewwe[6], ep[6];
    symsw(ee, wp, ewwe);
    for (int k = 0; k < 6; ++k) { ep[k] = dbarp[k] + ewwe[k]; h1[L.c_ep(c) + k] = ep[k] * itinc; }
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    double ep_dot = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    ep_dot = ep_dot * itinc;
    for (int k = 0; k < 6; ++k) ep[k] = ep[k] + ed[k];
    t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
    t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
    const double p_strain_inc = sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
    const double p_work_inc = stress[0] * ep[0] + stress[1] * ep[1] + stress[2] * ep[2] +
                              stress[3] * ep[3] + stress[4] * ep[4] + stress[5] * ep[5];
    for (int k = 0; k < 6; ++k) h1[L.c_eps(c) + k] = ee[k];
    const double ec_dot = p_strain_inc * itinc;
    double n_eff;
    if (ec_dot > 0.0) {
      for (int k = 0; k < 6; ++k) ep[k] = ep[k] * itinc;
      const double iec = 1.0 / ec_dot;
      n_eff = 0.0;
      const double dgn = w.dg * P.rate_n / tt;
      for (int i = 0; i < ns; ++i) {
        const double rs = sf.rs[i];
        const double dgdt = dgn * sf.pw[i] + cv;
        double m[6], q[3];
        slip_system(P, w, i, m, q);
        t1 = m[0] * ep[0] + m[1] * ep[1] + m[2] * ep[2];
        t2 = m[3] * ep[3] + m[4] * ep[4] + m[5] * ep[5];
        const double ec_slip = t1 + 0.5 * t2;
        t1 = rs * dgdt * iec;
        t2 = ec_slip * iec * itinc;
        n_eff = n_eff + (2.0 / 3.0) * t1 * t2;
      }
    } else {
      n_eff = 1.0e10;
    }
    double s_trace = (stress[0] + stress[1] + stress[2]) / 3.0;
    double se[6];
    for (int k = 0; k < 6; ++k) se[k] = stress[k];
    se[0] -= s_trace; se[1] -= s_trace; se[2] -= s_trace;
    t1 = se[0] * se[0] + se[1] * se[1] + se[2] * se[2];
    t2 = se[3] * se[3] + se[4] * se[4] + se[5] * se[5];
    s_trace = sqrt(1.5 * (t1 + 2.0 * t2));
    double B_eff;
    if (ec_dot < 1.0e-100) B_eff = 0.0;
    else if (n_eff > 100.0) B_eff = -1.0;
    else B_eff = ec_dot / pow(s_trace, n_eff);
    // total slip increments (plastic + viscous), the most active system and the number active
    double maxslip = 0.0;
    int sysID = 0;
    for (int i = 0; i < ns; ++i) {
      const double inc = sf.si[i] + sf.rs[i] * cv;
      h1[L.c_slip(c) + i] = inc;
      h1[L.o_sliphist() + i] = h1[L.o_sliphist() + i] + inc;
      const double cur = fabs(inc);
      if (cur > maxslip) { maxslip = cur; sysID = i + 1; }
    }
    for (int k = ns; k < L.len_slip; ++k) h1[L.c_slip(c) + k] = 0.0;
    int numAct = 0;
    const double thresh = 0.1 * maxslip;
    for (int i = 0; i < ns; ++i)
      if (fabs(sf.si[i] + sf.rs[i] * cv) >= thresh) numAct = numAct + 1;
    // uhist (15 used entries of len_u)
    for (int k = 0; k < L.len_u; ++k) h1[L.c_u(c) + k] = 0.0;
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
    double Rp[3][3];
    double W[3][3], A[3][3];
    W[0][0] = 0.0; W[0][1] = wb[2]; W[0][2] = wb[1];
    W[1][0] = -wb[2]; W[1][1] = 0.0; W[1][2] = wb[0];
    W[2][0] = -wb[1]; W[2][1] = -wb[0]; W[2][2] = 0.0;
    const double alpha = sqrt(W[1][2] * W[1][2] + W[0][2] * W[0][2] + W[0][1] * W[0][1]);
    if (alpha < 1.0e-16) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = 0.0;
    } else {
      double WW[3][3];
      mat33(WW, W, W);
      const double ca = (1.0 - cos(alpha)) / (alpha * alpha), sa = sin(alpha) / alpha;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = ca * WW[i][j] + sa * W[i][j];
    }
    A[0][0] += 1.0; A[1][1] += 1.0; A[2][2] += 1.0;
    mat33(Rp, A, w.n_Rp);
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) h1[L.c_Rp(c) + j * 3 + i] = Rp[i][j];
    const double pi = 3.141592653589793, one_eighty = 180.0;
    double w1[3][3], full[3][3];
    mat33_bt(w1, Rp, w.R);
    mat33(full, P.g, w1);
    double psi = atan2(full[2][1], full[2][0]);
    if (psi < 0.0) psi = psi + 2.0 * pi;
    double phi = atan2(full[1][2], full[0][2]);
    if (phi < 0.0) phi = phi + 2.0 * pi;
    if (full[2][2] > 1.0) full[2][2] = 1.0;
    const double theta = acos(full[2][2]);
    if ((int)P.angle_type_d == 1) {
      h1[L.c_euler(c) + 0] = one_eighty / pi * psi;
      h1[L.c_euler(c) + 1] = one_eighty / pi * theta;
      h1[L.c_euler(c) + 2] = one_eighty / pi * phi;
    } else {
      h1[L.c_euler(c) + 0] = psi; h1[L.c_euler(c) + 1] = theta; h1[L.c_euler(c) + 2] = phi;
    }
  }
  return false;
}

// mm10 (mm10_a.f:29): all crystals of one Gauss point.  hn / h1: history at n / n+1;
// props: ncrystals consecutive Props records of this element; Rur: rot_blk_n1 (R of F = R U at
// n+1, identity for small strain), column-major 9; uddt: un-rotated strain increment;
// cgn / cgn1: urcs at n / n+1 (9); sl / sls: slip_store_doubles<NS>() doubles of scratch with
// element stride sls (Work).  Returns true when a crystal requests a step cut.
template <int NS>
W3D_HD inline bool gauss_point_update(int step, int ncrystals, const Props<NS> *props, const Layout &L,
                                      const CHist &hn, const Hist &h1, const double *Rur,
                                      const double *uddt, double dt, bool iter_0_extrapolate_off,
                                      bool asymmetric, const double *cgn, double *cgn1, double *sl,
                                      int sls) {
  // At step 1 the reference initialises its block-local copy of history at n
  // (mm10_init_general_hist / _uout_hist / _slip_hist, mm10_init_cc_hist0) and never writes that
  // copy back; the same values are substituted here on the fly, history at n stays untouched.
  const bool init = (step == 1);
  PointSums acc;
  for (int k = 0; k < 6; ++k) acc.sig[k] = 0.0;
  acc.t_work = 0.0; acc.p_work = 0.0; acc.p_strain = 0.0;
  // the slip-history sum is accumulated directly into history at n+1 (zeroed first)
  for (int k = 0; k < L.len_slip; ++k) h1[L.o_sliphist() + k] = 0.0;

  Work<NS> w;
  w.sl = sl;
  w.sls = sls;
  for (int c = 0; c < ncrystals; ++c) {
    const Props<NS> &P = props[c];
    // mm10_copy_cc_hist
    if (init) {
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.n_Rp[i][j] = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < 6; ++k) { w.n_stress[k] = cgn[k]; w.n_D[k] = 0.0; }
      w.n_tt = P.tau_y + 1.0e-5;   // mm10_init_voche
      w.n_tt_rate = 0.0;
    } else {
      for (int k = 0; k < 6; ++k) w.n_stress[k] = hn[L.c_stress(c) + k];
      for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.n_Rp[i][j] = hn[L.c_Rp(c) + j * 3 + i];
      for (int k = 0; k < 6; ++k) w.n_D[k] = hn[L.c_D(c) + k];
      w.n_tt = hn[L.c_tt(c)];
      w.n_tt_rate = hn[L.c_rate(c)];
    }
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) w.R[i][j] = Rur[j * 3 + i];
    w.tt_rate = 0.0;
    if (crystal_update(P, w, uddt, dt, iter_0_extrapolate_off, asymmetric, init, L, c, hn, h1, acc))
      return true;
  }
  // mm10_a_crystal_avgs + mm10_a_store_crystal
  const double rn = (double)ncrystals;
  const double irn = 1.0 / rn;
  for (int k = 0; k < 6; ++k) cgn1[k] = acc.sig[k] * irn;
  if (ncrystals > 1)
    for (int k = 0; k < 36; ++k) h1[L.o_D() + k] = h1[L.o_D() + k] * irn;
  for (int k = 0; k < 9; ++k) h1[L.o_R() + k] = Rur[k];
  for (int k = 0; k < L.len_slip; ++k)
    h1[L.o_sliphist() + k] = (init ? 0.0 : hn[L.o_sliphist() + k]) + h1[L.o_sliphist() + k] * irn;
  const double t_work_inc = acc.t_work * irn, p_work_inc = acc.p_work * irn, p_strain_inc = acc.p_strain * irn;
  cgn1[6] = cgn[6] + t_work_inc;
  cgn1[7] = cgn[7] + p_work_inc;
  cgn1[8] = cgn[8] + p_strain_inc;
  h1[L.o_work() + 0] = (init ? 0.0 : hn[L.o_work() + 0]) + t_work_inc;
  h1[L.o_work() + 1] = (init ? 0.0 : hn[L.o_work() + 1]) + p_work_inc;
  h1[L.o_work() + 2] = (init ? 0.0 : hn[L.o_work() + 2]) + p_strain_inc;
  return false;
}

// mm10_calc_grads (mm10_a.f:1579) for the 2x2x2 rule: least-squares linear fit of
// Rp^T R (geonl) / Rp (small strain) over the 8 Gauss points; the DGELS solution of the fixed
// 8x4 system [xi eta zeta 1] has the closed form  grad_k = sum_i x_ik f_i / (8 l^2).
// rt[gp][9]: the fitted tensors column-major; jacinv: inverse of the reference's local_work%jac
// (the last Jacobian formed, undeformed configuration, Gauss point 8); out: grads(3,3,3) col-major.
W3D_HD inline void calc_grads(const double (*rt)[9], bool geonl, const double (&jacinv)[3][3],
                              double *out) {
  const double l = 0.5773502691896257645;
  const int sx[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  const int sy[8] = {-1, 1, -1, 1, -1, 1, -1, 1};
  const int sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a) {
      double g0 = 0.0, g1 = 0.0, g2 = 0.0;
      for (int i = 0; i < 8; ++i) {
        const double f = rt[i][b * 3 + a];
        g0 += sx[i] * f; g1 += sy[i] * f; g2 += sz[i] * f;
      }
      const double s = 1.0 / (8.0 * l);
      g0 *= s; g1 *= s; g2 *= s;
      double r0 = g0, r1 = g1, r2 = g2;
      if (geonl) {
        r0 = jacinv[0][0] * g0 + jacinv[0][1] * g1 + jacinv[0][2] * g2;
        r1 = jacinv[1][0] * g0 + jacinv[1][1] * g1 + jacinv[1][2] * g2;
        r2 = jacinv[2][0] * g0 + jacinv[2][1] * g1 + jacinv[2][2] * g2;
      }
      out[a + 3 * b + 9 * 0] = r0;
      out[a + 3 * b + 9 * 1] = r1;
      out[a + 3 * b + 9 * 2] = r2;
    }
}

// ---- host-side property setup (setup_mm10_rknstr, rknstr.f:1568-2027) -------------------------
// crystal-type record = the c_array(cnum) fields the path reads (mod_crystals.f:148-236)
struct CrystalType {
  double elast_stiff[6][6];        // row-major
  double bi[48][3], ni[48][3];     // slip directions / normals in crystal axes
  double harden_n, tau_y, tau_v, voche_m, theta_o, k_o, b, iD_v, eps_dot_o_y;
  double atol, atol1, rtol, rtol1;
  int nslip, num_hard, h_type, tang_calc, miter, solver, alter_mode;
};

// mm10_rotation_matrix (mm10_a.f:1310), Kocks convention
inline void rotation_matrix(const double *angles, int angle_type /*1 degrees, 2 radians*/,
                            double (&r)[3][3]) {
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

template <int NS>
inline void setup_props(const CrystalType &ct, const double *angles, int angle_type, Props<NS> &P) {
  P.rate_n = ct.harden_n; P.tau_y = ct.tau_y; P.tau_v = ct.tau_v; P.voche_m = ct.voche_m;
  P.theta_0 = ct.theta_o; P.k_0 = ct.k_o; P.burgers = ct.b; P.iD_v = ct.iD_v;
  P.eps_dot_0_y = ct.eps_dot_o_y;
  P.atol = ct.atol; P.atol1 = ct.atol1; P.rtol = ct.rtol; P.rtol1 = ct.rtol1;
  for (int k = 0; k < 3; ++k) P.init_angles[k] = angles[k];
  P.nslip_d = ct.nslip; P.miter_d = ct.miter; P.alter_mode_d = ct.alter_mode; P.angle_type_d = angle_type;
  rotation_matrix(angles, angle_type, P.g);
  double tr[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tr[i][j] = P.g[j][i];
  for (int s = 0; s < NS; ++s) {
    for (int k = 0; k < 6; ++k) P.ms[s][k] = 0.0;
    for (int k = 0; k < 3; ++k) { P.qs[s][k] = 0.0; P.ns[s][k] = 0.0; }
  }
  for (int s = 0; s < ct.nslip && s < NS; ++s) {
    double bs[3], nn[3], A[3][3];
    mv3(bs, tr, ct.bi[s]);
    mv3(nn, tr, ct.ni[s]);
    for (int k = 0; k < 3; ++k) P.ns[s][k] = nn[k];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = bs[i] * nn[j];
    double As[3][3], Aa[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { As[i][j] = 0.5 * (A[i][j] + A[j][i]); Aa[i][j] = 0.5 * (A[i][j] - A[j][i]); }
    P.ms[s][0] = As[0][0]; P.ms[s][1] = As[1][1]; P.ms[s][2] = As[2][2];       // mm10_ET2EV
    P.ms[s][3] = 2.0 * As[0][1]; P.ms[s][4] = 2.0 * As[1][2]; P.ms[s][5] = 2.0 * As[0][2];
    P.qs[s][0] = Aa[1][2]; P.qs[s][1] = Aa[0][2]; P.qs[s][2] = Aa[0][1];       // mm10_WT2WV
  }
  // rotate the stiffness: C = Rstiff C0 Rstiff^T with Rstiff = RT2RVE(g^T)
  double Rs[6][6], t66[6][6];
  rt2rve(tr, Rs);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += ct.elast_stiff[i][k] * Rs[j][k];
    t66[i][j] = s;
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += Rs[i][k] * t66[k][j];
    P.stiffness[i][j] = s;
  }
}

}  // namespace w3d_mm10
