// Test harness (tests/ only): one element block at one Gauss point through
//   (a) the independent restatement of mm03.f        oracle/w3d_oracle_mm03.h  (block routines)
//   (b) the engine's per-point core                  warp3d_b200/csrc/w3d_mm03_core.h (host build)
// on the same inputs.  Built on demand by tests/test_mm03_core.py with g++.
#include <vector>

#include "../../oracle/w3d_oracle_mm03.h"
#include "../../warp3d_b200/csrc/w3d_mm03_core.h"

// props: e, nu, sigyld (unreduced), h_fixed, n_power, m_power, eps_ref per element (7 doubles);
// arrays are [span][...] row-major.  rc[i]: 0 ok, 1 no convergence (per implementation).
extern "C" int mm03_block_both(int span, int step, int iter, double dt, const double *props, const double *hist_n,
                               const double *cgn, const double *uddt, double *h1_lit, double *cgn1_lit,
                               double *cep_lit, int *rc_lit, double *h1_core, double *cgn1_core, double *cep_core,
                               int *rc_core) {
  // ---- literal, block structured as drive_03_update_a does it (rstgp1.f:1898-1978)
  {
    using namespace w3do_mm03;
    std::vector<Props> P(span);
    std::vector<Point> pt(span);
    for (int i = 0; i < span; ++i) {
      const double *p = props + 7 * i;
      P[i].e = p[0]; P[i].nu = p[1]; P[i].sigyld = p[2]; P[i].h_fixed = p[3]; P[i].n_power = p[4];
      P[i].m_power = p[5]; P[i].eps_ref = p[6]; P[i].f0 = 0.0; P[i].q1 = P[i].q2 = P[i].q3 = 0.0;
      P[i].nucleation = false;
      if (P[i].n_power > 0.0) P[i].sigyld = P[i].sigyld * 0.95;
      for (int k = 0; k < 6; ++k) { pt[i].deps[k] = uddt[6 * i + k]; pt[i].stress_n1_elas[k] = 0.0; }
      for (int k = 0; k < 9; ++k) { pt[i].stress_n[k] = cgn[9 * i + k]; pt[i].stress_n1[k] = 0.0; }
      for (int k = 0; k < 15; ++k) { pt[i].history[k] = hist_n[15 * i + k]; pt[i].history1[k] = 0.0; }
      pt[i].null_point = pt[i].nonlinear_flag = false;
      pt[i].p_trial = pt[i].q_trial = pt[i].yld_func = 0.0;
      rc_lit[i] = 0;
    }
    if (mm03p(step, span, pt.data(), P.data()))
      for (int i = 0; i < span; ++i) {
        if (pt[i].null_point || !pt[i].nonlinear_flag) continue;
        rc_lit[i] = mm03(P[i], dt, pt[i]);
      }
    for (int i = 0; i < span; ++i) {
      double D[6][6];
      cnst3(iter, P[i], pt[i].stress_n1_elas, pt[i].history, pt[i].history1, D);
      int k = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= r; ++c) cep_lit[21 * i + k++] = D[r][c];
      for (int q = 0; q < 9; ++q) cgn1_lit[9 * i + q] = pt[i].stress_n1[q];
      for (int q = 0; q < 15; ++q) h1_lit[15 * i + q] = pt[i].history1[q];
    }
  }
  // ---- engine core, point by point; the block-level decision is the caller's (k_mm03 does it per CTA)
  {
    using namespace w3d_mm03;
    bool any_active = false;
    for (int i = 0; i < span; ++i) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += std::fabs(uddt[6 * i + k]);
      if (!(s <= 1.0e-10 * 0.002)) any_active = true;
    }
    for (int i = 0; i < span; ++i) {
      const double *p = props + 7 * i;
      Props P;
      P.e = p[0]; P.nu = p[1]; P.sigyld = p[4] > 0.0 ? p[2] * 0.95 : p[2]; P.h_fixed = p[3]; P.n_power = p[4];
      P.m_power = p[5]; P.eps_ref = p[6]; P.q1 = P.q2 = P.q3 = 0.0;
      double D[6][6];
      rc_core[i] = update(step, iter, P, dt, !any_active, cgn + 9 * i, uddt + 6 * i, hist_n + 15 * i,
                          cgn1_core + 9 * i, h1_core + 15 * i, D);
      int k = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= r; ++c) cep_core[21 * i + k++] = D[r][c];
    }
  }
  return 0;
}
