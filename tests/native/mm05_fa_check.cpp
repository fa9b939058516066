// Test harness (tests/ only): one element block at one Gauss point, either option of mm05 (sign of
// mm_props(4)), through the two restatements kept under oracle/:
//   (a) w3d_oracle_mm05.h     per point, the transcription the engine's device core descends from
//   (b) w3d_oracle_mm05_fa.h / _gp.h  block form, written a second time directly from mm05.f (round 2)
// Built on demand by tests/test_mm05_fa_second_restatement.py with g++.
#include <vector>

#include "../../oracle/w3d_oracle_mm05.h"
#include "../../oracle/w3d_oracle_mm05_fa.h"
#include "../../oracle/w3d_oracle_mm05_gp.h"

// props per element: e, nu, sigyld, mm_props(1:10); arrays [span][...] row-major.
extern "C" int mm05_fa_both(int span, int step, int iter, int adaptive, const double *props, const double *hist_n,
                            const double *cgn, const double *uddt, double *h1_a, double *cgn1_a, double *d_a,
                            double *trial_a, int *rc_a, double *h1_b, double *cgn1_b, double *d_b, double *trial_b,
                            int *rc_b) {
  {
    using namespace w3do_mm05;
    for (int i = 0; i < span; ++i) {
      const double *p = props + 13 * i;
      Props P;
      P.e = p[0]; P.nu = p[1]; P.sigyld = p[2];
      for (int k = 0; k < 10; ++k) P.mm[k] = p[3 + k];
      double h[14], h1[14], c0[9], c1[9], ud[6], tr[6], D[6][6];
      for (int k = 0; k < 14; ++k) { h[k] = hist_n[14 * i + k]; h1[k] = 0.0; }
      for (int k = 0; k < 9; ++k) { c0[k] = cgn[9 * i + k]; c1[k] = 0.0; }
      for (int k = 0; k < 6; ++k) { ud[k] = uddt[6 * i + k]; tr[k] = 0.0; }
      rc_a[i] = update(step, iter, P, c0, c1, ud, h, h1, tr, adaptive != 0);
      if (rc_a[i] == OK) cnst5(P, tr, h1, c1, D);
      for (int k = 0; k < 14; ++k) h1_a[14 * i + k] = h1[k];
      for (int k = 0; k < 9; ++k) cgn1_a[9 * i + k] = c1[k];
      for (int k = 0; k < 6; ++k) trial_a[6 * i + k] = tr[k];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) d_a[36 * i + 6 * r + c] = rc_a[i] == OK ? D[r][c] : 0.0;
    }
  }
  {
    namespace fa = w3do_mm05_fa;
    std::vector<double> mmp((size_t)span * 10), e(span), nu(span), sy(span), sn(cgn, cgn + (size_t)span * 9),
        dp(uddt, uddt + (size_t)span * 6), hn(hist_n, hist_n + (size_t)span * 14);
    for (int i = 0; i < span; ++i) {
      const double *p = props + 13 * i;
      e[i] = p[0]; nu[i] = p[1]; sy[i] = p[2];
      for (int k = 0; k < 10; ++k) mmp[(size_t)i * 10 + k] = p[3 + k];
      for (int k = 0; k < 14; ++k) h1_b[14 * i + k] = 0.0;
      for (int k = 0; k < 9; ++k) cgn1_b[9 * i + k] = 0.0;
    }
    fa::Block B;
    B.span = span;
    B.mm_props = reinterpret_cast<double (*)[10]>(mmp.data());
    B.e_vec = e.data(); B.nu_vec = nu.data(); B.sigyld_vec = sy.data();
    B.trial = reinterpret_cast<double (*)[6]>(trial_b);
    B.stress_n = reinterpret_cast<double (*)[9]>(sn.data());
    B.stress_np1 = reinterpret_cast<double (*)[9]>(cgn1_b);
    B.deps = reinterpret_cast<double (*)[6]>(dp.data());
    B.history_n = reinterpret_cast<double (*)[14]>(hn.data());
    B.history_np1 = reinterpret_cast<double (*)[14]>(h1_b);
    // option by the sign of mm_props(4) = matprp(58); generalized plasticity: h_u = matprp(55),
    // tau = matprp(56), beta_u = matprp(57), delta_u = matprp(59) (rknstr.f:1282-1293)
    namespace gp = w3do_mm05_gp;
    std::vector<double> hu(span), tau(span), bu(span), du(span);
    for (int i = 0; i < span; ++i) {
      hu[i] = mmp[(size_t)i * 10 + 0]; tau[i] = mmp[(size_t)i * 10 + 1];
      bu[i] = mmp[(size_t)i * 10 + 2]; du[i] = mmp[(size_t)i * 10 + 4];
    }
    gp::GpVectors V{e.data(), e.data(), nu.data(), nu.data(), sy.data(), sy.data(), hu.data(), hu.data(),
                    bu.data(), bu.data(), du.data(), du.data(), tau.data()};
    const bool generalized = mmp[3] < 0.0;
    const int rc = generalized ? gp::mm05_gp(step, iter, B, V, adaptive != 0) : fa::mm05_fa(step, iter, B, adaptive != 0);
    for (int i = 0; i < span; ++i) rc_b[i] = rc;
    if (rc == fa::OK) {
      if (generalized) gp::cnst5_gp(B, V, reinterpret_cast<double (*)[6][6]>(d_b));
      else fa::cnst5_fa(B, reinterpret_cast<double (*)[6][6]>(d_b));
    }
  }
  return 0;
}
