// Test harness (tests/ only): one Gauss-point update of mm10 through
//   (a) the independent statement-for-statement restatement  oracle/w3d_oracle_mm10.h
//   (b) the engine's fast lattice-frame core                 warp3d_b200/csrc/w3d_mm10_fast.h (host build)
// on the same inputs.  Built on demand by tests/test_mm10_fast_core.py with g++.
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../oracle/w3d_oracle_mm10.h"
#include "../../warp3d_b200/csrc/w3d_mm10_fast.h"

extern "C" int mm10_point_both(const double *ct /*344*/, int ncrystals, const double *angles, int angle_type,
                               int step, const double *hist_n, const double *Rur, const double *uddt, double dt,
                               int iter0off, int asym, const double *cgn, int taul, double *h1_lit,
                               double *cgn1_lit, double *h1_fast, double *cgn1_fast, int *cuts /*2*/) {
  const int nslip = (int)ct[337];
  if (nslip > 12) return -1;
  // ---- literal
  {
    using namespace w3do_mm10;
    CrystalType c;
    const size_t nd = 36 + 144 + 144 + 13;
    std::memcpy(&c, ct, nd * sizeof(double));
    const double *ti = ct + nd;
    c.nslip = (int)ti[0]; c.num_hard = (int)ti[1]; c.h_type = (int)ti[2]; c.tang_calc = (int)ti[3];
    c.miter = (int)ti[4]; c.solver = (int)ti[5]; c.alter_mode = (int)ti[6];
    std::vector<Props<12>> P(ncrystals);
    for (int k = 0; k < ncrystals; ++k) { std::memset(&P[k], 0, sizeof(Props<12>)); setup_props<12>(c, angles + 3 * k, angle_type, P[k]); }
    const Layout L = Layout::make(c.nslip, c.num_hard);
    const CHist hn{hist_n, 1};
    const Hist h1{h1_lit, 1};
    cuts[0] = gauss_point_update<12>(step, ncrystals, P.data(), L, hn, h1, Rur, uddt, dt, iter0off != 0, asym != 0, cgn, cgn1_lit) ? 1 : 0;
  }
  // ---- fast
  {
    using namespace w3d_mm10f;
    static CrystalConsts K;
    setup_consts(ct, K);
    K.angle_type = angle_type;
    g_host_ct = &K;
    std::vector<FProps> P(ncrystals);
    for (int k = 0; k < ncrystals; ++k)
      if (!setup_fprops(ct, angles + 3 * k, angle_type, P[k])) return -2;
    const LayoutF L = LayoutF::make(nslip);
    const CHistF hn{hist_n, 1};
    const HistF h1{h1_fast, 1};
    double ws[WORK_DOUBLES + NSF];
    const WorkF W{ws, 1};
    const PointInF in{Rur, uddt, cgn, 1};
    // as launch_mm10_fast picks the k_mm10f instantiation: one crystal per element -> tangent stored
    // ahead by the Newton passes; rate exponent 19 compiled in
    auto run = [&](auto taul_c, auto single_c, auto rate_c) {
      return gauss_point_update_fast<decltype(taul_c)::value, NoSync, decltype(single_c)::value, decltype(rate_c)::value>(
          step, ncrystals, P.data(), L, hn, h1, W, in, dt, iter0off != 0, asym != 0, cgn1_fast);
    };
    using T = std::true_type;
    using F = std::false_type;
    using RT = std::integral_constant<int, -1>;
    using E19 = std::integral_constant<int, 19>;
    const bool single = (ncrystals == 1);
    const bool e19 = (g_host_ct->rate_ie == 19);
    const bool cut = taul ? (single ? (e19 ? run(T{}, T{}, E19{}) : run(T{}, T{}, RT{})) : run(T{}, F{}, RT{}))
                          : (single ? (e19 ? run(F{}, T{}, E19{}) : run(F{}, T{}, RT{})) : run(F{}, F{}, RT{}));
    cuts[1] = cut ? 1 : 0;
  }
  return 0;
}
