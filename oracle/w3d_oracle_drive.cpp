/*
 * w3d_oracle_drive.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): block drivers for the
 * strain / stress / internal-force pass.  Follows reference src/:
 *   drive_eps_sig_internal_forces.f:20-387 (driver), :399-632 (do_nleps_block),
 *   :1653-1940 (dupstr_blocked), :1955-2000 (recstr_gastr), dupstr.f:18-88, rknstr.f:15-184,
 *   :455-547 (rknstr_geonl), :563-614 (rknstr_sm_displ), :1018-1066 (setup_mm01_rknstr),
 *   rstgp1.f:19-332 (rstgp1), :351-598 (rstgp2), :614-1034 (drive_01_update),
 *   :4983-5005, :5168-5241, :5360-5427, gtmat1.f:17-190 (gtmat1), rplstr.f:15-93,
 *   rknifv.f:15-215, gpifv1.f:16-165, addifv.f:86-113, update.f:75-93, mnralg.f:1258-1420.
 * PARITY STATUS: see w3d_oracle.h (mm01, mm05 and mm10 pinned end to end by reference known
 * answers at 1e-5; unpinned at function granularity).
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "w3d_oracle_internal.h"
#include "w3d_oracle_mm10.h"   // statement-for-statement mm10 restatement (independent of the product)
#ifdef W3D_MM10_STATS
extern "C" { long long w3d_mm10_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0}; }
#endif
#include "w3d_oracle_mm05.h"   // per-point mm05 restatement (independent of the product)
#include "w3d_oracle_mm05_fa.h"   // nonlinear_hardening option of mm05: second, block-form restatement
#include "w3d_oracle_mm05_gp.h"   // generalized_plasticity option: the same
#include "w3d_oracle_mm03.h"   // mm03 (mises) restatement (independent of the product)

namespace w3do {

// block-local work arrays == reference "local_work" (include_sig_up:121-272), hex8-sized where the
// extent is not layout-visible.
struct LocalWork {
  std::vector<double> ce_0, ce_n, ce_mid, ce_n1, ue, due, uenh, uen1;      // (mxvl, mxedof)
  std::vector<double> det_j, det_j_mid;                                     // (mxvl, mxgp)
  std::vector<double> nxi, neta, nzeta;                                     // (mxndel, mxgp)
  std::vector<double> gama, gama_mid;                                       // (mxvl,3,3,mxgp)
  std::vector<double> jac;                                                  // (mxvl,3,3)
  std::vector<double> fn, fn1, dfn1;                                        // (mxvl,3,3), (mxvl)
  std::vector<double> vol_block, volume_block;                              // (mxvl,8,3), (mxvl)
  std::vector<double> b;                                                    // (mxvl,mxedof,nstr)
  std::vector<double> urcs_blk_n, urcs_blk_n1;                              // (mxvl,nstrs,mxgp)
  std::vector<double> rot_blk_n1;                                           // (mxvl,9,mxgp)
  std::vector<double> rtse;                                                 // (mxvl,nstr,mxgp)
  std::vector<double> ddtse, strain_n;                                      // (mxvl,nstr,mxgp)
  std::vector<double> elem_hist, elem_hist1;                                // (span,hist,ngp)
  double weights[mxgp];
  double e_vec[mxvl], e_vec_n[mxvl], nu_vec[mxvl], nu_vec_n[mxvl], beta_vec[mxvl], h_vec[mxvl],
      tan_e_vec[mxvl], sigyld_vec[mxvl];
  bool lnelas_vec[mxvl];
  LocalWork()
      : ce_0(mxvl * mxedof), ce_n(mxvl * mxedof), ce_mid(mxvl * mxedof), ce_n1(mxvl * mxedof),
        ue(mxvl * mxedof), due(mxvl * mxedof), uenh(mxvl * mxedof), uen1(mxvl * mxedof),
        det_j(mxvl * mxgp), det_j_mid(mxvl * mxgp), nxi(mxndel * mxgp), neta(mxndel * mxgp),
        nzeta(mxndel * mxgp), gama(mxvl * 9 * mxgp), gama_mid(mxvl * 9 * mxgp), jac(mxvl * 9),
        fn(mxvl * 9), fn1(mxvl * 9), dfn1(mxvl), vol_block(mxvl * 24), volume_block(mxvl),
        b(mxvl * mxedof * nstr), urcs_blk_n(mxvl * nstrs * mxgp), urcs_blk_n1(mxvl * nstrs * mxgp),
        rot_blk_n1(mxvl * 9 * mxgp), rtse(mxvl * nstr * mxgp), ddtse(mxvl * nstr * mxgp),
        strain_n(mxvl * nstr * mxgp) {}
};

// recstr_gastr (drive_eps_sig_internal_forces.f:1955-2000): AoS (nprm,ngp,span) -> (mxvl,nprm,ngp)
static void gastr(double *mlocal, const double *mglobal, int ngp, int nprm, int span) {
  for (int k = 0; k < ngp; ++k)
    for (int j = 0; j < nprm; ++j)
      for (int i = 0; i < span; ++i)
        mlocal[((size_t)k * nprm + j) * mxvl + i] = mglobal[((size_t)i * ngp + k) * nprm + j];
}
// rp_scstr (rplstr.f:173-236): (mxvl,nprm,ngp) -> AoS (nprm,ngp,span)
static void scstr(const double *mlocal, double *mglobal, int ngp, int nprm, int span) {
  for (int k = 0; k < ngp; ++k)
    for (int j = 0; j < nprm; ++j)
      for (int i = 0; i < span; ++i)
        mglobal[((size_t)i * ngp + k) * nprm + j] = mlocal[((size_t)k * nprm + j) * mxvl + i];
}
// dptstf_copy_history / rp_scstr_history: global (hist,ngp,span) <-> local (span,hist,ngp)
static void copy_history_in(double *local, const double *global, int ngp, int hist, int span) {
  for (int k = 0; k < ngp; ++k)
    for (int j = 0; j < hist; ++j)
      for (int i = 0; i < span; ++i)
        local[((size_t)k * hist + j) * span + i] = global[((size_t)i * ngp + k) * hist + j];
}
static void copy_history_out(const double *local, double *global, int ngp, int hist, int span) {
  for (int k = 0; k < ngp; ++k)
    for (int j = 0; j < hist; ++j)
      for (int i = 0; i < span; ++i)
        global[((size_t)i * ngp + k) * hist + j] = local[((size_t)k * hist + j) * span + i];
}

// per-block pointers for every block, built once with running offsets (blocks are concatenated
// in block order in every array)
std::vector<BlockPtrs> all_block_ptrs(const w3do_model *m, const w3do_state *s) {
  std::vector<BlockPtrs> v(m->nelblk);
  int64_t oi = 0, oh = 0, oe = 0, ou = 0, orr = 0, oc = 0, ok = 0, ov = 0;
  for (int blk = 0; blk < m->nelblk; ++blk) {
    const BlockInfo b = block_info(m, blk);
    BlockPtrs &p = v[blk];
    p.edest = m->edest + oi;
    p.cdest = m->cdest + oi;
    p.history = s->history ? s->history + oh : nullptr;
    p.history1 = s->history1 ? s->history1 + oh : nullptr;
    p.eps_n = s->eps_n + oe;
    p.eps_n1 = s->eps_n1 + oe;
    p.urcs_n = s->urcs_n + ou;
    p.urcs_n1 = s->urcs_n1 + ou;
    p.rot_n = s->rot_n + orr;
    p.rot_n1 = s->rot_n1 + orr;
    p.cep = s->cep + oc;
    p.estiff = s->estiff ? s->estiff + ok : nullptr;
    p.element_vol = s->element_vol ? s->element_vol + ov : nullptr;
    oi += (int64_t)b.totdof * b.span;
    oh += (int64_t)b.hist_size * b.ngp * b.span;
    oe += (int64_t)nstr * b.ngp * b.span;
    ou += (int64_t)nstrs * b.ngp * b.span;
    orr += (int64_t)9 * b.ngp * b.span;
    oc += (int64_t)21 * b.span * b.ngp;
    ok += (int64_t)b.utsz * b.span;
    ov += b.span;
  }
  return v;
}

// rstgp1_store_cep (rstgp1.f:5360-5397): 21 lower-triangle terms by row, global (21,span,ngp)
static void store_cep(int span, int gpn, double *gbl /*(21,span,ngp)*/, const double *cep) {
  static const int rr[21] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6};
  static const int cc[21] = {1, 1, 2, 1, 2, 3, 1, 2, 3, 4, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 6};
  for (int i = 0; i < span; ++i)
    for (int k = 0; k < 21; ++k)
      gbl[((size_t)(gpn - 1) * span + i) * 21 + k] = Q6(cep, i, rr[k], cc[k]);
}

// gtmat1.f:17-190 (gtmat1), paths used by mm01/mm05 (compute_f_n / f_bar only for umat/cp)
static void gtmat1(const BlockInfo &bi, LocalWork &lw, int gpn, double *qnhalf, int &error) {
  const int span = bi.span, nnode = bi.nnode;
  std::vector<double> rnh(mxvl * 9, 0.0), fnh(mxvl * 9, 0.0), theta(mxvl * 9, 0.0);
  double dfh[mxvl];
  const double *nxi = &lw.nxi[(gpn - 1) * mxndel], *neta = &lw.neta[(gpn - 1) * mxndel],
               *nzeta = &lw.nzeta[(gpn - 1) * mxndel];
  double *gama = &lw.gama[(size_t)(gpn - 1) * 9 * mxvl];
  jacob1(span, lw.jac.data(), &lw.det_j[(gpn - 1) * mxvl], gama, nxi, neta, nzeta, lw.ce_0.data(),
         nnode);
  tcomp1(span, theta.data(), nxi, neta, nzeta, gama, lw.uen1.data(), nnode);
  fcomp1(span, lw.fn1.data(), lw.dfn1.data(), theta.data(), error);
  if (error == 1) return;
  rtcmp1(span, lw.fn1.data(), &lw.rot_blk_n1[(size_t)(gpn - 1) * 9 * mxvl]);
  tcomp1(span, theta.data(), nxi, neta, nzeta, gama, lw.uenh.data(), nnode);
  fcomp1(span, fnh.data(), dfh, theta.data(), error);
  if (error == 1) return;
  rtcmp1(span, fnh.data(), rnh.data());
  getrm1(span, qnhalf, rnh.data(), 1);
}

// rstgp1_b (rstgp1.f:5189-5241)
static void rstgp1_b(int span, double &internal_energy, double &plastic_work,
                     const double *gp_energies, const double *gp_plast_work, const double *det_j,
                     const double *dfn1, int itype, double beta_fact, double weight) {
  if (itype == 1) {
    for (int i = 0; i < span; ++i) {
      const double factor = dfn1[i] * det_j[i] * beta_fact * weight;
      internal_energy = internal_energy + gp_energies[i] * factor;
      plastic_work = plastic_work + gp_plast_work[i] * factor;
    }
  } else {
    for (int i = 0; i < span; ++i) {
      const double factor = det_j[i] * beta_fact * weight;
      internal_energy = internal_energy + gp_energies[i] * factor;
      plastic_work = plastic_work + gp_plast_work[i] * factor;
    }
  }
}

// drive_01_update (rstgp1.f:614-984) with drive_01_update_a/_b (942-1034)
static void drive_01_update(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn,
                            int step, int iter, const double *uddt_displ, double *gbl_cep,
                            int *fatal) {
  const int span = bi.span;
  std::vector<double> uddt(mxvl * nstr), cep(mxvl * 36, 0.0);
  double gp_dtemps[mxvl];
  for (int i = 0; i < mxvl; ++i) gp_dtemps[i] = 0.0;
  // uddt = uddt_displ + uddt_temps (temperatures off -> uddt_temps = 0)
  for (int k = 0; k < mxvl * nstr; ++k) uddt[k] = uddt_displ[k] + 0.0;
  const bool nonlinear_update = iter >= 1 || m->extrapolated_du != 0;
  double *urcs_n = &lw.urcs_blk_n[(size_t)(gpn - 1) * nstrs * mxvl];
  double *urcs_n1 = &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl];
  double *hist = &lw.elem_hist[(size_t)(gpn - 1) * bi.hist_size * span];
  double *hist1 = &lw.elem_hist1[(size_t)(gpn - 1) * bi.hist_size * span];
  double *rtse = &lw.rtse[(size_t)(gpn - 1) * nstr * mxvl];
  if (nonlinear_update) {
    mm01(span, step, lw.e_vec, lw.nu_vec, lw.beta_vec, lw.h_vec, lw.lnelas_vec, lw.sigyld_vec,
         urcs_n, urcs_n1, uddt.data(), hist, hist1, rtse, gp_dtemps, lw.e_vec_n, lw.nu_vec_n,
         fatal);
    cnst1(span, cep.data(), rtse, lw.nu_vec, lw.e_vec, &hist1[(2 - 1) * span],
          &hist1[(5 - 1) * span], lw.beta_vec, &hist1[(1 - 1) * span], &hist1[(4 - 1) * span]);
  } else {
    const double one = 1.0, two = 2.0;
    for (int i = 0; i < span; ++i) {
      const double e = lw.e_vec[i], nu = lw.nu_vec[i];
      const double c1 = (e / ((one + nu) * (one - two * nu)));
      const double c2 = (one - nu) * c1;
      const double c3 = ((one - two * nu) / two) * c1;
      const double c4 = nu * c1;
      Q6(cep, i, 1, 1) = c2; Q6(cep, i, 2, 2) = c2; Q6(cep, i, 3, 3) = c2;
      Q6(cep, i, 4, 4) = c3; Q6(cep, i, 5, 5) = c3; Q6(cep, i, 6, 6) = c3;
      Q6(cep, i, 1, 2) = c4; Q6(cep, i, 1, 3) = c4; Q6(cep, i, 2, 1) = c4;
      Q6(cep, i, 3, 1) = c4; Q6(cep, i, 2, 3) = c4; Q6(cep, i, 3, 2) = c4;
    }
    // drive_01_update_b: stress_np1 = stress_n (first 6 columns) then += cep(i,m,k)*uddt(i,m)
    for (int k = 1; k <= 6; ++k)
      for (int i = 0; i < mxvl; ++i) V2(urcs_n1, i, k) = V2(urcs_n, i, k);
    for (int k = 1; k <= 6; ++k)
      for (int mm = 1; mm <= 6; ++mm)
        for (int i = 0; i < span; ++i)
          V2(urcs_n1, i, k) = V2(urcs_n1, i, k) + Q6(cep, i, mm, k) * V2(uddt, i, mm);
  }
  store_cep(span, gpn, gbl_cep, cep.data());
}

// drive_05_update (rstgp1.f:2373-2700) -> mm05 (mm05.f:9) + cnst5 (:1566), nonlinear_hardening
// or generalized_plasticity option by the sign of matprp(58).  Per-point arithmetic = w3d_oracle_mm05.h.
static void drive_05_update(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn,
                            int step, int iter, const double *uddt_displ, double *gbl_cep,
                            bool &material_cut_step, int *fatal) {
  using namespace w3do_mm05;
  const int span = bi.span;
  std::vector<double> cep(mxvl * 36, 0.0);
  double *urcs_n = &lw.urcs_blk_n[(size_t)(gpn - 1) * nstrs * mxvl];
  double *urcs_n1 = &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl];
  double *hist = &lw.elem_hist[(size_t)(gpn - 1) * bi.hist_size * span];
  double *hist1 = &lw.elem_hist1[(size_t)(gpn - 1) * bi.hist_size * span];
  const bool nonlinear_update = iter >= 1 || m->extrapolated_du != 0;
  const bool adaptive_possible = m->adaptive_flag != 0 && step > 1;
  if (bi.hist_size != 14 || !m->matprp) { *fatal = 5; return; }
  if (matprp(m, 58, bi.matnum) == 0.0f) { *fatal = 3; return; }   // option flag not set (inmat.f:2220, :2272)
  if (nonlinear_update) {
    // both options through the block-form restatements (w3d_oracle_mm05_fa.h / _gp.h), arrays as
    // the reference passes them: drive_05_update_c (rstgp1.f:2627-2700) -> mm05 -> mm05_fa or
    // mm05_gp, then cnst5 / cnst5_gp
    namespace fa = w3do_mm05_fa;
    namespace gp = w3do_mm05_gp;
    const bool generalized_pl = matprp(m, 58, bi.matnum) < 0.0f;
    std::vector<double> mmp((size_t)span * 10), tr((size_t)span * 6), sn((size_t)span * 9), s1((size_t)span * 9, 0.0),
        dp((size_t)span * 6), hn((size_t)span * 14), h1((size_t)span * 14);
    for (int i = 0; i < span; ++i) {
      const int elem = bi.felem + i;
      lw.e_vec[i] = rprop(m, 7, elem); lw.nu_vec[i] = rprop(m, 8, elem); lw.sigyld_vec[i] = rprop(m, 23, elem);
      for (int k = 0; k < 10; ++k) mmp[(size_t)i * 10 + k] = matprp(m, 55 + k, bi.matnum);
      for (int k = 0; k < 9; ++k) sn[(size_t)i * 9 + k] = urcs_n[(size_t)k * mxvl + i];
      for (int k = 0; k < 6; ++k) dp[(size_t)i * 6 + k] = uddt_displ[(size_t)k * mxvl + i] + 0.0;
      for (int k = 0; k < 14; ++k) {
        hn[(size_t)i * 14 + k] = hist[(size_t)k * span + i];
        h1[(size_t)i * 14 + k] = hist1[(size_t)k * span + i];
      }
    }
    fa::Block B;
    B.span = span;
    B.mm_props = reinterpret_cast<double (*)[10]>(mmp.data());
    B.e_vec = lw.e_vec; B.nu_vec = lw.nu_vec; B.sigyld_vec = lw.sigyld_vec;
    B.trial = reinterpret_cast<double (*)[6]>(tr.data());
    B.stress_n = reinterpret_cast<double (*)[9]>(sn.data());
    B.stress_np1 = reinterpret_cast<double (*)[9]>(s1.data());
    B.deps = reinterpret_cast<double (*)[6]>(dp.data());
    B.history_n = reinterpret_cast<double (*)[14]>(hn.data());
    B.history_np1 = reinterpret_cast<double (*)[14]>(h1.data());
    // generalized plasticity: setup_mm05_rknstr (rknstr.f:1282-1293), gp_tau_vec = matprp(56)
    std::vector<double> hu(span, (double)matprp(m, 55, bi.matnum)), tau(span, (double)matprp(m, 56, bi.matnum)),
        bu(span, (double)matprp(m, 57, bi.matnum)), du(span, (double)matprp(m, 59, bi.matnum));
    gp::GpVectors V{lw.e_vec, lw.e_vec, lw.nu_vec, lw.nu_vec, lw.sigyld_vec, lw.sigyld_vec, hu.data(), hu.data(),
                    bu.data(), bu.data(), du.data(), du.data(), tau.data()};
    const int rc = generalized_pl ? gp::mm05_gp(step, iter, B, V, adaptive_possible)
                                  : fa::mm05_fa(step, iter, B, adaptive_possible);
    if (rc == fa::CUT) { material_cut_step = true; return; }
    if (rc != fa::OK) { *fatal = rc; return; }
    std::vector<double> dm((size_t)span * 36);
    if (generalized_pl) gp::cnst5_gp(B, V, reinterpret_cast<double (*)[6][6]>(dm.data()));
    else fa::cnst5_fa(B, reinterpret_cast<double (*)[6][6]>(dm.data()));
    for (int i = 0; i < span; ++i) {
      for (int k = 0; k < 9; ++k) urcs_n1[(size_t)k * mxvl + i] = s1[(size_t)i * 9 + k];
      for (int k = 0; k < 14; ++k) hist1[(size_t)k * span + i] = h1[(size_t)i * 14 + k];
      for (int r = 1; r <= 6; ++r)
        for (int c = 1; c <= 6; ++c) Q6(cep, i, r, c) = dm[(size_t)i * 36 + (r - 1) * 6 + (c - 1)];
    }
    store_cep(span, gpn, gbl_cep, cep.data());
    return;
  }
  for (int i = 0; i < span; ++i) {
    const int elem = bi.felem + i;
    Props P;
    P.e = rprop(m, 7, elem); P.nu = rprop(m, 8, elem); P.sigyld = rprop(m, 23, elem);
    for (int k = 0; k < 10; ++k) P.mm[k] = matprp(m, 55 + k, bi.matnum);
    double cgn[9], cgn1[9], ud[6], h[14], h1[14], trial[6];
    for (int k = 0; k < 9; ++k) { cgn[k] = urcs_n[(size_t)k * mxvl + i]; cgn1[k] = 0.0; }
    for (int k = 0; k < 6; ++k) ud[k] = uddt_displ[(size_t)k * mxvl + i] + 0.0;
    for (int k = 0; k < 14; ++k) { h[k] = hist[(size_t)k * span + i]; h1[k] = hist1[(size_t)k * span + i]; }
    double D[6][6];
    if (nonlinear_update) {
      const int rc = update(step, iter, P, cgn, cgn1, ud, h, h1, trial, adaptive_possible);
      if (rc == CUT) { material_cut_step = true; return; }
      if (rc != OK) { *fatal = rc; return; }
      cnst5(P, trial, h1, cgn1, D);
      for (int k = 0; k < 14; ++k) hist1[(size_t)k * span + i] = h1[k];
    } else {
      // drive_05_update_a/_b: linear-elastic [D] and stress for iteration 0 (no history update)
      const double one = 1.0, two = 2.0;
      const double c1 = (P.e / ((one + P.nu) * (one - two * P.nu)));
      const double c2 = (one - P.nu) * c1, c3 = ((one - two * P.nu) / two) * c1, c4 = P.nu * c1;
      for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) D[r][c] = 0.0;
      D[0][0] = c2; D[1][1] = c2; D[2][2] = c2; D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
      D[0][1] = c4; D[0][2] = c4; D[1][0] = c4; D[2][0] = c4; D[1][2] = c4; D[2][1] = c4;
      for (int k = 0; k < 6; ++k) {
        cgn1[k] = cgn[k];
        for (int mm = 0; mm < 6; ++mm) cgn1[k] = cgn1[k] + D[mm][k] * ud[mm];
      }
    }
    for (int k = 0; k < 9; ++k) urcs_n1[(size_t)k * mxvl + i] = cgn1[k];
    for (int r = 1; r <= 6; ++r) for (int c = 1; c <= 6; ++c) Q6(cep, i, r, c) = D[r - 1][c - 1];
  }
  store_cep(span, gpn, gbl_cep, cep.data());
}

// drive_03_update (rstgp1.f:1418-1990) -> mm03p (mm03.f:10), mm03 (:2063), cnst3 (:3252); mises
// material without voids, constant or power-law hardening (w3d_oracle_mm03.h).  Property
// vectors: setup_mm03_rknstr (rknstr.f:1142-1215).  Not segmental, no temperatures, no fgm.
static void drive_03_update(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn,
                            int step, int iter, const double *uddt_displ, double *gbl_cep,
                            bool &material_cut_step, int *fatal) {
  using namespace w3do_mm03;
  const int span = bi.span;
  const double trans_factor = 0.95;
  std::vector<double> cep(mxvl * 36, 0.0);
  double *urcs_n = &lw.urcs_blk_n[(size_t)(gpn - 1) * nstrs * mxvl];
  double *urcs_n1 = &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl];
  double *hist = &lw.elem_hist[(size_t)(gpn - 1) * bi.hist_size * span];
  double *hist1 = &lw.elem_hist1[(size_t)(gpn - 1) * bi.hist_size * span];
  double *rtse = &lw.rtse[(size_t)(gpn - 1) * nstr * mxvl];
  if (bi.hist_size != HIST) { *fatal = 5; return; }
  if ((iprop(m, 24, bi.felem) & 4) != 0) { *fatal = 3; return; }          // segmental curves
  std::vector<Props> P(span);
  std::vector<Point> pt(span);
  for (int i = 0; i < span; ++i) {
    const int elem = bi.felem + i;
    Props &p = P[i];
    p.e = rprop(m, 7, elem); p.nu = rprop(m, 8, elem); p.h_fixed = rprop(m, 15, elem);
    p.sigyld = rprop(m, 23, elem); p.n_power = rprop(m, 21, elem); p.eps_ref = rprop(m, 22, elem);
    p.m_power = rprop(m, 20, elem); p.f0 = rprop(m, 26, elem);
    p.q1 = rprop(m, 27, elem); p.q2 = rprop(m, 28, elem); p.q3 = rprop(m, 29, elem);
    p.nucleation = (iprop(m, 30, elem) & 1) != 0;
    if (p.f0 > 0.0 || p.nucleation) { *fatal = 3; return; }                // Gurson: not restated
    // copy_sigyld_vec (rstgp1.f:1898-1906): reduced for the cubic transition
    if (p.n_power > 0.0) p.sigyld = p.sigyld * trans_factor;
    for (int k = 0; k < 6; ++k) pt[i].deps[k] = uddt_displ[(size_t)k * mxvl + i] + 0.0;   // + uddt_temps
    for (int k = 0; k < 9; ++k) { pt[i].stress_n[k] = urcs_n[(size_t)k * mxvl + i]; pt[i].stress_n1[k] = 0.0; }
    for (int k = 0; k < HIST; ++k) {
      pt[i].history[k] = hist[(size_t)k * span + i];
      pt[i].history1[k] = hist1[(size_t)k * span + i];
    }
    for (int k = 0; k < 6; ++k) pt[i].stress_n1_elas[k] = rtse[(size_t)k * mxvl + i];
    pt[i].null_point = false; pt[i].nonlinear_flag = false;
    pt[i].p_trial = pt[i].q_trial = pt[i].yld_func = 0.0;
  }
  const bool nonlinear_update = iter >= 1 || m->extrapolated_du != 0;
  const bool adaptive = m->adaptive_flag != 0 && step > 1;
  const bool allow_cut = adaptive && (iprop(m, 24, bi.felem) & 2) != 0;    // rknstr.f:1213-1214
  if (nonlinear_update) {
    const bool process_block = mm03p(step, span, pt.data(), P.data());
    if (process_block) {
      for (int i = 0; i < span; ++i) {
        if (pt[i].null_point) continue;
        if (!pt[i].nonlinear_flag) continue;
        if (mm03(P[i], m->dt, pt[i])) {
          if (allow_cut) { material_cut_step = true; return; }
          *fatal = 7;                                                       // abort_job
          return;
        }
      }
    }
    for (int i = 0; i < span; ++i) {
      double D[6][6];
      cnst3(iter, P[i], pt[i].stress_n1_elas, pt[i].history, pt[i].history1, D);
      for (int r = 1; r <= 6; ++r) for (int c = 1; c <= 6; ++c) Q6(cep, i, r, c) = D[r - 1][c - 1];
    }
    for (int i = 0; i < span; ++i) {
      for (int k = 0; k < 9; ++k) urcs_n1[(size_t)k * mxvl + i] = pt[i].stress_n1[k];   // drive_03_update_c
      for (int k = 0; k < HIST; ++k) hist1[(size_t)k * span + i] = pt[i].history1[k];
      for (int k = 0; k < 6; ++k) rtse[(size_t)k * mxvl + i] = pt[i].stress_n1_elas[k];
    }
  } else {
    // drive_03_update_b / _e: linear-elastic [D], no history update
    const double one = 1.0, two = 2.0;
    for (int i = 0; i < span; ++i) {
      const double e = P[i].e, nu = P[i].nu;
      const double c1 = (e / ((one + nu) * (one - two * nu)));
      const double c2 = (one - nu) * c1;
      const double c3 = ((one - two * nu) / two) * c1;
      const double c4 = nu * c1;
      Q6(cep, i, 1, 1) = c2; Q6(cep, i, 2, 2) = c2; Q6(cep, i, 3, 3) = c2;
      Q6(cep, i, 4, 4) = c3; Q6(cep, i, 5, 5) = c3; Q6(cep, i, 6, 6) = c3;
      Q6(cep, i, 1, 2) = c4; Q6(cep, i, 1, 3) = c4; Q6(cep, i, 2, 1) = c4;
      Q6(cep, i, 3, 1) = c4; Q6(cep, i, 2, 3) = c4; Q6(cep, i, 3, 2) = c4;
    }
    for (int k = 1; k <= 6; ++k)
      for (int i = 0; i < mxvl; ++i) V2(urcs_n1, i, k) = V2(urcs_n, i, k);
    for (int k = 1; k <= 6; ++k)
      for (int mm = 1; mm <= 6; ++mm)
        for (int i = 0; i < span; ++i)
          V2(urcs_n1, i, k) = V2(urcs_n1, i, k) + Q6(cep, i, mm, k) * pt[i].deps[mm - 1];
  }
  store_cep(span, gpn, gbl_cep, cep.data());
}

// drive_10_update (rstgp1.f:4327-4513) -> mm10 (mm10_a.f:29).  The per-Gauss-point crystal update
// is the statement-for-statement restatement in w3d_oracle_mm10.h; this wrapper supplies the
// reference's block-local layouts.
// NS = capacity of the slip-system arrays: 12 (fcc, bcc, ...: what the engine builds) or 48
// (bcc48: CPU only, used to check the shared core against the reference's published results)
template <int NS>
static void drive_10_update_ns(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn,
                               int step, int iter, const double *uddt_displ, bool &material_cut_step,
                               int *fatal) {
  using namespace w3do_mm10;
  const int span = bi.span;
  const w3do_cp *cp = (const w3do_cp *)m->cp;
  if (!cp || cp->nslip > NS || cp->num_hard != 1) { *fatal = 3; return; }
  const Layout L = Layout::make(cp->nslip, cp->num_hard);
  if (L.size(cp->ncrystals) != bi.hist_size) { *fatal = 5; return; }
  // small displacement analysis: rotations set to identity at the first point (:4408-4421)
  if (!bi.geonl && gpn == 1) {
    std::fill(lw.rot_blk_n1.begin(), lw.rot_blk_n1.end(), 0.0);
    for (int igp = 0; igp < bi.ngp; ++igp)
      for (int i = 0; i < mxvl; ++i) {
        lw.rot_blk_n1[((size_t)igp * 9 + 0) * mxvl + i] = 1.0;
        lw.rot_blk_n1[((size_t)igp * 9 + 4) * mxvl + i] = 1.0;
        lw.rot_blk_n1[((size_t)igp * 9 + 8) * mxvl + i] = 1.0;
      }
  }
  bool iter_0_extrapolate_off = false;
  if (iter == 0) {
    iter_0_extrapolate_off = m->extrapolated_du == 0;
    if (step == 1) iter_0_extrapolate_off = true;
  }
  double *urcs_n = &lw.urcs_blk_n[(size_t)(gpn - 1) * nstrs * mxvl];
  double *urcs_n1 = &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl];
  const double *rot = &lw.rot_blk_n1[(size_t)(gpn - 1) * 9 * mxvl];
  double *hist = &lw.elem_hist[(size_t)(gpn - 1) * bi.hist_size * span];
  double *hist1 = &lw.elem_hist1[(size_t)(gpn - 1) * bi.hist_size * span];
  const size_t prec = sizeof(Props<NS>) / sizeof(double);
  for (int i = 0; i < span; ++i) {
    const int elem = bi.felem + i;
    const Props<NS> *P =
        (const Props<NS> *)(cp->props + (size_t)(elem - 1) * cp->ncrystals * prec);
    double Rur[9], ud[6], cgn[9], cgn1[9];
    for (int k = 0; k < 9; ++k) Rur[k] = rot[(size_t)k * mxvl + i];
    for (int k = 0; k < 6; ++k) ud[k] = uddt_displ[(size_t)k * mxvl + i] + 0.0;
    for (int k = 0; k < 9; ++k) cgn[k] = urcs_n[(size_t)k * mxvl + i];
    const CHist hn{hist + i, span};
    const Hist h1{hist1 + i, span};
    const bool cut = gauss_point_update<NS>(step, cp->ncrystals, P, L, hn, h1, Rur, ud, m->dt,
                                            iter_0_extrapolate_off, cp->asymmetric_assembly != 0,
                                            cgn, cgn1);
    if (cut) { material_cut_step = true; return; }
    for (int k = 0; k < 9; ++k) urcs_n1[(size_t)k * mxvl + i] = cgn1[k];
  }
}
static void drive_10_update(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn,
                            int step, int iter, const double *uddt_displ, bool &material_cut_step,
                            int *fatal) {
  const w3do_cp *cp = (const w3do_cp *)m->cp;
  if (cp && cp->nslip > 12)
    drive_10_update_ns<48>(m, bi, lw, gpn, step, iter, uddt_displ, material_cut_step, fatal);
  else
    drive_10_update_ns<12>(m, bi, lw, gpn, step, iter, uddt_displ, material_cut_step, fatal);
}

// rknstr_finish_cp (rknstr.f:268-296) -> mm10_calc_grads (mm10_a.f:1579): gradient of the elastic
// rotation by a linear fit over the Gauss points, stored in history at n+1 for every point.
static void finish_cp(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int step) {
  using namespace w3do_mm10;
  const w3do_cp *cp = (const w3do_cp *)m->cp;
  const int span = bi.span, ngp = bi.ngp, H = bi.hist_size;
  const Layout L = Layout::make(cp->nslip, cp->num_hard);
  for (int i = 0; i < span; ++i) {
    double grads[27];
    if (cp->ncrystals > 1) {
      for (int k = 0; k < 27; ++k) grads[k] = 0.0;
    } else {
      double rt[8][9], jacinv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int gp = 0; gp < ngp; ++gp) {
        double Rp[9], rot[9];
        for (int k = 0; k < 9; ++k) {
          // the block-local history at n was initialised at step 1 (Rp = I)
          Rp[k] = (step == 1) ? ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0)
                              : lw.elem_hist[((size_t)gp * H + L.c_Rp(0) + k) * span + i];
          rot[k] = lw.rot_blk_n1[((size_t)gp * 9 + k) * mxvl + i];
        }
        if (bi.geonl) {
          for (int a = 0; a < 3; ++a)      // mult_type_3t: Rt = Rp * rot^T
            for (int b = 0; b < 3; ++b)
              rt[gp][b * 3 + a] = Rp[0 * 3 + a] * rot[0 * 3 + b] + Rp[1 * 3 + a] * rot[1 * 3 + b] +
                                  Rp[2 * 3 + a] * rot[2 * 3 + b];
        }
      }
      if (!bi.geonl) {
        // mm10_a_copy_vector(Rt3d, Rps, ngp*9) (mm10_a.f:1622): the FIRST ngp*9 contiguous words
        // of the block-local array starting at elem_hist(1, rs, 1) -- i.e. word rs + k / span of
        // element 1 + mod(k, span) at Gauss point 1 -- whatever now_elem is.  Every element of a
        // small-strain block therefore fits the same 72 numbers (for span >= 72: Rp(1,1) of the
        // block's first 72 elements).  Reproduced literally: the bar is the reference's result.
        double *flat = &rt[0][0];
        for (int k = 0; k < ngp * 9; ++k) {
          const int word = L.c_Rp(0) + k / span, el = k % span;   // word >= H: next Gauss-point plane
          if (step == 1) {
            const int q = word % H - L.c_Rp(0);   // block-local history at n initialised at step 1:
            flat[k] = (q == 0 || q == 4 || q == 8) ? 1.0 : 0.0;   // Rp = I (other words taken as 0)
          } else {
            flat[k] = lw.elem_hist[(size_t)word * span + el];
          }
        }
      }
      if (bi.geonl) {
        // local_work%jac holds the last Jacobian formed: undeformed configuration, last point
        double a[3][3];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) a[r][c] = lw.jac[(size_t)(c * 3 + r) * mxvl + i];
        const double det = a[0][0] * a[1][1] * a[2][2] - a[0][0] * a[1][2] * a[2][1] -
                           a[0][1] * a[1][0] * a[2][2] + a[0][1] * a[1][2] * a[2][0] +
                           a[0][2] * a[1][0] * a[2][1] - a[0][2] * a[1][1] * a[2][0];
        double cof[3][3];
        cof[0][0] = +(a[1][1] * a[2][2] - a[1][2] * a[2][1]);
        cof[0][1] = -(a[1][0] * a[2][2] - a[1][2] * a[2][0]);
        cof[0][2] = +(a[1][0] * a[2][1] - a[1][1] * a[2][0]);
        cof[1][0] = -(a[0][1] * a[2][2] - a[0][2] * a[2][1]);
        cof[1][1] = +(a[0][0] * a[2][2] - a[0][2] * a[2][0]);
        cof[1][2] = -(a[0][0] * a[2][1] - a[0][1] * a[2][0]);
        cof[2][0] = +(a[0][1] * a[1][2] - a[0][2] * a[1][1]);
        cof[2][1] = -(a[0][0] * a[1][2] - a[0][2] * a[1][0]);
        cof[2][2] = +(a[0][0] * a[1][1] - a[0][1] * a[1][0]);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) jacinv[r][c] = (std::fabs(det) <= 1.0e-10) ? 0.0 : cof[c][r] / det;
      }
      calc_grads(rt, bi.geonl, jacinv, grads);
    }
    for (int gp = 0; gp < ngp; ++gp)
      for (int k = 0; k < 27; ++k)
        lw.elem_hist1[((size_t)gp * H + L.o_gradfe() + k) * span + i] = grads[k];
  }
}

// rstgp1 (rstgp1.f:19-332) geonl ; rstgp2 (rstgp1.f:351-598) small strain
static void rstgp(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int gpn, int step,
                  int iter, double *gbl_cep, double &block_energy, double &block_plastic_work,
                  bool &material_cut_step, int *fatal) {
  const int span = bi.span, nnode = bi.nnode;
  double internal_energy = block_energy, plastic_work = block_plastic_work;
  std::vector<double> ddt(mxvl * nstr, 0.0), uddt(mxvl * nstr, 0.0), qnhalf(mxvl * 36, 0.0);
  const double *nxi = &lw.nxi[(gpn - 1) * mxndel], *neta = &lw.neta[(gpn - 1) * mxndel],
               *nzeta = &lw.nzeta[(gpn - 1) * mxndel];
  double *ddtse = &lw.ddtse[(size_t)(gpn - 1) * nstr * mxvl];
  if (bi.geonl) {
    // rstgp1_a (rstgp1.f:5168-5187)
    const double half = 0.5;
    for (int j = 1; j <= bi.ndof * nnode; ++j)
      for (int i = 0; i < span; ++i) {
        V2(lw.uenh, i, j) = std::fma(half, V2(lw.due, i, j), V2(lw.ue, i, j));   // contraction convention, w3d_oracle.cpp
        V2(lw.uen1, i, j) = V2(lw.ue, i, j) + V2(lw.due, i, j);
      }
    int error = 0;
    gtmat1(bi, lw, gpn, qnhalf.data(), error);
    if (error == 1) {
      // rstgp1.f:136-145: adaptive -> request step cut; else abort_job
      const bool adaptive = m->adaptive_flag != 0 && step > 1;
      if (!adaptive) *fatal = 2;
      material_cut_step = true;
      return;
    }
    gtlsn1(span, nnode, lw.due.data(), ddt.data(), &lw.gama_mid[(size_t)(gpn - 1) * 9 * mxvl], nxi,
           neta, nzeta, lw.vol_block.data(), bi.bbar, m->eps_bbar, lw.b.data());
    qmply1(span, qnhalf.data(), ddt.data(), uddt.data());
  } else {
    gtlsn1(span, nnode, lw.due.data(), uddt.data(), &lw.gama[(size_t)(gpn - 1) * 9 * mxvl], nxi,
           neta, nzeta, lw.vol_block.data(), bi.bbar, m->eps_bbar, lw.b.data());
  }
  // rstgp1_update_strains (rstgp1.f:4983-5005)
  for (int k = 1; k <= 6; ++k)
    for (int i = 0; i < span; ++i) V2(ddtse, i, k) = V2(ddtse, i, k) + V2(uddt, i, k);

  switch (bi.mat_type) {
    case 1:
      drive_01_update(m, bi, lw, gpn, step, iter, uddt.data(), gbl_cep, fatal);
      break;
    case 5:
      drive_05_update(m, bi, lw, gpn, step, iter, uddt.data(), gbl_cep, material_cut_step, fatal);
      if (material_cut_step || *fatal) return;
      break;
    case 3:
      drive_03_update(m, bi, lw, gpn, step, iter, uddt.data(), gbl_cep, material_cut_step, fatal);
      if (material_cut_step || *fatal) return;
      break;
    case 10:
      drive_10_update(m, bi, lw, gpn, step, iter, uddt.data(), material_cut_step, fatal);
      if (material_cut_step || *fatal) return;
      break;
    default:
      *fatal = 3;  // material model not in the oracle
      return;
  }
  if (iter != 0) {
    const double *u7 = &lw.urcs_blk_n1[((size_t)(gpn - 1) * nstrs + 6) * mxvl];
    const double *u8 = &lw.urcs_blk_n1[((size_t)(gpn - 1) * nstrs + 7) * mxvl];
    rstgp1_b(span, internal_energy, plastic_work, u7, u8, &lw.det_j[(gpn - 1) * mxvl],
             lw.dfn1.data(), bi.geonl ? 1 : 2, m->beta_fact, lw.weights[gpn - 1]);
    block_energy = internal_energy;
    block_plastic_work = plastic_work;
  }
}

// rknstr (rknstr.f:15-184) incl. rknstr_geonl (455-547), rknstr_sm_displ (563-614),
// setup_mm01_rknstr (1018-1066)
static void rknstr(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, int step, int iter,
                   double *gbl_cep, double &block_energy, double &block_plastic_work,
                   bool &material_cut_step, int *fatal) {
  const int span = bi.span, nnode = bi.nnode, ngp = bi.ngp;
  if (bi.bbar && bi.elem_type == 2) {
    std::fill(lw.vol_block.begin(), lw.vol_block.end(), 0.0);
    std::fill(lw.volume_block.begin(), lw.volume_block.end(), 0.0);
  }
  for (int gpn = 1; gpn <= ngp; ++gpn) {
    double xi, eta, zeta;
    quad2(gpn, xi, eta, zeta, lw.weights[gpn - 1]);
    double *nxi = &lw.nxi[(gpn - 1) * mxndel], *neta = &lw.neta[(gpn - 1) * mxndel],
           *nzeta = &lw.nzeta[(gpn - 1) * mxndel];
    deriv2(xi, eta, zeta, nxi, neta, nzeta);
    if (bi.geonl) {
      jacob1(span, lw.jac.data(), &lw.det_j_mid[(gpn - 1) * mxvl],
             &lw.gama_mid[(size_t)(gpn - 1) * 9 * mxvl], nxi, neta, nzeta, lw.ce_mid.data(), nnode);
      if (bi.bbar)
        vol_terms(&lw.gama_mid[(size_t)(gpn - 1) * 9 * mxvl], &lw.det_j_mid[(gpn - 1) * mxvl],
                  lw.vol_block.data(), nxi, neta, nzeta, lw.volume_block.data(), span);
    } else {
      jacob1(span, lw.jac.data(), &lw.det_j[(gpn - 1) * mxvl],
             &lw.gama[(size_t)(gpn - 1) * 9 * mxvl], nxi, neta, nzeta, lw.ce_0.data(), nnode);
      if (bi.bbar)
        vol_terms(&lw.gama[(size_t)(gpn - 1) * 9 * mxvl], &lw.det_j[(gpn - 1) * mxvl],
                  lw.vol_block.data(), nxi, neta, nzeta, lw.volume_block.data(), span);
    }
  }
  if (bi.bbar && bi.elem_type == 2) vol_avg(lw.vol_block.data(), lw.volume_block.data(), span);

  // setup_mm01_rknstr (rknstr.f:1018-1066): single-precision props widened to double
  if (bi.mat_type == 1) {
    for (int i = 0; i < span; ++i) {
      const int elem = bi.felem + i;
      lw.e_vec[i] = rprop(m, 7, elem);
      lw.e_vec_n[i] = rprop(m, 7, elem);
      lw.nu_vec[i] = rprop(m, 8, elem);
      lw.nu_vec_n[i] = rprop(m, 8, elem);
      lw.beta_vec[i] = rprop(m, 14, elem);
      lw.h_vec[i] = rprop(m, 15, elem);
      lw.lnelas_vec[i] = lprop(m, 17, elem);
      lw.sigyld_vec[i] = rprop(m, 23, elem);
    }
  }
  for (int gpn = 1; gpn <= ngp; ++gpn) {
    rstgp(m, bi, lw, gpn, step, iter, gbl_cep, block_energy, block_plastic_work,
          material_cut_step, fatal);
    if (material_cut_step || *fatal) return;
  }
  if (bi.mat_type == 10) finish_cp(m, bi, lw, step);   // rknstr.f:182
}

// gpifv1 (gpifv1.f:16-165), solid branch
static void gpifv1(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, double *eleifv
                   /*(span,totdof)*/, int gpn, double *element_volumes) {
  const int span = bi.span, nnode = bi.nnode, totdof = bi.totdof;
  const double w = lw.weights[gpn - 1] * 1.0;
  const double *dj = &lw.det_j[(gpn - 1) * mxvl];
  double *b = lw.b.data();
  blcmp1(span, b, &lw.gama[(size_t)(gpn - 1) * 9 * mxvl], &lw.nxi[(gpn - 1) * mxndel],
         &lw.neta[(gpn - 1) * mxndel], &lw.nzeta[(gpn - 1) * mxndel], nnode);
  if (bi.bbar && bi.elem_type == 2) bmod(b, lw.vol_block.data(), span, m->eps_bbar);
  std::vector<double> qtn1(mxvl * 36), cs_blk_n1(mxvl * nstr);
  const double *sig = &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl];
  if (bi.geonl) {
    getrm1(span, qtn1.data(), &lw.rot_blk_n1[(size_t)(gpn - 1) * 9 * mxvl], 2);
    qmply1(span, qtn1.data(), sig, cs_blk_n1.data());
    sig = cs_blk_n1.data();
  }
  for (int j = 1; j <= totdof; ++j)
    for (int i = 0; i < span; ++i)
      eleifv[(size_t)(j - 1) * span + i] =
          eleifv[(size_t)(j - 1) * span + i] + BM(b, i, j, 1) * V2(sig, i, 1) * w * dj[i] +
          BM(b, i, j, 2) * V2(sig, i, 2) * w * dj[i] + BM(b, i, j, 3) * V2(sig, i, 3) * w * dj[i] +
          BM(b, i, j, 4) * V2(sig, i, 4) * w * dj[i] + BM(b, i, j, 5) * V2(sig, i, 5) * w * dj[i] +
          BM(b, i, j, 6) * V2(sig, i, 6) * w * dj[i];
  for (int i = 0; i < span; ++i) element_volumes[i] = element_volumes[i] + w * dj[i];
}

// rknifv (rknifv.f:15-100) + rknifv_std (141-215)
static void rknifv(const w3do_model *m, const BlockInfo &bi, LocalWork &lw, double *eleifv,
                   double *updated_element_volumes) {
  const int span = bi.span, nnode = bi.nnode, ngp = bi.ngp, totdof = bi.totdof;
  double element_volumes_for_blk[mxvl];
  for (int k = 0; k < span * totdof; ++k) eleifv[k] = 0.0;
  for (int i = 0; i < span; ++i) element_volumes_for_blk[i] = 0.0;
  if (bi.bbar && bi.elem_type == 2) {
    std::fill(lw.vol_block.begin(), lw.vol_block.end(), 0.0);
    std::fill(lw.volume_block.begin(), lw.volume_block.end(), 0.0);
  }
  for (int gpn = 1; gpn <= ngp; ++gpn) {
    double xi, eta, zeta;
    quad2(gpn, xi, eta, zeta, lw.weights[gpn - 1]);
    double *nxi = &lw.nxi[(gpn - 1) * mxndel], *neta = &lw.neta[(gpn - 1) * mxndel],
           *nzeta = &lw.nzeta[(gpn - 1) * mxndel];
    deriv2(xi, eta, zeta, nxi, neta, nzeta);
    jacob1(span, lw.jac.data(), &lw.det_j[(gpn - 1) * mxvl], &lw.gama[(size_t)(gpn - 1) * 9 * mxvl],
           nxi, neta, nzeta, lw.ce_n1.data(), nnode);
    if (bi.bbar && bi.elem_type == 2)
      vol_terms(&lw.gama[(size_t)(gpn - 1) * 9 * mxvl], &lw.det_j[(gpn - 1) * mxvl],
                lw.vol_block.data(), nxi, neta, nzeta, lw.volume_block.data(), span);
  }
  if (bi.bbar && bi.elem_type == 2) vol_avg(lw.vol_block.data(), lw.volume_block.data(), span);
  for (int gpn = 1; gpn <= ngp; ++gpn) gpifv1(m, bi, lw, eleifv, gpn, element_volumes_for_blk);
  // update_element_volumes (rknifv.f:233-251), growth_by_kill = .false.
  if (updated_element_volumes)
    for (int i = 0; i < span; ++i) updated_element_volumes[i] = element_volumes_for_blk[i];
}

// do_nleps_block (drive_eps_sig_internal_forces.f:399-632)
static void do_nleps_block(const w3do_model *m, const BlockPtrs &bp, int blk, int iter, int step,
                           const double *u, const double *du, bool &material_cut_step,
                           double &block_energy, double &block_plastic_work, double *einfvec
                           /*(span,totdof)*/, int *fatal) {
  const BlockInfo bi = block_info(m, blk);
  const int span = bi.span, totdof = bi.totdof, nnode = bi.nnode, ngp = bi.ngp;
  if (bi.elem_type != 2 || nnode != 8 || ngp != 8 || bi.ndof != 3 || span > mxvl) {
    *fatal = 4;  // only hex8 (l3disop) 2x2x2 in the oracle
    return;
  }
  LocalWork lw;
  lw.elem_hist.assign((size_t)span * bi.hist_size * ngp, 0.0);
  lw.elem_hist1.assign((size_t)span * bi.hist_size * ngp, 0.0);
  material_cut_step = false;
  block_energy = 0.0;
  block_plastic_work = 0.0;

  // dupstr (dupstr.f:82-88)
  for (int j = 1; j <= totdof; ++j)
    for (int i = 0; i < span; ++i) {
      const int dof = bp.edest[(size_t)i * totdof + (j - 1)];
      V2(lw.ue, i, j) = u[dof - 1];
      V2(lw.due, i, j) = du[dof - 1];
    }
  // dupstr_blocked (drive_eps_sig_internal_forces.f:1699-1754)
  {
    int k = 1;
    for (int j = 1; j <= nnode; ++j) {
      for (int i = 0; i < span; ++i) {
        V2(lw.ce_0, i, k) = m->c[bp.cdest[(size_t)i * totdof + (k - 1)] - 1];
        V2(lw.ce_0, i, k + 1) = m->c[bp.cdest[(size_t)i * totdof + (k)] - 1];
        V2(lw.ce_0, i, k + 2) = m->c[bp.cdest[(size_t)i * totdof + (k + 1)] - 1];
      }
      k = k + 3;
    }
    bool update = true;
    if (step == 1 && iter == 0) update = false;
    const bool update_coords = bi.geonl && update;
    const double half = 0.5;
    for (int j = 1; j <= totdof; ++j)
      for (int i = 0; i < span; ++i) {
        if (update_coords) {
          V2(lw.ce_n, i, j) = V2(lw.ce_0, i, j) + V2(lw.ue, i, j);
          V2(lw.ce_mid, i, j) = V2(lw.ce_0, i, j) + V2(lw.ue, i, j) + half * V2(lw.due, i, j);
          V2(lw.ce_n1, i, j) = V2(lw.ce_0, i, j) + V2(lw.ue, i, j) + V2(lw.due, i, j);
        } else {
          V2(lw.ce_n, i, j) = V2(lw.ce_0, i, j);
          V2(lw.ce_mid, i, j) = V2(lw.ce_0, i, j);
          V2(lw.ce_n1, i, j) = V2(lw.ce_0, i, j);
        }
      }
  }
  if (bi.hist_size > 0) copy_history_in(lw.elem_hist.data(), bp.history, ngp, bi.hist_size, span);
  gastr(lw.ddtse.data(), bp.eps_n, ngp, nstr, span);
  gastr(lw.strain_n.data(), bp.eps_n, ngp, nstr, span);
  gastr(lw.urcs_blk_n.data(), bp.urcs_n, ngp, nstrs, span);

  rknstr(m, bi, lw, step, iter, bp.cep, block_energy, block_plastic_work, material_cut_step,
         fatal);
  if (material_cut_step || *fatal) return;

  // rplstr (rplstr.f:15-93)
  scstr(lw.urcs_blk_n1.data(), bp.urcs_n1, ngp, nstrs, span);
  scstr(lw.ddtse.data(), bp.eps_n1, ngp, nstr, span);
  if (iter > 0 && bi.geonl) scstr(lw.rot_blk_n1.data(), bp.rot_n1, ngp, 9, span);
  const bool save_history_1 = iter > 0;
  const bool save_history_2 = bi.mat_type == 10;
  if ((save_history_1 || save_history_2) && bi.hist_size > 0)
    copy_history_out(lw.elem_hist1.data(), bp.history1, ngp, bi.hist_size, span);

  rknifv(m, bi, lw, einfvec, bp.element_vol);
}

}  // namespace w3do

using namespace w3do;

extern "C" {

int64_t w3do_block_offset(const w3do_model *m, int32_t blk, int32_t which) {
  return block_offset(m, blk, which);
}
void w3do_quad2(int32_t gpn, double *xi, double *eta, double *zeta, double *w) {
  quad2(gpn, *xi, *eta, *zeta, *w);
}
void w3do_deriv2(double xi, double eta, double zeta, double *nxi, double *neta, double *nzeta) {
  deriv2(xi, eta, zeta, nxi, neta, nzeta);
}
void w3do_cp_icp(int32_t n, int32_t *cp, int32_t *dcp, int32_t *icp) { make_cp_icp(n, cp, dcp, icp); }

int32_t w3do_mm10_props_doubles(void) { return (int32_t)(sizeof(w3do_mm10::Props<12>) / sizeof(double)); }
int32_t w3do_mm10_crystal_type_doubles(void) { return (int32_t)(sizeof(w3do_mm10::CrystalType) / sizeof(double)); }
int32_t w3do_mm10_hist_size(int32_t nslip, int32_t num_hard, int32_t ncrystals) {
  return w3do_mm10::Layout::make(nslip, num_hard).size(ncrystals);
}
// crystal_type: w3do_mm10::CrystalType with the trailing ints widened to doubles (oracle_py.py)
void w3do_mm10_setup_props(const double *ctd, const double *angles, int32_t angle_type,
                           double *props_out) {
  w3do_mm10::CrystalType ct;
  const size_t nd = 36 + 144 + 144 + 13;
  std::memcpy(&ct, ctd, nd * sizeof(double));
  const double *ti = ctd + nd;
  ct.nslip = (int)ti[0]; ct.num_hard = (int)ti[1]; ct.h_type = (int)ti[2]; ct.tang_calc = (int)ti[3];
  ct.miter = (int)ti[4]; ct.solver = (int)ti[5]; ct.alter_mode = (int)ti[6];
  if (ct.nslip > 12) {
    static w3do_mm10::Props<48> P48;   // 48-system record (bcc48), see drive_10_update
#pragma omp critical(w3do_props48)
    {
      std::memset(&P48, 0, sizeof P48);
      w3do_mm10::setup_props<48>(ct, angles, angle_type, P48);
      std::memcpy(props_out, &P48, sizeof P48);
    }
    return;
  }
  w3do_mm10::Props<12> P;
  std::memset(&P, 0, sizeof P);
  w3do_mm10::setup_props<12>(ct, angles, angle_type, P);
  std::memcpy(props_out, &P, sizeof P);
}
int32_t w3do_mm10_props_doubles_for(int32_t nslip) {
  return (int32_t)((nslip > 12 ? sizeof(w3do_mm10::Props<48>) : sizeof(w3do_mm10::Props<12>)) / sizeof(double));
}

// n records at once: cnum (1-based into the table of crystal types), angles (3 per record)
void w3do_mm10_setup_props_batch(const double *ct_table, const int32_t *cnum, const double *angles,
                                 int64_t n, int32_t angle_type, double *props_out) {
  const size_t ctd = sizeof(w3do_mm10::CrystalType) / sizeof(double) + 0;
  const size_t rec = 36 + 144 + 144 + 13 + 7;
  (void)ctd;
  // all crystal types of a model share one record size (the 48-system one if any type needs it)
  int nslip_max = 0;
  for (int64_t k = 0; k < n; ++k) {
    const int ns = (int)ct_table[(size_t)(cnum[k] - 1) * rec + 337];
    if (ns > nslip_max) nslip_max = ns;
  }
  const size_t pd = (size_t)w3do_mm10_props_doubles_for(nslip_max);
  for (int64_t k = 0; k < n; ++k)
    w3do_mm10_setup_props(ct_table + (size_t)(cnum[k] - 1) * rec, angles + 3 * k, angle_type,
                          props_out + (size_t)k * pd);
}

// drive_eps_sig_internal_forces.f:20-387
int32_t w3do_drive_eps_sig_internal_forces(const w3do_model *m, w3do_state *s, int32_t step,
                                           int32_t iter, const double *u, const double *du,
                                           double *ifv, w3do_scalars *out, int32_t nthreads) {
  const int nelblk = m->nelblk;
  std::vector<double> block_energies(nelblk, 0.0), block_plastic_work(nelblk, 0.0);
  std::vector<char> step_cut_flags(nelblk, 0);
  std::vector<int> fatals(nelblk, 0);
  // allocate_ifv(1): blocks of element internal force vectors (span,totdof)
  std::vector<int64_t> ifv_off(nelblk + 1, 0);
  for (int blk = 0; blk < nelblk; ++blk) {
    BlockInfo bi = block_info(m, blk);
    ifv_off[blk + 1] = ifv_off[blk] + (int64_t)bi.span * bi.totdof;
  }
  std::vector<double> einfvec(ifv_off[nelblk], 0.0);
  const std::vector<BlockPtrs> bps = all_block_ptrs(m, s);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int blk = 0; blk < nelblk; ++blk) {
    bool cut = false;
    do_nleps_block(m, bps[blk], blk, iter, step, u, du, cut, block_energies[blk], block_plastic_work[blk],
                   &einfvec[ifv_off[blk]], &fatals[blk]);
    step_cut_flags[blk] = cut ? 1 : 0;
  }
  int fatal = 0;
  bool material_cut_step = false;
  double internal_energy = 0.0, plastic_work = 0.0;
  for (int blk = 0; blk < nelblk; ++blk) {
    if (fatals[blk] && !fatal) fatal = fatals[blk];
    if (step_cut_flags[blk]) material_cut_step = true;
    internal_energy = internal_energy + block_energies[blk];
    plastic_work = plastic_work + block_plastic_work[blk];
  }
  out->material_cut_step = material_cut_step ? 1 : 0;
  out->internal_energy = internal_energy;
  out->plastic_work = plastic_work;
  for (int i = 0; i < m->nodof; ++i) ifv[i] = 0.0;
  out->sum_ifv = 0.0;
  out->num_term_ifv = 0;
  if (fatal) return fatal;
  if (material_cut_step) return 0;
  // addifv_standard (addifv.f:86-113)
  double sum_ifv = 0.0;
  int64_t num_term_ifv = 0;
  if (nthreads == 1) {
    for (int blk = 0; blk < nelblk; ++blk) {
      BlockInfo bi = block_info(m, blk);
      const int32_t *bedst = bps[blk].edest;
      const double *eleifv = &einfvec[ifv_off[blk]];
      for (int j = 1; j <= bi.totdof; ++j)
        for (int i = 0; i < bi.span; ++i)
          ifv[bedst[(size_t)i * bi.totdof + (j - 1)] - 1] += eleifv[(size_t)(j - 1) * bi.span + i];
      for (int j = 1; j <= bi.totdof; ++j)
        for (int i = 0; i < bi.span; ++i)
          sum_ifv = sum_ifv + std::fabs(eleifv[(size_t)(j - 1) * bi.span + i]);
      num_term_ifv += (int64_t)bi.totdof * bi.span;
    }
  } else {
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : sum_ifv, num_term_ifv)
    for (int blk = 0; blk < nelblk; ++blk) {
      BlockInfo bi = block_info(m, blk);
      const int32_t *bedst = bps[blk].edest;
      const double *eleifv = &einfvec[ifv_off[blk]];
      for (int j = 1; j <= bi.totdof; ++j)
        for (int i = 0; i < bi.span; ++i) {
          const double v = eleifv[(size_t)(j - 1) * bi.span + i];
#pragma omp atomic update
          ifv[bedst[(size_t)i * bi.totdof + (j - 1)] - 1] += v;
          sum_ifv += std::fabs(v);
        }
      num_term_ifv += (int64_t)bi.totdof * bi.span;
    }
  }
  out->sum_ifv = sum_ifv;
  out->num_term_ifv = num_term_ifv;
  // beta_fact scaling of ifv (drive_eps_sig_internal_forces.f:358-359)
  if (m->beta_fact != 1.0)
    for (int i = 0; i < m->nodof; ++i) ifv[i] = m->beta_fact * ifv[i];
  return 0;
}

static void copy_blocks(const w3do_model *m, double *dst, const double *src, int which) {
  if (!dst || !src) return;
  const int64_t n = block_offset(m, m->nelblk, which);
  std::memcpy(dst, src, sizeof(double) * (size_t)n);
}
// update.f:75-93
void w3do_update(const w3do_model *m, w3do_state *s) {
  copy_blocks(m, s->history, s->history1, 0);
  copy_blocks(m, s->eps_n, s->eps_n1, 1);
  copy_blocks(m, s->urcs_n, s->urcs_n1, 2);
}
// mnralg.f:1258-1310
void w3do_adaptive_save(const w3do_model *m, w3do_state *s) { copy_blocks(m, s->rot_n, s->rot_n1, 3); }
// mnralg.f:1335-1401
void w3do_adaptive_reset(const w3do_model *m, w3do_state *s) {
  copy_blocks(m, s->history1, s->history, 0);
  copy_blocks(m, s->rot_n1, s->rot_n, 3);
  copy_blocks(m, s->eps_n1, s->eps_n, 1);
  copy_blocks(m, s->urcs_n1, s->urcs_n, 2);
}

}  // extern "C"
