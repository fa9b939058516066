/*
 * w3d_oracle_tanstf.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): element tangent stiffness pass.
 * Follows reference src/: tanstf.f:17-95 (tanstf), :119-341 (do_nlek_block), :642-717
 * (dptstf_blocks), dptstf.f:16-215, rktstf.f:15-296, gptns1.f:16-342 (gptns1, gptns1_a),
 * :837-878 (drive_01_cnst), :1174-1283 (extract_D_symmetric), :1628-1718 (bdbtgen),
 * :1735-1970 (ctran1, ctran1_qbar), kg1.f:16-160 (kgstiff, symmetric assembly).
 * PARITY STATUS: see w3d_oracle.h (mm01, mm05 and mm10 pinned end to end by reference known
 * answers at 1e-5; unpinned at function granularity).
 */
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "w3d_oracle_internal.h"
#include "w3d_oracle_mm10.h"

namespace w3do {

struct EkWork {
  std::vector<double> ce, ce_0, ue, due;            // (mxvl, mxedof)
  std::vector<double> det_jac_block;                // (mxvl, mxgp)
  std::vector<double> nxi, neta, nzeta;             // (mxndel, mxgp)
  std::vector<double> gama_block;                   // (mxvl,3,3,mxgp)
  std::vector<double> jac_block;                    // (mxvl,3,3)
  std::vector<double> vol_block, volume_block;      // (mxvl,8,3), (mxvl)
  std::vector<double> b_block, bd_block;            // (mxvl,mxedof,nstr)
  std::vector<double> urcs_blk_n1, rot_blk_n1;      // (mxvl,nstrs,mxgp), (mxvl,9,mxgp)
  std::vector<double> cep, qn1;                     // (mxvl,6,6)
  std::vector<double> cs_blk_n1;                    // (mxvl,nstr)
  std::vector<double> ek_symm;                      // (span,utsz)
  std::vector<int> cp, icp;                         // (mxedof), (mxutsz,2)
  double weights[mxgp];
  EkWork()
      : ce(mxvl * mxedof), ce_0(mxvl * mxedof), ue(mxvl * mxedof), due(mxvl * mxedof),
        det_jac_block(mxvl * mxgp), nxi(mxndel * mxgp), neta(mxndel * mxgp), nzeta(mxndel * mxgp),
        gama_block(mxvl * 9 * mxgp), jac_block(mxvl * 9), vol_block(mxvl * 24),
        volume_block(mxvl), b_block(mxvl * mxedof * nstr), bd_block(mxvl * mxedof * nstr),
        urcs_blk_n1(mxvl * nstrs * mxgp), rot_blk_n1(mxvl * 9 * mxgp), cep(mxvl * 36),
        qn1(mxvl * 36), cs_blk_n1(mxvl * nstr), cp(mxedof), icp(mxutsz * 2) {}
};

// tanstf_gastr: global (nprm,ngp,span) -> local (mxvl,nprm,ngp)
static void gastr(double *mlocal, const double *mglobal, int ngp, int nprm, int span) {
  for (int k = 0; k < ngp; ++k)
    for (int j = 0; j < nprm; ++j)
      for (int i = 0; i < span; ++i)
        mlocal[((size_t)k * nprm + j) * mxvl + i] = mglobal[((size_t)i * ngp + k) * nprm + j];
}

// extract_D_symmetric (gptns1.f:1174-1283): global cep (21,span,ngp) * weight * det_jac -> 6x6
static void extract_D_symmetric(int span, int gpn, double weight, const double *det_jac /*gp*/,
                                const double *gbl_cep, double *cep) {
  static const int rr[21] = {1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6};
  static const int cc[21] = {1, 1, 2, 1, 2, 3, 1, 2, 3, 4, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 6};
  for (int k = 1; k <= 21; ++k)
    for (int ielem = 0; ielem < span; ++ielem) {
      const size_t sloc = (size_t)21 * span * (gpn - 1) + (size_t)21 * ielem;
      const double f = weight * det_jac[ielem];
      const double v = f * gbl_cep[sloc + (k - 1)];
      Q6(cep, ielem, rr[k - 1], cc[k - 1]) = v;
      Q6(cep, ielem, cc[k - 1], rr[k - 1]) = v;
    }
}

// ctran1 (gptns1.f:1735-1900) + ctran1_qbar (1905-1967)
static void ctran1(int span, double *cep, const double *qn1, const double *cs, bool qbar,
                   const double *dj, double w) {
  std::vector<double> tc(mxvl * 36);
  const double two = 2.0, half = 0.5;
  for (int j = 1; j <= 6; ++j)
    for (int i = 0; i < span; ++i)
      for (int k = 1; k <= 6; ++k)
        Q6(tc, i, j, k) = (Q6(qn1, i, j, 1) * Q6(cep, i, 1, k) + Q6(qn1, i, j, 2) * Q6(cep, i, 2, k) +
                           Q6(qn1, i, j, 3) * Q6(cep, i, 3, k) + Q6(qn1, i, j, 4) * Q6(cep, i, 4, k) +
                           Q6(qn1, i, j, 5) * Q6(cep, i, 5, k) + Q6(qn1, i, j, 6) * Q6(cep, i, 6, k));
  for (int j = 1; j <= 6; ++j)
    for (int i = 0; i < span; ++i)
      for (int k = 1; k <= 6; ++k)
        Q6(cep, i, j, k) = Q6(tc, i, j, 1) * Q6(qn1, i, k, 1) + Q6(tc, i, j, 2) * Q6(qn1, i, k, 2) +
                           Q6(tc, i, j, 3) * Q6(qn1, i, k, 3) + Q6(tc, i, j, 4) * Q6(qn1, i, k, 4) +
                           Q6(tc, i, j, 5) * Q6(qn1, i, k, 5) + Q6(tc, i, j, 6) * Q6(qn1, i, k, 6);
  if (!qbar) return;
  for (int i = 0; i < span; ++i) {
    const double wf = dj[i] * w;
    const double halfw = half * wf;
    const double xx = V2(cs, i, 1), yy = V2(cs, i, 2), zz = V2(cs, i, 3), xy = V2(cs, i, 4),
                 yz = V2(cs, i, 5), xz = V2(cs, i, 6);
    Q6(cep, i, 1, 1) = Q6(cep, i, 1, 1) - two * xx * wf;
    Q6(cep, i, 2, 2) = Q6(cep, i, 2, 2) - two * yy * wf;
    Q6(cep, i, 3, 3) = Q6(cep, i, 3, 3) - two * zz * wf;
    Q6(cep, i, 4, 1) = Q6(cep, i, 4, 1) - xy * wf;
    Q6(cep, i, 6, 1) = Q6(cep, i, 6, 1) - xz * wf;
    Q6(cep, i, 4, 2) = Q6(cep, i, 4, 2) - xy * wf;
    Q6(cep, i, 5, 2) = Q6(cep, i, 5, 2) - yz * wf;
    Q6(cep, i, 5, 3) = Q6(cep, i, 5, 3) - yz * wf;
    Q6(cep, i, 6, 3) = Q6(cep, i, 6, 3) - xz * wf;
    Q6(cep, i, 4, 4) = Q6(cep, i, 4, 4) - halfw * (xx + yy);
    Q6(cep, i, 5, 5) = Q6(cep, i, 5, 5) - halfw * (yy + zz);
    Q6(cep, i, 6, 6) = Q6(cep, i, 6, 6) - halfw * (xx + zz);
    Q6(cep, i, 5, 4) = Q6(cep, i, 5, 4) - wf * xz;
    Q6(cep, i, 6, 4) = Q6(cep, i, 6, 4) - wf * yz;
    Q6(cep, i, 6, 5) = Q6(cep, i, 6, 5) - wf * xy;
    Q6(cep, i, 1, 4) = Q6(cep, i, 4, 1);
    Q6(cep, i, 1, 6) = Q6(cep, i, 6, 1);
    Q6(cep, i, 2, 4) = Q6(cep, i, 4, 2);
    Q6(cep, i, 2, 5) = Q6(cep, i, 5, 2);
    Q6(cep, i, 3, 5) = Q6(cep, i, 5, 3);
    Q6(cep, i, 3, 6) = Q6(cep, i, 6, 3);
    Q6(cep, i, 4, 5) = Q6(cep, i, 5, 4);
    Q6(cep, i, 4, 6) = Q6(cep, i, 6, 4);
    Q6(cep, i, 5, 6) = Q6(cep, i, 6, 5);
  }
}

// bdbtgen (gptns1.f:1628-1718)
static void bdbtgen(int span, const int *icp, const double *b, double *bd, const double *d,
                    double *ek_symm /*(span,utsz)*/, int utsz, int totdof) {
  for (int j = 1; j <= totdof; ++j)
    for (int i = 0; i < span; ++i)
      for (int k = 1; k <= 6; ++k)
        BM(bd, i, j, k) = Q6(d, i, 1, k) * BM(b, i, j, 1) + Q6(d, i, 2, k) * BM(b, i, j, 2) +
                          Q6(d, i, 3, k) * BM(b, i, j, 3) + Q6(d, i, 4, k) * BM(b, i, j, 4) +
                          Q6(d, i, 5, k) * BM(b, i, j, 5) + Q6(d, i, 6, k) * BM(b, i, j, 6);
  for (int j = 1; j <= utsz; ++j) {
    const int row = icp[j - 1];
    const int col = icp[mxutsz + j - 1];
    for (int i = 0; i < span; ++i)
      ek_symm[(size_t)(j - 1) * span + i] =
          ek_symm[(size_t)(j - 1) * span + i] + BM(b, i, col, 1) * BM(bd, i, row, 1) +
          BM(b, i, col, 2) * BM(bd, i, row, 2) + BM(b, i, col, 3) * BM(bd, i, row, 3) +
          BM(b, i, col, 4) * BM(bd, i, row, 4) + BM(b, i, col, 5) * BM(bd, i, row, 5) +
          BM(b, i, col, 6) * BM(bd, i, row, 6);
  }
}

// kgstiff (kg1.f:16-110), symmetric assembly
static void kgstiff(int span, const int *cp, const int *icp, const double *gama, const double *nxi,
                    const double *neta, const double *nzeta, int nnode, const double *sig,
                    const double *dj, double w, double *ek_symm, const double *vol, bool bbar) {
  constexpr int mxnusz = (mxndel * mxndel + mxndel) / 2;
  std::vector<double> gtg((size_t)mxvl * mxnusz), gxi(mxvl * mxndel), geta(mxvl * mxndel),
      gzeta(mxvl * mxndel);
  if (bbar) {
    for (int j = 1; j <= nnode; ++j)
      for (int i = 0; i < span; ++i) {
        V2(gxi, i, j) = VOL(i, j, 1);
        V2(geta, i, j) = VOL(i, j, 2);
        V2(gzeta, i, j) = VOL(i, j, 3);
      }
  } else {
    for (int j = 1; j <= nnode; ++j)
      for (int i = 0; i < span; ++i) {
        V2(gxi, i, j) = GAM(i, 1, 1) * nxi[j - 1] + GAM(i, 1, 2) * neta[j - 1] + GAM(i, 1, 3) * nzeta[j - 1];
        V2(geta, i, j) = GAM(i, 2, 1) * nxi[j - 1] + GAM(i, 2, 2) * neta[j - 1] + GAM(i, 2, 3) * nzeta[j - 1];
        V2(gzeta, i, j) = GAM(i, 3, 1) * nxi[j - 1] + GAM(i, 3, 2) * neta[j - 1] + GAM(i, 3, 3) * nzeta[j - 1];
      }
  }
  const int nterm = cp[nnode - 1] + nnode;
  for (int j = 1; j <= nterm; ++j) {
    const int a = icp[j - 1], bq = icp[mxutsz + j - 1];
    for (int i = 0; i < span; ++i)
      V2(gtg, i, j) = (V2(gxi, i, a) * V2(gxi, i, bq) * V2(sig, i, 1) +
                       V2(geta, i, a) * V2(geta, i, bq) * V2(sig, i, 2) +
                       V2(gzeta, i, a) * V2(gzeta, i, bq) * V2(sig, i, 3) +
                       (V2(gxi, i, a) * V2(geta, i, bq) + V2(gxi, i, bq) * V2(geta, i, a)) * V2(sig, i, 4) +
                       (V2(geta, i, a) * V2(gzeta, i, bq) + V2(geta, i, bq) * V2(gzeta, i, a)) * V2(sig, i, 5) +
                       (V2(gxi, i, a) * V2(gzeta, i, bq) + V2(gxi, i, bq) * V2(gzeta, i, a)) * V2(sig, i, 6)) *
                      dj[i] * w;
  }
  for (int enode = 1; enode <= nnode; ++enode) {
    const int cp1 = cp[enode - 1];
    const int cp2 = cp[nnode + enode - 1] + nnode;
    const int cp3 = cp[2 * nnode + enode - 1] + 2 * nnode;
    for (int j = 1; j <= enode; ++j)
      for (int i = 0; i < span; ++i) {
        ek_symm[(size_t)(cp1 + j - 1) * span + i] += V2(gtg, i, cp1 + j);
        ek_symm[(size_t)(cp2 + j - 1) * span + i] += V2(gtg, i, cp1 + j);
        ek_symm[(size_t)(cp3 + j - 1) * span + i] += V2(gtg, i, cp1 + j);
      }
  }
}

// do_nlek_block (tanstf.f:119-341) -> dptstf, dptstf_blocks, rktstf, gptns1
static void do_nlek_block(const w3do_model *m, const BlockPtrs &bp, int blk, int step, int iter,
                          const double *u, const double *du, int *fatal) {
  (void)step; (void)iter;
  const BlockInfo bi = block_info(m, blk);
  const int span = bi.span, nnode = bi.nnode, totdof = bi.totdof, ngp = bi.ngp, utsz = bi.utsz;
  if (bi.elem_type != 2 || nnode != 8 || ngp != 8 || bi.ndof != 3 || span > mxvl) {
    *fatal = 4;
    return;
  }
  if (bi.mat_type != 1 && bi.mat_type != 3 && bi.mat_type != 5 && bi.mat_type != 10) {   // drive_03_cnst = extract_D_symmetric (gptns1.f:939-950)
    *fatal = 3;
    return;
  }
  double *ek = bp.estiff;
  for (size_t k = 0; k < (size_t)utsz * span; ++k) ek[k] = 0.0;  // tanstf.f:228
  EkWork lw;
  make_cp_icp(mxedof, lw.cp.data(), nullptr, lw.icp.data());  // local_cp / local_icp copies

  // dptstf (dptstf.f:64-156)
  {
    int k = 1;
    for (int j = 1; j <= nnode; ++j) {
      for (int i = 0; i < span; ++i)
        for (int t = 0; t < 3; ++t) {
          const double v = m->c[bp.cdest[(size_t)i * totdof + (k - 1 + t)] - 1];
          V2(lw.ce, i, k + t) = v;
          V2(lw.ce_0, i, k + t) = v;
        }
      k = k + 3;
    }
    if (bi.geonl) {
      for (int j = 1; j <= totdof; ++j)
        for (int i = 0; i < span; ++i) {
          const int dof = bp.edest[(size_t)i * totdof + (j - 1)];
          V2(lw.ue, i, j) = u[dof - 1];
          V2(lw.due, i, j) = du[dof - 1];
        }
      for (int j = 1; j <= totdof; ++j)
        for (int i = 0; i < span; ++i) V2(lw.ce, i, j) = V2(lw.ce, i, j) + V2(lw.ue, i, j) + V2(lw.due, i, j);
    }
  }
  // dptstf_blocks (tanstf.f:642-717)
  if (bi.geonl) gastr(lw.rot_blk_n1.data(), bp.rot_n1, ngp, 9, span);
  gastr(lw.urcs_blk_n1.data(), bp.urcs_n1, ngp, nstrs, span);

  // rktstf (rktstf.f:15-220)
  lw.ek_symm.assign((size_t)span * utsz, 0.0);
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
    jacob1(span, lw.jac_block.data(), &lw.det_jac_block[(gpn - 1) * mxvl],
           &lw.gama_block[(size_t)(gpn - 1) * 9 * mxvl], nxi, neta, nzeta, lw.ce.data(), nnode);
    if (bi.bbar)
      vol_terms(&lw.gama_block[(size_t)(gpn - 1) * 9 * mxvl], &lw.det_jac_block[(gpn - 1) * mxvl],
                lw.vol_block.data(), nxi, neta, nzeta, lw.volume_block.data(), span);
  }
  if (bi.bbar) vol_avg(lw.vol_block.data(), lw.volume_block.data(), span);
  std::fill(lw.b_block.begin(), lw.b_block.end(), 0.0);
  std::fill(lw.bd_block.begin(), lw.bd_block.end(), 0.0);

  for (int gpn = 1; gpn <= ngp; ++gpn) {
    // gptns1 (gptns1.f:16-342)
    const double weight = lw.weights[gpn - 1];
    const double *nxi = &lw.nxi[(gpn - 1) * mxndel], *neta = &lw.neta[(gpn - 1) * mxndel],
                 *nzeta = &lw.nzeta[(gpn - 1) * mxndel];
    const double *gama = &lw.gama_block[(size_t)(gpn - 1) * 9 * mxvl];
    const double *dj = &lw.det_jac_block[(gpn - 1) * mxvl];
    // gptns1_a
    if (bi.geonl) {
      getrm1(span, lw.qn1.data(), &lw.rot_blk_n1[(size_t)(gpn - 1) * 9 * mxvl], 2);
      qmply1(span, lw.qn1.data(), &lw.urcs_blk_n1[(size_t)(gpn - 1) * nstrs * mxvl],
             lw.cs_blk_n1.data());
    }
    blcmp1(span, lw.b_block.data(), gama, nxi, neta, nzeta, nnode);
    if (bi.bbar) bmod(lw.b_block.data(), lw.vol_block.data(), span, m->eps_bbar);
    if (bi.mat_type == 10) {
      // drive_10_cnst (gptns1.f:1371-1410): [D] = history1(1:36) (6x6 column-major) * w * detJ
      for (int i = 0; i < span; ++i) {
        const double f = weight * dj[i];
        const double *h = bp.history1 + ((size_t)i * ngp + (gpn - 1)) * bi.hist_size;
        for (int c = 1; c <= 6; ++c)
          for (int r = 1; r <= 6; ++r) Q6(lw.cep, i, r, c) = h[(c - 1) * 6 + (r - 1)] * f;
      }
    } else {
      // drive_01_cnst / drive_05_cnst -> extract_D_symmetric
      extract_D_symmetric(span, gpn, weight, dj, bp.cep, lw.cep.data());
    }
    if (bi.geonl)
      ctran1(span, lw.cep.data(), lw.qn1.data(), lw.cs_blk_n1.data(), m->qbar_flag != 0, dj, weight);
    bdbtgen(span, lw.icp.data(), lw.b_block.data(), lw.bd_block.data(), lw.cep.data(),
            lw.ek_symm.data(), utsz, totdof);
    if (bi.geonl)
      kgstiff(span, lw.cp.data(), lw.icp.data(), gama, nxi, neta, nzeta, nnode, lw.cs_blk_n1.data(),
              dj, weight, lw.ek_symm.data(), lw.vol_block.data(), bi.bbar);
  }
  // rktstf_do_transpose (rktstf.f:221-296): (span,utsz) -> (utsz,span)
  for (int j = 0; j < span; ++j)
    for (int i = 0; i < utsz; ++i) ek[(size_t)j * utsz + i] = lw.ek_symm[(size_t)i * span + j];
  if (m->beta_fact != 1.0)
    for (size_t k = 0; k < (size_t)utsz * span; ++k) ek[k] = m->beta_fact * ek[k];
}

}  // namespace w3do

using namespace w3do;

extern "C" int32_t w3do_tanstf(const w3do_model *m, w3do_state *s, int32_t step, int32_t iter,
                               const double *u, const double *du, int32_t nthreads) {
  const int nelblk = m->nelblk;
  std::vector<int> fatals(nelblk, 0);
  const std::vector<BlockPtrs> bps = all_block_ptrs(m, s);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int blk = 0; blk < nelblk; ++blk) do_nlek_block(m, bps[blk], blk, step, iter, u, du, &fatals[blk]);
  for (int blk = 0; blk < nelblk; ++blk)
    if (fatals[blk]) return fatals[blk];
  return 0;
}
