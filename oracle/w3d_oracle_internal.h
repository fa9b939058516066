/* w3d_oracle_internal.h -- shared declarations for the CPU ORACLE (test infrastructure only). */
#ifndef W3D_ORACLE_INTERNAL_H
#define W3D_ORACLE_INTERNAL_H
#include <cstdint>
#include <cstring>
#include "w3d_oracle.h"

namespace w3do {

// param_def:20-60 (only the values visible on this path)
constexpr int mxvl = W3DO_MXVL;
constexpr int mxelpr = W3DO_MXELPR;
constexpr int mxmtpr = W3DO_MXMTPR;
constexpr int mxndel = 20;
constexpr int mxedof = 60;
constexpr int mxgp = 14;
constexpr int nstr = 6;
constexpr int nstrs = 9;
constexpr int mxutsz = (mxedof * mxedof + mxedof) / 2;

// props / iprops / lprops are three views of one 4-byte table (mod_main.f:41-46)
inline int32_t iprop(const w3do_model *m, int row, int elem /*1-based*/) {
  int32_t v;
  std::memcpy(&v, (const char *)m->props + 4 * ((size_t)(elem - 1) * mxelpr + (row - 1)), 4);
  return v;
}
inline float rprop(const w3do_model *m, int row, int elem) {
  float v;
  std::memcpy(&v, (const char *)m->props + 4 * ((size_t)(elem - 1) * mxelpr + (row - 1)), 4);
  return v;
}
// ifort logicals: .true. has the low bit set (default -1); .false. = 0
inline bool lprop(const w3do_model *m, int row, int elem) { return (iprop(m, row, elem) & 1) != 0; }
inline float matprp(const w3do_model *m, int row, int matno /*1-based*/) {
  return m->matprp[(size_t)(matno - 1) * mxmtpr + (row - 1)];
}

struct BlockInfo {
  int blk, span, felem, elem_type, nnode, ndof, int_order, ngp, mat_type, matnum, totdof,
      hist_size, utsz;
  bool geonl, bbar;
};
BlockInfo block_info(const w3do_model *m, int blk);
int64_t block_offset(const w3do_model *m, int blk, int which);

// indexing macros for block-local arrays (leading dimension mxvl, i is 0-based, others 1-based)
#define GAM(i, r, c) gama[(((c)-1) * 3 + ((r)-1)) * w3do::mxvl + (i)]
#define VOL(i, j, k) vol[(((k)-1) * 8 + ((j)-1)) * w3do::mxvl + (i)]
#define F3(f, i, r, c) f[(((c)-1) * 3 + ((r)-1)) * w3do::mxvl + (i)]
#define Q6(q, i, r, c) q[(((c)-1) * 6 + ((r)-1)) * w3do::mxvl + (i)]
#define BM(b, i, j, k) b[(((k)-1) * w3do::mxedof + ((j)-1)) * w3do::mxvl + (i)]
#define V2(a, i, j) a[((j)-1) * w3do::mxvl + (i)]

struct BlockPtrs {
  const int32_t *edest, *cdest;
  double *history, *history1, *eps_n, *eps_n1, *urcs_n, *urcs_n1, *rot_n, *rot_n1, *cep, *estiff,
      *element_vol;
};
} // namespace w3do
#include <vector>
namespace w3do {
std::vector<BlockPtrs> all_block_ptrs(const w3do_model *m, const w3do_state *s);

// kinematics (w3d_oracle.cpp)
void quad2(int gpn, double &xi, double &eta, double &zeta, double &w);
void deriv2(double xi, double eta, double zeta, double *nxi, double *neta, double *nzeta);
void make_cp_icp(int n, int *cp, int *dcp, int *icp);
void jacob1(int span, double *jac, double *dj, double *gama, const double *nxi, const double *neta,
            const double *nzeta, const double *ce, int nnode);
void vol_terms(const double *in_jacob, const double *dj, double *vol, const double *nxi,
               const double *neta, const double *nzeta, double *volume, int span);
void vol_avg(double *vol, const double *volume, int span);
void tcomp1(int span, double *theta, const double *nxi, const double *neta, const double *nzeta,
            const double *gama, const double *ue, int nnode);
void fcomp1(int span, double *f, double *df, const double *theta, int &error);
void rtcmp1(int span, const double *f, double *r);
void getrm1(int span, double *q, const double *r, int opt);
void qmply1(int span, const double *q, const double *m1, double *m2);
void blcmp1(int span, double *b, const double *gama, const double *nxi, const double *neta,
            const double *nzeta, int nnode);
void bmod(double *b, const double *vol, int span, double eps_stab);
void gtlsn1(int span, int nnode, const double *due, double *deps, const double *gama,
            const double *nxi, const double *neta, const double *nzeta, const double *vol_block,
            bool bbar, double eps_bbar, double *b);

// mm01 (w3d_oracle_mm01.cpp)
void mm01(int span, int step, const double *ym_n1, const double *nu_n1, const double *beta,
          const double *hprime_n1, const bool *lnelas, const double *yld_n1, double *cgn,
          double *cgn1, const double *deps, double *history, double *history1, double *rtse,
          const double *dtemps, const double *ym_n, const double *nu_n, int *fatal);
void cnst1(int span, double *cep, const double *rtsg, const double *nu, const double *e,
           const double *kn1, const double *hprime, const double *beta, const double *ldt,
           const double *dstates);

}  // namespace w3do
#endif
