// w3d_oracle_residual.cpp -- CPU oracle (test infrastructure only, see w3d_oracle.h): literal,
// sequential restatement of the structure residual and its norms:
//   upres_iter_0         src/upres.f:13-77
//   upres                src/upres.f:121-226
//   cvtest residual loop src/cvtest.f:62-78
// Sums run over dof 1..nodof in order, exactly as the reference does.  NULL vectors stand for
// zero vectors (statics: no mass, no contact, no multipoint constraints).
#include <cmath>
#include <cstdint>

namespace {
inline double at(const double *v, int i) { return v ? v[i] : 0.0; }
}  // namespace

extern "C" {

// upres.f:57-72.  Returns norm_load.
double w3do_upres_iter_0(int32_t nodof, double dt, double nbeta, const int32_t *cstmap,
                         const double *pbar, const double *ifv, const double *mdiag,
                         const double *du, const double *contact_force, double *res) {
  const double nfac = 1.0 / (nbeta * dt * dt);
  double sum = 0.0;
  for (int i = 0; i < nodof; ++i) {
    if (cstmap[i] == 0)
      res[i] = (pbar[i] - at(mdiag, i) * du[i] * nfac + at(contact_force, i)) - ifv[i];
    else
      res[i] = 0.0;
    sum = sum + res[i] * res[i];
  }
  return std::sqrt(sum);
}

// upres.f:147-226.  out3 = {mgload, sum_loads}; returns num_term_loads.
int64_t w3do_upres(int32_t nodof, double dt, double nbeta, const int32_t *cstmap,
                   const double *pbar, const double *load, const double *ifv, const double *mdiag,
                   const double *du, const double *velocity, const double *accel,
                   const double *contact_force, const double *d_lagrange_forces, double *res,
                   double *out2) {
  const double nfac = 1.0 / (nbeta * dt * dt);
  const double nfac1 = (1.0 - 2.0 * nbeta) / (2.0 * nbeta);
  const double nfac2 = 1.0 / (nbeta * dt);
  if (d_lagrange_forces) {
    for (int i = 0; i < nodof; ++i)
      res[i] = (pbar[i] + at(contact_force, i) + d_lagrange_forces[i] - at(mdiag, i) * du[i] * nfac) - ifv[i];
  } else {
    for (int i = 0; i < nodof; ++i)
      res[i] = (pbar[i] + at(contact_force, i) - at(mdiag, i) * du[i] * nfac) - ifv[i];
  }
  double mgload = 0.0, sum_loads = 0.0;
  int64_t num_term_loads = 0;
  for (int i = 0; i < nodof; ++i) {
    const double accel_n1 = du[i] * nfac - nfac1 * at(accel, i) - nfac2 * at(velocity, i);
    const double inertia_contact = std::fabs(accel_n1) * at(mdiag, i) + std::fabs(at(contact_force, i));
    double term_load;
    if (cstmap[i] == 0) {
      term_load = inertia_contact + std::fabs(at(load, i));
    } else {
      term_load = inertia_contact + std::fabs(ifv[i]);
      res[i] = 0.0;
    }
    if (term_load != 0.0) {
      mgload = mgload + term_load * term_load;
      num_term_loads += 1;
      sum_loads = sum_loads + std::fabs(term_load);
    }
  }
  out2[0] = std::sqrt(mgload);
  out2[1] = sum_loads;
  return num_term_loads;
}

// cvtest.f:62-78.  out2 = {norm of the residual, res(res_max_dof)}; returns res_max_dof (1-based).
int32_t w3do_cvtest_residual(int32_t nodof, const double *res, const double *d_lagrange_forces,
                             double *out2) {
  double res_max = -100.0, sum = 0.0;
  int32_t res_max_dof = 0;
  for (int j = 0; j < nodof; ++j) {
    double resforce = res[j];
    if (d_lagrange_forces && d_lagrange_forces[j] != 0.0) resforce = 0.0;
    sum = sum + resforce * resforce;
    if (std::fabs(resforce) > res_max) {
      res_max = std::fabs(resforce);
      res_max_dof = j + 1;
    }
  }
  out2[0] = std::sqrt(sum);
  out2[1] = res_max_dof >= 1 ? res[res_max_dof - 1] : 0.0;
  return res_max_dof;
}

}  // extern "C"
