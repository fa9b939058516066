// w3d_residual_data.h -- device vectors of the structure-residual path (w3d_residual.cuh)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace w3d {

constexpr int RES_TPB = 256;
constexpr int RES_MAX_CTAS = 1184;  // 148 SMs x 8 (all 64 warps of an SM): partials of the two-stage, fixed-order reduction

struct ResidualPartial {            // one per CTA of k_residual, then one for the whole vector
  double sum_res2, mgload2, sum_loads, max_abs;
  long long num_term_loads;
  int max_dof;                      // 1-based dof of max_abs (lowest such dof), 0 = none
  int pad;
  double res_at_max;                // res(max_dof), signed: filled by k_residual_finish only
};

struct ResidualData {
  bool ready = false;
  int32_t *d_cstmap = nullptr;
  // nullptr = the reference vector is identically zero (statics, no contact, no MPCs)
  double *d_pbar = nullptr, *d_load = nullptr, *d_mdiag = nullptr, *d_vel = nullptr,
         *d_acc = nullptr, *d_cf = nullptr, *d_dlag = nullptr;
  double *d_res = nullptr;
  ResidualPartial *d_part = nullptr;   // RES_MAX_CTAS + 1
  ResidualPartial *h_part = nullptr;   // pinned, 1
  void release() {
    cudaFree(d_cstmap); cudaFree(d_pbar); cudaFree(d_load); cudaFree(d_mdiag); cudaFree(d_vel);
    cudaFree(d_acc); cudaFree(d_cf); cudaFree(d_dlag); cudaFree(d_res); cudaFree(d_part);
    if (h_part) cudaFreeHost(h_part);
    *this = ResidualData();
  }
};

}  // namespace w3d
