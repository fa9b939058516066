// w3d_mm03_kernel.h -- interface of the mm03 (mises / power-law) Gauss-point kernel and of the
// MAT = 3 instantiations of the strain / internal-force kernel (w3d_mm03.cu).
// mm03 runs like mm10 as three launches: PHASE 1 kinematics -> scratch (R, strain increment,
// energy factor), k_mm03_flags + k_mm03 material update + tangent, PHASE 2 internal force.  The
// material update is its own kernel because the reference's pre-processor takes one decision per
// ELEMENT BLOCK and Gauss point (mm03.f:54-79: a block whose points are all null copies the state
// at n): k_mm03_flags raises one flag per (block, Gauss point) that has a live point, k_mm03 (one
// thread per Gauss point) reads it.
#pragma once
#include <cuda_runtime.h>
#include "w3d_device.cuh"

namespace w3d {

struct Mm03Args {
  const CtaDesc *cta;        // the segment's 16-element groups (as for every other kernel)
  int *blk_active;           // [first slot of the block / 16][8]: any point of the block at this Gauss point not null
  const float *eprops;       // [slot][EPROP]: e, nu, -, h, sig_0, -, n_power, m_power, eps_ref, q1, q2, q3
  const double *hist_n, *urcs_n, *scratch;
  double *hist_n1, *urcs_n1, *cep;
  int *flags;                // [0] material_cut_step, [1] fatal code
  long long ng;
  double dt;
  int step, iter, allow_cut;
  int nonlinear_update;      // iter >= 1 or extrapolated du (rstgp1.f:1666); else linear-elastic [D]
};

struct StressArgs;
cudaError_t mm03_set_shape_table(const double *tab3, size_t bytes);
cudaError_t mm03_prepare_device();
cudaError_t launch_stress_mm03(bool geonl, bool bbar, int phase, int ncta, const StressArgs &a, cudaStream_t st);
cudaError_t launch_mm03(const Mm03Args &a, int ncta, cudaStream_t st);

}  // namespace w3d
