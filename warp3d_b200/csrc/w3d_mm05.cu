// w3d_mm05.cu -- the MAT = 5 (cyclic plasticity, mm05) instantiations of the fused strain /
// stress / internal-force kernel, in their own translation unit (compile time: the four
// instantiations carry both mm05 options).
//
// Reproducibility note (measured).  The generalized_plasticity option finds
// the elastic fraction of a step that starts outside the yield surface by a golden-section
// search for the MINIMUM of the yield function with |b - a| < 1e-10 (mm05.f:3290-3350).  Near a
// minimum the function is flat to second order, so the comparisons fx1 < fx2 turn into rounding
// noise once the bracket is below ~1e-8: the reference's own eta, and everything integrated from
// it, is reproducible only to ~1e-8 there (CPU restatement with vs without FMA contraction:
// 2e-8 in the plastic multiplier, 5e-10 in the tangent, 6e-11 in the stresses).  Compiling this
// unit with -fmad=false was tried and does not help: last-bit differences in the strain increment
// handed to the model are enough.  Points that do not take that branch agree to 1e-10.
#include "w3d_stress_kernel.cuh"

namespace w3d {

cudaError_t mm05_set_shape_table(const double *tab3, size_t bytes) {
  return cudaMemcpyToSymbol(c_shape, tab3, bytes);
}

cudaError_t mm05_prepare_device() {
  cudaError_t e = stress_opt_in(k_eps_sig_ifv<5, true, true, 0>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<5, true, false, 0>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<5, false, true, 0>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<5, false, false, 0>);
  return e;
}

cudaError_t launch_stress_mm05(bool geonl, bool bbar, int ncta, const StressArgs &a0, cudaStream_t st) {
  StressArgs a = a0;
  a.ncta = ncta;
  const int grid = (ncta + STRESS_GROUPS - 1) / STRESS_GROUPS;
  if (geonl) {
    if (bbar) return launch_stress_kernel(k_eps_sig_ifv<5, true, true, 0>, grid, a, st);
    return launch_stress_kernel(k_eps_sig_ifv<5, true, false, 0>, grid, a, st);
  }
  if (bbar) return launch_stress_kernel(k_eps_sig_ifv<5, false, true, 0>, grid, a, st);
  return launch_stress_kernel(k_eps_sig_ifv<5, false, false, 0>, grid, a, st);
}

}  // namespace w3d
