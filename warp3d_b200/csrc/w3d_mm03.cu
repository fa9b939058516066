// w3d_mm03.cu -- material model 3 on the device: k_mm03 (one CTA per element block) and the
// MAT = 3 instantiations of the fused kinematics / internal-force kernel, in their own
// translation unit.  Reference: drive_03_update (rstgp1.f:1418-1990), mm03p / mm03 / cnst3
// (mm03.f); per-point arithmetic in w3d_mm03_core.h.
#include "w3d_stress_kernel.cuh"
#include "w3d_mm03_kernel.h"
#include "w3d_mm03_core.h"

namespace w3d {

// mm03p's block-level decision first (one flag per element block and Gauss point: is any point
// of the block not null?), then one thread per Gauss point.  (A first version gave one CTA to each
// element block so that the decision was a shared-memory reduction: 137 CTAs on the seb deck, a
// stress pass of 0.63 ms against 0.07 ms for mm01 -- far too little parallelism.)
__global__ void __launch_bounds__(TPB) k_mm03_flags(const Mm03Args a) {
  const CtaDesc cd = a.cta[blockIdx.x];
  const int el = threadIdx.x >> 3, gp = threadIdx.x & 7;
  if (el >= cd.nvalid) return;
  const long long g = (long long)(cd.slot0 + el) * 8 + gp;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) s += fabs(a.scratch[(9 + k) * a.ng + g]);
  if (!(s <= 1.0e-10 * 0.002)) a.blk_active[(cd.blk_slot0 / EPC) * 8 + gp] = 1;   // benign race: all write 1
}

__global__ void __launch_bounds__(TPB) k_mm03(const Mm03Args a) {
  const CtaDesc cd = a.cta[blockIdx.x];
  const int el = threadIdx.x >> 3, gp = threadIdx.x & 7;
  if (el >= cd.nvalid) return;
  const long long ng = a.ng;
  const int slot = cd.slot0 + el;
  const long long g = (long long)slot * 8 + gp;
  {
    const double efac = a.scratch[15 * ng + g];
    if (!(efac == efac)) return;   // det F <= 0: step cut already requested by the kinematics
    const bool block_null = a.blk_active[(cd.blk_slot0 / EPC) * 8 + gp] == 0;
    const float *ep = a.eprops + (size_t)slot * EPROP;
    w3d_mm03::Props P;
    P.e = (double)ep[0]; P.nu = (double)ep[1]; P.h_fixed = (double)ep[3]; P.sigyld = (double)ep[4];
    P.n_power = (double)ep[6]; P.m_power = (double)ep[7]; P.eps_ref = (double)ep[8];
    P.q1 = (double)ep[9]; P.q2 = (double)ep[10]; P.q3 = (double)ep[11];
    if (P.n_power > 0.0) P.sigyld = P.sigyld * 0.95;     // trans_factor (rstgp1.f:1460, :1898-1906)
    double cgn[9], uddt[6], h[w3d_mm03::HIST], cgn1[9], h1[w3d_mm03::HIST], D[6][6];
#pragma unroll
    for (int k = 0; k < 9; ++k) cgn[k] = a.urcs_n[k * ng + g];
#pragma unroll
    for (int k = 0; k < 6; ++k) uddt[k] = a.scratch[(9 + k) * ng + g];
#pragma unroll
    for (int k = 0; k < w3d_mm03::HIST; ++k) h[k] = a.hist_n[k * ng + g];
    if (!a.nonlinear_update) {
      // drive_03_update_b / _e (rstgp1.f:1746-1790, :1999-2026): linear-elastic [D], no history
      w3d_mm03::elastic_D(P.e, P.nu, D);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double sk = cgn[k];
#pragma unroll
        for (int m = 0; m < 6; ++m) sk = sk + D[m][k] * uddt[m];
        a.urcs_n1[k * ng + g] = sk;
      }
#pragma unroll
      for (int k = 6; k < 9; ++k) a.urcs_n1[k * ng + g] = 0.0;   // local urcs_blk_n1 is zero-initialised
      int k21 = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) { a.cep[k21 * ng + g] = D[r][c]; ++k21; }
      return;
    }
    const int rc = w3d_mm03::update(a.step, a.iter, P, a.dt, block_null, cgn, uddt, h, cgn1, h1, D);
    if (rc != w3d_mm03::OK) {       // mm03.f:2425-2432: cut the step if allowed, else abort_job
      if (a.allow_cut) atomicOr(&a.flags[0], 1);
      else atomicMax(&a.flags[1], 7);
      return;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) a.urcs_n1[k * ng + g] = cgn1[k];
    if (a.iter > 0) {
#pragma unroll
      for (int k = 0; k < w3d_mm03::HIST; ++k) a.hist_n1[k * ng + g] = h1[k];
    }
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) { a.cep[k * ng + g] = D[r][c]; ++k; }
  }
}

// two launches: flags, then the points (the caller has zeroed blk_active on the same stream)
cudaError_t launch_mm03(const Mm03Args &a, int ncta, cudaStream_t st) {
  if (ncta <= 0) return cudaSuccess;
  if (a.nonlinear_update) k_mm03_flags<<<ncta, TPB, 0, st>>>(a);
  k_mm03<<<ncta, TPB, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t mm03_set_shape_table(const double *tab3, size_t bytes) {
  return cudaMemcpyToSymbol(c_shape, tab3, bytes);
}

cudaError_t mm03_prepare_device() {
  cudaError_t e = stress_opt_in(k_eps_sig_ifv<3, true, true, 1>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, true, false, 1>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, false, true, 1>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, false, false, 1>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, true, true, 2>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, true, false, 2>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, false, true, 2>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<3, false, false, 2>);
  return e;
}

template <int PHASE>
static cudaError_t launch_phase(bool geonl, bool bbar, int grid, const StressArgs &a, cudaStream_t st) {
  if (geonl) {
    if (bbar) return launch_stress_kernel(k_eps_sig_ifv<3, true, true, PHASE>, grid, a, st);
    return launch_stress_kernel(k_eps_sig_ifv<3, true, false, PHASE>, grid, a, st);
  }
  if (bbar) return launch_stress_kernel(k_eps_sig_ifv<3, false, true, PHASE>, grid, a, st);
  return launch_stress_kernel(k_eps_sig_ifv<3, false, false, PHASE>, grid, a, st);
}

// phase 1 / 2 around k_mm03
cudaError_t launch_stress_mm03(bool geonl, bool bbar, int phase, int ncta, const StressArgs &a0, cudaStream_t st) {
  StressArgs a = a0;
  a.ncta = ncta;
  const int grid = (ncta + STRESS_GROUPS - 1) / STRESS_GROUPS;
  if (phase == 1) return launch_phase<1>(geonl, bbar, grid, a, st);
  return launch_phase<2>(geonl, bbar, grid, a, st);
}

}  // namespace w3d
