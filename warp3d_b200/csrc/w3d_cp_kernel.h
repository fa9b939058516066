// w3d_cp_kernel.h -- interface of the crystal-plasticity (mm10) Gauss-point kernel, compiled in
// its own translation unit (w3d_cp.cu).  FMA contraction is on: the local Newton iteration is
// contractive, parity with the CPU build of the shared core holds at 1e-10 (tests/test_gpu_mm10.py).
#pragma once
#include <cuda_runtime.h>
#include "w3d_device.cuh"

namespace w3d {

struct CpArgs {
  const CtaDesc *cta;
  const double *props;        // [slot][ncrystals][props_doubles]  (w3d_mm10::Props<12> or <48>)
  const double *hist_n, *urcs_n, *scratch;
  double *hist_n1, *urcs_n1;
  int *flags;
  long long ng;
  double dt;
  int step, iter_0_extrapolate_off, asymmetric, ncrystals, nslip, num_hard;
  int ncta;                   // number of CTA descriptors (set by launch_mm10)
  // k_mm10f with a schedule (Mm10Schedule): the element slots in the order they are taken, 16 per
  // 128 threads; trips[position] receives the largest solver-loop trip count of the element
  const int *order;
  int *trips;
  int n_el;
};

// Order in which k_mm10f takes the elements: by the trip count of the previous stress pass,
// longest first (stable: equal counts keep their order, the first pass runs in block order).  A
// CTA walks the solver loop in lock step and is as slow as its slowest point, so points of like
// cost belong together; the result of a point does not depend on where it is computed.
struct Mm10Schedule;
cudaError_t mm10_schedule_create(const CtaDesc *host_cta, int ncta, Mm10Schedule **out);
void mm10_schedule_destroy(Mm10Schedule *s);

// Two instantiations of the shared core: up to 12 slip systems (fcc, bcc: the fast path, property
// records staged in shared memory) and up to 48 (bcc48 and the other long tables of
// mod_crystals.f; records read from global memory).
int mm10_props_doubles(int nslip);
// host: one property record from a crystal-type record (setup_mm10_rknstr, rknstr.f:1568-2027)
void mm10_setup_props_host(const double *crystal_type_rec, const double *angles, int angle_type,
                           double *props_out);
int mm10_hist_size(int nslip, int num_hard, int ncrystals);
void mm10_layout_offsets(int nslip, int num_hard, int *o_gradfe, int *o_Rp0);
// ---- fast path (w3d_mm10_fast.h): one crystal type with <= 12 slip systems, Voce hardening.
// eligible: decided once at create; props: 128 doubles per (element, crystal); consts: the
// crystal-type table that launch_mm10_fast uploads to __constant__ memory before the kernel
bool mm10_fast_eligible(const double *crystal_types, int n_crystal_types);
int mm10_fast_props_doubles();
bool mm10_fast_setup_props_host(const double *crystal_type_rec, const double *angles, int angle_type,
                                double *props_out);
int mm10_fast_consts_bytes();
void mm10_fast_setup_consts_host(const double *crystal_type_rec, int angle_type, void *consts_out);
cudaError_t launch_mm10_fast(const CpArgs &a, int ncta, const void *consts, bool taul, Mm10Schedule *sched,
                             cudaStream_t st);
// shared-memory opt-in of every instantiation on the current device (w3d_engine_create)
cudaError_t mm10_prepare_device();
cudaError_t launch_mm10(const CpArgs &a, int ncta, cudaStream_t st);

}  // namespace w3d
