// w3d_cp.cu -- mm10 (crystal plasticity) Gauss-point update kernel.  Replaces drive_10_update
// (rstgp1.f:4327-4513) -> mm10 (mm10_a.f:29): one thread per Gauss point runs all crystals of the
// point through the shared core (w3d_mm10_core.h): Schmid-tensor set-up, stress predictor and
// 7x7 Newton update with line search and adaptive sub-incrementation, consistent tangent by
// Schur complement, plastic rotation update and the output/history terms.
// Inputs from the PHASE 1 kinematics launch (scratch: R(n+1), strain increment, energy factor).
#include <cstring>

#include "w3d_cp_kernel.h"
#include "w3d_mm10_core.h"

namespace w3d {

using namespace w3d_mm10;

#ifndef MM10_GROUPS
// 16-element groups (CTA descriptors) per CTA: CTA = MM10_GROUPS x 128 threads.  2 = one 256-thread
// CTA per SM instead of two of 128: the warps share more of the instruction fetch (measured 4.16 ->
// 3.80 ms on the 64k-element check)
#define MM10_GROUPS 2
#endif
#ifndef MM10_MIN_CTAS
#define MM10_MIN_CTAS (2 / MM10_GROUPS)   // 8 resident warps per SM at 255 registers per thread
#endif
// property records staged per CTA when they fit (1 448 B per element and crystal for 12 systems)
constexpr int MM10_PROPS_SMEM_LIMIT = 64 * 1024;

template <int NS>
__global__ void __launch_bounds__(TPB * MM10_GROUPS, MM10_MIN_CTAS) k_mm10(const CpArgs a) {
  extern __shared__ double sPropsAll[];   // [group][element of the group][crystal][Props<NS>] or unused
  const int grp = threadIdx.x / TPB;
  const int tid = threadIdx.x % TPB;
  const int idesc = blockIdx.x * MM10_GROUPS + grp;
  const bool have = idesc < a.ncta;
  const CtaDesc cd = a.cta[have ? idesc : a.ncta - 1];
  const int el = tid >> 3, gp = tid & 7;
  constexpr int REC = (int)(sizeof(Props<NS>) / sizeof(double));
  const int erec = REC * a.ncrystals;                    // doubles per element
  const bool staged = (size_t)erec * EPC * MM10_GROUPS * sizeof(double) <= (size_t)MM10_PROPS_SMEM_LIMIT;
  double *sProps = sPropsAll + (size_t)grp * erec * EPC;
  if (staged) {
    // the 8 lanes of an element read the same record over and over in the slip-system loops:
    // one coalesced copy into shared memory, then conflict-free broadcast reads
    const double *src = a.props + (size_t)cd.slot0 * erec;
    const int n = cd.nvalid * erec;
    if (have)
      for (int k = tid; k < n; k += TPB) sProps[k] = src[k];
    __syncthreads();
  }
  if (!have || el >= cd.nvalid) return;
  const int slot = cd.slot0 + el;
  const long long ng = a.ng;
  const long long g = (long long)slot * 8 + gp;
  const double efac = a.scratch[15 * ng + g];
  if (!(efac == efac)) return;   // det F <= 0: step cut already requested by the kinematics
  double Rur[9], uddt[6], cgn[9], cgn1[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Rur[k] = a.scratch[k * ng + g];
#pragma unroll
  for (int k = 0; k < 6; ++k) uddt[k] = a.scratch[(9 + k) * ng + g];
#pragma unroll
  for (int k = 0; k < 9; ++k) cgn[k] = a.urcs_n[k * ng + g];
  const Layout L = Layout::make(a.nslip, a.num_hard);
  const Props<NS> *P = staged ? reinterpret_cast<const Props<NS> *>(sProps + (size_t)el * erec)
                              : reinterpret_cast<const Props<NS> *>(a.props + (size_t)slot * erec);
  const CHist hn{a.hist_n + g, ng};
  const Hist h1{a.hist_n1 + g, ng};
  // no slip store on the device: the core recomputes the rotated Schmid data where needed
  const bool cut = gauss_point_update<NS>(a.step, a.ncrystals, P, L, hn, h1, Rur, uddt, a.dt,
                                          a.iter_0_extrapolate_off != 0, a.asymmetric != 0, cgn, cgn1,
                                          nullptr, 0);
  if (cut) {
    atomicOr(&a.flags[0], 1);
    return;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) a.urcs_n1[k * ng + g] = cgn1[k];
}

int mm10_props_doubles(int nslip) {
  return (int)((nslip <= 12 ? sizeof(Props<12>) : sizeof(Props<48>)) / sizeof(double));
}
int mm10_hist_size(int nslip, int num_hard, int ncrystals) { return Layout::make(nslip, num_hard).size(ncrystals); }
void mm10_layout_offsets(int nslip, int num_hard, int *o_gradfe, int *o_Rp0) {
  const Layout L = Layout::make(nslip, num_hard);
  *o_gradfe = L.o_gradfe();
  *o_Rp0 = L.c_Rp(0);
}
void mm10_setup_props_host(const double *ctd, const double *angles, int angle_type, double *props_out) {
  CrystalType ct;
  const size_t nd = 36 + 144 + 144 + 13;
  std::memcpy(&ct, ctd, nd * sizeof(double));
  const double *ti = ctd + nd;
  ct.nslip = (int)ti[0]; ct.num_hard = (int)ti[1]; ct.h_type = (int)ti[2]; ct.tang_calc = (int)ti[3];
  ct.miter = (int)ti[4]; ct.solver = (int)ti[5]; ct.alter_mode = (int)ti[6];
  if (ct.nslip <= 12) {
    Props<12> P;
    std::memset(&P, 0, sizeof P);
    setup_props<12>(ct, angles, angle_type, P);
    std::memcpy(props_out, &P, sizeof P);
  } else {
    Props<48> P;
    std::memset(&P, 0, sizeof P);
    setup_props<48>(ct, angles, angle_type, P);
    std::memcpy(props_out, &P, sizeof P);
  }
}
template <int NS>
static cudaError_t launch_mm10_ns(const CpArgs &a, int ncta, cudaStream_t st) {
  const size_t need = sizeof(Props<NS>) * (size_t)a.ncrystals * EPC * MM10_GROUPS;
  const int smem = need <= (size_t)MM10_PROPS_SMEM_LIMIT ? (int)need : 0;
  CpArgs b = a;
  b.ncta = ncta;
  k_mm10<NS><<<(ncta + MM10_GROUPS - 1) / MM10_GROUPS, TPB * MM10_GROUPS, smem, st>>>(b);
  return cudaGetLastError();
}
cudaError_t mm10_prepare_device() {
  cudaError_t e = cudaFuncSetAttribute(k_mm10<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM10_PROPS_SMEM_LIMIT);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_mm10<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, MM10_PROPS_SMEM_LIMIT);
  return e;
}
cudaError_t launch_mm10(const CpArgs &a, int ncta, cudaStream_t st) {
  return a.nslip <= 12 ? launch_mm10_ns<12>(a, ncta, st) : launch_mm10_ns<48>(a, ncta, st);
}

}  // namespace w3d
