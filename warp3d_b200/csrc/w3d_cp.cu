// w3d_cp.cu -- mm10 (crystal plasticity) Gauss-point update kernel.  Replaces drive_10_update
// (rstgp1.f:4327-4513) -> mm10 (mm10_a.f:29): one thread per Gauss point runs all crystals of the
// point through the shared core (w3d_mm10_core.h): Schmid-tensor set-up, stress predictor and
// 7x7 Newton update with line search and adaptive sub-incrementation, consistent tangent by
// Schur complement, plastic rotation update and the output/history terms.
// Inputs from the PHASE 1 kinematics launch (scratch: R(n+1), strain increment, energy factor).
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "w3d_cp_kernel.h"
#include "w3d_mm10_core.h"
#include "w3d_mm10_fast.h"

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

// ------------------------------------------------------------------------------------------------
// Fast path (w3d_mm10_fast.h): one crystal type, <= 12 slip systems.  Thread = Gauss point, 256
// threads per CTA, one CTA per SM; no per-thread stack -- the slip tables are operands from
// __constant__ memory, B6 / C B6 / B3 and the vectors that must survive the Jacobian sit in a
// per-thread column of shared memory ([word][thread]: conflict-free).
// ------------------------------------------------------------------------------------------------
// CTA shape, measured on the 1 M-element mm10 workload (stress pass, 3 launches): 128 threads in
// lock step, 2 CTAs per SM 45.6 ms; 256 threads, 1 CTA 46.3 ms; 64 threads, 4 CTAs 58.2 ms (four
// instruction streams per SM); without the lock step 56.6 / 62.1 / 70.4 ms.  (Round-1 kernel:
// 54.1 ms.)  With geometric hardening the column is 12 words longer: one CTA of 128 per SM.
// the solver loop's lock step: one barrier per trip, loop ends when no thread of the CTA needs more
struct CtaSync {
  // barrier.red WITHOUT .aligned (what __syncthreads_or would emit): the lanes of a warp reach
  // the loop top at different times -- those already done skip the trip's body -- and an aligned
  // barrier executed by a diverged warp is undefined (it deadlocked).  Every thread of the CTA
  // arrives exactly once per trip; nobody leaves the kernel before the last trip.
  static __device__ __forceinline__ bool any(bool p) {
    unsigned r;
    __syncwarp();      // every lane of every warp comes through here once per trip
    asm volatile(
        "{\n\t"
        ".reg .pred pin, pout;\n\t"
        "setp.ne.u32 pin, %1, 0;\n\t"
        "barrier.cta.red.or.pred pout, 0, pin;\n\t"
        "selp.u32 %0, 1, 0, pout;\n\t"
        "}"
        : "=r"(r)
        : "r"((unsigned)(p ? 1u : 0u))
        : "memory");
    return r != 0;
  }
};
#ifndef W3D_MM10F_THREADS
#define W3D_MM10F_THREADS 128
#endif
#ifndef W3D_MM10F_LOCKSTEP
#define W3D_MM10F_LOCKSTEP 1
#endif
#if W3D_MM10F_LOCKSTEP
typedef CtaSync Mm10fSync;
#else
typedef w3d_mm10f::NoSync Mm10fSync;
#endif
template <bool TAUL, bool SINGLE>
struct Mm10fShape {
  static constexpr int THREADS = TAUL ? TPB : W3D_MM10F_THREADS;
  // geometric hardening: 112 words per thread (w3d_mm10_fast.h, tl_word), still two CTAs per SM
  static constexpr int MIN_CTAS = TAUL ? 2 : (W3D_MM10F_THREADS > 128 ? 1 : W3D_MM10F_THREADS > 64 ? 2 : 4);
  static constexpr int SMEM_BYTES = w3d_mm10f::work_doubles(TAUL) * THREADS * (int)sizeof(double);
};

// two CTAs of either shape per SM: 228 KB of shared memory less 1 KB reserved per CTA
static_assert(Mm10fShape<false, true>::MIN_CTAS * (Mm10fShape<false, true>::SMEM_BYTES + 1024) <= 228 * 1024,
              "k_mm10f workspace column too long for its CTAs per SM");
static_assert(Mm10fShape<true, true>::MIN_CTAS * (Mm10fShape<true, true>::SMEM_BYTES + 1024) <= 228 * 1024,
              "k_mm10f workspace column with tau_l too long for two CTAs per SM");

// SINGLE: one crystal per element -- the tangent columns stored ahead by the Newton passes go
// straight to their place (w3d_mm10_fast.h, crystal_update_fast: SINGLE)
// RATE_E: rate exponent harden_n - 1 fixed at compile time (19: harden_n 20 of verification/test82
// and test74), or -1 = read from the crystal table at run time
template <bool TAUL, bool SINGLE, int RATE_E>
__global__ void __launch_bounds__(Mm10fShape<TAUL, SINGLE>::THREADS, Mm10fShape<TAUL, SINGLE>::MIN_CTAS) k_mm10f(const CpArgs a) {
  extern __shared__ double sWork[];   // [word][thread]
  constexpr int MM10F_THREADS = Mm10fShape<TAUL, SINGLE>::THREADS;
  // a CTA takes THREADS / 128 whole 16-element groups, or (64 threads) half a group
  constexpr int GROUPS = MM10F_THREADS >= TPB ? MM10F_THREADS / TPB : 1;
  constexpr int SPLIT = MM10F_THREADS >= TPB ? 1 : TPB / MM10F_THREADS;
  const int grp = threadIdx.x / TPB;
  const int tid = threadIdx.x % TPB + (blockIdx.x % SPLIT) * MM10F_THREADS;
  const int idesc = (blockIdx.x / SPLIT) * GROUPS + grp;
  const int el = tid >> 3, gp = tid & 7;
  const int pos = idesc * EPC + el;          // position in the schedule (a.order)
  bool have;
  int slot;
  if (a.order) {
    have = pos < a.n_el;
    slot = a.order[have ? pos : a.n_el - 1];
  } else {
    const bool desc = idesc < a.ncta;
    const CtaDesc cd = a.cta[desc ? idesc : a.ncta - 1];
    have = desc && el < cd.nvalid;
    slot = cd.slot0 + (el < cd.nvalid ? el : cd.nvalid - 1);
  }
  const long long ng = a.ng;
  const long long g = (long long)slot * 8 + gp;
  const double efac = a.scratch[15 * ng + g];
  // padding lanes and points whose det F was not positive (NaN marker: the kinematics already
  // asked for a step cut) take no part in the update, only in the loop's barriers
  const bool active = have && (efac == efac);
  if (!active) {
    for (int c = 0; c < a.ncrystals; ++c)
      while (Mm10fSync::any(false)) {}
    return;
  }
  double cgn1[9];
  const w3d_mm10f::LayoutF L = w3d_mm10f::LayoutF::make(a.nslip);
  const w3d_mm10f::FProps *P =
      reinterpret_cast<const w3d_mm10f::FProps *>(a.props) + (size_t)slot * a.ncrystals;
  const w3d_mm10f::CHistF hn{a.hist_n + g, ng};
  const w3d_mm10f::HistF h1{a.hist_n1 + g, ng};
  const w3d_mm10f::WorkF W{sWork + threadIdx.x, MM10F_THREADS};
  const w3d_mm10f::PointInF in{a.scratch + g, a.scratch + 9 * ng + g, a.urcs_n + g, ng};
  unsigned trips = 0;
  const bool cut = w3d_mm10f::gauss_point_update_fast<TAUL, Mm10fSync, SINGLE, RATE_E>(a.step, a.ncrystals, P, L, hn, h1, W, in, a.dt,
                                                                     a.iter_0_extrapolate_off != 0, a.asymmetric != 0, cgn1,
                                                                     &trips);
  if (a.trips) atomicMax(&a.trips[pos], (int)trips);
  if (cut) {
    atomicOr(&a.flags[0], 1);
    return;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) a.urcs_n1[k * ng + g] = cgn1[k];
}

// ---- Mm10Schedule: see w3d_cp_kernel.h
struct Mm10Schedule {
  int n_el = 0;
  int *order[2] = {nullptr, nullptr};   // element slots, current order in order[cur]
  int *trips = nullptr, *keys_out = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int cur = 0;
  bool have_trips = false;
};
void mm10_schedule_destroy(Mm10Schedule *s) {
  if (!s) return;
  cudaFree(s->order[0]); cudaFree(s->order[1]); cudaFree(s->trips); cudaFree(s->keys_out); cudaFree(s->tmp);
  delete s;
}
cudaError_t mm10_schedule_create(const CtaDesc *host_cta, int ncta, Mm10Schedule **out) {
  *out = nullptr;
  std::vector<int> slots;
  for (int i = 0; i < ncta; ++i)
    for (int e = 0; e < host_cta[i].nvalid; ++e) slots.push_back(host_cta[i].slot0 + e);
  if (slots.empty()) return cudaSuccess;
  Mm10Schedule *s = new Mm10Schedule;
  s->n_el = (int)slots.size();
  const size_t bytes = sizeof(int) * slots.size();
  cudaError_t e = cudaMalloc((void **)&s->order[0], bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->order[1], bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->trips, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->keys_out, bytes);
  if (e == cudaSuccess)
    e = cub::DeviceRadixSort::SortPairsDescending(nullptr, s->tmp_bytes, s->trips, s->keys_out, s->order[0], s->order[1],
                                                  s->n_el, 0, 8);
  if (e == cudaSuccess) e = cudaMalloc(&s->tmp, s->tmp_bytes ? s->tmp_bytes : 1);
  if (e == cudaSuccess) e = cudaMemcpy(s->order[0], slots.data(), bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { mm10_schedule_destroy(s); return e; }
  *out = s;
  return cudaSuccess;
}

bool mm10_fast_eligible(const double *crystal_types, int n_crystal_types) {
  if (getenv("W3D_MM10_GENERIC")) return false;      // A/B switch: force the general kernel
  if (n_crystal_types != 1) return false;
  const int nslip = (int)crystal_types[337], num_hard = (int)crystal_types[338], h_type = (int)crystal_types[339];
  const int miter = (int)crystal_types[341];      // the loop's iteration counter holds 8 bits
  return nslip >= 1 && nslip <= w3d_mm10f::NSF && num_hard == 1 && h_type == 1 && miter <= 250;
}
int mm10_fast_props_doubles() { return (int)(sizeof(w3d_mm10f::FProps) / sizeof(double)); }
bool mm10_fast_setup_props_host(const double *ct, const double *angles, int angle_type, double *out) {
  w3d_mm10f::FProps P;
  if (!w3d_mm10f::setup_fprops(ct, angles, angle_type, P)) return false;
  std::memcpy(out, &P, sizeof P);
  return true;
}
int mm10_fast_consts_bytes() { return (int)sizeof(w3d_mm10f::CrystalConsts); }
void mm10_fast_setup_consts_host(const double *ct, int angle_type, void *out) {
  w3d_mm10f::CrystalConsts K;
  std::memset(&K, 0, sizeof K);
  w3d_mm10f::setup_consts(ct, K);
  K.angle_type = angle_type;
  std::memcpy(out, &K, sizeof K);
}
cudaError_t launch_mm10_fast(const CpArgs &a, int ncta, const void *consts, bool taul, Mm10Schedule *sched,
                             cudaStream_t st) {
  // the crystal-type table goes to __constant__ memory in stream order before every launch: two
  // engines of one process (sequentially, as one handle per GPU prescribes) never see each
  // other's table
  cudaError_t e = cudaMemcpyToSymbolAsync(w3d_mm10f::c_mm10f, consts, sizeof(w3d_mm10f::CrystalConsts), 0,
                                          cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  CpArgs b = a;
  b.ncta = ncta;
  b.order = nullptr; b.trips = nullptr; b.n_el = 0;
  if (sched) {
    if (sched->have_trips) {
      // keys: trip counts by position of the previous pass (saturated at 255: 8 bits)
      e = cub::DeviceRadixSort::SortPairsDescending(sched->tmp, sched->tmp_bytes, sched->trips, sched->keys_out,
                                                    sched->order[sched->cur], sched->order[1 - sched->cur],
                                                    sched->n_el, 0, 8, st);
      if (e != cudaSuccess) return e;
      sched->cur = 1 - sched->cur;
    }
    e = cudaMemsetAsync(sched->trips, 0, sizeof(int) * (size_t)sched->n_el, st);
    if (e != cudaSuccess) return e;
    sched->have_trips = true;
    b.order = sched->order[sched->cur]; b.trips = sched->trips; b.n_el = sched->n_el;
    ncta = (sched->n_el + EPC - 1) / EPC;
  }
#ifdef W3D_MM10F_HIST
  {
    int *hp = a.flags + 2;
    cudaMemcpyToSymbolAsync(w3d_mm10f::w3df_hist, &hp, sizeof hp, 0, cudaMemcpyHostToDevice, st);
  }
#endif
  const bool single = (a.ncrystals == 1);
  const bool e19 = (static_cast<const w3d_mm10f::CrystalConsts *>(consts)->rate_ie == 19);
  if (taul) {
    constexpr int T = Mm10fShape<true, true>::THREADS;
    if (single && e19) k_mm10f<true, true, 19><<<ncta, T, Mm10fShape<true, true>::SMEM_BYTES, st>>>(b);
    else if (single) k_mm10f<true, true, -1><<<ncta, T, Mm10fShape<true, true>::SMEM_BYTES, st>>>(b);
    else k_mm10f<true, false, -1><<<ncta, T, Mm10fShape<true, false>::SMEM_BYTES, st>>>(b);
  } else {
    constexpr int T = Mm10fShape<false, true>::THREADS, S = Mm10fShape<false, true>::SMEM_BYTES;
    const int grid = T >= TPB ? (ncta + T / TPB - 1) / (T / TPB) : ncta * (TPB / T);
    if (single && e19) k_mm10f<false, true, 19><<<grid, T, S, st>>>(b);
    else if (single) k_mm10f<false, true, -1><<<grid, T, S, st>>>(b);
    else k_mm10f<false, false, -1><<<grid, T, S, st>>>(b);
  }
  return cudaGetLastError();
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
  auto opt_in = [&](auto kernel, int bytes) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  };
  opt_in(k_mm10f<false, false, -1>, Mm10fShape<false, false>::SMEM_BYTES);
  opt_in(k_mm10f<false, true, -1>, Mm10fShape<false, true>::SMEM_BYTES);
  opt_in(k_mm10f<false, true, 19>, Mm10fShape<false, true>::SMEM_BYTES);
  opt_in(k_mm10f<true, false, -1>, Mm10fShape<true, false>::SMEM_BYTES);
  opt_in(k_mm10f<true, true, -1>, Mm10fShape<true, true>::SMEM_BYTES);
  opt_in(k_mm10f<true, true, 19>, Mm10fShape<true, true>::SMEM_BYTES);
  return e;
}
cudaError_t launch_mm10(const CpArgs &a, int ncta, cudaStream_t st) {
  return a.nslip <= 12 ? launch_mm10_ns<12>(a, ncta, st) : launch_mm10_ns<48>(a, ncta, st);
}

}  // namespace w3d
