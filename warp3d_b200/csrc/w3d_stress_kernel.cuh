// w3d_stress_kernel.cuh -- fused strain recovery / material update / internal force kernel.
//
// One launch replaces, for every owned hex8 block of one segment (same material model, geonl and
// B-bar flags), the reference chain
//   dupstr (dupstr.f:82-88) + dupstr_blocked (drive_eps_sig_internal_forces.f:1699-1754)
//   rknstr_geonl / rknstr_sm_displ (rknstr.f:455-614) with vol_terms/vol_avg
//   rstgp1 / rstgp2 (rstgp1.f:19-598): gtmat1, gtlsn1, qmply1, rstgp1_update_strains
//   drive_01_update (rstgp1.f:614-984): mm01 + cnst1 + rstgp1_store_cep, or linear-elastic iter 0
//   rstgp1_b energy integration (rstgp1.f:5189-5241)
//   rplstr (rplstr.f:15-93)
//   rknifv_std + gpifv1 (rknifv.f:141-215, gpifv1.f:16-165)
//   addifv_standard (addifv.f:86-113)
// Thread = (element, Gauss point).  HBM traffic per Gauss point: read eps_n 48 + urcs_n 72 +
// hist_n 88, write eps_n1 48 + urcs_n1 72 + rot_n1 72 + hist_n1 88 + cep 168 (B), plus the
// element's index maps / coordinates / displacements and 24 force atomics.
#pragma once
#include "w3d_device.cuh"
#include "w3d_mm01.cuh"
#include "w3d_mm10_core.h"
#include "w3d_mm05_core.h"

namespace w3d {

struct StressArgs {
  const CtaDesc *cta;
  const int32_t *edest, *cdest;   // [slot][24], 1-based
  const float *eprops;            // [slot][EPROP]: e, nu, beta, hprime, yld, lnelas
  const double *c, *u, *du;       // (nodof)
  double *ifv;                    // (nodof), atomically accumulated
  const double *eps_n, *urcs_n, *hist_n;
  double *eps_n1, *urcs_n1, *rot_n1, *hist_n1, *cep, *elem_vol;
  double *partials;               // [ncta][4]: internal energy, plastic work, sum |f|, -
  int *flags;                     // [0] material_cut_step, [1] fatal code
  long long ng;                   // component stride of the state arrays (Gauss-point slots)
  double eps_bbar, beta_fact;
  int step, iter;
  int nonlinear_update;           // iter >= 1 or extrapolated_du (rstgp1.f:835)
  int update_coords;              // geonl and not (step 1, iter 0)  (:1722-1754)
  int adaptive;                   // adaptive_flag and step > 1 (rstgp1.f:136-145)
  int hist_size;
  // crystal plasticity (mm10) runs as three launches: PHASE 1 kinematics -> scratch,
  // k_mm10 material update, PHASE 2 internal force + lattice-curvature fit
  double *scratch;                // [16][ng]: R(n+1) 9 (column-major), uddt 6, energy factor 1
  int cp_ncrystals, cp_o_gradfe, cp_o_Rp;   // history offsets (w3d_mm10::Layout)
  int ncta;                       // number of CTA descriptors of the segment (set by the launcher)
};

// parent-space shape function derivatives of the 8 Gauss points: [dir][node][gp], filled by the
// host with the reference's quad2 / deriv2 tables (getgpts.f:642-718, derivs.f:283-343).
// static: one copy per translation unit (w3d_engine.cu and w3d_mm05.cu each upload theirs)
static __constant__ double c_shape[3][8][8];

// MAT = 5 instantiations live in w3d_mm05.cu
cudaError_t mm05_set_shape_table(const double *tab3, size_t bytes);
cudaError_t mm05_prepare_device();   // shared-memory opt-in of the MAT = 5 instantiations on the current device
cudaError_t launch_stress_mm05(bool geonl, bool bbar, int ncta, const struct StressArgs &a, cudaStream_t st);

// A CTA takes STRESS_GROUPS consecutive 16-element groups (CTA descriptors): the warps of a CTA
// start together and walk the same ~70 KB of straight-line code, so a larger CTA shares more of
// the instruction fetch (the kernel's second largest stall), at the same 16 warps per SM.
// Measured on the 2 M-element mm01 mesh: 1 group (128 threads) 5.01 ms, 2 groups 4.71 ms.
#ifndef STRESS_GROUPS
#define STRESS_GROUPS 2
#endif
#ifndef STRESS_MIN_CTAS
#define STRESS_MIN_CTAS (4 / STRESS_GROUPS)   // register cap 128/thread -> 16 warps per SM
#endif
constexpr int STPB = TPB * STRESS_GROUPS;   // threads per CTA of the stress kernel
constexpr int SEPC = EPC * STRESS_GROUPS;   // elements per CTA
constexpr int NSTATE_IN = 6 + 9 + 14;       // eps_n, urcs_n, hist_n (11 mm01 / 14 mm05) per Gauss point
// shape table 192, c/u/du 3 x SEPC x 24, per-warp sums (+1 pad), state at n NSTATE_IN x STPB,
// element property records SEPC x EPROP floats, mid-step and n+1 coordinates 2 x SEPC x 24
constexpr int STRESS_SMEM_BYTES =
    (int)sizeof(double) * (192 + 3 * SEPC * 24 + (STPB / 32) * 3 + 1 + NSTATE_IN * STPB + SEPC * (EPROP / 2) +
                           2 * SEPC * 24);

// Dynamic shared memory above 48 KB is an opt-in per kernel AND per device: every instantiation is
// opted in by w3d_engine_create after cudaSetDevice (stress_prepare_device below), so a second
// engine on another device of the same process is covered and no process-wide cache is needed.
inline cudaError_t stress_opt_in(void (*k)(const struct StressArgs)) {
  if (STRESS_SMEM_BYTES <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STRESS_SMEM_BYTES);
}
inline cudaError_t launch_stress_kernel(void (*k)(const struct StressArgs), int grid,
                                        const struct StressArgs &a, cudaStream_t st) {
  k<<<grid, STPB, STRESS_SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

// 8-byte asynchronous global -> shared copy (LDGSTS): state at n streams in while the
// kinematics run, so no warp waits on HBM latency when the material update starts
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Jacobian with the node loop kept rolled: shape derivatives and coordinates are re-read from
// shared memory each trip, so only the 9 accumulators are live (register pressure, not smem
// bandwidth, limits this kernel).  MODE 0: X = c ; 1: X = (c+u)+du/2 ; 2: X = (c+u)+du.
template <bool EXACT, int MODE>
__device__ __forceinline__ void jacobian_smem(const double (*sN)[8][8], int gp, const double *C,
                                              const double *U, const double *DU,
                                              double (&G)[3][3], double &dj) {
  double J[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
#pragma unroll 4
  for (int j = 0; j < 8; ++j) {
    const double nx = sN[0][j][gp], ne = sN[1][j][gp], nz = sN[2][j][gp];
    double x = C[j], y = C[8 + j], z = C[16 + j];
    if (MODE == 1) {
      x = (x + U[j]) + 0.5 * DU[j]; y = (y + U[8 + j]) + 0.5 * DU[8 + j]; z = (z + U[16 + j]) + 0.5 * DU[16 + j];
    } else if (MODE == 2) {
      x = (x + U[j]) + DU[j]; y = (y + U[8 + j]) + DU[8 + j]; z = (z + U[16 + j]) + DU[16 + j];
    }
    if (EXACT) {     // contraction convention of the noise-amplifying chain (w3d_device.cuh)
      J[0][0] = xf(nx, x, J[0][0]); J[0][1] = xf(nx, y, J[0][1]); J[0][2] = xf(nx, z, J[0][2]);
      J[1][0] = xf(ne, x, J[1][0]); J[1][1] = xf(ne, y, J[1][1]); J[1][2] = xf(ne, z, J[1][2]);
      J[2][0] = xf(nz, x, J[2][0]); J[2][1] = xf(nz, y, J[2][1]); J[2][2] = xf(nz, z, J[2][2]);
    } else {
      J[0][0] = fma(nx, x, J[0][0]); J[0][1] = fma(nx, y, J[0][1]); J[0][2] = fma(nx, z, J[0][2]);
      J[1][0] = fma(ne, x, J[1][0]); J[1][1] = fma(ne, y, J[1][1]); J[1][2] = fma(ne, z, J[1][2]);
      J[2][0] = fma(nz, x, J[2][0]); J[2][1] = fma(nz, y, J[2][1]); J[2][2] = fma(nz, z, J[2][2]);
    }
  }
  if (EXACT) {
    const double j1 = xf(J[1][1], J[2][2], -xm(J[1][2], J[2][1]));
    const double j2 = xf(J[1][0], J[2][2], -xm(J[1][2], J[2][0]));
    const double j3 = xf(J[1][0], J[2][1], -xm(J[1][1], J[2][0]));
    dj = xf(J[0][2], j3, xf(-J[0][1], j2, xm(J[0][0], j1)));
    const double rdj = __ddiv_rn(1.0, dj);      // one reciprocal, nine products (convention)
    G[0][0] = xm(j1, rdj);
    G[1][0] = xm(-j2, rdj);
    G[2][0] = xm(j3, rdj);
    G[0][1] = xm(xf(J[2][1], J[0][2], -xm(J[0][1], J[2][2])), rdj);
    G[1][1] = xm(xf(J[0][0], J[2][2], -xm(J[2][0], J[0][2])), rdj);
    G[2][1] = xm(xf(J[0][1], J[2][0], -xm(J[0][0], J[2][1])), rdj);
    G[0][2] = xm(xf(J[0][1], J[1][2], -xm(J[0][2], J[1][1])), rdj);
    G[1][2] = xm(xf(J[0][2], J[1][0], -xm(J[0][0], J[1][2])), rdj);
    G[2][2] = xm(xf(J[0][0], J[1][1], -xm(J[0][1], J[1][0])), rdj);
  } else {
    const double j1 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double j2 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
    const double j3 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    dj = J[0][0] * j1 - J[0][1] * j2 + J[0][2] * j3;
    const double r = 1.0 / dj;
    G[0][0] = j1 * r; G[1][0] = -j2 * r; G[2][0] = j3 * r;
    G[0][1] = (J[2][1] * J[0][2] - J[0][1] * J[2][2]) * r;
    G[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * r;
    G[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    G[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    G[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    G[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
  }
}

// strain increment  B(G) * due  with the node loop rolled (gtlsn1.f:25-82 without the B-bar part)
__device__ __forceinline__ void strain_increment(const double (*sN)[8][8], int gp,
                                                 const double (&G)[3][3], const double *DU,
                                                 double (&d)[6]) {
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = 0.0;
#pragma unroll 4
  for (int j = 0; j < 8; ++j) {
    const double nx = sN[0][j][gp], ne = sN[1][j][gp], nz = sN[2][j][gp];
    const double bx = G[0][0] * nx + G[0][1] * ne + G[0][2] * nz;
    const double by = G[1][0] * nx + G[1][1] * ne + G[1][2] * nz;
    const double bz = G[2][0] * nx + G[2][1] * ne + G[2][2] * nz;
    const double du = DU[j], dv = DU[8 + j], dw = DU[16 + j];
    d[0] = fma(bx, du, d[0]);
    d[1] = fma(by, dv, d[1]);
    d[2] = fma(bz, dw, d[2]);
    d[3] = fma(by, du, fma(bx, dv, d[3]));
    d[4] = fma(bz, dv, fma(by, dw, d[4]));
    d[5] = fma(bz, du, fma(bx, dw, d[5]));
  }
}

// Jacobian matrix only, reference statement order (local_work%jac of jacob1) -- used by the
// mm10 lattice-curvature fit, which inverts it by cofactors (mm10_a_invert_33)
__device__ __forceinline__ void jacobian_matrix_exact(const double (*sN)[8][8], int gp,
                                                      const double *C, double (&J)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const double nx = sN[0][j][gp], ne = sN[1][j][gp], nz = sN[2][j][gp];
    const double x = C[j], y = C[8 + j], z = C[16 + j];
    J[0][0] = xf(nx, x, J[0][0]); J[0][1] = xf(nx, y, J[0][1]); J[0][2] = xf(nx, z, J[0][2]);
    J[1][0] = xf(ne, x, J[1][0]); J[1][1] = xf(ne, y, J[1][1]); J[1][2] = xf(ne, z, J[1][2]);
    J[2][0] = xf(nz, x, J[2][0]); J[2][1] = xf(nz, y, J[2][1]); J[2][2] = xf(nz, z, J[2][2]);
  }
}

template <int MAT, bool GEONL, bool BBAR, int PHASE = 0>
__global__ void __launch_bounds__(STPB, STRESS_MIN_CTAS) k_eps_sig_ifv(const StressArgs a) {
  // dynamic shared memory (STRESS_SMEM_BYTES; above the 48 KB static limit for STRESS_GROUPS = 2)
  extern __shared__ double dsm[];
  double (*sN)[8][8] = reinterpret_cast<double (*)[8][8]>(dsm);                       // [3][8][8]
  double (*sC)[24] = reinterpret_cast<double (*)[24]>(dsm + 192);                     // [SEPC][24]
  double (*sU)[24] = sC + SEPC, (*sDU)[24] = sC + 2 * SEPC;
  double (*sRed)[3] = reinterpret_cast<double (*)[3]>(dsm + 192 + 3 * SEPC * 24);     // [STPB/32][3]
  // state at n, [component][thread]: conflict free
  double (*sSt)[STPB] = reinterpret_cast<double (*)[STPB]>(dsm + 192 + 3 * SEPC * 24 + (STPB / 32) * 3 + 1);
  // the element's property record (EPROP floats), streamed with the state: 8 bytes per lane
  double (*sEp)[EPROP / 2] = reinterpret_cast<double (*)[EPROP / 2]>(&sSt[NSTATE_IN][0]);
  // nodal coordinates on the mid-step and n+1 configurations, formed once per element by the
  // gathering lanes instead of by each of its 8 Gauss points inside the Jacobian loops
  double (*sXm)[24] = reinterpret_cast<double (*)[24]>(&sEp[SEPC][0]);
  double (*sX1)[24] = sXm + SEPC;

  const int tid = threadIdx.x;
  for (int k = tid; k < 192; k += STPB) (&sN[0][0][0])[k] = (&c_shape[0][0][0])[k];
  __syncthreads();   // the shape table is the only CTA-wide exchange before the final partial sums:
                     // a warp owns 4 whole elements, their staging below needs a warp barrier only

  const int idesc = blockIdx.x * STRESS_GROUPS + tid / TPB;   // this thread's 16-element group
  const bool have = idesc < a.ncta;
  const CtaDesc cd = a.cta[have ? idesc : a.ncta - 1];
  const int el = tid >> 3, gp = tid & 7;                       // element within the CTA
  const int elg = el % EPC;                                    // ... within its group
  const bool valid = have && elg < cd.nvalid;
  const int slot = cd.slot0 + (elg < cd.nvalid ? elg : cd.nvalid - 1);
  const long long g = (long long)slot * 8 + gp;
  const long long ng = a.ng;

  // ---- gather, first half: lane gp fetches the dof / coordinate numbers of node gp of its
  //      element (dofs node, 8+node, 16+node); the values are fetched after the state stream
  //      below has been issued, so the index round trip overlaps it
  int dofs[3], cidx[3];
  {
    const int32_t *ed = a.edest + (size_t)slot * 24;
    const int32_t *cdst = a.cdest + (size_t)slot * 24;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dofs[k] = ed[k * 8 + gp] - 1;
      cidx[k] = cdst[k * 8 + gp] - 1;
    }
  }

  // ---- start streaming this Gauss point's state at n into shared memory
  if (PHASE != 2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) cp_async8(&sSt[k][tid], a.eps_n + k * ng + g);
  }
  if (MAT == 1 || MAT == 5) {
#pragma unroll
    for (int k = 0; k < 9; ++k) cp_async8(&sSt[6 + k][tid], a.urcs_n + k * ng + g);
#pragma unroll
    for (int k = 0; k < (MAT == 1 ? 11 : 14); ++k) cp_async8(&sSt[15 + k][tid], a.hist_n + k * ng + g);
    static_assert(EPROP == 16, "one 8-byte piece of the property record per Gauss-point lane");
    cp_async8(&sEp[el][gp], reinterpret_cast<const double *>(a.eprops + (size_t)slot * EPROP) + gp);
  }
  cp_async_commit();

  // ---- gather, second half: the 9 values of the lane's node
  {
    double vu[3], vdu[3], vc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vu[k] = a.u[dofs[k]];
      vdu[k] = a.du[dofs[k]];
      vc[k] = a.c[cidx[k]];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sU[el][k * 8 + gp] = vu[k];
      sDU[el][k * 8 + gp] = vdu[k];
      sC[el][k * 8 + gp] = vc[k];
      if (GEONL) {
        const double cu = vc[k] + vu[k];
        sXm[el][k * 8 + gp] = fma(0.5, vdu[k], cu);
        sX1[el][k * 8 + gp] = cu + vdu[k];
      }
    }
  }
  __syncwarp();

  const double *C = sC[el], *U = sU[el], *DU = sDU[el];
  const double beta_b = (1.0 - a.eps_bbar) * (1.0 / 3.0);   // bmod.f: beta = (1-eps)/3
  const double weight = 1.0;                                 // quad2: w = 1

  bool error = false;
  double R1[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double efac;          // detF * detJ0 * beta_fact * w  (rstgp1_b)
  double uddt[6];

  if (PHASE == 2) {
    // kinematics were done by the PHASE 1 launch: reload R(n+1) and the energy factor
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) R1[r][c] = a.scratch[(c * 3 + r) * ng + g];
    efac = a.scratch[15 * ng + g];
    error = !(efac == efac);   // NaN marks a point whose det F was not positive
#pragma unroll
    for (int k = 0; k < 6; ++k) uddt[k] = 0.0;
  } else if (GEONL) {
    // ---- gtmat1 (gtmat1.f:17-190): F at n+1 and n+1/2 from the undeformed configuration
    double Rh[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    {
      double G0[3][3], dj0;
      jacobian_smem<true, 0>(sN, gp, C, U, DU, G0, dj0);
      double t1[9], th[9];   // theta at n+1, n+1/2 (tcomp1, gtmat1.f:522-584)
#pragma unroll
      for (int k = 0; k < 9; ++k) { t1[k] = 0.0; th[k] = 0.0; }
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const double nx = sN[0][j][gp], ne = sN[1][j][gp], nz = sN[2][j][gp];
        const double tx = xdot3(G0[0][0], nx, G0[0][1], ne, G0[0][2], nz);
        const double ty = xdot3(G0[1][0], nx, G0[1][1], ne, G0[1][2], nz);
        const double tz = xdot3(G0[2][0], nx, G0[2][1], ne, G0[2][2], nz);
        const double u0 = U[j], v0 = U[8 + j], w0 = U[16 + j];
        const double du = DU[j], dv = DU[8 + j], dw = DU[16 + j];
        const double u1 = xa(u0, du), v1 = xa(v0, dv), w1 = xa(w0, dw);          // uen1
        const double uh = xf(0.5, du, u0), vh = xf(0.5, dv, v0), wh = xf(0.5, dw, w0);
        t1[0] = xf(tx, u1, t1[0]); t1[1] = xf(tx, v1, t1[1]); t1[2] = xf(tx, w1, t1[2]);
        t1[3] = xf(ty, u1, t1[3]); t1[4] = xf(ty, v1, t1[4]); t1[5] = xf(ty, w1, t1[5]);
        t1[6] = xf(tz, u1, t1[6]); t1[7] = xf(tz, v1, t1[7]); t1[8] = xf(tz, w1, t1[8]);
        th[0] = xf(tx, uh, th[0]); th[1] = xf(tx, vh, th[1]); th[2] = xf(tx, wh, th[2]);
        th[3] = xf(ty, uh, th[3]); th[4] = xf(ty, vh, th[4]); th[5] = xf(ty, wh, th[5]);
        th[6] = xf(tz, uh, th[6]); th[7] = xf(tz, vh, th[7]); th[8] = xf(tz, wh, th[8]);
      }
      // fcomp1 (gtmat1.f:601-674)
      double F1[3][3], Fh[3][3];
      F1[0][0] = xa(t1[0], 1.0); F1[0][1] = t1[3]; F1[0][2] = t1[6];
      F1[1][0] = t1[1]; F1[1][1] = xa(t1[4], 1.0); F1[1][2] = t1[7];
      F1[2][0] = t1[2]; F1[2][1] = t1[5]; F1[2][2] = xa(t1[8], 1.0);
      Fh[0][0] = xa(th[0], 1.0); Fh[0][1] = th[3]; Fh[0][2] = th[6];
      Fh[1][0] = th[1]; Fh[1][1] = xa(th[4], 1.0); Fh[1][2] = th[7];
      Fh[2][0] = th[2]; Fh[2][1] = th[5]; Fh[2][2] = xa(th[8], 1.0);
      const double f1 = xf(F1[1][1], F1[2][2], -xm(F1[1][2], F1[2][1]));
      const double f2 = xf(F1[1][0], F1[2][2], -xm(F1[1][2], F1[2][0]));
      const double f3 = xf(F1[1][0], F1[2][1], -xm(F1[1][1], F1[2][0]));
      const double dfn1 = xf(F1[0][2], f3, xf(-F1[0][1], f2, xm(F1[0][0], f1)));
      const double h1 = xf(Fh[1][1], Fh[2][2], -xm(Fh[1][2], Fh[2][1]));
      const double h2 = xf(Fh[1][0], Fh[2][2], -xm(Fh[1][2], Fh[2][0]));
      const double h3 = xf(Fh[1][0], Fh[2][1], -xm(Fh[1][1], Fh[2][0]));
      const double dfh = xf(Fh[0][2], h3, xf(-Fh[0][1], h2, xm(Fh[0][0], h1)));
      error = !(dfn1 > 1.0e-20) || !(dfh > 1.0e-20);
      efac = dfn1 * dj0 * a.beta_fact * weight;
      if (!error) {
        // both decompositions (n+1 and n+1/2) in flight together: two independent chains of
        // sqrt / atan2 / sincos for the FP64 pipe (2 M mesh: 4.00 -> 3.84 ms); W3D_POLAR_ROLLED
        // keeps one copy of the code executed twice
#if defined(W3D_POLAR_ROLLED)
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int w = 0; w < 2; ++w) {
          double F[3][3], R[3][3];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) F[r][c] = w ? Fh[r][c] : F1[r][c];
          polar_rotation(F, R);
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (w == 0) R1[r][c] = R[r][c];
              else Rh[r][c] = R[r][c];
            }
        }
      }
    }
    // ---- strain increment on the mid-step configuration (rknstr_geonl, gtlsn1)
    double ddt[6];
    {
      double Gm[3][3], djm;
      jacobian_smem<false, 0>(sN, gp, a.update_coords ? sXm[el] : C, U, DU, Gm, djm);
      strain_increment(sN, gp, Gm, DU, ddt);
      if (BBAR) {
        // mean-dilatation B-bar (vol_terms / vol_avg / bmod): eps_kk -> eps_kk + beta (thbar - th)
        const double thgp = ddt[0] + ddt[1] + ddt[2];
        const double vol = sum8(djm);
        const double thbar = sum8(djm * thgp) / vol;
        const double corr = beta_b * (thbar - thgp);
        ddt[0] += corr; ddt[1] += corr; ddt[2] += corr;
      }
    }
    double q[6][6];
    getrm1_opt1(Rh, q);
    qmply(q, ddt, uddt);
  } else {
    // ---- rknstr_sm_displ + rstgp2: B on the undeformed configuration
    double G0[3][3], dj0;
    jacobian_smem<false, 0>(sN, gp, C, U, DU, G0, dj0);
    efac = dj0 * a.beta_fact * weight;
    strain_increment(sN, gp, G0, DU, uddt);
    if (BBAR) {
      const double thgp = uddt[0] + uddt[1] + uddt[2];
      const double vol = sum8(dj0);
      const double thbar = sum8(dj0 * thgp) / vol;
      const double corr = beta_b * (thbar - thgp);
      uddt[0] += corr; uddt[1] += corr; uddt[2] += corr;
    }
  }

  // a Gauss point with det F <= 0 asks for a step cut; its state is not written
  const bool store = valid && !error;
  if (error && valid) {
    atomicOr(&a.flags[0], 1);
    if (!a.adaptive) atomicMax(&a.flags[1], 6);
  }

  // ---- rstgp1_update_strains (rstgp1.f:4983-5005): total strain at n+1, stored at once
  cp_async_wait_all();   // each thread reads back only what it copied itself ...
  if (MAT == 1 || MAT == 5) __syncwarp();   // ... except the property record, shared by the 8 lanes of an element
  if (PHASE != 2) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double e1 = sSt[k][tid] + uddt[k];
      if (store) a.eps_n1[k * ng + g] = e1;
    }
  }
  if (PHASE != 2 && store && GEONL && a.iter > 0) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) a.rot_n1[(c * 3 + r) * ng + g] = R1[r][c];
  }
  if (PHASE == 1) {
    // hand R(n+1), the un-rotated strain increment and the energy factor to k_mm10
    if (valid) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) a.scratch[(c * 3 + r) * ng + g] = R1[r][c];
#pragma unroll
      for (int k = 0; k < 6; ++k) a.scratch[(9 + k) * ng + g] = uddt[k];
      a.scratch[15 * ng + g] = error ? __longlong_as_double(0x7ff8000000000000LL) : efac;
    }
    return;
  }

  // ---- material update; rplstr stores (rplstr.f:15-93) + rstgp1_store_cep issued as soon as the
  //      values exist so their registers die early
  double cgn1[9];
  int fatal = 0;
  if (MAT == 1) {
    double cgn[9], cep[21];
#pragma unroll
    for (int k = 0; k < 9; ++k) cgn[k] = sSt[6 + k][tid];
    const float *ep = reinterpret_cast<const float *>(&sEp[el][0]);
    Mm01Props pr;
    pr.e = (double)ep[0]; pr.nu = (double)ep[1]; pr.beta = (double)ep[2];
    pr.hprime = (double)ep[3]; pr.yld = (double)ep[4]; pr.lnelas = ep[5] != 0.0f;
    if (a.nonlinear_update) {
      double h[11], h1[11], rtse[6];
#pragma unroll
      for (int k = 0; k < 11; ++k) h[k] = sSt[15 + k][tid];
      double h11_first = h[10];
      if (1.0 - pr.beta != 0.0) h11_first = a.hist_n[10 * ng + (long long)cd.blk_slot0 * 8 + gp];
      fatal = mm01_update(a.step, pr, cgn, cgn1, uddt, h, h1, rtse, h11_first);
      if (store && a.iter > 0) {
#pragma unroll
        for (int k = 0; k < 11; ++k) a.hist_n1[k * ng + g] = h1[k];
      }
      cnst1_cep21(pr, rtse, h1, cep);
    } else {
      // drive_01_update_a/_b: linear-elastic [D] and stress for iteration 0
      elastic_cep21(pr.e, pr.nu, cep);
      const double c2 = cep[0], c4 = cep[1], c3 = cep[9];
      cgn1[0] = cgn[0] + (c2 * uddt[0] + c4 * uddt[1] + c4 * uddt[2]);
      cgn1[1] = cgn[1] + (c4 * uddt[0] + c2 * uddt[1] + c4 * uddt[2]);
      cgn1[2] = cgn[2] + (c4 * uddt[0] + c4 * uddt[1] + c2 * uddt[2]);
      cgn1[3] = cgn[3] + c3 * uddt[3];
      cgn1[4] = cgn[4] + c3 * uddt[4];
      cgn1[5] = cgn[5] + c3 * uddt[5];
      cgn1[6] = 0.0; cgn1[7] = 0.0; cgn1[8] = 0.0;   // local urcs_blk_n1 is zero-initialised
    }
    if (store) {
#pragma unroll
      for (int k = 0; k < 21; ++k) a.cep[k * ng + g] = cep[k];
#pragma unroll
      for (int k = 0; k < 9; ++k) a.urcs_n1[k * ng + g] = cgn1[k];
    }
  }
  bool mm05_cut = false;
  if (MAT == 5) {
    // drive_05_update (rstgp1.f:2373-2700): mm05 nonlinear_hardening + cnst5, or linear-elastic
    double cgn[9], D[6][6];
#pragma unroll
    for (int k = 0; k < 9; ++k) cgn[k] = sSt[6 + k][tid];
    const float *ep = reinterpret_cast<const float *>(&sEp[el][0]);
    w3d_mm05::Props pr;
    pr.e = (double)ep[0]; pr.nu = (double)ep[1]; pr.sigyld = (double)ep[4];
#pragma unroll
    for (int k = 0; k < 10; ++k) pr.mm[k] = (double)ep[6 + k];
    if (a.nonlinear_update) {
      double h[14], h1[14], trial[6];
#pragma unroll
      for (int k = 0; k < 14; ++k) { h[k] = sSt[15 + k][tid]; h1[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < 9; ++k) cgn1[k] = 0.0;
      // nonlinear_hardening or generalized_plasticity by the sign of matprp(58) (mm05.f:232-254)
      const int rc = w3d_mm05::update(a.step, a.iter, pr, cgn, cgn1, uddt, h, h1, trial, a.adaptive != 0);
      if (rc == w3d_mm05::CUT) mm05_cut = true;
      else if (rc != w3d_mm05::OK) fatal = rc;
      w3d_mm05::cnst5(pr, trial, h1, cgn1, D);
      if (store && !mm05_cut && a.iter > 0) {
#pragma unroll
        for (int k = 0; k < 14; ++k) a.hist_n1[k * ng + g] = h1[k];
      }
    } else {
      double c21[21];
      elastic_cep21(pr.e, pr.nu, c21);
      const double c2 = c21[0], c4 = c21[1], c3 = c21[9];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) D[r][c] = 0.0;
      D[0][0] = c2; D[1][1] = c2; D[2][2] = c2; D[3][3] = c3; D[4][4] = c3; D[5][5] = c3;
      D[1][0] = c4; D[2][0] = c4; D[2][1] = c4;
      cgn1[0] = cgn[0] + (c2 * uddt[0] + c4 * uddt[1] + c4 * uddt[2]);
      cgn1[1] = cgn[1] + (c4 * uddt[0] + c2 * uddt[1] + c4 * uddt[2]);
      cgn1[2] = cgn[2] + (c4 * uddt[0] + c4 * uddt[1] + c2 * uddt[2]);
      cgn1[3] = cgn[3] + c3 * uddt[3];
      cgn1[4] = cgn[4] + c3 * uddt[4];
      cgn1[5] = cgn[5] + c3 * uddt[5];
      cgn1[6] = 0.0; cgn1[7] = 0.0; cgn1[8] = 0.0;
    }
    if (mm05_cut && valid) atomicOr(&a.flags[0], 1);
    if (store && !mm05_cut) {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) { a.cep[k * ng + g] = D[r][c]; ++k; }
#pragma unroll
      for (int q = 0; q < 9; ++q) a.urcs_n1[q * ng + g] = cgn1[q];
    }
    if (mm05_cut) {
#pragma unroll
      for (int q = 0; q < 9; ++q) cgn1[q] = 0.0;
    }
  }
  if (MAT == 3) {
    // PHASE 2: stresses at n+1 and [D] come from k_mm03 (w3d_mm03.cu)
#pragma unroll
    for (int k = 0; k < 9; ++k) cgn1[k] = a.urcs_n1[k * ng + g];
  }
  if (MAT == 10) {
    // PHASE 2: stresses at n+1 come from k_mm10
#pragma unroll
    for (int k = 0; k < 9; ++k) cgn1[k] = a.urcs_n1[k * ng + g];
    // rknstr_finish_cp / mm10_calc_grads: least-squares gradient of Rp^T R over the 8 points
    double *sRt = &sSt[0][0];                 // reuse: [el][gp][9]
    {
      double Rp[9];
#pragma unroll
      for (int k = 0; k < 9; ++k)
        Rp[k] = (a.step == 1) ? ((k == 0 || k == 4 || k == 8) ? 1.0 : 0.0)
                              : a.hist_n[(long long)(a.cp_o_Rp + k) * ng + g];
      double *rt = sRt + (el * 8 + gp) * 9;
      if (GEONL) {
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb)   // Rt = Rp * R^T (mult_type_3t), column-major storage
            rt[bb * 3 + aa] = Rp[0 * 3 + aa] * R1[bb][0] + Rp[1 * 3 + aa] * R1[bb][1] + Rp[2 * 3 + aa] * R1[bb][2];
      } else {
        // small strain: mm10_calc_grads copies the FIRST ngp*9 contiguous words of the block-local
        // history array starting at elem_hist(1, rs, 1) (mm10_a_copy_vector(Rt3d, Rps, ngp*9),
        // mm10_a.f:1622) -- word rs + k / span of block element 1 + mod(k, span) at Gauss point 1,
        // whichever element is being fitted.  Reproduced literally (the bar is the reference's
        // result): lane gp fills rt[gp][0..8] = flat[gp*9 ..].
#pragma unroll
        for (int k9 = 0; k9 < 9; ++k9) {
          const int k = gp * 9 + k9;
          const int i = k % cd.blk_span;
          int word = a.cp_o_Rp + k / cd.blk_span;          // runs on into the next Gauss-point
          const int plane = word / a.hist_size;            // plane of the (span, hist, ngp) array
          word -= plane * a.hist_size;                     // for blocks of very few elements
          const int q = word - a.cp_o_Rp;
          rt[k9] = (a.step == 1) ? ((q == 0 || q == 4 || q == 8) ? 1.0 : 0.0)
                                 : a.hist_n[(long long)word * ng + (long long)(cd.blk_slot0 + i) * 8 + plane];
        }
      }
    }
    __syncwarp();
    if (valid) {
      double jacinv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      if (GEONL) {
        double Jm[3][3];
        jacobian_matrix_exact(sN, 7, C, Jm);
        const double det = Jm[0][0] * Jm[1][1] * Jm[2][2] - Jm[0][0] * Jm[1][2] * Jm[2][1] -
                           Jm[0][1] * Jm[1][0] * Jm[2][2] + Jm[0][1] * Jm[1][2] * Jm[2][0] +
                           Jm[0][2] * Jm[1][0] * Jm[2][1] - Jm[0][2] * Jm[1][1] * Jm[2][0];
        double cof[3][3];
        cof[0][0] = +(Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]);
        cof[0][1] = -(Jm[1][0] * Jm[2][2] - Jm[1][2] * Jm[2][0]);
        cof[0][2] = +(Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]);
        cof[1][0] = -(Jm[0][1] * Jm[2][2] - Jm[0][2] * Jm[2][1]);
        cof[1][1] = +(Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]);
        cof[1][2] = -(Jm[0][0] * Jm[2][1] - Jm[0][1] * Jm[2][0]);
        cof[2][0] = +(Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]);
        cof[2][1] = -(Jm[0][0] * Jm[1][2] - Jm[0][2] * Jm[1][0]);
        cof[2][2] = +(Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) jacinv[r][c] = (fabs(det) <= 1.0e-10) ? 0.0 : cof[c][r] / det;
      }
      double grads[27];
      if (a.cp_ncrystals > 1) {
#pragma unroll
        for (int k = 0; k < 27; ++k) grads[k] = 0.0;
      } else {
        w3d_mm10::calc_grads((const double (*)[9])(sRt + el * 72), GEONL, jacinv, grads);
      }
      if (!error) {
#pragma unroll
        for (int k = 0; k < 27; ++k) a.hist_n1[(long long)(a.cp_o_gradfe + k) * ng + g] = grads[k];
      }
    }
  }
  if (fatal && store) atomicMax(&a.flags[1], fatal);

  // ---- rstgp1_b: energy integration over the undeformed volume
  double en = 0.0, pw = 0.0;
  if (a.iter != 0 && store) { en = cgn1[6] * efac; pw = cgn1[7] * efac; }

  // ---- internal force on the n+1 configuration (rknifv_std, gpifv1)
  double s[6];
  double evol;
  double G1[3][3];
  {
    double sig[6];
    if (GEONL) {
      double q[6][6];
      getrm1_opt2(R1, q);
      const double t6[6] = {cgn1[0], cgn1[1], cgn1[2], cgn1[3], cgn1[4], cgn1[5]};
      qmply(q, t6, sig);
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) sig[k] = cgn1[k];
    }
    if (error) {
#pragma unroll
      for (int k = 0; k < 6; ++k) sig[k] = 0.0;
    }
    double dj1;
    jacobian_smem<false, 0>(sN, gp, (GEONL && a.update_coords) ? sX1[el] : C, U, DU, G1, dj1);
    const double wdj = weight * dj1;
    evol = sum8(wdj);
    double add = 0.0;
    if (BBAR) {
      // B-bar^T sigma = B^T (sigma - beta p m) + beta (P/V) sum_gp B_vol^T : fold the mean
      // dilatation part into a modified stress so no nodal volume terms are exchanged
      const double p = sig[0] + sig[1] + sig[2];
      const double pbar = sum8(p * wdj) / sum8(dj1);
      add = beta_b * (pbar * dj1 - p * wdj);
    }
    s[0] = fma(sig[0], wdj, add); s[1] = fma(sig[1], wdj, add); s[2] = fma(sig[2], wdj, add);
    s[3] = sig[3] * wdj; s[4] = sig[4] * wdj; s[5] = sig[5] * wdj;
  }

  double f[24];   // [node*3 + comp]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double nx = sN[0][j][gp], ne = sN[1][j][gp], nz = sN[2][j][gp];
    const double bx = G1[0][0] * nx + G1[0][1] * ne + G1[0][2] * nz;
    const double by = G1[1][0] * nx + G1[1][1] * ne + G1[1][2] * nz;
    const double bz = G1[2][0] * nx + G1[2][1] * ne + G1[2][2] * nz;
    f[j * 3 + 0] = bx * s[0] + by * s[3] + bz * s[5];
    f[j * 3 + 1] = by * s[1] + bx * s[3] + bz * s[4];
    f[j * 3 + 2] = bz * s[2] + by * s[4] + bx * s[5];
  }
  // reduce-scatter over the 8 Gauss-point lanes: lane n ends with node n's 3 force components
  {
    const bool hi4 = (gp & 4) != 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const double send = hi4 ? f[k] : f[12 + k];
      const double keep = hi4 ? f[12 + k] : f[k];
      f[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const bool hi2 = (gp & 2) != 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double send = hi2 ? f[k] : f[6 + k];
      const double keep = hi2 ? f[6 + k] : f[k];
      f[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const bool hi1 = (gp & 1) != 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double send = hi1 ? f[k] : f[3 + k];
      const double keep = hi1 ? f[3 + k] : f[k];
      f[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  }
  // ---- addifv_standard: scatter with FP64 reductions at L2, sum |f| for the convergence test
  double sabs = 0.0;
  if (valid) {
    sabs = fabs(f[0]) + fabs(f[1]) + fabs(f[2]);
    const double bf = a.beta_fact;
    atomicAdd(&a.ifv[dofs[0]], bf * f[0]);
    atomicAdd(&a.ifv[dofs[1]], bf * f[1]);
    atomicAdd(&a.ifv[dofs[2]], bf * f[2]);
    if (gp == 0) a.elem_vol[slot] = evol;
  }

  // ---- per-CTA partial sums, reduced in a fixed order by k_reduce_partials
  en = warp_sum(en); pw = warp_sum(pw); sabs = warp_sum(sabs);
  if ((tid & 31) == 0) { sRed[tid >> 5][0] = en; sRed[tid >> 5][1] = pw; sRed[tid >> 5][2] = sabs; }
  __syncthreads();
  if (tid % TPB < 3 && have) {     // one partial per 16-element group, as the host reduces them
    const int q = tid % TPB, w0 = (tid / TPB) * (TPB / 32);
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < TPB / 32; ++w) v += sRed[w0 + w][q];
    a.partials[(size_t)idesc * 4 + q] = v;
  }
}

// first stage of the deterministic reduction of the per-CTA partials when there are many of them
// (125 000 CTAs on the 2 M-element mesh took one CTA 130 us): REDUCE_STAGE1_CTAS CTAs each sum a
// fixed contiguous chunk in a fixed order into tmp[cta*4 + k]; k_reduce_partials finishes
constexpr int REDUCE_STAGE1_CTAS = 128;
static __global__ void __launch_bounds__(256) k_reduce_partials_stage1(const double *partials, int ncta,
                                                                       double *tmp) {
  __shared__ double s[3][256];
  const int chunk = (ncta + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * chunk, hi = min(ncta, lo + chunk);
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    v0 += partials[(size_t)i * 4 + 0];
    v1 += partials[(size_t)i * 4 + 1];
    v2 += partials[(size_t)i * 4 + 2];
  }
  s[0][threadIdx.x] = v0; s[1][threadIdx.x] = v1; s[2][threadIdx.x] = v2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) tmp[(size_t)blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x][0];
}

// deterministic final reduction of the per-CTA partials: one CTA, fixed tree
static __global__ void __launch_bounds__(256) k_reduce_partials(const double *partials, int ncta,
                                                         double *out /*[4]*/, int accumulate) {
  __shared__ double s[3][256];
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  for (int i = threadIdx.x; i < ncta; i += 256) {
    v0 += partials[(size_t)i * 4 + 0];
    v1 += partials[(size_t)i * 4 + 1];
    v2 += partials[(size_t)i * 4 + 2];
  }
  s[0][threadIdx.x] = v0; s[1][threadIdx.x] = v1; s[2][threadIdx.x] = v2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    const double r = s[threadIdx.x][0];
    out[threadIdx.x] = accumulate ? out[threadIdx.x] + r : r;
  }
}

}  // namespace w3d
