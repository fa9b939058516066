// w3d_tanstf_kernel.cuh -- element tangent stiffness kernel (hex8, 300-term upper triangle).
//
// One launch replaces, for every owned block of a segment, the reference chain
//   dptstf (dptstf.f:64-156) + dptstf_blocks (tanstf.f:642-717)
//   rktstf (rktstf.f:15-220): jacob1, vol_terms, vol_avg, per Gauss point gptns1
//   gptns1 (gptns1.f:16-342): getrm1(opt 2)+qmply1, blcmp1, bmod, extract_D_symmetric (:1174),
//     ctran1 + ctran1_qbar (:1735-1970), bdbtgen (:1628-1718), kgstiff (kg1.f:16-160)
//   rktstf_do_transpose (rktstf.f:221-296)
// and writes estiff in the reference layout (utsz=300, span), blocks concatenated.
//
// Phase 1, thread = (element, Gauss point): geometry on the n+1 configuration, Cauchy stress,
//   rotated / q-bar corrected [D*] w detJ -> shared memory.
// Phase 2, thread = (element, node l): accumulates the 3x3 node-pair blocks
//   K(a,l) = sum_gp Bbar_a^T D* Bbar_l,  a = l, l+1, .. l+4 (mod 8)   (36 blocks over 8 lanes)
//   with Bbar = B + m g^T (m = [1 1 1 0 0 0]^T, g = beta (V - N)) expanded so each entry costs
//   4 FMAs.  Results are staged per warp in shared memory and stored with coalesced 16 B writes.
#pragma once
#include "w3d_device.cuh"

namespace w3d {

struct TanstfArgs {
  const CtaDesc *cta;
  const int32_t *edest, *cdest;
  const double *c, *u, *du;
  const double *urcs_n1, *rot_n1, *cep;
  double *estiff;        // (300, nelem) in unpadded element order
  long long ng;
  double eps_bbar, beta_fact;
  int qbar;
  // mm10 keeps its consistent [D] as a full 6x6 (column-major) in history(1:36) at n+1
  // (drive_10_cnst, gptns1.f:1371-1410); read the lower triangle from there when set
  const double *hist_n1;
  int d_from_hist;
};

// shared-memory record of one (element, Gauss point): 24 N-derivatives [node][3], 21 D* terms
// (lower triangle by rows), 6 = D* m, 6 = sigma w detJ, 1 = detJ
// (58 doubles used, 58..63 double as the landing zone of the cp.async stress prefetch).  Strides are chosen for the 32-bank shared memory: GPREC = 1 (mod 16) spreads
// the 8 Gauss points of an element over distinct bank pairs, ELREC = 8 (mod 16) puts the second
// element of a half-warp 16 banks away.
constexpr int GPREC = 65;
constexpr int ELREC = 8 * GPREC + 24 + 6 + 2;     // + V[node][3] + Sbar[6] (+2 pad) = 552
static_assert(GPREC >= 58 && GPREC % 16 == 1 && ELREC % 16 == 8, "bank layout");
static_assert(4 * ELREC >= 4 * 300, "staging must fit in the warp's phase-1 region");

// position of dof pair (r, c), 0-based, in the column-major upper triangle (setup.f:268-276)
__device__ __forceinline__ int utpos(int r, int c) {
  const int lo = min(r, c), hi = max(r, c);
  return (hi * (hi + 1)) / 2 + lo;
}

#ifndef TANSTF_MIN_CTAS
#define TANSTF_MIN_CTAS 3   // register cap 168/thread; 3 x 70 KB of shared memory per SM
#endif

template <bool GEONL, bool BBAR>
__global__ void __launch_bounds__(TPB, TANSTF_MIN_CTAS) k_tanstf(const TanstfArgs a) {
  extern __shared__ double smem[];
  __shared__ double sN[3][8][8];
  const int tid = threadIdx.x;
  for (int k = tid; k < 192; k += TPB) (&sN[0][0][0])[k] = (&c_shape[0][0][0])[k];
  __syncthreads();   // the only CTA-wide barrier: everything below is exchanged within a warp (4 whole elements)

  const CtaDesc cd = a.cta[blockIdx.x];
  const int el = tid >> 3, gp = tid & 7;
  const bool valid = el < cd.nvalid;
  const int slot = cd.slot0 + (valid ? el : cd.nvalid - 1);
  const long long g = (long long)slot * 8 + gp;
  const long long ng = a.ng;
  double *er = smem + (size_t)el * ELREC;     // this element's record
  double *gr = er + gp * GPREC;               // this Gauss point's record
  double *sV = er + 8 * GPREC;                // V[node][3]
  double *sS = sV + 24;                       // Sbar[6]

  // ---- dof / coordinate numbers of node `gp` first: their round trip overlaps the state stream
  int dofs[3], cidx[3];
  {
    const int32_t *ed = a.edest + (size_t)slot * 24;
    const int32_t *cdst = a.cdest + (size_t)slot * 24;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cidx[k] = cdst[k * 8 + gp] - 1;
      dofs[k] = GEONL ? ed[k * 8 + gp] - 1 : 0;
    }
  }
  // ---- start streaming this point's [D], R and stresses into its (still unused) record slots:
  //      cep -> gr[24..44] (overwritten in place by D* later), R -> gr[45..53], t -> gr[58..63]
  {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) {
        cp_async8(gr + 24 + k, a.d_from_hist ? a.hist_n1 + (long long)(c * 6 + r) * ng + g
                                             : a.cep + (long long)k * ng + g);
        ++k;
      }
    if (GEONL) {
#pragma unroll
      for (int q9 = 0; q9 < 9; ++q9) cp_async8(gr + 45 + q9, a.rot_n1 + (long long)q9 * ng + g);
#pragma unroll
      for (int q6 = 0; q6 < 6; ++q6) cp_async8(gr + 58 + q6, a.urcs_n1 + (long long)q6 * ng + g);
    }
    cp_async_commit();
  }
  // ---- L2 prefetch for a CTA that will start soon (AHEAD CTAs later = a slot offset: slots are
  //      laid out in CTA order; approximate where blocks are padded, it is only a hint): its
  //      state stream and index tables are then L2 hits instead of HBM round trips under load.
  //      Measured on the 2 M-element mesh (launch, ms): none 4.144; 74 / 148 / 222 CTAs ahead
  //      3.966 / 3.966 / 3.964; 444 3.987; 888 4.212 (evicted again before use).  The same hint
  //      in the stress kernel changed nothing (3.720 vs 3.725: its state is consumed after the
  //      long kinematics, the stream is already hidden) and was left out.
#ifndef TANSTF_PREFETCH_AHEAD
#define TANSTF_PREFETCH_AHEAD 148
#endif
  if (TANSTF_PREFETCH_AHEAD > 0) {
    const long long pslot = (long long)slot + (long long)TANSTF_PREFETCH_AHEAD * EPC;
    const long long pg = pslot * 8 + gp;
    if (pg < ng) {
      if ((tid & 3) == 0) {      // one request per 32-byte sector of the 8-byte SoA streams
        if (a.d_from_hist) {
#pragma unroll 1
          for (int k = 0; k < 36; ++k) prefetch_l2(a.hist_n1 + (long long)k * ng + pg);
        } else {
#pragma unroll
          for (int k = 0; k < 21; ++k) prefetch_l2(a.cep + (long long)k * ng + pg);
        }
        if (GEONL) {
#pragma unroll
          for (int q9 = 0; q9 < 9; ++q9) prefetch_l2(a.rot_n1 + (long long)q9 * ng + pg);
#pragma unroll
          for (int q6 = 0; q6 < 6; ++q6) prefetch_l2(a.urcs_n1 + (long long)q6 * ng + pg);
        }
      }
      if (gp < 3) {              // 96 B of dof numbers and 96 B of coordinate numbers per element
        prefetch_l2(a.cdest + (size_t)pslot * 24 + gp * 8);
        if (GEONL) prefetch_l2(a.edest + (size_t)pslot * 24 + gp * 8);
      }
    }
  }
  // ---- dptstf: coordinates of node `gp` on the n+1 configuration, staged in the N slots
  {
    double vc[3], vu[3], vdu[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vc[k] = a.c[cidx[k]];
      if (GEONL) { vu[k] = a.u[dofs[k]]; vdu[k] = a.du[dofs[k]]; }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double x = vc[k];
      if (GEONL) x = (x + vu[k]) + vdu[k];
      sV[k * 8 + gp] = x;   // temporary: coordinates [comp][node]
    }
  }
  __syncwarp();
  double nx[8], ne[8], nz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { nx[j] = sN[0][j][gp]; ne[j] = sN[1][j][gp]; nz[j] = sN[2][j][gp]; }

  // ---- phase 1
  const double beta_b = (1.0 - a.eps_bbar) * (1.0 / 3.0);
  double dj;
  {
    double X[8], Y[8], Z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { X[j] = sV[j]; Y[j] = sV[8 + j]; Z[j] = sV[16 + j]; }
    double G[3][3];
    jacobian<false>(nx, ne, nz, X, Y, Z, G, dj);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gr[j * 3 + 0] = G[0][0] * nx[j] + G[0][1] * ne[j] + G[0][2] * nz[j];
      gr[j * 3 + 1] = G[1][0] * nx[j] + G[1][1] * ne[j] + G[1][2] * nz[j];
      gr[j * 3 + 2] = G[2][0] * nx[j] + G[2][1] * ne[j] + G[2][2] * nz[j];
    }
  }
  const double wdj = 1.0 * dj;   // quad2 weight = 1
  {
    // extract_D_symmetric: D = w detJ cep(21 lower-triangle terms by rows)
    cp_async_wait_all();   // each thread reads only what it copied itself
    double D[6][6];
    {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
          const double v = wdj * gr[24 + k];
          D[r][c] = v; D[c][r] = v;
          ++k;
        }
    }
    double cs[6] = {0, 0, 0, 0, 0, 0};
    if (GEONL) {
      // gptns1_a: Cauchy stress; ctran1: D* = q D q^T (- Q-bar)
      double R[3][3], q[6][6];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = gr[45 + c * 3 + r];
      getrm1_opt2(R, q);
      const double t6[6] = {gr[58], gr[59], gr[60], gr[61], gr[62], gr[63]};
      qmply(q, t6, cs);
      double tc[6][6];
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int k = 0; k < 6; ++k)
          tc[j][k] = q[j][0] * D[0][k] + q[j][1] * D[1][k] + q[j][2] * D[2][k] +
                     q[j][3] * D[3][k] + q[j][4] * D[4][k] + q[j][5] * D[5][k];
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int k = 0; k <= j; ++k) {
          const double v = tc[j][0] * q[k][0] + tc[j][1] * q[k][1] + tc[j][2] * q[k][2] +
                           tc[j][3] * q[k][3] + tc[j][4] * q[k][4] + tc[j][5] * q[k][5];
          D[j][k] = v; D[k][j] = v;
        }
      if (a.qbar) {
        const double wf = wdj, halfw = 0.5 * wf;
        const double xx = cs[0], yy = cs[1], zz = cs[2], xy = cs[3], yz = cs[4], xz = cs[5];
        D[0][0] -= 2.0 * xx * wf; D[1][1] -= 2.0 * yy * wf; D[2][2] -= 2.0 * zz * wf;
        D[3][0] -= xy * wf; D[5][0] -= xz * wf; D[3][1] -= xy * wf;
        D[4][1] -= yz * wf; D[4][2] -= yz * wf; D[5][2] -= xz * wf;
        D[3][3] -= halfw * (xx + yy); D[4][4] -= halfw * (yy + zz); D[5][5] -= halfw * (xx + zz);
        D[4][3] -= wf * xz; D[5][3] -= wf * yz; D[5][4] -= wf * xy;
      }
    }
    {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) { gr[24 + k] = D[r][c]; ++k; }
    }
    // D* m (m = [1 1 1 0 0 0]); read from the lower triangle so the matrix stays symmetric
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double d0 = (r >= 0) ? D[r][0] : D[0][r];
      const double d1 = (r >= 1) ? D[r][1] : D[1][r];
      const double d2 = (r >= 2) ? D[r][2] : D[2][r];
      gr[45 + r] = d0 + d1 + d2;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) gr[51 + k] = cs[k] * wdj;
    gr[57] = dj;
  }
  __syncwarp();

  // ---- element means (vol_terms / vol_avg) and summed stress, lane = node l
  const int l = gp;
  double Vl[3] = {0, 0, 0};
  {
    double vx = 0, vy = 0, vz = 0, vol = 0, sb = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const double *pr = er + p * GPREC;
      const double d = pr[57];
      vx = fma(pr[l * 3 + 0], d, vx);
      vy = fma(pr[l * 3 + 1], d, vy);
      vz = fma(pr[l * 3 + 2], d, vz);
      vol += d;
      if (l < 6) sb += pr[51 + l];
    }
    __syncwarp();   // coordinates in sV are dead (phase 1 finished) before V overwrites them
    if (BBAR) {
      const double voli = 1.0 / vol;
      Vl[0] = vx * voli; Vl[1] = vy * voli; Vl[2] = vz * voli;
      sV[l * 3 + 0] = Vl[0]; sV[l * 3 + 1] = Vl[1]; sV[l * 3 + 2] = Vl[2];
    }
    if (l < 6) sS[l] = sb;
  }
  __syncwarp();

  // ---- phase 2
  double acc[5][3][3];
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[d][i][j] = 0.0;
  const int nblk = (l < 4) ? 5 : 4;   // pairs {l, l+4} belong to lanes 0..3
  double SV[3] = {0.0, 0.0, 0.0};     // sum over Gauss points of m^T D* Bbar_l  (B-bar only)

#ifndef TANSTF_P_UNROLL
#define TANSTF_P_UNROLL 1
#endif
#define W3D_PRAGMA_(x) _Pragma(#x)
#define W3D_UNROLL_N(n) W3D_PRAGMA_(unroll n)
  W3D_UNROLL_N(TANSTF_P_UNROLL)
  for (int p = 0; p < 8; ++p) {
    const double *pr = er + p * GPREC;
    // D* (lower triangle) and D* m
    double D[6][6];
    {
      int k = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) { const double v = pr[24 + k]; D[r][c] = v; D[c][r] = v; ++k; }
    }
    double Dm[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) Dm[r] = pr[45 + r];
    // own node: DB[:,cl] = D* Bbar_l[:,cl]
    const double Nx = pr[l * 3 + 0], Ny = pr[l * 3 + 1], Nz = pr[l * 3 + 2];
    double gx = 0, gy = 0, gz = 0;
    if (BBAR) { gx = beta_b * (Vl[0] - Nx); gy = beta_b * (Vl[1] - Ny); gz = beta_b * (Vl[2] - Nz); }
    double DB[3][6], sv[3];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      // Bbar_l u-column = [Nx 0 0 Ny 0 Nz] + gx m ; v = [0 Ny 0 Nx Nz 0] + gy m ; w = [0 0 Nz 0 Ny Nx] + gz m
      DB[0][r] = D[r][0] * Nx + D[r][3] * Ny + D[r][5] * Nz;
      DB[1][r] = D[r][1] * Ny + D[r][3] * Nx + D[r][4] * Nz;
      DB[2][r] = D[r][2] * Nz + D[r][4] * Ny + D[r][5] * Nx;
      if (BBAR) {
        DB[0][r] = fma(Dm[r], gx, DB[0][r]);
        DB[1][r] = fma(Dm[r], gy, DB[1][r]);
        DB[2][r] = fma(Dm[r], gz, DB[2][r]);
      }
    }
#pragma unroll
    for (int cl = 0; cl < 3; ++cl) sv[cl] = DB[cl][0] + DB[cl][1] + DB[cl][2];
    if (BBAR) {
      // partner row  Bbar_a = B_a + m g_a^T,  g_a = beta (V_a - N_a):  the N_a part of g_a is
      // folded into the first three entries of DB, the V_a part (constant over the Gauss points)
      // is applied once after the loop through SV = sum_gp sv.
#pragma unroll
      for (int cl = 0; cl < 3; ++cl) {
        SV[cl] += sv[cl];
        const double bs = beta_b * sv[cl];
        DB[cl][0] -= bs; DB[cl][1] -= bs; DB[cl][2] -= bs;
      }
    }
    // partner nodes: 27 FMAs per 3x3 block, accumulators are the FMA addends
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if (d < nblk) {
        const int an = (l + d) & 7;
        const double ax = pr[an * 3 + 0], ay = pr[an * 3 + 1], az = pr[an * 3 + 2];
#pragma unroll
        for (int cl = 0; cl < 3; ++cl) {
          const double *v = DB[cl];
          acc[d][0][cl] = fma(az, v[5], fma(ay, v[3], fma(ax, v[0], acc[d][0][cl])));
          acc[d][1][cl] = fma(az, v[4], fma(ax, v[3], fma(ay, v[1], acc[d][1][cl])));
          acc[d][2][cl] = fma(ax, v[5], fma(ay, v[4], fma(az, v[2], acc[d][2][cl])));
        }
        if (GEONL && !BBAR) {
          // kgstiff (kg1.f:16-160) without B-bar: per-point gradients
          const double s0 = pr[51], s1 = pr[52], s2 = pr[53], s3 = pr[54], s4 = pr[55], s5 = pr[56];
          const double gtg = ax * Nx * s0 + ay * Ny * s1 + az * Nz * s2 + (ax * Ny + Nx * ay) * s3 +
                             (ay * Nz + Ny * az) * s4 + (ax * Nz + Nx * az) * s5;
          acc[d][0][0] += gtg; acc[d][1][1] += gtg; acc[d][2][2] += gtg;
        }
      }
    }
  }
  if (BBAR) {
    // the V_a part of the partner rows: K(a,l) += beta V_a (x) SV
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if (d < nblk) {
        const int an = (l + d) & 7;
#pragma unroll
        for (int ca = 0; ca < 3; ++ca) {
          const double bv = beta_b * sV[an * 3 + ca];
#pragma unroll
          for (int cl = 0; cl < 3; ++cl) acc[d][ca][cl] = fma(bv, SV[cl], acc[d][ca][cl]);
        }
      }
    }
  }
  if (GEONL && BBAR) {
    // kgstiff with B-bar uses the element-mean gradients V, so the Gauss sum factorises:
    // gtg(a,l) = V_a^T (sum_gp sigma w detJ) V_l
    const double s0 = sS[0], s1 = sS[1], s2 = sS[2], s3 = sS[3], s4 = sS[4], s5 = sS[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      if (d < nblk) {
        const int an = (l + d) & 7;
        const double ax = sV[an * 3 + 0], ay = sV[an * 3 + 1], az = sV[an * 3 + 2];
        const double gtg = ax * Vl[0] * s0 + ay * Vl[1] * s1 + az * Vl[2] * s2 +
                           (ax * Vl[1] + Vl[0] * ay) * s3 + (ay * Vl[2] + Vl[1] * az) * s4 +
                           (ax * Vl[2] + Vl[0] * az) * s5;
        acc[d][0][0] += gtg; acc[d][1][1] += gtg; acc[d][2][2] += gtg;
      }
    }
  }
  __syncwarp();   // all lanes of the warp are done reading the phase-1 records

  // ---- stage the warp's 4 x 300 terms and store them contiguously
  const int warp = tid >> 5, lane = tid & 31;
  double *stage = smem + (size_t)(warp * 4) * ELREC;           // reuse this warp's records
  double *mine = stage + (el & 3) * 300;
  const double bf = a.beta_fact;
#pragma unroll
  for (int d = 0; d < 5; ++d) {
    if (d < nblk) {
      const int an = (l + d) & 7;
#pragma unroll
      for (int ca = 0; ca < 3; ++ca)
#pragma unroll
        for (int cl = 0; cl < 3; ++cl) {
          if (d == 0 && ca > cl) continue;   // diagonal node block: upper part only
          mine[utpos(ca * 8 + an, cl * 8 + l)] = bf * acc[d][ca][cl];
        }
    }
  }
  __syncwarp();
  {
    const int e0 = warp * 4;                                   // first element of this warp
    const int nel = min(4, cd.nvalid - e0);                    // real elements in this warp
    if (nel > 0) {
      const int n2 = nel * 150;                                // double2 count
      double2 *dst = reinterpret_cast<double2 *>(a.estiff + (size_t)(cd.epos0 + e0) * 300);
      const double2 *src = reinterpret_cast<const double2 *>(stage);
      for (int k = lane; k < n2; k += 32) dst[k] = src[k];
    }
  }
}

}  // namespace w3d
