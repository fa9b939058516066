// w3d_residual.cuh -- structure residual and its norms on the device (SURVEY 8(f) NEXT-2).
// Included at the end of w3d_engine.cu (it needs the engine struct).
//
// Replaces reference src/upres.f (upres_iter_0 :13-77, upres :121-226) and the residual loop of
// src/cvtest.f:62-78.  The internal forces of the last stress pass and the displacement increment
// are already in HBM (d_ifv, d_du), so the Newton driver needs only the residual -- the right-hand
// side of the next solve -- and seven scalars back; the line search or adaptive logic can test a
// trial without pulling ifv.
//
// res(i) is evaluated with the reference's expression and operation order, without contraction
// (bit-identical to the reference's for the same ifv).  The sums (norm of the residual, mgload,
// sum_loads) are two-stage tree sums in a fixed order: reproducible run to run, equal to the
// reference's sequential sums to rounding (~1e-15 relative); num_term_loads and the dof of the
// largest residual (lowest dof among equals, as the reference's strict '>' scan) are exact.
#pragma once

namespace {

struct ResidualArgs {
  int nodof, iter, have_mpc;
  double nfac, nfac1, nfac2;
  const int32_t *cstmap;
  const double *pbar, *load, *mdiag, *vel, *acc, *cf, *dlag, *du, *ifv;
  double *res;
  ResidualPartial *part;
  const unsigned char *owned;   // several domains, neighbour exchange: the dofs this rank owns (else null)
};

__device__ __forceinline__ double ld0(const double *v, int i) { return v ? v[i] : 0.0; }

__device__ __forceinline__ void combine(ResidualPartial &a, const ResidualPartial &b) {
  a.sum_res2 = __dadd_rn(a.sum_res2, b.sum_res2);
  a.mgload2 = __dadd_rn(a.mgload2, b.mgload2);
  a.sum_loads = __dadd_rn(a.sum_loads, b.sum_loads);
  a.num_term_loads += b.num_term_loads;
  // cvtest.f:70-73: strict '>' on an ascending scan == largest |res|, lowest dof among equals
  if (b.max_dof != 0 && (a.max_dof == 0 || b.max_abs > a.max_abs || (b.max_abs == a.max_abs && b.max_dof < a.max_dof))) {
    a.max_abs = b.max_abs;
    a.max_dof = b.max_dof;
  }
}

__device__ __forceinline__ ResidualPartial shfl_down(const ResidualPartial &p, int d) {
  ResidualPartial q;
  q.sum_res2 = __shfl_down_sync(0xffffffffu, p.sum_res2, d);
  q.mgload2 = __shfl_down_sync(0xffffffffu, p.mgload2, d);
  q.sum_loads = __shfl_down_sync(0xffffffffu, p.sum_loads, d);
  q.max_abs = __shfl_down_sync(0xffffffffu, p.max_abs, d);
  q.num_term_loads = __shfl_down_sync(0xffffffffu, p.num_term_loads, d);
  q.max_dof = __shfl_down_sync(0xffffffffu, p.max_dof, d);
  q.pad = 0;
  q.res_at_max = 0.0;
  return q;
}

__device__ __forceinline__ void cta_reduce(ResidualPartial &p, ResidualPartial *out) {
  __shared__ ResidualPartial s[RES_TPB / 32];
  for (int d = 16; d >= 1; d >>= 1) {
    const ResidualPartial q = shfl_down(p, d);
    combine(p, q);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    ResidualPartial t = s[0];
    for (int k = 1; k < RES_TPB / 32; ++k) combine(t, s[k]);
    *out = t;
  }
}

// upres_iter_0 / upres + the residual loop of cvtest, one thread per dof (grid-stride)
__global__ void __launch_bounds__(RES_TPB) k_residual(const ResidualArgs a) {
  ResidualPartial p{0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0.0};
  for (int i = blockIdx.x * RES_TPB + threadIdx.x; i < a.nodof; i += gridDim.x * RES_TPB) {
    if (a.owned && !a.owned[i]) {   // another rank's dof: its internal force is not complete here
      a.res[i] = 0.0;
      continue;
    }
    const bool constrained = a.cstmap[i] != 0;
    const double du = a.du[i], ifv = a.ifv[i], md = ld0(a.mdiag, i), cf = ld0(a.cf, i);
    const double inertia = __dmul_rn(__dmul_rn(md, du), a.nfac);
    double r;
    if (a.iter == 0) {
      // upres.f:63-68
      r = constrained ? 0.0 : __dsub_rn(__dadd_rn(__dsub_rn(a.pbar[i], inertia), cf), ifv);
    } else {
      // upres.f:170-180
      double ext = __dadd_rn(a.pbar[i], cf);
      if (a.have_mpc) ext = __dadd_rn(ext, a.dlag[i]);
      r = __dsub_rn(__dsub_rn(ext, inertia), ifv);
      // upres.f:203-218
      const double accel_n1 = __dsub_rn(__dsub_rn(__dmul_rn(du, a.nfac), __dmul_rn(a.nfac1, ld0(a.acc, i))),
                                        __dmul_rn(a.nfac2, ld0(a.vel, i)));
      const double inertia_contact = __dadd_rn(__dmul_rn(fabs(accel_n1), md), fabs(cf));
      double term_load;
      if (!constrained) {
        term_load = __dadd_rn(inertia_contact, fabs(ld0(a.load, i)));
      } else {
        term_load = __dadd_rn(inertia_contact, fabs(ifv));
        r = 0.0;
      }
      if (term_load != 0.0) {
        p.mgload2 = __dadd_rn(p.mgload2, __dmul_rn(term_load, term_load));
        p.num_term_loads += 1;
        p.sum_loads = __dadd_rn(p.sum_loads, fabs(term_load));
      }
    }
    a.res[i] = r;
    // cvtest.f:62-75
    double rf = r;
    if (a.have_mpc && a.dlag[i] != 0.0) rf = 0.0;
    p.sum_res2 = __dadd_rn(p.sum_res2, __dmul_rn(rf, rf));
    const double ab = fabs(rf);
    if (p.max_dof == 0 || ab > p.max_abs) {   // ascending i within a thread: keeps the lowest dof
      p.max_abs = ab;
      p.max_dof = i + 1;
    }
  }
  cta_reduce(p, a.part + blockIdx.x);
}

__global__ void __launch_bounds__(RES_TPB) k_residual_finish(ResidualPartial *part, int n, const double *res) {
  ResidualPartial p{0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0.0};
  for (int k = threadIdx.x; k < n; k += RES_TPB) combine(p, part[k]);
  cta_reduce(p, part + RES_MAX_CTAS);
  if (threadIdx.x == 0) {   // cvtest.f:77: the signed residual at that dof
    ResidualPartial &t = part[RES_MAX_CTAS];
    t.res_at_max = t.max_dof >= 1 ? res[t.max_dof - 1] : 0.0;
  }
}

template <typename T>
int res_upload(w3d_engine *h, T **dst, const T *src, size_t n) {
  if (!src) {
    cudaFree(*dst);
    *dst = nullptr;
    return W3D_OK;
  }
  if (!*dst && cudaMalloc((void **)dst, n * sizeof(T)) != cudaSuccess)
    return h->fail(W3D_ERR_CUDA, "residual: device allocation of %zu bytes failed", n * sizeof(T));
  if (cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
    return h->fail(W3D_ERR_CUDA, "residual: host to device copy failed");
  return W3D_OK;
}

// several domains: the norms of the structure from the ranks' owned parts.  [0..3] = the three sums
// and the number of load terms (added by ncclAllReduce), then one (max |res|, dof, res) triple per
// rank, zero except the rank's own (the same all-reduce gathers them)
__global__ void k_residual_pack(const ResidualPartial *p, double *red, int rank, int nranks) {
  for (int k = threadIdx.x; k < 4 + 3 * nranks; k += blockDim.x) red[k] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    red[0] = p->sum_res2; red[1] = p->mgload2; red[2] = p->sum_loads; red[3] = (double)p->num_term_loads;
    red[4 + 3 * rank] = p->max_abs; red[5 + 3 * rank] = (double)p->max_dof; red[6 + 3 * rank] = p->res_at_max;
  }
}
__global__ void k_residual_unpack(const double *red, int nranks, ResidualPartial *p) {
  ResidualPartial t{red[0], red[1], red[2], 0.0, (long long)red[3], 0, 0, 0.0};
  for (int r = 0; r < nranks; ++r) {     // cvtest.f:70-73 over the whole vector: largest, lowest dof among equals
    const double ab = red[4 + 3 * r];
    const int dof = (int)red[5 + 3 * r];
    if (dof != 0 && (t.max_dof == 0 || ab > t.max_abs || (ab == t.max_abs && dof < t.max_dof))) {
      t.max_abs = ab; t.max_dof = dof; t.res_at_max = red[6 + 3 * r];
    }
  }
  *p = t;
}

}  // namespace

extern "C" {

int32_t w3d_residual_set_loads(w3d_engine *h, const w3d_residual_loads *in) {
  if (!h || !in) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  if (!in->cstmap || !in->pbar) return h->fail(W3D_ERR_ARG, "residual: cstmap and pbar are required");
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  ResidualData &R = h->residual;
  const size_t n = (size_t)h->cfg.nodof;
  int rc;
  if ((rc = res_upload(h, &R.d_cstmap, in->cstmap, n))) return rc;
  if ((rc = res_upload(h, &R.d_pbar, in->pbar, n))) return rc;
  if ((rc = res_upload(h, &R.d_load, in->load, n))) return rc;
  if ((rc = res_upload(h, &R.d_mdiag, in->mdiag, n))) return rc;
  if ((rc = res_upload(h, &R.d_vel, in->velocity, n))) return rc;
  if ((rc = res_upload(h, &R.d_acc, in->accel, n))) return rc;
  if ((rc = res_upload(h, &R.d_cf, in->contact_force, n))) return rc;
  if ((rc = res_upload(h, &R.d_dlag, in->d_lagrange_forces, n))) return rc;
  if (!R.d_res) CUDA_TRY(cudaMalloc((void **)&R.d_res, n * sizeof(double)));
  if (!R.d_part) CUDA_TRY(cudaMalloc((void **)&R.d_part, (RES_MAX_CTAS + 1) * sizeof(ResidualPartial)));
  if (!R.h_part) CUDA_TRY(cudaMallocHost((void **)&R.h_part, sizeof(ResidualPartial)));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // the caller's buffers are free to change now
  R.ready = true;
  return W3D_OK;
}

int32_t w3d_residual(w3d_engine *h, int32_t iter, double dt, double nbeta, double *res,
                     w3d_residual_norms *out) {
  if (!h) return W3D_ERR_ARG;
  if (!out) return h->fail(W3D_ERR_ARG, "residual: norms output is required");
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  ResidualData &R = h->residual;
  if (!R.ready) return h->fail(W3D_ERR_ARG, "w3d_residual before w3d_residual_set_loads");
  if (!h->have_vectors) return h->fail(W3D_ERR_ARG, "w3d_residual before any stress pass (no internal forces)");
  if (iter < 0 || !(dt > 0.0) || !(nbeta > 0.0)) return h->fail(W3D_ERR_ARG, "residual: bad iter / dt / nbeta");
  CUDA_TRY(cudaSetDevice(h->device));
  { const int rj = comm_join(h); if (rj) return rj; }   // the internal forces must be complete
  ResidualArgs a{};
  a.nodof = h->cfg.nodof;
  a.iter = iter;
  a.have_mpc = R.d_dlag ? 1 : 0;
  a.nfac = 1.0 / (nbeta * dt * dt);
  a.nfac1 = (1.0 - 2.0 * nbeta) / (2.0 * nbeta);
  a.nfac2 = 1.0 / (nbeta * dt);
  a.cstmap = R.d_cstmap;
  a.pbar = R.d_pbar; a.load = R.d_load; a.mdiag = R.d_mdiag; a.vel = R.d_vel; a.acc = R.d_acc;
  a.cf = R.d_cf; a.dlag = R.d_dlag; a.du = h->d_du; a.ifv = h->d_ifv;
  a.res = R.d_res;
  a.part = R.d_part;
  // several domains after a neighbour exchange: every rank evaluates the dofs it owns (their
  // internal forces are complete there) and the norms are summed over the ranks; after the
  // all-reduce mode every rank holds the whole vector and evaluates all of it
  const bool split = h->comm.comm && h->comm.exchange == 1 && h->comm.d_owned;
  a.owned = split ? h->comm.d_owned : nullptr;
  const int ncta = std::min((a.nodof + RES_TPB - 1) / RES_TPB, RES_MAX_CTAS);
  k_residual<<<ncta, RES_TPB, 0, h->stream>>>(a);
  k_residual_finish<<<1, RES_TPB, 0, h->stream>>>(R.d_part, ncta, R.d_res);
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  if (split) {
    CommData &cm = h->comm;
    k_residual_pack<<<1, 64, 0, h->stream>>>(R.d_part + RES_MAX_CTAS, cm.d_rred, cm.rank, cm.nranks);
    NCCL_TRY(nccl_api()->all_reduce(cm.d_rred, cm.d_rred, 4 + 3 * (size_t)cm.nranks, NCCL_FLOAT64, NCCL_SUM, cm.comm, h->stream));
    k_residual_unpack<<<1, 1, 0, h->stream>>>(cm.d_rred, cm.nranks, R.d_part + RES_MAX_CTAS);
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaMemcpyAsync(R.h_part, R.d_part + RES_MAX_CTAS, sizeof(ResidualPartial), cudaMemcpyDeviceToHost, h->stream));
  if (res) CUDA_TRY(cudaMemcpyAsync(res, R.d_res, sizeof(double) * a.nodof, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  const ResidualPartial &p = *R.h_part;
  out->norm_res = std::sqrt(p.sum_res2);
  out->mgload = std::sqrt(p.mgload2);
  out->sum_loads = p.sum_loads;
  out->num_term_loads = p.num_term_loads;
  out->res_max_dof = p.max_dof;
  out->res_max = p.res_at_max;
  return W3D_OK;
}

int32_t w3d_residual_device(w3d_engine *h, double **d_res) {
  if (!h || !d_res) return W3D_ERR_ARG;
  if (!h->residual.ready) return h->fail(W3D_ERR_ARG, "w3d_residual_device before w3d_residual_set_loads");
  *d_res = h->residual.d_res;
  return W3D_OK;
}

}  // extern "C"
