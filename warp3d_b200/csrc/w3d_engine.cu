// w3d_engine.cu -- host side of the B200 element-block engine and its C ABI (include/w3d_engine.h).
//
// The engine mirrors the reference's block contract on the host (elblks, props, edest/cdest,
// history_blk_list) and keeps all blocked state resident in HBM as structure-of-arrays.  Blocks
// that share (material model, geonl, bbar, history size) form a *segment*; every pass is one
// kernel launch per segment.  There is no CPU code path for the arithmetic: if CUDA is not
// usable every entry point returns W3D_ERR_CUDA.
#include "../../include/w3d_engine.h"

#include <dlfcn.h>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "w3d_stress_kernel.cuh"
#include "w3d_tanstf_kernel.cuh"
#include "w3d_cp_kernel.h"
#include "w3d_mm03_kernel.h"
#include "w3d_assemble_data.h"
#include "w3d_residual_data.h"
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <type_traits>
#include <utility>

using namespace w3d;

namespace {

thread_local std::string g_create_error;

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      return fail(W3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                  __FILE__, __LINE__);                                                          \
    }                                                                                           \
  } while (0)

struct Segment {
  int mat_type = 0, hist_size = 0;
  bool geonl = false, bbar = false;
  int nslots = 0, ncta = 0;
  long long ng = 0;                 // nslots * 8
  std::vector<int> blocks;          // 0-based block numbers in block order
  // device
  CtaDesc *d_cta = nullptr;
  int32_t *d_edest = nullptr, *d_cdest = nullptr;
  float *d_eprops = nullptr;
  double *eps_n = nullptr, *eps_n1 = nullptr, *urcs_n = nullptr, *urcs_n1 = nullptr,
         *rot_n = nullptr, *rot_n1 = nullptr, *cep = nullptr, *hist_n = nullptr,
         *hist_n1 = nullptr, *elem_vol = nullptr, *partials = nullptr;
  // crystal plasticity (mm10)
  double *d_cprops = nullptr, *scratch = nullptr;
  int mm03_allow_cut = 0;
  int *d_mm03_active = nullptr;      // mm03: [slot / 16][8] live-point flags of the element blocks (k_mm03_flags)
  int ncrystals = 0, nslip = 0, num_hard = 0, o_gradfe = 0, o_Rp0 = 0;
  bool cp_fast = false, cp_taul = false;   // fast lattice-frame kernel (w3d_mm10_fast.h); k_0 != 0
  std::vector<char> cp_consts;             // its crystal-type table (host copy, uploaded per launch)
  Mm10Schedule *cp_sched = nullptr;        // element order of k_mm10f (w3d_cp_kernel.h); W3D_MM10_SCHEDULE=0: block order
};

struct BlockInfo {
  int span = 0, felem = 0, domain = 0;
  bool owned = false;
  int seg = -1;
  int slot0 = 0;        // segment-relative first slot
  long long epos0 = 0;  // unpadded element position among owned blocks
  int elem_type = 0, nnode = 0, ndof = 0, int_order = 0, ngp = 0, mat_type = 0, matnum = 0,
      hist_size = 0, totdof = 0, utsz = 0;
  bool geonl = false, bbar = false;
  long long tab_off = 0;  // offset (ints) of this block's edest/cdest table in the caller's arrays
};

}  // namespace

// exchange step between domains (w3d_comm.cuh)
struct CommData {
  void *comm = nullptr;              // ncclComm_t
  int rank = 0, nranks = 1;
  int exchange = 1, bcast_inputs = 0;
  std::vector<long long> counts;     // dofs shared with each rank (host, from create)
  std::vector<int32_t> dofs;         // concatenated lists, 1-based
  long long ntot = 0;
  int32_t *d_idx = nullptr;
  double *d_send = nullptr, *d_recv = nullptr, *d_red = nullptr;   // d_red[8]
  // dofs this rank OWNS (lowest domain among those whose blocks touch the dof; untouched dofs go
  // to rank 0): the ranks' owned sets tile 1..nodof, so sums over owned dofs all-reduce to the
  // structure's (w3d_residual after a neighbour exchange)
  std::vector<uint8_t> owned;        // host, from create
  uint8_t *d_owned = nullptr;
  double *d_rred = nullptr;          // 4 sums + 3 doubles per rank for the residual norms
  // the exchange runs on its own stream so that the stiffness pass (which reads neither ifv nor
  // the scalars) overlaps it; whoever touches ifv / the scalars next joins first (comm_join)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_done = nullptr;
  bool pending = false;
};

struct w3d_engine {
  w3d_engine_config cfg{};
  CommData comm;
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<BlockInfo> blocks;
  std::vector<Segment> segs;
  long long n_owned_elems = 0;
  int dof_lo = 0, dof_hi = 0;       // [dof_lo, dof_hi): the structure dofs this engine's blocks touch
  double *d_c = nullptr, *d_u = nullptr, *d_du = nullptr, *d_ifv = nullptr;
  double *d_estiff = nullptr;       // (300, n_owned_elems)
  double *d_scal = nullptr;         // [4]
  int *d_flags = nullptr;           // [2]
  double *d_stage = nullptr;        // push/pull staging (one block)
  size_t stage_doubles = 0;
  double *h_pin = nullptr;          // pinned: [0..3] scalars, then 2 ints of flags, [8..15] reduced scalars
  int64_t launches = 0;
  double dt = 1.0;
  w3d_scalars last{};
  bool have_vectors = false;
  AssemblyData assembly;            // assembled-equations path (w3d_assemble.cuh)
  ResidualData residual;            // structure residual and norms (w3d_residual.cuh)

  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

namespace {

inline int32_t iprop(const w3d_engine_config &c, int row, int elem) {
  int32_t v;
  std::memcpy(&v, (const char *)c.props + 4 * ((size_t)(elem - 1) * c.mxelpr + (row - 1)), 4);
  return v;
}
inline float rprop(const w3d_engine_config &c, int row, int elem) {
  float v;
  std::memcpy(&v, (const char *)c.props + 4 * ((size_t)(elem - 1) * c.mxelpr + (row - 1)), 4);
  return v;
}
// ifort logicals: .true. has the low bit set
inline bool lprop(const w3d_engine_config &c, int row, int elem) { return (iprop(c, row, elem) & 1) != 0; }

#include "w3d_comm.cuh"   // NCCL binding, neighbour plan, pack / add kernels

// quad2 (getgpts.f:642-718) and deriv2 (derivs.f:283-343): WARP3D point and node ordering
void shape_tables(double (*tab)[8][8]) {
  const double l1 = 0.5773502691896257645;
  static const int sx[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  static const int sy[8] = {-1, 1, -1, 1, -1, 1, -1, 1};
  static const int sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int gp = 0; gp < 8; ++gp) {
    const double xi = sx[gp] * l1, eta = sy[gp] * l1, zeta = sz[gp] * l1;
    const double one = 1.0, one25 = 0.125;
    const double xp = one + xi, ep = one + eta, zp = one + zeta;
    const double xm_ = one - xi, em = one - eta, zm = one - zeta;
    double nxi[8], neta[8], nzeta[8];
    nxi[0] = -em * zp * one25; neta[0] = -xm_ * zp * one25; nzeta[0] = xm_ * em * one25;
    nxi[1] = -em * zm * one25; neta[1] = -xm_ * zm * one25; nzeta[1] = -xm_ * em * one25;
    nxi[2] = -ep * zm * one25; neta[2] = xm_ * zm * one25;  nzeta[2] = -xm_ * ep * one25;
    nxi[3] = -ep * zp * one25; neta[3] = xm_ * zp * one25;  nzeta[3] = xm_ * ep * one25;
    nxi[4] = em * zp * one25;  neta[4] = -xp * zp * one25;  nzeta[4] = xp * em * one25;
    nxi[5] = em * zm * one25;  neta[5] = -xp * zm * one25;  nzeta[5] = -xp * em * one25;
    nxi[6] = ep * zm * one25;  neta[6] = xp * zm * one25;   nzeta[6] = -xp * ep * one25;
    nxi[7] = ep * zp * one25;  neta[7] = xp * zp * one25;   nzeta[7] = xp * ep * one25;
    for (int j = 0; j < 8; ++j) { tab[0][j][gp] = nxi[j]; tab[1][j][gp] = neta[j]; tab[2][j][gp] = nzeta[j]; }
  }
}

// ---- layout conversion kernels: one reference block <-> the segment's SoA arrays ---------------
// reference (nprm, ngp=8, span)  <->  soa[k*ng + (slot0+i)*8 + gp]
__global__ void k_block_to_soa(const double *blk, double *soa, long long ng, int slot0, int span,
                               int nprm, int cep_layout) {
  const int n = nprm * 8 * span;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    int k, gp, i;
    if (cep_layout) { k = t % nprm; i = (t / nprm) % span; gp = t / (nprm * span); }   // (21,span,ngp)
    else { k = t % nprm; gp = (t / nprm) % 8; i = t / (nprm * 8); }                    // (nprm,ngp,span)
    soa[(long long)k * ng + (long long)(slot0 + i) * 8 + gp] = blk[t];
  }
}
__global__ void k_soa_to_block(double *blk, const double *soa, long long ng, int slot0, int span,
                               int nprm, int cep_layout) {
  const int n = nprm * 8 * span;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    int k, gp, i;
    if (cep_layout) { k = t % nprm; i = (t / nprm) % span; gp = t / (nprm * span); }
    else { k = t % nprm; gp = (t / nprm) % 8; i = t / (nprm * 8); }
    blk[t] = soa[(long long)k * ng + (long long)(slot0 + i) * 8 + gp];
  }
}
// R = I for every Gauss point (rotation_init, reopen.f:1841-1960)
__global__ void k_rot_identity(double *rot, long long ng) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ng;
       t += (long long)gridDim.x * blockDim.x) {
    rot[0 * ng + t] = 1.0; rot[4 * ng + t] = 1.0; rot[8 * ng + t] = 1.0;
  }
}

template <typename T>
cudaError_t dalloc(T **p, size_t n) {
  cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, (n ? n : 1) * sizeof(T));
  return e;
}

int nprm_of(const w3d_engine *h, int which, const BlockInfo &b) {
  (void)h;
  switch (which) {
    case W3D_HIST_N: case W3D_HIST_N1: return b.hist_size;
    case W3D_EPS_N: case W3D_EPS_N1: return 6;
    case W3D_URCS_N: case W3D_URCS_N1: return 9;
    case W3D_ROT_N: case W3D_ROT_N1: return 9;
    case W3D_CEP: return 21;
    default: return 0;
  }
}
double *soa_of(Segment &s, int which) {
  switch (which) {
    case W3D_HIST_N: return s.hist_n;   case W3D_HIST_N1: return s.hist_n1;
    case W3D_EPS_N: return s.eps_n;     case W3D_EPS_N1: return s.eps_n1;
    case W3D_URCS_N: return s.urcs_n;   case W3D_URCS_N1: return s.urcs_n1;
    case W3D_ROT_N: return s.rot_n;     case W3D_ROT_N1: return s.rot_n1;
    case W3D_CEP: return s.cep;
    default: return nullptr;
  }
}
long long block_doubles(int which, const BlockInfo &b) {
  switch (which) {
    case W3D_HIST_N: case W3D_HIST_N1: return (long long)b.hist_size * b.ngp * b.span;
    case W3D_EPS_N: case W3D_EPS_N1: return 6LL * b.ngp * b.span;
    case W3D_URCS_N: case W3D_URCS_N1: case W3D_ROT_N: case W3D_ROT_N1: return 9LL * b.ngp * b.span;
    case W3D_CEP: return 21LL * b.ngp * b.span;
    case W3D_ESTIFF: return (long long)b.utsz * b.span;
    case W3D_ELEM_VOL: return b.span;
    default: return -1;
  }
}

void free_all(w3d_engine *h) {
  for (auto &s : h->segs) {
    cudaFree(s.d_cta); cudaFree(s.d_edest); cudaFree(s.d_cdest); cudaFree(s.d_eprops);
    cudaFree(s.eps_n); cudaFree(s.eps_n1); cudaFree(s.urcs_n); cudaFree(s.urcs_n1);
    cudaFree(s.rot_n); cudaFree(s.rot_n1); cudaFree(s.cep); cudaFree(s.hist_n);
    cudaFree(s.hist_n1); cudaFree(s.elem_vol); cudaFree(s.partials);
    cudaFree(s.d_cprops); cudaFree(s.scratch); cudaFree(s.d_mm03_active);
    mm10_schedule_destroy(s.cp_sched); s.cp_sched = nullptr;
  }
  h->assembly.release();
  h->residual.release();
  cudaFree(h->d_c); cudaFree(h->d_u); cudaFree(h->d_ifv);
  cudaFree(h->comm.d_idx); cudaFree(h->comm.d_send); cudaFree(h->comm.d_recv); cudaFree(h->comm.d_red);
  cudaFree(h->comm.d_owned); cudaFree(h->comm.d_rred);
  if (h->comm.stream) cudaStreamDestroy(h->comm.stream);
  if (h->comm.ev_start) cudaEventDestroy(h->comm.ev_start);
  if (h->comm.ev_done) cudaEventDestroy(h->comm.ev_done);
  if (h->comm.comm && nccl_api()->comm_destroy) nccl_api()->comm_destroy(h->comm.comm);
  cudaFree(h->d_estiff); cudaFree(h->d_scal); cudaFree(h->d_flags); cudaFree(h->d_stage);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->stream) cudaStreamDestroy(h->stream);
}

template <int MAT, int PHASE>
cudaError_t launch_stress(const Segment &sg, const StressArgs &a0, cudaStream_t st) {
  StressArgs a = a0;
  a.ncta = sg.ncta;                                                  // 16-element groups
  const int grid = (sg.ncta + STRESS_GROUPS - 1) / STRESS_GROUPS;    // CTAs of STRESS_GROUPS groups
  if (sg.geonl) {
    if (sg.bbar) return launch_stress_kernel(k_eps_sig_ifv<MAT, true, true, PHASE>, grid, a, st);
    return launch_stress_kernel(k_eps_sig_ifv<MAT, true, false, PHASE>, grid, a, st);
  }
  if (sg.bbar) return launch_stress_kernel(k_eps_sig_ifv<MAT, false, true, PHASE>, grid, a, st);
  return launch_stress_kernel(k_eps_sig_ifv<MAT, false, false, PHASE>, grid, a, st);
}
template <int MAT, int PHASE>
cudaError_t stress_prepare_device() {
  cudaError_t e = stress_opt_in(k_eps_sig_ifv<MAT, true, true, PHASE>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<MAT, true, false, PHASE>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<MAT, false, true, PHASE>);
  if (e == cudaSuccess) e = stress_opt_in(k_eps_sig_ifv<MAT, false, false, PHASE>);
  return e;
}

int create_impl(w3d_engine *h) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  const w3d_engine_config &c = h->cfg;
  if (c.abi_version != W3D_ABI_VERSION) return fail(W3D_ERR_CONFIG, "abi_version %d != %d", c.abi_version, W3D_ABI_VERSION);
  if (!c.elblks || !c.props || !c.c || !c.edest || !c.cdest || !c.history_blk_list)
    return fail(W3D_ERR_ARG, "null table in w3d_engine_config");
  if (c.nelblk <= 0 || c.noelem <= 0 || c.nodof <= 0 || c.mxelpr < 38)
    return fail(W3D_ERR_CONFIG, "bad sizes nelblk=%d noelem=%d nodof=%d mxelpr=%d", c.nelblk, c.noelem, c.nodof, c.mxelpr);

  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return fail(W3D_ERR_CUDA, "no CUDA device");
  if (c.device >= 0) { CUDA_TRY(cudaSetDevice(c.device)); }
  CUDA_TRY(cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) return fail(W3D_ERR_CUDA, "device %d is sm_%d%d; this engine is built for sm_100a only", h->device, prop.major, prop.minor);
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

  // ---- block table (do_nleps_block reads the same props rows, :448-461) and segments
  h->blocks.resize(c.nelblk);
  long long tab = 0, epos = 0;
  for (int blk = 0; blk < c.nelblk; ++blk) {
    BlockInfo &b = h->blocks[blk];
    b.span = c.elblks[4 * blk + 0];
    b.felem = c.elblks[4 * blk + 1];
    b.domain = c.elblks[4 * blk + 2];
    if (b.span <= 0 || b.felem < 1 || b.felem + b.span - 1 > c.noelem)
      return fail(W3D_ERR_CONFIG, "block %d: span %d felem %d outside 1..%d", blk + 1, b.span, b.felem, c.noelem);
    if (c.mxvl > 0 && b.span > c.mxvl) return fail(W3D_ERR_CONFIG, "block %d: span %d > mxvl %d", blk + 1, b.span, c.mxvl);
    const int fe = b.felem;
    b.elem_type = iprop(c, 1, fe); b.nnode = iprop(c, 2, fe); b.ndof = iprop(c, 4, fe);
    b.int_order = iprop(c, 5, fe); b.ngp = iprop(c, 6, fe);
    b.geonl = lprop(c, 18, fe); b.bbar = lprop(c, 19, fe);
    b.mat_type = iprop(c, 25, fe); b.matnum = iprop(c, 38, fe);
    b.totdof = b.nnode * b.ndof;
    b.utsz = (b.totdof * b.totdof - b.totdof) / 2 + b.totdof;
    b.hist_size = c.history_blk_list[blk];
    b.tab_off = tab;
    tab += (long long)b.totdof * b.span;
    b.owned = (c.numprocs <= 1) || (b.domain == c.myid);
    if (!b.owned) continue;
    if (b.elem_type != 2 || b.nnode != 8 || b.ndof != 3 || b.ngp != 8)
      return fail(W3D_ERR_UNSUPPORTED, "block %d: element type %d (%d nodes, %d points): only hex8 l3disop 2x2x2", blk + 1, b.elem_type, b.nnode, b.ngp);
    if (b.mat_type != 1 && b.mat_type != 3 && b.mat_type != 5 && b.mat_type != 10)
      return fail(W3D_ERR_UNSUPPORTED, "block %d: material model %d not built yet (mm01, mm03, mm05, mm10)", blk + 1, b.mat_type);
    if (b.mat_type == 3) {
      // the mises path of mm03: no voids, no nucleation, no segmental curves (mm03.f:2163, rknstr.f:816-819)
      if (b.hist_size != 15) return fail(W3D_ERR_CONFIG, "block %d: mm03 history size %d != 15", blk + 1, b.hist_size);
      for (int i = 0; i < b.span; ++i) {
        const int el = b.felem + i;
        if (rprop(c, 26, el) > 0.0f || (iprop(c, 30, el) & 1))
          return fail(W3D_ERR_UNSUPPORTED, "block %d element %d: Gurson material (f0 > 0 or nucleation) not built; mm03 mises only", blk + 1, el);
        if (iprop(c, 24, el) & 4)
          return fail(W3D_ERR_UNSUPPORTED, "block %d element %d: segmental stress-strain curves not built", blk + 1, el);
      }
    }
    if (b.mat_type == 5) {
      if (b.hist_size != 14) return fail(W3D_ERR_CONFIG, "block %d: mm05 history size %d != 14", blk + 1, b.hist_size);
      if (!c.matprp || b.matnum < 1 || b.matnum > c.nmat || c.mxmtpr < 64)
        return fail(W3D_ERR_CONFIG, "block %d: mm05 needs matprp(55:64) of material %d", blk + 1, b.matnum);
      // matprp(58) > 0: nonlinear_hardening, < 0: generalized_plasticity (inmat.f:2220, :2272)
      if (c.matprp[(size_t)(b.matnum - 1) * c.mxmtpr + 57] == 0.0f)
        return fail(W3D_ERR_CONFIG, "block %d: mm05 option flag matprp(58) of material %d is not set", blk + 1, b.matnum);
    }
    if (b.mat_type == 1 && b.hist_size != 11)
      return fail(W3D_ERR_CONFIG, "block %d: mm01 history size %d != 11", blk + 1, b.hist_size);
    if (b.mat_type == 10) {
      if (!c.crystal_types || !c.elem_cnum || !c.elem_angles || c.cp_ncrystals < 1 || c.n_crystal_types < 1)
        return fail(W3D_ERR_CONFIG, "block %d uses mm10 but no crystal tables were given", blk + 1);
      const double *ct0 = c.crystal_types;
      const int nslip = (int)ct0[337], num_hard = (int)ct0[338], h_type = (int)ct0[339];
      if (nslip < 1 || nslip > 48 || num_hard != 1 || h_type != 1)
        return fail(W3D_ERR_UNSUPPORTED, "block %d: mm10 with nslip %d, num_hard %d, hardening %d (built: <=48 systems, Voce)", blk + 1, nslip, num_hard, h_type);
      for (int t = 1; t < c.n_crystal_types; ++t)      // one kernel instantiation per segment
        if (((int)c.crystal_types[(size_t)t * W3D_CRYSTAL_TYPE_DOUBLES + 337] <= 12) != (nslip <= 12))
          return fail(W3D_ERR_UNSUPPORTED, "mm10: crystal types with <=12 and with >12 slip systems in one model");
      if (b.hist_size != mm10_hist_size(nslip, num_hard, c.cp_ncrystals))
        return fail(W3D_ERR_CONFIG, "block %d: mm10 history size %d != %d", blk + 1, b.hist_size, mm10_hist_size(nslip, num_hard, c.cp_ncrystals));
    }
    int sidx = -1;
    for (size_t s = 0; s < h->segs.size(); ++s) {
      const Segment &sg = h->segs[s];
      if (sg.mat_type == b.mat_type && sg.geonl == b.geonl && sg.bbar == b.bbar && sg.hist_size == b.hist_size) { sidx = (int)s; break; }
    }
    if (sidx < 0) {
      Segment sg;
      sg.mat_type = b.mat_type; sg.geonl = b.geonl; sg.bbar = b.bbar; sg.hist_size = b.hist_size;
      h->segs.push_back(sg);
      sidx = (int)h->segs.size() - 1;
    }
    Segment &sg = h->segs[sidx];
    b.seg = sidx;
    b.slot0 = sg.nslots;
    b.epos0 = epos;
    sg.nslots += ((b.span + EPC - 1) / EPC) * EPC;
    sg.blocks.push_back(blk);
    epos += b.span;
  }
  h->n_owned_elems = epos;
  {
    // the internal-force vector is only ever written inside the dof range of the owned blocks
    // (a slab of the structure when there are several domains): only that range is zeroed per pass
    int lo = c.nodof, hi = 0;
    for (const BlockInfo &b : h->blocks) {
      if (!b.owned) continue;
      for (long long k = 0; k < (long long)b.totdof * b.span; ++k) {
        const int32_t e = c.edest[b.tab_off + k];
        if (e >= 1 && e <= c.nodof) { lo = std::min(lo, e - 1); hi = std::max(hi, (int)e); }
      }
    }
    h->dof_lo = hi > lo ? lo : 0;
    h->dof_hi = hi > lo ? hi : 0;
  }
  if (c.numprocs > 1) {
    // dofs shared with every other domain (used by the neighbour exchange after w3d_comm_init)
    std::string perr;
    const int prc = comm_plan(c, c.myid, h->comm.counts, h->comm.dofs, perr);
    if (prc != W3D_OK && prc != W3D_ERR_UNSUPPORTED) return fail(prc, "interface plan: %s", perr.c_str());
    if (prc == W3D_ERR_UNSUPPORTED) { h->comm.counts.clear(); h->comm.dofs.clear(); h->comm.exchange = 0; }
    h->comm.owned.assign((size_t)c.nodof, 0);
    const int orc = w3d_comm_owned(&c, c.myid, h->comm.owned.data());
    if (orc != W3D_OK) return fail(orc, "ownership of the dofs: block table inconsistent");
  }

  // ---- global vectors
  CUDA_TRY(dalloc(&h->d_c, (size_t)c.nodof));
  CUDA_TRY(dalloc(&h->d_u, 2 * (size_t)c.nodof));     // u and du contiguous: one broadcast moves both
  h->d_du = h->d_u + c.nodof;
  CUDA_TRY(dalloc(&h->d_ifv, (size_t)c.nodof));
  CUDA_TRY(cudaMemcpy(h->d_c, c.c, sizeof(double) * c.nodof, cudaMemcpyHostToDevice));
  CUDA_TRY(dalloc(&h->d_estiff, (size_t)h->n_owned_elems * 300));
  CUDA_TRY(dalloc(&h->d_scal, 4));
  CUDA_TRY(dalloc(&h->d_flags, 64));   // [0] cut, [1] fatal, [2..] debug histogram (W3D_MM10F_HIST builds)
  CUDA_TRY(cudaMallocHost((void **)&h->h_pin, 16 * sizeof(double)));
  {
    double tab3[3][8][8];
    shape_tables(tab3);
    CUDA_TRY(cudaMemcpyToSymbol(c_shape, tab3, sizeof tab3));
    CUDA_TRY(mm05_set_shape_table(&tab3[0][0][0], sizeof tab3));
    CUDA_TRY(mm03_set_shape_table(&tab3[0][0][0], sizeof tab3));
  }

  // ---- per-segment tables and state
  size_t max_blk_doubles = 0;
  for (auto &sg : h->segs) {
    sg.ng = (long long)sg.nslots * 8;
    sg.ncta = sg.nslots / EPC;
    std::vector<CtaDesc> cta(sg.ncta);
    std::vector<int32_t> ed((size_t)sg.nslots * 24), cdv((size_t)sg.nslots * 24);
    std::vector<float> ep((size_t)sg.nslots * EPROP, 0.0f);
    int ci = 0;
    for (int blk : sg.blocks) {
      const BlockInfo &b = h->blocks[blk];
      for (int i0 = 0; i0 < b.span; i0 += EPC) {
        CtaDesc &d = cta[ci++];
        d.slot0 = b.slot0 + i0;
        d.nvalid = (b.span - i0 < EPC) ? b.span - i0 : EPC;
        d.blk_slot0 = b.slot0;
        d.epos0 = (int)(b.epos0 + i0);
        d.blk_span = b.span;
      }
      const int padded = ((b.span + EPC - 1) / EPC) * EPC;
      for (int i = 0; i < padded; ++i) {
        const int src = (i < b.span) ? i : b.span - 1;       // padding repeats the last element
        const int elem = b.felem + src;
        const size_t slot = (size_t)b.slot0 + i;
        for (int j = 0; j < 24; ++j) {
          const int32_t e = c.edest[b.tab_off + (size_t)src * 24 + j];
          const int32_t cc = c.cdest[b.tab_off + (size_t)src * 24 + j];
          if (e < 1 || e > c.nodof || cc < 1 || cc > c.nodof)
            return fail(W3D_ERR_CONFIG, "block %d element %d: edest/cdest entry %d outside 1..nodof", blk + 1, elem, j + 1);
          ed[slot * 24 + j] = e;
          cdv[slot * 24 + j] = cc;
        }
        // setup_mm01_rknstr (rknstr.f:1018-1066): real*4 props widened on the device
        float *p = &ep[slot * EPROP];
        p[0] = rprop(c, 7, elem); p[1] = rprop(c, 8, elem); p[2] = rprop(c, 14, elem);
        p[3] = rprop(c, 15, elem); p[4] = rprop(c, 23, elem); p[5] = lprop(c, 17, elem) ? 1.0f : 0.0f;
        if (b.mat_type == 5)   // setup_mm05_rknstr (rknstr.f:1234-1317): mm05_props = matprp(55:64)
          for (int k = 0; k < 10; ++k) p[6 + k] = c.matprp[(size_t)(b.matnum - 1) * c.mxmtpr + 54 + k];
        if (b.mat_type == 3) {   // setup_mm03_rknstr (rknstr.f:1142-1215)
          p[6] = rprop(c, 21, elem); p[7] = rprop(c, 20, elem); p[8] = rprop(c, 22, elem);
          p[9] = rprop(c, 27, elem); p[10] = rprop(c, 28, elem); p[11] = rprop(c, 29, elem);
        }
      }
      for (int w = 0; w <= W3D_ELEM_VOL; ++w) {
        const long long n = block_doubles(w, b);
        if (n > 0 && (size_t)n > max_blk_doubles) max_blk_doubles = (size_t)n;
      }
    }
    if (sg.mat_type == 10) {
      // setup_mm10_rknstr (rknstr.f:1568-2027) once per (element, crystal): rotation from the
      // Kocks angles, Schmid tensors and elastic stiffness in the reference frame
      sg.ncrystals = c.cp_ncrystals;
      sg.nslip = (int)c.crystal_types[337];
      sg.cp_fast = mm10_fast_eligible(c.crystal_types, c.n_crystal_types);
      if (sg.cp_fast) {
        sg.cp_taul = c.crystal_types[36 + 288 + 5] != 0.0;      // k_o: geometric hardening
        sg.cp_consts.resize(mm10_fast_consts_bytes());
        mm10_fast_setup_consts_host(c.crystal_types, c.cp_angle_type, sg.cp_consts.data());
        const char *sch = getenv("W3D_MM10_SCHEDULE");
        if (!(sch && sch[0] == '0')) CUDA_TRY(mm10_schedule_create(cta.data(), sg.ncta, &sg.cp_sched));
      }
      const int pd = sg.cp_fast ? mm10_fast_props_doubles() : mm10_props_doubles(sg.nslip);
      sg.num_hard = (int)c.crystal_types[338];
      mm10_layout_offsets(sg.nslip, sg.num_hard, &sg.o_gradfe, &sg.o_Rp0);
      std::vector<double> cp((size_t)sg.nslots * sg.ncrystals * pd, 0.0);
      for (int blk : sg.blocks) {
        const BlockInfo &b = h->blocks[blk];
        const int padded = ((b.span + EPC - 1) / EPC) * EPC;
        for (int i = 0; i < padded; ++i) {
          const int src = (i < b.span) ? i : b.span - 1;
          const int elem = b.felem + src;
          for (int cr = 0; cr < sg.ncrystals; ++cr) {
            const size_t k = (size_t)(elem - 1) * sg.ncrystals + cr;
            const int cnum = c.elem_cnum[k];
            if (cnum < 1 || cnum > c.n_crystal_types)
              return fail(W3D_ERR_CONFIG, "element %d crystal %d: crystal number %d outside 1..%d", elem, cr + 1, cnum, c.n_crystal_types);
            double *rec = &cp[((size_t)(b.slot0 + i) * sg.ncrystals + cr) * pd];
            const double *ctr = c.crystal_types + (size_t)(cnum - 1) * W3D_CRYSTAL_TYPE_DOUBLES;
            if (sg.cp_fast) {
              if (!mm10_fast_setup_props_host(ctr, c.elem_angles + 3 * k, c.cp_angle_type, rec))
                return fail(W3D_ERR_CONFIG, "element %d crystal %d: singular elastic stiffness", elem, cr + 1);
            } else {
              mm10_setup_props_host(ctr, c.elem_angles + 3 * k, c.cp_angle_type, rec);
            }
          }
        }
      }
      CUDA_TRY(dalloc(&sg.d_cprops, cp.size()));
      CUDA_TRY(cudaMemcpy(sg.d_cprops, cp.data(), sizeof(double) * cp.size(), cudaMemcpyHostToDevice));
      CUDA_TRY(dalloc(&sg.scratch, (size_t)16 * sg.nslots * 8));
    }
    CUDA_TRY(dalloc(&sg.d_cta, (size_t)sg.ncta));
    CUDA_TRY(cudaMemcpy(sg.d_cta, cta.data(), sizeof(CtaDesc) * sg.ncta, cudaMemcpyHostToDevice));
    CUDA_TRY(dalloc(&sg.d_edest, ed.size()));
    CUDA_TRY(cudaMemcpy(sg.d_edest, ed.data(), sizeof(int32_t) * ed.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(dalloc(&sg.d_cdest, cdv.size()));
    CUDA_TRY(cudaMemcpy(sg.d_cdest, cdv.data(), sizeof(int32_t) * cdv.size(), cudaMemcpyHostToDevice));
    if (sg.mat_type == 3) {
      CUDA_TRY(cudaMalloc((void **)&sg.d_mm03_active, sizeof(int) * 8 * (size_t)(sg.nslots / EPC)));
      CUDA_TRY(dalloc(&sg.scratch, (size_t)16 * sg.nslots * 8));
      // allow_cut (rknstr.f:1213-1214): bit 2 of iprops(24) of the material; one flag per segment
      sg.mm03_allow_cut = 1;
      for (int blk : sg.blocks)
        if (!(iprop(c, 24, h->blocks[blk].felem) & 2)) sg.mm03_allow_cut = 0;
    }
    CUDA_TRY(dalloc(&sg.d_eprops, ep.size()));
    CUDA_TRY(cudaMemcpy(sg.d_eprops, ep.data(), sizeof(float) * ep.size(), cudaMemcpyHostToDevice));
    const size_t ng = (size_t)sg.ng;
    CUDA_TRY(dalloc(&sg.eps_n, 6 * ng));   CUDA_TRY(dalloc(&sg.eps_n1, 6 * ng));
    CUDA_TRY(dalloc(&sg.urcs_n, 9 * ng));  CUDA_TRY(dalloc(&sg.urcs_n1, 9 * ng));
    CUDA_TRY(dalloc(&sg.rot_n, 9 * ng));   CUDA_TRY(dalloc(&sg.rot_n1, 9 * ng));
    CUDA_TRY(dalloc(&sg.cep, 21 * ng));
    CUDA_TRY(dalloc(&sg.hist_n, (size_t)sg.hist_size * ng));
    CUDA_TRY(dalloc(&sg.hist_n1, (size_t)sg.hist_size * ng));
    CUDA_TRY(dalloc(&sg.elem_vol, (size_t)sg.nslots));
    CUDA_TRY(dalloc(&sg.partials, (size_t)(sg.ncta + REDUCE_STAGE1_CTAS) * 4));   // + stage-1 sums
    k_rot_identity<<<148 * 4, 256, 0, h->stream>>>(sg.rot_n, sg.ng);
    k_rot_identity<<<148 * 4, 256, 0, h->stream>>>(sg.rot_n1, sg.ng);
    h->launches += 2;
  }
  h->stage_doubles = max_blk_doubles;
  CUDA_TRY(dalloc(&h->d_stage, h->stage_doubles));
  CUDA_TRY(cudaFuncSetAttribute(k_tanstf<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EPC * ELREC * 8));
  CUDA_TRY(cudaFuncSetAttribute(k_tanstf<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EPC * ELREC * 8));
  CUDA_TRY(cudaFuncSetAttribute(k_tanstf<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EPC * ELREC * 8));
  CUDA_TRY(cudaFuncSetAttribute(k_tanstf<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EPC * ELREC * 8));
  // opt-in of every kernel that needs more than 48 KB of dynamic shared memory, on THIS device
  CUDA_TRY((stress_prepare_device<1, 0>()));
  CUDA_TRY((stress_prepare_device<10, 1>()));
  CUDA_TRY((stress_prepare_device<10, 2>()));
  CUDA_TRY(mm05_prepare_device());
  CUDA_TRY(mm03_prepare_device());
  CUDA_TRY(mm10_prepare_device());
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaGetLastError());
  return W3D_OK;
}

// the exchange step after the stress kernels of all segments (stream-ordered, no host sync)
// deterministic fixed-order reduction of the per-CTA partial sums of every segment into d_scal
void reduce_scalars(w3d_engine *h, cudaStream_t st) {
  int first = 1;
  for (auto &sg : h->segs) {
    if (sg.ncta > 4096) {   // two fixed-order stages (still deterministic)
      double *tmp = sg.partials + (size_t)sg.ncta * 4;
      k_reduce_partials_stage1<<<REDUCE_STAGE1_CTAS, 256, 0, st>>>(sg.partials, sg.ncta, tmp);
      k_reduce_partials<<<1, 256, 0, st>>>(tmp, REDUCE_STAGE1_CTAS, h->d_scal, first ? 0 : 1);
      h->launches += 2;
    } else {
      k_reduce_partials<<<1, 256, 0, st>>>(sg.partials, sg.ncta, h->d_scal, first ? 0 : 1);
      h->launches += 1;
    }
    first = 0;
  }
}

// make the engine's stream wait for an exchange still in flight on the communication stream
int comm_join(w3d_engine *h) {
  if (!h->comm.pending) return W3D_OK;
  h->comm.pending = false;
  if (cudaStreamWaitEvent(h->stream, h->comm.ev_done, 0) != cudaSuccess)
    return h->fail(W3D_ERR_CUDA, "cudaStreamWaitEvent on the exchange failed");
  return W3D_OK;
}

int comm_exchange(w3d_engine *h) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CommData &cm = h->comm;
  NcclApi *api = nccl_api();
  const int nodof = h->cfg.nodof;
  // everything below goes to the communication stream, after the stress kernels of this pass
  CUDA_TRY(cudaEventRecord(cm.ev_start, h->stream));
  CUDA_TRY(cudaStreamWaitEvent(cm.stream, cm.ev_start, 0));
  cudaStream_t cs = cm.stream;
  reduce_scalars(h, cs);
  if (cm.exchange == 1 && cm.ntot >= 0 && !cm.counts.empty()) {
    if (cm.ntot > 0) {
      const int grid = (int)std::min<long long>((cm.ntot + 255) / 256, 148 * 8);
      k_iface_pack<<<grid, 256, 0, cs>>>(h->d_ifv, cm.d_idx, cm.d_send, cm.ntot);
      NCCL_TRY(api->group_start());
      long long off = 0;
      for (int r = 0; r < cm.nranks; ++r) {
        const long long n = cm.counts[r];
        if (n > 0) {
          NCCL_TRY(api->send(cm.d_send + off, (size_t)n, NCCL_FLOAT64, r, cm.comm, cs));
          NCCL_TRY(api->recv(cm.d_recv + off, (size_t)n, NCCL_FLOAT64, r, cm.comm, cs));
        }
        off += n;
      }
      NCCL_TRY(api->group_end());
      k_iface_add<<<grid, 256, 0, cs>>>(h->d_ifv, cm.d_idx, cm.d_recv, cm.ntot);
      h->launches += 2;
    }
  } else {
    NCCL_TRY(api->all_reduce(h->d_ifv, h->d_ifv, (size_t)nodof, NCCL_FLOAT64, NCCL_SUM, cm.comm, cs));
  }
  // internal_energy, plastic_work, sum_ifv, num_term_ifv and the cut flag: one all-reduce
  k_scal_pack<<<1, 1, 0, cs>>>(h->d_scal, h->d_flags, 24.0 * (double)h->n_owned_elems, cm.d_red);
  h->launches += 1;
  NCCL_TRY(api->all_reduce(cm.d_red, cm.d_red, 5, NCCL_FLOAT64, NCCL_SUM, cm.comm, cs));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(cm.ev_done, cs));
  cm.pending = true;
  return W3D_OK;
}

// host -> device copy of u, du; with bcast_inputs rank 0's copy goes to every rank over NVLink
int upload_inputs(w3d_engine *h, const double *u, const double *du) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  const size_t nb = sizeof(double) * h->cfg.nodof;
  CommData &cm = h->comm;
  const bool bc = cm.comm && cm.bcast_inputs;
  if (!bc || cm.rank == 0) {
    if (!u || !du) return fail(W3D_ERR_ARG, "u / du are NULL");
    CUDA_TRY(cudaMemcpyAsync(h->d_u, u, nb, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_du, du, nb, cudaMemcpyHostToDevice, h->stream));
  }
  if (bc) {
    const int rj = comm_join(h);     // one NCCL operation of this communicator at a time
    if (rj) return rj;
  }
  if (bc)
    NCCL_TRY(nccl_api()->broadcast(h->d_u, h->d_u, 2 * (size_t)h->cfg.nodof, NCCL_FLOAT64, 0, cm.comm, h->stream));
  h->have_vectors = true;
  return W3D_OK;
}

int stress_resident(w3d_engine *h, int step, int iter, int extrapolated_du) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  const w3d_engine_config &c = h->cfg;
  { const int rj = comm_join(h); if (rj) return rj; }
  if (h->comm.comm && h->comm.exchange == 0)      // the all-reduce leaves sums in every entry
    CUDA_TRY(cudaMemsetAsync(h->d_ifv, 0, sizeof(double) * c.nodof, h->stream));
  else if (h->dof_hi > h->dof_lo)
    CUDA_TRY(cudaMemsetAsync(h->d_ifv + h->dof_lo, 0, sizeof(double) * (size_t)(h->dof_hi - h->dof_lo), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, 64 * sizeof(int), h->stream));
  for (auto &sg : h->segs) {
    StressArgs a{};
    a.cta = sg.d_cta; a.edest = sg.d_edest; a.cdest = sg.d_cdest; a.eprops = sg.d_eprops;
    a.c = h->d_c; a.u = h->d_u; a.du = h->d_du; a.ifv = h->d_ifv;
    a.eps_n = sg.eps_n; a.urcs_n = sg.urcs_n; a.hist_n = sg.hist_n;
    a.eps_n1 = sg.eps_n1; a.urcs_n1 = sg.urcs_n1; a.rot_n1 = sg.rot_n1; a.hist_n1 = sg.hist_n1;
    a.cep = sg.cep; a.elem_vol = sg.elem_vol; a.partials = sg.partials; a.flags = h->d_flags;
    a.ng = sg.ng; a.eps_bbar = c.eps_bbar; a.beta_fact = c.beta_fact;
    a.step = step; a.iter = iter;
    a.nonlinear_update = (iter >= 1 || extrapolated_du) ? 1 : 0;
    a.update_coords = (sg.geonl && !(step == 1 && iter == 0)) ? 1 : 0;
    a.adaptive = (c.adaptive_flag && step > 1) ? 1 : 0;
    a.hist_size = sg.hist_size;
    if (sg.mat_type == 10) {
      a.scratch = sg.scratch; a.cp_ncrystals = sg.ncrystals; a.cp_o_gradfe = sg.o_gradfe; a.cp_o_Rp = sg.o_Rp0;
      CUDA_TRY((launch_stress<10, 1>(sg, a, h->stream)));
      CpArgs ca{};
      ca.cta = sg.d_cta; ca.props = sg.d_cprops; ca.hist_n = sg.hist_n; ca.urcs_n = sg.urcs_n;
      ca.scratch = sg.scratch; ca.hist_n1 = sg.hist_n1; ca.urcs_n1 = sg.urcs_n1; ca.flags = h->d_flags;
      ca.ng = sg.ng; ca.dt = h->dt; ca.step = step;
      ca.iter_0_extrapolate_off = (iter == 0 && (!extrapolated_du || step == 1)) ? 1 : 0;   // rstgp1.f:4453-4457
      ca.asymmetric = c.asymmetric_assembly ? 1 : 0;
      ca.ncrystals = sg.ncrystals; ca.nslip = sg.nslip; ca.num_hard = sg.num_hard;
      if (sg.cp_fast) CUDA_TRY(launch_mm10_fast(ca, sg.ncta, sg.cp_consts.data(), sg.cp_taul, sg.cp_sched, h->stream));
      else CUDA_TRY(launch_mm10(ca, sg.ncta, h->stream));
      CUDA_TRY((launch_stress<10, 2>(sg, a, h->stream)));
      h->launches += 3;
    } else if (sg.mat_type == 3) {
      // drive_03_update: kinematics -> k_mm03 (one CTA per element block) -> internal force
      a.scratch = sg.scratch;
      CUDA_TRY(launch_stress_mm03(sg.geonl, sg.bbar, 1, sg.ncta, a, h->stream));
      Mm03Args ma{};
      ma.cta = sg.d_cta; ma.blk_active = sg.d_mm03_active; ma.eprops = sg.d_eprops; ma.hist_n = sg.hist_n; ma.urcs_n = sg.urcs_n;
      ma.scratch = sg.scratch; ma.hist_n1 = sg.hist_n1; ma.urcs_n1 = sg.urcs_n1; ma.cep = sg.cep;
      ma.flags = h->d_flags; ma.ng = sg.ng; ma.dt = h->dt; ma.step = step; ma.iter = iter;
      ma.allow_cut = (a.adaptive && sg.mm03_allow_cut) ? 1 : 0;
      ma.nonlinear_update = a.nonlinear_update;
      if (ma.nonlinear_update)
        CUDA_TRY(cudaMemsetAsync(sg.d_mm03_active, 0, sizeof(int) * 8 * (size_t)(sg.nslots / EPC), h->stream));
      CUDA_TRY(launch_mm03(ma, sg.ncta, h->stream));
      CUDA_TRY(launch_stress_mm03(sg.geonl, sg.bbar, 2, sg.ncta, a, h->stream));
      h->launches += ma.nonlinear_update ? 4 : 3;
    } else if (sg.mat_type == 5) {
      CUDA_TRY(launch_stress_mm05(sg.geonl, sg.bbar, sg.ncta, a, h->stream));
      h->launches += 1;
    } else {
      CUDA_TRY((launch_stress<1, 0>(sg, a, h->stream)));
      h->launches += 1;
    }
  }
  // energy / plastic work / sum |f| of the pass: on the engine's stream, or -- with a communicator
  // -- on the communication stream in front of the scalar all-reduce (off the critical path)
  if (!h->comm.comm) reduce_scalars(h, h->stream);
  CUDA_TRY(cudaGetLastError());
  h->have_vectors = true;
  if (h->comm.comm) {
    const int rc = comm_exchange(h);
    if (rc) return rc;
  }
  return W3D_OK;
}

int fetch_scalars(w3d_engine *h, w3d_scalars *out) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  { const int rj = comm_join(h); if (rj) return rj; }
  CUDA_TRY(cudaMemcpyAsync(h->h_pin, h->d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->h_pin + 4, h->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (h->comm.comm)
    CUDA_TRY(cudaMemcpyAsync(h->h_pin + 8, h->comm.d_red, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  int flags[2];
  std::memcpy(flags, h->h_pin + 4, sizeof flags);
#ifdef W3D_MM10F_HIST
  {                                   // debug builds only: solver-loop trips per Gauss point (w3d_cp.cu)
    int hist[64];
    cudaMemcpy(hist, h->d_flags, sizeof hist, cudaMemcpyDeviceToHost);
    fprintf(stderr, "mm10f trips histogram:");
    for (int k = 2; k < 64; ++k) fprintf(stderr, " %d", hist[k]);
    fprintf(stderr, "\n");
  }
#endif
  w3d_scalars s{};
  if (h->comm.comm) {     // sums over all ranks (wmpi_reduce_vec / wmpi_redlog)
    const double *r = h->h_pin + 8;
    s.material_cut_step = r[4] > 0.0 ? 1 : 0;
    if (!s.material_cut_step) {
      s.internal_energy = r[0]; s.plastic_work = r[1]; s.sum_ifv = r[2];
      s.num_term_ifv = (int64_t)llround(r[3]);
    }
  } else {
    s.material_cut_step = flags[0] ? 1 : 0;
    if (!s.material_cut_step) {
      s.internal_energy = h->h_pin[0];
      s.plastic_work = h->h_pin[1];
      s.sum_ifv = h->h_pin[2];
      s.num_term_ifv = 24 * (int64_t)h->n_owned_elems;
    }
  }
  h->last = s;
  if (out) *out = s;
  if (flags[1] == W3D_ERR_MM01_LAMBDA) return fail(W3D_ERR_MM01_LAMBDA, "mm01_general: lambda_deltat is negative (mm01.f:878)");
  if (flags[1] == W3D_ERR_DETF) return fail(W3D_ERR_DETF, "det F <= 0 and the step is not adaptive (rstgp1.f:136-145)");
  if (flags[1]) return fail(flags[1], "device-side fatal condition %d", flags[1]);
  return W3D_OK;
}

int tanstf_resident(w3d_engine *h, int step, int iter) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  (void)step; (void)iter;
  const w3d_engine_config &c = h->cfg;
  for (auto &sg : h->segs) {
    TanstfArgs a{};
    a.cta = sg.d_cta; a.edest = sg.d_edest; a.cdest = sg.d_cdest;
    a.c = h->d_c; a.u = h->d_u; a.du = h->d_du;
    a.urcs_n1 = sg.urcs_n1; a.rot_n1 = sg.rot_n1; a.cep = sg.cep;
    a.estiff = h->d_estiff; a.ng = sg.ng; a.eps_bbar = c.eps_bbar; a.beta_fact = c.beta_fact;
    a.qbar = c.qbar_flag ? 1 : 0;
    a.hist_n1 = sg.hist_n1; a.d_from_hist = (sg.mat_type == 10) ? 1 : 0;
    const size_t shm = (size_t)EPC * ELREC * 8;
    if (sg.geonl) {
      if (sg.bbar) k_tanstf<true, true><<<sg.ncta, TPB, shm, h->stream>>>(a);
      else k_tanstf<true, false><<<sg.ncta, TPB, shm, h->stream>>>(a);
    } else {
      if (sg.bbar) k_tanstf<false, true><<<sg.ncta, TPB, shm, h->stream>>>(a);
      else k_tanstf<false, false><<<sg.ncta, TPB, shm, h->stream>>>(a);
    }
    h->launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  return W3D_OK;
}

int copy_state(w3d_engine *h, int from, int to) {
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  for (auto &sg : h->segs) {
    const double *src = soa_of(sg, from);
    double *dst = soa_of(sg, to);
    const size_t n = (size_t)nprm_of(h, from, h->blocks[sg.blocks[0]]) * (size_t)sg.ng;
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  return W3D_OK;
}

}  // namespace

extern "C" {

int32_t w3d_engine_create(const w3d_engine_config *cfg, w3d_engine **handle) {
  if (!cfg || !handle) { g_create_error = "null argument"; return W3D_ERR_ARG; }
  *handle = nullptr;
  w3d_engine *h = new w3d_engine();
  h->cfg = *cfg;
  const int rc = create_impl(h);
  if (rc != W3D_OK) {
    g_create_error = h->err;
    free_all(h);
    delete h;
    return rc;
  }
  // the engine keeps no pointer into caller memory after create
  h->cfg.elblks = nullptr; h->cfg.props = nullptr; h->cfg.matprp = nullptr; h->cfg.c = nullptr;
  h->cfg.edest = nullptr; h->cfg.cdest = nullptr; h->cfg.history_blk_list = nullptr;
  h->cfg.crystal_types = nullptr; h->cfg.elem_cnum = nullptr; h->cfg.elem_angles = nullptr;
  *handle = h;
  return W3D_OK;
}

int32_t w3d_engine_destroy(w3d_engine *h) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  if (h->comm.stream) cudaStreamSynchronize(h->comm.stream);
  cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
  return W3D_OK;
}

const char *w3d_engine_last_error(const w3d_engine *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t w3d_eps_sig_ifv(w3d_engine *h, int32_t step, int32_t iter, int32_t extrapolated_du,
                        const double *u, const double *du, double *ifv, w3d_scalars *out) {
  if (!h) return W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  // with a communicator ifv may be NULL (a rank that does not want the vector back; the owned
  // part is fetched with w3d_download_ifv_range) and, with bcast_inputs, u / du on ranks > 0
  if (!h->comm.comm && (!u || !du || !ifv)) return fail(W3D_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t nb = sizeof(double) * h->cfg.nodof;
  int rc = upload_inputs(h, u, du);
  if (rc) return rc;
  rc = stress_resident(h, step, iter, extrapolated_du);
  if (rc) return rc;
  { const int rj = comm_join(h); if (rj) return rj; }
  if (ifv) CUDA_TRY(cudaMemcpyAsync(ifv, h->d_ifv, nb, cudaMemcpyDeviceToHost, h->stream));
  w3d_scalars s;
  rc = fetch_scalars(h, &s);
  if (s.material_cut_step && ifv) std::memset(ifv, 0, nb);   // :253-263
  if (out) *out = s;
  return rc;
}

int32_t w3d_tanstf(w3d_engine *h, int32_t step, int32_t iter, const double *u, const double *du,
                   double *estiff) {
  if (!h) return W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t nb = sizeof(double) * h->cfg.nodof;
  (void)nb;
  if (u && du) {
    const int rcu = upload_inputs(h, u, du);
    if (rcu) return rcu;
  } else if (!h->have_vectors) {
    return fail(W3D_ERR_ARG, "w3d_tanstf: u/du are NULL and no displacement vectors are resident");
  }
  const int rc = tanstf_resident(h, step, iter);
  if (rc) return rc;
  if (estiff)
    CUDA_TRY(cudaMemcpyAsync(estiff, h->d_estiff, sizeof(double) * 300 * (size_t)h->n_owned_elems, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}

int32_t w3d_set_time_step(w3d_engine *h, double dt) {
  if (!h) return W3D_ERR_ARG;
  h->dt = dt;
  return W3D_OK;
}

int32_t w3d_update(w3d_engine *h) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  int rc = copy_state(h, W3D_HIST_N1, W3D_HIST_N);
  if (!rc) rc = copy_state(h, W3D_EPS_N1, W3D_EPS_N);
  if (!rc) rc = copy_state(h, W3D_URCS_N1, W3D_URCS_N);
  return rc;
}
int32_t w3d_adaptive_save(w3d_engine *h) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  return copy_state(h, W3D_ROT_N1, W3D_ROT_N);
}
int32_t w3d_adaptive_reset(w3d_engine *h) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  int rc = copy_state(h, W3D_HIST_N, W3D_HIST_N1);
  if (!rc) rc = copy_state(h, W3D_ROT_N, W3D_ROT_N1);
  if (!rc) rc = copy_state(h, W3D_EPS_N, W3D_EPS_N1);
  if (!rc) rc = copy_state(h, W3D_URCS_N, W3D_URCS_N1);
  return rc;
}

static int state_xfer(w3d_engine *h, int which, int blk1, double *host, bool push) {
  if (!h || !host) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  if (blk1 < 1 || blk1 > (int)h->blocks.size()) return fail(W3D_ERR_ARG, "block %d out of range", blk1);
  const BlockInfo &b = h->blocks[blk1 - 1];
  if (!b.owned) return fail(W3D_ERR_ARG, "block %d is not owned by this engine (domain %d)", blk1, b.domain);
  const long long n = block_doubles(which, b);
  if (n < 0) return fail(W3D_ERR_ARG, "unknown state selector %d", which);
  if (n == 0) return W3D_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  Segment &sg = h->segs[b.seg];
  const auto kind = push ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (which == W3D_ESTIFF || which == W3D_ELEM_VOL) {
    double *dev = (which == W3D_ESTIFF) ? h->d_estiff + (size_t)b.epos0 * 300 : sg.elem_vol + b.slot0;
    if (push) CUDA_TRY(cudaMemcpyAsync(dev, host, n * sizeof(double), kind, h->stream));
    else CUDA_TRY(cudaMemcpyAsync(host, dev, n * sizeof(double), kind, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return W3D_OK;
  }
  double *soa = soa_of(sg, which);
  const int nprm = nprm_of(h, which, b);
  const int grid = (int)((n + 255) / 256);
  if (push) {
    CUDA_TRY(cudaMemcpyAsync(h->d_stage, host, n * sizeof(double), kind, h->stream));
    k_block_to_soa<<<grid, 256, 0, h->stream>>>(h->d_stage, soa, sg.ng, b.slot0, b.span, nprm, which == W3D_CEP);
  } else {
    k_soa_to_block<<<grid, 256, 0, h->stream>>>(h->d_stage, soa, sg.ng, b.slot0, b.span, nprm, which == W3D_CEP);
    CUDA_TRY(cudaMemcpyAsync(host, h->d_stage, n * sizeof(double), kind, h->stream));
  }
  h->launches += 1;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaGetLastError());
  return W3D_OK;
}

int32_t w3d_push_state(w3d_engine *h, int32_t which, int32_t blk, const double *src) {
  return state_xfer(h, which, blk, const_cast<double *>(src), true);
}
int32_t w3d_pull_state(w3d_engine *h, int32_t which, int32_t blk, double *dst) {
  return state_xfer(h, which, blk, dst, false);
}

int64_t w3d_block_offset(const w3d_engine *h, int32_t which, int32_t blk) {
  if (!h || blk < 1 || blk > (int)h->blocks.size() + 1) return -1;
  int64_t off = 0;
  for (int k = 0; k < blk - 1; ++k) {
    if (!h->blocks[k].owned) continue;
    const long long n = block_doubles(which, h->blocks[k]);
    if (n < 0) return -1;
    off += n;
  }
  return off;
}

int32_t w3d_device_vectors(w3d_engine *h, double **d_u, double **d_du, double **d_ifv) {
  if (!h) return W3D_ERR_ARG;
  if (d_u) *d_u = h->d_u;
  if (d_du) *d_du = h->d_du;
  if (d_ifv) *d_ifv = h->d_ifv;
  comm_join(h);
  h->have_vectors = true;
  return W3D_OK;
}
int32_t w3d_upload_displacements(w3d_engine *h, const double *u, const double *du) {
  if (!h) return W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  const int rc = upload_inputs(h, u, du);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}
int32_t w3d_download_ifv_range(w3d_engine *h, int32_t first, int32_t n, double *dst) {
  if (!h || !dst) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  if (first < 1 || n < 0 || (long long)first + n - 1 > h->cfg.nodof) return fail(W3D_ERR_ARG, "dof range %d..%d outside 1..%d", first, first + n - 1, h->cfg.nodof);
  CUDA_TRY(cudaSetDevice(h->device));
  { const int rj = comm_join(h); if (rj) return rj; }
  CUDA_TRY(cudaMemcpyAsync(dst, h->d_ifv + (first - 1), sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}

int32_t w3d_comm_unique_id(uint8_t id[128]) {
  NcclApi *api = nccl_api();
  if (!id || !api->lib || !api->error.empty()) { g_create_error = api->error.empty() ? "null argument" : api->error; return W3D_ERR_COMM; }
  NcclUniqueId u;
  const int r = api->get_unique_id(&u);
  if (r != 0) { g_create_error = std::string("ncclGetUniqueId: ") + api->get_error_string(r); return W3D_ERR_COMM; }
  std::memcpy(id, u.internal, 128);
  return W3D_OK;
}
int32_t w3d_comm_init(w3d_engine *h, const uint8_t id[128], int32_t rank, int32_t nranks) {
  if (!h || !id) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  NcclApi *api = nccl_api();
  if (!api->lib || !api->error.empty()) return fail(W3D_ERR_COMM, "NCCL not usable: %s", api->error.c_str());
  if (h->comm.comm) return fail(W3D_ERR_COMM, "communicator already initialised");
  if (nranks != h->cfg.numprocs || rank != h->cfg.myid)
    return fail(W3D_ERR_CONFIG, "rank %d of %d does not match the engine's myid %d / numprocs %d", rank, nranks, h->cfg.myid, h->cfg.numprocs);
  CUDA_TRY(cudaSetDevice(h->device));
  NcclUniqueId u;
  std::memcpy(u.internal, id, 128);
  void *comm = nullptr;
  NCCL_TRY(api->comm_init_rank(&comm, nranks, u, rank));
  CommData &cm = h->comm;
  cm.comm = comm; cm.rank = rank; cm.nranks = nranks;
  cm.ntot = 0;
  for (long long n : cm.counts) cm.ntot += n;
  CUDA_TRY(dalloc(&cm.d_red, 8));
  CUDA_TRY(cudaStreamCreateWithFlags(&cm.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&cm.ev_start, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&cm.ev_done, cudaEventDisableTiming));
  if (cm.ntot > 0) {
    CUDA_TRY(dalloc(&cm.d_idx, (size_t)cm.ntot));
    CUDA_TRY(dalloc(&cm.d_send, (size_t)cm.ntot));
    CUDA_TRY(dalloc(&cm.d_recv, (size_t)cm.ntot));
    CUDA_TRY(cudaMemcpy(cm.d_idx, cm.dofs.data(), sizeof(int32_t) * cm.ntot, cudaMemcpyHostToDevice));
  }
  if (!cm.owned.empty()) {
    CUDA_TRY(cudaMalloc((void **)&cm.d_owned, cm.owned.size()));
    CUDA_TRY(cudaMemcpy(cm.d_owned, cm.owned.data(), cm.owned.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(dalloc(&cm.d_rred, 4 + 3 * (size_t)nranks));
  }
  int nnb = 0;
  for (long long n : cm.counts) nnb += n > 0;
  fprintf(stderr, "w3d_comm: rank %d of %d on device %d, %d neighbour domain(s), %lld shared dofs of %d\n", rank, nranks,
          h->device, nnb, cm.ntot, h->cfg.nodof);
  return W3D_OK;
}
int32_t w3d_comm_set_options(w3d_engine *h, int32_t exchange, int32_t bcast_inputs) {
  if (!h) return W3D_ERR_ARG;
  if (exchange != 0 && exchange != 1) return h->fail(W3D_ERR_ARG, "exchange must be 0 (all-reduce) or 1 (neighbours)");
  if (exchange == 1 && h->comm.counts.empty() && h->cfg.numprocs > 1) return h->fail(W3D_ERR_UNSUPPORTED, "no neighbour plan for this model (more than 64 domains)");
  h->comm.exchange = exchange;
  h->comm.bcast_inputs = bcast_inputs ? 1 : 0;
  cudaSetDevice(h->device);
  comm_join(h);
  cudaMemsetAsync(h->d_ifv, 0, sizeof(double) * h->cfg.nodof, h->stream);   // no stale sums of the other mode
  return W3D_OK;
}
int32_t w3d_comm_plan(const w3d_engine_config *cfg, int32_t rank, int64_t *counts, int32_t *dofs) {
  if (!cfg || !counts || rank < 0 || rank >= (cfg->numprocs > 0 ? cfg->numprocs : 1)) return W3D_ERR_ARG;
  std::vector<long long> cnt;
  std::vector<int32_t> d;
  std::string err;
  const int rc = comm_plan(*cfg, rank, cnt, d, err);
  if (rc) { g_create_error = err; return rc; }
  for (size_t r = 0; r < cnt.size(); ++r) counts[r] = cnt[r];
  if (dofs) std::memcpy(dofs, d.data(), sizeof(int32_t) * d.size());
  return W3D_OK;
}
int32_t w3d_comm_owned(const w3d_engine_config *cfg, int32_t rank, uint8_t *owned) {
  if (!cfg || !owned || !cfg->elblks || !cfg->edest || !cfg->props) return W3D_ERR_ARG;
  const w3d_engine_config &c = *cfg;
  const int np = c.numprocs > 0 ? c.numprocs : 1;
  std::vector<int32_t> low((size_t)c.nodof, np);
  long long tab = 0;
  for (int blk = 0; blk < c.nelblk; ++blk) {
    const int span = c.elblks[4 * blk + 0], felem = c.elblks[4 * blk + 1];
    const int dom = np > 1 ? c.elblks[4 * blk + 2] : 0;
    if (span <= 0 || felem < 1 || felem + span - 1 > c.noelem || dom < 0 || dom >= np) return W3D_ERR_CONFIG;
    const int totdof = iprop(c, 2, felem) * iprop(c, 4, felem);
    for (long long k = 0; k < (long long)totdof * span; ++k) {
      const int32_t e = c.edest[tab + k];
      if (e < 1 || e > c.nodof) return W3D_ERR_CONFIG;
      if (dom < low[(size_t)e - 1]) low[(size_t)e - 1] = dom;
    }
    tab += (long long)totdof * span;
  }
  // a dof no block touches (an unconnected node) belongs to rank 0
  for (int d = 0; d < c.nodof; ++d) owned[d] = (low[d] == rank || (low[d] == np && rank == 0)) ? 1 : 0;
  return W3D_OK;
}
int32_t w3d_download_ifv(w3d_engine *h, double *ifv) {
  if (!h || !ifv) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  { const int rj = comm_join(h); if (rj) return rj; }
  CUDA_TRY(cudaMemcpyAsync(ifv, h->d_ifv, sizeof(double) * h->cfg.nodof, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}
int32_t w3d_download_estiff(w3d_engine *h, double *estiff) {
  if (!h || !estiff) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(estiff, h->d_estiff, sizeof(double) * 300 * (size_t)h->n_owned_elems, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}
int32_t w3d_eps_sig_ifv_resident(w3d_engine *h, int32_t step, int32_t iter, int32_t extrapolated_du) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  return stress_resident(h, step, iter, extrapolated_du);
}
int32_t w3d_tanstf_resident(w3d_engine *h, int32_t step, int32_t iter) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  return tanstf_resident(h, step, iter);
}
int32_t w3d_last_scalars(w3d_engine *h, w3d_scalars *out) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  return fetch_scalars(h, out);
}
void *w3d_engine_stream(w3d_engine *h) { return h ? (void *)h->stream : nullptr; }
int32_t w3d_engine_synchronize(w3d_engine *h) {
  if (!h) return W3D_ERR_ARG;
  cudaSetDevice(h->device);
  if (comm_join(h)) return W3D_ERR_CUDA;
  return cudaStreamSynchronize(h->stream) == cudaSuccess ? W3D_OK : h->fail(W3D_ERR_CUDA, "stream synchronize failed");
}
int64_t w3d_engine_launch_count(const w3d_engine *h) { return h ? h->launches : -1; }

}  // extern "C"

#include "w3d_assemble.cuh"   // element stiffnesses -> assembled sparse equations (NEXT-1)
#include "w3d_residual.cuh"   // structure residual and convergence norms (NEXT-2)
