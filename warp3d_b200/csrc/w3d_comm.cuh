// w3d_comm.cuh -- the exchange step of the path behind the C ABI (included by w3d_engine.cu).
//
// Replaces, for one process per GPU, what the reference's MPI build does at the end of
// drive_eps_sig_internal_forces: wmpi_reduce_vec(ifv) (:338-340), the three scalar reductions and
// the logical OR of material_cut_step (:248-250).  NCCL is bound at run time (dlopen of
// libnccl.so.2 -- the copy already loaded in the process if there is one), so the engine library
// itself has no link-time dependency and loads on a box without NCCL.
//
// Exchange of the internal forces (w3d_comm_set_options, mode):
//   1 = neighbour exchange (default).  At create every engine works out, from the replicated block
//       tables, which dofs its own blocks share with each other rank (w3d_comm_plan: ascending dof
//       order, so both sides of a pair hold the same list).  Per iteration: one pack kernel gathers
//       this rank's partial sums on those dofs, one grouped ncclSend/ncclRecv per neighbour moves
//       them, one kernel adds what came in.  Every rank ends with the complete internal force on
//       every dof its own blocks touch; a dof shared by several ranks receives each of their
//       original partials once.  Message = the slab faces, not nodof.
//   0 = ncclAllReduce of the whole vector: the reference's semantics (every rank, hence root, holds
//       the full vector).
// The scalars and the cut flag travel as ONE 5-double all-reduce per pass.
// (Included by w3d_engine.cu inside its anonymous namespace, after iprop(); <dlfcn.h> comes from there.)
#pragma once

struct NcclUniqueId { char internal[128]; };
typedef int (*nccl_get_unique_id_t)(NcclUniqueId *);
typedef int (*nccl_comm_init_rank_t)(void **, int, NcclUniqueId, int);
typedef int (*nccl_comm_destroy_t)(void *);
typedef const char *(*nccl_get_error_string_t)(int);
typedef int (*nccl_all_reduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_broadcast_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_send_t)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_recv_t)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_group_t)(void);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // nccl.h: ncclFloat64, ncclSum

struct NcclApi {
  void *lib = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_get_error_string_t get_error_string = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  nccl_broadcast_t broadcast = nullptr;
  nccl_send_t send = nullptr;
  nccl_recv_t recv = nullptr;
  nccl_group_t group_start = nullptr, group_end = nullptr;
  std::string error;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char *names[3] = {getenv("W3D_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  // the copy already mapped into the process (e.g. by the MPI/torch layer that launched us) first
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  for (int k = 0; k < 3 && !api.lib; ++k)
    if (names[k]) api.lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) { api.error = "libnccl.so.2 not found (set W3D_NCCL_LIB)"; return &api; }
#define W3D_NCCL_SYM(field, name)                                              \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));     \
  if (!api.field) api.error = std::string("symbol missing in libnccl: ") + name;
  W3D_NCCL_SYM(get_unique_id, "ncclGetUniqueId")
  W3D_NCCL_SYM(comm_init_rank, "ncclCommInitRank")
  W3D_NCCL_SYM(comm_destroy, "ncclCommDestroy")
  W3D_NCCL_SYM(get_error_string, "ncclGetErrorString")
  W3D_NCCL_SYM(all_reduce, "ncclAllReduce")
  W3D_NCCL_SYM(broadcast, "ncclBroadcast")
  W3D_NCCL_SYM(send, "ncclSend")
  W3D_NCCL_SYM(recv, "ncclRecv")
  W3D_NCCL_SYM(group_start, "ncclGroupStart")
  W3D_NCCL_SYM(group_end, "ncclGroupEnd")
#undef W3D_NCCL_SYM
  return &api;
}

#define NCCL_TRY(expr)                                                                           \
  do {                                                                                           \
    const int _r = (expr);                                                                       \
    if (_r != 0)                                                                                 \
      return fail(W3D_ERR_COMM, "%s failed: %s (%s:%d)", #expr, nccl_api()->get_error_string(_r), \
                  __FILE__, __LINE__);                                                           \
  } while (0)

// dofs this rank's blocks share with every other rank, from the replicated block tables.
// counts[r] = number of dofs shared with rank r (0 for r == rank); dofs: the lists concatenated
// in rank order, 1-based, ascending within a list.  Host only.
int comm_plan(const w3d_engine_config &c, int rank, std::vector<long long> &counts,
              std::vector<int32_t> &dofs, std::string &err) {
  const int np = c.numprocs;
  counts.assign(np > 0 ? np : 1, 0);
  dofs.clear();
  if (np <= 1) return W3D_OK;
  if (np > 64) { err = "neighbour exchange supports at most 64 domains"; return W3D_ERR_UNSUPPORTED; }
  if (!c.elblks || !c.edest || !c.props) { err = "null table in w3d_engine_config"; return W3D_ERR_ARG; }
  std::vector<uint64_t> touched((size_t)c.nodof, 0);
  long long tab = 0;
  for (int blk = 0; blk < c.nelblk; ++blk) {
    const int span = c.elblks[4 * blk + 0], felem = c.elblks[4 * blk + 1], dom = c.elblks[4 * blk + 2];
    if (span <= 0 || felem < 1 || felem + span - 1 > c.noelem || dom < 0 || dom >= np) {
      err = "block table inconsistent (span / felem / domain)";
      return W3D_ERR_CONFIG;
    }
    const int totdof = iprop(c, 2, felem) * iprop(c, 4, felem);
    const uint64_t bit = 1ull << dom;
    for (long long k = 0; k < (long long)totdof * span; ++k) {
      const int32_t e = c.edest[tab + k];
      if (e < 1 || e > c.nodof) { err = "edest entry outside 1..nodof"; return W3D_ERR_CONFIG; }
      touched[(size_t)e - 1] |= bit;
    }
    tab += (long long)totdof * span;
  }
  const uint64_t mine = 1ull << rank;
  std::vector<std::vector<int32_t>> lists(np);
  for (int d = 0; d < c.nodof; ++d) {
    const uint64_t m = touched[d];
    if (!(m & mine) || m == mine) continue;
    for (int r = 0; r < np; ++r)
      if (r != rank && (m >> r & 1)) lists[r].push_back(d + 1);
  }
  for (int r = 0; r < np; ++r) {
    counts[r] = (long long)lists[r].size();
    dofs.insert(dofs.end(), lists[r].begin(), lists[r].end());
  }
  return W3D_OK;
}

__global__ void k_iface_pack(const double *ifv, const int32_t *idx, double *send, long long n) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    send[k] = ifv[idx[k] - 1];
}
// a dof shared with two neighbours appears in two lists: the adds go through the L2 atomics
__global__ void k_iface_add(double *ifv, const int32_t *idx, const double *recv, long long n) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    atomicAdd(&ifv[idx[k] - 1], recv[k]);
}
// [energy, plastic work, sum |f|, number of terms, cut flag] of this rank for the one all-reduce
__global__ void k_scal_pack(const double *scal, const int *flags, double num_term, double *red) {
  red[0] = scal[0]; red[1] = scal[1]; red[2] = scal[2];
  red[3] = num_term;
  red[4] = flags[0] ? 1.0 : 0.0;
}

