// w3d_assemble_data.h -- device / host tables of the assembled-equations path (w3d_assemble.cuh)
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace w3d {

struct AssemblyData {
  bool ready = false;
  int neqns = 0;                     // equations of the whole structure (dof_map)
  // rows held by this domain: the equations of the nodes its elements touch, a contiguous range
  // (equations are numbered in node order).  One domain: all of them.
  int eq0 = 0, neq_local = 0;        // first equation (0-based) and count
  int node_lo = 0, node_hi = 0;      // node range [node_lo, node_hi)
  long long ncoeff = 0;              // off-diagonal terms in those rows
  // host copies handed back by w3d_assemble_structure
  std::vector<int32_t> dof_eqn_map, eqn_node_map, k_ptrs, k_indexes;
  // device
  int32_t *d_eqn_node = nullptr;     // (neqns)  0-based node of each equation
  int8_t *d_eqn_comp = nullptr;      // (neqns)  dof component 0..2
  long long *d_row_start = nullptr;  // (neq_local + 1) offsets into k_indexes / k_coeffs, local rows
  int32_t *d_kidx = nullptr;         // (ncoeff)
  int32_t *d_inv_ptr = nullptr;      // (nonode + 1) into the two lists below
  int32_t *d_inv_epos = nullptr;     // position of the element in d_estiff / d_eeqn
  int8_t *d_inv_lnode = nullptr;     // local node number (0..7) of the node in that element
  uint32_t *d_posmap = nullptr;      // [incidence - inc0][24]: row positions of the 3 x 24 terms (k_assemble_build_map)
  int inc0 = 0;                      // first incidence (d_inv_ptr[node_lo]) the map covers
  int32_t *d_dof_eqn = nullptr;      // (nodof) equation number of each structure dof, 0 = constrained
  static constexpr int NCHUNK = 8;   // node ranges whose copy to the host overlaps the next range's assembly
  std::vector<long long> h_row_start;   // (neq_local + 1) host copy of d_row_start
  std::vector<int32_t> node_first_eqn;  // (nonode + 1) first equation (0-based) at or after each node
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t chunk_done[NCHUNK] = {};
  int max_row = 0;                   // longest row (terms right of the diagonal)
  int32_t *d_eeqn = nullptr;         // (24, n_owned_elems) equation number of each element dof, 0 = constrained
  double *d_kdiag = nullptr, *d_kcoef = nullptr;
  // full structurally symmetric CSR (convert_vss_csr), built on first request
  bool csr_ready = false;
  long long csr_nnz = 0;
  std::vector<int32_t> csr_ptrs, csr_indexes;   // (neqns + 1) 1-based offsets, (nnz) 1-based columns
  int32_t *d_csr_src = nullptr;                 // (nnz) >= 0: index into k_coeffs, < 0: -(row + 1) of k_diag
  double *d_csr_val = nullptr;
  void release() {
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (int k = 0; k < NCHUNK; ++k)
      if (chunk_done[k]) cudaEventDestroy(chunk_done[k]);
    cudaFree(d_eqn_node); cudaFree(d_eqn_comp); cudaFree(d_row_start); cudaFree(d_kidx);
    cudaFree(d_inv_ptr); cudaFree(d_inv_epos); cudaFree(d_inv_lnode); cudaFree(d_posmap); cudaFree(d_eeqn); cudaFree(d_dof_eqn);
    cudaFree(d_kdiag); cudaFree(d_kcoef); cudaFree(d_csr_src); cudaFree(d_csr_val);
    *this = AssemblyData();
  }
};


}  // namespace w3d
