/*
 * w3d_oracle_assemble.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY): assembly of the element
 * stiffnesses into the reference's sparse upper-triangle ("VSS") vectors, symmetric case.
 * Follows reference src/: setup.f:158-190 (inverse incidences, ascending element numbers),
 * assemble_code.f:561-600 (dof_map), :20-250 (count_profile_symmetric[_srow]: flag-and-list
 * marking of the columns of a row), :260-500 (build_col_sparse_symm[_srow] + the ascending sort
 * of each row's columns), :659-935 (assem_by_row / assem_a_row / assem_a_row_24: scatter into a
 * full work row, copy out through k_indexes).  hex8 (24 dof, no repeated nodes) only.
 * PARITY STATUS: integer structure checked against an independent scipy assembly in
 * tests/test_assembly.py; parity unpinned against the Fortran (see w3d_oracle.h).
 */
#include <algorithm>
#include <cstdint>
#include <vector>

#include "w3d_oracle_internal.h"

namespace w3do {

struct InverseIncidences {   // main_data inverse_incidences(node)%element_count / %element_list
  std::vector<std::vector<int>> element_list;   // 1-based element numbers, in setup.f order
};

// element -> (block, relative column) = elems_to_blocks, and the offset of its (300) column in the
// packed estiff blocks; also the 24 structure dof numbers of the element (edest_blocks)
struct ElemTables {
  std::vector<int64_t> estiff_off;       // (noelem) offset of estiff(1, rel_col) of the element
  std::vector<const int32_t *> edest;    // (noelem) -> 24 dof numbers
};

static ElemTables elem_tables(const w3do_model *m) {
  ElemTables t;
  t.estiff_off.assign(m->noelem, 0);
  t.edest.assign(m->noelem, nullptr);
  int64_t oi = 0, ok = 0;
  for (int blk = 0; blk < m->nelblk; ++blk) {
    const BlockInfo b = block_info(m, blk);
    for (int i = 0; i < b.span; ++i) {
      const int elem = b.felem + i;
      t.estiff_off[elem - 1] = ok + (int64_t)i * b.utsz;
      t.edest[elem - 1] = m->edest + oi + (int64_t)i * b.totdof;
    }
    oi += (int64_t)b.totdof * b.span;
    ok += (int64_t)b.utsz * b.span;
  }
  return t;
}

// setup.f:158-190 with incid(j) = node of the element's j-th u-dof
static InverseIncidences inverse_incidences(const w3do_model *m, const ElemTables &t) {
  InverseIncidences inv;
  inv.element_list.assign(m->nonode, {});
  for (int elem = 1; elem <= m->noelem; ++elem) {
    const int32_t *ed = t.edest[elem - 1];
    for (int elnd = 1; elnd <= 8; ++elnd) {
      const int stnd = (ed[elnd - 1] - 1) / 3 + 1;
      bool seen = false;
      for (int chknod = 1; chknod <= elnd - 1; ++chknod)
        if ((ed[chknod - 1] - 1) / 3 + 1 == stnd) seen = true;
      if (seen) continue;
      inv.element_list[stnd - 1].push_back(elem);
    }
  }
  return inv;
}

}  // namespace w3do

using namespace w3do;

extern "C" {

// dof_map (assemble_code.f:561-600); returns neqns
int32_t w3do_dof_map(const int32_t *cstmap, int32_t nonode, int32_t ndof, int32_t *dof_eqn_map,
                     int32_t *eqn_node_map) {
  int eqn_counter = 0;
  for (int dof = 1; dof <= nonode * ndof; ++dof) { eqn_node_map[dof - 1] = 0; dof_eqn_map[dof - 1] = 0; }
  for (int dof = 1; dof <= ndof * nonode; ++dof) {
    if (cstmap[dof - 1] != 0) continue;
    eqn_counter = eqn_counter + 1;
    dof_eqn_map[dof - 1] = eqn_counter;
    eqn_node_map[eqn_counter - 1] = (dof - 1) / ndof + 1;
  }
  return eqn_counter;
}

// count_profile_symmetric (k_indexes == NULL: k_ptrs and the returned ncoeff only) and
// build_col_sparse_symm (k_indexes filled, each row ascending).  k_ptrs has neqns entries.
int64_t w3do_assemble_structure(const w3do_model *m, int32_t neqns, const int32_t *eqn_node_map,
                                const int32_t *dof_eqn_map, int32_t *k_ptrs, int32_t *k_indexes) {
  const ElemTables t = elem_tables(m);
  const InverseIncidences inv = inverse_incidences(m, t);
  std::vector<int> scol_flags(neqns + 1, 0), scol_list;
  int64_t ncoeff = 0;
  for (int srow = 1; srow <= neqns - 1; ++srow) {
    const int snode = eqn_node_map[srow - 1];
    scol_list.clear();
    for (int ele_on_snode : inv.element_list[snode - 1]) {
      const int32_t *ed = t.edest[ele_on_snode - 1];
      const int totdof = 24;
      for (int erow = 1; erow <= totdof; ++erow) {
        if (dof_eqn_map[ed[erow - 1] - 1] != srow) continue;
        for (int ecol = 1; ecol <= totdof; ++ecol) {
          const int scol = dof_eqn_map[ed[ecol - 1] - 1];
          if (scol <= srow) continue;
          if (scol_flags[scol] == 0) {
            scol_flags[scol] = 1;
            scol_list.push_back(scol);
          }
        }
        break;   // .not. repeat_incid
      }
    }
    for (int s : scol_list) scol_flags[s] = 0;
    k_ptrs[srow - 1] = (int32_t)scol_list.size();
    if (k_indexes) {
      std::sort(scol_list.begin(), scol_list.end());      // build_col_sparse_sort
      for (size_t j = 0; j < scol_list.size(); ++j) k_indexes[ncoeff + (int64_t)j] = scol_list[j];
    }
    ncoeff += (int64_t)scol_list.size();
  }
  k_ptrs[neqns - 1] = 0;
  return ncoeff;
}

// assem_by_row / assem_a_row / assem_a_row_24 (assemble_code.f:659-935).  estiff: the packed
// (300, span) blocks in block order; k_diag is overwritten (assembled from zero).
void w3do_assemble(const w3do_model *m, const double *estiff, int32_t neqns,
                   const int32_t *eqn_node_map, const int32_t *dof_eqn_map, const int32_t *k_ptrs,
                   const int32_t *k_indexes, double *k_diag, double *k_coeffs) {
  const ElemTables t = elem_tables(m);
  const InverseIncidences inv = inverse_incidences(m, t);
  int32_t cp[24], dcp[24];
  for (int j = 1; j <= 24; ++j) { cp[j - 1] = j * (j - 1) / 2; dcp[j - 1] = cp[j - 1] + j; }   // setup.f:268-276
  std::vector<double> coeff_row(neqns + 1, 0.0);
  std::vector<int64_t> row_start_index(neqns + 1, 1);
  for (int i = 2; i <= neqns; ++i) row_start_index[i] = row_start_index[i - 1] + k_ptrs[i - 2];
  for (int srow = 1; srow <= neqns; ++srow) {
    const int snode = eqn_node_map[srow - 1];
    for (int ele_on_snode : inv.element_list[snode - 1]) {
      const int32_t *ed = t.edest[ele_on_snode - 1];
      const double *emat = estiff + t.estiff_off[ele_on_snode - 1];
      int local_scol[24];
      for (int k = 0; k < 24; ++k) local_scol[k] = dof_eqn_map[ed[k] - 1];
      for (int erow = 1; erow <= 24; ++erow) {
        if (local_scol[erow - 1] != srow) continue;
        for (int ecol = 1; ecol <= 24; ++ecol) {
          const int scol = local_scol[ecol - 1];
          if (scol < srow) continue;
          const int kk = dcp[std::max(ecol, erow) - 1] - std::abs(ecol - erow);
          coeff_row[scol] = coeff_row[scol] + emat[kk - 1];
        }
        break;   // .not. repeated
      }
    }
    k_diag[srow - 1] = 0.0 + coeff_row[srow];
    coeff_row[srow] = 0.0;
    if (srow == neqns) continue;
    const int64_t start_loc = row_start_index[srow] - 1;
    for (int j = 1; j <= k_ptrs[srow - 1]; ++j) {
      k_coeffs[start_loc + j - 1] = coeff_row[k_indexes[start_loc + j - 1]];
      coeff_row[k_indexes[start_loc + j - 1]] = 0.0;
    }
  }
}


// convert_vss_csr (drive_assemble_solve.f:1501-1590), statement for statement (1-based arrays kept
// through the V() accessors): upper-triangle VSS pattern -> full structurally symmetric CSR.
// ptrs(neqns+1), indexes(2*ncoeff+neqns); returns nnz.
int64_t w3do_convert_vss_csr(int32_t neqns, int64_t ncoeff, const int32_t *vss_ptrs,
                             const int32_t *vss_indexes, int32_t *ptrs, int32_t *indexes) {
#define VP(i) vss_ptrs[(i)-1]
#define VI(i) vss_indexes[(i)-1]
#define P(i) ptrs[(i)-1]
#define IX(i) indexes[(i)-1]
  const int64_t nnz = (int64_t)neqns + 2 * ncoeff;
  std::vector<int32_t> workv((size_t)neqns, 0);
#define WK(i) workv[(i)-1]
  for (int i = 1; i <= neqns + 1; ++i) P(i) = 0;
  for (int i = 1; i <= neqns; ++i) P(i) = VP(i);
  for (int i = 1; i <= neqns; ++i) P(i) = P(i) + 1;
  for (int64_t i = 1; i <= ncoeff; ++i) {
    P(VI(i)) = P(VI(i)) + 1;
    WK(VI(i)) = WK(VI(i)) + 1;
  }
  for (int i = 2; i <= neqns; ++i) P(i) = P(i - 1) + P(i);
  for (int i = neqns + 1; i >= 2; --i) P(i) = P(i - 1) + 1;
  P(1) = 1;
  for (int64_t i = 1; i <= nnz; ++i) IX(i) = 0;
  int64_t sft = 1, j = 1;
  for (int i = 1; i <= neqns; ++i) {
    int64_t o = P(i) + WK(i);
    IX(o) = i;
    o = o + 1;
    for (j = sft; j <= sft + VP(i) - 1; ++j) {
      IX(o) = VI(j);
      o = o + 1;
    }
    sft = j;
  }
  for (int i = 1; i <= neqns; ++i) WK(i) = 0;
  sft = 1;
  for (int c = 1; c <= neqns; ++c) {
    for (j = sft; j <= sft + VP(c) - 1; ++j) {
      const int r = VI(j);
      const int64_t o = P(r) + WK(r);
      IX(o) = c;
      WK(r) = WK(r) + 1;
    }
    sft = j;
  }
#undef VP
#undef VI
#undef P
#undef IX
#undef WK
  return nnz;
}

}  // extern "C"
