// w3d_assemble.cuh -- element stiffnesses -> assembled sparse equilibrium equations on the device.
// Included at the end of w3d_engine.cu (it needs the engine struct).
//
// Replaces, for the symmetric case, reference src/assemble_code.f: dof_map (:561-600),
// count_profile_symmetric (:20-250), build_col_sparse_symm (:260-500) and assem_by_row /
// assem_a_row / assem_a_row_24 (:659-935), with the inverse incidences of setup.f:158-190.
// Output = the reference's "NASA-VSS" upper-triangle format (assemble_code.f:603-620):
//   k_ptrs(neqns)    number of non-zero terms right of the diagonal on each row (last = 0)
//   k_indexes(ncoef) their column (equation) numbers, ascending within a row, 1-based
//   k_coeffs(ncoef)  the terms, k_diag(neqns) the diagonal
// Equations are the dofs without an absolute constraint, numbered in dof order (dof_map).
//
// Why on the device: the element stiffnesses are 2 400 B per element (4.8 GB for the 2 M-element
// mesh) and the host copy of them is what bounds the end-to-end Newton iteration; assembled, the
// same information is ~2.0 GB.
//
// The structure is built once per constraint set on the host (integer work, seconds); the values
// are assembled every iteration by one thread per equation, which walks the elements on the
// equation's node in the reference's order (ascending element number) and the 24 element columns
// in order, so every sum is taken in the reference's order: no atomics, bit-reproducible.
#pragma once

namespace {

// assem_a_row + assem_a_row_24 (assemble_code.f:785-935) for every equation at once
__global__ void __launch_bounds__(128) k_assemble_rows(
    int neqns, int eq0, const int32_t *__restrict__ eqn_node, const int8_t *__restrict__ eqn_comp,
    const int32_t *__restrict__ inv_ptr, const int32_t *__restrict__ inv_epos,
    const int8_t *__restrict__ inv_lnode, const int32_t *__restrict__ eeqn,
    const double *__restrict__ estiff, const long long *__restrict__ row_start,
    const int32_t *__restrict__ kidx, double *__restrict__ kdiag, double *__restrict__ kcoef) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;   // row of this domain; equation eq0 + r + 1
  if (r >= neqns) return;
  const int srow = eq0 + r + 1;
  const int node = eqn_node[eq0 + r], comp = eqn_comp[eq0 + r];
  const long long rs = row_start[r];
  const int n = (int)(row_start[r + 1] - rs);
  double *row = kcoef + rs;
  const int32_t *cols = kidx + rs;
  for (int k = 0; k < n; ++k) row[k] = 0.0;
  double diag = 0.0;
  for (int j = inv_ptr[node]; j < inv_ptr[node + 1]; ++j) {   // elements on the node, ascending
    const long long e = inv_epos[j];
    const int erow = comp * 8 + inv_lnode[j];
    const int32_t *q = eeqn + e * 24;
    const double *ek = estiff + e * 300;
#pragma unroll 4
    for (int ecol = 0; ecol < 24; ++ecol) {
      const int scol = q[ecol];
      if (scol < srow) continue;                              // lower triangle or constrained
      const int hi = max(ecol, erow), lo = min(ecol, erow);
      const double v = ek[(hi * (hi + 1)) / 2 + lo];          // dcp(max) - |ecol - erow|
      if (scol == srow) {
        diag += v;
      } else {
        int a = 0, b = n;                                     // first position with cols[] >= scol
        while (a < b) {
          const int m = (a + b) >> 1;
          if (cols[m] < scol) a = m + 1; else b = m;
        }
        row[a] += v;
      }
    }
  }
  kdiag[r] = diag;
}

// The same sums with one WARP per structure node (its up to three equations at once): lanes 0..23
// take the 24 columns of an element, the elements on the node are walked in ascending order with
// a warp barrier between them, and the rows are accumulated in shared memory and written out
// coalesced.  Within one element every column lands on a different term (no repeated nodes), so
// the per-term order of additions is again the reference's: identical bits to k_assemble_rows.
// Measured on the 2 M-element mesh: 44 ms (thread per equation, scattered global
// read-modify-writes) -> this kernel.
constexpr int ASM_WARPS = 4;       // warps (nodes) per CTA
constexpr int ASM_MAXROW = 96;     // longest row this kernel takes (81 on a structured hex8 mesh); 14 KB of
                                   // shared memory per CTA -> 16 CTAs = 64 warps per SM
__global__ void __launch_bounds__(ASM_WARPS * 32) k_assemble_nodes(
    int node0, int nonode, int eq0, const int32_t *__restrict__ dof_eqn, const int32_t *__restrict__ inv_ptr,
    const int32_t *__restrict__ inv_epos, const int8_t *__restrict__ inv_lnode,
    const int32_t *__restrict__ eeqn, const double *__restrict__ estiff,
    const long long *__restrict__ row_start, const int32_t *__restrict__ kidx,
    double *__restrict__ kdiag, double *__restrict__ kcoef) {
  __shared__ double sval[ASM_WARPS][3][ASM_MAXROW];
  __shared__ int32_t scols[ASM_WARPS][3][ASM_MAXROW];
  __shared__ double sdiag[ASM_WARPS][3];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int node = node0 + blockIdx.x * ASM_WARPS + w;   // nodes [node0, nonode)
  if (node >= nonode) return;
  int srow[3], n[3];
  long long rs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    srow[c] = dof_eqn[3 * node + c];
    rs[c] = srow[c] ? row_start[srow[c] - 1 - eq0] : 0;              // rows are stored from eq0 on
    n[c] = srow[c] ? (int)(row_start[srow[c] - eq0] - rs[c]) : 0;
    for (int k = lane; k < n[c]; k += 32) { scols[w][c][k] = kidx[rs[c] + k]; sval[w][c][k] = 0.0; }
    if (lane == 0) sdiag[w][c] = 0.0;
  }
  __syncwarp();
  // elements on the node, ascending.  The loads of element j+1 (its equation numbers and the three
  // stiffness terms of this lane's column) are issued before the shared-memory search-and-add of
  // element j: the chain index -> equation numbers / terms -> search -> add no longer runs at
  // global-memory latency once per element.
  const int jbeg = inv_ptr[node], jend = inv_ptr[node + 1];
  int scol_n = 0;
  double v_n[3] = {0.0, 0.0, 0.0};
  auto fetch = [&](int j) {
    const long long e = inv_epos[j];
    const int ln = inv_lnode[j];
    if (lane < 24) {
      scol_n = eeqn[e * 24 + lane];
      const double *ek = estiff + e * 300;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int erow = c * 8 + ln;
        const int hi = max(lane, erow), lo = min(lane, erow);
        v_n[c] = ek[(hi * (hi + 1)) / 2 + lo];                      // dcp(max) - |ecol - erow|
      }
    }
  };
  if (jbeg < jend) fetch(jbeg);
  for (int j = jbeg; j < jend; ++j) {
    const int scol = scol_n;
    const double v3[3] = {v_n[0], v_n[1], v_n[2]};
    if (j + 1 < jend) fetch(j + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (srow[c] == 0 || lane >= 24 || scol < srow[c]) continue;   // constrained / lower triangle
      const double v = v3[c];
      if (scol == srow[c]) {
        sdiag[w][c] += v;                                           // one lane per element
      } else {
        int a = 0, b = n[c];
        while (a < b) {
          const int m = (a + b) >> 1;
          if (scols[w][c][m] < scol) a = m + 1; else b = m;
        }
        sval[w][c][a] += v;                                         // distinct terms within an element
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (srow[c] == 0) continue;
    for (int k = lane; k < n[c]; k += 32) kcoef[rs[c] + k] = sval[w][c][k];
    if (lane == 0) kdiag[srow[c] - 1 - eq0] = sdiag[w][c];
  }
}

// Where every element term lands, worked out once per constraint set: for incidence j (element j of
// a node's list), lane = element column, one byte per equation of the node -- the position in the
// row (0..ASM_MAXROW-1), 0xFE = the diagonal, 0xFF = not stored (lower triangle / constrained).
// k_assemble_nodes spent its time on the three binary searches per lane and element (issue slots
// 69 % busy, profiles/r02_assemble_256k_details.csv); with the map the same additions in the same
// order need one 4-byte load instead.
__global__ void __launch_bounds__(ASM_WARPS * 32) k_assemble_build_map(
    int node0, int nonode, int eq0, const int32_t *__restrict__ dof_eqn, const int32_t *__restrict__ inv_ptr,
    const int32_t *__restrict__ inv_epos, const int32_t *__restrict__ eeqn,
    const long long *__restrict__ row_start, const int32_t *__restrict__ kidx, int inc0,
    uint32_t *__restrict__ posmap) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int node = node0 + blockIdx.x * ASM_WARPS + w;
  if (node >= nonode || lane >= 24) return;
  int srow[3], n[3];
  long long rs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    srow[c] = dof_eqn[3 * node + c];
    rs[c] = srow[c] ? row_start[srow[c] - 1 - eq0] : 0;
    n[c] = srow[c] ? (int)(row_start[srow[c] - eq0] - rs[c]) : 0;
  }
  for (int j = inv_ptr[node]; j < inv_ptr[node + 1]; ++j) {
    const int scol = eeqn[(long long)inv_epos[j] * 24 + lane];
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t k = 0xFFu;
      if (srow[c] != 0 && scol >= srow[c]) {
        if (scol == srow[c]) {
          k = 0xFEu;
        } else {
          int a = 0, b = n[c];
          while (a < b) {
            const int m = (a + b) >> 1;
            if (kidx[rs[c] + m] < scol) a = m + 1; else b = m;
          }
          k = (uint32_t)a;
        }
      }
      code |= k << (8 * c);
    }
    posmap[(long long)(j - inc0) * 24 + lane] = code;
  }
}

__global__ void __launch_bounds__(ASM_WARPS * 32) k_assemble_nodes_mapped(
    int node0, int nonode, int eq0, const int32_t *__restrict__ dof_eqn, const int32_t *__restrict__ inv_ptr,
    const int32_t *__restrict__ inv_epos, const int8_t *__restrict__ inv_lnode, const double *__restrict__ estiff,
    const long long *__restrict__ row_start, int inc0, const uint32_t *__restrict__ posmap,
    double *__restrict__ kdiag, double *__restrict__ kcoef) {
  __shared__ double sval[ASM_WARPS][3][ASM_MAXROW];
  __shared__ double sdiag[ASM_WARPS][3];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int node = node0 + blockIdx.x * ASM_WARPS + w;
  if (node >= nonode) return;
  int srow[3], n[3];
  long long rs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    srow[c] = dof_eqn[3 * node + c];
    rs[c] = srow[c] ? row_start[srow[c] - 1 - eq0] : 0;
    n[c] = srow[c] ? (int)(row_start[srow[c] - eq0] - rs[c]) : 0;
    for (int k = lane; k < n[c]; k += 32) sval[w][c][k] = 0.0;
    if (lane == 0) sdiag[w][c] = 0.0;
  }
  __syncwarp();
  const int jbeg = inv_ptr[node], jend = inv_ptr[node + 1];
  uint32_t code_n = 0xFFFFFFu;
  double v_n[3] = {0.0, 0.0, 0.0};
  auto fetch = [&](int j) {
    if (lane < 24) {
      code_n = posmap[(long long)(j - inc0) * 24 + lane];
      const int ln = inv_lnode[j];
      const double *ek = estiff + (long long)inv_epos[j] * 300;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int erow = c * 8 + ln;
        const int hi = max(lane, erow), lo = min(lane, erow);
        v_n[c] = ek[(hi * (hi + 1)) / 2 + lo];
      }
    }
  };
  if (jbeg < jend) fetch(jbeg);
  for (int j = jbeg; j < jend; ++j) {       // elements on the node, ascending: the reference's order of additions
    const uint32_t code = code_n;
    const double v3[3] = {v_n[0], v_n[1], v_n[2]};
    if (j + 1 < jend) fetch(j + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t k = (code >> (8 * c)) & 0xFFu;
      if (lane >= 24 || k == 0xFFu) continue;
      if (k == 0xFEu) sdiag[w][c] += v3[c];          // one lane per element and equation
      else sval[w][c][k] += v3[c];                   // distinct terms within an element
    }
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (srow[c] == 0) continue;
    for (int k = lane; k < n[c]; k += 32) kcoef[rs[c] + k] = sval[w][c][k];
    if (lane == 0) kdiag[srow[c] - 1 - eq0] = sdiag[w][c];
  }
}

// values of the assembled upper triangle into the full CSR pattern (diagonal, upper and the
// mirrored lower terms): one gather per stored term
__global__ void k_csr_fill(const int32_t *__restrict__ src, const double *__restrict__ kdiag,
                           const double *__restrict__ kcoef, double *__restrict__ val, long long nnz) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
    const int32_t s = src[k];
    val[k] = (s < 0) ? kdiag[-(s + 1)] : kcoef[s];
  }
}

}  // namespace

extern "C" {

// dof_map + inverse incidences + count_profile_symmetric + build_col_sparse_symm.
// cstmap(nodof): 0 = free dof (as the reference's cstmap); edest: the same packed per-block
// (totdof, span) tables as in the configuration.
int32_t w3d_assemble_setup(w3d_engine *h, const int32_t *cstmap, const int32_t *edest,
                           int32_t *neqns_out, int64_t *ncoeff_out) {
  if (!h || !cstmap || !edest) return h ? h->fail(W3D_ERR_ARG, "null argument") : W3D_ERR_ARG;
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  const w3d_engine_config &c = h->cfg;
  CUDA_TRY(cudaSetDevice(h->device));
  AssemblyData &A = h->assembly;
  A.release();
  const int nodof = c.nodof, nonode = c.nonode;
  if (nodof != 3 * nonode) return h->fail(W3D_ERR_CONFIG, "assembly expects 3 dofs per node");
  // ---- dof_map (assemble_code.f:561-600)
  A.dof_eqn_map.assign(nodof, 0);
  A.eqn_node_map.clear();
  std::vector<int8_t> eqn_comp;
  int neq = 0;
  for (int dof = 0; dof < nodof; ++dof) {
    if (cstmap[dof] != 0) continue;
    A.dof_eqn_map[dof] = ++neq;
    A.eqn_node_map.push_back(dof / 3 + 1);
    eqn_comp.push_back((int8_t)(dof % 3));
  }
  A.neqns = neq;
  if (neq == 0) return h->fail(W3D_ERR_CONFIG, "no free dof");
  // ---- element tables, inverse incidences (setup.f:158-190: ascending elements).  Two sets:
  //      every element of the model for the sparsity (the rows are those of the whole structure,
  //      whichever domain holds the contributing elements), this domain's elements in d_estiff
  //      order for the values.
  const long long nel = h->n_owned_elems;
  const int noelem = c.noelem;
  std::vector<int32_t> eeqn((size_t)nel * 24);          // owned, by position in d_estiff
  std::vector<int32_t> eeqn_all((size_t)noelem * 24);   // all, by element number
  std::vector<int32_t> enode_all((size_t)noelem * 8);
  std::vector<int32_t> cnt_all(nonode + 1, 0), cnt_own(nonode + 1, 0);
  std::vector<long long> pos_of((size_t)noelem, -1);     // element number - 1 -> position, -1 = not owned
  int node_lo = nonode, node_hi = 0;
  for (const BlockInfo &b : h->blocks) {
    if (b.totdof != 24) return h->fail(W3D_ERR_UNSUPPORTED, "assembly: 24-dof elements only");
    for (int i = 0; i < b.span; ++i) {
      const int elem0 = b.felem - 1 + i;
      if (elem0 < 0 || elem0 >= noelem) return h->fail(W3D_ERR_CONFIG, "element number out of range in block table");
      const int32_t *ed = edest + b.tab_off + (size_t)i * 24;
      int32_t *q = &eeqn_all[(size_t)elem0 * 24];
      int32_t *en = &enode_all[(size_t)elem0 * 8];
      for (int k = 0; k < 24; ++k) {
        const int d = ed[k];
        if (d < 1 || d > nodof) return h->fail(W3D_ERR_CONFIG, "edest entry %d out of range", d);
        q[k] = A.dof_eqn_map[d - 1];
      }
      for (int k = 0; k < 8; ++k) {
        const int nd = (ed[k] - 1) / 3;   // 0-based node of local node k (u-dof = dstmap(node))
        if (ed[8 + k] != ed[k] + 1 || ed[16 + k] != ed[k] + 2)
          return h->fail(W3D_ERR_CONFIG, "edest is not in the u(8), v(8), w(8) order of a node-major dof map");
        for (int p = 0; p < k; ++p)
          if (en[p] == nd)
            return h->fail(W3D_ERR_UNSUPPORTED, "assembly: repeated nodes in an element (collapsed elements) not built");
        en[k] = nd;
        ++cnt_all[nd + 1];
        if (b.owned) {
          ++cnt_own[nd + 1];
          node_lo = std::min(node_lo, nd);
          node_hi = std::max(node_hi, nd + 1);
        }
      }
      if (b.owned) {
        pos_of[elem0] = b.epos0 + i;
        std::memcpy(&eeqn[(size_t)(b.epos0 + i) * 24], q, sizeof(int32_t) * 24);
      }
    }
  }
  if (node_hi <= node_lo) return h->fail(W3D_ERR_CONFIG, "assembly: this domain owns no element");
  // inverse incidences by ascending element number
  std::vector<int32_t> ptr_all(nonode + 1, 0), inv_ptr(nonode + 1, 0);
  for (int nd = 0; nd < nonode; ++nd) {
    ptr_all[nd + 1] = ptr_all[nd] + cnt_all[nd + 1];
    inv_ptr[nd + 1] = inv_ptr[nd] + cnt_own[nd + 1];
  }
  std::vector<int32_t> fill_all(ptr_all.begin(), ptr_all.end() - 1), fill(inv_ptr.begin(), inv_ptr.end() - 1);
  std::vector<int32_t> inv_all((size_t)ptr_all[nonode]);      // element numbers - 1
  std::vector<int32_t> inv_epos((size_t)inv_ptr[nonode]);
  std::vector<int8_t> inv_lnode((size_t)inv_ptr[nonode]);
  for (int elem0 = 0; elem0 < noelem; ++elem0) {
    const long long pos = pos_of[elem0];
    for (int k = 0; k < 8; ++k) {
      const int nd = enode_all[(size_t)elem0 * 8 + k];
      inv_all[fill_all[nd]++] = elem0;
      if (pos >= 0) {
        inv_epos[fill[nd]] = (int32_t)pos;
        inv_lnode[fill[nd]] = (int8_t)k;
        ++fill[nd];
      }
    }
  }
  // ---- the rows of this domain: equations of the nodes [node_lo, node_hi)
  A.node_first_eqn.assign(nonode + 1, neq);
  for (int r = neq - 1; r >= 0; --r) A.node_first_eqn[A.eqn_node_map[r] - 1] = r;
  for (int nd = nonode - 1; nd >= 0; --nd)          // nodes without a free dof take the next node's
    if (A.node_first_eqn[nd] > A.node_first_eqn[nd + 1]) A.node_first_eqn[nd] = A.node_first_eqn[nd + 1];
  A.node_lo = node_lo; A.node_hi = node_hi;
  A.eq0 = A.node_first_eqn[node_lo];
  A.neq_local = A.node_first_eqn[node_hi] - A.eq0;
  if (A.neq_local <= 0) return h->fail(W3D_ERR_CONFIG, "assembly: no free dof on this domain's nodes");
  const int nloc = A.neq_local;
  // ---- sparsity: per equation the distinct columns right of the diagonal, ascending.  The up
  //      to three equations of a node see the same elements, so the node's sorted list of distinct
  //      equations is formed once and every row takes its tail; node ranges are dealt to host
  //      threads (6.1 M equations: 4.6 s single-threaded in round 1).
  A.k_ptrs.assign(nloc, 0);
  std::vector<long long> row_start(nloc + 1, 0);
  {
    const int nnode_loc = node_hi - node_lo;
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(1u, std::min(nthr, 32u));
    if (nnode_loc < 4096) nthr = 1;
    std::vector<std::vector<int32_t>> part(nthr);
    std::vector<std::thread> pool;
    auto work = [&](unsigned t) {
      const int na = node_lo + (int)((long long)nnode_loc * t / nthr), nb = node_lo + (int)((long long)nnode_loc * (t + 1) / nthr);
      std::vector<int32_t> &out = part[t];
      out.reserve((size_t)(nb - na) * 3 * 40);
      std::vector<int32_t> cand;
      cand.reserve(27 * 24);
      for (int nd = na; nd < nb; ++nd) {
        const int r0 = A.node_first_eqn[nd], r1 = A.node_first_eqn[nd + 1];   // equations of this node
        if (r1 <= r0) continue;
        cand.clear();
        for (int j = ptr_all[nd]; j < ptr_all[nd + 1]; ++j) {
          const int32_t *q = &eeqn_all[(size_t)inv_all[j] * 24];
          for (int k = 0; k < 24; ++k)
            if (q[k] > r0 + 1) cand.push_back(q[k]);
        }
        std::sort(cand.begin(), cand.end());
        cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
        for (int r = r0; r < r1; ++r) {
          const int srow = r + 1;
          const auto first = std::upper_bound(cand.begin(), cand.end(), srow);
          A.k_ptrs[r - A.eq0] = (int32_t)(cand.end() - first);
          out.insert(out.end(), first, cand.end());
        }
      }
    };
    for (unsigned t = 1; t < nthr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int r = 0; r < nloc; ++r) row_start[r + 1] = row_start[r] + A.k_ptrs[r];
    A.k_indexes.resize((size_t)row_start[nloc]);
    size_t off = 0;
    for (unsigned t = 0; t < nthr; ++t) {
      if (!part[t].empty()) std::memcpy(A.k_indexes.data() + off, part[t].data(), sizeof(int32_t) * part[t].size());
      off += part[t].size();
      std::vector<int32_t>().swap(part[t]);
    }
  }
  A.ncoeff = row_start[nloc];
  // ---- upload
  std::vector<int32_t> eqn_node0(neq);
  for (int r = 0; r < neq; ++r) eqn_node0[r] = A.eqn_node_map[r] - 1;
  auto up = [&](auto **dptr, const auto &v) -> cudaError_t {
    using T = typename std::remove_reference<decltype(v[0])>::type;
    const size_t bytes = sizeof(T) * (v.size() ? v.size() : 1);
    cudaError_t e = cudaMalloc((void **)dptr, bytes);
    if (e != cudaSuccess) return e;
    return v.size() ? cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice) : cudaSuccess;
  };
  CUDA_TRY(up(&A.d_eqn_node, eqn_node0));
  CUDA_TRY(up(&A.d_eqn_comp, eqn_comp));
  CUDA_TRY(up(&A.d_row_start, row_start));
  CUDA_TRY(up(&A.d_kidx, A.k_indexes));
  CUDA_TRY(up(&A.d_inv_ptr, inv_ptr));
  CUDA_TRY(up(&A.d_inv_epos, inv_epos));
  CUDA_TRY(up(&A.d_inv_lnode, inv_lnode));
  CUDA_TRY(up(&A.d_eeqn, eeqn));
  CUDA_TRY(up(&A.d_dof_eqn, A.dof_eqn_map));
  A.h_row_start = row_start;
  CUDA_TRY(cudaStreamCreateWithFlags(&A.copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < AssemblyData::NCHUNK; ++k)
    CUDA_TRY(cudaEventCreateWithFlags(&A.chunk_done[k], cudaEventDisableTiming));
  A.max_row = 0;
  for (int r = 0; r < nloc; ++r) A.max_row = std::max(A.max_row, (int)A.k_ptrs[r]);
  CUDA_TRY(cudaMalloc((void **)&A.d_kdiag, sizeof(double) * (size_t)nloc));
  CUDA_TRY(cudaMalloc((void **)&A.d_kcoef, sizeof(double) * (size_t)(A.ncoeff ? A.ncoeff : 1)));
  // position map of every element term (k_assemble_build_map), 96 B per (node, element) incidence;
  // skipped when it would not fit comfortably (the searching kernel is used then)
  cudaFree(A.d_posmap); A.d_posmap = nullptr;
  if (A.max_row <= ASM_MAXROW && A.node_hi > A.node_lo) {
    A.inc0 = inv_ptr[A.node_lo];
    const long long ninc = (long long)inv_ptr[A.node_hi] - A.inc0;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const size_t need = sizeof(uint32_t) * 24 * (size_t)(ninc > 0 ? ninc : 1);
    if (need < free_b / 4 && cudaMalloc((void **)&A.d_posmap, need) == cudaSuccess) {
      const int nn = A.node_hi - A.node_lo;
      k_assemble_build_map<<<(nn + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, h->stream>>>(
          A.node_lo, A.node_hi, A.eq0, A.d_dof_eqn, A.d_inv_ptr, A.d_inv_epos, A.d_eeqn, A.d_row_start, A.d_kidx,
          A.inc0, A.d_posmap);
      h->launches += 1;
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else {
      cudaGetLastError();
      A.d_posmap = nullptr;
    }
  }
  A.ready = true;
  if (neqns_out) *neqns_out = neq;
  if (ncoeff_out) *ncoeff_out = A.ncoeff;
  return W3D_OK;
}

// the rows this domain assembles: equations first_eqn .. first_eqn + n_eqns - 1 (1-based)
int32_t w3d_assemble_rows(w3d_engine *h, int32_t *first_eqn, int32_t *n_eqns) {
  if (!h) return W3D_ERR_ARG;
  const AssemblyData &A = h->assembly;
  if (!A.ready) return h->fail(W3D_ERR_ARG, "w3d_assemble_setup has not been called");
  if (first_eqn) *first_eqn = A.eq0 + 1;
  if (n_eqns) *n_eqns = A.neq_local;
  return W3D_OK;
}

// the integer structure built by w3d_assemble_setup; any pointer may be NULL
int32_t w3d_assemble_structure(w3d_engine *h, int32_t *dof_eqn_map, int32_t *eqn_node_map,
                               int32_t *k_ptrs, int32_t *k_indexes) {
  if (!h) return W3D_ERR_ARG;
  const AssemblyData &A = h->assembly;
  if (!A.ready) return h->fail(W3D_ERR_ARG, "w3d_assemble_setup has not been called");
  if (dof_eqn_map) std::memcpy(dof_eqn_map, A.dof_eqn_map.data(), sizeof(int32_t) * A.dof_eqn_map.size());
  if (eqn_node_map) std::memcpy(eqn_node_map, A.eqn_node_map.data(), sizeof(int32_t) * A.eqn_node_map.size());
  if (k_ptrs) std::memcpy(k_ptrs, A.k_ptrs.data(), sizeof(int32_t) * A.k_ptrs.size());
  if (k_indexes) std::memcpy(k_indexes, A.k_indexes.data(), sizeof(int32_t) * A.k_indexes.size());
  return W3D_OK;
}

// assem_by_row on the element stiffnesses of the last w3d_tanstf / w3d_tanstf_resident.
// k_diag(neqns), k_coeffs(ncoeff): host buffers, or both NULL to leave the result on the device.
int32_t w3d_assemble(w3d_engine *h, double *k_diag, double *k_coeffs) {
  if (!h) return W3D_ERR_ARG;
  AssemblyData &A = h->assembly;
  if (!A.ready) return h->fail(W3D_ERR_ARG, "w3d_assemble_setup has not been called");
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  CUDA_TRY(cudaSetDevice(h->device));
  if (A.max_row <= ASM_MAXROW && k_diag && k_coeffs && A.copy_stream) {
    // host output: assemble node ranges one after the other and start the copy of each range's
    // rows (equations are numbered in node order, so they are contiguous) on a second stream as
    // soon as its kernel is done -- the assembly hides behind the PCIe transfer
    const int nnode = A.node_hi - A.node_lo;
    const int nchunk = AssemblyData::NCHUNK;
    for (int k = 0; k < nchunk; ++k) {
      const int na = A.node_lo + (int)((long long)nnode * k / nchunk), nb = A.node_lo + (int)((long long)nnode * (k + 1) / nchunk);
      if (nb <= na) continue;
      if (A.d_posmap)
        k_assemble_nodes_mapped<<<(nb - na + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, h->stream>>>(
            na, nb, A.eq0, A.d_dof_eqn, A.d_inv_ptr, A.d_inv_epos, A.d_inv_lnode, h->d_estiff, A.d_row_start, A.inc0,
            A.d_posmap, A.d_kdiag, A.d_kcoef);
      else
        k_assemble_nodes<<<(nb - na + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, h->stream>>>(
            na, nb, A.eq0, A.d_dof_eqn, A.d_inv_ptr, A.d_inv_epos, A.d_inv_lnode, A.d_eeqn, h->d_estiff,
            A.d_row_start, A.d_kidx, A.d_kdiag, A.d_kcoef);
      CUDA_TRY(cudaGetLastError());
      h->launches += 1;
      CUDA_TRY(cudaEventRecord(A.chunk_done[k], h->stream));
      CUDA_TRY(cudaStreamWaitEvent(A.copy_stream, A.chunk_done[k], 0));
      const int ea = A.node_first_eqn[na] - A.eq0, eb = A.node_first_eqn[nb] - A.eq0;   // local rows [ea, eb)
      if (eb > ea) {
        const long long ca = A.h_row_start[ea], cb = A.h_row_start[eb];
        CUDA_TRY(cudaMemcpyAsync(k_diag + ea, A.d_kdiag + ea, sizeof(double) * (size_t)(eb - ea), cudaMemcpyDeviceToHost, A.copy_stream));
        if (cb > ca)
          CUDA_TRY(cudaMemcpyAsync(k_coeffs + ca, A.d_kcoef + ca, sizeof(double) * (size_t)(cb - ca), cudaMemcpyDeviceToHost, A.copy_stream));
      }
    }
    CUDA_TRY(cudaStreamSynchronize(A.copy_stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return W3D_OK;
  }
  if (A.max_row <= ASM_MAXROW && A.d_posmap) {
    k_assemble_nodes_mapped<<<(A.node_hi - A.node_lo + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, h->stream>>>(
        A.node_lo, A.node_hi, A.eq0, A.d_dof_eqn, A.d_inv_ptr, A.d_inv_epos, A.d_inv_lnode, h->d_estiff, A.d_row_start,
        A.inc0, A.d_posmap, A.d_kdiag, A.d_kcoef);
  } else if (A.max_row <= ASM_MAXROW) {
    k_assemble_nodes<<<(A.node_hi - A.node_lo + ASM_WARPS - 1) / ASM_WARPS, ASM_WARPS * 32, 0, h->stream>>>(
        A.node_lo, A.node_hi, A.eq0, A.d_dof_eqn, A.d_inv_ptr, A.d_inv_epos, A.d_inv_lnode, A.d_eeqn, h->d_estiff,
        A.d_row_start, A.d_kidx, A.d_kdiag, A.d_kcoef);
  } else {   // very long rows (nodes shared by many elements): thread per equation
    const int tpb = 128;
    k_assemble_rows<<<(A.neq_local + tpb - 1) / tpb, tpb, 0, h->stream>>>(
        A.neq_local, A.eq0, A.d_eqn_node, A.d_eqn_comp, A.d_inv_ptr, A.d_inv_epos, A.d_inv_lnode, A.d_eeqn,
        h->d_estiff, A.d_row_start, A.d_kidx, A.d_kdiag, A.d_kcoef);
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  if (k_diag) CUDA_TRY(cudaMemcpyAsync(k_diag, A.d_kdiag, sizeof(double) * (size_t)A.neq_local, cudaMemcpyDeviceToHost, h->stream));
  if (k_coeffs && A.ncoeff) CUDA_TRY(cudaMemcpyAsync(k_coeffs, A.d_kcoef, sizeof(double) * (size_t)A.ncoeff, cudaMemcpyDeviceToHost, h->stream));
  if (k_diag || k_coeffs) CUDA_TRY(cudaStreamSynchronize(h->stream));
  return W3D_OK;
}


// convert_vss_csr (drive_assemble_solve.f:1501-1590): the upper-triangle VSS pattern turned into
// the full, structurally symmetric CSR pattern (per row: lower terms, diagonal, upper terms;
// columns ascending; 1-based offsets in k_ptrs(neqns+1)) that the reference hands to hypre and to
// the asymmetric Pardiso path -- same statements, same order, so the integers are the
// reference's.  k_coeffs (nnz) receives the values of the last w3d_assemble in that pattern
// (upper terms mirrored into the lower triangle: the symmetric stiffness in full storage).  Any
// pointer may be NULL; the pattern is built on the first call after w3d_assemble_setup.  One
// domain only (the reference converts on root after the merge).
int32_t w3d_assemble_csr(w3d_engine *h, int64_t *nnz_out, int32_t *k_ptrs, int32_t *k_indexes, double *k_coeffs) {
  if (!h) return W3D_ERR_ARG;
  AssemblyData &A = h->assembly;
  if (!A.ready) return h->fail(W3D_ERR_ARG, "w3d_assemble_setup has not been called");
  auto fail = [&](int code, const char *fmt, auto... args) { return h->fail(code, fmt, args...); };
  if (A.neq_local != A.neqns) return fail(W3D_ERR_UNSUPPORTED, "w3d_assemble_csr: one domain only (rows %d of %d held here)", A.neq_local, A.neqns);
  CUDA_TRY(cudaSetDevice(h->device));
  const int neqns = A.neqns;
  const long long ncoeff = A.ncoeff;
  if (!A.csr_ready) {
    const long long nnz = (long long)neqns + 2 * ncoeff;
    if (nnz > 2147483000LL) return fail(W3D_ERR_UNSUPPORTED, "w3d_assemble_csr: %lld terms exceed default-integer offsets", nnz);
    const int32_t *vss_ptrs = A.k_ptrs.data();
    const int32_t *vss_indexes = A.k_indexes.data();
    std::vector<int32_t> &ptrs = A.csr_ptrs, &indexes = A.csr_indexes;
    std::vector<int32_t> src((size_t)nnz), work((size_t)neqns, 0);
    ptrs.assign((size_t)neqns + 1, 0);
    indexes.assign((size_t)nnz, 0);
    // number of terms per row: upper terms + diagonal + terms per column
    for (int i = 0; i < neqns; ++i) ptrs[i] = vss_ptrs[i] + 1;
    for (long long i = 0; i < ncoeff; ++i) {
      ptrs[vss_indexes[i] - 1] += 1;
      work[vss_indexes[i] - 1] += 1;
    }
    // convert to 1-based offsets
    for (int i = 1; i < neqns; ++i) ptrs[i] = ptrs[i - 1] + ptrs[i];
    for (int i = neqns; i >= 1; --i) ptrs[i] = ptrs[i - 1] + 1;
    ptrs[0] = 1;
    // upper triangle and diagonal
    long long sidx = 0;
    for (int i = 0; i < neqns; ++i) {
      long long o = (long long)ptrs[i] - 1 + work[i];      // the diagonal's slot (0-based)
      indexes[o] = i + 1;
      src[o] = -(i + 1);
      ++o;
      for (int j = 0; j < vss_ptrs[i]; ++j, ++sidx, ++o) {
        indexes[o] = vss_indexes[sidx];
        src[o] = (int32_t)sidx;
      }
    }
    // lower triangle: row r, column c, columns strictly increasing
    std::fill(work.begin(), work.end(), 0);
    sidx = 0;
    for (int cc = 0; cc < neqns; ++cc)
      for (int j = 0; j < vss_ptrs[cc]; ++j, ++sidx) {
        const int r = vss_indexes[sidx] - 1;
        const long long o = (long long)ptrs[r] - 1 + work[r];
        indexes[o] = cc + 1;
        src[o] = (int32_t)sidx;
        work[r] += 1;
      }
    A.csr_nnz = nnz;
    CUDA_TRY(cudaMalloc((void **)&A.d_csr_src, sizeof(int32_t) * (size_t)nnz));
    CUDA_TRY(cudaMemcpy(A.d_csr_src, src.data(), sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void **)&A.d_csr_val, sizeof(double) * (size_t)nnz));
    A.csr_ready = true;
  }
  if (nnz_out) *nnz_out = A.csr_nnz;
  if (k_ptrs) std::memcpy(k_ptrs, A.csr_ptrs.data(), sizeof(int32_t) * A.csr_ptrs.size());
  if (k_indexes) std::memcpy(k_indexes, A.csr_indexes.data(), sizeof(int32_t) * A.csr_indexes.size());
  if (k_coeffs) {
    const int grid = (int)std::min<long long>((A.csr_nnz + 255) / 256, 148LL * 16);
    k_csr_fill<<<grid, 256, 0, h->stream>>>(A.d_csr_src, A.d_kdiag, A.d_kcoef, A.d_csr_val, A.csr_nnz);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(k_coeffs, A.d_csr_val, sizeof(double) * (size_t)A.csr_nnz, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return W3D_OK;
}

}  // extern "C"
