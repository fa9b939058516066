"""Assembly of the element stiffnesses into the reference's sparse upper-triangle vectors
(assemble_code.f): the CPU restatement against an independent scipy assembly (CPU), and the
device assembly against the restatement, bit for bit (GPU)."""
import numpy as np
import pytest
import scipy.sparse as sp

from common import load_path
from newton_driver import assemble as scipy_assemble
from oracle import oracle_py as op
from warp3d_b200 import build_structured_hex8, new_state


def _case(nx=4, ny=3, nz=3, seed=5, block_size=7):
    m = build_structured_hex8(nx, ny, nz, coord_jitter=0.1, seed=seed, block_size=block_size)
    rng = np.random.default_rng(seed)
    cst = np.zeros(m.nodof, dtype=np.int32)
    x = m.c.reshape(-1, 3)
    for node in np.nonzero(x[:, 0] < 1e-9)[0]:            # a clamped face ...
        cst[3 * node:3 * node + 3] = 1
    cst[rng.choice(m.nodof, size=m.nodof // 9, replace=False)] = 1   # ... and scattered constraints
    return m, cst


def _vss_to_csr(s, kd, kc):
    n = s["neqns"]
    rows = np.repeat(np.arange(n), s["k_ptrs"])
    up = sp.csr_matrix((kc, (rows, s["k_indexes"] - 1)), shape=(n, n))
    return up + up.T + sp.diags(kd)


def test_restatement_matches_an_independent_assembly():
    m, cst = _case()
    o, st = op.Oracle(m), new_state(m)
    (step, iters), = list(load_path(m, nsteps=1, strain=3.0e-3))
    for itr, u, du in iters:
        o.eps_sig_ifv(st, step, itr, u, du)
    ek = o.tanstf(st, step, itr, u, du).copy()
    s = op.assemble_structure(o, cst)
    free = np.nonzero(cst == 0)[0]
    # dof_map: free dofs numbered in dof order, node of each equation
    assert s["neqns"] == free.size
    assert np.array_equal(s["dof_eqn_map"][free], np.arange(1, free.size + 1)) and not s["dof_eqn_map"][cst != 0].any()
    assert np.array_equal(s["eqn_node_map"], free // 3 + 1)
    K = scipy_assemble(m, ek)[free][:, free].tocsr()
    K.sort_indices()
    up = sp.triu(K, k=1).tocsr()
    up.sort_indices()
    # the reference keeps every structural non-zero, also where the value happens to vanish
    assert s["k_ptrs"][-1] == 0 and s["ncoeff"] == int(s["k_ptrs"].sum())
    rows = np.repeat(np.arange(s["neqns"]), s["k_ptrs"])
    assert np.all(s["k_indexes"] - 1 > rows)                                   # strictly right of the diagonal
    starts = np.concatenate([[0], np.cumsum(s["k_ptrs"])])
    for r in range(s["neqns"]):
        seg = s["k_indexes"][starts[r]:starts[r + 1]]
        assert np.all(np.diff(seg) > 0)                                        # ascending, no duplicates
        assert set(up.indices[up.indptr[r]:up.indptr[r + 1]] + 1) <= set(seg)  # covers every numeric non-zero
    kd, kc = op.assemble(o, ek, s)
    A = _vss_to_csr(s, kd, kc)
    scale = abs(K).max()
    assert abs(A - K).max() <= 1e-14 * scale
    # a free structure: no constraints at all
    s0 = op.assemble_structure(o, np.zeros(m.nodof, dtype=np.int32))
    kd0, kc0 = op.assemble(o, ek, s0)
    K0 = scipy_assemble(m, ek)
    assert s0["neqns"] == m.nodof and abs(_vss_to_csr(s0, kd0, kc0) - K0).max() <= 1e-14 * abs(K0).max()


@pytest.mark.gpu
@pytest.mark.parametrize("constrained", [True, False])
def test_device_assembly_matches_the_restatement_bit_for_bit(constrained):
    from warp3d_b200.engine import Engine
    m, cst = _case(nx=6, ny=5, nz=4, seed=11, block_size=13)
    if not constrained:
        cst[:] = 0
    e, o, st = Engine(m), op.Oracle(m), new_state(m)
    (step, iters), = list(load_path(m, nsteps=1, strain=3.0e-3))
    for itr, u, du in iters:
        e.drive_eps_sig_internal_forces(step, itr, u, du)
        o.eps_sig_ifv(st, step, itr, u, du)
    ek_dev = e.tanstf(step, itr, u, du).copy()         # also leaves the stiffnesses on the device
    sg, so = e.assemble_setup(cst), op.assemble_structure(o, cst)
    assert sg["neqns"] == so["neqns"] and sg["ncoeff"] == so["ncoeff"]
    for k in ("dof_eqn_map", "eqn_node_map", "k_ptrs", "k_indexes"):
        assert np.array_equal(sg[k], so[k]), k                                  # integers: bit exact
    kd, kc = e.assemble()
    # same element stiffnesses in, same order of additions: identical bits out
    kd_o, kc_o = op.assemble(o, ek_dev, so)
    assert np.array_equal(kd, kd_o) and np.array_equal(kc, kc_o)
    # and against the restatement's own element stiffnesses at the parity bar
    kd_r, kc_r = op.assemble(o, o.tanstf(st, step, itr, u, du), so)
    scale = max(abs(kd_r).max(), abs(kc_r).max())
    assert abs(kd - kd_r).max() <= 1e-10 * scale and abs(kc - kc_r).max() <= 1e-10 * scale
    # a second pass (resident, no host copy) leaves the same result behind
    e.tanstf(step, itr, u, du, to_host=False)
    e.assemble(to_host=False)
    kd2, kc2 = e.assemble()
    assert np.array_equal(kd2, kd) and np.array_equal(kc2, kc)
    e.close()


@pytest.mark.gpu
def test_device_assembly_refuses_what_is_not_built():
    from warp3d_b200.engine import Engine, EngineError
    m, cst = _case(nx=4, ny=3, nz=3, seed=3, block_size=6)
    e = Engine(m)
    with pytest.raises(EngineError) as err:          # assemble before set-up
        e._neqns, e._ncoeff = 1, 1
        e.assemble()
    assert "w3d_assemble_setup" in str(err.value)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ndom", [2, 3])
def test_device_assembly_by_domains_adds_up_to_the_one_domain_result(ndom):
    """numprocs > 1: every engine returns the rows of the nodes its elements touch, in the layout of
    the whole structure; rows of interface nodes come from two domains and add."""
    from warp3d_b200 import slab_domains
    from warp3d_b200.engine import Engine
    m, cst = _case(nx=5, ny=4, nz=6, seed=13, block_size=10)
    (step, iters), = list(load_path(m, nsteps=1, strain=3.0e-3))
    itr, u, du = iters[-1]

    def run(e):
        for it, uu, dd in iters:
            e.drive_eps_sig_internal_forces(step, it, uu, dd)
        e.tanstf(step, itr, u, du, to_host=False)

    e1 = Engine(m)
    run(e1)
    s1 = e1.assemble_setup(cst)
    kd1, kc1 = e1.assemble()
    assert s1["first_eqn"] == 1 and s1["n_eqns"] == s1["neqns"]
    start1 = np.concatenate([[0], np.cumsum(s1["k_ptrs"])])
    e1.close()

    m.elblks[:, 2] = slab_domains(m, ndom)
    kd_sum, kc_sum = np.zeros_like(kd1), np.zeros_like(kc1)
    covered = np.zeros(s1["neqns"], dtype=np.int32)
    for rank in range(ndom):
        e = Engine(m, myid=rank, numprocs=ndom)
        run(e)
        s = e.assemble_setup(cst)
        kd, kc = e.assemble()
        lo, n = s["first_eqn"] - 1, s["n_eqns"]
        assert 0 < n < s1["neqns"] and s["neqns"] == s1["neqns"]
        assert np.array_equal(s["dof_eqn_map"], s1["dof_eqn_map"]) and np.array_equal(s["eqn_node_map"], s1["eqn_node_map"])
        # the rows carry the structure of the whole model, bit for bit
        assert np.array_equal(s["k_ptrs"], s1["k_ptrs"][lo:lo + n])
        assert np.array_equal(s["k_indexes"], s1["k_indexes"][start1[lo]:start1[lo + n]])
        assert kd.size == n and kc.size == s["ncoeff"] == start1[lo + n] - start1[lo]
        kd_sum[lo:lo + n] += kd
        kc_sum[start1[lo]:start1[lo + n]] += kc
        covered[lo:lo + n] += 1
        e.close()
    assert covered.min() == 1 and covered.max() == 2            # interface rows come twice
    once = covered == 1
    assert np.array_equal(kd_sum[once], kd1[once])              # rows of one domain: the same bits
    rows = np.repeat(np.arange(s1["neqns"]), s1["k_ptrs"])
    assert np.array_equal(kc_sum[once[rows]], kc1[once[rows]])
    scale = max(abs(kd1).max(), abs(kc1).max())                 # interface rows: one more rounding
    assert abs(kd_sum - kd1).max() <= 1e-14 * scale and abs(kc_sum - kc1).max() <= 1e-14 * scale


def _full_csr(ptrs, idx, val, n):
    return sp.csr_matrix((val, idx.astype(np.int64) - 1, ptrs.astype(np.int64) - 1), shape=(n, n))


def test_convert_vss_csr_restatement_is_the_full_sorted_pattern():
    """convert_vss_csr (drive_assemble_solve.f:1501): the restatement's CSR pattern equals the
    pattern of the symmetric matrix scipy assembles -- offsets 1-based, columns ascending, lower
    terms before the diagonal before the upper terms, nnz = neqns + 2 ncoeff."""
    m, cst = _case(nx=5, ny=4, nz=3, seed=21, block_size=9)
    o, st = op.Oracle(m), new_state(m)
    (step, iters), = list(load_path(m, nsteps=1, strain=3.0e-3))
    for itr, u, du in iters:
        o.eps_sig_ifv(st, step, itr, u, du)
    ek = o.tanstf(st, step, itr, u, du)
    s = op.assemble_structure(o, cst)
    ptrs, idx = op.convert_vss_csr(s)
    neq = s["neqns"]
    assert ptrs[0] == 1 and ptrs[-1] == neq + 2 * s["ncoeff"] + 1 and len(idx) == neq + 2 * s["ncoeff"]
    for r in range(neq):
        cols = idx[ptrs[r] - 1:ptrs[r + 1] - 1]
        assert np.all(np.diff(cols) > 0) and (r + 1) in cols
    # the same pattern built independently: upper terms, their mirror images and the diagonal
    rows = np.repeat(np.arange(neq), s["k_ptrs"])
    cols = s["k_indexes"].astype(np.int64) - 1
    full = sp.coo_matrix((np.ones(3 * 0 + 2 * len(cols) + neq),
                          (np.concatenate([rows, cols, np.arange(neq)]), np.concatenate([cols, rows, np.arange(neq)]))),
                         shape=(neq, neq)).tocsr()
    full.sort_indices()
    assert np.array_equal(full.indptr + 1, ptrs) and np.array_equal(full.indices + 1, idx)
    assert full.data.max() == 1.0          # no term listed twice


@pytest.mark.gpu
def test_device_csr_pattern_bit_exact_and_values():
    """w3d_assemble_csr: the integers equal the restatement of convert_vss_csr bit for bit; the
    values are the symmetric stiffness in full storage (upper terms mirrored)."""
    from warp3d_b200.engine import Engine
    m, cst = _case(nx=6, ny=5, nz=4, seed=11, block_size=13)
    e, o, st = Engine(m), op.Oracle(m), new_state(m)
    (step, iters), = list(load_path(m, nsteps=1, strain=3.0e-3))
    for itr, u, du in iters:
        e.drive_eps_sig_internal_forces(step, itr, u, du)
    e.tanstf(step, itr, u, du, to_host=False)
    sg = e.assemble_setup(cst)
    kd, kc = e.assemble()
    ptrs, idx, val = e.assemble_csr()
    so = op.assemble_structure(o, cst)
    ptrs_o, idx_o = op.convert_vss_csr(so)
    assert np.array_equal(ptrs, ptrs_o) and np.array_equal(idx, idx_o)        # integers: bit exact
    neq = sg["neqns"]
    K = _full_csr(ptrs, idx, val, neq)
    A = _vss_to_csr(sg, kd, kc)                                                # upper + upper^T + diag
    assert abs(K - A).max() == 0.0
    assert abs(K - K.T).max() == 0.0
    e.close()
