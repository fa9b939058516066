"""BASELINE.json configs[1]: the reference's fracture_models/seb deck run with mm01 bilinear Mises
plasticity (17 318 hex8, 137 user-defined blocks with two MPI domains, geonl + B-bar,
`consistent q off`).  The mesh comes from tests/golden/seb_deck.npz (written by
tests/golden/make_seb_fixture.py from the deck files); nothing here reads /root/reference.

CPU half: the deck's tables as the reference would hold them, and the restatement on the real
(unstructured, strongly graded: element edges from 4.4e-7 m at the notch to 3e-3 m) mesh.
GPU half: the CUDA engine through the C ABI against the restatement at 1e-10, one domain and the
deck's two domains, and a Newton solve of the deck's first load steps with identical iteration
counts."""
import numpy as np
import pytest

from common import RTOL, history_views, load_path, relerr
from warp3d_b200 import new_state, seb_bend_field, seb_model


_bend_field = seb_bend_field


def test_seb_deck_tables():
    m, cons = seb_model()
    assert (m.noelem, m.nonode, m.nodof, m.nelblk) == (17318, 19935, 59805, 137)
    # incid_and_blocking:17325+: 133 full blocks and four ragged ones (77, 60, 58, 99 elements)
    spans = m.elblks[:, 0]
    assert np.count_nonzero(spans == 128) == 133 and sorted(spans[spans != 128].tolist()) == [58, 60, 77, 99]
    assert (spans[63], spans[66], spans[69], spans[136]) == (77, 60, 58, 99)
    assert np.array_equal(m.elblks[:, 1], 1 + np.concatenate([[0], np.cumsum(spans)[:-1]]))
    assert np.bincount(m.elblks[:, 2]).tolist() == [67, 70]          # the deck's two MPI domains
    assert not m.qbar_flag and m.eps_bbar == 0.0 and m.adaptive_flag   # consistent q off, ...
    assert (m.hist_size == 11).all()                                  # mm01
    assert int(m.props[0, 24]) == 1 and int(m.props[0, 17]) == -1 and int(m.props[0, 18]) == -1
    assert m.edest.min() >= 1 and m.edest.max() == m.nodof
    assert len(cons["dofs"]) == 2289 and len(np.unique(cons["dofs"])) == 2289
    assert np.count_nonzero(cons["values"]) == 15 and np.all(cons["values"][cons["values"] != 0] == -6.35e-6)


def test_seb_restatement_on_the_deck_mesh():
    """Every element of the graded mesh has a positive Jacobian at all 8 points (no step cut, the
    summed element volumes equal the specimen's quarter-model volume), the internal forces of a
    stress state are self-equilibrated, and the notch yields under the bending field."""
    from oracle.oracle_py import Oracle
    m, _ = seb_model()
    o, st = Oracle(m, nthreads=4), new_state(m)
    u, du = np.zeros(m.nodof), _bend_field(m)
    o.eps_sig_ifv(st, 1, 0, u, du)
    f, sc = o.eps_sig_ifv(st, 1, 1, u, du)
    assert not sc["material_cut_step"] and sc["num_term_ifv"] == 24 * m.noelem
    vol = st.element_vol.sum()
    x = m.c.reshape(-1, 3)
    assert vol > 0 and vol < np.prod(x.max(axis=0) - x.min(axis=0)) * 1.02
    f3 = f.reshape(-1, 3)
    assert np.abs(f3.sum(axis=0)).max() <= 1e-9 * np.abs(f3).sum()
    _, state = history_views(m, st.history1)
    nplast = int((state == 1).sum())
    assert 0 < nplast < m.noelem * 8
    k = o.tanstf(st, 1, 1, u, du).reshape(-1, 300)
    assert np.isfinite(k).all() and (k[:, [0, 2, 5]] > 0).all()


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
STATE_FP = ("eps_n1", "urcs_n1", "rot_n1", "cep", "element_vol", "eps_n", "urcs_n", "rot_n")


def _compare(m, sg, so, tag):
    for name in STATE_FP:
        e = relerr(getattr(sg, name), getattr(so, name))
        assert e <= RTOL, f"{tag}: {name} relerr {e:.3e}"
    for name in ("history", "history1"):
        fg, ig = history_views(m, getattr(sg, name))
        fo, io = history_views(m, getattr(so, name))
        assert np.array_equal(ig, io), f"{tag}: {name} integer state differs"
        for col in range(fg.shape[1]):
            e = relerr(fg[:, col], fo[:, col])
            assert e <= RTOL, f"{tag}: {name} column {col} relerr {e:.3e}"


@pytest.mark.gpu
def test_seb_engine_parity_1e10():
    """Two load steps of the bending field (iteration 0, two corrections, rollover): forces,
    scalars, every state array and all 17 318 element stiffnesses at 1e-10."""
    from oracle.oracle_py import Oracle
    from warp3d_b200.engine import Engine
    m, _ = seb_model()
    eng, orc = Engine(m), Oracle(m, nthreads=8)
    so, sg = new_state(m), new_state(m)
    u = np.zeros(m.nodof)
    nplast = 0
    for step, scale in ((1, 0.6), (2, 0.5)):
        tgt = _bend_field(m, scale)
        for itr, du in ((0, 0.3 * tgt), (1, 0.9 * tgt), (2, tgt)):
            fo, sco = orc.eps_sig_ifv(so, step, itr, u, du)
            fg, scg = eng.drive_eps_sig_internal_forces(step, itr, u, du)
            tag = f"seb step {step} iter {itr}"
            assert relerr(fg, fo) <= RTOL, f"{tag}: ifv relerr {relerr(fg, fo):.3e}"
            assert scg["num_term_ifv"] == sco["num_term_ifv"] == 24 * m.noelem
            assert scg["material_cut_step"] == sco["material_cut_step"] is False
            for k in ("internal_energy", "plastic_work", "sum_ifv"):
                assert abs(scg[k] - sco[k]) <= RTOL * max(abs(sco[k]), 1e-300), f"{tag}: {k}"
            eng.pull_state(sg)
            _compare(m, sg, so, tag)
            ko = orc.tanstf(so, step, itr, u, du).reshape(-1, 300)
            kg = eng.tanstf(step, itr, u, du).reshape(-1, 300)
            e = (np.abs(kg - ko).max(axis=1) / np.abs(ko).max(axis=1)).max()
            assert e <= RTOL, f"{tag}: estiff relerr {e:.3e}"
        _, state = history_views(m, so.history1)
        nplast += int((state == 1).sum())
        orc.update(so); orc.adaptive_save(so)
        eng.update(); eng.adaptive_save()
        u = u + tgt
    assert nplast > 1000
    eng.close()


@pytest.mark.gpu
def test_seb_two_domains_sum_to_one():
    """The deck's own domain column (blocks 1-67 -> rank 0, 68-137 -> rank 1): the internal-force
    vectors of Engine(myid=0,1; numprocs=2) add up to the one-domain vector, the scalars add, and
    each rank's state equals the one-domain state on the blocks it owns."""
    from warp3d_b200.engine import Engine
    m, _ = seb_model()
    u, du = np.zeros(m.nodof), _bend_field(m, 0.8)
    one = Engine(m)
    one.drive_eps_sig_internal_forces(1, 0, u, du)
    f1, s1 = one.drive_eps_sig_internal_forces(1, 1, u, du)
    st1 = one.pull_state(new_state(m))
    k1 = one.tanstf(1, 1, u, du).reshape(-1, 300)
    one.close()
    fsum, ssum, nterm = np.zeros(m.nodof), dict(internal_energy=0.0, plastic_work=0.0, sum_ifv=0.0), 0
    off_u, off_k = m.block_offsets(72), m.block_offsets(1)
    for r in (0, 1):
        e = Engine(m, myid=r, numprocs=2)
        e.drive_eps_sig_internal_forces(1, 0, u, du)
        f, s = e.drive_eps_sig_internal_forces(1, 1, u, du)
        fsum += f
        for k in ssum:
            ssum[k] += s[k]
        nterm += s["num_term_ifv"]
        st = e.pull_state(new_state(m))
        kr = e.tanstf(1, 1, u, du).reshape(-1, 300)
        pos = 0
        for b in e.owned_blocks:
            assert np.array_equal(st.urcs_n1[off_u[b]:off_u[b + 1]], st1.urcs_n1[off_u[b]:off_u[b + 1]])
            n = int(m.elblks[b, 0])
            assert np.array_equal(kr[pos:pos + n], k1[off_k[b]:off_k[b + 1]])
            pos += n
        e.close()
    assert nterm == s1["num_term_ifv"]
    assert relerr(fsum, f1) <= 1e-13          # atomics: order of additions differs
    for k in ssum:
        assert abs(ssum[k] - s1[k]) <= 1e-12 * abs(s1[k])


@pytest.mark.gpu
@pytest.mark.parametrize("material", ["bilinear", "mises"])
def test_seb_newton_iteration_counts(material):
    """(material = "mises": the deck AS SHIPPED, mm03 with n_power 10; "bilinear": BASELINE's
    configuration of it, mm01.)  The deck's first two load steps (displacement increments of the `constraints` file) by a
    modified Newton iteration -- one factorisation of the iteration-0 stiffness, formed by the CUDA
    engine, shared by both runs (a full-Newton factorisation of the 57 516 equations costs ~30 s in
    scipy) -- once with the restatement and once with the engine behind it: identical iteration
    counts, residual norms to 1e-8 of the force scale, displacements to 1e-9.  Convergence: the
    deck's `norm res tol 0.005` tightened tenfold (modified Newton converges linearly once the
    notch yields, ~2 000 and ~8 000 plastic points after the two steps)."""
    import scipy.sparse.linalg as spla
    from newton_driver import assemble
    from oracle.oracle_py import Oracle
    from warp3d_b200.engine import Engine
    m, cons = seb_model(material=material)
    fixed = cons["dofs"]
    free = np.setdiff1d(np.arange(m.nodof), fixed)
    eng = Engine(m)
    z = np.zeros(m.nodof)
    eng.drive_eps_sig_internal_forces(1, 0, z, z)
    K = assemble(m, eng.tanstf(1, 0, z, z))
    eng.close()
    lu = spla.splu(K[free][:, free].tocsc(), permc_spec="MMD_AT_PLUS_A", options=dict(SymmetricMode=True))
    Kfc = K[free][:, fixed]

    def run(backend_stress, update):
        u, log = np.zeros(m.nodof), []
        for step in (1, 2):
            du = np.zeros(m.nodof)
            f, _ = backend_stress(step, 0, u, du)
            norms = []
            for it in range(1, 40):
                delta = np.zeros(m.nodof)
                if it == 1:
                    delta[fixed] = 4.0 * cons["values"]         # 4 deck steps in one: plasticity at the notch
                delta[free] = lu.solve(-f[free] - Kfc @ delta[fixed])
                du = du + delta
                f, sc = backend_stress(step, it, u, du)
                assert not sc["material_cut_step"]
                rn, ref = float(np.linalg.norm(f[free])), float(np.linalg.norm(f[fixed]))
                norms.append((rn, ref))
                if rn <= 5e-4 * ref:                            # the deck's own test is 5e-3 (warp3d.inp:63)
                    break
            log.append(norms)
            u = u + du
            update()
        return u, log

    orc, so = Oracle(m, nthreads=8), new_state(m)
    uo, logo = run(lambda s, i, u, du: orc.eps_sig_ifv(so, s, i, u, du),
                   lambda: (orc.update(so), orc.adaptive_save(so)))
    eng = Engine(m)
    ug, logg = run(eng.drive_eps_sig_internal_forces, lambda: (eng.update(), eng.adaptive_save()))
    eng.close()
    assert [len(a) for a in logo] == [len(a) for a in logg], (logo, logg)
    for a, b in zip(logo, logg):
        a, b = np.array(a), np.array(b)
        assert np.all(np.abs(a - b) <= 1e-8 * a[:, 1:2]), (a, b)
    assert np.abs(ug - uo).max() <= 1e-9 * np.abs(uo).max()
    if material == "mises":
        state = so.history.reshape(-1, 15)[:, 5].view(np.int64) & 0xFFFFFFFF
    else:
        _, state = history_views(m, so.history)
    assert int((state == 1).sum()) > 0      # the notch did yield
