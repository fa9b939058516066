"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Tolerance 1e-10 relative (BASELINE.json north_star); integer history state
bit-exact."""
import numpy as np
import pytest

from common import RTOL, history_views, load_path, relerr
from warp3d_b200 import Material, build_structured_hex8, new_state, smooth_displacement

pytestmark = pytest.mark.gpu

STATE_FP = ("eps_n1", "urcs_n1", "rot_n1", "cep", "element_vol", "eps_n", "urcs_n", "rot_n")


def _engine(model):
    from warp3d_b200.engine import Engine
    return Engine(model)


def _oracle(model):
    from oracle.oracle_py import Oracle
    return Oracle(model)


def _compare_states(model, sg, so, tag, tol=RTOL):
    for name in STATE_FP:
        e = relerr(getattr(sg, name), getattr(so, name))
        assert e <= tol, f"{tag}: {name} relerr {e:.3e}"
    for name in ("history", "history1"):
        fg, ig = history_views(model, getattr(sg, name))
        fo, io = history_views(model, getattr(so, name))
        assert np.array_equal(ig, io), f"{tag}: {name} integer state differs"
        # compare each history column on its own scale (lambda, k, ebarp, H', backstress)
        for col in range(fg.shape[1]):
            e = relerr(fg[:, col], fo[:, col])
            assert e <= tol, f"{tag}: {name} column {col} relerr {e:.3e}"


def _run_path(model, nsteps=3, strain=2.0e-3, tol=RTOL):
    eng, orc = _engine(model), _oracle(model)
    so = new_state(model)
    sg = new_state(model)
    eng.push_state(sg)
    nplast = 0
    for step, iters in load_path(model, nsteps=nsteps, strain=strain):
        for itr, u, du in iters:
            fo, sco = orc.eps_sig_ifv(so, step, itr, u, du)
            fg, scg = eng.drive_eps_sig_internal_forces(step, itr, u, du)
            tag = f"step {step} iter {itr}"
            assert relerr(fg, fo) <= tol, f"{tag}: ifv relerr {relerr(fg, fo):.3e}"
            assert scg["num_term_ifv"] == sco["num_term_ifv"]
            assert scg["material_cut_step"] == sco["material_cut_step"]
            for k in ("internal_energy", "plastic_work", "sum_ifv"):
                assert abs(scg[k] - sco[k]) <= tol * max(abs(sco[k]), 1e-300), f"{tag}: {k} {scg[k]} vs {sco[k]}"
            eng.pull_state(sg)
            _compare_states(model, sg, so, tag, tol)
            ko = orc.tanstf(so, step, itr, u, du)
            kg = eng.tanstf(step, itr, u, du)
            # element stiffness: every element on its own scale
            ko2, kg2 = ko.reshape(-1, 300), kg.reshape(-1, 300)
            e = (np.abs(kg2 - ko2).max(axis=1) / np.abs(ko2).max(axis=1)).max()
            assert e <= tol, f"{tag}: estiff relerr {e:.3e}"
        _, st = history_views(model, so.history1)
        nplast += int((st == 1).sum())
        orc.update(so); orc.adaptive_save(so)
        eng.update(); eng.adaptive_save()
    eng.pull_state(sg)
    _compare_states(model, sg, so, "after update", tol)
    eng.close()
    return nplast


def test_single_block_geonl_bbar():
    m = build_structured_hex8(5, 4, 3, coord_jitter=0.15, seed=3)
    assert _run_path(m) > 0


def test_ragged_blocks_geonl_bbar():
    # 7*5*4 = 140 elements in blocks of 48 -> spans 48, 48, 44 (ragged, not a multiple of 16)
    m = build_structured_hex8(7, 5, 4, coord_jitter=0.1, seed=5, block_size=48)
    assert list(m.elblks[:, 0]) == [48, 48, 44]
    assert _run_path(m, strain=3.0e-3) > 0


def test_full_blocks_two_materials_kinematic():
    # two materials (second with mixed hardening beta=0.5) -> blocks split on material change
    mats = [Material(), Material(e=70.0e9, nu=0.33, yld_pt=250.0e6, tan_e=1.0e9, beta=0.5)]
    nx, ny, nz = 8, 8, 5
    em = np.ones(nx * ny * nz, dtype=np.int32)
    em[200:] = 2
    m = build_structured_hex8(nx, ny, nz, materials=mats, elem_material=em, coord_jitter=0.1)
    assert list(m.elblks[:, 0]) == [128, 72, 120]
    assert _run_path(m, strain=4.0e-3) > 0


@pytest.mark.parametrize("geonl,bbar", [(False, True), (False, False), (True, False)])
def test_formulation_variants(geonl, bbar):
    m = build_structured_hex8(4, 4, 3, geonl=geonl, bbar=bbar, coord_jitter=0.1, seed=11)
    assert _run_path(m, strain=3.0e-3) > 0


def test_qbar_off_and_eps_bbar():
    m = build_structured_hex8(4, 3, 3, coord_jitter=0.1, qbar_flag=False, eps_bbar=0.05)
    assert _run_path(m, strain=3.0e-3) > 0


def test_large_strain():
    m = build_structured_hex8(4, 3, 3, coord_jitter=0.1)
    assert _run_path(m, strain=4.0e-2) > 0


def test_small_strain_level_polar_noise():
    # strains ~1e-3: the reference's Cardano polar decomposition is noise dominated here; the
    # engine must still agree because it reproduces the reference's operation order exactly
    m = build_structured_hex8(6, 5, 4, coord_jitter=0.1)
    _run_path(m, strain=1.0e-3)


def test_linear_elastic_material():
    m = build_structured_hex8(3, 3, 3, materials=[Material(linear_elastic=True)], coord_jitter=0.1)
    assert _run_path(m, strain=5.0e-3) == 0


def test_step_cut_on_inverted_element():
    """det F <= 0 at a Gauss point: material_cut_step is raised and ifv zeroed
    (drive_eps_sig_internal_forces.f:253-263)."""
    m = build_structured_hex8(3, 3, 3)
    eng, orc = _engine(m), _oracle(m)
    so = new_state(m)
    u = np.zeros(m.nodof)
    du = np.zeros(m.nodof)
    # step 2 so that the request is adaptive, not fatal
    du.reshape(-1, 3)[:, 0] = -1.5 * m.c.reshape(-1, 3)[:, 0]   # collapse through zero volume
    fo, sco = orc.eps_sig_ifv(so, 2, 1, u, du)
    fg, scg = eng.drive_eps_sig_internal_forces(2, 1, u, du)
    assert sco["material_cut_step"] and scg["material_cut_step"]
    assert not fg.any() and not fo.any()
    eng.close()


def test_push_pull_roundtrip():
    m = build_structured_hex8(5, 4, 3, block_size=37)
    eng = _engine(m)
    rng = np.random.default_rng(0)
    s = new_state(m)
    for n in s.NAMES:
        a = getattr(s, n)
        a[:] = rng.standard_normal(a.shape)
    eng.push_state(s)
    t = new_state(m)
    eng.pull_state(t, names=[n for n in s.NAMES if n != "estiff"])
    for n in s.NAMES:
        if n == "estiff":
            continue
        assert np.array_equal(getattr(s, n), getattr(t, n)), n
    eng.close()
