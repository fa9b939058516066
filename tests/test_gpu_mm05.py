"""GPU parity for mm05 (cyclic plasticity, nonlinear_hardening / Frederick-Armstrong and
generalized_plasticity options): engine through the C ABI vs the CPU build of the single-source
model under the oracle driver."""
import numpy as np
import pytest

from common import RTOL, load_path, relerr
from warp3d_b200 import build_structured_hex8, cyclic_fa_material, cyclic_gp_material, new_state

pytestmark = pytest.mark.gpu


def _run(m, nsteps=4, strain=3.0e-3, tol=RTOL, tol_after_first_yield=None):
    """tol_after_first_yield: bound for load steps that start from a plastic state, for models whose
    update is not reproducible to 1e-10 from such a state even on one machine (generalized
    plasticity, see test_cyclic_gp_*)."""
    from oracle.oracle_py import Oracle
    from warp3d_b200.engine import Engine
    eng, orc = Engine(m), Oracle(m)
    so, sg = new_state(m), new_state(m)
    nsub = 0
    for step, iters in load_path(m, nsteps=nsteps, strain=strain):
        if step > 1 and tol_after_first_yield is not None:
            tol = tol_after_first_yield
        for itr, u, du in iters:
            fo, sco = orc.eps_sig_ifv(so, step, itr, u, du)
            fg, scg = eng.drive_eps_sig_internal_forces(step, itr, u, du)
            tag = f"step {step} iter {itr}"
            assert scg["material_cut_step"] == sco["material_cut_step"], tag
            if sco["material_cut_step"]:
                orc.adaptive_reset(so); eng.adaptive_reset()
                continue
            assert relerr(fg, fo) <= tol, f"{tag}: ifv {relerr(fg, fo):.3e}"
            for k in ("internal_energy", "plastic_work", "sum_ifv"):
                assert abs(scg[k] - sco[k]) <= tol * max(abs(sco[k]), 1e-300), (tag, k)
            eng.pull_state(sg)
            for name in ("eps_n1", "urcs_n1", "rot_n1", "cep", "element_vol"):
                e = relerr(getattr(sg, name), getattr(so, name))
                assert e <= tol, f"{tag}: {name} {e:.3e}"
            hg, ho = sg.history1.reshape(-1, 14), so.history1.reshape(-1, 14)
            assert np.array_equal(hg[:, 3].view(np.int64) & 0xFFFFFFFF, ho[:, 3].view(np.int64) & 0xFFFFFFFF), tag
            assert np.array_equal(hg[:, 11], ho[:, 11]), f"{tag}: sub-increment counts differ"
            for col in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 12):
                assert relerr(hg[:, col], ho[:, col]) <= tol, f"{tag}: history1 column {col}"
            # history(14) = dHi, the change of the isotropic modulus between sub-increments: zero up
            # to rounding on this isothermal path, so it is compared on the scale of the modulus
            h_scale = float(np.abs(m.matprp[0, 54]))
            assert np.abs(hg[:, 13] - ho[:, 13]).max() <= tol * h_scale, f"{tag}: history1 column 13"
            ko = orc.tanstf(so, step, itr, u, du).reshape(-1, 300)
            kg = eng.tanstf(step, itr, u, du).reshape(-1, 300)
            assert (np.abs(kg - ko).max(axis=1) / np.abs(ko).max(axis=1)).max() <= tol, f"{tag}: estiff"
            nsub = max(nsub, int(ho[:, 11].max()))
        orc.update(so); orc.adaptive_save(so)
        eng.update(); eng.adaptive_save()
    eng.close()
    return nsub


def test_cyclic_fa_geonl_bbar():
    m = build_structured_hex8(4, 3, 2, materials=[cyclic_fa_material()], coord_jitter=0.1, seed=21, block_size=10)
    assert _run(m) >= 3       # Richardson extrapolation with >= 3 sub-increments exercised


def test_cyclic_fa_small_strain():
    m = build_structured_hex8(3, 3, 2, materials=[cyclic_fa_material(gamma_u=120.0, h_u=30.0e9)], geonl=False,
                              coord_jitter=0.1, seed=22)
    assert _run(m, strain=4.0e-3) >= 2


# Generalized plasticity: a step that starts outside the yield surface (every step from a plastic
# state, the normal situation in this model) finds its elastic fraction by a golden-section search
# for the minimum of the yield function down to a bracket of 1e-10 (mm05.f:3290-3350).  Comparisons
# of function values that close to a minimum are rounding noise below ~1e-8, so the result depends
# on the last bits of the input: the CPU restatement compiled with and without FMA contraction
# differs from itself by 2e-8 in the plastic multiplier, 5e-10 in the tangent, 6e-11 in the
# stresses on such steps.  The first load step (elastic -> plastic, Ridder's
# bracketed root) is held to 1e-10; later steps to 2e-7, ten times the measured self-noise.
GP_NOISE = 2.0e-7


def test_cyclic_gp_geonl_bbar():
    # the reference's own generalized_plasticity parameters (verification/test39, testP; ksi)
    m = build_structured_hex8(4, 3, 2, materials=[cyclic_gp_material()], coord_jitter=0.1, seed=23, block_size=10)
    assert _run(m, strain=4.0e-3, tol_after_first_yield=GP_NOISE) >= 3


def test_cyclic_gp_small_strain_mixed_hardening():
    m = build_structured_hex8(3, 3, 2, materials=[cyclic_gp_material(gp_tau=0.4, gp_h_u=900.0, gp_delta_u=25.0)],
                              geonl=False, coord_jitter=0.1, seed=24)
    assert _run(m, strain=5.0e-3, tol_after_first_yield=GP_NOISE) >= 2
