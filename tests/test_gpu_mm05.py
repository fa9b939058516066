"""GPU parity for mm05 (cyclic plasticity, nonlinear_hardening / Frederick-Armstrong option):
engine through the C ABI vs the CPU build of the single-source model under the oracle driver."""
import numpy as np
import pytest

from common import RTOL, load_path, relerr
from warp3d_b200 import build_structured_hex8, cyclic_fa_material, new_state

pytestmark = pytest.mark.gpu


def _run(m, nsteps=4, strain=3.0e-3, tol=RTOL):
    from oracle.oracle_py import Oracle
    from warp3d_b200.engine import Engine
    eng, orc = Engine(m), Oracle(m)
    so, sg = new_state(m), new_state(m)
    nsub = 0
    for step, iters in load_path(m, nsteps=nsteps, strain=strain):
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
            for col in (0, 1, 2, 4, 5, 6, 7, 8, 9, 10):
                assert relerr(hg[:, col], ho[:, col]) <= tol, f"{tag}: history1 column {col}"
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
