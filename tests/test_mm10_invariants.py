"""CPU checks of the crystal-plasticity (mm10, Voce) core through the oracle's block driver:
consistent tangent vs finite differences, Taylor averaging over crystals, elastic limit, rigid
rotation, rate dependence.  (The core is single-source with the engine; these tests are what
argues its fidelity to mm10_a.f / mm10_b.f besides the line citations.)"""
import numpy as np
import pytest

from common import relerr
from newton_driver import assemble
from oracle import oracle_py as op
from warp3d_b200 import (Material, attach_crystals, build_structured_hex8, make_crystal_type,
                         mm10_hist_size, new_state, smooth_displacement)

CT = dict(e=100000.0, nu=0.3, harden_n=20, tau_y=100.0, tau_v=50.0, theta_o=500.0, voche_m=1.0)


def _model(n=(1, 1, 1), ncrystals=1, geonl=False, angles=None, dt=1.0, **ct):
    kw = dict(CT); kw.update(ct)
    m = build_structured_hex8(*n, materials=[Material(mat_type=10)], geonl=geonl, bbar=True, dt=dt)
    return attach_crystals(m, make_crystal_type(**kw), ncrystals=ncrystals, angles=angles)


def _uniform(m, H):
    return (m.c.reshape(-1, 3) @ np.asarray(H).T).reshape(-1)


def test_history_size_matches_mm10_set_history_locs():
    assert mm10_hist_size(12, 1) == 87 + 71 == 158          # fcc, 1 crystal
    assert mm10_hist_size(48, 25) == 123 + 25 * 234 == 5973  # simple_voce: bcc48 x 25 crystals
    assert op.lib().w3do_mm10_hist_size(12, 1, 3) == 87 + 3 * 71


def test_elastic_response_is_the_rotated_cubic_stiffness():
    """far below tau_y the update is C : eps with C the crystal stiffness rotated by the Kocks
    angles; for an isotropic crystal the orientation must not matter."""
    H = np.diag([1.0e-5, -3.0e-6, -3.0e-6]) + 1.0e-6 * np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]])
    sig = []
    for ang in ([[[0.0, 0.0, 0.0]]], [[[35.0, 70.0, 110.0]]]):
        m = _model(angles=np.array(ang))
        o, st = op.Oracle(m), new_state(m)
        o.eps_sig_ifv(st, 1, 1, np.zeros(m.nodof), _uniform(m, H))
        sig.append(st.urcs_n1.reshape(-1, 9)[:, :6].copy())
    E, nu = CT["e"], CT["nu"]
    lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    eps = np.array([H[0, 0], H[1, 1], H[2, 2], 2 * H[0, 1], 0.0, 0.0])
    ref = np.r_[2 * G * eps[:3] + lam * eps[:3].sum(), G * eps[3:]]
    assert relerr(sig[0], np.tile(ref, (8, 1))) < 1e-9       # slip at 1e-5 strain is ~1e-40
    assert relerr(sig[1], sig[0]) < 1e-12


def test_taylor_average_over_crystals():
    ang = np.array([[[10.0, 20.0, 30.0], [100.0, 50.0, 200.0], [300.0, 120.0, 45.0]]])
    H = np.diag([3.0e-3, -1.35e-3, -1.35e-3])
    m3 = _model(ncrystals=3, angles=ang)
    o3, s3 = op.Oracle(m3), new_state(m3)
    o3.eps_sig_ifv(s3, 1, 1, np.zeros(m3.nodof), _uniform(m3, H))
    singles = []
    for c in range(3):
        m1 = _model(angles=ang[:, c:c + 1])
        o1, s1 = op.Oracle(m1), new_state(m1)
        o1.eps_sig_ifv(s1, 1, 1, np.zeros(m1.nodof), _uniform(m1, H))
        singles.append(s1.urcs_n1.reshape(-1, 9)[:, :6].copy())
        assert s1.urcs_n1.reshape(-1, 9)[0, 8] > 1e-4        # plastic
    assert relerr(s3.urcs_n1.reshape(-1, 9)[:, :6], sum(singles) / 3.0) < 1e-13


@pytest.mark.parametrize("alter,bound", [(False, 2.0e-2), (True, 3.0e-3)])
def test_tangent_against_finite_differences(alter, bound):
    """mm10_tangent is the Schur complement of the local Jacobian.  It is exact up to (i) the
    Jacobian being the one of the last Newton pass, (ii) the symmetrisation (mm10_a_make_symm_1)
    and (iii) dgdd = 0 (mm10_dgdd_voche), i.e. the dependence of the slip pre-factor
    dg = |d eps|_eff on the strain increment is dropped by the reference: ~1 % here.  With the
    Voce 'alter_mode' (fixed pre-factor) only (i)+(ii) remain: ~0.1 %."""
    m = _model(n=(2, 1, 1), alter_mode=alter, eps_dot_o_y=2.5e-3)
    base = smooth_displacement(m, strain=3.0e-3, seed=4)
    o = op.Oracle(m)
    z = np.zeros(m.nodof)
    st = new_state(m)
    f0, _ = o.eps_sig_ifv(st, 1, 1, z, base)
    K = assemble(m, o.tanstf(st, 1, 1, z, base))
    assert st.urcs_n1.reshape(-1, 9)[:, 8].min() > 1e-4       # every point plastic
    rng = np.random.default_rng(2)
    d = rng.standard_normal(m.nodof)
    h = 1e-6 * np.abs(base).max()
    st2 = new_state(m)
    f1, _ = o.eps_sig_ifv(st2, 1, 1, z, base + h * d)
    err = np.linalg.norm((f1 - f0) - K @ (h * d)) / np.linalg.norm(f1 - f0)
    assert err < bound, err


def test_rate_dependence_and_hardening():
    """same strain, shorter time step -> higher flow stress (rate exponent n = 20); the Voce
    hardening variable grows from tau_y towards tau_y + tau_v."""
    H = np.diag([3.0e-3, -1.35e-3, -1.35e-3])
    out = []
    for dt in (1.0, 1.0e-3):
        m = _model(dt=dt, angles=np.array([[[20.0, 40.0, 60.0]]]))
        o, st = op.Oracle(m), new_state(m)
        o.eps_sig_ifv(st, 1, 1, np.zeros(m.nodof), _uniform(m, H))
        hs = st.history1.reshape(8, -1)
        out.append((st.urcs_n1.reshape(-1, 9)[0, 0], hs[0, 87 + 30 + 12]))
    (s_slow, tt_slow), (s_fast, tt_fast) = out
    assert s_fast == pytest.approx(s_slow, rel=1e-12)     # mm10 Voce: dg = eff. strain incr, rate-independent form
    assert CT["tau_y"] < tt_slow < CT["tau_y"] + CT["tau_v"]


def test_rigid_rotation_leaves_the_crystal_stress_free():
    m = _model(n=(2, 2, 1), geonl=True)
    a = np.array([0.3, -0.5, 0.8]); a /= np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(0.4) * K + (1 - np.cos(0.4)) * K @ K
    x = m.c.reshape(-1, 3)
    du = (x @ R.T - x).reshape(-1)
    o, st = op.Oracle(m), new_state(m)
    f, sc = o.eps_sig_ifv(st, 1, 1, np.zeros(m.nodof), du)   # step 1: history initialised
    assert not sc["material_cut_step"]
    assert np.abs(st.urcs_n1.reshape(-1, 9)[:, :6]).max() < 1e-8
    H = int(m.hist_size[0])
    rot = st.history1.reshape(-1, H)[:, 63:72].reshape(-1, 3, 3).transpose(0, 2, 1)
    assert np.abs(rot - R).max() < 1e-12
