"""CPU checks of the mm05 nonlinear_hardening (Frederick-Armstrong) core through the oracle driver."""
import numpy as np

from common import relerr
from newton_driver import assemble
from oracle import oracle_py as op
from warp3d_b200 import build_structured_hex8, cyclic_fa_material, new_state, smooth_displacement


def test_matprp_rows_follow_inmat():
    m = build_structured_hex8(1, 1, 1, materials=[cyclic_fa_material(q_u=50.0e6, b_u=10.0, h_u=3.0e9, gamma_u=30.0)])
    r = m.matprp[0, 54:60]
    f = np.float32
    assert r[0] == f(50.0e6) and r[3] == f(1.0) and r[5] == f(0.001)
    assert r[1] == f(f(10.0) * np.sqrt(f(2.0) / f(3.0)).astype(f))
    assert r[2] == f(f(3.0e9) * (f(2.0) / f(3.0)))
    assert int(m.hist_size[0]) == 14 and int(m.props[0, 24]) == 5


def test_elastic_below_yield_and_yield_surface_consistency():
    mat = cyclic_fa_material()
    m = build_structured_hex8(2, 2, 1, materials=[mat], geonl=False, coord_jitter=0.1)
    o, st = op.Oracle(m), new_state(m)
    z = np.zeros(m.nodof)
    o.eps_sig_ifv(st, 1, 1, z, smooth_displacement(m, strain=5.0e-4))      # E eps = 100 MPa < 400
    h1 = st.history1.reshape(-1, 14)
    assert np.all((h1[:, 3].view(np.int64) & 0xFFFFFFFF) == 3) and np.all(h1[:, 0] == 0.0)
    st = new_state(m)
    o.eps_sig_ifv(st, 1, 1, z, smooth_displacement(m, strain=4.0e-3))
    h1 = st.history1.reshape(-1, 14)
    s = st.urcs_n1.reshape(-1, 9)[:, :6]
    assert np.all((h1[:, 3].view(np.int64) & 0xFFFFFFFF) == 1)
    # on the updated yield surface: || dev(sigma) - alpha || = k   (to the local-solve tolerance)
    dev = s.copy(); dev[:, :3] -= s[:, :3].mean(axis=1, keepdims=True)
    xi = dev - h1[:, 5:11]
    nrm = np.sqrt((xi[:, :3] ** 2).sum(1) + 2 * (xi[:, 3:] ** 2).sum(1))
    assert np.abs(nrm / h1[:, 1] - 1.0).max() < 2e-3


def test_consistent_tangent_vs_finite_differences():
    m = build_structured_hex8(2, 1, 1, materials=[cyclic_fa_material()], geonl=False, coord_jitter=0.05)
    o = op.Oracle(m)
    z = np.zeros(m.nodof)
    # two converged steps so the state is on the yield surface, then a plastic increment
    st = new_state(m)
    u1 = smooth_displacement(m, strain=3.0e-3, seed=2)
    for it in (0, 1, 2):
        o.eps_sig_ifv(st, 1, it, z, u1)
    o.update(st); o.adaptive_save(st)
    du = smooth_displacement(m, strain=1.5e-3, seed=3)
    s0 = st.copy()
    f0, _ = o.eps_sig_ifv(s0, 2, 2, u1, du)
    K = assemble(m, o.tanstf(s0, 2, 2, u1, du))
    rng = np.random.default_rng(1)
    d = rng.standard_normal(m.nodof)
    h = 1e-6 * np.abs(du).max()
    s1 = st.copy()
    f1, _ = o.eps_sig_ifv(s1, 2, 2, u1, du + h * d)
    err = np.linalg.norm((f1 - f0) - K @ (h * d)) / np.linalg.norm(f1 - f0)
    # cnst5 is the tangent of a single backward-Euler step; the update uses Richardson
    # extrapolation over sub-increments, so agreement is to a few percent, not to round-off
    assert err < 8e-2, err


def test_cyclic_hardening_saturates():
    """uniaxial cycling: isotropic part k grows towards sqrt(2/3)(sigy + q_u)."""
    mat = cyclic_fa_material()
    m = build_structured_hex8(1, 1, 1, materials=[mat], geonl=False)
    o, st = op.Oracle(m), new_state(m)
    x = m.c.reshape(-1, 3)
    u = np.zeros(m.nodof)
    kmax = 0.0
    for step in range(1, 25):
        sign = 1.0 if (step - 1) // 3 % 2 == 0 else -1.0
        du = np.zeros((m.nonode, 3)); du[:, 0] = sign * 2.0e-3 * x[:, 0]; du[:, 1] = -0.4 * sign * 2.0e-3 * x[:, 1]; du[:, 2] = -0.4 * sign * 2.0e-3 * x[:, 2]
        du = du.reshape(-1)
        for it in (0, 1, 2):
            _, sc = o.eps_sig_ifv(st, step, it, u, du)
            assert not sc["material_cut_step"]
        o.update(st); o.adaptive_save(st)
        u = u + du
        kmax = max(kmax, st.history.reshape(-1, 14)[:, 1].max())
    k0 = 0.81649658 * 400.0e6
    ksat = 0.81649658 * (400.0e6 + 100.0e6)
    assert k0 < kmax <= ksat * (1 + 1e-9)
    assert kmax > k0 + 0.2 * (ksat - k0)


# ---- generalized_plasticity option ---------------------------------------------------------------
from warp3d_b200 import cyclic_gp_material  # noqa: E402


def test_gp_matprp_rows_follow_inmat():
    m = build_structured_hex8(1, 1, 1, materials=[cyclic_gp_material(gp_h_u=350.0, gp_tau=0.25, gp_beta_u=180.0,
                                                                      gp_delta_u=0.02, sig_tol=0.0005)])
    r = m.matprp[0, 54:60]
    f = np.float32
    # uniaxial values stored unconverted, option flag -1 (inmat.f:2017-2036, :2272)
    assert list(r) == [f(350.0), f(0.25), f(180.0), f(-1.0), f(0.02), f(0.0005)]
    assert int(m.hist_size[0]) == 14 and int(m.props[0, 24]) == 5


def test_gp_elastic_below_yield_then_smooth_approach_to_the_limit_surface():
    """Generalized plasticity: elastic inside the yield surface k = sqrt(2/3) sigma_0; beyond it the
    stress leaves the yield surface and approaches the limit surface (radius k + sqrt(2/3) beta_u,
    translating with H) with a monotonically decreasing tangent."""
    mat = cyclic_gp_material()
    m = build_structured_hex8(1, 1, 1, materials=[mat], geonl=False)
    o, st = op.Oracle(m), new_state(m)
    x = m.c.reshape(-1, 3)
    u = np.zeros(m.nodof)
    mises, epsp = [], []
    for step in range(1, 16):
        de = 1.0e-3                     # yield strain = 60 / 30000 = 2e-3
        du = np.zeros((m.nonode, 3))
        du[:, 0] = de * x[:, 0]; du[:, 1] = -0.45 * de * x[:, 1]; du[:, 2] = -0.45 * de * x[:, 2]
        du = du.reshape(-1)
        for it in (0, 1, 2):
            _, sc = o.eps_sig_ifv(st, step, it, u, du)
            assert not sc["material_cut_step"], step
        o.update(st); o.adaptive_save(st)
        u = u + du
        h = st.history.reshape(-1, 14)
        s = st.urcs_n.reshape(-1, 9)[:, :6]
        dev = s.copy(); dev[:, :3] -= s[:, :3].mean(axis=1, keepdims=True)
        mises.append(np.sqrt(1.5) * np.sqrt((dev[:, :3] ** 2).sum(1) + 2 * (dev[:, 3:] ** 2).sum(1)).max())
        epsp.append(h[:, 2].max())
        state = h[:, 3].view(np.int64) & 0xFFFFFFFF
        if step == 1:
            assert np.all(state == 3) and np.all(h[:, 0] == 0.0) and epsp[-1] == 0.0
        else:
            assert np.all(state == 1) and np.all(h[:, 0] > 0.0)
            assert 2 <= h[:, 11].max() <= 10                      # sub-increment count, capped at mxsubinc
            # tau = 0: purely kinematic, the yield surface keeps its size
            assert np.allclose(h[:, 1], 0.81649658 * 60.0, rtol=1e-12)
            xi = dev - h[:, 5:11]
            nrm = np.sqrt((xi[:, :3] ** 2).sum(1) + 2 * (xi[:, 3:] ** 2).sum(1))
            assert np.all(nrm > h[:, 1])                          # outside the yield surface ...
            assert np.all(nrm < h[:, 1] + 0.81649658 * 200.0)     # ... inside the limit surface
    mises, epsp = np.array(mises), np.array(epsp)
    assert np.all(np.diff(mises) > 0) and np.all(np.diff(epsp[1:]) > 0)
    slope = np.diff(mises[1:]) / np.diff(epsp[1:])
    assert np.all(np.diff(slope) < 0)                             # smooth, ever softer approach
    assert mises[-1] < 60.0 + 200.0 + 400.0 * np.sqrt(2.0 / 3.0) * epsp[-1]


def test_gp_consistent_tangent_vs_finite_differences():
    """cnst5_gp is the tangent of ONE backward-Euler step of the GP return, the update runs up to 10
    sub-increments with Richardson extrapolation: for a small plastic increment (2 sub-increments)
    the two agree to 1e-3, which checks the transcriptions of update and tangent against each
    other; for a large one (10 sub-increments) only to some 20 %, as in the reference."""
    m = build_structured_hex8(2, 1, 1, materials=[cyclic_gp_material()], geonl=False, coord_jitter=0.05)
    o = op.Oracle(m)
    z = np.zeros(m.nodof)
    st = new_state(m)
    u1 = smooth_displacement(m, strain=4.0e-3, seed=2)
    for it in (0, 1, 2):
        o.eps_sig_ifv(st, 1, it, z, u1)
    o.update(st); o.adaptive_save(st)
    d = np.random.default_rng(1).standard_normal(m.nodof)
    for amp, nsub_max, bound in ((1.5e-4, 2, 5e-3), (1.5e-3, 10, 0.25)):
        du = smooth_displacement(m, strain=amp, seed=3)
        s0 = st.copy()
        f0, sc = o.eps_sig_ifv(s0, 2, 2, u1, du)
        assert not sc["material_cut_step"]
        h1 = s0.history1.reshape(-1, 14)
        assert np.all((h1[:, 3].view(np.int64) & 0xFFFFFFFF) == 1) and h1[:, 11].max() == nsub_max
        K = assemble(m, o.tanstf(s0, 2, 2, u1, du))
        assert np.abs(K - K.T).max() <= 1e-12 * np.abs(K).max()
        hstep = 1e-6 * np.abs(du).max()
        s1 = st.copy()
        f1, _ = o.eps_sig_ifv(s1, 2, 2, u1, du + hstep * d)
        err = np.linalg.norm((f1 - f0) - K @ (hstep * d)) / np.linalg.norm(f1 - f0)
        assert err < bound, (amp, err)
