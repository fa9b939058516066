"""CPU tests that pin the oracle (oracle/) by physics invariants and by the reference's literal
tables.  The reference ships no function-level golden vectors for this path and cannot be built
here, so these are the strongest pins available: table identities, rigid-body
invariance, patch tests, symmetric/consistent tangents, closed-form uniaxial plasticity."""
import numpy as np
import pytest

from common import relerr
from newton_driver import OracleBackend, assemble, icp_table, solve_steps
from oracle import oracle_py as op
from warp3d_b200 import Material, build_structured_hex8, new_state, smooth_displacement
from warp3d_b200.model import HEX8_NODE_SIGNS


def test_quad2_points_and_weights():
    """getgpts.f:642-718: weight 1, |coordinate| = 1/sqrt(3), WARP3D point order."""
    order = [(-1, -1, -1), (-1, 1, -1), (1, -1, -1), (1, 1, -1), (-1, -1, 1), (-1, 1, 1), (1, -1, 1), (1, 1, 1)]
    for gp, sg in enumerate(order, start=1):
        xi, eta, zeta, w = op.quad2(gp)
        assert w == 1.0
        assert np.allclose([xi, eta, zeta], np.array(sg) * 0.5773502691896257645, rtol=0, atol=0)


def test_deriv2_matches_trilinear_shape_functions():
    """derivs.f:283-343: dN/dxi of N_a = (1+xi xi_a)(1+eta eta_a)(1+zeta zeta_a)/8 in WARP3D's
    node order, partition of unity, and d(x)/d(xi) of the parent cube = identity."""
    rng = np.random.default_rng(0)
    for _ in range(5):
        xi, eta, zeta = rng.uniform(-1, 1, 3)
        nxi, neta, nzeta = op.deriv2(xi, eta, zeta)
        s = HEX8_NODE_SIGNS
        assert np.allclose(nxi, s[:, 0] * (1 + eta * s[:, 1]) * (1 + zeta * s[:, 2]) / 8, atol=1e-16)
        assert np.allclose(neta, s[:, 1] * (1 + xi * s[:, 0]) * (1 + zeta * s[:, 2]) / 8, atol=1e-16)
        assert np.allclose(nzeta, s[:, 2] * (1 + xi * s[:, 0]) * (1 + eta * s[:, 1]) / 8, atol=1e-16)
        assert abs(nxi.sum()) < 1e-16 and abs(neta.sum()) < 1e-16 and abs(nzeta.sum()) < 1e-16
        J = np.array([nxi @ s, neta @ s, nzeta @ s], dtype=float)
        assert np.allclose(J, np.eye(3), atol=1e-15)


def test_cp_dcp_icp_packing_is_bit_exact():
    """setup.f:268-276, 407-416: cp(j) = j(j-1)/2, dcp(j) = cp(j)+j, icp(cp(j)+i) = (i, j)."""
    cp, dcp, icp = op.cp_icp(24)
    j = np.arange(1, 25)
    assert np.array_equal(cp, j * (j - 1) // 2)
    assert np.array_equal(dcp, cp + j)
    r, c = icp_table(24)
    assert np.array_equal(icp[0], r + 1) and np.array_equal(icp[1], c + 1)
    # the reference's row rule for assembly: k = dcp(max) - |ecol - erow|  (assemble_code.f:926)
    for erow in (1, 5, 24):
        for ecol in (1, 7, 24):
            k = dcp[max(erow, ecol) - 1] - abs(ecol - erow)
            assert (icp[0, k - 1], icp[1, k - 1]) == (min(erow, ecol), max(erow, ecol))


def _rigid(model, angle=0.3, axis=(1.0, 2.0, -1.0), shift=(0.1, -0.2, 0.05)):
    a = np.array(axis) / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    x = model.c.reshape(-1, 3)
    return ((x @ R.T + np.array(shift)) - x).reshape(-1), R


def test_rigid_body_motion_gives_zero_strain_and_the_rotation():
    m = build_structured_hex8(3, 2, 2, coord_jitter=0.15, seed=2)
    o, st = op.Oracle(m), new_state(m)
    du, R = _rigid(m)
    # step 2: coordinates are updated (geonl), finite rotation in one increment
    f, sc = o.eps_sig_ifv(st, 2, 1, np.zeros(m.nodof), du)
    assert np.abs(st.eps_n1).max() < 1e-13
    assert np.abs(st.urcs_n1.reshape(-1, 9)[:, :6]).max() < 1e-13 * 200e9
    rot = st.rot_n1.reshape(-1, 3, 3).transpose(0, 2, 1)       # stored column-major
    assert np.abs(rot - R).max() < 1e-12
    assert np.abs(f).max() < 1e-3          # forces ~ 1e-13 * E * area


@pytest.mark.parametrize("bbar", [True, False])
def test_small_strain_patch_test_linear_elastic(bbar):
    """Uniform strain on a distorted mesh: every Gauss point carries D*eps exactly and interior
    nodes are in equilibrium (also with B-bar, whose mean-dilatation equals the local one)."""
    mat = Material(linear_elastic=True)
    m = build_structured_hex8(3, 3, 3, materials=[mat], geonl=False, bbar=bbar, coord_jitter=0.2, seed=4)
    H = np.array([[1.0e-3, 2.0e-4, -1.0e-4], [0.0, -5.0e-4, 3.0e-4], [2.0e-4, 0.0, 7.0e-4]])
    x = m.c.reshape(-1, 3)
    du = (x @ H.T).reshape(-1)
    o, st = op.Oracle(m), new_state(m)
    f, _ = o.eps_sig_ifv(st, 1, 1, np.zeros(m.nodof), du)
    E, nu = float(np.float32(mat.e)), float(np.float32(mat.nu))
    eps = np.array([H[0, 0], H[1, 1], H[2, 2], H[0, 1] + H[1, 0], H[1, 2] + H[2, 1], H[0, 2] + H[2, 0]])
    lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    sig = np.r_[2 * G * eps[:3] + lam * eps[:3].sum(), G * eps[3:]]
    assert relerr(st.eps_n1.reshape(-1, 6), np.tile(eps, (m.noelem * 8, 1))) < 1e-12
    assert relerr(st.urcs_n1.reshape(-1, 9)[:, :6], np.tile(sig, (m.noelem * 8, 1))) < 1e-12
    xi = x.round(12)
    interior = np.all((x > x.min(0) + 1e-9) & (x < x.max(0) - 1e-9), axis=1)
    assert np.abs(f.reshape(-1, 3)[interior]).max() < 1e-9 * np.abs(f).max()


def test_elastic_stiffness_is_symmetric_psd_with_six_rigid_modes_and_ku_equals_f():
    mat = Material(linear_elastic=True)
    m = build_structured_hex8(2, 2, 2, materials=[mat], geonl=False, bbar=True, coord_jitter=0.15, seed=6)
    o, st = op.Oracle(m), new_state(m)
    rng = np.random.default_rng(1)
    du = rng.standard_normal(m.nodof) * 1e-4
    f, _ = o.eps_sig_ifv(st, 1, 1, np.zeros(m.nodof), du)
    K = assemble(m, o.tanstf(st, 1, 1, np.zeros(m.nodof), du)).toarray()
    assert np.abs(K - K.T).max() <= 1e-15 * np.abs(K).max()   # duplicate sums differ in order only
    w = np.linalg.eigvalsh(K)
    assert w[0] > -1e-6 * w[-1] and (np.abs(w) < 1e-9 * w[-1]).sum() == 6
    assert relerr(K @ du, f) < 1e-11


def test_consistent_tangent_matches_finite_differences_small_strain_plastic():
    """cnst1 is the derivative of the mm01 return map: K du ~ f(u+du) - f(u) to O(du^2)."""
    m = build_structured_hex8(2, 2, 1, geonl=False, bbar=True, coord_jitter=0.1, seed=8)
    o = op.Oracle(m)
    base = smooth_displacement(m, strain=4.0e-3, seed=3)
    st = new_state(m)
    z = np.zeros(m.nodof)
    f0, _ = o.eps_sig_ifv(st, 1, 1, z, base)
    K = assemble(m, o.tanstf(st, 1, 1, z, base))
    assert (st.history1.reshape(-1, 11)[:, 0] > 0).mean() > 0.9      # plastic everywhere
    rng = np.random.default_rng(5)
    d = rng.standard_normal(m.nodof)
    errs = []
    for h in (1e-6, 1e-7):
        st2 = new_state(m)
        f1, _ = o.eps_sig_ifv(st2, 1, 1, z, base + h * base.max() * d)
        errs.append(np.linalg.norm((f1 - f0) - K @ (h * base.max() * d)) / np.linalg.norm(f1 - f0))
    assert errs[0] < 1e-4 and errs[1] < errs[0]


def test_uniaxial_bilinear_closed_form_and_quadratic_newton_convergence():
    """Displacement-controlled bar, small strain: sigma = sigy + Et (eps - epsy) past yield;
    lateral strain from plastic incompressibility; Newton residuals contract quadratically."""
    mat = Material(e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=20.0e9, beta=1.0)
    m = build_structured_hex8(3, 2, 2, materials=[mat], geonl=False, bbar=True)
    x = m.c.reshape(-1, 3)
    fixed = []
    for n in range(m.nonode):
        if x[n, 0] == 0.0: fixed.append(3 * n)
        if x[n, 0] == 3.0: fixed.append(3 * n)
        if x[n, 1] == 0.0: fixed.append(3 * n + 1)
        if x[n, 2] == 0.0: fixed.append(3 * n + 2)
    fixed = np.array(sorted(set(fixed)))
    eps_steps = [1.0e-3, 2.0e-3, 2.0e-3]           # total 5e-3, yield strain 2e-3
    incs = []
    for de in eps_steps:
        v = np.where((fixed % 3 == 0) & (x[fixed // 3, 0] == 3.0), 3.0 * de, 0.0)
        incs.append(v)
    st = new_state(m)
    u, log = solve_steps(m, OracleBackend(m, st), fixed, incs, nsteps=3, tol=1e-12)
    E, Et, sy = float(np.float32(mat.e)), float(np.float32(mat.tan_e)), float(np.float32(mat.yld_pt))
    eps = 5.0e-3
    sig = sy + Et * (eps - sy / E)
    s = st.urcs_n.reshape(-1, 9)
    assert relerr(s[:, 0], np.full(len(s), sig)) < 1e-6      # props are real*4 -> H' rounding
    assert np.abs(s[:, 1:6]).max() < 1e-6 * sig
    assert relerr(st.eps_n.reshape(-1, 6)[:, 0], np.full(len(s), eps)) < 1e-10
    for norms in log:
        assert len(norms) <= 6
    # quadratic contraction in the plastic steps: r_{k+1} <~ C r_k^2 (relative)
    r = np.array(log[1])[:, 0]
    assert r[-1] < 1e-9 * r[0] or len(r) <= 3
