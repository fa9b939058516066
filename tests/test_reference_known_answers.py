"""Known answers published with the reference, reproduced with the CPU restatement (oracle/):
the only place where results of this repository can be laid next to numbers the Fortran itself
produced (it cannot be built here).  Five of them are within reach of the path and cover every
material model it builds: simple_voce and verification/test82/voche_model (mm10, small strain),
verification/test24 (mm01, finite strain) and verification/test60 a and b (mm05, both options,
finite strain, 420 cyclic steps).  All but simple_voce (48 slip systems) also run through the CUDA
engine (gpu-marked tests at the end).

example_problems_cp/simple_voce (rhdodds/warp3d): a unit cube stretched in z for 5 h and then held
for 75 h while the stress relaxes; crystal plasticity with Voce hardening, bcc48 slip, 25 crystal
orientations per integration point (Taylor), alter_mode on.  The deck says every integration point
of every element has the same solution, so one element with the deck's boundary conditions carries
it.  results_for_plotting gives the axial stress at 79 times with 9 digits.  Data (angles, curve)
are in tests/golden/simple_voce_reference.json; nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from newton_driver import assemble
from oracle import oracle_py as op
from warp3d_b200 import Material, attach_crystals, build_structured_hex8, make_crystal_type, new_state

HERE = os.path.dirname(os.path.abspath(__file__))


class OracleBackend:
    """The CPU restatement behind the four calls the drivers below make."""

    def __init__(self, m):
        self.m, self.o, self.st = m, op.Oracle(m, nthreads=4), new_state(m)

    def sync(self):
        self.o.sync_globals()

    def eps_sig_ifv(self, step, it, u, du):
        return self.o.eps_sig_ifv(self.st, step, it, u, du)

    def tanstf(self, step, it, u, du):
        return self.o.tanstf(self.st, step, it, u, du)

    def update(self):
        self.o.update(self.st); self.o.adaptive_save(self.st)

    def reset(self):
        self.o.adaptive_reset(self.st)

    def state(self):
        return self.st

    def set_loads(self, cstmap, pbar):
        self.cst, self.pbar = cstmap, pbar

    def residual(self, it, ifv, du):
        """upres + the cvtest residual loop (upres.f:121, cvtest.f:62-78)"""
        res, n = op.upres(self.cst, self.pbar, self.pbar, ifv, du)
        n.update(op.cvtest_residual(res))
        return res, n


class EngineBackend:
    """The CUDA engine through the C ABI: the same drivers, the product instead of the oracle."""

    def __init__(self, m):
        from warp3d_b200.engine import Engine
        self.m, self.e = m, Engine(m)

    def sync(self):
        self.e.set_time_step(self.m.dt)

    def eps_sig_ifv(self, step, it, u, du):
        return self.e.drive_eps_sig_internal_forces(step, it, u, du)

    def tanstf(self, step, it, u, du):
        return self.e.tanstf(step, it, u, du)

    def update(self):
        self.e.update(); self.e.adaptive_save()

    def reset(self):
        self.e.adaptive_reset()

    def state(self):
        return self.e.pull_state(new_state(self.m))

    def set_loads(self, cstmap, pbar):
        self.e.residual_set_loads(cstmap, pbar, load=pbar)

    def residual(self, it, ifv, du):
        """the same on the device, from the internal forces the stress pass left there"""
        return self.e.residual(it)


def _run_simple_voce(nsteps, nel=3, out_elem=21, backend=None):
    ref = json.load(open(os.path.join(HERE, "golden", "simple_voce_reference.json")))
    m = build_structured_hex8(nel, nel, nel, materials=[Material(mat_type=10)], geonl=False, bbar=True,
                              block_size=4)                   # "blocking automatic size = 4"
    # the deck's coordinates: multiples of 1/3 as printed single-precision numbers
    grid = {1: [0.0, 1.0], 3: [0.0, 0.333333343, 0.666666687, 1.0]}[nel]
    idx = np.rint(m.c.reshape(-1, 3)).astype(int)
    m.c[:] = np.asarray(grid)[idx].reshape(-1)
    ct = make_crystal_type(e=200000.0, nu=0.3, slip="bcc48", harden_n=5.0, tau_y=205.0, tau_v=5000.0,
                           voche_m=1.0, theta_o=0.01, eps_dot_o_y=4.0e-05, alter_mode=True)
    ang = np.broadcast_to(np.array(ref["angles_deg"]), (m.noelem, 25, 3))
    attach_crystals(m, ct, ncrystals=25, angles=ang, angle_type=1)
    m.extrapolated_du = 0                                   # "extrapolate off"
    x = m.c.reshape(-1, 3)
    fixed, top = [], []
    for node in range(m.nonode):
        for k in range(3):
            if x[node, k] == 0.0:
                fixed.append(3 * node + k)                  # symmetry planes x = y = z = 0
        if x[node, 2] == 1.0:
            top.append(3 * node + 2)                        # plane z = 1: w imposed
    fixed, top = np.array(fixed), np.array(top)
    cons = np.concatenate([fixed, top])
    free = np.setdiff1d(np.arange(m.nodof), cons)
    be = (backend or OracleBackend)(m)
    u = np.zeros(m.nodof)
    stress, times, t, spread = [], [], 0.0, []
    for step in range(1, nsteps + 1):
        dt = 0.5 if step <= 10 else 0.1 if step <= 30 else 0.5 if step <= 40 else 1.0 if step <= 50 else 2.0
        dw = 0.0003 if step <= 10 else 1.0e-15
        m.dt = dt
        be.sync()
        du = np.zeros(m.nodof)
        ifv, _ = be.eps_sig_ifv(step, 0, u, du)
        K = assemble(m, be.tanstf(step, 0, u, du))
        for it in range(1, 12):
            delta = np.zeros(m.nodof)
            if it == 1:
                delta[top] = dw
            rhs = (-ifv)[free] - K[free][:, cons] @ delta[cons]
            delta[free] = spla.spsolve(K[free][:, free].tocsc(), rhs)
            du = du + delta
            ifv, sc = be.eps_sig_ifv(step, it, u, du)
            assert not sc["material_cut_step"], (step, it)
            rn = np.linalg.norm(ifv[free]); rf = np.linalg.norm(ifv[cons]) or 1.0
            if it >= 2 and rn <= 1.0e-9 * rf:              # the deck: at least 2 iterations
                break
            K = assemble(m, be.tanstf(step, it, u, du))
        u = u + du
        be.update()
        t += dt
        s = be.state().urcs_n.reshape(-1, 8, 9)             # (element, Gauss point, component)
        spread.append(float(np.ptp(s[:, :, 2]) / abs(s[:, :, 2]).max()))
        # "output stresses 21" with the elements' `center` option: the element value is the
        # average over its integration points
        e = min(out_elem, m.noelem) - 1
        stress.append(float(s[e, :, 2].mean())); times.append(t)
    return np.array(times), np.array(stress), ref, max(spread)


def test_simple_voce_stretch_and_relaxation_curve():
    nsteps = 79
    times, stress, ref, spread = _run_simple_voce(nsteps)
    tr, sr = np.array(ref["time_h"][1:nsteps + 1]), np.array(ref["stress_cp"][1:nsteps + 1])
    assert np.allclose(times, tr, rtol=0, atol=1e-9)
    rel = np.abs(stress - sr) / sr
    # Measured: 1.1e-5 at the first step, 4e-6 at peak stress (523.26 MPa), 1e-7 at 80 h
    # (264.364496 here, 264.364523 published).  The reference stops its Newton loops at relative
    # tolerances of 5e-5 (crystal) and 1e-4 (structure) and its own regression tests accept 1e-4.
    # The whole difference is a constant -6.7e-4 MPa that enters in step 1 and relaxes away: the
    # published first point is exactly elastic (E x 0.0003 stored as real*4 = 60.000003) while
    # backward Euler at 60 MPa relaxes 6.7e-4 MPa in 0.5 h -- the same rate law reproduces the
    # published relaxation of step 2 (0.021526 MPa) to six digits.
    assert rel.max() < 2.0e-5, (rel.max(), int(rel.argmax()), stress[int(rel.argmax())], sr[int(rel.argmax())])
    assert rel[9:].max() < 6.0e-6 and rel[-1] < 3.0e-7
    assert 0.005 < spread < 0.1          # not homogeneous: element 21 of the 3 x 3 x 3 mesh matters


def test_simple_voce_one_element_is_not_the_same_problem():
    """The deck's remark that all points share one solution holds only approximately (25 crystals
    are not an isotropic aggregate): a single element reproduces the elastic range but drifts to
    1e-3 once the crystals yield -- the 27-element mesh above is needed for 1e-5."""
    times, stress, ref, _ = _run_simple_voce(12, nel=1, out_elem=1)
    sr = np.array(ref["stress_cp"][1:13])
    rel = np.abs(stress - sr) / sr
    assert rel[:3].max() < 2.0e-5 and 5.0e-4 < rel[9:].max() < 3.0e-3


# ---- verification/test24: hollow sphere sector, mm01 bilinear, geonl + B-bar, skew constraints -----
def _run_test24(nsteps=21, backend=OracleBackend):
    ref = json.load(open(os.path.join(HERE, "golden", "test24_reference.json")))
    mat = Material(e=30000.0, nu=0.3, yld_pt=60.0, tan_e=500.0, beta=1.0)
    from warp3d_b200 import build_hex8_model
    m = build_hex8_model(ref["coords"], ref["incid"], materials=[mat], geonl=True, bbar=True)   # 45 elements
    m.extrapolated_du = 0
    n = m.nodof
    x = m.c.reshape(-1, 3)
    # node frames: u_local = T u_global (rows of T = local axes); fixed local dofs
    Q = np.eye(n)
    fixed = set()
    for g in ref["skew"]:
        T = np.array(g["T"])
        for node in g["nodes"]:
            s = 3 * (node - 1)
            Q[s:s + 3, s:s + 3] = T
            fixed.update(s + k for k in g["fixed"])
    for node in range(m.nonode):
        if x[node, 2] == 0.0:
            fixed.add(3 * node + 2)        # plane z = 0 symmetry
        if x[node, 1] == 0.0:
            fixed.add(3 * node + 1)        # plane y = 0 symmetry
    fixed = np.array(sorted(fixed))
    free = np.setdiff1d(np.arange(n), fixed)
    P1 = np.zeros(n)
    for node, f in ref["loads"].items():
        P1[3 * (int(node) - 1):3 * int(node)] = f
    be = backend(m)
    u = np.zeros(n)
    factor = 0.0
    for step in range(1, nsteps + 1):
        factor += 1.0 if step == 1 else 0.2
        P = factor * P1
        du = np.zeros(n)
        ifv, _ = be.eps_sig_ifv(step, 0, u, du)
        K = assemble(m, be.tanstf(step, 0, u, du)).toarray()
        for it in range(1, 40):
            r = Q @ (P - ifv)                                   # residual in the node frames
            Kl = Q @ K @ Q.T
            dl = np.zeros(n)
            dl[free] = np.linalg.solve(Kl[np.ix_(free, free)], r[free])
            du = du + Q.T @ dl
            ifv, sc = be.eps_sig_ifv(step, it, u, du)
            assert not sc["material_cut_step"], (step, it)
            rn = np.linalg.norm((Q @ (P - ifv))[free])
            # (the reference's polar decomposition leaves ~1e-9 of noise in the forces: DESIGN.md 4)
            if rn <= 1.0e-7 * np.linalg.norm(P):
                break
            K = assemble(m, be.tanstf(step, it, u, du)).toarray()
        else:
            raise AssertionError(f"step {step} did not converge: {rn}")
        u = u + du
        be.update()
    # "output stresses 1": Cauchy stress = R t R^T at each point, element-centre value = mean
    st = be.state()
    t = st.urcs_n.reshape(-1, 8, 9)[0]
    R = st.rot_n.reshape(-1, 8, 9)[0]
    sxx = []
    for gp in range(8):
        Rm = R[gp].reshape(3, 3).T                              # stored column-major
        tm = np.array([[t[gp, 0], t[gp, 3], t[gp, 5]], [t[gp, 3], t[gp, 1], t[gp, 4]], [t[gp, 5], t[gp, 4], t[gp, 2]]])
        sxx.append((Rm @ tm @ Rm.T)[0, 0])
    return float(np.mean(sxx)), ref, u, m


def test_verification_test24_sphere_with_skew_constraints():
    """mm01 (bilinear, isotropic hardening) with the finite-strain kinematics and B-bar: the
    reference's known answer for verification/test24, sigma_xx of element 1 after 21 load steps
    of internal pressure on a hollow sphere sector (the inner radius grows by half)."""
    sxx, ref, u, m = _run_test24()
    want = ref["answer"]["value"]
    assert abs(sxx - want) <= 1.0e-4 * abs(want), (sxx, want)


# ---- verification/test60: cyclic loading of a notched bar, mm05 (both options), geonl + B-bar -------
def _run_test60(option, nsteps=420, backend=OracleBackend):
    import scipy.sparse as sp
    from warp3d_b200 import build_hex8_model, cyclic_fa_material, cyclic_gp_material
    ref = json.load(open(os.path.join(HERE, "golden", "test60_reference.json")))
    if option == "a":     # nonlinear_hardening
        mat = cyclic_fa_material(e=30000.0, nu=0.3, yld_pt=60.0, q_u=0.0, b_u=0.0, h_u=30000.0, gamma_u=735.0,
                                 sig_tol=0.001)
    else:                 # generalized_plasticity
        mat = cyclic_gp_material(e=30000.0, nu=0.3, yld_pt=60.0, gp_h_u=400.0, gp_tau=0.25, gp_beta_u=100.0,
                                 gp_delta_u=800.0, sig_tol=0.001)
    m = build_hex8_model(ref["coords"], ref["incid"], materials=[mat], geonl=True, bbar=True)
    m.extrapolated_du = 0
    m.adaptive_flag = True                                       # "adaptive on"
    n = m.nodof
    T = np.array(ref["skew"]["T"])
    blocks = {node: T for node in ref["skew"]["nodes"]}
    Q = sp.csr_matrix(sp.block_diag([sp.coo_matrix(blocks.get(node + 1, np.eye(3))) for node in range(m.nonode)]))
    fixed_val = {}
    for node, dof, val in ref["constraints"]:
        fixed_val[3 * (node - 1) + dof] = val
    fixed = np.array(sorted(fixed_val))
    dfix = np.array([fixed_val[k] for k in fixed])
    free = np.setdiff1d(np.arange(n), fixed)
    mult = np.zeros(nsteps + 1)
    for a, b, f in ref["steps"]:
        mult[a:min(b, nsteps) + 1] = f
    ry = np.array([3 * (node - 1) + 1 for node in ref["reaction_nodes"]])
    be = backend(m)
    u = np.zeros(n)
    total = None
    QT = Q.T.tocsr()

    def increment(step, u, frac, depth=0):
        """One (sub-)increment of the step's imposed displacements; halves itself when a material
        point asks for a step cut, as the reference's adaptive solution does.  Returns (du, ifv)."""
        du = np.zeros(n)
        ifv, _ = be.eps_sig_ifv(step, 0, u, du)
        K = assemble(m, be.tanstf(step, 0, u, du))
        for it in range(1, 40):
            r = -(Q @ ifv)
            Kl = (Q @ K @ QT).tocsr()
            dl = np.zeros(n)
            if it == 1:
                dl[fixed] = frac * mult[step] * dfix
            rhs = r[free] - Kl[free][:, fixed] @ dl[fixed]
            dl[free] = spla.spsolve(Kl[free][:, free].tocsc(), rhs)
            du = du + QT @ dl
            ifv, sc = be.eps_sig_ifv(step, it, u, du)
            if sc["material_cut_step"]:
                assert depth < 4, (step, it, depth)
                be.reset()
                du1, _ = increment(step, u, 0.5 * frac, depth + 1)
                be.update()
                du2, ifv = increment(step, u + du1, 0.5 * frac, depth + 1)
                return du1 + du2, ifv
            fl = Q @ ifv
            rn = np.linalg.norm(fl[free]); rf = np.linalg.norm(fl[fixed]) or 1.0
            # the reference's own test: norm res tol 0.0001 (percent); the generalized-plasticity
            # update carries ~1e-8 of self-noise, 1e-7 is not always reachable
            if rn <= 1.0e-6 * rf or (it >= 12 and rn <= 1.0e-5 * rf):
                return du, ifv
            K = assemble(m, be.tanstf(step, it, u, du))
        raise AssertionError(f"step {step} did not converge: {rn / rf}")

    for step in range(1, nsteps + 1):
        du, ifv = increment(step, u, 1.0)
        u = u + du
        be.update()
        total = float(ifv[ry].sum())
    return total, ref


@pytest.mark.parametrize("option", ["a", "b"])
def test_verification_test60_cyclic_notched_bar(option):
    """mm05 with the finite-strain kinematics and B-bar over 420 load steps (2 1/2 cycles between
    unsymmetric end displacements): the reference's known answers for verification/test60, the total
    axial reaction on the loaded end.  a = nonlinear_hardening, b = generalized_plasticity.
    Measured: a 0.0210364867 against 0.021036833 (1.6e-5), b 0.0232109041 against 0.023211111
    (8.9e-6); the reference accepts 1e-4 and converges its steps to 1e-6 of the force norm."""
    total, ref = _run_test60(option)
    want = ref["answers"][option]["value"]
    assert abs(total - want) <= 5.0e-5 * abs(want), (option, total, want)


# ---- verification/test82/voche_model: cube with a void, mm10 (bcc, 12 systems), small strain ---------
def _run_test82_voche(nsteps=8, backend=OracleBackend):
    from warp3d_b200 import build_hex8_model
    ref = json.load(open(os.path.join(HERE, "golden", "test82_voche_reference.json")))
    m = build_hex8_model(ref["coords"], ref["incid"], materials=[Material(mat_type=10)], geonl=False, bbar=True,
                         block_size=20, dt=1.0e6)                      # "blocking automatic size = 20"
    ct = make_crystal_type(e=150000.0, nu=0.28, slip="bcc", harden_n=20.0, voche_m=5.0, theta_o=400.0,
                           tau_y=300.0, tau_v=1200.0)
    ang = np.broadcast_to(np.array(ref["angles_deg"]), (m.noelem, 1, 3))
    attach_crystals(m, ct, ncrystals=1, angles=ang, angle_type=1)
    m.extrapolated_du = 0
    n = m.nodof
    x = m.c.reshape(-1, 3)
    inc = np.array(ref["incid"])
    # equivalent nodal loads of the face pressures (elem_load_b.f hex_face_nloads): the faces are
    # plane squares, so each of the 4 face nodes takes p A / 4 along the inward normal
    P1 = np.zeros(n)
    for elem, face, p in ref["face_pressures"]:
        nodes = inc[elem - 1][np.array(ref["hex8_face_nodes"][face - 1]) - 1] - 1
        xf = x[nodes]
        nrm = np.cross(xf[1] - xf[0], xf[3] - xf[0])
        area = np.linalg.norm(nrm)
        nrm = nrm / area
        if np.dot(nrm, xf.mean(axis=0) - x[inc[elem - 1] - 1].mean(axis=0)) < 0:
            nrm = -nrm                                                 # outward
        for nd in nodes:
            P1[3 * nd:3 * nd + 3] += -p * area / 4.0 * nrm             # positive pressure pushes inward
    fixed = np.array(sorted({3 * nd + k for nd in range(m.nonode) for k in range(3) if x[nd, k] == 0.0}))
    free = np.setdiff1d(np.arange(n), fixed)
    cst = np.zeros(n, dtype=np.int32)
    cst[fixed] = 1
    be = backend(m)
    be.sync()
    u = np.zeros(n)
    for step in range(1, nsteps + 1):
        P = ref["load_factor_per_step"] * step * P1
        be.set_loads(cst, P)
        du = np.zeros(n)
        ifv, _ = be.eps_sig_ifv(step, 0, u, du)
        K = assemble(m, be.tanstf(step, 0, u, du))
        r = np.where(cst == 0, P - ifv, 0.0)
        for it in range(1, 40):
            delta = np.zeros(n)
            delta[free] = spla.spsolve(K[free][:, free].tocsc(), r[free])
            du = du + delta
            ifv, sc = be.eps_sig_ifv(step, it, u, du)
            assert not sc["material_cut_step"], (step, it)
            r, norms = be.residual(it, ifv, du)                        # upres + cvtest norms
            rn = norms["norm_res"]
            assert abs(rn - np.linalg.norm((P - ifv)[free])) <= 1e-12 * np.linalg.norm(P)
            if it >= 2 and rn <= 1.0e-9 * np.linalg.norm(P):
                break
            K = assemble(m, be.tanstf(step, it, u, du))
        else:
            raise AssertionError(f"step {step} did not converge: {rn}")
        u = u + du
        be.update()
    s = be.state().urcs_n.reshape(-1, 8, 9)
    return float(s[63, :, 0].mean()), ref                              # element 64, centre value


def test_verification_test82_voche_model():
    """mm10 with the 12-system bcc table (the size the engine builds) and the non-trivial Voce
    exponent: the reference's known answer for verification/test82/voche_model."""
    sxx, ref = _run_test82_voche()
    want = ref["answer"]["value"]
    assert abs(sxx - want) <= 1.0e-4 * abs(want), (sxx, want)


# ---- the same known answers through the CUDA engine (the product, not its checker) ------------------
@pytest.mark.gpu
def test_engine_reproduces_the_simple_voce_curve():
    """25 crystals per point with the 48-system bcc table (k_mm10<48>): the whole published curve."""
    nsteps = 79
    times, stress, ref, spread = _run_simple_voce(nsteps, backend=EngineBackend)
    sr = np.array(ref["stress_cp"][1:nsteps + 1])
    rel = np.abs(stress - sr) / sr
    assert rel.max() < 2.0e-5 and rel[9:].max() < 6.0e-6 and rel[-1] < 3.0e-7, (rel.max(), rel[-1])


@pytest.mark.gpu
def test_engine_reproduces_verification_test24():
    sxx, ref, u, m = _run_test24(backend=EngineBackend)
    want = ref["answer"]["value"]
    assert abs(sxx - want) <= 1.0e-4 * abs(want), (sxx, want)


@pytest.mark.gpu
@pytest.mark.parametrize("option", ["a", "b"])
def test_engine_reproduces_verification_test60(option):
    total, ref = _run_test60(option, backend=EngineBackend)
    want = ref["answers"][option]["value"]
    assert abs(total - want) <= 5.0e-5 * abs(want), (option, total, want)


@pytest.mark.gpu
def test_engine_reproduces_verification_test82_voche_model():
    sxx, ref = _run_test82_voche(backend=EngineBackend)
    want = ref["answer"]["value"]
    assert abs(sxx - want) <= 1.0e-4 * abs(want), (sxx, want)
