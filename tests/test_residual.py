"""Structure residual and its norms (upres.f, cvtest.f:62-78): the sequential CPU restatement
against vectorised numpy (CPU), and the device kernels against the restatement (GPU)."""
import numpy as np
import pytest

from common import load_path
from oracle import oracle_py as op
from warp3d_b200 import build_structured_hex8, new_state


def _vectors(n, seed, dynamic, mpc, contact):
    rng = np.random.default_rng(seed)
    v = dict(cstmap=(rng.random(n) < 0.15).astype(np.int32), pbar=rng.normal(size=n) * 50.0,
             load=np.where(rng.random(n) < 0.3, rng.normal(size=n) * 40.0, 0.0),
             ifv=rng.normal(size=n) * 50.0, du=rng.normal(size=n) * 1e-3,
             mdiag=None, velocity=None, accel=None, contact_force=None, d_lagrange_forces=None)
    if dynamic:
        v.update(mdiag=rng.random(n) * 2.0e-4, velocity=rng.normal(size=n), accel=rng.normal(size=n) * 10.0)
    if contact:
        v["contact_force"] = np.where(rng.random(n) < 0.1, rng.normal(size=n) * 5.0, 0.0)
    if mpc:
        v["d_lagrange_forces"] = np.where(rng.random(n) < 0.1, rng.normal(size=n) * 5.0, 0.0)
    return v


def _z(v, n):
    return np.zeros(n) if v is None else v


def _numpy_upres(v, dt, nbeta):
    """upres.f:147-226 + cvtest.f:62-78 with numpy array expressions (same operation order per dof)."""
    n = v["pbar"].size
    md, cf, vel, acc = (_z(v[k], n) for k in ("mdiag", "contact_force", "velocity", "accel"))
    nfac, nfac1, nfac2 = 1.0 / (nbeta * dt * dt), (1.0 - 2.0 * nbeta) / (2.0 * nbeta), 1.0 / (nbeta * dt)
    ext = v["pbar"] + cf
    if v["d_lagrange_forces"] is not None:
        ext = ext + v["d_lagrange_forces"]
    res = (ext - md * v["du"] * nfac) - v["ifv"]
    a1 = v["du"] * nfac - nfac1 * acc - nfac2 * vel
    ic = abs(a1) * md + abs(cf)
    free = v["cstmap"] == 0
    term = np.where(free, ic + abs(v["load"]), ic + abs(v["ifv"]))
    res = np.where(free, res, 0.0)
    rf = res if v["d_lagrange_forces"] is None else np.where(v["d_lagrange_forces"] != 0.0, 0.0, res)
    dof = int(np.argmax(abs(rf))) + 1           # argmax returns the first of equals
    return res, dict(mgload=np.sqrt(np.sum(term * term)), sum_loads=np.sum(abs(term)),
                     num_term_loads=int(np.count_nonzero(term)), norm_res=np.sqrt(np.sum(rf * rf)),
                     res_max=res[dof - 1], res_max_dof=dof)


def _oracle_upres(v, dt, nbeta):
    res, n1 = op.upres(v["cstmap"], v["pbar"], v["load"], v["ifv"], v["du"], dt, nbeta, v["mdiag"],
                       v["velocity"], v["accel"], v["contact_force"], v["d_lagrange_forces"])
    n1.update(op.cvtest_residual(res, v["d_lagrange_forces"]))
    return res, n1


def _close(a, b, rtol=1e-13):
    return abs(a - b) <= rtol * max(abs(a), abs(b), 1e-300)


@pytest.mark.parametrize("dynamic,mpc,contact", [(False, False, False), (True, False, True), (True, True, True)])
def test_restatement_against_numpy(dynamic, mpc, contact):
    v = _vectors(3001, 7, dynamic, mpc, contact)
    dt, nbeta = 0.02, 0.25
    res, n1 = _oracle_upres(v, dt, nbeta)
    res_np, n_np = _numpy_upres(v, dt, nbeta)
    assert np.array_equal(res, res_np)                       # element-wise: same operations, same bits
    for k in ("mgload", "sum_loads", "norm_res"):
        assert _close(n1[k], n_np[k]), k                     # sequential vs pairwise sums
    assert n1["num_term_loads"] == n_np["num_term_loads"] and n1["res_max_dof"] == n_np["res_max_dof"]
    assert n1["res_max"] == n_np["res_max"]
    # iteration 0 (upres.f:57-72): constrained dofs zero, norm of the rest
    res0, norm0 = op.upres_iter_0(v["cstmap"], v["pbar"], v["ifv"], v["du"], dt, nbeta, v["mdiag"], v["contact_force"])
    md, cf = _z(v["mdiag"], res0.size), _z(v["contact_force"], res0.size)
    ref0 = np.where(v["cstmap"] == 0, (v["pbar"] - md * v["du"] * (1.0 / (nbeta * dt * dt)) + cf) - v["ifv"], 0.0)
    assert np.array_equal(res0, ref0) and _close(norm0, np.sqrt(np.sum(ref0 * ref0)))


def test_restatement_edge_cases():
    n = 12
    z = np.zeros(n)
    cst = np.zeros(n, dtype=np.int32)
    # everything zero: the strict '>' scan starting from -100 stops at dof 1 (cvtest.f:62-73)
    res, n1 = op.upres(cst, z, z, z, z)
    assert not res.any() and n1 == dict(mgload=0.0, sum_loads=0.0, num_term_loads=0)
    assert op.cvtest_residual(res) == dict(norm_res=0.0, res_max=0.0, res_max_dof=1)
    # equal magnitudes: the lowest dof wins, the value keeps its sign
    r = z.copy(); r[[3, 7]] = [-2.5, 2.5]
    assert op.cvtest_residual(r) == dict(norm_res=np.sqrt(12.5), res_max=-2.5, res_max_dof=4)
    # a dof carrying a multipoint-constraint force is left out of the norms (cvtest.f:66-68)
    dl = z.copy(); dl[3] = 1.0
    assert op.cvtest_residual(r, dl) == dict(norm_res=2.5, res_max=2.5, res_max_dof=8)
    # reactions enter the load measure at constrained dofs (upres.f:210-213)
    cst[5] = 1
    f = z.copy(); f[5] = -4.0
    res, n1 = op.upres(cst, z, z, f, z)
    assert res[5] == 0.0 and n1 == dict(mgload=4.0, sum_loads=4.0, num_term_loads=1)


def _engine_case(seed=3):
    m = build_structured_hex8(5, 4, 3, coord_jitter=0.1, seed=seed, block_size=9)
    (step, iters), = list(load_path(m, nsteps=1, strain=4.0e-3))
    return m, step, iters


@pytest.mark.gpu
@pytest.mark.parametrize("dynamic,mpc,contact", [(False, False, False), (True, False, True), (True, True, True)])
def test_device_residual_matches_the_restatement(dynamic, mpc, contact):
    from warp3d_b200.engine import Engine
    m, step, iters = _engine_case()
    e = Engine(m)
    v = _vectors(m.nodof, 21, dynamic, mpc, contact)
    dt, nbeta = 0.05, 0.3
    e.residual_set_loads(v["cstmap"], v["pbar"], v["load"], v["mdiag"], v["velocity"], v["accel"],
                         v["contact_force"], v["d_lagrange_forces"])
    for itr, u, du in iters:
        ifv = e.drive_eps_sig_internal_forces(step, itr, u, du)[0].copy()
        v["ifv"], v["du"] = ifv, du      # the engine works on its own ifv: the restatement gets exactly that vector
        if itr == 0:
            res_o, norm0 = op.upres_iter_0(v["cstmap"], v["pbar"], ifv, du, dt, nbeta, v["mdiag"], v["contact_force"])
            n_o = dict(mgload=0.0, sum_loads=0.0, num_term_loads=0)
            n_o.update(op.cvtest_residual(res_o, v["d_lagrange_forces"]))
            if not mpc:
                assert n_o["norm_res"] == norm0
        else:
            res_o, n_o = _oracle_upres(v, dt, nbeta)
        res, n = e.residual(itr, dt, nbeta)
        assert np.array_equal(res, res_o), itr              # same expression, same order: same bits
        for k in ("mgload", "sum_loads", "norm_res"):
            assert _close(n[k], n_o[k]), (itr, k)
        assert n["num_term_loads"] == n_o["num_term_loads"] and n["res_max_dof"] == n_o["res_max_dof"]
        assert n["res_max"] == n_o["res_max"]
        # resident form: same norms, nothing copied but the scalars; and reproducible
        _, n2 = e.residual(itr, dt, nbeta, to_host=False)
        assert n2 == n
    e.close()


@pytest.mark.gpu
def test_device_residual_edge_cases_and_refusals():
    from warp3d_b200.engine import Engine, EngineError
    m, step, iters = _engine_case(seed=4)
    e = Engine(m)
    with pytest.raises(EngineError) as err:
        e.residual(1)
    assert "w3d_residual_set_loads" in str(err.value)
    z = np.zeros(m.nodof)
    e.residual_set_loads(np.zeros(m.nodof, dtype=np.int32), z)
    with pytest.raises(EngineError) as err:          # no stress pass yet: there are no internal forces
        e.residual(1)
    assert "stress pass" in str(err.value)
    e.drive_eps_sig_internal_forces(1, 0, z, z)          # zero displacements: ifv == 0
    res, n = e.residual(1)
    assert not res.any() and n == dict(norm_res=0.0, mgload=0.0, sum_loads=0.0, num_term_loads=0,
                                       res_max=0.0, res_max_dof=1)
    with pytest.raises(EngineError):
        e.residual(1, dt=0.0)
    with pytest.raises(ValueError):
        e.residual_set_loads(np.zeros(3, dtype=np.int32), z)
    e.close()


@pytest.mark.parametrize("n", [1, 2, 31, 257, 1025])
def test_restatement_on_ragged_sizes(n):
    """sizes around the reduction widths of the device kernel (warp, CTA), one dof included"""
    for seed, (dyn, mpc, con) in enumerate([(False, False, False), (True, True, True)]):
        v = _vectors(n, 100 + seed, dyn, mpc, con)
        res, n1 = _oracle_upres(v, 0.1, 0.25)
        res_np, n_np = _numpy_upres(v, 0.1, 0.25)
        assert np.array_equal(res, res_np)
        assert n1["num_term_loads"] == n_np["num_term_loads"] and n1["res_max_dof"] == n_np["res_max_dof"]
        assert n1["res_max"] == n_np["res_max"]
        for k in ("mgload", "sum_loads", "norm_res"):
            assert _close(n1[k], n_np[k]), (n, k)
