"""mm03 (mises, constant / power-law hardening, optional rate sensitivity): the engine's per-point
core (warp3d_b200/csrc/w3d_mm03_core.h, compiled for the HOST here; the same source runs on the
GPU in k_mm03) against the independent block-structured restatement of mm03.f
(oracle/w3d_oracle_mm03.h) on the same inputs -- random multi-step strain paths with loading,
elastic unloading and reloading, blocks with null points, all-null blocks, user initial stresses on
step 1.  Bar: 1e-10 on every stress / history / tangent word, the state flag word bit for bit.
Physical checks on the restatement itself: the updated stress sits on the yield surface, the
flow stress follows the uniaxial curve, the consistent tangent matches a finite difference."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "mm03_point_check.cpp")
LIB = os.path.join(ROOT, "build", "tests", "libmm03_point_check.so")


def _lib():
    deps = [SRC, os.path.join(ROOT, "oracle", "w3d_oracle_mm03.h"),
            os.path.join(ROOT, "warp3d_b200", "csrc", "w3d_mm03_core.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC],
                       check=True)
    L = C.CDLL(LIB)
    L.mm03_block_both.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def block_both(props, step, it, dt, hist, cgn, uddt):
    L = _lib()
    span = props.shape[0]
    out = {}
    for tag in ("lit", "core"):
        out[tag] = dict(h1=np.zeros((span, 15)), cgn1=np.zeros((span, 9)), cep=np.zeros((span, 21)),
                        rc=np.zeros(span, dtype=np.int32))
    rc = L.mm03_block_both(span, step, it, C.c_double(dt), _p(props), _p(hist), _p(cgn), _p(uddt),
                           _p(out["lit"]["h1"]), _p(out["lit"]["cgn1"]), _p(out["lit"]["cep"]), _p(out["lit"]["rc"]),
                           _p(out["core"]["h1"]), _p(out["core"]["cgn1"]), _p(out["core"]["cep"]), _p(out["core"]["rc"]))
    assert rc == 0
    return out["lit"], out["core"]


def _compare(lit, core, worst):
    assert (lit["rc"] == core["rc"]).all()
    ok = lit["rc"] == 0
    # state flag word: bit for bit
    assert (lit["h1"][ok, 5].view(np.int64) == core["h1"][ok, 5].view(np.int64)).all()
    groups = {"stress": (slice(0, 6), "cgn1"), "energy": (slice(6, 9), "cgn1"), "cep": (slice(0, 21), "cep"),
              "ebarp": (slice(0, 1), "h1"), "sbar_p_h": (slice(1, 4), "h1"), "dep_deq_q": (slice(6, 9), "h1"),
              "eps_el": (slice(9, 15), "h1")}
    for name, (sl, key) in groups.items():
        a, b = lit[key][ok][:, sl], core[key][ok][:, sl]
        if a.size == 0:
            continue
        scale = max(np.abs(a).max(), 1e-300)
        worst[name] = max(worst.get(name, 0.0), float(np.abs(a - b).max() / scale))


CASES = {
    "power_law_n10": dict(e=200.0e9, nu=0.3, yld=400.0e6, h=0.0, n=10.0, m=0.0, eta=0.0),
    "power_law_n4": dict(e=70.0e9, nu=0.33, yld=250.0e6, h=0.0, n=4.0, m=0.0, eta=0.0),
    "constant_h": dict(e=200.0e9, nu=0.3, yld=400.0e6, h=2.0e9 * 200.0e9 / (200.0e9 - 2.0e9), n=0.0, m=0.0, eta=0.0),
    "perfectly_plastic": dict(e=30000.0, nu=0.3, yld=60.0, h=0.0, n=0.0, m=0.0, eta=0.0),
    "rate_power_law": dict(e=200.0e9, nu=0.3, yld=400.0e6, h=0.0, n=10.0, m=15.0, eta=1.0 / 0.5),
    "rate_constant_h": dict(e=200.0e9, nu=0.3, yld=400.0e6, h=4.0e9, n=0.0, m=8.0, eta=1.0 / 2.0),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_core_matches_restatement_on_strain_paths(case):
    k = CASES[case]
    rng = np.random.default_rng(zlib.crc32(case.encode()) % 10000)
    span = 24
    props = np.tile(np.array([k["e"], k["nu"], k["yld"], k["h"], k["n"], k["m"], k["eta"]]), (span, 1)).copy()
    eps_y = k["yld"] / k["e"]
    hist = np.zeros((span, 15)); cgn = np.zeros((span, 9))
    # user initial stresses on a few points (step 1 builds the elastic strains from them)
    cgn[::5, :6] = 0.2 * k["yld"] * rng.normal(size=(len(cgn[::5]), 6))
    direction = rng.normal(size=(span, 6)); direction[:, :3] -= 0.8 * direction[:, :3].mean(axis=1, keepdims=True)
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    worst, nplastic, nunload = {}, 0, 0
    dt = 0.7
    for step in range(1, 9):
        sgn = -1.0 if step in (4, 5) else 1.0             # unload, reverse, reload
        amp = (0.8 if step == 1 else 1.7) * eps_y
        for it in ((0, 1, 2) if step > 1 else (1, 2)):
            uddt = sgn * amp * (direction + 0.15 * rng.normal(size=(span, 6))) * (0.4 if it == 0 else 1.0)
            uddt[3] = 0.0                                  # a null point in a live block
            uddt[7] = 1.0e-14 * rng.normal(size=6)         # ... and one just inside the null threshold
            lit, core = block_both(props, step, it, dt, hist, cgn, uddt)
            assert (lit["rc"] == 0).all()
            _compare(lit, core, worst)
            st = lit["h1"][:, 5].view(np.int64) & 0xFFFFFFFF
            nplastic += int((st == 1).sum())
            nunload += int(((st == 0xFFFFFFFF) & (lit["h1"][:, 0] > 0)).sum())
            # on the yield surface, flow stress = q of the updated stress
            pl = st == 1
            if pl.any():
                s = lit["cgn1"][pl]
                q = np.sqrt(0.5 * ((s[:, 0] - s[:, 1]) ** 2 + (s[:, 1] - s[:, 2]) ** 2 + (s[:, 0] - s[:, 2]) ** 2) +
                            3.0 * (s[:, 3] ** 2 + s[:, 4] ** 2 + s[:, 5] ** 2))
                assert np.abs(q / lit["h1"][pl, 1] - 1.0).max() < 5e-6
        hist, cgn = lit["h1"].copy(), lit["cgn1"].copy()   # converged: n+1 -> n
    assert nplastic > 50 and nunload > 5, (nplastic, nunload)
    bad = {n: v for n, v in worst.items() if v > 1e-10}
    assert not bad, (bad, worst)


def test_all_null_block_copies_the_state_and_marks_points_elastic():
    k = CASES["power_law_n10"]
    span = 5
    props = np.tile(np.array([k["e"], k["nu"], k["yld"], k["h"], k["n"], k["m"], k["eta"]]), (span, 1)).copy()
    rng = np.random.default_rng(3)
    hist = np.zeros((span, 15)); cgn = np.zeros((span, 9))
    uddt = 2.5e-3 * rng.normal(size=(span, 6))
    lit, core = block_both(props, 1, 1, 1.0, hist, cgn, uddt)
    hist, cgn = lit["h1"].copy(), lit["cgn1"].copy()
    assert ((hist[:, 5].view(np.int64) & 0xFFFFFFFF) == 1).any()
    lit, core = block_both(props, 2, 1, 1.0, hist, cgn, np.zeros((span, 6)))
    for out in (lit, core):
        assert (out["cgn1"] == cgn).all()
        assert (out["h1"][:, :5] == hist[:, :5]).all() and (out["h1"][:, 9:] == hist[:, 9:]).all()
        assert ((out["h1"][:, 5].view(np.int64) & 0xFFFFFFFF) == 0xFFFFFFFF).all()
        assert (out["h1"][:, 6:9] == 0).all()
    assert (lit["cep"] == core["cep"]).all()
    # the same zero increment on a block that has one live point: plastic points are skipped
    uddt = np.zeros((span, 6)); uddt[0] = 1.0e-4
    lit, core = block_both(props, 2, 1, 1.0, hist, cgn, uddt)
    worst = {}
    _compare(lit, core, worst)
    assert all(v < 1e-10 for v in worst.values()), worst
    was_plastic = hist[:, 0] > 0
    assert (lit["cgn1"][1:][was_plastic[1:], 6:9] == 0).all()


def test_uniaxial_curve_and_consistent_tangent():
    """Strain-driven uniaxial-stress-like path: the flow stress follows sig_o (eps/eps_o)^(1/n) beyond
    the transition, and cep of iteration >= 2 is the derivative of the update (finite difference)."""
    k = CASES["power_law_n10"]
    props = np.array([[k["e"], k["nu"], k["yld"], k["h"], k["n"], k["m"], k["eta"]]])
    hist = np.zeros((1, 15)); cgn = np.zeros((1, 9))
    d0 = np.array([[1.0, -0.45, -0.45, 0.2, 0.0, 0.1]]) * 6.0e-3
    lit, core = block_both(props, 1, 2, 1.0, hist, cgn, d0)
    ebarp, sbar = lit["h1"][0, 0], lit["h1"][0, 1]
    eps_o = k["yld"] / k["e"]
    assert ebarp > 0
    assert abs(sbar - k["yld"] * ((ebarp + sbar / k["e"]) / eps_o) ** (1.0 / k["n"])) < 1e-6 * sbar
    D = np.zeros((6, 6)); idx = 0
    for r in range(6):
        for c in range(r + 1):
            D[r, c] = D[c, r] = lit["cep"][0, idx]; idx += 1
    for c in range(6):
        h = 1.0e-8
        dp, dm = d0.copy(), d0.copy()
        dp[0, c] += h; dm[0, c] -= h
        sp = block_both(props, 1, 2, 1.0, hist, cgn, dp)[0]["cgn1"][0, :6]
        sm = block_both(props, 1, 2, 1.0, hist, cgn, dm)[0]["cgn1"][0, :6]
        fd = (sp - sm) / (2 * h)
        assert np.abs(fd - D[:, c]).max() < 2e-4 * np.abs(D).max(), (c, fd, D[:, c])


def test_driver_patch_test_uniform_strain():
    """mm03 under the oracle's block driver: a homogeneous displacement gradient on a distorted
    mesh gives the same stress and history at every Gauss point and self-equilibrated internal
    forces (interior nodes carry none)."""
    from oracle.oracle_py import Oracle
    from warp3d_b200 import Material, build_structured_hex8, new_state
    mat = Material(mat_type=3, e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=0.0, extra={"n_power": 10.0})
    m = build_structured_hex8(4, 3, 3, materials=[mat], geonl=False, bbar=False, coord_jitter=0.15, seed=5, block_size=11)
    assert (m.hist_size == 15).all()
    x = m.c.reshape(-1, 3)
    G = np.array([[4.0e-3, 1.0e-3, 0.0], [0.0, -1.5e-3, 5.0e-4], [2.0e-4, 0.0, -1.2e-3]])
    du = (x @ G.T).reshape(-1)
    o, s = Oracle(m), new_state(m)
    f, sc = o.eps_sig_ifv(s, 1, 1, np.zeros(m.nodof), du)
    sig = s.urcs_n1.reshape(-1, 9)
    h1 = s.history1.reshape(-1, 15)
    assert (h1[:, 0] > 0).all() and sc["plastic_work"] > 0
    assert np.abs(sig[:, :6] - sig[0, :6]).max() < 1e-9 * np.abs(sig[0, :6]).max()
    assert np.abs(h1[:, 0] - h1[0, 0]).max() < 1e-9 * h1[0, 0]
    xr = m.c.reshape(-1, 3)
    g = m.grid
    interior = np.ones(m.nonode, dtype=bool)
    for d in range(3):
        interior &= (xr[:, d] > xr[:, d].min() + 0.3) & (xr[:, d] < xr[:, d].max() - 0.3)
    assert interior.any()
    assert np.abs(f.reshape(-1, 3)[interior]).max() < 1e-9 * np.abs(f).max()
