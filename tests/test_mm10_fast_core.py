"""The engine's fast mm10 core (warp3d_b200/csrc/w3d_mm10_fast.h: lattice-frame slip sums, one
flat solver loop, tangent from the LU of the full Jacobian) compiled for the HOST and run point by
point against the independent statement-for-statement restatement (oracle/w3d_oracle_mm10.h) on the
same inputs: random orientations, rotations and strain paths over several load steps, elastic
iteration-0 shortcut, several crystals per point, geometric hardening (k_0 != 0), general (pow)
and whole-number exponents.  Bar: 1e-10 per history word group (north_star).  The same core runs
on the GPU as k_mm10f; tests/test_gpu_mm10.py is the device-side check."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from warp3d_b200 import make_crystal_type, mm10_hist_size

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "mm10_point_check.cpp")
LIB = os.path.join(ROOT, "build", "tests", "libmm10_point_check.so")


def _lib():
    deps = [SRC, os.path.join(ROOT, "oracle", "w3d_oracle_mm10.h"),
            os.path.join(ROOT, "warp3d_b200", "csrc", "w3d_mm10_fast.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC],
                       check=True)
    L = C.CDLL(LIB)
    L.mm10_point_both.restype = C.c_int
    return L


def _rot(rng, angle):
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run_paths(ct, ncrystals=1, npoints=40, nsteps=4, strain=2.5e-3, dt=1.0, taul=False, seed=0, rot=0.02, asym=0):
    L = _lib()
    rng = np.random.default_rng(seed)
    nslip = int(ct[337])
    H = mm10_hist_size(nslip, ncrystals)
    worst = {}
    ncut = 0
    nplastic = 0
    for pt in range(npoints):
        ang = np.stack([rng.uniform(0, 360, ncrystals), rng.uniform(0, 180, ncrystals),
                        rng.uniform(0, 360, ncrystals)], axis=1).copy()
        hist = np.zeros(H)
        cgn = np.zeros(9)
        direction = rng.normal(size=6); direction[:3] -= direction[:3].mean() * 0.7
        direction /= np.linalg.norm(direction)
        for step in range(1, nsteps + 1):
            sgn = -1.0 if step == nsteps else 1.0            # unload at the end
            for it0 in (1, 0):                               # iteration 0 (elastic shortcut), then a real one
                uddt = sgn * strain * (direction + 0.2 * rng.normal(size=6)) * (0.3 if it0 else 1.0)
                Rur = _rot(rng, rot * step).T.reshape(-1).copy()   # column-major R
                if taul and step > 1:                        # some lattice curvature in history at n
                    hist[36:63] = 5.0 * rng.normal(size=27)      # tau_l of a few MPa
                h_lit, h_fast = np.zeros(H), np.zeros(H)
                c_lit, c_fast = np.zeros(9), np.zeros(9)
                cuts = (C.c_int * 2)()
                rc = L.mm10_point_both(_p(ct), ncrystals, _p(ang), 1, step, _p(hist), _p(Rur), _p(uddt),
                                       C.c_double(dt), it0, asym, _p(cgn), int(taul), _p(h_lit), _p(c_lit),
                                       _p(h_fast), _p(c_fast), cuts)
                assert rc == 0
                assert cuts[0] == cuts[1], (pt, step, it0)
                if cuts[0]:
                    ncut += 1
                    break
                groups = {"stress": (c_lit[:6], c_fast[:6]), "work": (c_lit[6:], c_fast[6:]),
                          "D": (h_lit[:36], h_fast[:36]), "R": (h_lit[63:72], h_fast[63:72]),
                          "work_h": (h_lit[72:75], h_fast[72:75]), "sliphist": (h_lit[75:75 + nslip], h_fast[75:75 + nslip])}
                c0 = 75 + nslip
                cs = 30 + nslip + 1 + 15 + 1 + 12
                for c in range(ncrystals):
                    b = c0 + c * cs
                    for name, lo, hi in (("c_stress", 0, 6), ("c_euler", 6, 9), ("c_Rp", 9, 18), ("c_D", 18, 24),
                                         ("c_eps", 24, 30), ("c_slip", 30, 30 + nslip), ("c_tt", 30 + nslip, 31 + nslip),
                                         ("c_rate", 46 + nslip, 47 + nslip), ("c_ep", 47 + nslip, 53 + nslip),
                                         ("c_ed", 53 + nslip, 59 + nslip)):
                        groups[f"{name}"] = (h_lit[b + lo:b + hi], h_fast[b + lo:b + hi]) if c == 0 else \
                            (np.r_[groups[name][0], h_lit[b + lo:b + hi]], np.r_[groups[name][1], h_fast[b + lo:b + hi]])
                    u_l, u_f = h_lit[b + 31 + nslip:b + 46 + nslip], h_fast[b + 31 + nslip:b + 46 + nslip]
                    # u(7), u(8): integers (most active system, number of active systems) -- exact
                    assert u_l[6] == u_f[6] and u_l[7] == u_f[7], (pt, step, it0, u_l, u_f)
                    # the user words nothing is stored in stay zero (several crystals per point: 12 of
                    # them hold tangent columns during the solve)
                    assert not u_l[[0, 1, 2, 3, 4, 8, 9]].any() and not u_f[[0, 1, 2, 3, 4, 8, 9]].any(), (pt, step, u_f)
                    for k in (5, 10, 11, 12, 13, 14):
                        if abs(u_l[k]) < 1e9:                # n_eff = 1e10 sentinel compared exactly below
                            groups[f"u{k}"] = (np.atleast_1d(u_l[k]), np.atleast_1d(u_f[k]))
                        else:
                            assert u_l[k] == u_f[k]
                for name, (a, b_) in groups.items():
                    scale = np.abs(a).max()
                    if name in ("c_euler",):
                        scale = 360.0
                    if name in ("c_slip", "sliphist", "c_ep", "c_ed", "work", "work_h", "c_rate"):
                        scale = max(scale, 1e-30)
                    err = np.abs(a - b_).max() / scale if scale > 0 else np.abs(b_).max()
                    worst[name] = max(worst.get(name, 0.0), err)
                if not it0:
                    nplastic += int(np.abs(h_lit[c0 + 30:c0 + 30 + nslip]).max() > 1e-8)
                    hist, cgn = h_lit.copy(), c_lit.copy()   # roll over with the literal state
            else:
                continue
            break
    return worst, ncut, nplastic


FCC = dict(e=150000.0, nu=0.28, harden_n=20.0, voche_m=5.0, theta_o=400.0, tau_y=300.0, tau_v=1200.0)


@pytest.mark.parametrize("case", ["bench_fcc", "voche_m1", "bcc", "general_exponents", "polycrystal3",
                                  "geometric_hardening", "alter_mode", "asymmetric_tangent",
                                  "polycrystal3_asymmetric_tangent"])
def test_fast_core_matches_literal_restatement(case):
    kw = dict(FCC)
    run = dict(strain=6.0e-3)
    if case == "voche_m1":
        kw = dict(e=100000.0, nu=0.3, harden_n=20, tau_y=100.0, tau_v=50.0, theta_o=500.0, voche_m=1.0)
        run = dict(strain=2.5e-3, dt=0.5)
    elif case == "bcc":
        kw.update(slip="bcc")
    elif case == "general_exponents":
        kw.update(harden_n=12.5, voche_m=1.7)
    elif case == "polycrystal3":
        run.update(ncrystals=3, npoints=15)
    elif case == "asymmetric_tangent":            # asymmetric assembly: mm10_a_make_symm_1 skipped
        run.update(asym=1, npoints=15)
    elif case == "polycrystal3_asymmetric_tangent":
        run.update(ncrystals=3, npoints=10, asym=1)
    elif case == "geometric_hardening":
        kw.update(k_o=1.0)
        run.update(taul=True)
    elif case == "alter_mode":
        kw.update(alter_mode=True, eps_dot_o_y=4.0e-3, harden_n=5.0)
    ct = make_crystal_type(**kw)
    worst, ncut, nplastic = _run_paths(ct, seed=zlib.crc32(case.encode()) % 1000, **run)
    assert nplastic > 10, "the paths must reach the plastic range"
    bad = {k: v for k, v in worst.items() if v > 1e-10}
    assert not bad, (bad, worst)
