"""Generates tests/golden/*.npz from the CPU oracle (oracle/, gcc -O2 -ffp-contract=off; for mm10
and mm05 that is the statement-for-statement restatement in oracle/w3d_oracle_mm10.h / _mm05.h,
not the engine's re-ordered cores).

These are *regression* vectors: the reference (ifort-only Fortran + MKL) cannot be built or run
in this environment and ships no function-level golden vectors for the path, so the fixtures pin
the oracle's own output, not the Fortran's; what ties the oracle to the Fortran is in
tests/test_reference_known_answers.py (status per model: oracle/w3d_oracle.h).
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from common import load_path  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from warp3d_b200 import (Material, attach_crystals, build_structured_hex8, make_crystal_type,  # noqa: E402
                         new_state)

CASES = {
    # name: (builder kwargs, load-path kwargs)
    "hex8_geonl_bbar_mm01": (dict(nx=3, ny=2, nz=2, coord_jitter=0.15, seed=3, block_size=7),
                             dict(nsteps=2, strain=3.0e-3)),
    "hex8_small_mm01_kinematic": (dict(nx=2, ny=2, nz=2, coord_jitter=0.1, seed=9, geonl=False, bbar=False,
                                       materials=[Material(e=70.0e9, nu=0.33, yld_pt=250.0e6, tan_e=1.0e9, beta=0.25)]),
                                  dict(nsteps=2, strain=5.0e-3)),
    # crystal plasticity: fcc, Voce, 2 crystals per Gauss point
    "hex8_geonl_mm10_voce": (dict(nx=2, ny=1, nz=1, coord_jitter=0.1, seed=4, dt=0.5,
                                  materials=[Material(mat_type=10)], crystals=2),
                             dict(nsteps=2, strain=2.5e-3)),
}
CRYSTAL = dict(e=100000.0, nu=0.3, harden_n=20, tau_y=100.0, tau_v=50.0, theta_o=500.0, voche_m=1.0)


def build(name):
    bk, lk = CASES[name]
    bk = dict(bk)
    nx, ny, nz = bk.pop("nx"), bk.pop("ny"), bk.pop("nz")
    ncrys = bk.pop("crystals", 0)
    m = build_structured_hex8(nx, ny, nz, **bk)
    if ncrys:
        attach_crystals(m, make_crystal_type(**CRYSTAL), ncrystals=ncrys)
    return m, load_path(m, **lk)


def run(name, backend_factory):
    """Runs the case; returns dict of arrays keyed 's{step}i{iter}_{what}'."""
    m, path = build(name)
    be = backend_factory(m)
    out = {}
    for step, iters in path:
        for itr, u, du in iters:
            ifv, sc = be.eps_sig_ifv(step, itr, u, du)
            k = be.tanstf(step, itr, u, du)
            tag = f"s{step}i{itr}"
            out[tag + "_ifv"] = ifv.copy()
            out[tag + "_estiff"] = np.array(k, copy=True)
            out[tag + "_scalars"] = np.array([sc["internal_energy"], sc["plastic_work"], sc["sum_ifv"],
                                              float(sc["num_term_ifv"])])
            st = be.state()
            for nme in ("eps_n1", "urcs_n1", "rot_n1", "history1", "cep"):
                out[tag + "_" + nme] = getattr(st, nme).copy()
        be.update()
    return out


class OracleRun:
    def __init__(self, m):
        self.o, self.s = Oracle(m), new_state(m)

    def eps_sig_ifv(self, *a):
        return self.o.eps_sig_ifv(self.s, *a)

    def tanstf(self, *a):
        return self.o.tanstf(self.s, *a)

    def state(self):
        return self.s

    def update(self):
        self.o.update(self.s); self.o.adaptive_save(self.s)


if __name__ == "__main__":
    for name in CASES:
        out = run(name, OracleRun)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, len(out), "arrays", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")
