"""Models of the reference's own decks that BASELINE.json names, rebuilt from fixtures.

``seb_model`` = fracture_models/seb (BASELINE.json configs[1]): a 3-D SE(B) fracture specimen,
17 318 hex8 (l3disop, geonl, 2x2x2, B-bar), 19 935 nodes, the deck's USER blocking table (137
blocks, 133 x 128 + 77, 60, 58, 99) with its 4th column as MPI domains 0 / 1, `consistent q off`, `bbar
stabilization factor 0.0`, `adaptive on`, `extrapolate on` (warp3d.inp:47-72).  The fixture
(tests/golden/seb_deck.npz) is written by tests/golden/make_seb_fixture.py from the deck files.

The deck ships with material `mises` (mm03); BASELINE's config runs it with mm01, i.e. with the
card `properties bilinear e 200.0e9 nu 0.3 yld_pt 400.0e6 tan_e 2.0e9` (src/inmat.f:793-796
sets matprp(9) = 1 for `bilinear`).  `material="mises"` keeps the shipped card (mm03, n_power 10).
"""
from __future__ import annotations

import os

import numpy as np

from .model import Material, build_hex8_model

SEB_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                           "seb_deck.npz")


def seb_model(material="bilinear", fixture=None, domains=True):
    """Returns (model, constraints) with constraints = dict(dofs (0-based structure dofs),
    values (imposed displacement per load step, `step 1-160 constraints 1.0`))."""
    d = np.load(fixture or SEB_FIXTURE)
    if material == "bilinear":
        mat = Material(mat_type=1, e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=2.0e9, beta=1.0)
    elif material == "mises":
        mat = Material(mat_type=3, e=200.0e9, nu=0.3, yld_pt=400.0e6, tan_e=0.0, beta=1.0,
                       extra={"n_power": 10.0})
    else:
        raise ValueError("material: 'bilinear' (mm01, BASELINE config) or 'mises' (mm03, as shipped)")
    blocking = d["blocking"].copy()
    if not domains:
        blocking[:, 2] = 0
    m = build_hex8_model(d["coords"], d["incid"], materials=[mat], geonl=True, bbar=True,
                         blocking=blocking, eps_bbar=0.0, qbar_flag=False, adaptive_flag=True,
                         extrapolated_du=False, dt=1.0e6)   # extrapolate on acts from step 2; callers set it
    dofs = 3 * (d["cons_node"].astype(np.int64) - 1) + d["cons_dof"].astype(np.int64)
    return m, dict(dofs=dofs, values=d["cons_val"].copy())


def seb_bend_field(m, scale=1.0):
    """A smooth bending-like displacement field over the specimen (x: width direction, -W/2..W/2;
    y: along the span, 0 = crack plane; z: thickness): strains up to ~2.6e-3 at scale 1, i.e. part
    of the points beyond the yield strain 2e-3 of the deck's steel.  Synthetic load for parity
    tests and for the bench workload `seb_mm01` (the deck's own load needs the global solve)."""
    x = m.c.reshape(-1, 3)
    L = x[:, 0].max() - x[:, 0].min()
    xi = (x[:, 0] - x[:, 0].min()) / L
    yi = (x[:, 1] - x[:, 1].min()) / (x[:, 1].max() - x[:, 1].min())
    u = np.zeros_like(x)
    u[:, 1] = -7.0e-4 * scale * (1.0 - xi) ** 2 * L
    u[:, 0] = 5.25e-3 * scale * (yi - 0.5) * xi * L
    u[:, 2] = -0.3 * 3.5e-3 * scale * x[:, 2] * (yi - 0.5)
    return u.reshape(-1).copy()
