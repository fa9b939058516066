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


# example_problems_threads/test_1 (BASELINE.json configs[0]): incidences exactly as the deck lists
# them (WARP3D's l3disop node order), 20 elements on a 6 x 3 x 3 grid of 54 nodes
_TEST_1_INCID = (
    (10, 1, 4, 13, 11, 2, 5, 14), (28, 19, 22, 31, 29, 20, 23, 32), (46, 37, 40, 49, 47, 38, 41, 50),
    (11, 2, 5, 14, 12, 3, 6, 15), (29, 20, 23, 32, 30, 21, 24, 33), (47, 38, 41, 50, 48, 39, 42, 51),
    (13, 4, 7, 16, 14, 5, 8, 17), (31, 22, 25, 34, 32, 23, 26, 35), (49, 40, 43, 52, 50, 41, 44, 53),
    (14, 5, 8, 17, 15, 6, 9, 18), (32, 23, 26, 35, 33, 24, 27, 36), (50, 41, 44, 53, 51, 42, 45, 54),
    (19, 10, 13, 22, 20, 11, 14, 23), (37, 28, 31, 40, 38, 29, 32, 41), (20, 11, 14, 23, 21, 12, 15, 24),
    (38, 29, 32, 41, 39, 30, 33, 42), (22, 13, 16, 25, 23, 14, 17, 26), (40, 31, 34, 43, 41, 32, 35, 44),
    (23, 14, 17, 26, 24, 15, 18, 27), (41, 32, 35, 44, 42, 33, 36, 45))


def test_1_model():
    """example_problems_threads/test_1 as shipped: a cantilever of 20 l3disop elements (`nonlinear`
    = geonl, 2x2x2, bbar), 54 nodes, material `mises e 30000 nu 0.3 yld_pt 60.0 n_power 10.` (mm03),
    `blocking automatic` (one block), `bbar stabilization factor 0.0`, `time step 1.0e10`,
    consistent q-matrix on (the default).  Returns (model, deck) with deck = dict(fixed (0-based
    dofs: plane x=0 u=0 plus the listed v / w constraints), unit_load (nodof: the `unit` nodal
    force pattern), factors (load multipliers of steps 1-7 of loading end_shear))."""
    coords = np.array([[2.0 * ix, 0.5 * iy, -0.5 * iz] for ix in range(6) for iy in range(3) for iz in range(3)])
    mat = Material(mat_type=3, e=30000.0, nu=0.3, yld_pt=60.0, tan_e=0.0, beta=1.0, extra={"n_power": 10.0})
    m = build_hex8_model(coords, np.array(_TEST_1_INCID, dtype=np.int32), materials=[mat], geonl=True, bbar=True,
                         eps_bbar=0.0, qbar_flag=True, adaptive_flag=True, extrapolated_du=False, dt=1.0e10)
    fixed = [3 * n for n in range(9)]                         # plane x=0 u=0: nodes 1-9
    fixed += [3 * (2 - 1) + 2, 3 * (4 - 1) + 1, 3 * (5 - 1) + 1, 3 * (5 - 1) + 2, 3 * (6 - 1) + 1, 3 * (8 - 1) + 2]
    unit = np.zeros(m.nodof)
    for nodes, f in (((46, 48, 52, 54), 0.0625), ((47, 49, 51, 53), 0.125), ((50,), 0.25)):
        for n in nodes:
            unit[3 * (n - 1) + 1] = f
    return m, dict(fixed=np.array(sorted(fixed)), unit_load=unit, factors=[2.0, 1.0, 0.2, 0.2, 0.2, 0.2, 0.2])
