"""Convert the reference's fracture_models/seb deck (BASELINE.json configs[1]) into a fixture.

Run here, in the container that has /root/reference; the GPU box never reads the reference tree:

    python tests/golden/make_seb_fixture.py            # writes tests/golden/seb_deck.npz

What is taken from the deck (file:line refer to /root/reference/fracture_models/seb/):
  coordinates            19 935 nodes, x y z as printed (9 significant digits)
  incid_and_blocking     :3-17320   17 318 hex8 incidences, WARP3D node order, 1-based
                         :17325+    the USER blocking table  blk span felem domain  (137 blocks:
                                    133 x 128 + 77, 60, 58, 99; 4th column = MPI domain 0 / 1)
  constraints            node dof value lines (w = 0 symmetry plane, v = 0 crack plane, u = 0
                         support line, u = -6.35e-6 applied displacement increment per step)
  warp3d.inp             :37-39 material card; :47-49 element card (l3disop nonlinear = geonl,
                         2x2x2, bbar); :60-72 solution parameters (consistent q off, adaptive on,
                         bbar stabilization factor 0.0, extrapolate on, time step 1.0e6)

The material card as shipped is `mises` (matprp(9) = 3 -> mm03, src/inmat.f:816-823).  BASELINE's
config runs the deck with mm01: the card becomes `properties bilinear e 200.0e9 nu 0.3 yld_pt
400.0e6 tan_e 2.0e9` (matprp(9) = 1, src/inmat.f:793-796); that substitution is made by the
loader (warp3d_b200.decks.seb_model), not here -- the fixture holds only what the deck holds.
"""
import os
import sys

import numpy as np

SRC = "/root/reference/fracture_models/seb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "seb_deck.npz")


def data_lines(path):
    with open(path) as f:
        for ln in f:
            s = ln.strip()
            if not s or s.startswith(("c ", "c\t", "*")) or s == "c":
                continue
            yield s


def main():
    nonode, noelem = 19935, 17318            # warp3d.inp:41
    coords = np.zeros((nonode, 3))
    seen = 0
    for s in data_lines(os.path.join(SRC, "coordinates")):
        t = s.split()
        if t[0] == "coordinates":
            continue
        coords[int(t[0]) - 1] = [float(x) for x in t[1:4]]
        seen += 1
    assert seen == nonode, seen

    incid = np.zeros((noelem, 8), dtype=np.int32)
    blocking = []
    mode = None
    for s in data_lines(os.path.join(SRC, "incid_and_blocking")):
        t = s.split()
        if t[0] == "incidences":
            mode = "incid"; continue
        if t[0] == "blocking":
            mode = "blocking"; continue
        if mode == "incid":
            incid[int(t[0]) - 1] = [int(x) for x in t[1:9]]
        elif mode == "blocking":
            blk, span, felem = int(t[0]), int(t[1]), int(t[2])
            dom = int(t[3]) if len(t) > 3 else 0
            assert blk == len(blocking) + 1
            blocking.append((span, felem, dom))
    blocking = np.array(blocking, dtype=np.int32)
    assert incid.min() >= 1 and incid.max() <= nonode
    assert blocking[:, 0].sum() == noelem and len(blocking) == 137

    cnode, cdof, cval = [], [], []
    for s in data_lines(os.path.join(SRC, "constraints")):
        t = s.split()
        if t[0] == "constraints":
            continue
        node = int(t[0])
        for k in range(1, len(t), 2):
            cnode.append(node); cdof.append("uvw".index(t[k])); cval.append(float(t[k + 1]))
    np.savez_compressed(
        OUT, coords=coords, incid=incid, blocking=blocking,
        cons_node=np.array(cnode, dtype=np.int32), cons_dof=np.array(cdof, dtype=np.int8),
        cons_val=np.array(cval, dtype=np.float64),
        provenance=np.array("rhdodds/warp3d 18.4.2 fracture_models/seb: coordinates, incid_and_blocking, "
                            "constraints; converted by tests/golden/make_seb_fixture.py"))
    print(OUT, os.path.getsize(OUT), "bytes;", len(cnode), "constraints;",
          "domains", np.bincount(blocking[:, 2]).tolist())


if __name__ == "__main__":
    if not os.path.isdir(SRC):
        sys.exit("needs /root/reference (run in the build container)")
    main()
