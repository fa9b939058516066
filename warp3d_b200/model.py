"""Host-side mirror of the reference's global data for the element-block hot path.

Everything here is *input data* for the engine, laid out exactly as the reference's Fortran
modules hold it (column-major arrays, 1-based index tables), so the same arrays can be handed
to the C ABI (``include/w3d_engine.h``) that a Fortran caller would bind:

* ``elblks(0:3, nelblk)``            -- src/inelbk.f:47-52, auto blocking src/inelbk.f:282-345
* ``props/iprops/lprops(mxelpr, noelem)`` (one 4-byte table, three views) -- src/mod_main.f:41-46,
  rows filled as src/elprp.f:715-935 (elprp2, l3disop) does
* ``edest/cdest (totdof, span)``     -- src/reopen.f:1200-1408 (all u, all v, all w ordering)
* blocked state arrays               -- src/mod_eleblocks.f:52-131

The synthetic structured hex8 meshes follow SURVEY.md section 8(d).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MXVL = 128     # param_def: mxvl
MXELPR = 43    # param_def: mxelpr
MXMTPR = 300   # param_def: mxmtpr
NSTR = 6
NSTRS = 9
HIST_SIZE = {1: 11, 3: 15, 5: 14}   # mm01.f:976, mm03.f:1623-1640, mm05.f:4136
CEP_SIZE = 21

# WARP3D hex8 (l3disop) node ordering in (xi, eta, zeta) -- src/derivs.f:283-343
HEX8_NODE_SIGNS = np.array(
    [(-1, -1, +1), (-1, -1, -1), (-1, +1, -1), (-1, +1, +1),
     (+1, -1, +1), (+1, -1, -1), (+1, +1, -1), (+1, +1, +1)], dtype=np.int64)


@dataclass
class Material:
    """One entry of the reference's material table (src/inmat.f), bilinear (mm01) subset."""
    mat_type: int = 1
    e: float = 200.0e9
    nu: float = 0.3
    yld_pt: float = 400.0e6
    tan_e: float = 2.0e9
    beta: float = 1.0
    linear_elastic: bool = False
    extra: dict = field(default_factory=dict)   # mm05: matprp(55:64); mm10: crystal data


@dataclass
class Model:
    nonode: int
    noelem: int
    nodof: int
    nelblk: int
    c: np.ndarray          # (nodof,) float64 coordinates
    incid: np.ndarray      # (noelem, 8) int32, 1-based nodes, WARP3D ordering
    elblks: np.ndarray     # (nelblk, 4) int32 rows = [span, felem, domain, 0]  (Fortran (0:3,blk))
    props: np.ndarray      # (noelem, MXELPR) int32 words (Fortran (mxelpr, noelem))
    matprp: np.ndarray     # (nmat, MXMTPR) float32
    edest: np.ndarray      # concatenated per block (span, 24) int32 (Fortran (24, span))
    cdest: np.ndarray
    hist_size: np.ndarray  # (nelblk,) int32
    materials: list
    eps_bbar: float = 0.0
    qbar_flag: bool = True
    dt: float = 1.0
    total_model_time: float = 0.0
    beta_fact: float = 1.0
    extrapolated_du: bool = False
    adaptive_flag: bool = True
    grid: tuple = ()
    # crystal plasticity (mm10) input, see attach_crystals
    crystal_types: np.ndarray = None
    elem_cnum: np.ndarray = None
    elem_angles: np.ndarray = None
    ncrystals: int = 0
    angle_type: int = 1
    asymmetric_assembly: bool = False

    # ---- sizes of the blocked arrays (doubles) ----
    def span(self, blk):
        return int(self.elblks[blk, 0])

    def felem(self, blk):
        return int(self.elblks[blk, 1])

    def sizes(self):
        spans = self.elblks[:, 0].astype(np.int64)
        return dict(
            history=int((spans * 8 * self.hist_size).sum()),
            eps=int(spans.sum() * 8 * NSTR), urcs=int(spans.sum() * 8 * NSTRS),
            rot=int(spans.sum() * 8 * 9), cep=int(spans.sum() * 8 * CEP_SIZE),
            estiff=int(spans.sum() * 300), element_vol=int(spans.sum()))

    def block_offsets(self, per_elem):
        """offsets (doubles) of each block in an array with per_elem doubles per element."""
        spans = self.elblks[:, 0].astype(np.int64)
        per = np.asarray(per_elem, dtype=np.int64) * spans
        off = np.zeros(self.nelblk + 1, dtype=np.int64)
        np.cumsum(per, out=off[1:])
        return off


@dataclass
class State:
    """Blocked state in the reference's layouts, blocks concatenated in block order."""
    history: np.ndarray
    history1: np.ndarray
    eps_n: np.ndarray
    eps_n1: np.ndarray
    urcs_n: np.ndarray
    urcs_n1: np.ndarray
    rot_n: np.ndarray
    rot_n1: np.ndarray
    cep: np.ndarray
    estiff: np.ndarray
    element_vol: np.ndarray

    NAMES = ("history", "history1", "eps_n", "eps_n1", "urcs_n", "urcs_n1", "rot_n", "rot_n1",
             "cep", "estiff", "element_vol")

    def copy(self):
        return State(**{k: getattr(self, k).copy() for k in self.NAMES})


def new_state(model: Model) -> State:
    """history_cep_init / rotation_init (src/reopen.f:1015-1176, 1841-1960): zeros, R = I."""
    sz = model.sizes()
    z = lambda n: np.zeros(n, dtype=np.float64)
    rot = np.zeros((sz["rot"] // 9, 9), dtype=np.float64)
    rot[:, 0] = 1.0
    rot[:, 4] = 1.0
    rot[:, 8] = 1.0
    return State(history=z(sz["history"]), history1=z(sz["history"]), eps_n=z(sz["eps"]),
                 eps_n1=z(sz["eps"]), urcs_n=z(sz["urcs"]), urcs_n1=z(sz["urcs"]),
                 rot_n=rot.reshape(-1).copy(), rot_n1=rot.reshape(-1).copy(), cep=z(sz["cep"]),
                 estiff=z(sz["estiff"]), element_vol=z(sz["element_vol"]))


def cyclic_fa_material(e=200.0e9, nu=0.3, yld_pt=400.0e6, q_u=100.0e6, b_u=20.0, h_u=10.0e9,
                       gamma_u=50.0, sig_tol=0.001):
    """mm05 'cyclic' material, nonlinear_hardening (Frederick-Armstrong) option: the matprp(55:60)
    rows exactly as src/inmat.f:2003-2040, :2220, :2287 stores them (real*4 arithmetic):
    q_u, b = sqrt(2/3) b_u, h = (2/3) h_u, option flag +1, gamma = sqrt(2/3) gamma_u, sig_tol."""
    f = np.float32
    root23 = np.sqrt(f(2.0) / f(3.0)).astype(np.float32)
    twothrds = f(2.0) / f(3.0)
    rows = {55: f(q_u), 56: f(f(b_u) * root23), 57: f(f(h_u) * twothrds), 58: f(1.0),
            59: f(f(gamma_u) * root23), 60: f(sig_tol)}
    return Material(mat_type=5, e=e, nu=nu, yld_pt=yld_pt, tan_e=0.0, beta=1.0, extra={"matprp": rows})


def cyclic_gp_material(e=30000.0, nu=0.3, yld_pt=60.0, gp_h_u=400.0, gp_tau=0.0, gp_beta_u=200.0,
                       gp_delta_u=0.010, sig_tol=0.0001):
    """mm05 'cyclic' material, generalized_plasticity option: matprp(55:60) as src/inmat.f:2017-2036,
    :2272, :2312-2395 stores them -- the uniaxial values unconverted, option flag -1:
    gp_h_u, gp_tau, gp_beta_u, -1, gp_delta_u, sig_tol; yld_pt is gp_sigma_0 (props(23)).
    Defaults = the reference's own decks (verification/test39, testP: ksi units).  The option's
    quadratic for the plastic multiplier is accepted only if |a l^2 + b l + c| <= 1e-6 in
    absolute terms (mm05.f:2962-2965), which moduli in Pa cannot meet in double precision: with
    SI-sized numbers the model asks for a step cut, exactly as the Fortran would."""
    f = np.float32
    rows = {55: f(gp_h_u), 56: f(gp_tau), 57: f(gp_beta_u), 58: f(-1.0), 59: f(gp_delta_u), 60: f(sig_tol)}
    return Material(mat_type=5, e=e, nu=nu, yld_pt=yld_pt, tan_e=0.0, beta=1.0, extra={"matprp": rows})


def _f32_bits(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).view(np.int32)


def simple_blocking(noelem, keys, auto_size=MXVL):
    """inelbk_simple_blocking (src/inelbk.f:282-345, 497-548): sequential pass over the elements;
    a new block opens when the current one is full or when (etype, matnum, geonl, intord, intnum,
    bbar) differ from the block's first element.  keys: (noelem, k) ints.  Returns a list of
    (span, felem) with 1-based felem.  Vectorised: runs of equal keys are cut into full blocks."""
    keys = np.asarray(keys)
    change = np.ones(noelem, dtype=bool)
    if noelem > 1:
        change[1:] = (keys[1:] != keys[:-1]).any(axis=1)
    starts = np.nonzero(change)[0]
    ends = np.append(starts[1:], noelem)
    blocks = []
    for s, e in zip(starts.tolist(), ends.tolist()):
        for f in range(s, e, auto_size):
            blocks.append((min(auto_size, e - f), f + 1))
    return blocks


def build_structured_hex8(nx, ny, nz, materials=None, elem_material=None, geonl=True, bbar=True,
                          h=1.0, block_size=MXVL, numprocs=1, coord_jitter=0.0, seed=0,
                          **globals_) -> Model:
    """Structured box of nx*ny*nz hex8 unit cubes, nodes and elements numbered x-fastest."""
    materials = materials or [Material()]
    nnx, nny, nnz = nx + 1, ny + 1, nz + 1
    nonode = nnx * nny * nnz
    noelem = nx * ny * nz
    nodof = 3 * nonode
    k, j, i = np.meshgrid(np.arange(nnz), np.arange(nny), np.arange(nnx), indexing="ij")
    coords = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64) * h
    if coord_jitter:
        rng = np.random.default_rng(seed)
        interior = ((i.ravel() > 0) & (i.ravel() < nx) & (j.ravel() > 0) & (j.ravel() < ny)
                    & (k.ravel() > 0) & (k.ravel() < nz))
        coords[interior] += rng.uniform(-coord_jitter, coord_jitter, (int(interior.sum()), 3)) * h
    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ei, ej, ek = ei.ravel(), ej.ravel(), ek.ravel()
    incid = np.empty((noelem, 8), dtype=np.int32)
    for a, (sx, sy, sz) in enumerate(HEX8_NODE_SIGNS):
        ni = ei + (sx > 0)
        nj = ej + (sy > 0)
        nk = ek + (sz > 0)
        incid[:, a] = (nk * nny + nj) * nnx + ni + 1

    return build_hex8_model(coords, incid, materials=materials, elem_material=elem_material, geonl=geonl,
                            bbar=bbar, block_size=block_size, numprocs=numprocs, grid=(nx, ny, nz),
                            **globals_)


def build_hex8_model(coords, incid, materials=None, elem_material=None, geonl=True, bbar=True,
                     block_size=MXVL, numprocs=1, grid=None, blocking=None, **globals_) -> Model:
    """The reference's tables for an arbitrary mesh of l3disop elements: coords (nonode, 3),
    incid (noelem, 8) 1-based in WARP3D's node order.  blocking: a user blocking table
    (nelblk, 3) = span, felem, domain as the `blocking` input command gives it
    (src/inelbk.f:60-250: user-defined blocks, 4th column = MPI domain); None = automatic."""
    materials = materials or [Material()]
    coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    incid = np.ascontiguousarray(np.asarray(incid, dtype=np.int32).reshape(-1, 8))
    nonode, noelem = coords.shape[0], incid.shape[0]
    nodof = 3 * nonode
    c = coords.reshape(-1).copy()
    if elem_material is None:
        elem_material = np.ones(noelem, dtype=np.int32)
    elem_material = np.asarray(elem_material, dtype=np.int32)

    # ---- material table (src/inmat.f rows) and element property table (src/elprp.f:715-935)
    matprp = np.zeros((len(materials), MXMTPR), dtype=np.float32)
    for mi, mat in enumerate(materials):
        matprp[mi, 0] = mat.e
        matprp[mi, 1] = mat.nu
        matprp[mi, 2] = mat.beta
        matprp[mi, 3] = mat.tan_e
        matprp[mi, 4] = mat.yld_pt
        matprp[mi, 8] = mat.mat_type
        if mat.mat_type == 3:
            # `mises` card (inmat.f:816-823, defaults :307-327): m_power, n_power, ref_eps rows
            # 10-12; f0 = 0, q1 = q2 = q3 = 0, no nucleation
            matprp[mi, 9] = mat.extra.get("m_power", 0.0)
            matprp[mi, 10] = mat.extra.get("n_power", 0.0)
            matprp[mi, 11] = mat.extra.get("ref_eps", 0.0)
        for row, val in mat.extra.get("matprp", {}).items():
            matprp[mi, row - 1] = val
    props = np.zeros((noelem, MXELPR), dtype=np.int32)
    props[:, 0] = 2              # type l3disop
    props[:, 1] = 8              # nodes
    props[:, 2] = 24 * 65536 + 6
    props[:, 3] = 3              # dof / node
    props[:, 4] = 1              # integration order 2x2x2
    props[:, 5] = 8              # number of gauss points
    props[:, 10] = 6
    mi = elem_material - 1
    e32 = matprp[mi, 0]
    et32 = matprp[mi, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        h32 = ((e32 * et32) / (e32 - et32)).astype(np.float32)   # real*4 arithmetic, elprp.f:866
    props[:, 6] = _f32_bits(e32)
    props[:, 7] = _f32_bits(matprp[mi, 1])
    props[:, 13] = _f32_bits(matprp[mi, 2])
    props[:, 14] = _f32_bits(h32)
    props[:, 16] = np.where(np.array([materials[k].linear_elastic for k in mi]), -1, 0)
    props[:, 17] = -1 if geonl else 0      # ifort .true. = -1
    props[:, 18] = -1 if bbar else 0
    props[:, 22] = _f32_bits(matprp[mi, 4])
    props[:, 24] = matprp[mi, 8].astype(np.int32)
    is3 = props[:, 24] == 3
    if is3.any():      # elprp.f:869-897: the rows mm03 reads
        d = matprp[mi, 11].astype(np.float64)
        eta32 = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 0.0).astype(np.float32)
        for word, col in ((19, matprp[mi, 9]), (20, matprp[mi, 10]), (21, eta32), (25, matprp[mi, 13]),
                          (26, matprp[mi, 14]), (27, matprp[mi, 15]), (28, matprp[mi, 16]),
                          (30, matprp[mi, 18]), (31, matprp[mi, 19]), (32, matprp[mi, 20])):
            props[:, word] = np.where(is3, _f32_bits(col.astype(np.float32)), props[:, word])
    props[:, 37] = elem_material
    props[:, 41] = -1

    keys = np.stack([props[:, 0], props[:, 37], props[:, 17], props[:, 4], props[:, 5],
                     props[:, 18]], axis=1)
    if blocking is not None:
        blocking = np.asarray(blocking, dtype=np.int32).reshape(-1, 3)
        blocks = [(int(sp), int(fe)) for sp, fe, _ in blocking]
        # the checks of inelbk (src/inelbk.f:170-250): blocks cover 1..noelem in order, span <= mxvl,
        # and every element of a block shares the block keys of its first element
        nxt = 1
        for sp, fe in blocks:
            if fe != nxt or sp < 1 or sp > MXVL:
                raise ValueError(f"user blocking: block at element {fe} (span {sp}) breaks the 1..noelem cover")
            if (keys[fe - 1:fe - 1 + sp] != keys[fe - 1]).any():
                raise ValueError(f"user blocking: block at element {fe} mixes element kinds")
            nxt = fe + sp
        if nxt != noelem + 1:
            raise ValueError("user blocking does not cover every element")
    else:
        blocks = simple_blocking(noelem, keys, block_size)
    nelblk = len(blocks)
    elblks = np.zeros((nelblk, 4), dtype=np.int32)
    for b, (span, felem) in enumerate(blocks):
        elblks[b, 0] = span
        elblks[b, 1] = felem
    if blocking is not None:
        elblks[:, 2] = blocking[:, 2]
    elif numprocs > 1:                   # round-robin, src/inelbk.f:368-378
        elblks[:, 2] = np.arange(nelblk) % numprocs

    # edest / cdest: dstmap(node) = 3*(node-1)+1 ; ordering all u, all v, all w
    base = 3 * (incid.astype(np.int64) - 1) + 1        # (noelem, 8)
    dest = np.concatenate([base, base + 1, base + 2], axis=1).astype(np.int32)  # (noelem, 24)
    edest = dest.reshape(-1).copy()
    cdest = edest.copy()
    hist_size = np.array([HIST_SIZE.get(int(props[f - 1, 24]), 0) for _, f in blocks],
                         dtype=np.int32)
    for b, (_, f) in enumerate(blocks):
        mt = materials[int(props[f - 1, 37]) - 1]
        if "hist_size" in mt.extra:
            hist_size[b] = mt.extra["hist_size"]
    return Model(nonode=nonode, noelem=noelem, nodof=nodof, nelblk=nelblk, c=c, incid=incid,
                 elblks=elblks, props=props, matprp=matprp, edest=edest, cdest=cdest,
                 hist_size=hist_size, materials=materials, grid=grid, **globals_)


def smooth_displacement(model: Model, strain=1.0e-3, seed=1, noise=1.0e-4, axis=0):
    """Uniaxial stretch along `axis` with lateral Poisson contraction plus a seeded perturbation
    (SURVEY 8(d)).  Returns a (nodof,) vector."""
    xyz = model.c.reshape(-1, 3)
    u = np.zeros_like(xyz)
    nu = 0.3
    for a in range(3):
        u[:, a] = (strain if a == axis else -nu * strain) * xyz[:, a]
    if noise:
        rng = np.random.default_rng(seed)
        u += rng.uniform(-1.0, 1.0, u.shape) * noise * strain * 10.0
    return u.reshape(-1).copy()


def graded_stretch(model: Model, eps_lo=1.0e-3, eps_hi=3.0e-3, seed=1, noise=1.0e-4, nu=0.3):
    """Benchmark displacement field: uniaxial stretch along x whose axial strain
    grows linearly from eps_lo at x=0 to eps_hi at x=L (so that about half of the Gauss points of
    the default mm01 material, yield strain 2e-3, are plastic), Poisson contraction in y and z,
    plus a seeded perturbation of relative size `noise`.  Returns a (nodof,) vector."""
    xyz = model.c.reshape(-1, 3)
    L = float(xyz[:, 0].max()) or 1.0
    x = xyz[:, 0]
    eps = eps_lo + (eps_hi - eps_lo) * x / L
    u = np.empty_like(xyz)
    u[:, 0] = eps_lo * x + (eps_hi - eps_lo) * x * x / (2.0 * L)
    u[:, 1] = -nu * eps * xyz[:, 1]
    u[:, 2] = -nu * eps * xyz[:, 2]
    if noise:
        rng = np.random.default_rng(seed)
        u += rng.uniform(-1.0, 1.0, u.shape) * noise * eps_hi
    return u.reshape(-1).copy()


def slab_domains(model: Model, numprocs: int) -> np.ndarray:
    """Contiguous block -> rank map (x-fastest element numbering makes these z-slabs of the
    structured meshes), balanced by element count.  Replaces the reference's round-robin default
    (src/inelbk.f:368-378), which makes every node an interface node.  Returns elblks(2,:)."""
    spans = model.elblks[:, 0].astype(np.int64)
    cum = np.cumsum(spans) - spans // 2
    dom = np.minimum((cum * numprocs) // int(spans.sum()), numprocs - 1)
    return dom.astype(np.int32)


# ------------------------------------------------------------------------------------------------
# crystal plasticity (mm10) input data: the c_array(cnum) table of src/mod_crystals.f and the
# per-element crystal numbers / Kocks angles of src/rknstr.f:1690-1720
# ------------------------------------------------------------------------------------------------
CRYSTAL_TYPE_DOUBLES = 36 + 144 + 144 + 13 + 7     # w3d_mm10::CrystalType, ints widened to double


def fcc_slip_systems():
    """slip directions bi and plane normals ni of slip_type 1 (fcc), src/mod_crystals.f:441-518."""
    f = 1.0 / np.sqrt(2.0)
    bi = np.array([[0, -f, f], [f, 0, -f], [-f, f, 0], [0, f, f], [f, 0, f], [f, -f, 0],
                   [0, -f, f], [-f, 0, -f], [f, f, 0], [0, f, f], [f, 0, -f], [-f, -f, 0]])
    g = 1.0 / np.sqrt(3.0)
    ni = np.array([[g, g, g]] * 3 + [[-g, -g, g]] * 3 + [[-g, g, g]] * 3 + [[g, -g, g]] * 3)
    return bi, ni


def bcc48_slip_systems():
    """The 48 bcc systems of slip_type 8 (src/mod_crystals.f:815-1210): <111> directions on the
    {110} (12), {112} (12) and {123} (24) planes, unit vectors.  Generated from the families, so
    the ORDER and the sign of each (b, n) pair are not the reference's table: the stress response
    and every sum over systems are invariant to both, only the per-system history columns
    (slip increments) would be permuted."""
    import itertools
    dirs = [np.array(d, dtype=float) for d in ((1, 1, 1), (-1, 1, 1), (1, -1, 1), (1, 1, -1))]
    out_b, out_n = [], []
    for fam in ((1, 1, 0), (1, 1, 2), (1, 2, 3)):
        planes = set()
        for perm in set(itertools.permutations(fam)):
            for sg in itertools.product((1, -1), repeat=3):
                v = tuple(int(s * c) for s, c in zip(sg, perm))
                if v != (0, 0, 0) and tuple(-x for x in v) not in planes:
                    planes.add(v)
        for pl in sorted(planes):
            nrm = np.array(pl, dtype=float)
            for d in dirs:
                if abs(nrm @ d) < 1e-12:
                    out_b.append(d / np.linalg.norm(d)); out_n.append(nrm / np.linalg.norm(nrm))
    bi, ni = np.array(out_b), np.array(out_n)
    assert bi.shape == (48, 3)
    return bi, ni


def make_crystal_type(e=69000.0, nu=0.33, mu=None, elastic="isotropic", slip="fcc", harden_n=20.0,
                      tau_y=0.0, tau_v=0.0, voche_m=1.0, theta_o=100.0, k_o=0.0, b=2.87e-7,
                      iD_v=0.0, eps_dot_o_y=1.0e10, atol=1.0e-5, atol1=1.0e-5, rtol=5.0e-5,
                      rtol1=1.0e-5, miter=30, alter_mode=False):
    """One crystal definition = c_array(cnum) after finalize_new_crystal
    (src/mod_crystals.f:238-417 defaults, :421-1934 derived data).  Returns a
    (CRYSTAL_TYPE_DOUBLES,) float64 record."""
    if slip == "fcc":
        bi, ni = fcc_slip_systems()
    elif slip == "bcc":            # slip_type 2 (src/mod_crystals.f:519-596): the fcc table with
        ni, bi = fcc_slip_systems()    # directions and normals exchanged, {110}<111>, 12 systems
    elif slip == "bcc48":          # 48 systems: CPU restatement only (the engine builds <= 12)
        bi, ni = bcc48_slip_systems()
    else:
        raise NotImplementedError("slip tables built: fcc, bcc, bcc48")
    nslip = bi.shape[0]
    if mu is None:
        mu = e / 2.0 / (1.0 + nu)
    flex = np.zeros((6, 6))
    flex[:3, :3] = -nu / e
    flex[0, 0] = flex[1, 1] = flex[2, 2] = 1.0 / e
    sh = 2.0 * (1.0 + nu) / e if elastic == "isotropic" else 1.0 / mu     # elastic_type 1 / 2 (cubic)
    flex[3, 3] = flex[4, 4] = flex[5, 5] = sh
    stiff = np.linalg.inv(flex)
    stiff = 0.5 * (stiff + stiff.T)
    rec = np.zeros(CRYSTAL_TYPE_DOUBLES)
    rec[0:36] = stiff.reshape(-1)
    B = np.zeros((48, 3)); N = np.zeros((48, 3))
    B[:nslip] = bi; N[:nslip] = ni
    rec[36:180] = B.reshape(-1)
    rec[180:324] = N.reshape(-1)
    rec[324:337] = [harden_n, tau_y, tau_v, voche_m, theta_o, k_o, b, iD_v, eps_dot_o_y, atol, atol1,
                    rtol, rtol1]
    rec[337:344] = [nslip, 1, 1, 0, miter, 1, 1 if alter_mode else 0]   # nslip, num_hard, h_type, tang_calc, miter, solver, alter_mode
    return rec


def mm10_hist_size(nslip, ncrystals, num_hard=1):
    """history words per Gauss point, mm10_set_history_locs (src/mm10_d.f:153-330)."""
    use_max = nslip == 48 or num_hard == 48
    ls = 48 if use_max else nslip
    lt = 48 if use_max else num_hard
    lu = 48 if use_max else 15
    return (36 + 27 + 9 + 3 + ls) + ncrystals * (30 + ls + lt + lu + lt + 12)


def attach_crystals(model: Model, crystal_types, ncrystals=1, seed=20261017, angles=None,
                    elem_cnum=None, angle_type=1):
    """Give every element of an mm10 model `ncrystals` crystals: crystal numbers (1-based into
    crystal_types) and Kocks angles in degrees drawn U(0,360) x U(0,180) x U(0,360).
    Also sets history_blk_list for the mm10 blocks."""
    ct = np.atleast_2d(np.asarray(crystal_types, dtype=np.float64))
    model.crystal_types = np.ascontiguousarray(ct)
    model.ncrystals = int(ncrystals)
    model.angle_type = int(angle_type)
    if angles is None:
        rng = np.random.default_rng(seed)
        angles = np.stack([rng.uniform(0, 360, (model.noelem, ncrystals)),
                           rng.uniform(0, 180, (model.noelem, ncrystals)),
                           rng.uniform(0, 360, (model.noelem, ncrystals))], axis=2)
    model.elem_angles = np.ascontiguousarray(np.asarray(angles, dtype=np.float64).reshape(model.noelem, ncrystals, 3))
    if elem_cnum is None:
        elem_cnum = np.ones((model.noelem, ncrystals), dtype=np.int32)
    model.elem_cnum = np.ascontiguousarray(np.asarray(elem_cnum, dtype=np.int32).reshape(model.noelem, ncrystals))
    nslip = int(ct[0, 337])
    for b in range(model.nelblk):
        f = model.felem(b)
        if int(model.props[f - 1, 24]) == 10:
            model.hist_size[b] = mm10_hist_size(nslip, ncrystals)
    model.asymmetric_assembly = False
    return model
