"""Host-side mirror of the reference's block drivers, bound to the C ABI in include/w3d_engine.h.

The method names follow the reference subroutines they stand in for
(``drive_eps_sig_internal_forces`` src/drive_eps_sig_internal_forces.f:20, ``tanstf``
src/tanstf.f:17, ``update`` src/update.f:75-93, ``adaptive_save`` / ``adaptive_reset``
src/mnralg.f:1258-1420).  Everything numerical happens in ``libw3d_engine.so`` (hand-written
sm_100a CUDA); this module only marshals the reference-layout arrays through ctypes exactly as a
Fortran ISO_C_BINDING caller would.  There is no CPU fallback: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .model import MXELPR, MXMTPR, MXVL, Model, State

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("W3D_ENGINE_LIB") or os.path.join(_HERE, "libw3d_engine.so")
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 3
(HIST_N, HIST_N1, EPS_N, EPS_N1, URCS_N, URCS_N1, ROT_N, ROT_N1, CEP, ESTIFF, ELEM_VOL) = range(11)
_STATE_WHICH = dict(history=HIST_N, history1=HIST_N1, eps_n=EPS_N, eps_n1=EPS_N1, urcs_n=URCS_N,
                    urcs_n1=URCS_N1, rot_n=ROT_N, rot_n1=ROT_N1, cep=CEP, estiff=ESTIFF,
                    element_vol=ELEM_VOL)

EXPORTS = (
    "w3d_engine_create", "w3d_engine_destroy", "w3d_engine_last_error", "w3d_eps_sig_ifv",
    "w3d_tanstf", "w3d_update", "w3d_adaptive_save", "w3d_adaptive_reset", "w3d_push_state",
    "w3d_pull_state", "w3d_block_offset", "w3d_device_vectors", "w3d_eps_sig_ifv_resident",
    "w3d_tanstf_resident", "w3d_last_scalars", "w3d_engine_stream", "w3d_engine_synchronize",
    "w3d_engine_launch_count", "w3d_upload_displacements", "w3d_download_ifv",
    "w3d_download_estiff", "w3d_set_time_step",
    "w3d_assemble_setup", "w3d_assemble_structure", "w3d_assemble", "w3d_assemble_rows", "w3d_assemble_csr",
    "w3d_residual_set_loads", "w3d_residual", "w3d_residual_device",
    "w3d_comm_unique_id", "w3d_comm_init", "w3d_comm_set_options", "w3d_comm_plan", "w3d_comm_owned",
    "w3d_download_ifv_range",
)


class _ResidualLoads(C.Structure):
    """w3d_residual_loads (include/w3d_engine.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("cstmap", "pbar", "load", "mdiag", "velocity", "accel",
                                           "contact_force", "d_lagrange_forces")]


class _ResidualNorms(C.Structure):
    """w3d_residual_norms (include/w3d_engine.h)"""
    _fields_ = [("norm_res", C.c_double), ("mgload", C.c_double), ("sum_loads", C.c_double),
                ("num_term_loads", C.c_int64), ("res_max", C.c_double), ("res_max_dof", C.c_int32)]


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"w3d_engine error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("mxvl", C.c_int32), ("mxelpr", C.c_int32),
        ("mxmtpr", C.c_int32), ("nelblk", C.c_int32), ("noelem", C.c_int32),
        ("nonode", C.c_int32), ("nodof", C.c_int32), ("myid", C.c_int32),
        ("numprocs", C.c_int32), ("elblks", C.c_void_p), ("props", C.c_void_p),
        ("matprp", C.c_void_p), ("nmat", C.c_int32), ("c", C.c_void_p), ("edest", C.c_void_p),
        ("cdest", C.c_void_p), ("history_blk_list", C.c_void_p), ("eps_bbar", C.c_double),
        ("qbar_flag", C.c_int32), ("beta_fact", C.c_double), ("adaptive_flag", C.c_int32),
        ("device", C.c_int32),
        ("crystal_types", C.c_void_p), ("n_crystal_types", C.c_int32), ("cp_ncrystals", C.c_int32),
        ("elem_cnum", C.c_void_p), ("elem_angles", C.c_void_p), ("cp_angle_type", C.c_int32),
        ("asymmetric_assembly", C.c_int32),
    ]


class _Scalars(C.Structure):
    _fields_ = [("internal_energy", C.c_double), ("plastic_work", C.c_double),
                ("sum_ifv", C.c_double), ("num_term_ifv", C.c_int64),
                ("material_cut_step", C.c_int32), ("reserved", C.c_int32)]


def build(force=False, verbose=False):
    """Compile libw3d_engine.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", "Makefile"))]
    srcs.append(os.path.join(_HERE, "..", "include", "w3d_engine.h"))
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libw3d_engine.so failed")
    return LIB_PATH


_lib = None


def lib():
    """Load the engine library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                  "(there is no CPU fallback for this path)")
        L = C.CDLL(LIB_PATH)
        L.w3d_engine_last_error.restype = C.c_char_p
        L.w3d_engine_last_error.argtypes = [C.c_void_p]
        L.w3d_block_offset.restype = C.c_int64
        L.w3d_block_offset.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.w3d_engine_launch_count.restype = C.c_int64
        L.w3d_engine_launch_count.argtypes = [C.c_void_p]
        L.w3d_engine_stream.restype = C.c_void_p
        L.w3d_engine_stream.argtypes = [C.c_void_p]
        for name in ("w3d_engine_create", "w3d_engine_destroy", "w3d_eps_sig_ifv", "w3d_tanstf",
                     "w3d_update", "w3d_adaptive_save", "w3d_adaptive_reset", "w3d_push_state",
                     "w3d_pull_state", "w3d_device_vectors", "w3d_eps_sig_ifv_resident",
                     "w3d_tanstf_resident", "w3d_last_scalars", "w3d_engine_synchronize",
                     "w3d_upload_displacements", "w3d_download_ifv", "w3d_download_estiff"):
            getattr(L, name).restype = C.c_int32
        L.w3d_engine_create.argtypes = [C.c_void_p, C.c_void_p]
        L.w3d_engine_destroy.argtypes = [C.c_void_p]
        L.w3d_eps_sig_ifv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.w3d_tanstf.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        for name in ("w3d_update", "w3d_adaptive_save", "w3d_adaptive_reset",
                     "w3d_engine_synchronize"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.w3d_push_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.w3d_pull_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.w3d_device_vectors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.w3d_eps_sig_ifv_resident.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.w3d_tanstf_resident.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.w3d_last_scalars.argtypes = [C.c_void_p, C.c_void_p]
        L.w3d_upload_displacements.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.w3d_download_ifv.argtypes = [C.c_void_p, C.c_void_p]
        L.w3d_download_estiff.argtypes = [C.c_void_p, C.c_void_p]
        L.w3d_set_time_step.restype = C.c_int32
        L.w3d_set_time_step.argtypes = [C.c_void_p, C.c_double]
        for name in ("w3d_comm_unique_id", "w3d_comm_init", "w3d_comm_set_options", "w3d_comm_plan",
                     "w3d_comm_owned", "w3d_download_ifv_range"):
            getattr(L, name).restype = C.c_int32
        L.w3d_comm_unique_id.argtypes = [C.c_void_p]
        L.w3d_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.w3d_comm_set_options.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.w3d_comm_plan.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.w3d_comm_owned.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.w3d_download_ifv_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def make_config(m: Model, device: int = -1, myid: int = 0, numprocs: int = 1):
    """w3d_engine_config for a model (the tables a Fortran caller would pass after its *_init
    routines).  Returns (cfg, keep-alive list of the arrays the pointers refer to)."""
    keep = [np.ascontiguousarray(x) for x in
            (m.elblks.astype(np.int32), m.props.astype(np.int32), m.matprp.astype(np.float32),
             m.c.astype(np.float64), m.edest.astype(np.int32), m.cdest.astype(np.int32),
             m.hist_size.astype(np.int32))]
    cfg = _Config()
    cfg.abi_version = ABI_VERSION
    cfg.mxvl, cfg.mxelpr, cfg.mxmtpr = max(MXVL, int(m.elblks[:, 0].max())), MXELPR, MXMTPR
    cfg.nelblk, cfg.noelem, cfg.nonode, cfg.nodof = m.nelblk, m.noelem, m.nonode, m.nodof
    cfg.myid, cfg.numprocs = myid, numprocs
    (cfg.elblks, cfg.props, cfg.matprp, cfg.c, cfg.edest, cfg.cdest,
     cfg.history_blk_list) = [_ptr(x) for x in keep]
    cfg.nmat = keep[2].shape[0]
    cfg.eps_bbar = m.eps_bbar
    cfg.qbar_flag = int(m.qbar_flag)
    cfg.beta_fact = m.beta_fact
    cfg.adaptive_flag = int(m.adaptive_flag)
    cfg.device = device
    if getattr(m, "crystal_types", None) is not None:      # mm10 input tables
        ct = np.ascontiguousarray(m.crystal_types, dtype=np.float64)
        cn = np.ascontiguousarray(m.elem_cnum, dtype=np.int32)
        ang = np.ascontiguousarray(m.elem_angles, dtype=np.float64)
        keep += [ct, cn, ang]
        cfg.crystal_types, cfg.elem_cnum, cfg.elem_angles = _ptr(ct), _ptr(cn), _ptr(ang)
        cfg.n_crystal_types, cfg.cp_ncrystals = ct.shape[0], int(m.ncrystals)
        cfg.cp_angle_type = int(m.angle_type)
        cfg.asymmetric_assembly = int(m.asymmetric_assembly)
    return cfg, keep


def comm_plan(model: Model, rank: int, numprocs: int):
    """w3d_comm_plan (host only, no GPU): the dofs domain `rank` shares with every other domain,
    as a list of 1-based ascending int32 arrays indexed by rank."""
    L = lib()
    cfg, keep = make_config(model, -1, rank, numprocs)
    counts = np.zeros(numprocs, dtype=np.int64)
    rc = L.w3d_comm_plan(C.byref(cfg), C.c_int32(rank), _ptr(counts), None)
    if rc:
        raise EngineError(rc, L.w3d_engine_last_error(None).decode())
    dofs = np.zeros(max(int(counts.sum()), 1), dtype=np.int32)
    rc = L.w3d_comm_plan(C.byref(cfg), C.c_int32(rank), _ptr(counts), _ptr(dofs))
    if rc:
        raise EngineError(rc, L.w3d_engine_last_error(None).decode())
    off = np.concatenate([[0], np.cumsum(counts)])
    return [dofs[off[r]:off[r + 1]].copy() for r in range(numprocs)]


def comm_owned(model: Model, rank: int, numprocs: int):
    """w3d_comm_owned: boolean mask of the dofs `rank` owns (lowest rank whose blocks touch it)."""
    L = lib()
    cfg, keep = make_config(model, -1, rank, numprocs)
    owned = np.zeros(model.nodof, dtype=np.uint8)
    rc = L.w3d_comm_owned(C.byref(cfg), C.c_int32(rank), _ptr(owned))
    if rc:
        raise EngineError(rc, "w3d_comm_owned failed")
    return owned.astype(bool)


def comm_unique_id():
    """w3d_comm_unique_id (rank 0): 128 bytes to hand to every rank's comm_init."""
    L = lib()
    buf = (C.c_uint8 * 128)()
    rc = L.w3d_comm_unique_id(buf)
    if rc:
        raise EngineError(rc, L.w3d_engine_last_error(None).decode())
    return bytes(buf)


class Engine:
    """One engine instance = one GPU = one MPI rank's share of the element blocks."""

    def __init__(self, model: Model, device: int = -1, myid: int = 0, numprocs: int = 1):
        self.model = model
        self._L = lib()
        m = model
        cfg, keep = make_config(m, device, myid, numprocs)
        h = C.c_void_p()
        rc = self._L.w3d_engine_create(C.byref(cfg), C.byref(h))
        if rc:
            raise EngineError(rc, self._L.w3d_engine_last_error(None).decode())
        self._h = h
        self._check(self._L.w3d_set_time_step(self._h, float(m.dt)))
        self.myid, self.numprocs = myid, numprocs
        owned = np.ones(m.nelblk, dtype=bool) if numprocs <= 1 else (m.elblks[:, 2] == myid)
        self.owned_blocks = np.nonzero(owned)[0]
        self.n_owned_elems = int(m.elblks[owned, 0].sum())

    # ---- plumbing
    def _check(self, rc):
        if rc:
            raise EngineError(rc, self._L.w3d_engine_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.w3d_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._L.w3d_engine_launch_count(self._h))

    @property
    def stream(self):
        return int(self._L.w3d_engine_stream(self._h) or 0)

    def synchronize(self):
        self._check(self._L.w3d_engine_synchronize(self._h))

    # ---- exchange step between domains (NCCL behind the C ABI)
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._L.w3d_comm_init(self._h, buf, C.c_int32(rank), C.c_int32(nranks)))
        self.comm_active = True

    def comm_set_options(self, exchange="neighbours", bcast_inputs=False):
        mode = {"neighbours": 1, "interface": 1, "full": 0, "allreduce": 0}[exchange]
        self._check(self._L.w3d_comm_set_options(self._h, C.c_int32(mode), C.c_int32(int(bcast_inputs))))

    def download_ifv_range(self, first, n, out=None):
        """ifv(first .. first+n-1), first 1-based."""
        if out is None:
            out = np.empty(n)
        self._check(self._L.w3d_download_ifv_range(self._h, C.c_int32(first), C.c_int32(n), _ptr(out)))
        return out

    # ---- the reference's drivers
    def set_time_step(self, dt):
        self._check(self._L.w3d_set_time_step(self._h, float(dt)))

    def drive_eps_sig_internal_forces(self, step, itr, u, du, ifv=None):
        """Strain recovery, material update, internal forces for all owned blocks.
        Returns (ifv, scalars); ifv is this rank's contribution."""
        u = None if u is None else _f64(u)       # None: only with a communicator (see w3d_engine.h)
        du = None if du is None else _f64(du)
        if ifv is None:
            ifv = np.empty(self.model.nodof)
        elif ifv is False:                       # communicator: this rank does not want ifv back
            ifv = None
        sc = _Scalars()
        rc = self._L.w3d_eps_sig_ifv(self._h, step, itr, int(self.model.extrapolated_du), _ptr(u),
                                     _ptr(du), _ptr(ifv), C.byref(sc))
        self._check(rc)
        return ifv, self._scalars(sc)

    @staticmethod
    def _scalars(sc):
        return dict(internal_energy=sc.internal_energy, plastic_work=sc.plastic_work,
                    sum_ifv=sc.sum_ifv, num_term_ifv=sc.num_term_ifv,
                    material_cut_step=bool(sc.material_cut_step))

    def tanstf(self, step, itr, u=None, du=None, out=None, to_host=True):
        """Element tangent stiffnesses (300, span) per owned block, blocks concatenated."""
        if u is not None:
            u, du = _f64(u), _f64(du)
        if to_host and out is None:
            out = np.empty(self.n_owned_elems * 300)
        rc = self._L.w3d_tanstf(self._h, step, itr, _ptr(u), _ptr(du),
                                _ptr(out) if to_host else None)
        self._check(rc)
        return out

    # ---- assembled equilibrium equations (assemble_code.f), symmetric, one domain
    def assemble_setup(self, cstmap):
        """dof_map + sparsity of the upper triangle for the constraint map `cstmap` (nodof ints,
        0 = free).  Returns dict(neqns, ncoeff, dof_eqn_map, eqn_node_map, k_ptrs, k_indexes)."""
        cst = np.ascontiguousarray(cstmap, dtype=np.int32)
        ed = np.ascontiguousarray(self.model.edest, dtype=np.int32)
        neq, nco = C.c_int32(), C.c_int64()
        self._check(self._L.w3d_assemble_setup(self._h, _ptr(cst), _ptr(ed), C.byref(neq), C.byref(nco)))
        first, nloc = C.c_int32(), C.c_int32()
        self._check(self._L.w3d_assemble_rows(self._h, C.byref(first), C.byref(nloc)))
        # one domain: first_eqn = 1 and n_eqns = neqns; several: this domain's row range
        self._neqns, self._ncoeff = nloc.value, nco.value
        out = dict(neqns=neq.value, ncoeff=nco.value, first_eqn=first.value, n_eqns=nloc.value,
                   dof_eqn_map=np.empty(self.model.nodof, dtype=np.int32),
                   eqn_node_map=np.empty(neq.value, dtype=np.int32),
                   k_ptrs=np.empty(nloc.value, dtype=np.int32),
                   k_indexes=np.empty(max(nco.value, 1), dtype=np.int32))
        self._check(self._L.w3d_assemble_structure(self._h, _ptr(out["dof_eqn_map"]), _ptr(out["eqn_node_map"]),
                                                   _ptr(out["k_ptrs"]), _ptr(out["k_indexes"])))
        out["k_indexes"] = out["k_indexes"][:nco.value]
        return out

    def assemble(self, k_diag=None, k_coeffs=None, to_host=True):
        """assem_by_row on the element stiffnesses of the last tanstf pass -> (k_diag, k_coeffs)."""
        if to_host:
            if k_diag is None:
                k_diag = np.empty(self._neqns)
            if k_coeffs is None:
                k_coeffs = np.empty(max(self._ncoeff, 1))
            self._check(self._L.w3d_assemble(self._h, _ptr(k_diag), _ptr(k_coeffs)))
            return k_diag, k_coeffs[:self._ncoeff]
        self._check(self._L.w3d_assemble(self._h, None, None))
        return None, None

    def assemble_csr(self, values=True):
        """convert_vss_csr: (k_ptrs (neqns+1, 1-based), k_indexes (nnz, 1-based), k_coeffs (nnz) or
        None) -- the full structurally symmetric CSR of the last assemble()."""
        nnz = C.c_int64()
        self._check(self._L.w3d_assemble_csr(self._h, C.byref(nnz), None, None, None))
        ptrs = np.empty(self._neqns + 1, dtype=np.int32)
        idx = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value) if values else None
        self._check(self._L.w3d_assemble_csr(self._h, C.byref(nnz), _ptr(ptrs), _ptr(idx), _ptr(val)))
        return ptrs, idx, val

    # ---- structure residual and its norms (upres.f, cvtest.f:62-78)
    def residual_set_loads(self, cstmap, pbar, load=None, mdiag=None, velocity=None, accel=None,
                           contact_force=None, d_lagrange_forces=None):
        """Copies the step's load vectors to the device (None = identically zero)."""
        keep = [np.ascontiguousarray(cstmap, dtype=np.int32)]
        keep += [None if v is None else _f64(v)
                 for v in (pbar, load, mdiag, velocity, accel, contact_force, d_lagrange_forces)]
        for v in keep:
            if v is not None and v.size != self.model.nodof:
                raise ValueError("residual vectors must have nodof entries")
        arg = _ResidualLoads(*[None if v is None else v.ctypes.data for v in keep])
        self._check(self._L.w3d_residual_set_loads(self._h, C.byref(arg)))

    def residual(self, iter, dt=1.0, nbeta=0.25, to_host=True):
        """upres_iter_0 (iter 0) or upres (iter >= 1) + the cvtest residual loop on the internal
        forces of the last stress pass.  Returns (res or None, dict of norms)."""
        res = np.empty(self.model.nodof) if to_host else None
        out = _ResidualNorms()
        self._check(self._L.w3d_residual(self._h, C.c_int32(iter), C.c_double(dt), C.c_double(nbeta),
                                         _ptr(res) if to_host else None, C.byref(out)))
        return res, dict(norm_res=out.norm_res, mgload=out.mgload, sum_loads=out.sum_loads,
                         num_term_loads=out.num_term_loads, res_max=out.res_max,
                         res_max_dof=out.res_max_dof)

    def update(self):
        self._check(self._L.w3d_update(self._h))

    def adaptive_save(self):
        self._check(self._L.w3d_adaptive_save(self._h))

    def adaptive_reset(self):
        self._check(self._L.w3d_adaptive_reset(self._h))

    # ---- device-resident passes (bench, NCCL layer)
    def device_vectors(self):
        pu, pdu, pf = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.w3d_device_vectors(self._h, C.byref(pu), C.byref(pdu), C.byref(pf)))
        return pu.value, pdu.value, pf.value

    def upload_displacements(self, u, du):
        u = None if u is None else _f64(u)
        du = None if du is None else _f64(du)
        self._check(self._L.w3d_upload_displacements(self._h, _ptr(u), _ptr(du)))

    def download_ifv(self, ifv=None):
        if ifv is None:
            ifv = np.empty(self.model.nodof)
        self._check(self._L.w3d_download_ifv(self._h, _ptr(ifv)))
        return ifv

    def download_estiff(self, out=None):
        if out is None:
            out = np.empty(self.n_owned_elems * 300)
        self._check(self._L.w3d_download_estiff(self._h, _ptr(out)))
        return out

    def eps_sig_ifv_resident(self, step, itr):
        self._check(self._L.w3d_eps_sig_ifv_resident(self._h, step, itr,
                                                     int(self.model.extrapolated_du)))

    def tanstf_resident(self, step, itr):
        self._check(self._L.w3d_tanstf_resident(self._h, step, itr))

    def last_scalars(self):
        sc = _Scalars()
        self._check(self._L.w3d_last_scalars(self._h, C.byref(sc)))
        return self._scalars(sc)

    # ---- blocked state in the reference layouts
    def block_offset(self, which, blk1):
        return int(self._L.w3d_block_offset(self._h, which, blk1))

    def _owned_offsets(self, name):
        """offsets of every block inside the model-wide concatenated array `name` (State)."""
        m = self.model
        per = dict(history=8 * m.hist_size.astype(np.int64), history1=8 * m.hist_size.astype(np.int64),
                   eps_n=48, eps_n1=48, urcs_n=72, urcs_n1=72, rot_n=72, rot_n1=72, cep=168,
                   estiff=300, element_vol=1)[name]
        return m.block_offsets(per)

    def push_state(self, state: State, names=None):
        for name in names or State.NAMES:
            if name == "estiff":
                continue
            a = getattr(state, name)
            off = self._owned_offsets(name)
            for blk in self.owned_blocks:
                seg = a[off[blk]:off[blk + 1]]
                if seg.size:
                    self._check(self._L.w3d_push_state(self._h, _STATE_WHICH[name], int(blk) + 1,
                                                       _ptr(seg)))

    def pull_block(self, name, blk):
        """One block (0-based) of the state array `name` in the reference's layout, flattened."""
        off = self._owned_offsets(name)
        out = np.empty(int(off[blk + 1] - off[blk]))
        self._check(self._L.w3d_pull_state(self._h, _STATE_WHICH[name], int(blk) + 1, _ptr(out)))
        return out

    def pull_state(self, state: State, names=None):
        for name in names or State.NAMES:
            a = getattr(state, name)
            off = self._owned_offsets(name)
            for blk in self.owned_blocks:
                seg = a[off[blk]:off[blk + 1]]
                if seg.size:
                    self._check(self._L.w3d_pull_state(self._h, _STATE_WHICH[name], int(blk) + 1,
                                                       _ptr(seg)))
        return state
