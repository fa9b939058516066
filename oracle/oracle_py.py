"""ctypes binding for the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# W3D_ORACLE_STATS=1 selects the build with mm10 operation counters (w3d_mm10_core.h, W3D_STAT)
_STATS = bool(os.environ.get("W3D_ORACLE_STATS"))
# W3D_ORACLE_FAST=1 selects the -O3 / FMA build used ONLY as the timed CPU baseline (bench.py)
_FAST = bool(os.environ.get("W3D_ORACLE_FAST")) and not _STATS
_LIB = os.path.join(_HERE, "libw3d_oracle_stats.so" if _STATS else
                    "libw3d_oracle_fast.so" if _FAST else "libw3d_oracle.so")


class _Model(C.Structure):
    _fields_ = [
        ("nelblk", C.c_int32), ("noelem", C.c_int32), ("nonode", C.c_int32), ("nodof", C.c_int32),
        ("elblks", C.c_void_p), ("props", C.c_void_p), ("matprp", C.c_void_p), ("nmat", C.c_int32),
        ("c", C.c_void_p), ("edest", C.c_void_p), ("cdest", C.c_void_p), ("hist_size", C.c_void_p),
        ("eps_bbar", C.c_double), ("qbar_flag", C.c_int32), ("dt", C.c_double),
        ("total_model_time", C.c_double), ("beta_fact", C.c_double),
        ("extrapolated_du", C.c_int32), ("adaptive_flag", C.c_int32),
        ("fpconstant_data", C.c_int32), ("cp", C.c_void_p),
    ]


class _Cp(C.Structure):
    _fields_ = [("ncrystals", C.c_int32), ("nslip", C.c_int32), ("num_hard", C.c_int32),
                ("asymmetric_assembly", C.c_int32), ("props", C.c_void_p)]


class _State(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("history", "history1", "eps_n", "eps_n1", "urcs_n", "urcs_n1", "rot_n", "rot_n1",
                 "cep", "estiff", "element_vol")]


class _Scalars(C.Structure):
    _fields_ = [("internal_energy", C.c_double), ("plastic_work", C.c_double),
                ("sum_ifv", C.c_double), ("num_term_ifv", C.c_int64),
                ("material_cut_step", C.c_int32)]


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.run(["make", "-C", _HERE, os.path.basename(_LIB)], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.w3do_drive_eps_sig_internal_forces.restype = C.c_int32
        _lib.w3do_tanstf.restype = C.c_int32
        _lib.w3do_block_offset.restype = C.c_int64
        _lib.w3do_mm10_props_doubles.restype = C.c_int32
        _lib.w3do_mm10_hist_size.restype = C.c_int32
    return _lib


def mm10_stats(reset=False):
    """mm10 operation counters of the stats build: slip evaluations, Jacobians, dense solves,
    crystal updates that ran the solver, sub-increment solves."""
    if not _STATS:
        raise RuntimeError("set W3D_ORACLE_STATS=1 before importing the oracle")
    arr = (C.c_longlong * 8).in_dll(lib(), "w3d_mm10_stats")
    out = dict(zip(("slip_evals", "jacobians", "solves", "crystal_solves", "subincrement_solves",
                    "pivoted_solves"), list(arr)[:6]))
    if reset:
        for k in range(8):
            arr[k] = 0
    return out


def mm10_props(model):
    """per (element, crystal) property records = setup_mm10_rknstr (rknstr.f:1568-2027)."""
    L = lib()
    nslip_max = int(np.asarray(model.crystal_types)[:, 337].max())
    L.w3do_mm10_props_doubles_for.restype = C.c_int32
    pd = L.w3do_mm10_props_doubles_for(C.c_int32(nslip_max))
    n = model.noelem * model.ncrystals
    out = np.zeros((n, pd))
    cn = np.ascontiguousarray(model.elem_cnum.reshape(-1), dtype=np.int32)
    ang = np.ascontiguousarray(model.elem_angles.reshape(-1), dtype=np.float64)
    ct = np.ascontiguousarray(model.crystal_types, dtype=np.float64)
    L.w3do_mm10_setup_props_batch(_ptr(ct), _ptr(cn), _ptr(ang), C.c_int64(n),
                                  C.c_int32(model.angle_type), _ptr(out))
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Runs the oracle on the engine package's Model + State tables (arrays are used in place)."""

    def __init__(self, model, nthreads=1):
        self.model = model
        self.nthreads = int(nthreads)
        m = model
        self._keep = [np.ascontiguousarray(x) for x in
                      (m.elblks, m.props, m.matprp, m.c, m.edest, m.cdest, m.hist_size)]
        self._cm = _Model()
        self._cp = None
        if getattr(m, "crystal_types", None) is not None:
            self._cp_props = mm10_props(m)
            self._cp = _Cp()
            self._cp.ncrystals = m.ncrystals
            self._cp.nslip = int(m.crystal_types[0, 337])
            self._cp.num_hard = int(m.crystal_types[0, 338])
            self._cp.asymmetric_assembly = int(m.asymmetric_assembly)
            self._cp.props = _ptr(self._cp_props)
        self.sync_globals()

    def sync_globals(self):
        m, cm = self.model, self._cm
        cm.nelblk, cm.noelem, cm.nonode, cm.nodof = m.nelblk, m.noelem, m.nonode, m.nodof
        (cm.elblks, cm.props, cm.matprp, cm.c, cm.edest, cm.cdest,
         cm.hist_size) = [_ptr(x) for x in self._keep]
        cm.nmat = self._keep[2].shape[0]
        cm.eps_bbar = m.eps_bbar
        cm.qbar_flag = int(m.qbar_flag)
        cm.dt = m.dt
        cm.total_model_time = m.total_model_time
        cm.beta_fact = m.beta_fact
        cm.extrapolated_du = int(m.extrapolated_du)
        cm.adaptive_flag = int(m.adaptive_flag)
        cm.fpconstant_data = 1
        cm.cp = C.cast(C.pointer(self._cp), C.c_void_p) if self._cp is not None else None

    @staticmethod
    def _cstate(st):
        cs = _State()
        for n in st.NAMES:
            a = getattr(st, n)
            assert a.dtype == np.float64 and a.flags.c_contiguous
            setattr(cs, n, _ptr(a))
        return cs

    def eps_sig_ifv(self, state, step, itr, u, du):
        """drive_eps_sig_internal_forces: returns (ifv, scalars dict)."""
        self.sync_globals()
        ifv = np.zeros(self.model.nodof)
        sc = _Scalars()
        cs = self._cstate(state)
        u = np.ascontiguousarray(u, dtype=np.float64)
        du = np.ascontiguousarray(du, dtype=np.float64)
        rc = lib().w3do_drive_eps_sig_internal_forces(
            C.byref(self._cm), C.byref(cs), C.c_int32(step), C.c_int32(itr), _ptr(u), _ptr(du),
            _ptr(ifv), C.byref(sc), C.c_int32(self.nthreads))
        if rc:
            raise RuntimeError(f"oracle fatal condition {rc}")
        return ifv, dict(internal_energy=sc.internal_energy, plastic_work=sc.plastic_work,
                         sum_ifv=sc.sum_ifv, num_term_ifv=sc.num_term_ifv,
                         material_cut_step=bool(sc.material_cut_step))

    def tanstf(self, state, step, itr, u, du):
        self.sync_globals()
        cs = self._cstate(state)
        u = np.ascontiguousarray(u, dtype=np.float64)
        du = np.ascontiguousarray(du, dtype=np.float64)
        rc = lib().w3do_tanstf(C.byref(self._cm), C.byref(cs), C.c_int32(step), C.c_int32(itr),
                               _ptr(u), _ptr(du), C.c_int32(self.nthreads))
        if rc:
            raise RuntimeError(f"oracle fatal condition {rc}")
        return state.estiff

    def update(self, state):
        cs = self._cstate(state)
        lib().w3do_update(C.byref(self._cm), C.byref(cs))

    def adaptive_save(self, state):
        cs = self._cstate(state)
        lib().w3do_adaptive_save(C.byref(self._cm), C.byref(cs))

    def adaptive_reset(self, state):
        cs = self._cstate(state)
        lib().w3do_adaptive_reset(C.byref(self._cm), C.byref(cs))


def quad2(gpn):
    v = [C.c_double() for _ in range(4)]
    lib().w3do_quad2(C.c_int32(gpn), *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def deriv2(xi, eta, zeta):
    out = [np.zeros(8) for _ in range(3)]
    lib().w3do_deriv2(C.c_double(xi), C.c_double(eta), C.c_double(zeta), *[_ptr(x) for x in out])
    return out


def cp_icp(n):
    utsz = (n * n + n) // 2
    cp = np.zeros(n, dtype=np.int32)
    dcp = np.zeros(n, dtype=np.int32)
    icp = np.zeros((2, utsz), dtype=np.int32)   # Fortran (utsz, 2)
    lib().w3do_cp_icp(C.c_int32(n), _ptr(cp), _ptr(dcp), _ptr(icp))
    return cp, dcp, icp


# ---- assembled equilibrium equations (assemble_code.f), symmetric ---------------------------------
def assemble_structure(oracle: "Oracle", cstmap):
    """dof_map, count_profile_symmetric and build_col_sparse_symm for the constraint map `cstmap`
    (nodof ints, 0 = free).  Returns dict(neqns, ncoeff, dof_eqn_map, eqn_node_map, k_ptrs, k_indexes)."""
    L, m = lib(), oracle.model
    cst = np.ascontiguousarray(cstmap, dtype=np.int32)
    dmap = np.zeros(m.nodof, dtype=np.int32)
    emap = np.zeros(m.nodof, dtype=np.int32)
    L.w3do_dof_map.restype = C.c_int32
    L.w3do_assemble_structure.restype = C.c_int64
    neq = L.w3do_dof_map(_ptr(cst), C.c_int32(m.nonode), C.c_int32(3), _ptr(dmap), _ptr(emap))
    emap = np.ascontiguousarray(emap[:neq])
    kp = np.zeros(neq, dtype=np.int32)
    oracle.sync_globals()
    nco = L.w3do_assemble_structure(C.byref(oracle._cm), C.c_int32(neq), _ptr(emap), _ptr(dmap), _ptr(kp), None)
    ki = np.zeros(max(nco, 1), dtype=np.int32)
    nco2 = L.w3do_assemble_structure(C.byref(oracle._cm), C.c_int32(neq), _ptr(emap), _ptr(dmap), _ptr(kp), _ptr(ki))
    assert nco2 == nco
    return dict(neqns=int(neq), ncoeff=int(nco), dof_eqn_map=dmap, eqn_node_map=emap, k_ptrs=kp,
                k_indexes=ki[:nco])


def assemble(oracle: "Oracle", estiff, s):
    """assem_by_row: (k_diag, k_coeffs) from the packed element stiffness blocks and a structure
    returned by assemble_structure."""
    L = lib()
    ek = np.ascontiguousarray(estiff, dtype=np.float64).reshape(-1)
    kd = np.zeros(s["neqns"])
    kc = np.zeros(max(s["ncoeff"], 1))
    ki = np.ascontiguousarray(s["k_indexes"] if s["ncoeff"] else np.zeros(1, dtype=np.int32))
    oracle.sync_globals()
    L.w3do_assemble(C.byref(oracle._cm), _ptr(ek), C.c_int32(s["neqns"]), _ptr(s["eqn_node_map"]),
                    _ptr(s["dof_eqn_map"]), _ptr(s["k_ptrs"]), _ptr(ki), _ptr(kd), _ptr(kc))
    return kd, kc[:s["ncoeff"]]


def convert_vss_csr(s):
    """convert_vss_csr (drive_assemble_solve.f:1501): (ptrs (neqns+1), indexes (nnz)), 1-based."""
    L = lib()
    L.w3do_convert_vss_csr.restype = C.c_int64
    neq, nco = s["neqns"], s["ncoeff"]
    ptrs = np.zeros(neq + 1, dtype=np.int32)
    idx = np.zeros(neq + 2 * nco, dtype=np.int32)
    ki = np.ascontiguousarray(s["k_indexes"] if nco else np.zeros(1, dtype=np.int32))
    nnz = L.w3do_convert_vss_csr(C.c_int32(neq), C.c_int64(nco), _ptr(s["k_ptrs"]), _ptr(ki), _ptr(ptrs), _ptr(idx))
    assert nnz == neq + 2 * nco
    return ptrs, idx


# ---- structure residual and its norms (upres.f, cvtest.f:62-78) -----------------------------------
def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def upres_iter_0(cstmap, pbar, ifv, du, dt=1.0, nbeta=0.25, mdiag=None, contact_force=None):
    """upres_iter_0 (upres.f:13): returns (res, norm_load)."""
    L = lib()
    cst = np.ascontiguousarray(cstmap, dtype=np.int32)
    pbar, ifv, du, mdiag, cf = map(_f64, (pbar, ifv, du, mdiag, contact_force))
    res = np.zeros(cst.size)
    L.w3do_upres_iter_0.restype = C.c_double
    nl = L.w3do_upres_iter_0(C.c_int32(cst.size), C.c_double(dt), C.c_double(nbeta), _ptr(cst), _ptr(pbar),
                             _ptr(ifv), _ptr(mdiag), _ptr(du), _ptr(cf), _ptr(res))
    return res, float(nl)


def upres(cstmap, pbar, load, ifv, du, dt=1.0, nbeta=0.25, mdiag=None, velocity=None, accel=None,
          contact_force=None, d_lagrange_forces=None):
    """upres (upres.f:121): returns (res, dict(mgload, sum_loads, num_term_loads))."""
    L = lib()
    cst = np.ascontiguousarray(cstmap, dtype=np.int32)
    pbar, load, ifv, du, mdiag, vel, acc, cf, dl = map(
        _f64, (pbar, load, ifv, du, mdiag, velocity, accel, contact_force, d_lagrange_forces))
    res = np.zeros(cst.size)
    out2 = np.zeros(2)
    L.w3do_upres.restype = C.c_int64
    n = L.w3do_upres(C.c_int32(cst.size), C.c_double(dt), C.c_double(nbeta), _ptr(cst), _ptr(pbar), _ptr(load),
                     _ptr(ifv), _ptr(mdiag), _ptr(du), _ptr(vel), _ptr(acc), _ptr(cf), _ptr(dl), _ptr(res),
                     _ptr(out2))
    return res, dict(mgload=float(out2[0]), sum_loads=float(out2[1]), num_term_loads=int(n))


def cvtest_residual(res, d_lagrange_forces=None):
    """The residual loop of cvtest (cvtest.f:62-78): dict(norm_res, res_max, res_max_dof)."""
    L = lib()
    res, dl = _f64(res), _f64(d_lagrange_forces)
    out2 = np.zeros(2)
    L.w3do_cvtest_residual.restype = C.c_int32
    dof = L.w3do_cvtest_residual(C.c_int32(res.size), _ptr(res), _ptr(dl), _ptr(out2))
    return dict(norm_res=float(out2[0]), res_max=float(out2[1]), res_max_dof=int(dof))
