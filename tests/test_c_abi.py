"""The C ABI driven from plain C (tests/c_abi_smoke.c, gcc): the calls, argument order and array
layouts of the ISO_C_BINDING interface in fortran/mod_w3d_engine.f, with every table built in C
from scratch -- no Python host layer on that side.  CPU: it compiles and links against
libw3d_engine.so, and without a device w3d_engine_create fails loudly (W3D_ERR_CUDA, no fallback).
GPU: its check sums equal those of the same mesh driven through ctypes."""
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(ROOT, "build", "tests", "c_abi_smoke")


def _build():
    from warp3d_b200 import engine
    engine.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "warp3d_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(HERE, "c_abi_smoke.c"), "-o", EXE, "-L", libdir, "-lw3d_engine",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-lm"], check=True)
    return EXE


def _run(args):
    return subprocess.run([_build()] + [str(a) for a in args], capture_output=True, text=True, timeout=300)


def _gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_harness_builds_and_fails_loudly_without_a_device():
    r = _run([4, 3, 2, 7])
    if _gpu():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
        assert "no usable device" in r.stdout and len(r.stdout.strip()) > 30      # carries the engine's message


@pytest.mark.gpu
def test_c_harness_matches_the_ctypes_path():
    nx, ny, nz, bsz = 7, 5, 4, 13
    r = _run([nx, ny, nz, bsz])
    assert r.returncode == 0, r.stdout + r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("c_abi_smoke: noelem")][0]
    got = {k: float(v) for k, v in re.findall(r"(\w+) (-?[\d.]+(?:e[+-]?\d+)?)", line)}
    from warp3d_b200 import Material, build_structured_hex8
    from warp3d_b200.engine import Engine
    m = build_structured_hex8(nx, ny, nz, materials=[Material()], block_size=bsz)
    x = m.c.reshape(-1, 3)
    u = np.zeros(m.nodof)
    du = (x * np.array([3.0e-3, -0.9e-3, -0.9e-3])).reshape(-1)
    e = Engine(m)
    e.drive_eps_sig_internal_forces(1, 0, u, du)
    f, sc = e.drive_eps_sig_internal_forces(1, 1, u, du)
    ek = e.tanstf(1, 1)
    cst = np.zeros(m.nodof, dtype=np.int32)
    cst.reshape(-1, 3)[x[:, 0] == 0.0] = 1
    e.tanstf(1, 1, to_host=False)
    s = e.assemble_setup(cst)
    kd, kc = e.assemble()
    e.residual_set_loads(cst, pbar=np.zeros(m.nodof), load=np.zeros(m.nodof))
    _, nrm = e.residual(1)
    e.close()
    assert got["noelem"] == m.noelem and got["nelblk"] == m.nelblk and got["cut"] == 0
    assert got["num_term_ifv"] == sc["num_term_ifv"] and got["neqns"] == s["neqns"] and got["ncoeff"] == s["ncoeff"]
    assert got["nnz"] == s["neqns"] + 2 * s["ncoeff"]
    assert sc["plastic_work"] > 0
    for name, ref in (("sum_abs_ifv", np.abs(f).sum()), ("plastic_work", sc["plastic_work"]),
                      ("sum_abs_estiff", np.abs(ek).sum()), ("sum_kdiag", kd.sum()),
                      ("sum_abs_kcoeffs", np.abs(kc).sum()), ("norm_res", nrm["norm_res"])):
        assert abs(got[name] - ref) <= 1e-12 * abs(ref), (name, got[name], ref)
