"""The C-ABI library builds, loads without a GPU, and exports every symbol include/w3d_engine.h
declares (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "w3d_engine.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(w3d_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from warp3d_b200 import engine
    engine.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in w3d_engine.h but not exported"
    assert set(engine.EXPORTS) == set(names)


def test_create_fails_loudly_without_gpu_or_with_bad_config():
    """No CPU fallback: without a usable device create returns W3D_ERR_CUDA; with a device a
    NULL config table returns W3D_ERR_ARG.  Either way a non-zero code and a message."""
    from warp3d_b200 import engine
    L = engine.lib()
    cfg = engine._Config()
    cfg.abi_version = engine.ABI_VERSION
    h = ctypes.c_void_p()
    rc = L.w3d_engine_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert L.w3d_engine_last_error(None)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "warp3d_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                hit = re.search(r"(import\s+oracle|from\s+oracle|oracle_py|libw3d_oracle|w3do_|w3d_oracle)", src)
                assert not hit, f"{f} reaches into oracle/: {hit.group(0)}"


def test_oracle_never_includes_product_source():
    """The checker is independent of what it checks: no file under oracle/ includes, imports or
    names anything of the engine package (mm05 and mm10 have their own statement-for-statement
    restatements, oracle/w3d_oracle_mm05.h and w3d_oracle_mm10.h)."""
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".py", ".cpp", ".h", ".c")) or f == "Makefile":
            src = open(os.path.join(ROOT, "oracle", f)).read()
            assert "warp3d_b200" not in src and "csrc/" not in src.replace("(csrc/", "("), f
