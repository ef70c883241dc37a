"""CPU tests (no GPU, no compute calls): the C-ABI library loads, exports every symbol include/smoke_b200.h
declares, exports the reference's nine C++ entry points, and fails loudly without a device."""
import ctypes
import os
import subprocess

import pytest


def test_library_exports_every_declared_symbol(smk):
    if not os.path.exists(smk.LIB_PATH):
        smk.build_library()
    L = smk.load_library()
    missing = [n for n in smk.declared_symbols() if not hasattr(L, n)]
    assert not missing, missing
    assert len(smk.declared_symbols()) >= 30
    assert L.smk_abi_version() == 1


def test_reference_entry_points_are_exported_with_reference_mangling(smk):
    """the nine Itanium-mangled symbols main.cpp / boundingBox.cpp link against."""
    want = ["_Z16getGPUPropertiesv", "_Z8simulatePff", "_Z16initializeVolumePfjjj", "_Z12deleteVolumev",
            "_Z11addObstaclefffffff", "_Z14addSmokeSourceffff", "_Z15updateObjectPosifff", "_Z11getBuoyancyv",
            "_Z10getGravityv"]
    L = ctypes.CDLL(smk.LIB_PATH)
    for n in want:
        assert hasattr(L, n), n
    # parameter pointers are stable and hold the reference defaults (cu:28-29)
    g = ctypes.cast(getattr(L, "_Z10getGravityv"), ctypes.CFUNCTYPE(ctypes.POINTER(ctypes.c_float)))()
    b = ctypes.cast(getattr(L, "_Z11getBuoyancyv"), ctypes.CFUNCTYPE(ctypes.POINTER(ctypes.c_float)))()
    assert abs(g[0] + 9.82) < 1e-6 and b[0] == 2.0


def test_product_does_not_reference_the_oracle(smk):
    """The product library must not link or embed anything from oracle/."""
    out = subprocess.run(["ldd", smk.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out
    syms = subprocess.run(["nm", "-D", smk.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms and "refcpu_" not in syms and "refgpu_" not in syms
    src_dir = os.path.join(os.path.dirname(smk.LIB_PATH))
    for root, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "pyoracle" not in txt and "smoke_oracle" not in txt and "liboracle" not in txt, f


def test_fails_loudly_without_gpu(smk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(smk.SmokeError):
        smk.SmokeSim(8, 8, 8)
