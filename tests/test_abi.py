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


def test_a_caller_compiled_against_the_reference_header_links(smk, tmp_path):
    """A translation unit that includes the REFERENCE's own project/smokeSimulation.cuh (read where it lies; build
    container only) and calls all nine entry points the way main.cpp / boundingBox.cpp do links against
    libsmoke_b200.so with a plain host compiler: no source change on the caller's side."""
    ref_hdr_dir = "/root/reference/project"
    if not os.path.exists(os.path.join(ref_hdr_dir, "smokeSimulation.cuh")):
        pytest.skip("reference tree not present (GPU box)")
    src = tmp_path / "caller.cpp"
    src.write_text(
        '#include "smokeSimulation.cuh"\n'
        "int main(int argc, char**) {\n"
        "  if (argc > 100) {\n"                      # never executed: this is a link test (no GPU here)
        "    static float grid[8 * 8 * 8];\n"
        "    getGPUProperties();\n"
        "    initializeVolume(grid, 8u, 8u, 8u);\n"
        "    int a = addSmokeSource(4.f, 4.f, 4.f, 2.f);\n"
        "    int b = addObstacle(2.f, 2.f, 2.f, 0.f, 0.f, 0.f, 1.f);\n"
        "    updateObjectPos(a, 3.f, 3.f, 3.f); (void)b;\n"
        "    *getGravity() = -9.82f; *getBuoyancy() = 2.0f;\n"
        "    simulate(grid, 0.01f);\n"
        "    deleteVolume();\n"
        "  }\n"
        "  return 0;\n"
        "}\n")
    exe = tmp_path / "caller"
    libdir = os.path.dirname(smk.LIB_PATH)
    r = subprocess.run(["/usr/bin/g++", "-O1", "-x", "c++", str(src), "-I", ref_hdr_dir, "-L", libdir, "-lsmoke_b200",
                        "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)], capture_output=True).returncode == 0   # loads the library, calls nothing


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


def test_smoke_run_builds_against_the_c_abi_and_rejects_bad_scenes(tmp_path):
    """host/smoke_run.cpp is plain C++ over include/smoke_b200.h (no CUDA header); without a GPU it must fail
    loudly in smk_create, and a malformed scene file is reported with its line."""
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(ROOT, "smoke-simulation_b200", "host")
    exe = str(tmp_path / "smoke_run")
    subprocess.run(["/usr/bin/g++", "-O2", "-Wall", "-Werror", os.path.join(host, "smoke_run.cpp"), "-I" + os.path.join(ROOT, "include"),
                    "-L" + os.path.dirname(host), "-lsmoke_b200", "-Wl,-rpath," + os.path.dirname(host), "-o", exe], check=True)
    bad = tmp_path / "bad.scene"
    bad.write_text("grid 8 8 8\nsauce 1 2 3 4\n")
    r = subprocess.run([exe, "--scene", str(bad)], capture_output=True, text=True)
    assert r.returncode == 2 and "bad.scene:2" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "--scene", "C1", "--ticks", "1"], capture_output=True, text=True)
        assert r.returncode == 2 and "smk_create" in r.stderr
