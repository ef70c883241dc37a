"""GPU test of the drop-in boundary: the reference's C++ entry points (host/smokeSimulation.cuh) driven from plain
host C++ (host/headless_main.cpp, the application's call sequence), compared with the oracle driven the same way."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "smoke-simulation_b200", "host")


def build_headless():
    exe = os.path.join(HOST, "headless")
    src = os.path.join(HOST, "headless_main.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["/usr/bin/g++", "-O2", src, "-I" + HOST, "-L" + os.path.dirname(HOST), "-lsmoke_b200",
                        "-Wl,-rpath,$ORIGIN/..", "-o", exe], check=True)
    return exe


def test_headless_application_sequence_matches_oracle(po):
    n, ticks = 48, 14
    exe = build_headless()
    out = subprocess.run([exe, str(ticks)], env=dict(os.environ, SMOKE_N=str(n)), capture_output=True, text=True, check=True).stdout
    m = re.search(r"sum\(density\) = ([0-9.eE+-]+), max = ([0-9.eE+-]+)", out)
    assert m, out
    assert "CUDA Device - ID 0" in out  # getGPUProperties() printed the device block (cu:70-76)
    s = n / 80.0
    f32 = np.float32
    o = po.Oracle(n, n, n, contract=1)
    src = o.add_source(f32(40 * s), f32(40 * s), f32(40 * s), f32(5 * s))
    o.add_obstacle(f32(60 * s), f32(10 * s), f32(60 * s), 0, 0, 0, f32(13 * s))
    o.step(0.01)
    for t in range(1, ticks):
        if t == ticks // 2:
            o.set_params(-9.0, 4.0)
        if t > ticks // 2:
            o.update_object_pos(src, float(f32(40 * s) + f32(0.25) * f32(t - ticks // 2)), f32(40 * s), f32(40 * s))
        o.step(0.05)
    d = o.get_field(po.SMOKE, po.PAST).astype(np.float64)
    assert abs(float(m.group(1)) - d.sum()) <= 1e-3 * max(1.0, d.sum()) * 1e-2, (m.group(1), d.sum())
    assert abs(float(m.group(2)) - d.max()) <= 1e-6
