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


def build_smoke_run():
    exe = os.path.join(HOST, "smoke_run")
    src = os.path.join(HOST, "smoke_run.cpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["/usr/bin/g++", "-O2", src, "-I" + os.path.join(ROOT, "include"), "-L" + os.path.dirname(HOST), "-lsmoke_b200",
                        "-Wl,-rpath,$ORIGIN/..", "-o", exe], check=True)
    return exe


def parse_scene(path):
    """The runner's scene format, parsed independently for the oracle side."""
    sc = dict(grid=(80, 80, 80), gravity=-9.82, buoyancy=2.0, iterations=30, ticks=20, dt0=0.01, dt=0.05, objs=[], moves=[])
    for line in open(path):
        t = line.split("#")[0].split()
        if not t:
            continue
        k, v = t[0], t[1:]
        if k == "grid":
            sc["grid"] = tuple(int(x) for x in v)
        elif k in ("gravity", "buoyancy", "dt0", "dt"):
            sc[k] = float(np.float32(v[0]))
        elif k in ("iterations", "ticks"):
            sc[k] = int(v[0])
        elif k in ("source", "obstacle"):
            sc["objs"].append((k, [float(np.float32(x)) for x in v]))
        elif k == "move":
            sc["moves"].append((int(v[0]), int(v[1]), [float(np.float32(x)) for x in v[2:]]))
    return sc


@pytest.mark.parametrize("scene,ticks", [("ragged_moving.scene", None), ("C1", 12)])
def test_smoke_run_dumps_match_oracle_field_by_field(po, tmp_path, scene, ticks):
    """the C++ runner replays a scene file (ragged grid, overlapping obstacles, moving objects) and a built-in scene
    with fixed ticks and dumps raw fields; every dumped field -- density, u, v, w, mask -- must equal the oracle driven
    through the same calls, element by element."""
    exe = build_smoke_run()
    path = os.path.join(ROOT, "smoke-simulation_b200", "scenes_txt", scene) if scene.endswith(".scene") else scene
    out = tmp_path / "dump"
    cmd = [exe, "--scene", path, "--dump", str(out), "--dump-every", "4"] + (["--ticks", str(ticks)] if ticks else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sc = parse_scene(path) if scene.endswith(".scene") else parse_scene(os.path.join(ROOT, "smoke-simulation_b200", "scenes_txt", scene + ".scene"))
    if ticks:
        sc["ticks"] = ticks
    W, H, D = sc["grid"]
    o = po.Oracle(W, H, D, contract=1)
    o.set_params(sc["gravity"], sc["buoyancy"]); o.set_iterations(sc["iterations"])
    for kind, v in sc["objs"]:
        if kind == "source":
            o.add_source(*v)
        else:
            o.add_obstacle(v[0], v[1], v[2], 0.0, 0.0, 0.0, v[3])

    def check(d, what):
        assert np.array_equal(np.fromfile(d / "mask.u8", dtype=np.uint8).reshape(D, H, W), o.get_field(po.MASK)), what + " mask"
        assert np.array_equal(np.fromfile(d / "density.f32", dtype=np.float32).reshape(D, H, W), o.get_field(po.SMOKE, po.PAST)), what + " density"
        for name, f in (("u", po.U), ("v", po.V), ("w", po.W)):
            got = np.fromfile(d / (name + ".f32"), dtype=np.float32).reshape(D + 1, H + 1, W + 1)
            assert np.array_equal(got, o.get_field(f, po.NOW)), f"{what} {name}"

    for t in range(sc["ticks"]):
        for (oid, tick, pos) in sc["moves"]:
            if tick == t:
                o.update_object_pos(oid, *pos)
        o.step(sc["dt0"] if t == 0 else sc["dt"])
        if (t + 1) % 4 == 0 and t != sc["ticks"] - 1:
            check(out / f"tick_{t + 1}", f"tick {t + 1}")
    check(out, "final")
    import json
    meta = json.load(open(out / "meta.json"))
    assert meta["grid"] == [W, H, D] and meta["ticks_done"] == sc["ticks"]
    assert np.float32(meta["max_abs_divergence"]) == np.float32(o.max_divergence())   # (%.9g round-trips a binary32 value)
