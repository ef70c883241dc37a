"""The deterministic synthetic scenes of SURVEY.md section 8(d) and the tick schedule of the reference application
(first tick dt = 0.01, main.cpp:293; then 1/20 s, main.cpp:93).  Shared by bench.py, the headless runs and the tests;
every engine with the reference's scene calls (``set_params``, ``add_source``, ``add_obstacle``) can be set up from them.
"""

SCENES = {
    # name: (W, H, D, gravity, alpha, sources[(x,y,z,r)], obstacles[(x,y,z,r)])
    "C1": (80, 80, 80, -9.82, 2.0, [(40, 40, 40, 5)], [(60, 10, 60, 13)]),          # main.cpp:87, 288-291
    "C2": (256, 256, 256, -9.82, 15.0, [(128, 32, 128, 16)], []),
    "C3": (512, 512, 512, -9.82, 15.0, [(256, 64, 256, 32)], [(256, 192, 256, 48)]),
    "C4": (1024, 1024, 1024, 9.82, 2.0, [(512, 128, 512, 64)], []),
}


def scaled_scene(name, n):
    """Scene `name` shrunk to an n^3 grid (positions / radii scaled, rounded to integers)."""
    W_, H_, D_, g, a, src, obs = SCENES[name]
    f = n / W_
    sc = lambda t: tuple(float(round(v * f)) for v in t[:3]) + (float(max(2, round(t[3] * f))),)
    return (n, n, n, g, a, [sc(t) for t in src], [sc(t) for t in obs])


def setup_scene(engine, scene):
    """addSmokeSource / addObstacle in the order of main.cpp:288-291 (sources first); returns the object ids."""
    _, _, _, g, a, src, obs = scene
    engine.set_params(g, a)
    ids = []
    for (x, y, z, r) in src:
        ids.append(engine.add_source(x, y, z, r))
    for (x, y, z, r) in obs:
        ids.append(engine.add_obstacle(x, y, z, 0.0, 0.0, 0.0, r))
    return ids


def tick_dt(t):
    """tick 0 uses dt = 0.01 (main.cpp:293), later ticks dt = 0.05 (= 1/20 s, main.cpp:93)."""
    return 0.01 if t == 0 else 0.05
