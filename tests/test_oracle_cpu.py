"""CPU tests (no GPU): the oracle restatement against the reference's own kernel bodies and against the
committed golden fixtures generated from them (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, all_fields, inject, random_state, rel_err

SMALL_SCENE = (20, 18, 16, -9.82, 2.0, [(10.0, 8.0, 8.0, 3.0)], [(13.0, 4.0, 9.0, 3.0), (6.0, 6.0, 5.0, 2.5)])
SMALL_RANDOM = (17, 21, 19, -9.82, 6.0, [(8.0, 9.0, 9.0, 2.5)], [])


def run_oracle(po, scene, ticks, contract, state=None, cls=None):
    e = (cls or po.Oracle)(*scene[:3]) if cls else po.Oracle(*scene[:3], contract=contract)
    po.setup_scene(e, scene)
    if state is not None:
        inject(po, e, state)
    for t in range(ticks):
        e.step(po.tick_dt(t))
    return e


def assert_bit_equal(a, b, what):
    for k in a:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), f"{what}: field {k} differs"


@pytest.mark.parametrize("name,scene,rand", [("small_scene", SMALL_SCENE, False), ("small_random", SMALL_RANDOM, True)])
def test_oracle_matches_golden_bit_exact(po, name, scene, rand):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    st = random_state(po, *scene[:3]) if rand else None
    e = run_oracle(po, scene, int(g["ticks"]), contract=0, state=st)
    got = all_fields(po, e)
    for k in got:
        assert np.array_equal(got[k].view(np.uint8), g[k].view(np.uint8)), f"{name}: {k}"


def test_oracle_c1_matches_golden(po):
    """80^3 default scene, 20 ticks: mask bit-exact (15 058 solid cells), density bit-exact (contract=0)."""
    g = np.load(os.path.join(GOLDEN, "c1_80.npz"))
    sc = po.SCENES["C1"]
    e = run_oracle(po, sc, int(g["ticks"]), contract=0)
    mask = e.get_field(po.MASK)
    assert np.array_equal(np.packbits(mask), g["mask_bits"])
    assert int((mask == 0).sum()) == 15058  # 6 400 floor + 8 658 sphere
    assert np.array_equal(e.get_field(po.SMOKE, po.PAST), g["density"])
    u, v, w = (e.get_field(f, po.NOW) for f in (po.U, po.V, po.W))
    assert np.array_equal(u[:, 40, :], g["u_y40"]) and np.array_equal(v[40], g["v_z40"]) and np.array_equal(w[:, 40, :], g["w_y40"])
    assert abs(e.max_divergence() - float(g["maxdiv"])) == 0.0
    assert abs(float(e.get_field(po.SMOKE, po.PAST).sum()) - 631.56) < 0.01  # survey-time sanity value


def test_contracted_mode_within_tolerance_of_golden(po):
    """contract=1 mirrors nvcc's FMA contraction: it must stay within the 1e-5 parity tolerance of the
    uncontracted reference bodies on the default scene."""
    g = np.load(os.path.join(GOLDEN, "c1_80.npz"))
    e = run_oracle(po, po.SCENES["C1"], int(g["ticks"]), contract=1)
    assert np.array_equal(np.packbits(e.get_field(po.MASK)), g["mask_bits"])
    assert rel_err(e.get_field(po.SMOKE, po.PAST), g["density"]) <= 1e-5
    u, v, w = (e.get_field(f, po.NOW) for f in (po.U, po.V, po.W))
    assert rel_err(u[40], g["u_z40"]) <= 1e-5 and rel_err(v[:, 40, :], g["v_y40"]) <= 1e-5 and rel_err(w[40], g["w_z40"]) <= 1e-5


def test_oracle_vs_reference_bodies_live(po):
    """Where oracle/_ref/libref_cpu.so is present: restatement == reference kernel bodies, stage by stage,
    on random fields with a random mask (bit-exact), including ragged non-cubic sizes."""
    if not po.have_ref_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    for (W, H, D) in [(17, 21, 19), (8, 9, 33), (3, 3, 3), (4, 5, 3)]:
        scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.0)], [])
        st = random_state(po, W, H, D, seed=7)
        a = po.Oracle(W, H, D, contract=0); b = po.RefCPU(W, H, D)
        for e in (a, b):
            po.setup_scene(e, scene); inject(po, e, st)
        dt = 0.05
        for stage in ("flip", "fill", "integrate", "clamp", "p0", "p1", "p0", "advect_velocity", "advect_smoke"):
            for e in (a, b):
                if stage in ("flip", "fill"):
                    getattr(e, stage)()
                elif stage[0] == "p":
                    e.pressure_halfsweep(int(stage[1]))
                else:
                    getattr(e, stage)(dt)
            assert_bit_equal(all_fields(po, a), all_fields(po, b), f"{W}x{H}x{D} after {stage}")


def test_last_obstacle_wins_and_moving_objects(po):
    """fillObstacle rewrites every interior cell per obstacle (cu:304-310): only the last obstacle survives;
    updateObjectPos moves it (cu:106-109)."""
    W = H = D = 16
    e = po.Oracle(W, H, D)
    e.add_obstacle(5, 5, 5, 0, 0, 0, 2.5)
    i2 = e.add_obstacle(10, 10, 10, 0, 0, 0, 2.5)
    e.step(0.01)
    m = e.get_field(po.MASK)
    assert m[5, 5, 5] == 1 and m[10, 10, 10] == 0
    e.update_object_pos(i2, 6, 6, 6)
    e.step(0.05)
    m = e.get_field(po.MASK)
    assert m[10, 10, 10] == 1 and m[6, 6, 6] == 0
    assert (m[:, 0, :] == 0).all() and m[0, 1, 0] == 1  # floor solid, outer shell otherwise fluid


def test_empty_scene_is_identity(po):
    """No sources, no obstacles, zero fields: everything stays zero and the mask stays floor-only."""
    e = po.Oracle(9, 7, 5)
    for t in range(3):
        e.step(po.tick_dt(t))
    f = all_fields(po, e)
    assert all(float(np.abs(f[k]).max()) == 0.0 for k in f if k != "mask")
    assert int((f["mask"] == 0).sum()) == 9 * 5


def test_pressure_reduces_divergence(po):
    W = H = D = 24
    st = random_state(po, W, H, D, seed=3)
    st["mask"][:] = 1; st["mask"][:, 0, :] = 0
    e = po.Oracle(W, H, D); inject(po, e, st); e.flip()
    d0 = e.max_divergence()
    for i in range(30):
        e.pressure_halfsweep(0); e.pressure_halfsweep(1)
    d1 = e.max_divergence()
    assert d1 < 0.05 * d0, (d0, d1)


def test_jacobi_extension_reduces_divergence_and_is_local(po):
    """The Jacobi extension (no reference counterpart) is specified by the oracle: check what a pressure iteration must
    do -- the divergence residual falls, solid faces and boundary faces never move -- and that the whole step runs."""
    W, H, D = 20, 18, 16
    st = random_state(po, W, H, D, seed=4)
    e = po.Oracle(W, H, D)
    inject(po, e, st)
    e.flip()
    before = {k: e.get_field(getattr(po, k.upper()), po.NOW).copy() for k in ("u", "v", "w")}
    d0 = e.max_divergence()
    for _ in range(40):
        e.jacobi_iteration()
    assert e.max_divergence() < 0.1 * d0
    m = st["mask"].reshape(D, H, W)
    u = e.get_field(po.U, po.NOW).reshape(D + 1, H + 1, W + 1)
    ub = before["u"].reshape(D + 1, H + 1, W + 1)
    solid_face = (m[:, :, 1:] == 0) | (m[:, :, :-1] == 0)  # u face between cells x-1 | x
    assert np.array_equal(u[:D, :H, 1:W][solid_face], ub[:D, :H, 1:W][solid_face])
    assert np.array_equal(u[:, :, 0], ub[:, :, 0]) and np.array_equal(u[:, :, W], ub[:, :, W])
    f = po.Oracle(16, 16, 16)
    po.setup_scene(f, po.scaled_scene("C1", 16))
    f.set_solver(1, 40)
    for t in range(3):
        f.step(po.tick_dt(t))
    assert np.isfinite(f.get_field(po.SMOKE, po.NOW)).all()


def test_union_obstacle_extension_is_the_or_of_single_obstacle_masks(po):
    """N3 extension, specified by the oracle: in union mode a cell is solid iff it is solid for ANY single obstacle
    (each evaluated with the reference's own last-wins rule, which for one obstacle is just 'inside')."""
    W, H, D = 22, 18, 16
    obs = [(7, 9, 8, 3.0), (13, 9, 8, 3.5), (11, 5, 6, 2.0)]

    def mask_for(spheres, union):
        e = po.Oracle(W, H, D)
        for s in spheres:
            e.add_obstacle(s[0], s[1], s[2], 0, 0, 0, s[3])
        e.set_obstacle_mode(union)
        e.step(0.01)
        return e.get_field(po.MASK)

    singles = [mask_for([s], 0) for s in obs]
    want = np.minimum.reduce(singles)
    assert np.array_equal(mask_for(obs, 1), want)
    assert not np.array_equal(mask_for(obs, 0), want)  # the reference rule really differs on this scene


def test_binary32_omega_product_host_twin_sampled(tmp_path):
    """The CPU twin of the pressure pass's binary32 evaluation of `(float)((double)q * -1.9)` (cu:384;
    tools/experiments/omega_fp32_exhaustive.c) on every 251st bit pattern -- the full 2^32 sweep takes a minute on 8 cores and
    runs on the device in the GPU suite (smk_selfcheck_omega).  Exit status 0 = no input in the valid range differs from
    the double-precision product; the tie-to-even rule must fire on 1/19 of the inputs, and the constants printed are the ones
    the kernel uses."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "omega_check")
    subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", exe,
                    os.path.join(root, "tools", "experiments", "omega_fp32_exhaustive.c"), "-lm"], check=True)
    r = subprocess.run([exe, "251"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "different from the reference: 0 " in r.stdout
    assert "0xbff33333" in r.stdout and "0xb2ccc8cd" in r.stdout and "0xb2ccd0cd" in r.stdout
    kernel = open(os.path.join(root, "smoke-simulation_b200", "csrc", "kernels_pressure_tma.cuh")).read()
    for c in ("-0x1.e66666p+0f", "-0x1.99919ap-26f", "-0x1.99a19ap-26f"):
        assert c in kernel
    ties = float(r.stdout.split("ties broken to even): ")[1].split("= ")[1].split(" %")[0])
    assert 4.5 < ties < 4.8
