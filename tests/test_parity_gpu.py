"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libsmoke_b200.so.  Checkers: oracle/liboracle.so (contract=1 = the FMA pattern nvcc applies to the
reference), the committed golden fixtures (reference kernel bodies, uncontracted) and -- when the
snapshot carries it -- the reference step itself built headless for sm_100a (oracle/_ref/libref_gpu.so).

Bars: mask bit-exact; fields max|a-b|/max|b| <= 1e-5; against the contracted oracle and
the GPU reference the kernels are written to be bit-identical, which is asserted as exact equality."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, all_fields, inject, random_state, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5
SMALL_SCENE = (20, 18, 16, -9.82, 2.0, [(10.0, 8.0, 8.0, 3.0)], [(13.0, 4.0, 9.0, 3.0), (6.0, 6.0, 5.0, 2.5)])
SMALL_RANDOM = (17, 21, 19, -9.82, 6.0, [(8.0, 9.0, 9.0, 2.5)], [])


def make_pair(po, smk, scene, state=None, contract=1):
    W, H, D = scene[:3]
    a = smk.SmokeSim(W, H, D)
    b = po.Oracle(W, H, D, contract=contract)
    for e in (a, b):
        po.setup_scene(e, scene)
        if state is not None:
            inject(po, e, state)
    return a, b


def compare(po, a, b, what, exact=True):
    fa, fb = all_fields(po, a), all_fields(po, b)
    assert np.array_equal(fa["mask"], fb["mask"]), f"{what}: mask differs in {(fa['mask'] != fb['mask']).sum()} cells"
    worst = 0.0
    for k in fa:
        if k == "mask":
            continue
        if exact:
            bad = fa[k] != fb[k]
            assert not bad.any(), f"{what}: {k} differs in {int(bad.sum())} entries, rel {rel_err(fa[k], fb[k]):.3e}"
        else:
            e = rel_err(fa[k], fb[k])
            worst = max(worst, e)
            assert e <= TOL, f"{what}: {k} rel err {e:.3e} > {TOL}"
    return worst


@pytest.mark.parametrize("dims", [(17, 21, 19), (8, 9, 33), (3, 3, 3), (4, 5, 3), (33, 16, 10), (64, 20, 12)])
def test_stage_by_stage_vs_oracle_random_fields(po, smk, dims):
    """Per-kernel parity on random u/v/w/density and a random mask, ragged non-cubic sizes included."""
    W, H, D = dims
    scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.0)], [])
    st = random_state(po, W, H, D, seed=11)
    a, b = make_pair(po, smk, scene, st)
    dt = 0.05
    a.flip(); b.flip()
    a.fill(); b.fill()
    compare(po, a, b, f"{dims} fill")
    a.force_clamp(dt); b.integrate(dt); b.clamp(dt)
    compare(po, a, b, f"{dims} force+clamp")
    for off in (0, 1, 0, 1):
        a.pressure_halfsweep(off); b.pressure_halfsweep(off)
        compare(po, a, b, f"{dims} halfsweep {off}")
    a.advect_velocity(dt); b.advect_velocity(dt)
    compare(po, a, b, f"{dims} advect velocity")
    a.advect_smoke(dt); b.advect_smoke(dt)
    compare(po, a, b, f"{dims} advect smoke")
    a.close()


def test_clamp_engages(po, smk):
    """Large velocities so that the max-velocity clamp (cu:331-352) actually fires."""
    W, H, D = 12, 10, 9
    st = random_state(po, W, H, D, seed=5)
    for k in ("u", "v", "w"):
        st[k] *= 20.0
    a, b = make_pair(po, smk, (W, H, D, -9.82, 2.0, [], []), st)
    a.flip(); b.flip(); a.fill(); b.fill()
    before = a.get_field(po.U, po.NOW).copy()
    a.force_clamp(0.05); b.integrate(0.05); b.clamp(0.05)
    assert (a.get_field(po.U, po.NOW) != before).any()
    compare(po, a, b, "clamp")
    a.close()


@pytest.mark.parametrize("scene,ticks,rand", [(SMALL_SCENE, 6, False), (SMALL_RANDOM, 3, True)])
def test_full_step_vs_oracle_and_golden(po, smk, scene, ticks, rand):
    st = random_state(po, *scene[:3]) if rand else None
    a, b = make_pair(po, smk, scene, st)
    host = np.zeros((scene[2], scene[1], scene[0]), dtype=np.float32)
    for t in range(ticks):
        a.step(po.tick_dt(t), host); b.step(po.tick_dt(t))
    compare(po, a, b, "full step vs oracle(contract=1)")
    assert np.array_equal(host, a.get_field(po.SMOKE, po.PAST)), "host readback != device density"
    g = np.load(os.path.join(GOLDEN, ("small_random" if rand else "small_scene") + ".npz"))
    fa = all_fields(po, a)
    assert np.array_equal(fa["mask"], g["mask"])
    for k in fa:
        if k != "mask":
            assert rel_err(fa[k], g[k]) <= TOL, (k, rel_err(fa[k], g[k]))
    a.close()


def test_c1_default_scene_20_ticks(po, smk):
    """BASELINE configs[0]: 80^3 default scene, 20 ticks, against the oracle (exact) and the golden fixture
    from the reference bodies (mask bit-exact, fields <= 1e-5, same final divergence residual)."""
    sc = po.SCENES["C1"]
    a, b = make_pair(po, smk, sc)
    for t in range(20):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "C1 vs oracle(contract=1)")
    g = np.load(os.path.join(GOLDEN, "c1_80.npz"))
    mask = a.get_field(po.MASK)
    assert np.array_equal(np.packbits(mask), g["mask_bits"]) and int((mask == 0).sum()) == 15058
    assert rel_err(a.get_field(po.SMOKE, po.PAST), g["density"]) <= TOL
    u, v, w = (a.get_field(f, po.NOW) for f in (po.U, po.V, po.W))
    for got, key in ((u[:, 40, :], "u_y40"), (v[:, 40, :], "v_y40"), (w[:, 40, :], "w_y40"), (u[40], "u_z40"), (v[40], "v_z40"), (w[40], "w_z40")):
        assert rel_err(got, g[key]) <= TOL, key
    res = a.max_divergence()
    assert abs(res - b.max_divergence()) == 0.0
    assert abs(res - float(g["maxdiv"])) <= 1e-5 * max(1.0, float(g["absmax"].max()))
    a.close()


def test_vs_reference_gpu_build(po, smk):
    """The reference step itself on the same GPU (oracle/_ref/libref_gpu.so): C1 for 20 ticks and a random
    state stage by stage.  Mask bit-exact, fields identical."""
    if not po.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not in the snapshot")
    sc = po.SCENES["C1"]
    a = smk.SmokeSim(*sc[:3]); po.setup_scene(a, sc)
    r = po.RefGPU(*sc[:3]); po.setup_scene(r, sc)
    try:
        for t in range(20):
            a.step(po.tick_dt(t)); r.step(po.tick_dt(t))
        compare(po, a, r, "C1 vs reference on GPU")
        assert np.array_equal(r.host, a.get_field(po.SMOKE, po.PAST))
    finally:
        r.close(); a.close()
    W, H, D = 33, 20, 17
    scene = (W, H, D, -9.82, 3.0, [(16.0, 10.0, 8.0, 3.0)], [])
    st = random_state(po, W, H, D, seed=21)
    a = smk.SmokeSim(W, H, D); po.setup_scene(a, scene); inject(po, a, st)
    r = po.RefGPU(W, H, D); po.setup_scene(r, scene); inject(po, r, st)
    try:
        a.flip(); r.flip(); a.fill(); r.fill()
        a.force_clamp(0.05); r.integrate(0.05); r.clamp(0.05)
        compare(po, a, r, "random: force+clamp vs reference GPU")
        for off in (0, 1, 0, 1):
            a.pressure_halfsweep(off); r.pressure_halfsweep(off)
        compare(po, a, r, "random: half-sweeps vs reference GPU")
        a.advect_velocity(0.05); r.advect_velocity(0.05)
        a.advect_smoke(0.05); r.advect_smoke(0.05)
        compare(po, a, r, "random: advection vs reference GPU")
    finally:
        r.close(); a.close()


def test_moving_obstacle_and_last_wins(po, smk):
    W = H = D = 16
    a = smk.SmokeSim(W, H, D); b = po.Oracle(W, H, D)
    ids = []
    for e in (a, b):
        e.add_source(8, 4, 8, 2.0)
        e.add_obstacle(5, 5, 5, 0, 0, 0, 2.5)
        ids.append(e.add_obstacle(10, 10, 10, 0, 0, 0, 2.5))
    assert ids[0] == ids[1] == 2
    for t in range(3):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    for e in (a, b):
        e.update_object_pos(2, 6, 9, 6)
    for t in range(3):
        a.step(0.05); b.step(0.05)
    compare(po, a, b, "moving obstacle")
    m = a.get_field(po.MASK)
    assert m[5, 5, 5] == 1 and m[10, 10, 10] == 1 and m[6, 9, 6] == 0
    a.close()


def test_sources_by_bounding_box_and_cached_mask(po, smk):
    """From the second step on the fill stamps the sources over their bounding boxes and reuses mask + stencil codes:
    spheres partly outside the grid, overlapping, zero / negative / huge radius, non-integer centres, a moving source,
    and a mask injected between steps (must be re-derived) -- all against the oracle, which fills the whole grid."""
    W, H, D = 30, 22, 18
    a = smk.SmokeSim(W, H, D); b = po.Oracle(W, H, D)
    srcs = [(-2.0, 5.0, 5.0, 4.0), (28.5, 20.25, 16.75, 3.5), (10.0, 10.0, 9.0, 0.0), (12.0, 8.0, 9.0, -2.0),
            (15.5, 11.5, 9.5, 2.0), (16.0, 12.0, 9.0, 2.5), (100.0, 100.0, 100.0, 3.0)]
    ids = []
    for e in (a, b):
        ids = [e.add_source(*s) for s in srcs]
        e.add_obstacle(20, 6, 9, 0, 0, 0, 3.0)
    for t in range(3):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "sources, step 3")
    for e in (a, b):
        e.update_object_pos(ids[4], 7.25, 15.0, 6.5)
    a.step(0.05); b.step(0.05)
    compare(po, a, b, "moved source")
    m = random_state(po, W, H, D, seed=8)["mask"]
    for e in (a, b):
        e.set_field(po.MASK, po.NOW, m)
    a.step(0.05); b.step(0.05)
    compare(po, a, b, "injected mask, obstacle list re-applied")
    big = smk.SmokeSim(W, H, D); bo = po.Oracle(W, H, D)
    for e in (big, bo):
        e.add_source(15.0, 11.0, 9.0, 500.0)
    for t in range(2):
        big.step(po.tick_dt(t)); bo.step(po.tick_dt(t))
    compare(po, big, bo, "source larger than the grid")
    a.close(); big.close()


def test_union_obstacles_and_uploaded_mask(po, smk):
    """SURVEY N3 extensions against the oracle: union semantics for overlapping / consecutive obstacles (the reference
    lets the last one decide), and an uploaded voxel mask that persists when there is no sphere obstacle."""
    W, H, D = 24, 20, 18
    a = smk.SmokeSim(W, H, D); b = po.Oracle(W, H, D)
    for e in (a, b):
        e.add_source(12, 5, 9, 3.0)
        e.add_obstacle(8, 10, 9, 0, 0, 0, 3.0)
        e.add_obstacle(15, 10, 9, 0, 0, 0, 3.0)
        e.set_obstacle_mode(1)
    for t in range(3):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "union obstacles")
    m = a.get_field(po.MASK)
    assert m[9, 10, 8] == 0 and m[9, 10, 15] == 0          # both spheres solid (last-wins would free the first)
    for e in (a, b):
        e.set_obstacle_mode(0)
    a.step(0.05); b.step(0.05)
    compare(po, a, b, "back to the reference semantics")
    assert a.get_field(po.MASK)[9, 10, 8] == 1
    a.close()
    # voxelised solid: no sphere obstacles, mask uploaded once
    a = smk.SmokeSim(W, H, D); b = po.Oracle(W, H, D)
    vox = np.ones((D, H, W), dtype=np.uint8); vox[:, 0, :] = 0; vox[6:12, 8:11, 5:19] = 0
    for e in (a, b):
        e.add_source(12, 4, 9, 3.0)
        e.set_field(po.MASK, po.NOW, vox)
    for t in range(4):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "uploaded voxel mask")
    assert np.array_equal(a.get_field(po.MASK), vox)
    a.close()


def test_half_precision_readback_is_the_rounded_density(po, smk):
    """binary16 readback == round-to-nearest-even of the float density (numpy's conversion)."""
    sc = po.scaled_scene("C1", 33)
    a, b = make_pair(po, smk, sc)
    for t in range(5):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    h = a.read_density_half()
    want = b.get_field(po.SMOKE, po.PAST).astype(np.float16)
    assert h.dtype == np.float16 and np.array_equal(h.view(np.uint16), want.view(np.uint16))
    assert float(h.astype(np.float32).max()) > 0.5
    a.close()


def test_empty_scene_and_parameters(po, smk):
    a = smk.SmokeSim(9, 7, 5)
    for t in range(2):
        a.step(po.tick_dt(t))
    f = all_fields(po, a)
    assert all(float(np.abs(f[k]).max()) == 0.0 for k in f if k != "mask")
    assert abs(a.gravity + 9.82) < 1e-6 and a.buoyancy == 2.0
    a.gravity = 3.0
    assert a.gravity == 3.0
    assert a.launch_count() > 0
    a.close()


def test_c2_256_one_tick_property_checks(po, smk):
    """BASELINE configs[1] size (256^3): two ticks against the oracle (exact), plus size-independent
    properties: density stays in [0,1], the residual drops, untouched boundary faces stay zero."""
    sc = po.SCENES["C2"]
    a, b = make_pair(po, smk, sc)
    for t in range(2):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "C2 256^3, 2 ticks")
    d = a.get_field(po.SMOKE, po.PAST)
    assert d.min() >= 0.0 and d.max() <= 1.0 + 1e-6  # trilinear weights sum to 1 within an ulp
    u = a.get_field(po.U, po.NOW)
    assert np.abs(u[:, :, 0]).max() == 0.0 and np.abs(u[:, :, -1]).max() == 0.0 and np.abs(u[0]).max() == 0.0
    a.close()


def test_c3_512_two_ticks_identical_to_reference_gpu_build(po, smk):
    """BASELINE configs[2] at full size (512^3, source + solid sphere): two ticks of the default schedule (fused passes,
    TMA advection, forcing riding on the first pass) against the reference step itself built for sm_100a -- mask
    bit-exact, every field identical, field by field to bound host memory -- plus size-independent properties."""
    if not po.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not in the snapshot")
    sc = po.SCENES["C3"]
    a = smk.SmokeSim(*sc[:3]); po.setup_scene(a, sc)
    r = po.RefGPU(*sc[:3]); po.setup_scene(r, sc)
    try:
        for t in range(2):
            a.step(po.tick_dt(t)); r.step(po.tick_dt(t))
        ma = a.get_field(po.MASK)
        assert np.array_equal(ma, r.get_field(po.MASK))
        assert int((ma == 0).sum()) > 512 * 512            # the floor plane and the sphere
        for f in (po.SMOKE, po.U, po.V, po.W):
            for which in (po.NOW, po.PAST):
                x, y = a.get_field(f, which), r.get_field(f, which)
                assert np.array_equal(x, y), (f, which, rel_err(x, y))
                if f == po.SMOKE:
                    assert x.min() >= 0.0 and x.max() <= 1.0 + 1e-6
                del x, y
        assert a.max_divergence() < 1.0
    finally:
        r.close(); a.close()


@pytest.mark.parametrize("fuse", [1, 2, 4])
@pytest.mark.parametrize("dims", [(20, 18, 16), (120, 50, 40), (57, 41, 9), (130, 100, 70)])
def test_fused_pressure_passes_identical(po, smk, dims, fuse):
    """Temporal blocking only reorders WHICH cell is updated WHEN, never the data dependencies of the
    red/black sweeps: K fused half-sweeps per launch must give the same bits as K launches (and as the
    oracle).  Sizes straddle several 56x40 output tiles and z-chunks; random mask and fields."""
    W, H, D = dims
    st = random_state(po, W, H, D, seed=3)
    scene = (W, H, D, -9.82, 3.0, [], [])
    a, b = make_pair(po, smk, scene, st)
    a.set_solver(0, 7, fuse)   # 14 half-sweeps: 3 passes of 4 + one of 2 (fuse=4), 7 of 2 (fuse=2)
    a.flip(); b.flip(); a.fill(); b.fill()
    a.pressure()
    for i in range(7):
        b.pressure_halfsweep(0); b.pressure_halfsweep(1)
    compare(po, a, b, f"{dims} fuse={fuse}")
    a.close()


@pytest.mark.parametrize("nctas", [0, 1, 3, 37, 148])
@pytest.mark.parametrize("dims,fuse", [((130, 100, 70), 4), ((57, 41, 9), 4), ((120, 50, 40), 2)])
def test_balanced_piece_lists_identical(po, smk, dims, fuse, nctas):
    """The schedule of a fused pass is free (csrc/pass_schedule.h): the (tile, z-chunk) grid (0, the default) and balanced piece
    lists for 1, 3, 37, 148 CTAs -- CTAs that work through several pieces, pieces of a few planes -- give the bits of
    separate half-sweep launches and of the oracle.  Random mask and fields; the last pass of 7 iterations has K = 2."""
    W, H, D = dims
    st = random_state(po, W, H, D, seed=5)
    scene = (W, H, D, -9.82, 3.0, [], [])
    a, b = make_pair(po, smk, scene, st)
    a.set_solver(0, 7, fuse)
    a.set_pass_ctas(nctas)
    a.flip(); b.flip(); a.fill(); b.fill()
    a.pressure()
    for i in range(7):
        b.pressure_halfsweep(0); b.pressure_halfsweep(1)
    compare(po, a, b, f"{dims} fuse={fuse} nctas={nctas}")
    a.close()


def test_balanced_full_steps_with_fused_forcing(po, smk):
    """Full steps (forcing + clamp riding on the first pass) with balanced piece lists on 5 CTAs vs the oracle."""
    sc = po.scaled_scene("C1", 64)
    a, b = make_pair(po, smk, sc)
    a.set_pass_ctas(5)
    for t in range(4):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "C1/64 balanced, 5 CTAs")
    a.close()


def test_fused_full_steps_c1_obstacle(po, smk):
    """C1 (80^3, obstacle) for 5 ticks with the default fused schedule (15 passes of 4) vs the oracle."""
    sc = po.SCENES["C1"]
    a, b = make_pair(po, smk, sc)
    a.set_solver(0, 30, 4)
    for t in range(5):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "C1 fused x4")
    a.close()


def test_pipelined_readback_delivers_every_step(po, smk):
    """smk_step_async with a host buffer: snapshot + copy on a second stream; after smk_sync the buffer holds the
    density of the last enqueued step, and intermediate steps were delivered in order."""
    import ctypes
    sc = po.scaled_scene("C1", 40)
    a, b = make_pair(po, smk, sc)
    host = np.zeros((40, 40, 40), dtype=np.float32)
    for t in range(5):
        a.step_async(po.tick_dt(t), host.ctypes.data_as(ctypes.c_void_p)); b.step(po.tick_dt(t))
        if t == 2:
            a.sync()
            assert np.array_equal(host, b.get_field(po.SMOKE, po.PAST))
    a.sync()
    assert np.array_equal(host, b.get_field(po.SMOKE, po.PAST))
    assert np.array_equal(host, a.get_field(po.SMOKE, po.PAST))
    a.close()


def test_density_to_cuda_array_without_host_round_trip(po, smk):
    """simulate(nullptr, dt) + device-to-array copy == what the reference uploads with glTexSubImage3D."""
    sc = po.scaled_scene("C1", 32)
    a, b = make_pair(po, smk, sc)
    L = smk.load_library()
    arr = L.smk_test_array_create(32, 32, 32)
    assert arr
    for t in range(4):
        a.step(po.tick_dt(t), None); b.step(po.tick_dt(t))
        a.copy_density_to_array(arr)
    a.sync()
    out = np.zeros((32, 32, 32), dtype=np.float32)
    assert L.smk_test_array_read(arr, out.ctypes.data_as(__import__("ctypes").c_void_p), 32, 32, 32) == 0
    assert np.array_equal(out, b.get_field(po.SMOKE, po.PAST))
    L.smk_test_array_destroy(arr)
    a.close()


# ---- extension: damped-Jacobi pressure iteration (no reference counterpart; specified by the oracle) --------------
@pytest.mark.parametrize("dims", [(24, 20, 16), (40, 33, 19), (3, 3, 3), (64, 64, 64), (128, 10, 6), (133, 9, 7), (260, 12, 5), (256, 17, 9)])
def test_jacobi_stage_matches_oracle(po, smk, dims):
    W, H, D = dims
    st = random_state(po, W, H, D, seed=11)
    a, b = make_pair(po, smk, (W, H, D, -9.82, 3.0, [], []), st)
    a.flip(); b.flip()
    a.set_solver(1, 7, 0)
    a.pressure()
    for _ in range(7):
        b.jacobi_iteration()
    compare(po, a, b, f"{dims} 7 jacobi iterations")
    a.close()


@pytest.mark.parametrize("nctas", [0, 1, 7, 296])
@pytest.mark.parametrize("dims", [(133, 41, 37), (260, 12, 5), (64, 64, 64)])
def test_jacobi_balanced_piece_lists_identical(po, smk, dims, nctas):
    """Jacobi iterations scheduled as a (tile, z-chunk) grid (0) or as balanced piece lists on 1 / 7 / 296 CTAs."""
    W, H, D = dims
    st = random_state(po, W, H, D, seed=12)
    a, b = make_pair(po, smk, (W, H, D, -9.82, 3.0, [], []), st)
    a.flip(); b.flip()
    a.set_solver(1, 5, 0)
    a.set_pass_ctas(nctas)
    a.pressure()
    for _ in range(5):
        b.jacobi_iteration()
    compare(po, a, b, f"{dims} jacobi nctas={nctas}")
    a.close()


def test_jacobi_tiny_values_match_oracle(po, smk):
    """Velocities in the denormal range (the decaying front of the iteration): the exact-quotient path for acc = 6."""
    W, H, D = 33, 18, 12
    st = random_state(po, W, H, D, seed=2)
    rng = np.random.default_rng(9)
    for k in ("u", "v", "w"):
        st[k] = (st[k] * np.float32(2.0) ** rng.integers(-149, -118, st[k].shape).astype(np.float32)).astype(np.float32)
    st["mask"][:] = 1
    a, b = make_pair(po, smk, (W, H, D, -9.82, 3.0, [], []), st)
    a.flip(); b.flip()
    a.set_solver(1, 3, 0)
    a.pressure()
    for _ in range(3):
        b.jacobi_iteration()
    compare(po, a, b, "jacobi, denormal-range fields")
    a.close()


def test_jacobi_full_step_matches_oracle(po, smk):
    sc = po.scaled_scene("C1", 48)
    a, b = make_pair(po, smk, sc)
    a.set_solver(1, 40, 0); b.set_solver(1, 40)
    for t in range(4):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "jacobi full step")
    assert a.max_divergence() == b.max_divergence()
    a.close()
