"""GPU parity tests added in round 2 for the gaps the round-1 review named (VERDICT.md "What's weak" 1-3):

* advection with LONG backtraces (|vel|*dt up to 6 cells): the staged-tile sampler's fallback to the global sampler,
  the box test at tile edges and the domain clamp far from the node -- reference: velocityAdvectionU/V/W cu:527-615,
  advectSmoke cu:617-638, sampleSmoke cu:451-484 (no bound on the backtrace length anywhere in the reference);
* the default simulate() path for grids with D >= 34: density advection in z-chunks, every chunk copied to the host
  while the next is computed (smk_api.cu enqueue_step) -- host buffer == oracle, bit for bit;
* long horizons: C1 x 100 ticks and C2 256^3 x 60 ticks (clamp engaged, developed plume) identical to the reference
  step itself built for sm_100a (oracle/_ref/libref_gpu.so);
* the host-buffer registration contract (smk_register_host / smk_unregister_host, re-validated cache).

Everything goes through the C ABI.  Equality is exact (`==`; +0 and -0 compare equal)."""
import numpy as np
import pytest

from conftest import all_fields, inject, random_state, rel_err
from test_parity_gpu import compare, make_pair

pytestmark = pytest.mark.gpu


def scaled_state(po, W, H, D, reach_cells, dt, seed):
    """random_state with u,v,w ~ U(-4,4) rescaled so that max|vel|*dt == reach_cells."""
    st = random_state(po, W, H, D, seed=seed)
    f = np.float32(reach_cells / (4.0 * dt))
    for k in ("u", "v", "w"):
        st[k] = (st[k] * f).astype(np.float32)
    return st


@pytest.mark.parametrize("reach,dt", [(0.9, 0.05), (1.5, 0.05), (2.5, 0.5), (6.0, 0.5), (40.0, 0.5)])
@pytest.mark.parametrize("dims", [(70, 45, 40), (33, 20, 17)])
def test_long_backtrace_advection(po, smk, dims, reach, dt):
    """Stage-level: velocity + density advection on random fields whose backtraces reach `reach` cells (40 = far outside
    the grid: every sample is clamped to the domain), against the oracle and the reference's own kernels on the GPU."""
    W, H, D = dims
    scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.0)], [])
    st = scaled_state(po, W, H, D, reach, dt, seed=31)
    a, b = make_pair(po, smk, scene, st)
    a.flip(); b.flip(); a.fill(); b.fill()
    a.advect_velocity(dt); b.advect_velocity(dt)
    compare(po, a, b, f"{dims} reach {reach}: velocity advection vs oracle")
    a.advect_smoke(dt); b.advect_smoke(dt)
    compare(po, a, b, f"{dims} reach {reach}: density advection vs oracle")
    moved = np.abs(a.get_field(po.U, po.PAST) - st["u"]).max()
    assert moved > 0.0
    if po.have_ref_gpu():
        r = po.RefGPU(W, H, D); po.setup_scene(r, scene); inject(po, r, st)
        try:
            r.flip(); r.fill(); r.advect_velocity(dt); r.advect_smoke(dt)
            compare(po, a, r, f"{dims} reach {reach}: advection vs the reference kernels on the GPU")
        finally:
            r.close()
    a.close()


def test_long_backtrace_full_steps_after_a_hitch(po, smk):
    """Full steps where one tick has dt = 0.5 (the reference passes wall-clock time, main.cpp:891-895): forcing, clamp
    (fires: 3/sqrt(dt) = 4.2), the fused passes and advection with ~2-cell backtraces, against the oracle."""
    sc = po.scaled_scene("C1", 48)
    st = scaled_state(po, 48, 48, 48, 3.0, 0.5, seed=7)
    a, b = make_pair(po, smk, sc, st)
    for dt in (0.05, 0.5, 0.05):
        a.step(dt); b.step(dt)
    compare(po, a, b, "hitch tick dt=0.5")
    a.close()


@pytest.mark.parametrize("n", [48, 40, 67])
def test_default_simulate_path_chunked_readback(po, smk, n):
    """simulate(host, dt) on grids with D >= 34 takes the chunked path (4 density-advection launches + copy stream);
    the HOST buffer must equal the oracle's density exactly, every tick."""
    sc = po.scaled_scene("C1", n)
    a, b = make_pair(po, smk, sc)
    host = np.full((n, n, n), -7.0, dtype=np.float32)
    for t in range(6):
        a.step(po.tick_dt(t), host); b.step(po.tick_dt(t))
        assert np.array_equal(host, b.get_field(po.SMOKE, po.PAST)), f"host readback differs at tick {t}"
    compare(po, a, b, f"C1/{n} with host readback")
    a.close()


def test_host_registration_contract(po, smk):
    """A buffer the library page-locked is re-validated on every use: a second, larger buffer, a buffer registered
    explicitly, and one unregistered between steps all receive the full density."""
    sc = po.scaled_scene("C1", 40)
    a, b = make_pair(po, smk, sc)
    big = np.zeros(2 * 40 ** 3, dtype=np.float32)
    h1 = big[:40 ** 3]
    a.step(0.01, h1); b.step(0.01)
    want = b.get_field(po.SMOKE, po.PAST).ravel()
    assert np.array_equal(h1, want)
    a.unregister_host(h1)
    h2 = big[40 ** 3 // 2: 40 ** 3 // 2 + 40 ** 3]      # overlaps the old range, other base address
    a.register_host(h2)
    a.step(0.05, h2); b.step(0.05)
    assert np.array_equal(h2, b.get_field(po.SMOKE, po.PAST).ravel())
    a.unregister_host(h2)
    a.step(0.05, h1); b.step(0.05)                      # implicit registration again, after an explicit release
    assert np.array_equal(h1, b.get_field(po.SMOKE, po.PAST).ravel())
    a.close()


def test_c1_100_ticks_identical_to_reference_gpu_build(po, smk):
    """BASELINE configs[0] over a long horizon: the clamp (cu:331-352) engages around tick 49, the plume hits the
    obstacle; every field of tick 100 identical to the reference step on the same GPU, host readback included."""
    if not po.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not in the snapshot")
    sc = po.SCENES["C1"]
    a = smk.SmokeSim(*sc[:3]); po.setup_scene(a, sc)
    r = po.RefGPU(*sc[:3]); po.setup_scene(r, sc)
    host = np.zeros((80, 80, 80), dtype=np.float32)
    try:
        vmax = 0.0
        for t in range(100):
            a.step(po.tick_dt(t), host); r.step(po.tick_dt(t))
            if t in (48, 60, 99):
                assert np.array_equal(host, r.host), f"tick {t}"
                vmax = max(vmax, float(np.abs(a.get_field(po.V, po.NOW)).max()))
        compare(po, a, r, "C1 x100 vs reference on GPU")
        assert vmax * vmax * 0.05 > 4.0, f"the clamp should be close to engaging (max|v| {vmax})"
    finally:
        r.close(); a.close()


def test_c2_256_60_ticks_identical_to_reference_gpu_build(po, smk):
    """BASELINE configs[1] at the ticks bench.py times (and beyond): 256^3 rising plume, 60 ticks, identical to the
    reference step on the same GPU -- u, v, w (both buffers), density (both buffers), mask."""
    if not po.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not in the snapshot")
    sc = po.SCENES["C2"]
    a = smk.SmokeSim(*sc[:3]); po.setup_scene(a, sc)
    r = po.RefGPU(*sc[:3]); po.setup_scene(r, sc)
    try:
        for t in range(60):
            a.step(po.tick_dt(t)); r.step(po.tick_dt(t))
            if t in (4, 24):
                x, y = a.get_field(po.SMOKE, po.PAST), r.get_field(po.SMOKE, po.PAST)
                assert np.array_equal(x, y), f"density at tick {t}"
        assert np.array_equal(a.get_field(po.MASK), r.get_field(po.MASK))
        for f in (po.SMOKE, po.U, po.V, po.W):
            for which in (po.NOW, po.PAST):
                x, y = a.get_field(f, which), r.get_field(f, which)
                assert np.array_equal(x, y), (f, which, rel_err(x, y))
                del x, y
        v = a.get_field(po.V, po.NOW)
        assert float(np.abs(v).max()) > 5.0          # a developed plume, not the first puff
    finally:
        r.close(); a.close()


def test_density_surface_write_equals_the_density_buffer(po, smk):
    """SURVEY N1 as specified: with a 3-D array bound (smk_bind_density_array) the density advection kernel itself writes the
    new density with surf3Dwrite; after every step the array must equal the "past" density buffer in EVERY cell -- the
    boundary shell and solid cells advection never writes included -- and the oracle's density.  simulate(nullptr, dt)."""
    import ctypes
    n = 40
    sc = po.scaled_scene("C1", n)
    a, b = make_pair(po, smk, sc)
    L = smk.load_library()
    arr = L.smk_test_array_create(n, n, n)
    assert arr
    a.bind_density_array(arr)
    out = np.full((n, n, n), -3.0, dtype=np.float32)
    for t in range(5):
        a.step(po.tick_dt(t), None); b.step(po.tick_dt(t))
        a.sync()
        assert L.smk_test_array_read(arr, out.ctypes.data_as(ctypes.c_void_p), n, n, n) == 0
        assert np.array_equal(out, a.get_field(po.SMOKE, po.PAST)), f"array != device density at tick {t}"
        assert np.array_equal(out, b.get_field(po.SMOKE, po.PAST)), f"array != oracle density at tick {t}"
    compare(po, a, b, "surface-write variant")
    a.bind_density_array(None)
    a.step(0.05, None); b.step(0.05)
    compare(po, a, b, "after unbinding")
    L.smk_test_array_destroy(arr)
    a.close()


def test_bit_packed_mask_round_trip_and_step(po, smk):
    """a voxelised solid uploaded as one bit per cell == the same mask uploaded as bytes (steps identical
    to the oracle); reading the packed mask back returns the bits."""
    W, H, D = 24, 20, 18
    vox = np.ones((D, H, W), dtype=np.uint8); vox[:, 0, :] = 0; vox[6:12, 8:11, 5:19] = 0; vox[3, 5, 7] = 0
    bits = np.packbits(vox.ravel(), bitorder="little")
    a = smk.SmokeSim(W, H, D); b = po.Oracle(W, H, D)
    for e in (a, b):
        e.add_source(12, 4, 9, 3.0)
    a.set_mask_bits(bits); b.set_field(po.MASK, po.NOW, vox)
    assert np.array_equal(a.get_field(po.MASK), vox)
    assert np.array_equal(a.get_mask_bits(), bits)
    for t in range(4):
        a.step(po.tick_dt(t)); b.step(po.tick_dt(t))
    compare(po, a, b, "bit-packed voxel mask")
    assert np.array_equal(np.unpackbits(a.get_mask_bits(), bitorder="little")[:W * H * D].reshape(D, H, W), vox)
    a.close()


def test_sparse_readback_leaves_the_host_buffer_identical(po, smk):
    """smk_set_readback_box: the caller's buffer after every step == the device density, through a growing plume, a
    source that jumps (fallback to the full copy), an injected density, a switch to a second buffer and back, and the
    copies get smaller than the field once the plume is known."""
    W, H, D = 64, 96, 80
    a = smk.SmokeSim(W, H, D)
    a.set_params(-9.82, 15.0)
    sid = a.add_source(32.0, 12.0, 40.0, 5.0)
    a.add_obstacle(32.0, 40.0, 40.0, 0, 0, 0, 6.0)
    a.set_readback_box(1)
    h1 = np.full((D, H, W), 7.0, dtype=np.float32)   # garbage the first (full) copy must overwrite
    h2 = np.full((D, H, W), -3.0, dtype=np.float32)
    full = W * H * D * 4
    small = 0
    before = a.readback_bytes()
    for t in range(40):
        if t == 20:
            a.update_object_pos(sid, 20.0, 70.0, 60.0)            # far outside the known rows
        if t == 28:
            rng = np.random.default_rng(1)
            a.set_field(po.SMOKE, po.NOW, (rng.uniform(0, 1, (D, H, W)) < 0.001).astype(np.float32))
        buf = h2 if 12 <= t < 16 else h1
        a.step(po.tick_dt(t), buf)
        assert np.array_equal(buf, a.get_field(po.SMOKE, po.PAST)), t
        now = a.readback_bytes()
        small += (now - before) < full
        before = now
    assert small >= 10
    a.close()




def test_omega_product_in_binary32_equals_the_double_product_for_every_input(smk):
    """The TMA-staged pressure pass evaluates `(float)((double)q * -1.9)` (cu:384) with four packed binary32 operations and
    a tie-to-even select instead of F2F / DMUL / F2F.  Every one of the 2^32 bit patterns (inf / nan and the denormal-range
    band that never reaches it excluded) through both on the device: zero mismatches, and the tie rule must actually fire
    (1/19 of the inputs are exact ties of 19 q / 10)."""
    bad, ties = smk.selfcheck_omega()
    assert bad == 0
    assert 190_000_000 < ties < 210_000_000        # 199 648 560 on the CPU twin


def test_pass_kernel_choice_is_automatic_by_grid_size_and_identical_either_way(po, smk):
    """smk_set_pass_kernel(AUTO): the TMA-staged pass from 5 M nodes per slab, the round-1 kernel below (launch-bound grids);
    forcing either kernel on the same scene gives the same bits (device hashes of every field)."""
    small = smk.SmokeSim(64, 64, 64); po.setup_scene(small, (64, 64, 64, -9.82, 4.0, [(32, 12, 32, 5)], []))
    small.set_pass_kernel("auto"); small.step(0.01)
    assert small.last_pass_kernel() == "reg"
    small.close()
    n = 176   # 177^3 = 5.5 M nodes
    scene = (n, n, n, -9.82, 6.0, [(n / 2, n / 6, n / 2, n / 12)], [(n / 2, n / 2, n / 2, n / 10)])
    hashes = {}
    for kind in ("auto", "reg", "tma"):
        s = smk.SmokeSim(n, n, n); po.setup_scene(s, scene); s.set_pass_kernel(kind)
        for t in range(2):
            s.step(po.tick_dt(t))
        hashes[kind] = (s.last_pass_kernel(), s.hash_owned())
        s.close()
    assert hashes["auto"][0] == "tma" and hashes["reg"][0] == "reg" and hashes["tma"][0] == "tma"
    assert hashes["auto"][1] == hashes["reg"][1] == hashes["tma"][1]
