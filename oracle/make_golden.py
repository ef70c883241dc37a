"""Generate tests/golden/*.npz from the REFERENCE's own kernel bodies (oracle/_ref/libref_cpu.so).

Run in the build container (needs /root/reference to build oracle/_ref):   python oracle/make_golden.py
The reference ships no golden vectors; these fixtures are outputs of its own code
(g++ -O2, no FMA contraction), committed so that the GPU box -- where /root/reference does not exist --
can check both the oracle restatement (bit-exact) and the CUDA path (<= 1e-5 relative).

Fixtures
  small_scene.npz   20x18x16 grid, 1 source + 2 obstacles (last-obstacle-wins), 6 ticks, all fields
  small_random.npz  17x21x19 grid, random u/v/w/density/mask injected, no obstacle, 3 ticks, all fields
  c1_80.npz         the 80^3 default scene (C1), 20 ticks: mask (bit-packed), density, and u/v/w planes
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL_SCENE = (20, 18, 16, -9.82, 2.0, [(10.0, 8.0, 8.0, 3.0)], [(13.0, 4.0, 9.0, 3.0), (6.0, 6.0, 5.0, 2.5)])
SMALL_RANDOM = (17, 21, 19, -9.82, 6.0, [(8.0, 9.0, 9.0, 2.5)], [])


def random_state(W, H, D, seed=1234):
    rng = np.random.default_rng(seed)
    st = {}
    for f, n in ((po.U, "u"), (po.V, "v"), (po.W, "w")):
        st[n] = rng.uniform(-4, 4, po.field_shape(f, W, H, D)).astype(np.float32)
    st["smoke"] = rng.uniform(0, 1, (D, H, W)).astype(np.float32)
    m = (rng.uniform(0, 1, (D, H, W)) < 0.9).astype(np.uint8)
    m[:, 0, :] = 0
    st["mask"] = m
    return st


def inject(e, st):
    e.set_field(po.U, po.BUF0, st["u"]); e.set_field(po.V, po.BUF0, st["v"]); e.set_field(po.W, po.BUF0, st["w"])
    e.set_field(po.SMOKE, po.BUF0, st["smoke"]); e.set_field(po.MASK, po.NOW, st["mask"])


def all_fields(e):
    d = {"mask": e.get_field(po.MASK)}
    for f, n in ((po.SMOKE, "smoke"), (po.U, "u"), (po.V, "v"), (po.W, "w")):
        d[n + "_now"] = e.get_field(f, po.NOW)
        d[n + "_past"] = e.get_field(f, po.PAST)
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    po.build(ref=True)

    e = po.RefCPU(*SMALL_SCENE[:3]); po.setup_scene(e, SMALL_SCENE)
    for t in range(6):
        e.step(po.tick_dt(t))
    np.savez_compressed(os.path.join(OUT, "small_scene.npz"), ticks=6, **all_fields(e))

    e = po.RefCPU(*SMALL_RANDOM[:3]); po.setup_scene(e, SMALL_RANDOM)
    inject(e, random_state(*SMALL_RANDOM[:3]))
    for t in range(3):
        e.step(po.tick_dt(t))
    np.savez_compressed(os.path.join(OUT, "small_random.npz"), ticks=3, **all_fields(e))

    sc = po.SCENES["C1"]
    e = po.RefCPU(*sc[:3]); po.setup_scene(e, sc)
    for t in range(20):
        e.step(po.tick_dt(t))
    u, v, w = (e.get_field(f, po.NOW) for f in (po.U, po.V, po.W))
    np.savez_compressed(
        os.path.join(OUT, "c1_80.npz"), ticks=20,
        mask_bits=np.packbits(e.get_field(po.MASK)), density=e.get_field(po.SMOKE, po.PAST),
        u_y40=u[:, 40, :], v_y40=v[:, 40, :], w_y40=w[:, 40, :], u_z40=u[40], v_z40=v[40], w_z40=w[40],
        absmax=np.array([np.abs(u).max(), np.abs(v).max(), np.abs(w).max()], dtype=np.float32),
        sums=np.array([u.astype(np.float64).sum(), v.astype(np.float64).sum(), w.astype(np.float64).sum()]),
        maxdiv=np.float32(e.max_divergence()))
    for n in sorted(os.listdir(OUT)):
        print(n, os.path.getsize(os.path.join(OUT, n)))


if __name__ == "__main__":
    main()
