"""CPU tests of the multi-GPU z-slab path (no GPU): the decomposition, schedule and halo regions come from the
product library's host-only C ABI (csrc/slab_plan.h -- the very code smk_step executes) and are run here with the
oracle as the compute engine: (1) all ranks emulated in one process, (2) two real processes under a gloo group
exchanging halos through smoke_simulation_b200.slab.  Bar: every rank's owned planes bit-identical to the
single-domain run."""
import os
import sys

import numpy as np
import pytest

from conftest import inject, random_state

FIELDS = (("u", 1), ("v", 2), ("w", 3))


def make_rank_engines(po, slab, scene, st, world, ghost):
    W, H, D = scene[:3]
    ranks = []
    for r in range(world):
        g = slab.geometry(W, H, D, world, r, ghost)
        assert g["ok"] == 1, g
        e = po.Oracle(W, H, D, contract=1)
        po.setup_scene(e, scene)
        inject(po, e, st)
        # poison everything this rank does not store: it must never influence the owned planes
        for f in (po.U, po.V, po.W):
            for b in (po.BUF0, po.BUF1):
                a = e.get_field(f, b); a[:g["zlo"]] = 1e30; a[g["zhc"] + 1:] = -1e30; e.set_field(f, b, a)
        for b in (po.BUF0, po.BUF1):
            a = e.get_field(po.SMOKE, b); a[:g["zlo"]] = 1e30; a[g["zhc"]:] = -1e30; e.set_field(po.SMOKE, b, a)
        ranks.append((g, e))
    return ranks


def run_op(po, e, g, D, op, dt):
    name, a, b, p0, p1 = op
    if name == "flip":
        e.flip()
    elif name == "fill":
        e.fill()
    elif name == "force":
        e.integrate_r(dt, a, b); e.clamp_r(dt, a, b)
    elif name == "pressure":
        for j in range(p1):
            e.pressure_halfsweep_r((p0 + j) & 1, max(1, g["zlo"]), min(D - 1, g["zhc"]))
    elif name == "advect_vel":
        e.advect_velocity_r(dt, a, b)
    elif name == "advect_smoke":
        e.advect_smoke_r(dt, a, b)
    else:
        raise AssertionError(name)


def compare_owned(po, g, e, ref, what):
    nlo, nhi = g["own_node_lo"], g["own_node_hi"]
    for f in (po.U, po.V, po.W):
        for which in (po.NOW, po.PAST):
            x, y = e.get_field(f, which)[nlo:nhi + 1], ref.get_field(f, which)[nlo:nhi + 1]
            assert np.array_equal(x, y), f"{what}: field {f} buffer {which} differs on owned node planes [{nlo},{nhi}]"
    for which in (po.NOW, po.PAST):
        x, y = e.get_field(po.SMOKE, which)[g["c0"]:g["c1"]], ref.get_field(po.SMOKE, which)[g["c0"]:g["c1"]]
        assert np.array_equal(x, y), f"{what}: density buffer {which} differs on owned cell planes"


@pytest.mark.parametrize("world,ghost,fuse,dims", [(2, 4, 4, (12, 10, 24)), (2, 8, 4, (10, 9, 40)), (3, 5, 2, (9, 8, 33)),
                                                   (4, 4, 1, (8, 8, 32)), (4, 9, 4, (8, 7, 64)), (2, 6, 4, (16, 12, 17))])
def test_slab_schedule_emulated_ranks_bit_identical(po, smk, world, ghost, fuse, dims):
    from smoke_simulation_b200 import slab
    W, H, D = dims
    iterations, steps, dt = 7, 3, 0.05
    scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.5)], [(W / 2, H / 3, D / 3, 2.0)])
    st = random_state(po, W, H, D, seed=5)
    ref = po.Oracle(W, H, D, contract=1); po.setup_scene(ref, scene); inject(po, ref, st); ref.set_iterations(iterations)
    ranks = make_rank_engines(po, slab, scene, st, world, ghost)
    plans = [slab.plan(W, H, D, world, r, ghost, iterations, fuse, steps) for r in range(world)]
    n = len(plans[0])
    assert all(len(p) == n for p in plans), [len(p) for p in plans]
    n_exchange = 0
    step_ops = n // steps
    for i in range(n):
        kinds = {p[i][0] for p in plans}
        if "exchange" in kinds:
            assert kinds == {"exchange"}, f"op {i}: ranks disagree on the exchange point: {[p[i] for p in plans]}"
            set_id = plans[0][i][1]
            n_exchange += 1
            for r, (g, e) in enumerate(ranks):
                for (side, send_lo, send_n, recv_lo, recv_n) in slab.regions(W, H, D, world, r, ghost, set_id):
                    peer = ranks[r - 1 if side == 0 else r + 1][1]
                    # what the peer sends me = the peer's region on the opposite side
                    (pside, ps_lo, ps_n, _, _), = [q for q in slab.regions(W, H, D, world, r - 1 if side == 0 else r + 1, ghost, set_id) if q[0] == 1 - side]
                    assert ps_n == recv_n and ps_lo == recv_lo, (r, side, ps_lo, ps_n, recv_lo, recv_n)
                    for f in ((po.U, po.V, po.W) if set_id == 0 else (po.SMOKE,)):
                        mine = e.get_field(f, po.NOW); mine[recv_lo:recv_lo + recv_n] = peer.get_field(f, po.NOW)[ps_lo:ps_lo + ps_n]
                        e.set_field(f, po.NOW, mine)
        else:
            for (g, e), p in zip(ranks, plans):
                run_op(po, e, g, D, p[i], dt)
        if (i + 1) % step_ops == 0 and len(set(len(p) for p in plans)) == 1:
            pass
    for t in range(steps):
        ref.step(dt)
    for r, (g, e) in enumerate(ranks):
        compare_owned(po, g, e, ref, f"world={world} ghost={ghost} fuse={fuse} rank {r}")
    if world > 1:
        assert n_exchange >= steps  # at least the pressure passes need ghosts


def test_geometry_and_regions_are_consistent(smk):
    from smoke_simulation_b200 import slab
    for (W, H, D, world, ghost) in [(256, 256, 256, 8, 8), (512, 512, 512, 8, 12), (80, 80, 80, 2, 4), (64, 64, 4096, 8, 8)]:
        covered = []
        for r in range(world):
            g = slab.geometry(W, H, D, world, r, ghost)
            assert g["ok"] == 1
            covered += list(range(g["c0"], g["c1"]))
            assert g["zlo"] == max(0, g["c0"] - ghost) and g["zhc"] == min(D, g["c1"] + ghost)
            for set_id in (0, 1):
                for (side, send_lo, send_n, recv_lo, recv_n) in slab.regions(W, H, D, world, r, ghost, set_id):
                    own_hi = g["own_node_hi"] if set_id == 0 else g["c1"] - 1
                    assert g["c0"] <= send_lo and send_lo + send_n - 1 <= own_hi, "only owned planes are sent"
                    assert recv_lo + recv_n - 1 <= (g["zhc"] if set_id == 0 else g["zhc"] - 1) and recv_lo >= g["zlo"]
                    assert (recv_lo + recv_n <= g["c0"]) or (recv_lo >= g["c1"]), "ghost planes only"
        assert covered == list(range(D))
    assert slab.geometry(64, 64, 64, 16, 3, 8)["ok"] == 0  # 4-plane slabs cannot serve 8 ghost planes
    p1 = slab.plan(64, 64, 64, 1, 0, 8, 30, 4, 2)
    assert not any(op[0] == "exchange" for op in p1) and sum(op[0] == "pressure" for op in p1) == 30


def _gloo_worker(rank, world, port, dims, ghost, fuse, iterations, steps, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        import pyoracle as po
        from conftest import inject as inj, random_state as rs
        from smoke_simulation_b200 import slab
        dist.init_process_group("gloo", rank=rank, world_size=world)
        W, H, D = dims
        dt = 0.05
        scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.5)], [])
        st = rs(po, W, H, D, seed=9)
        g = slab.geometry(W, H, D, world, rank, ghost)
        e = po.Oracle(W, H, D, contract=1); po.setup_scene(e, scene); inj(po, e, st)
        for op in slab.plan(W, H, D, world, rank, ghost, iterations, fuse, steps):
            if op[0] == "exchange":
                fields = (po.U, po.V, po.W) if op[1] == 0 else (po.SMOKE,)
                ts = [torch.from_numpy(e.get_field(f, po.NOW)) for f in fields]
                slab.exchange_arrays(dist, rank, slab.regions(W, H, D, world, rank, ghost, op[1]), ts)
                for f, t in zip(fields, ts):
                    e.set_field(f, po.NOW, t.numpy())
            else:
                run_op(po, e, g, D, op, dt)
        ref = po.Oracle(W, H, D, contract=1); po.setup_scene(ref, scene); inj(po, ref, st); ref.set_iterations(iterations)
        for t in range(steps):
            ref.step(dt)
        compare_owned(po, g, e, ref, f"gloo rank {rank}")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


def test_slab_two_processes_gloo(po, smk):
    """world_size 2 over gloo: real processes, halos moved by smoke_simulation_b200.slab.exchange_arrays."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, (10, 9, 28), 5, 4, 6, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
