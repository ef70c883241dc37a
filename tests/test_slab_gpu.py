"""GPU test of the z-slab path with VIRTUAL slabs: several slab instances on one GPU, driven in lock-step through
the plan (smk_exec_op), halos moved by device-to-device copies between the instances.  Exercises exactly the kernels,
ranges, ghost handling and exchange regions of a multi-GPU run; bar: owned planes bit-identical to the 1-GPU run."""
import numpy as np
import pytest

from conftest import inject, random_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,ghost,fuse,dims", [(2, 8, 4, (70, 40, 48)), (4, 4, 4, (40, 30, 64)), (3, 5, 2, (33, 20, 40)),
                                                   (2, 6, 1, (20, 18, 30))])
def test_virtual_slabs_bit_identical_to_single_gpu(po, smk, world, ghost, fuse, dims):
    import torch
    from smoke_simulation_b200 import slab
    from smoke_simulation_b200.slab import _DevBuf
    W, H, D = dims
    iterations, steps, dt = 7, 3, 0.05
    scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.5)], [(W / 2, H / 3, D / 3, 2.0)])
    st = random_state(po, W, H, D, seed=5)
    ref = smk.SmokeSim(W, H, D); po.setup_scene(ref, scene); inject(po, ref, st); ref.set_solver(0, iterations, fuse)
    sims, geoms, pending = [], [], {}
    for r in range(world):
        s = smk.SmokeSim(W, H, D, slab=(r, world), ghost=ghost)
        po.setup_scene(s, scene); inject(po, s, st); s.set_solver(0, iterations, fuse)
        s.set_exchange(lambda set_id, regs, stream, r=r: pending.__setitem__(r, regs) or 0)  # record, copy below
        sims.append(s); geoms.append(slab.geometry(W, H, D, world, r, ghost))
    plans = [slab.plan(W, H, D, world, r, ghost, iterations, fuse, steps) for r in range(world)]
    assert len({len(p) for p in plans}) == 1
    n_ex = 0
    for i in range(len(plans[0])):
        for s, p in zip(sims, plans):
            s.exec_op(p[i], dt)
        if plans[0][i][0] == "exchange":
            n_ex += 1
            for s in sims:
                s.sync()
            nf = 3 if plans[0][i][1] == 0 else 1
            for r in range(world):
                for k, (side, send_ptr, recv_ptr, send_bytes, recv_bytes) in enumerate(pending[r]):
                    peer = r - 1 if side == 0 else r + 1
                    # the peer's region towards me, same field slot
                    cand = [q for q in pending[peer] if q[0] == 1 - side]
                    q = cand[k % nf]
                    assert q[3] == recv_bytes, (q[3], recv_bytes)
                    dst = torch.as_tensor(_DevBuf(recv_ptr, recv_bytes), device="cuda")
                    src = torch.as_tensor(_DevBuf(q[1], q[3]), device="cuda")
                    dst.copy_(src)
            torch.cuda.synchronize()
            pending.clear()
    for t in range(steps):
        ref.step(dt)
    assert n_ex >= steps
    for r, (s, g) in enumerate(zip(sims, geoms)):
        s.sync()
        for f in (po.U, po.V, po.W):
            for which in (po.NOW, po.PAST):
                x = s.get_field(f, which)[g["own_node_lo"]:g["own_node_hi"] + 1]
                y = ref.get_field(f, which)[g["own_node_lo"]:g["own_node_hi"] + 1]
                assert np.array_equal(x, y), (r, f, which)
        for which in (po.NOW, po.PAST):
            assert np.array_equal(s.get_field(po.SMOKE, which)[g["c0"]:g["c1"]], ref.get_field(po.SMOKE, which)[g["c0"]:g["c1"]]), (r, which)
        assert np.array_equal(s.get_field(po.MASK)[g["zlo"]:g["zhc"]], ref.get_field(po.MASK)[g["zlo"]:g["zhc"]])
        s.close()
    ref.close()


@pytest.mark.parametrize("world,ghost,fuse,dims", [(2, 8, 4, (70, 40, 48)), (4, 4, 4, (40, 30, 64)), (3, 5, 4, (60, 37, 41)),
                                                   (2, 6, 2, (20, 18, 30)), (2, 8, 4, (256, 256, 160))])
@pytest.mark.parametrize("ctas", [0, 3, 148])
def test_peer_memory_path_bit_identical_to_single_gpu(po, smk, world, ghost, fuse, dims, ctas):
    _peer_memory_run(po, smk, world, ghost, fuse, dims, ctas, [0.05, 0.05, 0.05])


@pytest.mark.parametrize("world,ghost,dims,dts,expect_reach", [(2, 8, (70, 40, 48), [0.05, 0.5, 0.05], False), (3, 6, (40, 30, 60), [0.5, 0.05], False),
                                                              (2, 8, (70, 40, 48), [0.05, 16.0], True)])
def test_slabs_survive_a_frame_hitch(po, smk, world, ghost, dims, dts, expect_reach):
    """dt is wall-clock time in the reference (main.cpp:891-895).  A tick with dt = 0.5 on random u, v, w ~ U(-4, 4) backtraces 2-3
    planes across the slab boundary: the ghost refresh in front of advection covers the whole ghost depth, so the slabs stay
    bit-identical to the single-GPU run.  Only a reach beyond the ghost allocation itself (dt = 16: the clamp bounds the reach by 3 sqrt(dt) = 12 planes, ghost 8) is an
    error, and it is reported (SMK_ERR_REACH), never a silent stale read."""
    if expect_reach:
        with pytest.raises(smk.SmokeError, match="error 4"):
            _peer_memory_run(po, smk, world, ghost, 4, dims, 0, dts)
    else:
        _peer_memory_run(po, smk, world, ghost, 4, dims, 0, dts)


def _peer_memory_run(po, smk, world, ghost, fuse, dims, ctas, dts):
    """The B200-native transport: pressure passes read the neighbours' boundary planes straight from their memory
    (epoch handshake per pass, no ghost copies), halo refreshes before advection are pulls over the mapped memory.
    Virtual slabs in one process (pointer attach instead of CUDA IPC); every rank just runs smk_step_async.
    ctas: schedule of a pass -- the (tile, z-chunk) grid (0, default) or balanced piece lists on 3 / 148 CTAs (boundary
    pieces first, in-kernel handshake per piece).
    256x256x160 on two slabs is the case ADVICE r1 names: the slabs own 80 and 81 node planes, which chunk differently
    (2 chunks -> stream-level handshake on one rank, 5 chunks -> in-kernel handshake on the other); both must consume the
    same epoch values per pass."""
    from smoke_simulation_b200 import slab
    W, H, D = dims
    iterations, steps = 7, len(dts)
    scene = (W, H, D, -9.82, 3.0, [(W / 2, H / 2, D / 2, 2.5)], [(W / 2, H / 3, D / 3, 2.0)])
    st = random_state(po, W, H, D, seed=5)
    ref = smk.SmokeSim(W, H, D); po.setup_scene(ref, scene); inject(po, ref, st); ref.set_solver(0, iterations, fuse)
    sims = []
    for r in range(world):
        s = smk.SmokeSim(W, H, D, slab=(r, world), ghost=ghost)
        po.setup_scene(s, scene); inject(po, s, st); s.set_solver(0, iterations, fuse); s.set_pass_ctas(ctas)
        sims.append(s)
    slab.attach_peers_local(sims)
    # One host thread drives all slabs on one GPU: run the plan op by op and enqueue every slab's epoch signal before any
    # slab's wait (streams may share a hardware queue; in the real multi-process run each GPU simply calls smk_step).
    plans = [slab.plan_p2p(W, H, D, world, r, ghost, iterations, fuse, steps) for r in range(world)]
    assert len({len(p) for p in plans}) == 1
    step_i = -1
    for i in range(len(plans[0])):
        op = plans[0][i]
        if op[0] == "flip":
            step_i += 1
        if op[0] == "exchange" or (op[0] == "pressure" and op[4] == 4):
            for s in sims:
                s.p2p_presignal()
        for s, p in zip(sims, plans):
            s.exec_op(p[i], dts[step_i])
    for t in range(steps):
        ref.step(dts[t])
    for s in sims:
        s.sync()
    assert all(s.exchange_count() >= steps for s in sims)
    for r, s in enumerate(sims):
        g = slab.geometry(W, H, D, world, r, ghost)
        for f in (po.U, po.V, po.W):
            for which in (po.NOW, po.PAST):
                x = s.get_field(f, which)[g["own_node_lo"]:g["own_node_hi"] + 1]
                y = ref.get_field(f, which)[g["own_node_lo"]:g["own_node_hi"] + 1]
                assert np.array_equal(x, y), (r, f, which, float(np.abs(x - y).max()))
        for which in (po.NOW, po.PAST):
            assert np.array_equal(s.get_field(po.SMOKE, which)[g["c0"]:g["c1"]], ref.get_field(po.SMOKE, which)[g["c0"]:g["c1"]]), (r, which)
    for s in sims:
        s.close()
    ref.close()
