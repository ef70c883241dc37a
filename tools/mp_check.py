#!/usr/bin/env python
"""Multi-process slab check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mp_check.py [p2p|nccl] [W H Dper]

Every rank steps its z-slab of a small scene through smk_step (the same call bench.py times) AND the whole grid on its
own GPU, then compares its owned planes bit for bit: u, v, w (now / past), density (now / past) and the host readback.
Exit status 0 = every rank identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pyoracle as po                      # noqa: E402  (scene helpers and field ids only)
import smoke_simulation_b200 as smk        # noqa: E402
from conftest import inject, random_state  # noqa: E402


def main():
    transport = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    W, H, dper = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (64, 48, 40)
    mode = sys.argv[5] if len(sys.argv) > 5 else "plain"   # plain | hitch (one tick with dt = 0.5) | reach (dt too large for the ghost planes)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D, ghost, iterations, steps = dper * world, 8, 7, 4
    scene = (W, H, D, -9.82, 6.0, [(W / 2, H / 3, D / 2, 5.0), (W / 3, H / 2, dper - 1.0, 4.0)], [(W / 2, H / 2, dper + 0.5, 6.0)])
    st = random_state(po, W, H, D, seed=17)
    ref = smk.SmokeSim(W, H, D)
    sim = smk.SmokeSim(W, H, D, slab=(rank, world), ghost=ghost)
    for e in (ref, sim):
        po.setup_scene(e, scene); inject(po, e, st); e.set_solver(0, iterations, 0)
    keep = None
    if transport == "p2p":
        smk.slab.attach_peers_ipc(sim, rank, world, dist, torch.device("cuda", local))
    else:
        keep = smk.slab.TorchTransport(rank, world); sim.set_exchange(keep)
    dist.barrier()
    host = np.zeros((D, H, W), dtype=np.float32); ref_host = np.zeros_like(host)
    dts = [po.tick_dt(t) for t in range(steps)]
    if mode == "hitch":
        dts[2] = 0.5          # u, v, w ~ U(-4, 4) after projection: backtraces of 2-3 planes (the reference passes wall-clock dt)
    if mode == "reach":
        dts[1] = 16.0         # the clamp bounds the reach by 3 sqrt(dt) = 12 planes: beyond ghost - 1 -> the slabs whose backtraces leave their valid planes raise SMK_ERR_REACH, nobody returns stale data
    g = smk.slab.geometry(W, H, D, world, rank, ghost)

    def mismatches():
        bad = []
        for f, name in ((po.U, "u"), (po.V, "v"), (po.W, "w")):
            for which in (po.NOW, po.PAST):
                sl = slice(g["own_node_lo"], g["own_node_hi"] + 1)
                if not np.array_equal(sim.get_field(f, which)[sl], ref.get_field(f, which)[sl]):
                    bad.append((name, which))
        for which in (po.NOW, po.PAST):
            if not np.array_equal(sim.get_field(po.SMOKE, which)[g["c0"]:g["c1"]], ref.get_field(po.SMOKE, which)[g["c0"]:g["c1"]]):
                bad.append(("smoke", which))
        return bad

    if mode == "reach":
        # SMK_ERR_REACH is per rank: a slab raises it when ITS backtraces left its valid planes.  The ranks agree on stopping
        # after every step (what an application's driver has to do: a rank that carried on would wait for neighbours that
        # stopped); a rank that did NOT raise must hold exactly the single-GPU result of that step -- never stale data.
        raised_on, stale = [], 0
        for t in range(steps):
            mine = 0
            try:
                sim.step(dts[t], host)
            except smk.SmokeError as ex:
                if "error 4" not in str(ex):
                    print(f"mp_check rank {rank}: step {t} failed: {ex}", flush=True)
                    raise
                mine = 1
            ref.step(dts[t], ref_host)
            flags = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
            dist.all_gather(flags, torch.tensor([mine], dtype=torch.int32, device="cuda"))
            raised_on = [r for r in range(world) if int(flags[r].item())]
            if raised_on:
                if not mine and mismatches():
                    stale = 1
                    print(f"mp_check rank {rank}: no SMK_ERR_REACH but the slab differs from the single-GPU run: {mismatches()}", flush=True)
                break
        tot = torch.tensor([stale], device="cuda"); dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        ok = bool(raised_on) and not int(tot.item())
        if rank == 0:
            print(f"mp_check {transport} world={world} grid={W}x{H}x{D} reach: " +
                  (f"OK (SMK_ERR_REACH raised on ranks {raised_on}; every other rank identical to the single-GPU run)" if ok
                   else "FAILED (no SMK_ERR_REACH)" if not raised_on else "FAILED (stale data without SMK_ERR_REACH)"), flush=True)
        sim.close(); ref.close(); dist.barrier(); dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    for t in range(steps):
        sim.step(dts[t], host)
        ref.step(dts[t], ref_host)
    bad = mismatches()
    if not np.array_equal(host[g["c0"]:g["c1"]], ref_host[g["c0"]:g["c1"]]):
        bad.append(("host readback", 0))
    ok = torch.tensor([0 if bad else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if bad:
        print(f"rank {rank}: MISMATCH {bad}", flush=True)
    if rank == 0:
        print(f"mp_check {transport} world={world} grid={W}x{H}x{D} mode={mode} exchanges/step={sim.exchange_count() / steps:.1f}:",
              "OK (bit-identical to the single-GPU run)" if int(ok.item()) else "FAILED", flush=True)
    sim.close(); ref.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) else 1)


if __name__ == "__main__":
    main()
