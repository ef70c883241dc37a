#!/usr/bin/env python
"""Small workloads for compute-sanitizer (tools/sanitize.sh): a single-GPU scene with an obstacle through smk_step with host
readback (fused passes with forcing, TMA-staged velocity advection, chunked readback path), the stage-level kernels on random
fields, and two virtual slabs over the peer-memory path (in-kernel neighbour reads + epoch handshake, ghost pulls)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import smoke_simulation_b200 as smk
from smoke_simulation_b200 import scenes as po, slab

what = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(3)
if what in ("all", "single"):
    W, H, D = 72, 40, 44
    a = smk.SmokeSim(W, H, D); a.set_params(-9.82, 6.0); a.add_source(30, 8, 20, 4.0); a.add_obstacle(36, 22, 22, 0, 0, 0, 5.0)
    host = np.zeros((D, H, W), dtype=np.float32)
    for t in range(3):
        a.step(po.tick_dt(t), host)
    a.set_solver(0, 3, 1); a.step(0.05, host)          # one launch per half-sweep
    a.set_solver(1, 4, 0); a.step(0.05)                # Jacobi extension
    print("single: sum", float(host.sum()), "max|div|", a.max_divergence()); a.close()
if what in ("all", "slabs"):
    W, H, D, world, ghost = 70, 40, 48, 2, 8
    sims = [smk.SmokeSim(W, H, D, slab=(r, world), ghost=ghost) for r in range(world)]
    for s in sims:
        s.set_params(-9.82, 6.0); s.add_source(30, 10, 24, 5.0); s.set_solver(0, 4, 4)
    slab.attach_peers_local(sims)
    plans = [slab.plan_p2p(W, H, D, world, r, ghost, 4, 4, 2) for r in range(world)]
    for i in range(len(plans[0])):
        op = plans[0][i]
        if op[0] == "exchange" or (op[0] == "pressure" and op[4] == 4):
            for s in sims: s.p2p_presignal()
        for s, p in zip(sims, plans): s.exec_op(p[i], 0.05)
    for s in sims: s.sync()
    print("slabs: hashes", [hex(s.hash_owned()[6]) for s in sims])
    for s in sims: s.close()
