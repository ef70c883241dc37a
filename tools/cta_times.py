#!/usr/bin/env python
"""Per-CTA timing of the fused pressure pass (development aid):  SMK_PASS_DEBUG=1 SMK_PASS_KERNEL=tma python tools/cta_times.py [C2|C3]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smoke_simulation_b200 as smk
from smoke_simulation_b200 import scenes as po
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
sc = po.SCENES[name]
sim = smk.SmokeSim(*sc[:3]); po.setup_scene(sim, sc)
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for t in range(NT):
    sim.step(po.tick_dt(t))
L = smk.load_library()
L.smk_debug_pass_ctas.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
buf = np.zeros((65536, 4), dtype=np.int64)
n = L.smk_debug_pass_ctas(sim.h, buf.ctypes.data_as(C.c_void_p), 65536)
a = buf[:n]
if n <= 0:
    print("no data", n); sys.exit(0)
t0 = a[:, 0].min()
start = a[:, 0] - t0; dur = a[:, 1]; sm = a[:, 2]; var = a[:, 3] // 1000; planes = a[:, 3] % 1000
end = (start + dur).max()
print(f"{name}: {n} CTAs, pass span {end} cycles = {end / 1.965e3:.1f} us (at 1965 MHz)")
for v in (0, 1):
    m = var == v
    if m.any():
        print(f"  variant {'general' if v else 'lean'}: {m.sum()} CTAs, cycles min/med/max {dur[m].min()} / {int(np.median(dur[m]))} / {dur[m].max()}, per plane-step med {np.median(dur[m] / (planes[m] + 8)):.0f}")
busy = {}
for s_, d_ in zip(sm, dur):
    busy[s_] = busy.get(s_, 0) + d_
b = np.array(list(busy.values()))
print(f"  SMs used {len(b)}, busy cycles per SM min/mean/max {b.min()} / {int(b.mean())} / {b.max()}; ideal balanced span {int(dur.sum() / 148)}")
late = np.sort(start)[148:] if n > 148 else []
if len(late): print(f"  second-wave starts: first {late[0]}, last {late[-1]}")

# ---- per-step trace of one lean CTA (tile (1,2), chunk 1): lane 0 of every warp stamps clock64 after the step barrier
# wait + TMA issue (0), after its sweeps (1) and at the end of the step (2)
tr = np.zeros((34, 80, 8), dtype=np.int64)
if L.smk_debug_pass_ctas(sim.h, tr.ctypes.data_as(C.c_void_p), -1) > 0 and tr[5, 20, 0] > 0:
    steps = range(20, 50)
    for w in (0, 3, 5, 10, 15):
        a0 = tr[w, 20:50, 0]; a1 = tr[w, 20:50, 1]; a2 = tr[w, 20:50, 2]; nxt = tr[w, 21:51, 0]
        print(f"  warp {w:2d}: sweeps {np.median(a1 - a0):6.0f}  arrive->end of step {np.median(a2 - a1):6.0f}  end->next step's sweeps start (wait) {np.median(nxt - a2):6.0f}  step period {np.median(nxt - a0):6.0f}")
    print("  warp 0: wait done -> TMA issued (producer lane), median cycles:", int(np.median(tr[0, 21:50, 0] - tr[0, 21:50, 3])))
    iss = tr[16, 24:50, :]
    print("  producer: expect_tx", int(np.median(iss[:, 1] - iss[:, 0])), " TMA 4-D (u,v,w)", int(np.median(iss[:, 2] - iss[:, 1])), " TMA 3-D (codes)", int(np.median(iss[:, 3] - iss[:, 2])), "cycles")
    allw = tr[:16, 20:50, :]
    print("  over warps: last to finish sweeps minus first, per step (median):", int(np.median(allw[:, :, 1].max(axis=0) - allw[:, :, 1].min(axis=0))))
    rel = []
    for st in range(20, 50):
        last_arrive = allw[:, st - 20, 1].max(); who = int(allw[:, st - 20, 1].argmax())
        release = tr[:16, st + 1, 0].min()
        rel.append((release - last_arrive, who))
    print("  barrier release latency (first warp past the wait minus last arrive), median:", int(np.median([r[0] for r in rel])), " last-arriving warps:", [r[1] for r in rel][:16])
    st = 30
    if os.environ.get("TRACE_ALL"):
        for w in range(16):
            a0 = tr[w, 20:50, 0]; a1 = tr[w, 20:50, 1]; a2 = tr[w, 20:50, 2]; nxt = tr[w, 21:51, 0]
            print(f"  warp {w:2d}: sweeps {np.median(a1 - a0):6.0f}  post {np.median(a2 - a1):6.0f}  wait {np.median(nxt - a2):6.0f}")
    base = tr[:16, st, 0].min()
    print("  step 30, per warp [start, sweeps done, step end] relative to the first start:")
    for w in range(16):
        print(f"    warp {w:2d}: {tr[w, st, 0] - base:5d} {tr[w, st, 1] - base:5d} {tr[w, st, 2] - base:5d}")
    if os.environ.get("TRACE_ALL"):
        for w in (5, 6, 7):
            print(f"  warp {w} sweeps per step 20..35:", [int(v) for v in (tr[w, 20:36, 1] - tr[w, 20:36, 0])])

    if os.environ.get("TRACE_ALL"):
        for w in (5, 6, 7, 8):
            seg = np.median(np.diff(tr[w, 24:50, [0, 4, 5, 6, 7]].T, axis=1), axis=0)
            print(f"  warp {w}: start->sweep1 {seg[0]:.0f}, sweep2 {seg[1]:.0f}, sweep3 {seg[2]:.0f}, sweep4 {seg[3]:.0f}")


    if os.environ.get("TRACE_FINE"):
        ft = tr[17:33]
        for w in (5, 6, 7, 8, 9, 10):
            seg = ft[w, 24:50]
            ok = seg[:, 4] > 0
            if ok.any():
                d = seg[ok]
                print(f"  warp {w} sweep 2: loads+div {np.median(d[:,1]-d[:,0]):.0f}  q {np.median(d[:,2]-d[:,1]):.0f}  tiny check(+fix) {np.median(d[:,6]-d[:,2]):.0f}  P=f(q) {np.median(d[:,3]-d[:,6]):.0f}  updates+sts {np.median(d[:,4]-d[:,3]):.0f}  tiny fixes {int((d[:,5]>0).sum())}/{len(d)}")
