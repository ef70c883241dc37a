import sys, time, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import torch, numpy as np
import pyoracle as po, smoke_simulation_b200 as smk
sc = po.SCENES['C2']; W,H,D = sc[:3]
sim = smk.SmokeSim(W,H,D); po.setup_scene(sim, sc)
host = torch.empty((D,H,W), dtype=torch.float32, pin_memory=True); hp = host.data_ptr()
for t in range(4): sim.step_async(po.tick_dt(t), hp)
sim.sync()
def run(label, fn, K=30):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(K): fn()
    sim.sync(); torch.cuda.synchronize()
    print(label, '%.3f ms/step' % ((time.perf_counter()-t0)*1e3/K))
run('async no host', lambda: sim.step_async(0.05, None))
run('async pipelined host', lambda: sim.step_async(0.05, hp))
run('blocking host', lambda: sim.step_ptr(0.05, hp))
# enqueue time only
t0=time.perf_counter()
for i in range(30): sim.step_async(0.05, hp)
t1=time.perf_counter(); sim.sync(); t2=time.perf_counter()
print('enqueue %.3f ms/step, total %.3f' % ((t1-t0)*1e3/30, (t2-t0)*1e3/30))
