#!/usr/bin/env python
"""bench.py -- voxel-steps/s of the full smoke step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C3|C1|NxNxN]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full simulation tick (source/mask fill, forcing, clamp, 30 x 2 red-black SOR half-sweeps,
u/v/w advection, density advection) on synthetic plume scenes (no RNG: fields start at
zero, sources/obstacles are analytic spheres).  N = 1 runs configs[1] (256^3 rising plume) with the
reference's own solver schedule (RBGS, omega 1.9, 30 iterations -- the only schedule that has a reference
to be identical to; BASELINE's "40 Jacobi iterations" has no counterpart in the reference, SURVEY.md point 1).

One JSON line on stdout (rank 0):
  value     whole-job voxel-steps/s, fields resident in HBM, no host traffic in the timed region
  e2e       the same through the reference-facing call (simulate(): smk_step with a HOST density buffer,
            the device->host copy of every step inside the timed region)
  roofline  the dominant kernel (pressure half-sweep): algorithmic bytes per launch / its mean duration
            (CUDA events recorded by the library on its stream during the timed region) / measured HBM peak
  cpu_baseline   the reference's kernel bodies as an OpenMP host loop (oracle/_ref/libref_cpu.so; "port" =
            oracle/liboracle.so if the former is absent) on a bounded sample, rank 0, N = 1 only
--impl reference times that CPU implementation alone (all host threads) and prints the same line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "voxel_steps_per_sec"
UNIT = "voxel-steps/s"
BYTES_PER_CELL_HALFSWEEP = 25  # u,v,w read+write (24 B) + 1 B mask information per cell


def ncu_traffic(wname, hs_per_launch, kname="reg"):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` captures: profiles/ncu_traffic.json maps "<workload>:<half-sweeps per launch>" to
    {"bytes": ..., "source": "profiles/<capture summary>"}.  None when no capture of that configuration is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        key = f"{wname}:{int(round(hs_per_launch))}"
        return t.get(f"{key}:{kname}") if kname != "reg" else t.get(key)
    except Exception:
        return None


def parse_workload(name, gpus):
    """N = 1: configs[1] (C2, 256^3) -- the configuration BASELINE.json's metric is quoted on that fits one GPU.
    N > 1: configs[4] (C5): weak scaling, 512 x 512 x (512 N), one 512^3 slab per GPU."""
    from smoke_simulation_b200 import scenes as po
    if name is None:
        name = "C2" if gpus <= 1 else "C5"
    if name == "C5":
        G = max(gpus, 1)
        sc = (512, 512, 512 * G, -9.82, 15.0, [(256, 64, 256 * G, 32)], [])
        return name, sc, f"C5 weak scaling 512x512x{512 * G}: {G} z-slabs of 512 planes, source (256,64,{256 * G}) r32, alpha 15"
    if name in po.SCENES:
        sc = po.SCENES[name]
        label = {"C1": "C1 80^3 default scene (source r5 + solid sphere r13)",
                 "C2": "C2 256^3 open-boundary rising plume (source (128,32,128) r16, alpha 15)",
                 "C3": "C3 512^3 plume with solid-sphere obstacle",
                 "C4": "C4 1024^3 inverted-gravity plume"}[name]
        return name, sc, label
    dims = [int(v) for v in name.lower().split("x")]
    W, H, D = dims if len(dims) == 3 else (dims[0],) * 3
    sc = (W, H, D, -9.82, 15.0, [(W / 2, H / 8, D / 2, max(2.0, W / 16))], [])
    return name, sc, f"{W}x{H}x{D} rising plume"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, 100 ms)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v and (v & (v - 1)) == 0:
                    names[v] = n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit and "None" not in n and "GpuIdle" not in n:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_engine():
    """The reference's own CPU implementation of the path (preferred) or the oracle port."""
    import pyoracle as po
    if po.have_ref_cpu():
        return po.RefCPU, "reference", po
    return (lambda W, H, D: po.Oracle(W, H, D, contract=0)), "port", po


def time_cpu(scene, budget_s, steps=None, warmup=0, solver=None):
    """Time the CPU implementation on a bounded sample of `scene`: full x,y extent, the first Ds cell planes
    (Ds halved until the estimated run fits `budget_s`).  Returns (voxel-steps/s, cores, kind, sample, ms_per_step)."""
    eng, kind, po = cpu_engine()
    if solver is not None:   # the Jacobi extension has no reference implementation: time the oracle port of it
        kind = "port"
        def eng(W_, H_, D_, _s=solver):
            e_ = po.Oracle(W_, H_, D_); e_.set_solver(1, _s[1]); return e_
    W, H, D = scene[:3]
    cores = os.cpu_count() or 1
    if kind == "reference":
        po.RefCPU.set_threads(cores)
    else:
        po.Oracle.set_threads(cores)
    # probe on a thin slab to estimate the cost per voxel-step
    Dp = min(D, 16)
    sub = (W, H, Dp) + tuple(scene[3:])
    e = eng(W, H, Dp); po.setup_scene(e, sub)
    e.step(0.01)
    t0 = time.perf_counter(); e.step(0.05); est = (time.perf_counter() - t0) / (W * H * Dp)
    if hasattr(e, "close"):
        e.close()
    n_steps = (steps or 0) + warmup
    Ds = D
    if steps is None:  # choose the number of ticks for ~budget on the full volume, else shrink the volume
        while Ds > 16 and est * W * H * Ds * 2 > budget_s:
            Ds //= 2
        steps = max(1, min(20, int(budget_s / (est * W * H * Ds)) - 1))
        warmup = 1
    else:
        while Ds > 16 and est * W * H * Ds * n_steps > budget_s:
            Ds //= 2
    # sources / obstacles above the sampled planes are moved into the sample (C5's source sits in the middle of the domain)
    fit = lambda objs: [(x, y, z if z < Ds - 2 else Ds / 2, r) for (x, y, z, r) in objs]
    sub = (W, H, Ds, scene[3], scene[4], fit(scene[5]), fit(scene[6]))
    e = eng(W, H, Ds); po.setup_scene(e, sub)
    for t in range(warmup):
        e.step(po.tick_dt(t))
    t0 = time.perf_counter()
    for t in range(steps):
        e.step(po.tick_dt(t + warmup))
    dt = time.perf_counter() - t0
    if hasattr(e, "close"):
        e.close()
    sample = (f"{steps} ticks of {W}x{H}x{Ds}" + ("" if Ds == D else f" (first {Ds} of {D} z-planes of the workload)")
              + f", {cores} OpenMP threads")
    return W * H * Ds * steps / dt, cores, kind, sample, dt / steps * 1e3


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: pin this process (and so the first-touch placement of its pinned readback buffer) to the CPU
    cores NVML reports as local to its GPU, so that 8 concurrent device->host copies do not all land on one socket's
    memory.  Best effort; returns the number of cores bound to or None."""
    if os.environ.get("SMK_BENCH_NO_AFFINITY"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def time_reference_gpu(scene, ticks=10):
    """The reference's OWN kernels (unmodified smokeSimulation.cu built headless for sm_100a, oracle/_ref/libref_gpu.so) on
    the same GPU, kernels only, same scene -- the "beat THAT kernel on the same box" bar of SURVEY.md section 8(d)."""
    import pyoracle as po
    if not po.have_ref_gpu() or scene[0] * scene[1] * scene[2] > 600 ** 3:   # the reference overflows int sizes from 812^3
        return None
    try:
        r = po.RefGPU(*scene[:3]); po.setup_scene(r, scene)
        r.step(0.01); r.step(0.05); r.sync()
        ms = r.time_kernels(0.05, ticks) / ticks
        r.close()
        W, H, D = scene[:3]
        return {"value": W * H * D / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "what": "68 launches per step of the reference's kernels, device time by CUDA events, no host round trip"}
    except Exception as ex:   # pragma: no cover
        return {"unavailable": str(ex)}


def workload_config(label, scene, solver, iters, world, explicit):
    """The `config` object: identical for the b200 arm and the reference arm (same workload, same schedule)."""
    W, H, D = scene[:3]
    per_gpu = (2 * W * H * D * 4 + 9 * (W + 1) * (H + 1) * (D + 1) * 4 + 2 * W * H * D) / max(world, 1) / 1e6
    return {"workload": label + (f"; EXTENSION damped Jacobi (2/3) x{iters}, not a reference path" if solver == "jacobi"
                                 else f"; reference schedule RBGS omega=1.9 x{iters}"),
            "grid": [W, H, D], "solver": solver, "iterations": iters,
            "dt": "tick 0: 0.01, then 0.05 (main.cpp:293, :93)", "parallelism": f"zslab{world}",
            "scaling_pattern": "one GPU" if world == 1 else ("strong (fixed grid)" if explicit else "weak (C5: one 512^3 slab per GPU)"),
            "l2": (f"state per GPU {per_gpu:.0f} MB (> 126 MB L2): inputs larger than L2, no explicit flush" if per_gpu > 126 else
                   f"state per GPU {per_gpu:.0f} MB fits the 126 MB L2: launch/L2-bound case (SURVEY H7), no flush, reported as such")}


def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores, same config / metric / unit as the b200 arm.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    if rank != 0:
        return
    n = max(world, args.gpus)
    wname, scene, label = parse_workload(args.workload, n)
    explicit = args.workload is not None
    val, cores, kind, sample, ms = time_cpu(scene, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if (explicit and n > 1) else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(label, scene, "rbgs", args.iterations or 30, n, explicit),
        "what": "the reference's kernel bodies (cu:251-705) as an OpenMP host loop, all host threads, on a bounded sample",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pressure_roofline(times, K, sweeps, cells_local, wname, world, peak, peak_src, jac, bal, kname="reg"):
    """Roofline object of the dominant kernel (the pressure pass) from the library's own CUDA-event stage timers."""
    p_ms, p_launches = times["pressure"]
    per_launch_ms = p_ms / max(p_launches, 1)
    hs_per_launch = sweeps * K / max(p_launches, 1)
    compulsory_bytes = BYTES_PER_CELL_HALFSWEEP * cells_local   # u,v,w read + written once, 1 B of mask information
    alg_bytes = compulsory_bytes * hs_per_launch                # section 8(d): 25 B per cell and HALF-SWEEP x half-sweeps per launch
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    tr = ncu_traffic(wname, hs_per_launch, kname) if world == 1 else None
    return {
        "bound": "hbm",
        "kernel": ((f"k_jacobi_bal ({bal} CTAs, balanced piece lists)" if bal else "k_jacobi") +
                   ": one damped-Jacobi iteration per launch (extension)" if jac else
                   (f"balanced piece lists on {bal} CTAs, " if bal else "") +
                   ("k_pressure_tma<4,16> (TMA-staged planes, kernels_pressure_tma.cuh)" if kname == "tma" else
                    "k_pressure_reg<4,16> (kernels_pressure_reg.cuh)") +
                   ": fused pressure pass, 4 red/black SOR half-sweeps per launch (temporal blocking, register-resident u,w)"
                   if hs_per_launch > 1.5 else "k_pressure_half: one red/black SOR half-sweep per launch"),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
        "peak_source": peak_src, "launch_ms": per_launch_ms, "halfsweeps_per_launch": hs_per_launch,
        "algorithmic_bytes_per_launch": alg_bytes,
        "compulsory_bytes_per_launch": compulsory_bytes,
        "compulsory_GBps": achieved / hs_per_launch, "frac_compulsory": achieved / hs_per_launch / peak,
        "note": "achieved = 25 B/cell/half-sweep (SURVEY 8(d)) x half-sweeps per launch / launch time; a fused launch moves "
                "only the compulsory bytes through HBM (traffic), so achieved can exceed the HBM peak: that is the "
                "temporal-blocking win; frac_compulsory is the measured-DRAM fraction the north star's 70 % refers to",
        "stage_ms_per_step": {k: v[0] / K for k, v in times.items()},
    }


def time_single_gpu(smk, po, torch, scene, stream, K, Wm, with_e2e, iters, jac, fuse):
    """A fresh single-GPU simulation of `scene`: Wm warm-up ticks, K timed ticks device-resident and (optionally) K through
    the blocking host-readback call.  Returns (ms_per_step, e2e_ms_per_step or None, stage times, sim)."""
    W, H, D = scene[:3]
    sim = smk.SmokeSim(W, H, D)
    po.setup_scene(sim, scene)
    sim.set_solver(1 if jac else 0, iters, fuse)
    sim.set_stream(stream.cuda_stream)
    tick = 0
    for _ in range(Wm):
        sim.step_async(po.tick_dt(tick)); tick += 1
    sim.sync(); sim.reset_timers()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        sim.step_async(po.tick_dt(tick)); tick += 1
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    times = sim.stage_times()
    ms_e2e = None
    if with_e2e:
        host = torch.empty((D, H, W), dtype=torch.float32, pin_memory=True)
        sim.step_ptr(po.tick_dt(tick), host.data_ptr()); tick += 1
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(K):
            sim.step_ptr(po.tick_dt(tick), host.data_ptr()); tick += 1
        ms_e2e = (time.perf_counter() - w0) * 1e3 / K
        del host
    return ms, ms_e2e, times, sim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None,
                    help="C1|C2|C3|C4|C5 or NxNxN (default: C2 on one GPU; C5 = 512x512x(512 N) weak scaling on N > 1). "
                         "An explicit workload on N > 1 GPUs is a STRONG-scaling run (the grid is cut into N z-slabs)")
    ap.add_argument("--solver", default="rbgs", choices=["rbgs", "jacobi"],
                    help="rbgs = the reference schedule (headline); jacobi = damped-Jacobi extension, not a reference path")
    ap.add_argument("--iterations", type=int, default=None, help="default 30 (rbgs, cu:797) / 40 (jacobi, BASELINE configs[1])")
    ap.add_argument("--fuse", type=int, default=0, help="half-sweeps fused per pressure launch (0 = library default)")
    ap.add_argument("--ghost", type=int, default=8, help="ghost planes per interior slab side (multi-GPU)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="multi-GPU halo transport")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the bit-identity check against a single-GPU run of the same domain")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (c3_512, halo baseline, pageable / pipelined e2e)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import smoke_simulation_b200 as smk
    from smoke_simulation_b200 import scenes as po   # scene definitions only; the oracle is loaded by the cpu_baseline leg alone

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the smoke step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    explicit = args.workload is not None
    wname, scene, label = parse_workload(args.workload, world)
    W, H, D = scene[:3]
    # the library runs the whole step on ONE stream; hand it a real (non-default) torch stream so that
    # torch.cuda.Event brackets exactly the work of the step and the NCCL halo traffic is ordered on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    transport = None
    transport_name = None
    if world > 1:
        sim = smk.SmokeSim(W, H, D, slab=(rank, world), ghost=args.ghost)
        transport_name = args.transport
        if args.transport == "p2p":     # peer-mapped memory over NVLink (CUDA IPC): no Python, no NCCL in the step
            ok = 1
            try:
                smk.slab.attach_peers_ipc(sim, rank, world, dist, torch.device("cuda", local_rank))
            except Exception as ex:   # e.g. no peer access between two GPUs: every rank falls back together
                print(f"[bench rank {rank}] peer-memory attach failed ({ex}); falling back to NCCL", file=sys.stderr)
                ok = 0
            tok = torch.tensor([ok], device="cuda"); dist.all_reduce(tok, op=dist.ReduceOp.MIN)
            if int(tok.item()) == 0:
                sim.close()
                sim = smk.SmokeSim(W, H, D, slab=(rank, world), ghost=args.ghost)
                transport_name = "nccl"
        if transport_name == "nccl":    # torch.distributed send/recv (NCCL) through the transport callback
            transport = smk.slab.TorchTransport(rank, world)
            sim.set_exchange(transport)
    else:
        sim = smk.SmokeSim(W, H, D)
    po.setup_scene(sim, scene)
    jac = args.solver == "jacobi"
    iters = args.iterations if args.iterations is not None else (40 if jac else 30)
    sweeps = iters if jac else 2 * iters          # launches of the unfused solver per step
    sim.set_solver(1 if jac else 0, iters, args.fuse)
    sim.set_stream(stream.cuda_stream)
    c0, c1 = D * rank // world, D * (rank + 1) // world   # owned cell planes of this rank
    host = torch.empty((c1 - c0, H, W), dtype=torch.float32, pin_memory=True)
    host_ptr = host.data_ptr() - c0 * W * H * 4            # smk_step writes the OWNED planes at their global offset

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
        return v

    K, Wm = args.steps, args.warmup
    tick = 0
    for _ in range(Wm):
        sim.step_async(po.tick_dt(tick)); tick += 1
    sim.sync()

    # ---- timed region 1: device-resident throughput ("value") ------------------------------------------------
    sim.reset_timers()
    l0 = sim.launch_count()
    x0 = sim.exchange_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for _ in range(K):
            sim.step_async(po.tick_dt(tick)); tick += 1
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sim.launch_count() - l0
    exchanges = sim.exchange_count() - x0
    times = sim.stage_times()
    value = W * H * D * K / (ms * 1e-3)

    # ---- timed region 2: end to end through the reference-facing call (host buffer, D2H every step) --------
    for _ in range(min(Wm, 2)):
        sim.step_ptr(po.tick_dt(tick), host_ptr); tick += 1
    barrier()
    rb0 = sim.readback_bytes()
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        sim.step_ptr(po.tick_dt(tick), host_ptr); tick += 1   # blocking, like simulate() (cu:814)
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3))  # blocking call: host time is user-visible cost
    e2e = W * H * D * K / (ms_e2e * 1e-3)
    d2h = (sim.readback_bytes() - rb0) / K                     # what the library actually copied, this rank
    if world > 1:
        t = torch.tensor([d2h], device="cuda", dtype=torch.float64); dist.all_reduce(t); d2h = float(t.item())

    checksum = float(host.sum(dtype=torch.float64))  # after the blocking loop: the density of its last step
    if world > 1:
        tc = torch.tensor([checksum], device="cuda", dtype=torch.float64); dist.all_reduce(tc); checksum = float(tc.item())

    extras = {}
    if not args.no_extras:
        # ---- extra: the same call with a PAGEABLE caller buffer, which is what the reference's caller owns
        # (std::vector<float> m_grid, boundingBox.h:41); the library page-locks it on first use (smoke_b200.h)
        pageable = np.empty((c1 - c0, H, W), dtype=np.float32)
        pg_ptr = pageable.ctypes.data - c0 * W * H * 4
        sim.step_ptr(po.tick_dt(tick), pg_ptr); tick += 1
        barrier()
        w0 = time.perf_counter()
        for _ in range(K):
            sim.step_ptr(po.tick_dt(tick), pg_ptr); tick += 1
        barrier()
        ms_pg = max_over_ranks((time.perf_counter() - w0) * 1e3)
        extras["pageable_caller_buffer"] = {"value": W * H * D * K / (ms_pg * 1e-3), "ms_per_step": ms_pg / K,
                                            "api": "smk_step(sim, dt, numpy/std::vector memory): registered by the library on first use"}
        sim.unregister_host(pageable)
        # ---- extra: pipelined readback of smk_step_async (snapshot + copy on a second stream)
        sim.step_async(po.tick_dt(tick), host_ptr); tick += 1   # creates the copy stream / snapshot buffer
        sim.sync()
        barrier()
        w0 = time.perf_counter()
        for _ in range(K):
            sim.step_async(po.tick_dt(tick), host_ptr); tick += 1
        sim.sync()
        barrier()
        ms_pipe = max_over_ranks((time.perf_counter() - w0) * 1e3)
        extras["pipelined"] = {"value": W * H * D * K / (ms_pipe * 1e-3), "ms_per_step": ms_pipe / K,
                               "api": "smk_step_async(sim, dt, host_density): device snapshot + D2H on a second stream, "
                                      "overlapped with the next step; wall clock around K steps + smk_sync"}

        # ---- extra (one GPU): sparse blocking readback (smk_set_readback_box): only the rows that can hold smoke are
        # copied, the caller's buffer ends up identical to the full copy as long as the caller does not write to it
        if world == 1:
            sim.set_readback_box(1)
            for _ in range(2):
                sim.step_ptr(po.tick_dt(tick), host_ptr); tick += 1   # the first call copies everything
            rbs = sim.readback_bytes()
            w0 = time.perf_counter()
            for _ in range(K):
                sim.step_ptr(po.tick_dt(tick), host_ptr); tick += 1
            ms_sp = (time.perf_counter() - w0) * 1e3
            extras["sparse_readback"] = {"value": W * H * D * K / (ms_sp * 1e-3), "ms_per_step": ms_sp / K,
                                         "d2h_bytes_per_step": (sim.readback_bytes() - rbs) / K,
                                         "api": "smk_set_readback_box(sim, 1) + smk_step(sim, dt, host_density): blocking, same buffer "
                                                "contents as the full copy (tests/test_parity_gaps_gpu.py), opt-in because the caller "
                                                "must not write to the buffer between steps"}
            sim.set_readback_box(0)

    # ---- N > 1: bit-identity with the single-GPU run of the SAME domain, at the bench size --------------------
    # Every rank hashes its owned planes on the device (smk_hash_owned); rank 0 then steps the whole domain on its own
    # GPU for the same ticks and hashes the same plane ranges (smk_hash_range).
    parity = None
    halo = None
    if world > 1:
        mine = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in sim.hash_owned()], device="cuda", dtype=torch.int64)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        geoms = [smk.slab.geometry(W, H, D, world, r, args.ghost) for r in range(world)]
        if rank == 0 and not args.no_verify:
            try:
                ref = smk.SmokeSim(W, H, D)
                po.setup_scene(ref, scene); ref.set_solver(1 if jac else 0, iters, args.fuse); ref.set_stream(stream.cuda_stream)
                for t in range(tick):
                    ref.step_async(po.tick_dt(t))
                ref.sync()
                bad = []
                for r, g in enumerate(geoms):
                    want = ref.hash_range(g["own_node_lo"], g["own_node_hi"] + 1, g["c0"], g["c1"])
                    got = [int(v) & ((1 << 64) - 1) for v in allh[r].tolist()]
                    bad += [f"rank{r}:{n}" for n, a_, b_ in zip(smk.SmokeSim.HASH_NAMES, got, want) if a_ != b_]
                ref.close()
                parity = {"result": "identical" if not bad else "MISMATCH", "mismatches": bad, "ticks": tick,
                          "what": f"64-bit device hashes of every rank's owned planes of u,v,w (now, past) and the density after {tick} "
                                  f"ticks == the same plane ranges of a single-GPU run of the whole {W}x{H}x{D} domain",
                          "hash_rank0": [f"{int(v) & ((1 << 64) - 1):016x}" for v in allh[0].tolist()]}
            except Exception as ex:   # e.g. the whole domain does not fit one GPU
                parity = {"result": "unverified", "why": str(ex)}
        # ---- halo cost: the same per-GPU slab as a stand-alone single-GPU problem, on rank 0's GPU ------------------
        if rank == 0 and not args.no_extras:
            Dl = c1 - c0
            sub = (W, H, Dl, scene[3], scene[4], [(x, y, min(z, Dl - 2.0), r) for (x, y, z, r) in scene[5]], [])
            ms1, _, t1, s1 = time_single_gpu(smk, po, torch, sub, stream, max(3, K // 2), 2, False, iters, jac, args.fuse)
            s1.close()
            nb = 2 * (world - 1)     # slab faces with a neighbour
            halo = {"ms_per_step": ms / K - ms1, "single_gpu_same_slab_ms_per_step": ms1,
                    "pressure_ms_per_step": times["pressure"][0] / K, "pressure_ms_per_step_single": t1["pressure"][0] / max(3, K // 2),
                    "advect_ms_per_step": (times["advect_vel"][0] + times["advect_smoke"][0]) / K,
                    "advect_ms_per_step_single": (t1["advect_vel"][0] + t1["advect_smoke"][0]) / max(3, K // 2),
                    "exchanges_per_step": exchanges / K,
                    "peer_read_bytes_per_pass_per_face": 4 * 3 * (W + 8) * (H + 1) * 4,
                    "pulled_bytes_per_step_per_face": (args.ghost + 1) * 3 * (W + 8) * (H + 1) * 4 + args.ghost * W * H * 4,
                    "faces": nb,
                    "what": "ms_per_step = step time on N GPUs (max over ranks) minus the time of ONE slab-sized domain stepped alone "
                            "on one GPU: in-kernel neighbour reads of the passes, epoch handshakes, ghost pulls and boundary strips"}
    barrier()

    if rank == 0:
        peak, peak_src = load_peaks()
        cells = W * H * D
        cells_local = W * H * (c1 - c0)                      # one launch of the dominant kernel covers one slab
        bal = sim.last_pass_ctas()
        roofline = pressure_roofline(times, K, sweeps, cells_local, wname, world, peak, peak_src, jac, bal, sim.last_pass_kernel())
        cfg = workload_config(label, scene, args.solver, iters, world, explicit)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if (explicit and world > 1) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "impl_config": {"fuse": args.fuse, "ghost": args.ghost if world > 1 else 0, "transport": transport_name,
                            "cpu_affinity_cores": numa, "halo_exchanges_per_step": exchanges / K},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 544, "d2h_bytes_per_step": d2h,
                    "d2h_bytes_counted": "by the library where it enqueues the copies (smk_readback_bytes), summed over ranks",
                    "ms_per_step": ms_e2e / K, "api": "smk_step(sim, dt, host_density) == simulate(smoke_grid, dt)",
                    "note": "per-step inputs are the scene objects + dt/gravity/buoyancy, passed as kernel parameters; "
                            "every rank copies its owned planes of the new density to pinned host memory",
                    "density_checksum": checksum, **extras},
            "gpu_launches": launches * world,
            "clocks": clk.summary(),
            "roofline": roofline,
        }
        if parity is not None:
            line["parity"] = parity
        if halo is not None:
            line["halo"] = halo
        if world == 1 and wname == "C2" and not args.no_extras and not jac:
            # BASELINE configs[2] on one GPU (the north star's "512^3"): value, e2e and the pass's roofline
            try:
                sc3 = po.SCENES["C3"]
                k3 = max(3, K // 4)
                ms3, ms3e, t3, s3 = time_single_gpu(smk, po, torch, sc3, stream, k3, 2, True, iters, False, args.fuse)
                k3name = s3.last_pass_kernel()
                s3.close()
                n3 = sc3[0] * sc3[1] * sc3[2]
                line["c3_512"] = {"workload": "C3 512^3 plume with solid-sphere obstacle; reference schedule RBGS omega=1.9 x30",
                                  "value": n3 / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3, "steps": k3, "warmup": 2,
                                  "e2e": {"value": n3 / (ms3e * 1e-3), "ms_per_step": ms3e, "d2h_bytes_per_step": n3 * 4},
                                  "roofline": pressure_roofline(t3, k3, sweeps, n3, "C3", 1, peak, peak_src, False, 0, k3name)}
            except Exception as ex:   # pragma: no cover
                line["c3_512"] = {"unavailable": str(ex)}
        if not args.no_cpu_baseline and world == 1:
            val, cores, kind, sample, _ = time_cpu(scene, budget_s=20.0, solver=("jacobi", iters) if jac else None)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
            ref_gpu = None if jac else time_reference_gpu(scene)
            if ref_gpu:
                line["reference_gpu_kernels"] = ref_gpu
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
