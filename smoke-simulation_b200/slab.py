"""Host-side plumbing of the multi-GPU z-slab run: one process per GPU.

The decomposition and the per-step schedule are decided by the C++ library (csrc/slab_plan.h); this module only
  * mirrors them for callers (``geometry``, ``plan``, ``regions`` -- host-only C ABI calls, no GPU needed),
  * provides the halo transport the library calls back into: ``TorchTransport`` moves the contiguous plane ranges
    with ``torch.distributed`` point-to-point operations (NCCL over NVLink on the B200 box, gloo in the CPU tests).

The reference has no multi-GPU path; correctness bar: results bit-identical to the single-GPU run.
"""
import ctypes as C

import numpy as np

from . import binding

OP_NAMES = ("flip", "fill", "force", "pressure", "advect_vel", "advect_smoke", "exchange")
SET_VEL_NOW, SET_SMOKE_NOW = 0, 1


def geometry(W, H, D, world, rank, ghost):
    """dict(c0, c1, zlo, zhc, ghost, ok, own_node_lo, own_node_hi) of slab `rank`."""
    L = binding.load_library()
    out = (C.c_int * 8)()
    rc = L.smk_slab_geometry(W, H, D, world, rank, ghost, out)
    if rc != 0:
        raise binding.SmokeError(f"smk_slab_geometry failed ({rc})")
    keys = ("c0", "c1", "zlo", "zhc", "ghost", "ok", "own_node_lo", "own_node_hi")
    return dict(zip(keys, (int(v) for v in out)))


def plan(W, H, D, world, rank, ghost, iterations=30, fuse=4, steps=1):
    """The operation list smk_step executes: [(name, a, b, p0, p1), ...] for `steps` consecutive steps."""
    L = binding.load_library()
    n = L.smk_slab_plan(W, H, D, world, rank, ghost, iterations, fuse, steps, None, 0)
    if n < 0:
        raise binding.SmokeError(f"smk_slab_plan failed ({n})")
    buf = (C.c_int * (5 * n))()
    L.smk_slab_plan(W, H, D, world, rank, ghost, iterations, fuse, steps, buf, n)
    a = np.frombuffer(buf, dtype=np.int32).reshape(n, 5)
    return [(OP_NAMES[int(r[0])], int(r[1]), int(r[2]), int(r[3]), int(r[4])) for r in a]


def plan_p2p(W, H, D, world, rank, ghost, iterations=30, fuse=4, steps=1):
    """Like plan(), for the peer-memory path (pressure passes read the neighbours directly: no exchange for them)."""
    L = binding.load_library()
    n = L.smk_slab_plan_p2p(W, H, D, world, rank, ghost, iterations, fuse, steps, None, 0)
    if n < 0:
        raise binding.SmokeError(f"smk_slab_plan_p2p failed ({n})")
    buf = (C.c_int * (5 * n))()
    L.smk_slab_plan_p2p(W, H, D, world, rank, ghost, iterations, fuse, steps, buf, n)
    a = np.frombuffer(buf, dtype=np.int32).reshape(n, 5)
    return [(OP_NAMES[int(r[0])], int(r[1]), int(r[2]), int(r[3]), int(r[4])) for r in a]


def pass_schedule(W, H, out_lo, out_hi, K=4, nctas=148):
    """How a fused pressure pass over the node planes [out_lo, out_hi) is cut into pieces of equal cost for `nctas`
    CTAs (csrc/pass_schedule.h, host-only): dict(tiles=(tx, ty), cost=z-steps of the busiest CTA,
    ctas=[[(tile x, tile y, zo0, zo1), ...] per CTA])."""
    L = binding.load_library()
    info = (C.c_int * 4)()
    n = L.smk_pass_schedule(W, H, out_lo, out_hi, K, nctas, None, 0, None, 0, info)
    if n < 0:
        raise binding.SmokeError(f"smk_pass_schedule failed ({n})")
    pieces = (C.c_int * (4 * max(n, 1)))()
    first = (C.c_int * (int(info[3]) + 1))()
    L.smk_pass_schedule(W, H, out_lo, out_hi, K, nctas, pieces, n, first, int(info[3]) + 1, info)
    pc = np.frombuffer(pieces, dtype=np.int32).reshape(-1, 4)[:n]
    fs = [int(v) for v in first]
    ctas = [[tuple(int(v) for v in pc[i]) for i in range(fs[b], fs[b + 1])] for b in range(int(info[3]))]
    return dict(tiles=(int(info[0]), int(info[1])), cost=int(info[2]), ctas=ctas)


def regions(W, H, D, world, rank, ghost, set_id):
    """Halo regions of one exchange: [(side, send_lo, send_n, recv_lo, recv_n), ...] in global plane indices."""
    L = binding.load_library()
    buf = (C.c_int * 10)()
    n = L.smk_slab_regions(W, H, D, world, rank, ghost, set_id, buf, 2)
    a = np.frombuffer(buf, dtype=np.int32).reshape(2, 5)[:n]
    return [tuple(int(v) for v in r) for r in a]


def attach_peers_ipc(sim, rank, world, dist, device):
    """Multi-process setup of the peer-memory halo path: all-gather the 64-byte CUDA IPC handles of the slabs' arenas
    (torch.distributed is only the plumbing) and map the two neighbours.  After this smk_step needs no transport."""
    import torch
    mine = torch.frombuffer(bytearray(sim.p2p_export()), dtype=torch.uint8).to(device)
    allh = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(allh, mine)
    if rank > 0:
        sim.p2p_attach_ipc(0, bytes(allh[rank - 1].cpu().numpy().tobytes()))
    if rank < world - 1:
        sim.p2p_attach_ipc(1, bytes(allh[rank + 1].cpu().numpy().tobytes()))


def attach_peers_local(sims):
    """Same for slabs that live in ONE process (virtual slabs on one GPU, tests): plain pointers."""
    for r, s in enumerate(sims):
        if r > 0:
            s.p2p_attach_ptr(0, sims[r - 1].p2p_arena())
        if r < len(sims) - 1:
            s.p2p_attach_ptr(1, sims[r + 1].p2p_arena())


class _DevBuf:
    """A raw device range exposed through __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class TorchTransport:
    """Halo exchange over torch.distributed (backend nccl on GPUs).  Neighbours: rank-1 (side 0), rank+1 (side 1).
    All sends/receives of one exchange are posted as one batch, ordered on the library's stream."""

    def __init__(self, rank, world, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.rank, self.world, self.group = torch, dist, rank, world, group
        self.exchanges = 0
        self.bytes_sent = 0

    def __call__(self, set_id, regs, stream_ptr):
        torch, dist = self.torch, self.dist
        ext = torch.cuda.ExternalStream(int(stream_ptr)) if stream_ptr else torch.cuda.current_stream()
        ops = []
        with torch.cuda.stream(ext):
            for k, (side, send_ptr, recv_ptr, send_bytes, recv_bytes) in enumerate(regs):
                peer = self.rank - 1 if side == 0 else self.rank + 1
                if send_bytes:
                    ops.append(dist.P2POp(dist.isend, torch.as_tensor(_DevBuf(send_ptr, send_bytes), device="cuda"), peer, self.group))
                    self.bytes_sent += send_bytes
                if recv_bytes:
                    ops.append(dist.P2POp(dist.irecv, torch.as_tensor(_DevBuf(recv_ptr, recv_bytes), device="cuda"), peer, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        self.exchanges += 1
        return 0


def exchange_arrays(dist, rank, regs, fields, group=None):
    """CPU / gloo flavour used by the tests: `fields` are full-domain torch CPU tensors indexed [z, ...]; the regions
    are the plane ranges of ``regions()``.  Lower neighbour first / upper neighbour second ordering avoids deadlock
    with blocking gloo send/recv by posting everything as one batch."""
    import torch
    ops, recvs = [], []
    for (side, send_lo, send_n, recv_lo, recv_n) in regs:
        peer = rank - 1 if side == 0 else rank + 1
        for f in fields:
            if send_n:
                ops.append(dist.P2POp(dist.isend, f[send_lo:send_lo + send_n].contiguous(), peer, group))
            if recv_n:
                buf = torch.empty_like(f[recv_lo:recv_lo + recv_n])
                recvs.append((f, recv_lo, recv_n, buf))
                ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for f, lo, n, buf in recvs:
        f[lo:lo + n] = buf
