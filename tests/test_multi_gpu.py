"""Real multi-process slab runs (one rank per GPU, NCCL for the plumbing).  Skipped on boxes with fewer than two GPUs;
run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`.  The single-GPU suite covers the same
arithmetic with virtual slabs (test_slab_gpu.py); this covers what only separate processes exercise: CUDA IPC mapping,
the epoch handshake between GPUs, the overlapped halo pulls of smk_step, the NCCL fallback transport."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.parametrize("transport,dims,mode", [("p2p", (64, 48, 40), "plain"), ("p2p", (70, 33, 24), "plain"), ("nccl", (64, 48, 40), "plain"),
                                                 ("p2p", (256, 256, 80), "plain"), ("p2p", (96, 64, 48), "hitch"), ("nccl", (96, 64, 48), "hitch"),
                                                 ("p2p", (96, 64, 48), "reach")])
def test_slabs_across_processes_bit_identical(transport, dims, mode, pass_kernel):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mp_check.py"), transport] + [str(v) for v in dims] + [mode]
    env = dict(os.environ, SMK_PASS_KERNEL=pass_kernel)   # the ranks are separate processes: the process-wide switch
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and ("OK (bit-identical" in r.stdout or "OK (SMK_ERR_REACH" in r.stdout), r.stdout[-2000:] + r.stderr[-2000:]
