"""GPU parity across REAL processes — needs two GPUs on the box, skipped otherwise.

One process per GPU under torchrun (tools/check_multigpu.py): CUDA-IPC peer pointers, the library's own cross-GPU flags
(frame push with acknowledgements, the barriers folded into k_gather_foreign / k_raycast_sharded / k_model_maps) and the
foreign-block cache must reproduce, bit for bit and frame by frame, a single context that holds the whole scene.  The
one-GPU emulation (test_gpu_sharding.py) cannot cover that protocol: kernels of different ranks wait on one another."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def gpu_count() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("pipeline", ["1", "0"])
def test_two_processes_equal_single_context(pipeline):
    if gpu_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, TFB_CHECK_FRAMES="8", TFB_SHARD_PIPELINE=pipeline)   # 0: staged calls + tfb_shard_barrier
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tools", "check_multigpu.py")]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "mismatches: 0" in p.stdout
    assert p.stdout.count("pose == raycast == updates ==") == 8
