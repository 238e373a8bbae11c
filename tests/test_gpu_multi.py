"""GPU, >= 2 devices: the sharded pipeline under torchrun/NCCL equals the single-GPU pipeline (tools/check_multi_gpu.py).
Skipped on single-GPU boxes; the exchange kernels are covered there by test_gpu_pipeline.py on one device and the host
logic by the gloo test."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_pipeline_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 4 if n >= 4 else 2
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-GPU parity: OK" in out.stdout
