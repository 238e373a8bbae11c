"""GPU: bench.py end to end at a small size -- the JSON contract (one line, the keys the driver reads) on one GPU, and on
>= 2 GPUs both multi-GPU forms under torchrun with equal result digests."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--db-seqs", "200000", "--queries", "48", "--max-candidates", "500", "--steps", "1", "--warmup", "3", "--no-cpu-baseline"]


def _line(out):
    lines = [l for l in out.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, out.stdout[-2000:] + out.stderr[-3000:]
    return json.loads(lines[0])


def test_bench_line_on_one_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    d = _line(out)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "clocks",
              "roofline", "roofline_align", "roofline_prefilter", "roofline_prefilter_l2", "e2e", "gpu_launches", "parity"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["parity"]["sharded_equals_single"] and len(d["parity"]["digest"]) == 64


def test_bench_both_multi_gpu_forms():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    digests = {}
    for form in ("striped", "exchange"):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
                              os.path.join(ROOT, "bench.py"), "--gpus", "2", "--multi-gpu", form] + SMALL, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stderr[-3000:]
        d = _line(out)
        assert d["n_gpus"] == 2 and d["value"] > 0 and d["parity"]["sharded_equals_single"]
        assert ("striped" in d["config"]["sharding"]) == (form == "striped")
        digests[form] = d["parity"]["digest"]
    assert digests["striped"] == digests["exchange"]
