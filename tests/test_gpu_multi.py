"""Multi-GPU parity inside pytest: tests/multi_gpu_check.py under torchrun on every GPU of the box (skipped on a
one-GPU box; the same checks at N ranks also run inside bench.py's `parity` object)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_paths_against_the_oracle_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    port = str(30500 + os.getpid() % 2000)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                          "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert out.stdout.count("multi-GPU parity ok") >= 6, out.stdout[-3000:]
