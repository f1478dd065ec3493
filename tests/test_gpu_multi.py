"""Multi-GPU parity: CFG branch split x Ulysses token shards for the denoise loop, H-slab VAE decode with
halo rows stored into the neighbour's padded buffer -- each compared with the single-GPU path on the same inputs by
tools/mgpu_check.py, one process per GPU under torch.distributed.run.  Skipped on boxes with a single GPU (the
world_size-2 host logic is covered on CPU by tests/test_parallel_cpu.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_paths_match_single_gpu(cuda, n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs, {torch.cuda.device_count()} visible")
    env = dict(os.environ)
    env.pop("RANK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(29540 + n), str(ROOT / "tools" / "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
    assert "ALL OK" in r.stdout
