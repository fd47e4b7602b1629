"""GPU parity of the sharded path: ShardedEnsemble (lookup in send layout -> all-to-all -> unpack;
pack -> reverse all-to-all -> update) must equal the single-GPU PreallocationStrategy path bit for
bit.  With one visible GPU this runs as a world-size-1 NCCL group (exercises the kernels and the
layout); tests/dist_gpu_check.py is the same check under torchrun on >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_world1_equals_preallocation():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_gpu_check.py")], capture_output=True, text=True,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                                  MASTER_PORT="29655"), timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dist check ok" in out.stdout


def test_sharded_world2_if_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29656",
                          os.path.join(ROOT, "tests", "dist_gpu_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dist check ok") >= 1
