"""GPU, needs >= 2 devices (skipped on the 1-GPU box): sharded pipeline == single-GPU pipeline."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_identical_to_one():
	import torch
	if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
		pytest.skip("needs 2 GPUs")
	cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
	  "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "mgpu_check.py")]
	r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
	assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
