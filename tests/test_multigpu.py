"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the sharded pipeline == the single-GPU pipeline,
through both front ends -- one process per GPU with NCCL (muscle_b200/dist.py, torchrun) and one
process driving all GPUs (mb200_group_*, tests/test_consistency_gpu.py::test_group_all_devices_equal_engine)."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
	import torch
	return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 3, 8])
def test_ranks_identical_to_one(world):
	if _ngpu() < world:
		pytest.skip("needs %d GPUs" % world)
	cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
	  "--master-addr", "127.0.0.1", "--master-port", str(29511 + world), os.path.join(ROOT, "tests", "mgpu_check.py")]
	r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
	assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
