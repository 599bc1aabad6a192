"""Launched under torchrun with >= 2 GPUs by tests/test_multigpu.py (or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Checks that the sharded pipeline (posteriors on pair ranges -> NCCL all-gather of the sparse store
-> sharded consistency iterations with values-only all-gather) gives, on EVERY rank, bit-identical
results to a single-GPU run of the same engine."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_tables          # noqa: E402
from muscle_b200 import synth             # noqa: E402
from muscle_b200.mpcflat import MPCFlat   # noqa: E402


def main():
	rank = int(os.environ["RANK"])
	local = int(os.environ["LOCAL_RANK"])
	torch.cuda.set_device(local)
	dist.init_process_group("nccl", device_id=torch.device("cuda", local))
	n = int(os.environ.get("MGPU_N", "24"))
	seqs = synth.make_family(n, 110, 30, seed=17)
	t = load_tables()
	M = MPCFlat(t, device=local)
	M.InitSeqs(seqs); M.InitPairs(); M.InitDistMx()
	M.CalcPosteriors()
	offs, ents0 = M.engine.export_all()
	M.Consistency()
	_, ents2 = M.engine.export_all()
	ea = M.m_DistMx.copy()
	# single-GPU reference run of the same library on this rank (no process group involvement)
	S = MPCFlat(t, device=local)
	S._rank, S._world = 0, 1
	S.InitSeqs(seqs); S.InitPairs(); S.InitDistMx()
	S.CalcPosteriors()
	soffs, sents0 = S.engine.export_all()
	S.Consistency()
	_, sents2 = S.engine.export_all()
	ok_ea = np.array_equal(ea, S.m_DistMx)
	bad_off = [p for p in range(len(offs)) if not np.array_equal(offs[p], soffs[p])]
	bad0 = [p for p in range(len(offs)) if ents0[p].tobytes() != sents0[p].tobytes()]
	bad2 = [p for p in range(len(offs)) if ents2[p].tobytes() != sents2[p].tobytes()]
	ok = ok_ea and not bad_off and not bad0 and not bad2
	if not ok:
		print("rank", rank, "ranges", M._ranges, "ea", ok_ea, "bad_off", bad_off[:5], len(bad_off), "bad0", bad0[:5], len(bad0),
		  "bad2", bad2[:5], len(bad2), flush=True)
	flag = torch.tensor([1 if ok else 0], device="cuda")
	dist.all_reduce(flag, op=dist.ReduceOp.MIN)
	if rank == 0:
		print("MGPU_CHECK", "OK" if int(flag.item()) == 1 else "FAIL", "world", dist.get_world_size(), "pairs", len(offs),
		  "ranges", M._ranges, "timings", {k: round(float(v), 3) for k, v in M.timings.items()})
	dist.barrier()
	dist.destroy_process_group()
	sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
	main()
