"""CPU (gloo, world_size 2): the host logic of the multi-GPU path -- cell-balanced sharding and the
variable-length all-gather that assembles the store image in rank order."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from muscle_b200 import dist as mdist
from muscle_b200 import synth


def test_shard_ranges_cover_and_balance():
	seqs = synth.make_family(40, 100, 30, seed=3)
	lens = [len(s) for s in seqs]
	for world in (1, 2, 3, 8):
		ranges, cells, total = mdist.shard_ranges(lens, world)
		assert ranges[0][0] == 0 and ranges[-1][1] == 40*39//2
		for a, b in zip(ranges[:-1], ranges[1:]):
			assert a[1] == b[0]
		assert abs(sum(cells) - total) < 1e-6*total
		assert max(cells) <= total/world*1.15 + max(lens)**2
	assert synth.total_cells(seqs) == int(total)


def _free_port():
	s = socket.socket()
	s.bind(("127.0.0.1", 0))
	p = s.getsockname()[1]
	s.close()
	return p


def _worker(rank, world, port, q):
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	# each rank owns a different-length slice of a known vector, like a store shard
	full = torch.arange(1000, dtype=torch.int64)*3 + 1
	cuts = [0, 137, 1000]
	mine = full[cuts[rank]:cuts[rank + 1]]
	got, sizes = mdist.allgather_v(mine)
	ok = bool((got == full).all()) and sizes == [137, 863]
	# EA gather: per-rank EA vectors of a 6-sequence problem
	ranges = [(0, 7), (7, 15)]
	ea_all = np.linspace(0.1, 0.9, 15).astype(np.float32)
	lo, hi = ranges[rank]
	m = mdist.gather_ea(ea_all[lo:hi], ranges, 6)
	iu = np.triu_indices(6, 1)
	ok = ok and np.array_equal(m[iu], ea_all) and np.array_equal(m, m.T)
	# empty shard on one rank
	e, sz = mdist.allgather_v(torch.zeros(0 if rank == 0 else 5, dtype=torch.float32))
	ok = ok and e.numel() == 5 and sz == [0, 5]
	# in-place all-gather-v (what the store / entry-range exchanges use): own part in place, then
	# one broadcast per source rank; an empty part in the middle must not hang anybody
	bounds = [0, 137, 137 if False else 1000]
	buf = torch.zeros(1000, dtype=torch.int64)
	buf[bounds[rank]:bounds[rank + 1]] = full[bounds[rank]:bounds[rank + 1]]
	mdist.allgather_v_inplace(buf, bounds)
	ok = ok and bool((buf == full).all())
	b2 = [0, 0, 10] if rank >= 0 else None          # rank 0 owns nothing
	buf2 = torch.zeros(10, dtype=torch.float32)
	if rank == 1:
		buf2[:] = 7.0
	mdist.allgather_v_inplace(buf2, b2)
	ok = ok and bool((buf2 == 7.0).all())
	q.put((rank, ok))
	dist.destroy_process_group()


def test_allgather_v_gloo_world2():
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs:
		p.start()
	res = [q.get(timeout=120) for _ in range(2)]
	for p in procs:
		p.join(60)
	assert sorted(res) == [(0, True), (1, True)]
