"""Multi-GPU plumbing for the MPCFlat pair engine.

One process per GPU (torchrun), NCCL over NVLink for the two exchange steps the path really has:
  1. after the posterior stage: all-gather of the variable-length sparse blocks (row offsets and
     {P,col} entries of each rank's contiguous pair range), so every rank holds the full store;
  2. after each consistency iteration: all-gather of the updated VALUES only (the pattern is
     invariant, mysparsemx.cpp:87-113).
The posterior stage itself needs no collective: pairs are independent and are sharded in
contiguous, cell-balanced ranges of the reference's row-major pair order, which makes the
all-gather-v a plain concatenation in rank order.
All functions take torch tensors and a process group, so the host logic is testable on CPU with
the gloo backend (tests/test_dist_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(lens, world):
	"""Contiguous ranges [lo,hi) of the row-major (i<j) pair list with ~equal DP cells per rank.
	-> (ranges, cells_per_rank, total_cells)"""
	L = np.asarray(lens, np.float64)
	n = len(L)
	iu, ju = np.triu_indices(n, 1)
	cost = np.cumsum(L[iu]*L[ju])
	total = float(cost[-1])
	cuts = [0]
	for r in range(1, world):
		cuts.append(int(np.searchsorted(cost, total*r/world)))
	cuts.append(len(iu))
	for r in range(1, len(cuts)):
		cuts[r] = max(cuts[r], cuts[r - 1])
	cells = []
	for r in range(world):
		lo, hi = cuts[r], cuts[r + 1]
		cells.append(float(cost[hi - 1] - (cost[lo - 1] if lo > 0 else 0.0)) if hi > lo else 0.0)
	return [(cuts[r], cuts[r + 1]) for r in range(world)], cells, total


def allgather_v(t, group=None):
	"""all-gather of 1-D tensors of different lengths -> (concatenation in rank order, sizes).
	Sizes are exchanged first; payloads are padded to the longest shard for one all_gather."""
	world = dist.get_world_size(group)
	n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
	sizes = [torch.zeros_like(n) for _ in range(world)]
	dist.all_gather(sizes, n, group=group)
	sizes = [int(s.item()) for s in sizes]
	mx = max(sizes)
	pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
	pad[:t.numel()] = t
	parts = [torch.empty(mx, dtype=t.dtype, device=t.device) for _ in range(world)]
	dist.all_gather(parts, pad, group=group)
	return torch.cat([parts[r][:sizes[r]] for r in range(world)]), sizes


class _DevView:
	"""zero-copy torch view of library-owned device memory (__cuda_array_interface__)"""

	def __init__(self, ptr, nbytes):
		self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
		  "version": 2, "strides": None}


def device_view(ptr, nbytes, device):
	if nbytes == 0:
		return torch.empty(0, dtype=torch.uint8, device=device)
	return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def allgather_v_inplace(buf, bounds, group=None):
	"""In-place all-gather-v: `buf` is the full-size destination on every rank, rank r's part
	buf[bounds[r]:bounds[r+1]] is already in place; afterwards every rank holds every part.  One
	broadcast per source rank, all in flight together (NCCL: ncclBroadcast straight into the final
	position, no padding and no staging copy); empty parts are skipped on every rank alike."""
	world = dist.get_world_size(group)
	works = []
	for r in range(world):
		lo, hi = int(bounds[r]), int(bounds[r + 1])
		if hi > lo:
			works.append(dist.broadcast(buf[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r,
			  group=group, async_op=True))
	for w in works:
		w.wait()


def _sizes(n, device, group):
	t = torch.tensor([int(n)], dtype=torch.int64, device=device)
	out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
	dist.all_gather(out, t, group=group)
	return [int(x.item()) for x in out]


def gather_store(engine, group=None, have_store=True):
	"""Exchange step 1: every rank contributes the packed image of its pair range; afterwards every
	rank's engine holds the full all-pairs store.  The images are received directly into the
	library's final buffers (mb200_store_exchange_begin/_commit).  A rank whose range is empty
	(have_store False) contributes nothing.  Returns (bytes received per rank, device seconds)."""
	dev = torch.device("cuda", torch.cuda.current_device())
	rank = dist.get_rank(group)
	world = dist.get_world_size(group)
	if have_store:
		po, no, pe, ne = engine.store_pack_ptrs()
	else:
		po = pe = 0
		no = ne = 0
	n_off = _sizes(no, dev, group)
	n_ent = _sizes(ne, dev, group)
	ob = np.concatenate([[0], np.cumsum(n_off)]).astype(np.int64)
	eb = np.concatenate([[0], np.cumsum(n_ent)]).astype(np.int64)
	do, de = engine.store_exchange_begin(int(ob[-1]), int(eb[-1]))
	all_offs = device_view(do, int(ob[-1])*4, dev).view(torch.int32)
	all_ents = device_view(de, int(eb[-1])*8, dev).view(torch.int64)
	# own part into place (device-to-device, on torch's stream like the collectives)
	if no:
		all_offs[int(ob[rank]):int(ob[rank + 1])].copy_(device_view(po, no*4, dev).view(torch.int32))
	if ne:
		all_ents[int(eb[rank]):int(eb[rank + 1])].copy_(device_view(pe, ne*8, dev).view(torch.int64))
	ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	ev0.record()
	allgather_v_inplace(all_offs, ob, group)
	allgather_v_inplace(all_ents, eb, group)
	ev1.record()
	# the library reads these buffers on its own (non-blocking) stream: finish the collective first
	torch.cuda.current_stream().synchronize()
	engine.store_exchange_commit()
	return int(ob[-1])*4 + int(eb[-1])*8, ev0.elapsed_time(ev1)*1e-3


def gather_values(engine, entry_ranges, rank, group=None):
	"""Exchange step 2: after a sharded consistency iteration each rank owns new values for the
	entries of its pair range [entry_ranges[rank]); all-gather the entry ranges in place inside the
	library's packed store.  Returns (bytes, device seconds)."""
	dev = torch.device("cuda", torch.cuda.current_device())
	pe, ne = engine.store_entries_ptr()
	ents = device_view(pe, ne*8, dev).view(torch.int64)
	bounds = [entry_ranges[0][0]] + [hi for (_, hi) in entry_ranges]
	assert bounds[0] == 0 and bounds[-1] == ne, (bounds[0], bounds[-1], ne)
	ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	ev0.record()
	allgather_v_inplace(ents, bounds, group)
	ev1.record()
	torch.cuda.current_stream().synchronize()        # see gather_store
	engine.store_values_changed()
	return ne*8, ev0.elapsed_time(ev1)*1e-3


def gather_ea(ea_local, ranges, n, group=None, distributed=None):
	"""EA of every pair on every rank (N*N floats after symmetrisation), as the guide tree needs."""
	t = torch.as_tensor(np.ascontiguousarray(ea_local, np.float32))
	if distributed is None:
		distributed = dist.is_initialized() and dist.get_world_size(group) > 1
	if distributed:
		if dist.get_backend(group) == "nccl":
			t = t.cuda()
		t, _ = allgather_v(t, group)
		t = t.cpu()
	ea = t.numpy()
	m = np.zeros((n, n), np.float32)
	iu = np.triu_indices(n, 1)
	m[iu] = ea
	m.T[iu] = ea
	return m
