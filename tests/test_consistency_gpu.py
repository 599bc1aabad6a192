"""GPU parity of the consistency (relax) iteration, the packed store, the multi-device exchange
and the posterior decoding kernels against the CPU oracle and the compiled reference's goldens.
The relax arithmetic is fp32 mul/add in the reference's order and the posterior stage is itself
bit-exact (glibc-exact expf on the device), so everything here is compared with tolerance 0."""
import os
import numpy as np
import pytest
from conftest import GOLDEN
from muscle_b200 import synth

pytestmark = pytest.mark.gpu


def _pairs(n):
	return [(i, j) for i in range(n) for j in range(i + 1, n)]


def test_export_all_matches_export_pair(engine):
	seqs = synth.make_family(9, 70, 15, seed=41)
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	nnz, tot = engine.store_nnz()
	singles = [engine.export_pair(k, int(nnz[k])) for k in range(len(nnz))]
	offs, ents = engine.export_all()
	assert sum(len(e) for e in ents) == tot
	for k in range(len(nnz)):
		assert (offs[k] == singles[k][0]).all() and ents[k].tobytes() == singles[k][1].tobytes()
		# columns ascending inside every row (MySparseMx invariant)
		for i in range(len(offs[k]) - 1):
			c = ents[k]["col"][offs[k][i]:offs[k][i + 1]]
			assert (np.diff(c.astype(np.int64)) > 0).all()


@pytest.mark.parametrize("n,L,seed", [(3, 50, 1), (8, 70, 2), (20, 120, 3)])
def test_consistency_bitexact_vs_oracle(engine, oracle, n, L, seed):
	seqs = synth.make_family(n, L, L//5, seed=seed)
	lens = [len(s) for s in seqs]
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	offs, ents = engine.export_all()
	for it in range(2):
		want = [oracle.conspair(lens, x, y, offs, ents) for (x, y) in _pairs(n)]
		engine.consistency_iter()
		offs2, got = engine.export_all()
		for p in range(len(want)):
			assert (offs2[p] == offs[p]).all()
			assert got[p].tobytes() == want[p].tobytes(), "iter %d pair %d" % (it, p)
		ents = got


def test_consistency_golden_family8(engine):
	"""against the compiled reference's own output (tests/golden/family8.npz): bit-exact"""
	z = np.load(os.path.join(GOLDEN, "family8.npz"))
	seqs = [z["seq%d" % i].tobytes() for i in range(int(z["n"]))]
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	for it in (1, 2):
		engine.consistency_iter()
		offs, ents = engine.export_all()
		for p in range(len(offs)):
			assert z["ent%d_%d" % (it, p)].tobytes() == ents[p].tobytes(), (it, p)


def test_consistency_partial_range_and_values_exchange(engine, oracle):
	"""sharded update: two half-range calls on copies == one full call (what two ranks would do)"""
	import torch
	seqs = synth.make_family(7, 60, 10, seed=5)
	n = len(seqs)
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	offs, ents0 = engine.export_all()
	lens = [len(s) for s in seqs]
	want = [oracle.conspair(lens, x, y, offs, ents0) for (x, y) in _pairs(n)]
	npairs = n*(n - 1)//2
	half = npairs//2
	# rank A updates [0,half); grab its new values, then restore and let "rank B" update [half,npairs)
	nnz, tot = engine.store_nnz()
	base = np.concatenate([[0], np.cumsum(nnz)]).astype(np.int64)
	engine.consistency_iter(0, half)
	va = engine.store_values_torch()
	engine.posteriors_allpairs()
	engine.consistency_iter(half, npairs)
	# exchange: bring rank A's values for its range
	engine.store_set_values_torch(va[:base[half]], 0)
	_, got = engine.export_all()
	for p in range(npairs):
		assert got[p].tobytes() == want[p].tobytes(), p


def test_align_pairs_vs_oracle(engine, oracle):
	seqs = synth.make_family(6, 90, 25, seed=13) + ["ACDEFGHIKLMNPQRSTVWY"*20]     # LY=400 > 256 threads
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	offs, ents = engine.export_all()
	n = len(seqs)
	pairs = _pairs(n)
	sel = list(range(len(pairs)))
	scores, paths = engine.align_pairs(sel)
	for k in sel:
		x, y = pairs[k]
		dense = np.zeros((len(seqs[x]), len(seqs[y])), np.float32)
		for i in range(len(seqs[x])):
			for e in range(offs[k][i], offs[k][i + 1]):
				dense[i, ents[k]["col"][e]] = ents[k]["p"][e]
		s, path = oracle.calcaln(dense)
		assert np.float32(s) == scores[k] and path == paths[k], k


@pytest.mark.parametrize("bp_sort_min", ["1000000000000", "0"])      # ordered row walk / contributions sorted by cell
def test_align_groups_vs_oracle(engine, oracle, monkeypatch, bp_sort_min):
	monkeypatch.setenv("MB200_BP_SORT_MIN", bp_sort_min)
	seqs = synth.make_family(7, 70, 10, seed=21)
	lens = [len(s) for s in seqs]
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	engine.consistency_iter()
	offs, ents = engine.export_all()
	c1 = max(len(seqs[0]), len(seqs[1]) + 2, len(seqs[5]) + 1)
	c2 = max(len(seqs[2]), len(seqs[3]) + 1, len(seqs[4]))
	rows1 = [seqs[0] + "-"*(c1 - len(seqs[0])), "--" + seqs[1] + "-"*(c1 - 2 - len(seqs[1])), "-" + seqs[5] + "-"*(c1 - 1 - len(seqs[5]))]
	rows2 = [seqs[2] + "-"*(c2 - len(seqs[2])), "-" + seqs[3] + "-"*(c2 - 1 - len(seqs[3])), seqs[4] + "-"*(c2 - len(seqs[4]))]

	def p2c(row):
		return np.array([c for c, ch in enumerate(row) if ch != "-"], np.uint32)
	ids1, ids2 = [0, 1, 5], [2, 3, 4]
	want = oracle.buildpost(lens, ids1, [p2c(r) for r in rows1], c1, ids2, [p2c(r) for r in rows2], c2, offs, ents)
	score, path, post = engine.align_groups(ids1, [p2c(r) for r in rows1], c1, ids2, [p2c(r) for r in rows2], c2, want_post=True)
	assert post.tobytes() == want.tobytes()
	s, pth = oracle.calcaln(want)
	assert np.float32(s) == np.float32(score) and pth == path
	# the transposed orientation (group ids reversed: stored pair is (t,s))
	want2 = oracle.buildpost(lens, ids2, [p2c(r) for r in rows2], c2, ids1, [p2c(r) for r in rows1], c1, offs, ents)
	score2, path2, post2 = engine.align_groups(ids2, [p2c(r) for r in rows2], c2, ids1, [p2c(r) for r in rows1], c1, want_post=True)
	assert post2.tobytes() == want2.tobytes()
	s2, pth2 = oracle.calcaln(want2)
	assert np.float32(s2) == np.float32(score2) and pth2 == path2


def test_align_groups_large_join_paths_agree(engine, monkeypatch):
	"""A join above the library's own threshold (>= 1e6 (residue, member) rows) takes the bucketed-by-cell
	BuildPost on its own; the column-posterior matrix, the path and the score must equal the ordered row walk
	bit for bit (two staggered ungapped groups of 64 sequences)."""
	seqs = synth.make_family(128, 260, 40, seed=77)
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	rng = np.random.default_rng(5)
	ids1, ids2 = list(range(0, 128, 2)), list(range(1, 128, 2))
	sh1 = {i: int(rng.integers(0, 12)) for i in ids1}
	sh2 = {i: int(rng.integers(0, 12)) for i in ids2}
	c1 = max(len(seqs[i]) + sh1[i] for i in ids1)
	c2 = max(len(seqs[i]) + sh2[i] for i in ids2)
	p1 = [np.arange(len(seqs[i]), dtype=np.uint32) + sh1[i] for i in ids1]
	p2 = [np.arange(len(seqs[i]), dtype=np.uint32) + sh2[i] for i in ids2]
	assert sum(len(seqs[i]) for i in ids1)*len(ids2) >= 1000000
	monkeypatch.setenv("MB200_BP_SORT_MIN", "1000000000000")
	s_walk, path_walk, post_walk = engine.align_groups(ids1, p1, c1, ids2, p2, c2, want_post=True)
	monkeypatch.delenv("MB200_BP_SORT_MIN")
	s_sort, path_sort, post_sort = engine.align_groups(ids1, p1, c1, ids2, p2, c2, want_post=True)
	assert post_sort.tobytes() == post_walk.tobytes()
	assert path_sort == path_walk and np.float32(s_sort) == np.float32(s_walk)
	assert float(post_sort.max()) > 1.0          # conserved columns: long ordered sums


@pytest.mark.parametrize("world", [3, 8])
def test_virtual_ranks_pipeline(tables, oracle, world):
	"""the sharded pipeline with `world` virtual ranks on one GPU (no NCCL): every rank computes its
	pair range, the packed images are concatenated as the all-gather would, every rank loads the full
	store, runs its range of the consistency iteration, values are exchanged -- and all ranks must end
	bit-identical to a single-engine run (exercises middle ranges of every library call)."""
	import torch
	from muscle_b200.engine import Engine
	from muscle_b200 import dist as mdist
	seqs = synth.make_family(26, 90, 25, seed=9)
	n = len(seqs)
	lens = [len(s) for s in seqs]
	npairs = n*(n - 1)//2
	ranges, _, _ = mdist.shard_ranges(lens, world)
	single = Engine(0); single.set_hmm(tables); single.set_seqs(seqs)
	ea_s = single.posteriors_allpairs()
	single.consistency_iter(); single.consistency_iter()
	offs_s, ents_s = single.export_all()
	engines = []
	for r in range(world):
		e = Engine(0); e.set_hmm(tables); e.set_seqs(seqs)
		engines.append(e)
	dev = torch.device("cuda", 0)
	img_off, img_ent, eas = [], [], []
	for r, (lo, hi) in enumerate(ranges):
		eas.append(engines[r].posteriors_allpairs(lo, hi))
		po, no, pe, ne = engines[r].store_pack_ptrs()
		img_off.append(mdist.device_view(po, no*4, dev).view(torch.int32).clone())
		img_ent.append(mdist.device_view(pe, ne*8, dev).view(torch.int64).clone())
	assert np.array_equal(np.concatenate(eas), ea_s)
	all_off, all_ent = torch.cat(img_off), torch.cat(img_ent)
	# even ranks: copying load (mb200_store_load_allpairs); odd ranks: in-place exchange buffers
	# (mb200_store_exchange_begin/_commit), filled here by a device copy as the collective would
	for r, e in enumerate(engines):
		if r % 2 == 0:
			e.store_load_allpairs(0, npairs, all_off.data_ptr(), all_off.numel(), all_ent.data_ptr(), all_ent.numel())
		else:
			do, de = e.store_exchange_begin(all_off.numel(), all_ent.numel())
			mdist.device_view(do, all_off.numel()*4, dev).view(torch.int32).copy_(all_off)
			mdist.device_view(de, all_ent.numel()*8, dev).view(torch.int64).copy_(all_ent)
			torch.cuda.synchronize()
			e.store_exchange_commit()
	nnz, _ = engines[0].store_nnz()
	base = np.concatenate([[0], np.cumsum(nnz.astype(np.int64))])
	for it in range(2):
		parts = []
		for r, (lo, hi) in enumerate(ranges):
			engines[r].consistency_iter(lo, hi)
			pe, ne = engines[r].store_entries_ptr()
			v = mdist.device_view(pe, ne*8, dev).view(torch.int64)
			parts.append(v[int(base[lo]):int(base[hi])].clone())
		allv = torch.cat(parts)
		for e in engines:
			pe, ne = e.store_entries_ptr()
			mdist.device_view(pe, ne*8, dev).view(torch.int64).copy_(allv)
			torch.cuda.synchronize()
			e.store_values_changed()
	for r, e in enumerate(engines):
		offs, ents = e.export_all()
		for p in range(npairs):
			assert np.array_equal(offs[p], offs_s[p]) and ents[p].tobytes() == ents_s[p].tobytes(), (r, p)
		e.close()
	single.close()


def test_group_one_device_equals_engine(tables):
	"""the single-process multi-device front end with one member must equal a plain engine"""
	from muscle_b200.engine import Engine, Group
	seqs = synth.make_family(12, 80, 20, seed=33)
	e = Engine(0); e.set_hmm(tables); e.set_seqs(seqs)
	ea = e.posteriors_allpairs()
	e.consistency_iter(); e.consistency_iter()
	offs, ents = e.export_all()
	g = Group([0]); g.set_hmm(tables); g.set_seqs(seqs)
	ea_g = g.posteriors_allpairs()
	g.consistency_iter(); g.consistency_iter()
	offs_g, ents_g = g.engine(0).export_all()
	assert np.array_equal(ea, ea_g)
	for p in range(len(offs)):
		assert np.array_equal(offs[p], offs_g[p]) and ents[p].tobytes() == ents_g[p].tobytes(), p
	assert g.stats()["ndev"] == 1 and g.stats()["relax_kernel_ms"] > 0
	g.close(); e.close()


def test_group_all_devices_equal_engine(tables):
	""">= 2 GPUs: sharded posteriors + peer-memory all-gather-v + sharded relax + entry exchange
	must leave EVERY device bit-identical to a single-device run"""
	import torch
	if torch.cuda.device_count() < 2:
		pytest.skip("needs 2 GPUs")
	from muscle_b200.engine import Engine, Group
	seqs = synth.make_family(30, 100, 25, seed=35)
	e = Engine(0); e.set_hmm(tables); e.set_seqs(seqs)
	ea = e.posteriors_allpairs()
	e.consistency_iter(); e.consistency_iter()
	offs, ents = e.export_all()
	g = Group(None); g.set_hmm(tables); g.set_seqs(seqs)
	ea_g = g.posteriors_allpairs()
	g.consistency_iter(); g.consistency_iter()
	assert np.array_equal(ea, ea_g)
	st = g.stats()
	assert st["ndev"] == torch.cuda.device_count() and st["exchange1_bytes_per_dev"] > 0 and st["exchange2_bytes_per_dev"] > 0
	for r in range(g.size):
		offs_g, ents_g = g.engine(r).export_all()
		for p in range(len(offs)):
			assert np.array_equal(offs[p], offs_g[p]) and ents[p].tobytes() == ents_g[p].tobytes(), (r, p)
	g.close(); e.close()
