"""CPU: the C restatement against the compiled reference on fresh seeded inputs (bit-exact).
Skipped where oracle/_ref has not been built."""
import numpy as np
from muscle_b200 import synth


def test_tables_match_golden(ref, tables):
	t = ref.tables()
	for k in ("start", "trans", "ins", "match"):
		assert t[k].tobytes() == np.asarray(tables[k], np.float32).tobytes()
	assert np.float32(t["min_sparse_score"]) == np.float32(tables["min_sparse_score"])


def test_pairs_bitexact(ref, oracle):
	seqs = synth.make_family(6, 90, 25, seed=5) + synth.random_unrelated(2, 70, seed=6) + ["W", "ACDEFGHIKLMNPQRSTVWYXBZ"]
	for a in range(0, len(seqs) - 1):
		X, Y = seqs[a], seqs[a + 1]
		assert oracle.fwd(X, Y).tobytes() == ref.fwd(X, Y).tobytes()
		assert oracle.bwd(X, Y).tobytes() == ref.bwd(X, Y).tobytes()
		p = oracle.post(X, Y)
		assert p.tobytes() == ref.post(X, Y).tobytes()
		o1, e1 = oracle.sparse(p)
		o2, e2 = ref.sparse(p)
		assert (o1 == o2).all() and e1.tobytes() == e2.tobytes()
		assert oracle.alnscore(p) == ref.alnscore(p)
		assert oracle.calcaln(p) == ref.calcaln(p)


def test_mpc_pipeline_bitexact(ref, oracle):
	seqs = synth.make_family(7, 70, 10, seed=21)
	n = len(seqs)
	M = ref.mpc(seqs)
	M.posteriors()
	r = oracle.all_pairs(seqs, threads=2)
	ea = M.distmx()
	np.fill_diagonal(ea, 0)
	assert r["ea"].tobytes() == ea.tobytes()
	offs, ents = M.export_all()
	for p in range(len(offs)):
		assert (offs[p] == r["row_off"][p]).all() and ents[p].tobytes() == r["entries"][p].tobytes()
	lens = [len(s) for s in seqs]
	pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
	M.consiter()
	_, ents1 = M.export_all()
	for p, (x, y) in enumerate(pairs):
		assert oracle.conspair(lens, x, y, offs, ents).tobytes() == ents1[p].tobytes()
	# BuildPost + CalcAlnFlat of a join of two gapped groups
	c1 = max(len(seqs[0]), len(seqs[1]) + 2)
	c2 = max(len(seqs[2]), len(seqs[3]) + 1, len(seqs[4]))
	rows1 = [seqs[0] + "-"*(c1 - len(seqs[0])), "--" + seqs[1] + "-"*(c1 - 2 - len(seqs[1]))]
	rows2 = [seqs[2] + "-"*(c2 - len(seqs[2])), "-" + seqs[3] + "-"*(c2 - 1 - len(seqs[3])), seqs[4] + "-"*(c2 - len(seqs[4]))]
	score, path, post = M.alignalns([0, 1], rows1, [2, 3, 4], rows2)
	def p2c(row):
		return np.array([c for c, ch in enumerate(row) if ch != "-"], np.uint32)
	mine = oracle.buildpost(lens, [0, 1], [p2c(r) for r in rows1], c1, [2, 3, 4], [p2c(r) for r in rows2], c2, offs, ents1)
	assert mine.tobytes() == post.tobytes()
	s2, path2 = oracle.calcaln(mine)
	assert np.float32(s2) == np.float32(score) and path2 == path
	M.close()


def test_committed_goldens_are_what_the_reference_produces(ref):
	"""tests/golden/kat_pairs.npz and family8.npz (read on the GPU box, where /root/reference does not
	exist) are regenerated from the compiled reference and must match the committed bytes."""
	import os
	from conftest import GOLDEN
	z = np.load(os.path.join(GOLDEN, "kat_pairs.npz"))
	for k in range(int(z["n"])):
		X, Y = z["x%d" % k].tobytes(), z["y%d" % k].tobytes()
		f, b = ref.fwd(X, Y), ref.bwd(X, Y)
		assert f.tobytes() == z["fwd%d" % k].tobytes() and b.tobytes() == z["bwd%d" % k].tobytes()
		p = ref.post(X, Y)
		assert p.tobytes() == z["post%d" % k].tobytes()
		sc, path = ref.calcaln(p)
		assert np.float32(sc) == z["calcaln%d" % k] and path.encode() == z["path%d" % k].tobytes()
	fz = np.load(os.path.join(GOLDEN, "family8.npz"))
	seqs = [fz["seq%d" % i].tobytes() for i in range(int(fz["n"]))]
	M = ref.mpc(seqs)
	M.posteriors()
	assert M.distmx().tobytes() == fz["ea"].tobytes()
	offs, ents = M.export_all()
	for p in range(len(offs)):
		assert (offs[p] == fz["off%d" % p]).all() and ents[p].tobytes() == fz["ent0_%d" % p].tobytes()
	M.consiter()
	_, ents1 = M.export_all()
	for p in range(len(offs)):
		assert ents1[p].tobytes() == fz["ent1_%d" % p].tobytes()
	M.close()


def test_random_small_pairs_bitexact(ref, oracle):
	"""200 random pairs, lengths 1..40, alphabet with lower case, wildcards and non-letters:
	every function of the restatement equals the compiled reference bit for bit."""
	rng = np.random.default_rng(12345)
	alpha = list("ACDEFGHIKLMNPQRSTVWY"*3 + "acdxBZUO*-.")
	for _ in range(200):
		X = "".join(rng.choice(alpha, size=int(rng.integers(1, 41))))
		Y = "".join(rng.choice(alpha, size=int(rng.integers(1, 41))))
		assert oracle.fwd(X, Y).tobytes() == ref.fwd(X, Y).tobytes(), (X, Y)
		assert oracle.bwd(X, Y).tobytes() == ref.bwd(X, Y).tobytes(), (X, Y)
		p = oracle.post(X, Y)
		assert p.tobytes() == ref.post(X, Y).tobytes(), (X, Y)
		o1, e1 = oracle.sparse(p)
		o2, e2 = ref.sparse(p)
		assert (o1 == o2).all() and e1.tobytes() == e2.tobytes()
		assert oracle.alnscore(p) == ref.alnscore(p)
		assert oracle.calcaln(p) == ref.calcaln(p)


def test_upgma_oracle_vs_reference(oracle, ref):
	"""guide tree: the C restatement of UPGMA5::FixEADistMx + Run against the compiled reference, every
	linkage, random and tie-heavy EA matrices (ties exercise the scan-order and nearest-neighbour quirks)"""
	rng = np.random.default_rng(11)
	for trial in range(120):
		n = int(rng.integers(2, 48))
		npair = n*(n - 1)//2
		ea = (rng.integers(0, 5, npair).astype(np.float32)/4) if trial % 3 == 0 else rng.random(npair).astype(np.float32)
		for link in (1, 2, 3, 4):
			a, b = oracle.upgma(n, ea, link), ref.upgma(n, ea, link)
			for x, y in zip(a, b):
				assert x.tobytes() == y.tobytes(), (trial, n, link)
