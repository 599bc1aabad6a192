"""CPU: the C restatement (oracle/) against golden vectors produced by the compiled reference."""
import numpy as np
from conftest import GOLDEN
import os


def test_hmm_tables_shape(tables):
	assert tables["match"].size == 65536 and tables["ins"].size == 256
	assert abs(float(tables["min_sparse_score"]) - np.log(0.01)) < 1e-6


def test_logexp1_pieces(oracle):
	# piece boundaries of LOGEXP1 (scoretype.h:95-105) and the 7.5 cut of LOG_ADD
	for x, y in [(0.0, 0.0), (-1.0, 0.0), (-2.5, 0.0), (-4.5, 0.0), (-7.4999, 0.0), (-7.5, 0.0), (-2e20, -3.0),
	  (-2e20, -2e20)]:
		r = oracle.log_add(x, y)
		assert r == oracle.log_add(y, x)
		if x <= -7.5 + y or x == -2e20:
			assert r == np.float32(y)
	assert abs(oracle.log_add(0.0, 0.0) - np.log(2.0)) < 3e-4


def test_kat_pairs_bitexact(oracle):
	z = np.load(os.path.join(GOLDEN, "kat_pairs.npz"))
	for k in range(int(z["n"])):
		X, Y = z["x%d" % k].tobytes(), z["y%d" % k].tobytes()
		f, b = oracle.fwd(X, Y), oracle.bwd(X, Y)
		assert f.tobytes() == z["fwd%d" % k].tobytes()
		assert b.tobytes() == z["bwd%d" % k].tobytes()
		assert np.float32(oracle.total(f, b)) == z["total%d" % k]
		p = oracle.post(X, Y)
		assert p.tobytes() == z["post%d" % k].tobytes()
		off, ent = oracle.sparse(p)
		assert (off == z["off%d" % k]).all() and ent.tobytes() == z["ent%d" % k].tobytes()
		assert np.float32(oracle.alnscore(p)) == z["alnscore%d" % k]
		sc, path = oracle.calcaln(p)
		assert np.float32(sc) == z["calcaln%d" % k] and path.encode() == z["path%d" % k].tobytes()


def _family():
	z = np.load(os.path.join(GOLDEN, "family8.npz"))
	seqs = [z["seq%d" % i].tobytes() for i in range(int(z["n"]))]
	return z, seqs


def test_family_allpairs_bitexact(oracle):
	z, seqs = _family()
	r = oracle.all_pairs(seqs, threads=2)
	assert r["ea"].tobytes() == np.where(np.eye(len(seqs), dtype=bool), 0, z["ea"]).astype(np.float32).tobytes()
	for p in range(len(r["row_off"])):
		assert (r["row_off"][p] == z["off%d" % p]).all()
		assert r["entries"][p].tobytes() == z["ent0_%d" % p].tobytes()


def test_family_consistency_bitexact(oracle):
	z, seqs = _family()
	n = len(seqs)
	lens = [len(s) for s in seqs]
	pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
	offs = [z["off%d" % p] for p in range(len(pairs))]
	ents = [z["ent0_%d" % p] for p in range(len(pairs))]
	for it in (1, 2):
		new = [oracle.conspair(lens, x, y, offs, ents) for (x, y) in pairs]
		for p in range(len(pairs)):
			assert new[p].tobytes() == z["ent%d_%d" % (it, p)].tobytes(), (it, p)
		ents = new


def test_family_decode(oracle):
	z, seqs = _family()
	sc, path = oracle.calcaln(z["post01"])
	assert np.float32(sc) == z["score01"] and path.encode() == z["path01"].tobytes()


def test_nucleotide_tables_and_kat():
	"""nucleotide alphabet (4x4 table, U==T, wildcard for N/R/Y): oracle vs compiled-reference goldens"""
	from oracle.pyoracle import Oracle
	z = np.load(os.path.join(GOLDEN, "hmm_nucleo.npz"))
	t = {k: z[k] for k in ("start", "trans", "ins", "match", "min_sparse_score")}
	assert t["match"].reshape(256, 256)[ord("U"), ord("A")] == t["match"].reshape(256, 256)[ord("T"), ord("A")]
	O = Oracle(t)
	k = np.load(os.path.join(GOLDEN, "kat_nucleo.npz"))
	for i in range(int(k["n"])):
		X, Y = k["x%d" % i].tobytes(), k["y%d" % i].tobytes()
		f, b = O.fwd(X, Y), O.bwd(X, Y)
		assert np.ascontiguousarray(f[1:, 1:, 0]).tobytes() == k["fwdm%d" % i].tobytes()
		assert np.ascontiguousarray(b[1:, 1:, 0]).tobytes() == k["bwdm%d" % i].tobytes()
		assert np.float32(O.total(f, b)) == k["total%d" % i]
		assert O.post(X, Y).tobytes() == k["post%d" % i].tobytes()
