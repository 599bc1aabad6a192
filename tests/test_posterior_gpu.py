"""GPU parity of the fused posterior kernel (through the C ABI) against the CPU oracle.

Tolerance: NONE (north_star allows 1e-4).  Forward-M, Backward-M and the total log-probability are
IEEE add/mul in the reference's order; the posterior's expf is a restatement of glibc's algorithm
that matches it on every float of the score range (muscle_b200/csrc/common.cuh), so dense
posteriors, the sparse store and the EA scores are compared BIT-EXACT as well.
"""
import os
import numpy as np
import pytest
from conftest import GOLDEN
from muscle_b200 import synth

pytestmark = pytest.mark.gpu

POST_TOL = 0.0        # bit-exact
EA_TOL = 0.0


def _dense_check(engine, oracle, X, Y, force_c=0):
	engine.set_seqs([X, Y])
	post, fwd, bwd, tot = engine.calc_post_dense(0, 1, force_c=force_c)
	f, b = oracle.fwd(X, Y), oracle.bwd(X, Y)
	assert fwd.tobytes() == np.ascontiguousarray(f[1:, 1:, 0]).tobytes(), "Forward M not bit-exact"
	assert bwd.tobytes() == np.ascontiguousarray(b[1:, 1:, 0]).tobytes(), "Backward M not bit-exact"
	assert np.float32(tot) == np.float32(oracle.total(f, b)), "total not bit-exact"
	p = oracle.post(X, Y)
	assert np.abs(post - p).max() <= POST_TOL
	assert ((post == 0) == (p == 0)).all(), "threshold at log(0.01) decided differently"
	return post, p


def _sparse_equal(off_g, ent_g, off_o, ent_o, tag=""):
	"""identical pattern and values within POST_TOL; membership may differ only at the 0.01 cut"""
	if (off_g == off_o).all() and (ent_g["col"] == ent_o["col"]).all():
		assert np.abs(ent_g["p"] - ent_o["p"]).max(initial=0) <= POST_TOL, tag
		return
	LX = len(off_o) - 1
	for i in range(LX):
		g = {int(c): float(p) for p, c in ent_g[off_g[i]:off_g[i + 1]]}
		o = {int(c): float(p) for p, c in ent_o[off_o[i]:off_o[i + 1]]}
		for c in set(g) | set(o):
			if c in g and c in o:
				assert abs(g[c] - o[c]) <= POST_TOL, tag
			else:
				v = g.get(c, o.get(c))
				assert abs(v - 0.01) <= POST_TOL, "%s row %d col %d only on one side with p=%g" % (tag, i, c, v)


def test_kat_dense(engine, oracle):
	z = np.load(os.path.join(GOLDEN, "kat_pairs.npz"))
	for k in range(int(z["n"])):
		X, Y = z["x%d" % k].tobytes(), z["y%d" % k].tobytes()
		post, _ = _dense_check(engine, oracle, X, Y)
		assert np.abs(post - z["post%d" % k]).max() <= POST_TOL     # golden from the compiled reference


@pytest.mark.parametrize("force_c", [0, 1, 2, 3, 5, 8, 16])
def test_family_dense_all_widths(engine, oracle, force_c):
	seqs = synth.make_family(4, 140, 30, seed=31)
	_dense_check(engine, oracle, seqs[0], seqs[1], force_c)
	_dense_check(engine, oracle, seqs[2], seqs[3], force_c)


def test_ragged_and_tiny(engine, oracle):
	s = synth.make_family(3, 200, 10, seed=8)
	for X, Y in [("A", "C"), ("A", s[0]), (s[0], "W"), (s[1][:33], s[2][:32]), (s[1][:32], s[2][:33]),
	  (s[0][:64], s[1][:65]), ("ACDXBZ*", "acdefg")]:
		_dense_check(engine, oracle, X, Y)


def test_long_multistrip(engine, oracle):
	# LY > 512 forces several 32*C column strips even at C=16
	s = synth.make_family(2, 700, 40, seed=9)
	_dense_check(engine, oracle, s[0], s[1])
	_dense_check(engine, oracle, s[0][:100], s[1], force_c=4)


def test_allpairs_c1_sparse_and_ea(engine, oracle):
	seqs = synth.make_config("C1")
	r = oracle.all_pairs(seqs, threads=0)
	engine.set_seqs(seqs)
	ea = engine.posteriors_allpairs()
	n = len(seqs)
	iu = np.triu_indices(n, 1)
	assert np.abs(ea - r["ea"][iu]).max() <= EA_TOL
	nnz, tot = engine.store_nnz()
	for k in range(len(nnz)):
		off, ent = engine.export_pair(k, int(nnz[k]))
		_sparse_equal(off, ent, r["row_off"][k], r["entries"][k], "pair %d" % k)
	assert int(tot) == int(r["nnz"].sum())


def test_family8_golden(engine):
	z = np.load(os.path.join(GOLDEN, "family8.npz"))
	seqs = [z["seq%d" % i].tobytes() for i in range(int(z["n"]))]
	engine.set_seqs(seqs)
	ea = engine.posteriors_allpairs()
	n = len(seqs)
	iu = np.triu_indices(n, 1)
	assert np.abs(ea - z["ea"][iu]).max() <= EA_TOL
	nnz, _ = engine.store_nnz()
	for k in range(len(nnz)):
		off, ent = engine.export_pair(k, int(nnz[k]))
		_sparse_equal(off, ent, z["off%d" % k], z["ent0_%d" % k], "pair %d" % k)


def test_pair_list_api_and_order(engine, oracle):
	seqs = synth.make_family(5, 80, 20, seed=77)
	engine.set_seqs(seqs)
	px, py = [3, 0, 2, 4], [1, 4, 3, 0]          # arbitrary orientation, X indexes rows
	ea = engine.posteriors(px, py)
	nnz, _ = engine.store_nnz()
	for k, (x, y) in enumerate(zip(px, py)):
		p = oracle.post(seqs[x], seqs[y])
		off_o, ent_o = oracle.sparse(p)
		off, ent = engine.export_pair(k, int(nnz[k]))
		_sparse_equal(off, ent, off_o, ent_o)
		assert abs(ea[k] - oracle.alnscore(p)/min(len(seqs[x]), len(seqs[y]))) <= EA_TOL


def test_errors(engine):
	from muscle_b200.engine import MB200Error
	engine.set_seqs(["ACD", "EFG"])
	with pytest.raises(MB200Error):
		engine.posteriors([0], [5])


def test_config4_scale_long_pair(engine, oracle):
	"""BASELINE config 4 scale: one 1900 x 2600 pair (6 strips of 512 columns, 4.9e6 cells).
	Forward/Backward/total bit-exact, posterior within tolerance, EA and sparse store consistent."""
	s = synth.make_family(2, 2250, 350, seed=44)
	X, Y = s[0][:1900], (s[1] + s[1])[:2600]
	post, p = _dense_check(engine, oracle, X, Y)
	engine.set_seqs([X, Y])
	ea = engine.posteriors([0], [1])
	off, ent = engine.export_pair(0)
	off_o, ent_o = oracle.sparse(p)
	_sparse_equal(off, ent, off_o, ent_o, "long pair")
	assert abs(ea[0] - oracle.alnscore(p)/min(len(X), len(Y))) <= EA_TOL


def test_nucleotide_alphabet(tables):
	"""a second set of HMM tables (nucleotides, 6 residue classes instead of 21) through the same kernels"""
	from muscle_b200.engine import Engine
	z = np.load(os.path.join(GOLDEN, "hmm_nucleo.npz"))
	t = {k: z[k] for k in ("start", "trans", "ins", "match", "min_sparse_score")}
	k = np.load(os.path.join(GOLDEN, "kat_nucleo.npz"))
	e = Engine(0)
	e.set_hmm(t)
	for i in range(int(k["n"])):
		X, Y = k["x%d" % i].tobytes(), k["y%d" % i].tobytes()
		e.set_seqs([X, Y])
		post, fwd, bwd, tot = e.calc_post_dense(0, 1)
		assert fwd.tobytes() == k["fwdm%d" % i].tobytes() and bwd.tobytes() == k["bwdm%d" % i].tobytes()
		assert np.float32(tot) == k["total%d" % i]
		assert np.abs(post - k["post%d" % i]).max() <= POST_TOL
		ea = e.posteriors([0], [1])
		assert abs(ea[0] - float(k["alnscore%d" % i])/min(len(X), len(Y))) <= EA_TOL
	# tables can be swapped on a live context (the reference rewrites them between replicates)
	e.set_hmm(tables)
	e.close()


def test_guards_and_bad_input(engine):
	"""reference guard LX*LY*5+100 > INT_MAX (fwdflat3.cpp:17) and argument checking: errors, no compute"""
	from muscle_b200.engine import MB200Error
	long_a, long_b = "A"*21000, "C"*21000
	engine.set_seqs([long_a, long_b, "ACD"])
	with pytest.raises(MB200Error) as ei:
		engine.posteriors([0], [1])
	assert ei.value.code == -5 and "HMM overflow" in str(ei.value)
	engine.posteriors([0], [2])                           # 21000 x 3 is fine
	with pytest.raises(MB200Error):
		engine.set_seqs(["ACD", ""])                       # empty sequence: the reference would read X[0]
	engine.set_seqs(["ACD", "EFG"])
	with pytest.raises(MB200Error):
		engine.consistency_iter()                          # store is not an all-pairs store yet
