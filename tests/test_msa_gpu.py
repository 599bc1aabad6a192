"""GPU: device-resident MSAs (mb200_msa_reset / _join / _export) against the same progressive
alignment and refinement steps done the reference's way on the host: BuildPost + CalcAlnFlat through
mb200_align_groups (itself bit-exact against the oracle, test_consistency_gpu.py), gap insertion by
Sequence::AddGapsPath (sequence.cpp:115-140) and projection by MultiSequence::Project
(project.cpp:16-69) restated in Python below.  Paths, scores and final rows must be identical."""
import random
import numpy as np
import pytest
from muscle_b200 import synth
from muscle_b200.mpcflat import add_gaps_path

pytestmark = pytest.mark.gpu


def p2c(row):
	return np.array([c for c, ch in enumerate(row) if ch != "-"], np.uint32)


def project(rows):
	"""drop the columns that are gaps in every row (project.cpp:41-66)"""
	cols = len(rows[0])
	keep = [c for c in range(cols) if any(r[c] != "-" for r in rows)]
	return ["".join(r[c] for c in keep) for r in rows]


def host_join(engine, msa1, msa2):
	"""msa = list of (seq id, gapped row); returns (joined msa, path, score)"""
	ids1, rows1 = [i for i, _ in msa1], [r for _, r in msa1]
	ids2, rows2 = [i for i, _ in msa2], [r for _, r in msa2]
	score, path, _ = engine.align_groups(ids1, [p2c(r) for r in rows1], len(rows1[0]), ids2, [p2c(r) for r in rows2], len(rows2[0]))
	out = [(i, add_gaps_path(r, path, "X")) for i, r in msa1] + [(i, add_gaps_path(r, path, "Y")) for i, r in msa2]
	return out, path, score


def rows_from_device(engine, seqs, ids):
	maps, cols = engine.msa_export(ids)
	out = []
	for i, m, c in zip(ids, maps, cols):
		row = ["-"]*int(c)
		for pos, col in enumerate(m):
			row[int(col)] = seqs[i][pos]
		out.append((i, "".join(row)))
	return out


@pytest.mark.parametrize("bp_sort_min", ["1000000000000", "0"])      # both BuildPost paths of the library
@pytest.mark.parametrize("n,L,seed", [(9, 70, 7), (14, 110, 8)])
def test_progressive_and_refinement_on_device(engine, n, L, seed, monkeypatch, bp_sort_min):
	monkeypatch.setenv("MB200_BP_SORT_MIN", bp_sort_min)
	seqs = synth.make_family(n, L, L//4, seed=seed)
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	engine.consistency_iter()
	engine.msa_reset()
	rng = random.Random(seed)
	host = {i: [(i, seqs[i])] for i in range(n)}             # node -> msa
	nodes = list(range(n))
	nxt = n
	while len(nodes) > 1:                                     # a random join order (any binary tree will do)
		a, b = rng.sample(nodes, 2)
		joined, path, score = host_join(engine, host[a], host[b])
		ids_a, ids_b = [i for i, _ in host[a]], [i for i, _ in host[b]]
		cols, dscore, dpath = engine.msa_join(ids_a, ids_b, want_path=True)
		assert dpath == path and np.float32(dscore) == np.float32(score)
		assert cols == len(joined[0][1])
		nodes = [x for x in nodes if x not in (a, b)] + [nxt]
		host[nxt] = joined
		nxt += 1
	msa = host[nodes[0]]
	order = [i for i, _ in msa]
	assert rows_from_device(engine, seqs, order) == msa
	# refinement: random bipartitions of the rows, projected, re-aligned (refineflat.cpp:4-31)
	for it in range(6):
		pick = [rng.random() < 0.5 for _ in msa]
		g1 = [x for x, f in zip(msa, pick) if f]
		g2 = [x for x, f in zip(msa, pick) if not f]
		if not g1 or not g2:
			continue
		r1, r2 = project([r for _, r in g1]), project([r for _, r in g2])
		m1 = [(i, r) for (i, _), r in zip(g1, r1)]
		m2 = [(i, r) for (i, _), r in zip(g2, r2)]
		msa, path, score = host_join(engine, m1, m2)
		cols, dscore, dpath = engine.msa_join([i for i, _ in m1], [i for i, _ in m2], want_path=True)
		assert dpath == path and np.float32(dscore) == np.float32(score), it
		assert rows_from_device(engine, seqs, [i for i, _ in msa]) == msa, it


def test_msa_join_rejects_bad_groups(engine):
	from muscle_b200.engine import MB200Error
	seqs = synth.make_family(5, 50, 8, seed=3)
	engine.set_seqs(seqs)
	engine.posteriors_allpairs()
	engine.msa_reset()
	engine.msa_join([0], [1])
	with pytest.raises(MB200Error):
		engine.msa_join([0, 2], [3])          # 0 and 2 are not in one MSA
	with pytest.raises(MB200Error):
		engine.msa_join([0, 1], [1])          # overlap (the reference asserts SMI_1 != SMI_2)
	with pytest.raises(MB200Error):
		engine.align_groups([0], [np.arange(len(seqs[0]), dtype=np.uint32)], len(seqs[0]),
		  [0], [np.arange(len(seqs[0]), dtype=np.uint32)], len(seqs[0]))


@pytest.mark.parametrize("L", [1300, 2600])
def test_decoder_wide_and_tall(engine, oracle, L):
	"""k_aln_wave beyond one CTA (16 strips of 128 columns = 2048) and with the traceback bytes in global
	memory: a pair of long sequences through mb200_align_pairs against the oracle's CalcAlnFlat"""
	a = synth.make_family(2, L, L//8, seed=77)
	engine.set_seqs(a)
	engine.posteriors_allpairs()
	offs, ents = engine.export_all()
	scores, paths = engine.align_pairs([0])
	dense = np.zeros((len(a[0]), len(a[1])), np.float32)
	rows = np.repeat(np.arange(len(a[0])), np.diff(offs[0].astype(np.int64)))
	dense[rows, ents[0]["col"]] = ents[0]["p"]
	s, path = oracle.calcaln(dense)
	assert np.float32(s) == scores[0] and path == paths[0]
