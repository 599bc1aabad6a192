"""GPU: the device guide tree (mb200_guide_tree = UPGMA5::FixEADistMx + UPGMA5::Run) against the CPU
oracle (itself pinned bit-for-bit against the compiled reference, tests/test_oracle_vs_ref.py):
children and branch lengths of every internal node, tolerance 0."""
import numpy as np
import pytest
from muscle_b200 import synth

pytestmark = pytest.mark.gpu


def _same(a, b):
	return all(x.tobytes() == y.tobytes() for x, y in zip(a, b))


def test_guide_tree_from_store_ea(engine, oracle):
	seqs = synth.make_family(40, 90, 20, seed=12)
	engine.set_seqs(seqs)
	ea = engine.posteriors_allpairs()
	got = engine.guide_tree()                       # EA vector left on the device by the posterior stage
	assert _same(got, oracle.upgma(len(seqs), ea, 4))
	assert _same(engine.guide_tree(ea), got)        # host-supplied EA (what the multi-GPU front end passes)


@pytest.mark.parametrize("n", [2, 3, 33, 257, 1200])
def test_guide_tree_ties_and_sizes(engine, oracle, n):
	rng = np.random.default_rng(n)
	engine.set_seqs(["ACDEFGHIKL"]*n)
	npair = n*(n - 1)//2
	for ea in (rng.random(npair).astype(np.float32), rng.integers(0, 4, npair).astype(np.float32)/4):
		for link in (1, 2, 3, 4):
			assert _same(engine.guide_tree(ea, link), oracle.upgma(n, ea, link)), (n, link)


def test_guide_tree_rejects_bad_ea(engine):
	from muscle_b200.engine import MB200Error
	engine.set_seqs(["ACDEFGHIKL"]*5)
	ea = np.full(10, 0.5, np.float32)
	ea[3] = 1.5
	with pytest.raises(MB200Error):
		engine.guide_tree(ea)
