"""GPU: Mega (Muscle-3D feature profile) emission mode of the posterior kernel (mb200_set_seqs_mega)
against the dense posteriors the COMPILED REFERENCE computed with Mega::CalcFwdFlat_mega /
CalcBwdFlat_mega / CalcPostFlat (tests/golden/mega_bb11001.npz), and the sparse store + EA against
the oracle fed with those posteriors.  Tolerance 0."""
import numpy as np
import pytest
import mega_util

pytestmark = pytest.mark.gpu


def test_mega_dense_posteriors_vs_reference(engine):
	model, profiles, posts = mega_util.load_model()
	engine.set_seqs_mega(model, profiles)
	for (i, j), want in posts.items():
		for force_c in (0, 1, 2):
			post, _, _, _ = engine.calc_post_dense(i, j, force_c=force_c)
			assert post.tobytes() == want.tobytes(), (i, j, force_c, float(np.abs(post - want).max()))


def test_mega_allpairs_sparse_and_ea(engine, oracle):
	model, profiles, posts = mega_util.load_model()
	engine.set_seqs_mega(model, profiles)
	ea = engine.posteriors_allpairs()
	offs, ents = engine.export_all()
	n = len(profiles)
	k = 0
	for i in range(n):
		for j in range(i + 1, n):
			want = posts[(i, j)]
			o, e = oracle.sparse(want)
			assert np.array_equal(offs[k], o) and ents[k].tobytes() == e.tobytes(), (i, j)
			want_ea = np.float32(oracle.alnscore(want))/np.float32(min(want.shape))
			assert np.float32(ea[k]) == np.float32(want_ea), (i, j)
			k += 1
	# and back to plain residues on the same context
	engine.set_seqs(["ACDEFGHIKLMNPQRSTVWY"*3, "ACDEFGHIKLMNPQRSTVWY"*3])
	assert engine.posteriors_allpairs()[0] > 0.9


def test_mega_rejects_bad_model(engine):
	from muscle_b200.engine import MB200Error
	model, profiles, _ = mega_util.load_model()
	bad = [p.copy() for p in profiles]
	bad[0][3, 1] = 200                         # letter outside the 16-letter alphabet of feature 1
	with pytest.raises(MB200Error):
		engine.set_seqs_mega(model, bad)
