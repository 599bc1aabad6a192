"""CPU: the Mega emission restatement (tests/mega_util.py) pushed through the pinned plain oracle
reproduces, bit for bit, the dense posteriors the COMPILED REFERENCE computed with
Mega::CalcFwdFlat_mega / CalcBwdFlat_mega / CalcPostFlat for every pair of BB11001.mega
(tests/golden/mega_bb11001.npz, made by tests/golden/make_golden_mega.py)."""
import numpy as np
from conftest import load_tables
import mega_util


def test_mega_restatement_matches_reference_goldens():
	from oracle.pyoracle import Oracle
	base = load_tables()
	model, profiles, posts = mega_util.load_model()
	assert list(model["alpha"]) == [20, 16, 16, 16, 16, 16, 16, 16]
	for (i, j), want in posts.items():
		t, X, Y = mega_util.synthetic_tables(base, model, profiles[i], profiles[j])
		got = Oracle(t).post(X, Y)
		assert got.tobytes() == want.tobytes(), (i, j, float(np.abs(got - want).max()))
