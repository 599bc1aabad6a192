"""Host-side model of the ordering argument behind the bucketed BuildPost (muscle_b200/csrc/align.cu,
k_bpc_*): the reference adds the terms of a column-posterior cell in (s, t) order (buildpostflat.cpp:18-105).
The device writes the terms of residue r = (s, pos) at a position taken from a prefix sum over the residues
ordered by (row, s) -- row = column of alignment A that holds the residue -- and then sorts STABLY on the
column bits only.  This test replays exactly that with numpy on random alignments and checks that every
cell's terms end up contiguous and in (s, t) order, i.e. that two radix passes are enough."""
import numpy as np
import pytest


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_prefix_by_row_then_stable_sort_on_columns_keeps_st_order(seed):
	rng = np.random.default_rng(seed)
	na, nb, cols_a, cols_b = 7, 5, 40, 37
	colbits = int(np.ceil(np.log2(cols_b)))
	# alignment A: every member occupies a random subset of the columns (ascending = positions in order)
	rowof, seqof = [], []
	for s in range(na):
		L = int(rng.integers(5, cols_a))
		cols = np.sort(rng.choice(cols_a, L, replace=False))
		rowof += list(cols)
		seqof += [s]*L
	rowof, seqof = np.array(rowof), np.array(seqof)
	nres = len(rowof)
	# terms of (residue r, member t): a short ascending list of columns of B with a value that encodes (s, t)
	cnt = rng.integers(0, 6, size=(nres, nb))
	off = np.concatenate([[0], np.cumsum(cnt.ravel())])                 # exclusive scan in (r, t) order
	M = int(off[-1])
	# residues ordered by (row, s): stable sort of the s-major residue list by row
	perm = np.argsort(rowof, kind="stable")
	nterms = np.array([off[(r + 1)*nb] - off[r*nb] for r in perm])
	start = np.concatenate([[0], np.cumsum(nterms)])[:-1]
	base = np.empty(nres, np.int64)
	base[perm] = start
	keys = np.full(M, -1, np.int64)
	tag = np.zeros((M, 2), np.int64)                                    # (s, t) of every term
	for r in range(nres):
		for t in range(nb):
			n = int(cnt[r, t])
			cols = np.sort(rng.choice(cols_b, n, replace=False))
			o = int(base[r] + (off[r*nb + t] - off[r*nb]))
			keys[o:o + n] = (rowof[r] << colbits) | cols
			tag[o:o + n] = (seqof[r], t)
	assert (keys >= 0).all()
	# the stream is in row order before the sort
	assert (np.diff(keys >> colbits) >= 0).all()
	order = np.argsort(keys & ((1 << colbits) - 1), kind="stable")       # the two radix passes
	k2, tag2 = keys[order], tag[order]
	# every cell is one contiguous run ...
	change = np.flatnonzero(np.diff(k2) != 0) + 1
	runs = np.split(np.arange(M), change)
	assert len({int(k2[r[0]]) for r in runs}) == len(runs)
	# ... whose terms are in (s, t) order, one term per (s, t)
	for r in runs:
		st = tag2[r, 0]*nb + tag2[r, 1]
		assert (np.diff(st) > 0).all()
