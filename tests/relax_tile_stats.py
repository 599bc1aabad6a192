"""Analysis helper (not a test): what a 16x16-tile tensor-core formulation of the consistency transform
would have to compute.  Uses the CPU oracle on a subset of a synthetic configuration, counts for every
(x, y, z) the tile-level products C[I,J] += A[I,K] * B[K,J] that a blocked kernel must issue to cover all
scalar products that land on stored entries of XY, and compares with the scalar products themselves.
    python tests/relax_tile_stats.py [C2] [nseq] [tile]"""
import os
import sys
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_tables          # noqa: E402
from oracle import pyoracle               # noqa: E402
from muscle_b200 import synth             # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
def _fa(p):
	out=[]; cur=[]
	for l in open(p):
		if l.startswith(">"):
			if cur: out.append("".join(cur)); cur=[]
		else: cur.append(l.strip())
	if cur: out.append("".join(cur))
	return out
seqs = (_fa(os.path.join(ROOT, "tests", "golden", "rdrp64.fa")) if cfg == "rdrp" else synth.make_config(cfg))[:nseq]
o = pyoracle.Oracle(load_tables())
r = o.all_pairs(seqs, threads=8)
L = [len(s) for s in seqs]
pairs = [(i, j) for i in range(nseq) for j in range(i + 1, nseq)]
M = {}          # (a, b) -> pattern matrix rows = positions of a, cols = positions of b (both orientations)
for (i, j), off, ent in zip(pairs, r["row_off"], r["entries"]):
	m = sp.csr_matrix((np.ones(len(ent), np.float32), ent["col"].astype(np.int64), off.astype(np.int64)), shape=(L[i], L[j]))
	M[(i, j)] = m
	M[(j, i)] = m.T.tocsr()


def tiles(m):
	c = m.tocoo()
	t = sp.coo_matrix((np.ones(len(c.row), np.float32), (c.row//T, c.col//T)),
	  shape=((m.shape[0] + T - 1)//T, (m.shape[1] + T - 1)//T)).tocsr()
	t.data[:] = 1.0
	return t


TM = {k: tiles(v) for k, v in M.items()}
n_entry_z = n_scalar = n_tile_mma = n_ctile = 0
nnz_tiles = sum(TM[p].nnz for p in pairs)
nnz_ent = sum(M[p].nnz for p in pairs)
for (x, y) in pairs:
	pat = M[(x, y)]
	tpat = TM[(x, y)]
	for z in range(nseq):
		if z == x or z == y:
			continue
		prod = (M[(x, z)] @ M[(z, y)]).multiply(pat)            # scalar products that land on stored entries
		n_scalar += int(prod.sum())
		tprod = (TM[(x, z)] @ TM[(z, y)]).multiply(tpat)        # tile products needed to cover them (upper bound: pattern tiles)
		n_tile_mma += int(tprod.sum())
		n_ctile += tpat.nnz
		n_entry_z += pat.nnz
flop_per_mma = 2*T*T*T
print("config %s, %d sequences (mean length %.0f), %d pairs, tile %dx%d" % (cfg, nseq, np.mean(L), len(pairs), T, T))
print("stored entries per pair %.0f, non-empty tiles per pair %.1f, entries per non-empty tile %.2f (of %d cells: %.1f %% fill)"
  % (nnz_ent/len(pairs), nnz_tiles/len(pairs), nnz_ent/nnz_tiles, T*T, 100.0*nnz_ent/nnz_tiles/(T*T)))
print("(entry, z) steps                      %d" % n_entry_z)
print("scalar products needed                %d  (%.2f per (entry, z))" % (n_scalar, n_scalar/n_entry_z))
print("tile MMAs needed (%dx%dx%d)           %d  (%.2f per (pair, z))" % (T, T, T, n_tile_mma, n_tile_mma/(len(pairs)*(nseq - 2))))
print("dense flop in those MMAs / useful flop  %.0f x" % (n_tile_mma*flop_per_mma/(2.0*n_scalar)))
print("flop per (pair, z): scalar %.3g, tiles %.3g (x3 for a 3xTF32 split: %.3g)"
  % (2.0*n_scalar/(len(pairs)*(nseq - 2)), n_tile_mma*flop_per_mma/(len(pairs)*(nseq - 2)),
     3.0*n_tile_mma*flop_per_mma/(len(pairs)*(nseq - 2))))
