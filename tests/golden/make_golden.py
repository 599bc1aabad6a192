"""Generate the golden fixtures in tests/golden/ from the COMPILED, UNMODIFIED reference
(oracle/_ref/libmuscle_ref.so = /root/reference/src built with -ffp-contract=off, see oracle/Makefile).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures are small .npz files; tests on the GPU box read only these (never /root/reference).
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref            # noqa: E402
from muscle_b200 import synth              # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# named known-answer inputs of the reference's disabled developer harness (src/testfb.cpp:369-405)
KAT = [("MQTIF", "MSIF"), ("GATTACA", "MQTIF"), ("ABC", "DEF"), ("LQNGSEQVENCE", "QTHERSEQVENCEINSERT"),
  ("A", "C"), ("A", "ACDEFGHIKL"), ("ACDEFGHIKL", "W"), ("MKV", "MKV")]


def main():
	R = Ref()
	t = R.tables()
	# the 256x256 match table has 21 distinct rows; npz compression makes it tiny
	np.savez_compressed(os.path.join(OUT, "hmm_amino.npz"), start=t["start"], trans=t["trans"], ins=t["ins"],
	  match=t["match"], min_sparse_score=np.float32(t["min_sparse_score"]))

	kat = {}
	for k, (X, Y) in enumerate(KAT):
		f, b = R.fwd(X, Y), R.bwd(X, Y)
		p = R.post(X, Y)
		off, ent = R.sparse(p)
		score, path = R.calcaln(p)
		kat["x%d" % k] = np.frombuffer(X.encode(), np.uint8)
		kat["y%d" % k] = np.frombuffer(Y.encode(), np.uint8)
		kat["fwd%d" % k] = f
		kat["bwd%d" % k] = b
		kat["total%d" % k] = np.float32(R.total(f, b))
		kat["post%d" % k] = p
		kat["off%d" % k] = off
		kat["ent%d" % k] = ent
		kat["alnscore%d" % k] = np.float32(R.alnscore(p))
		kat["calcaln%d" % k] = np.float32(score)
		kat["path%d" % k] = np.frombuffer(path.encode(), np.uint8)
	kat["n"] = np.int32(len(KAT))
	np.savez_compressed(os.path.join(OUT, "kat_pairs.npz"), **kat)

	# a small family through the whole MPCFlat pipeline
	seqs = synth.make_family(8, 60, 8, seed=11)
	fam = {"n": np.int32(len(seqs))}
	for i, s in enumerate(seqs):
		fam["seq%d" % i] = np.frombuffer(s.encode(), np.uint8)
	M = R.mpc(seqs)
	M.posteriors()
	fam["ea"] = M.distmx()
	offs, ents = M.export_all()
	for p in range(len(offs)):
		fam["off%d" % p] = offs[p]
		fam["ent0_%d" % p] = ents[p]
	# dense (pre-sparsification) posterior of pair 0 and its decoding
	post01 = R.post(seqs[0], seqs[1])
	fam["post01"] = post01
	sc, path = R.calcaln(post01)
	fam["path01"] = np.frombuffer(path.encode(), np.uint8)
	fam["score01"] = np.float32(sc)
	for it in (1, 2):
		M.consiter()
		_, ents = M.export_all()
		for p in range(len(offs)):
			fam["ent%d_%d" % (it, p)] = ents[p]
	M.close()
	# final MSA with the reference's defaults (2 consistency + 100 refinement iterations)
	M = R.mpc(seqs)
	M.posteriors()
	rows = M.finish(2, 100)
	fam["msa_idx"] = np.array([r[0] for r in rows], np.int32)
	fam["msa_rows"] = np.array([np.frombuffer(r[1].encode(), np.uint8) for r in rows])
	# one progressive-alignment join on the consistency-transformed store: groups {0,1,2} + {3,4}
	M.close()
	np.savez_compressed(os.path.join(OUT, "family8.npz"), **fam)
	print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
	main()
