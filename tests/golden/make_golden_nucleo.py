"""Nucleotide-alphabet goldens (separate process: the reference initialises its alphabet once).
    python tests/golden/make_golden_nucleo.py
Writes tests/golden/hmm_nucleo.npz (PairHMM tables for ALPHA_Nucleo incl. the U==T fix,
hmmparams.cpp:407-425) and tests/golden/kat_nucleo.npz (dense results of the compiled reference)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PAIRS = [("GATTACA", "GATCACA"), ("ACGUACGUAACCGGUU", "ACGTACGTTACCGGTT"), ("ACGTNNACGT", "acgtryacgt"),
  ("ACGT"*40 + "GGCC", "ACGA"*38 + "TTGGCC")]


def main():
	R = Ref(nucleo=True)
	t = R.tables()
	np.savez_compressed(os.path.join(OUT, "hmm_nucleo.npz"), start=t["start"], trans=t["trans"], ins=t["ins"],
	  match=t["match"], min_sparse_score=np.float32(t["min_sparse_score"]))
	kat = {"n": np.int32(len(PAIRS))}
	for k, (X, Y) in enumerate(PAIRS):
		f, b = R.fwd(X, Y), R.bwd(X, Y)
		p = R.post(X, Y)
		kat["x%d" % k] = np.frombuffer(X.encode(), np.uint8)
		kat["y%d" % k] = np.frombuffer(Y.encode(), np.uint8)
		kat["fwdm%d" % k] = np.ascontiguousarray(f[1:, 1:, 0])
		kat["bwdm%d" % k] = np.ascontiguousarray(b[1:, 1:, 0])
		kat["total%d" % k] = np.float32(R.total(f, b))
		kat["post%d" % k] = p
		kat["alnscore%d" % k] = np.float32(R.alnscore(p))
	np.savez_compressed(os.path.join(OUT, "kat_nucleo.npz"), **kat)
	print("wrote hmm_nucleo.npz kat_nucleo.npz")


if __name__ == "__main__":
	main()
