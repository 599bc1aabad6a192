"""Golden vectors for the Mega (Muscle-3D feature profile) emission mode, produced by the COMPILED
REFERENCE (oracle/_ref): the model Mega::FromFile derives from tests/golden/e2e/BB11001.mega (the
reference's own test_data/mega/BB11001.mega), the feature profiles, and the dense posterior of every
pair computed by Mega::CalcFwdFlat_mega + CalcBwdFlat_mega + CalcPostFlat (calcpost.cpp:14-35).
    python tests/golden/make_golden_mega.py
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle   # noqa: E402

MEGA = os.path.join(ROOT, "tests", "golden", "e2e", "BB11001.mega")
OUT = os.path.join(ROOT, "tests", "golden", "mega_bb11001.npz")


def main():
	R = pyoracle.Ref()
	L = R.lib
	n = L.ref_mega_load(MEGA.encode())
	F = L.ref_mega_nfeat()
	alpha = np.zeros(F, np.uint32)
	w = np.zeros(F, np.float32)
	lp = np.zeros(4096, np.float32)
	lpm = np.zeros(65536, np.float32)
	L.ref_mega_model(C.c_void_p(alpha.ctypes.data), C.c_void_p(w.ctypes.data), C.c_void_p(lp.ctypes.data), C.c_void_p(lpm.ctypes.data))
	lp = lp[:int(alpha.sum())].copy()
	lpm = lpm[:int((alpha.astype(np.int64)**2).sum())].copy()
	out = {"n": n, "nfeat": F, "alpha": alpha, "weights": w, "logprobs": lp, "logprobmx": lpm}
	lens = []
	for i in range(n):
		Li = L.ref_mega_profile_len(i)
		let = np.zeros((Li, F), np.uint8)
		seq = C.create_string_buffer(Li + 1)
		L.ref_mega_profile(i, C.c_void_p(let.ctypes.data), seq)
		out["letters%d" % i] = let
		out["seq%d" % i] = np.frombuffer(seq.value, np.uint8)
		lens.append(Li)
	out["lens"] = np.array(lens, np.uint32)
	for i in range(n):
		for j in range(i + 1, n):
			post = np.zeros((lens[i], lens[j]), np.float32)
			L.ref_mega_calcpost(i, j, C.c_void_p(post.ctypes.data))
			out["post_%d_%d" % (i, j)] = post
	np.savez_compressed(OUT, **out)
	print("wrote", OUT, "n", n, "features", F, "alpha", alpha, "lens", lens)
	# end-to-end golden: the reference CLI on the same file
	subprocess.run([pyoracle.REF_CLI, "-align", MEGA, "-output", os.path.join(ROOT, "tests", "golden", "e2e", "BB11001.mega.ref.afa"),
	  "-quiet"], check=True)


if __name__ == "__main__":
	main()
