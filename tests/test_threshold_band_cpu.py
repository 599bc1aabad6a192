"""CPU: the reference thresholds the posterior twice -- CalcPostFlat keeps a cell when
score >= logf(0.01f) (calcposteriorflat.cpp:14-22) and MySparseMx::FromPost keeps it when
expf(score) >= 0.01f (mysparsemx.cpp:139-141).  AlignPairFlat_SparsePost decodes from the DENSE matrix
(alignpairflat.cpp:8-13) while the GPU engine decodes from the stored SPARSE matrix, so the two would
differ if a score could satisfy the first test and fail the second.  With glibc's expf that band is
empty: expf(logf(0.01f)) == 0.01f and expf is monotonic above it -- checked here on the floats right
above the cut (and on a coarse sweep of the whole score range), which is what makes the sparse decode
exact."""
import ctypes
import ctypes.util
import numpy as np
from conftest import load_tables


def test_no_score_passes_the_log_cut_and_fails_the_prob_cut():
	libm = ctypes.CDLL(ctypes.util.find_library("m"))
	libm.expf.restype = ctypes.c_float
	libm.expf.argtypes = [ctypes.c_float]
	libm.logf.restype = ctypes.c_float
	libm.logf.argtypes = [ctypes.c_float]
	cut = np.float32(load_tables()["min_sparse_score"])
	assert np.float32(libm.logf(ctypes.c_float(0.01))) == cut            # the table carries the host's value
	p01 = np.float32(0.01)
	s = cut
	prev = np.float32(0)
	for _ in range(200000):                                              # every float in [cut, cut + ~0.095]
		e = np.float32(libm.expf(ctypes.c_float(float(s))))
		assert e >= p01, (float(s), float(e))
		assert e >= prev                                                    # monotonic
		prev = e
		s = np.nextafter(s, np.float32(0))
	for s in np.linspace(float(cut), 0.0, 20001, dtype=np.float32):
		assert np.float32(libm.expf(ctypes.c_float(float(s)))) >= p01
