"""CPU: the C-ABI library loads and exports every symbol include/muscle_b200.h declares; without a
GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
	src = open(os.path.join(ROOT, "include", "muscle_b200.h")).read()
	src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
	return sorted(set(re.findall(r"\b(mb200_[a-z_0-9]+)\s*\(", src)))


def test_exports_match_header():
	from muscle_b200.engine import load_library, EXPORTS
	lib = load_library()
	names = _declared()
	assert len(names) >= 20
	for n in names:
		assert hasattr(lib, n), "library does not export %s" % n
	assert set(EXPORTS) <= set(names)


def test_no_cpu_fallback():
	import torch
	from muscle_b200.engine import load_library
	lib = load_library()
	assert b"sm_100a" in lib.mb200_version()
	if torch.cuda.is_available():
		return
	h = C.c_void_p()
	rc = lib.mb200_create(0, C.byref(h))
	assert rc == -2 and not h.value          # MB200_ENODEV
	assert b"no CPU path" in lib.mb200_last_error(None)


def test_product_never_touches_oracle():
	"""the package must not import or dlopen anything under oracle/"""
	pkg = os.path.join(ROOT, "muscle_b200")
	for dp, _, files in os.walk(pkg):
		for f in files:
			if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
				txt = open(os.path.join(dp, f), errors="ignore").read()
				assert "pyoracle" not in txt and "liboracle" not in txt and "libmuscle_ref" not in txt, f


def test_residue_classes_host_logic():
	"""host-side table compaction (no GPU needed): 20 amino letters x 2 cases + wildcard -> 21 classes;
	nucleotides: A C G T(=U) + wildcard -> 5 classes"""
	import numpy as np
	from muscle_b200.engine import load_library
	lib = load_library()
	for name, want, same, diff in (("hmm_amino.npz", 21, [("A", "a"), ("X", "*"), ("B", "Z")], [("A", "C"), ("L", "I")]),
	  ("hmm_nucleo.npz", 5, [("T", "U"), ("t", "u"), ("N", "R"), ("A", "a")], [("A", "C"), ("G", "T")])):
		z = np.load(os.path.join(ROOT, "tests", "golden", name))
		ins = np.ascontiguousarray(z["ins"], np.float32)
		match = np.ascontiguousarray(z["match"], np.float32).reshape(-1)
		b2c = np.zeros(256, np.uint8)
		n = C.c_int()
		rc = lib.mb200_residue_classes(C.c_void_p(ins.ctypes.data), C.c_void_p(match.ctypes.data),
		  C.c_void_p(b2c.ctypes.data), C.byref(n), None)
		assert rc == 0 and n.value == want, (name, n.value)
		for a, b in same:
			assert b2c[ord(a)] == b2c[ord(b)], (name, a, b)
		for a, b in diff:
			assert b2c[ord(a)] != b2c[ord(b)], (name, a, b)
		# a class reproduces the table entries of every member byte
		m = match.reshape(256, 256)
		for a in b"ACGTUNacdwy*":
			for b in b"ACGTUNlkxz":
				ra = int(np.flatnonzero(b2c == b2c[a])[0])
				rb = int(np.flatnonzero(b2c == b2c[b])[0])
				assert m[a, b] == m[ra, rb] and ins[a] == ins[ra]
