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
