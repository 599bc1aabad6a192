import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_tables():
	z = np.load(os.path.join(GOLDEN, "hmm_amino.npz"))
	return {k: z[k] for k in ("start", "trans", "ins", "match", "min_sparse_score")}


@pytest.fixture(scope="session")
def tables():
	return load_tables()


@pytest.fixture(scope="session")
def oracle(tables):
	from oracle.pyoracle import Oracle, build
	build(ref=False)
	return Oracle(tables)


@pytest.fixture(scope="session")
def ref():
	"""The compiled reference; present where oracle/_ref was built (this container; it also
	travels to the GPU box as a prebuilt .so)."""
	from oracle.pyoracle import Ref, REF_SO
	if not os.path.exists(REF_SO):
		pytest.skip("oracle/_ref/libmuscle_ref.so not built")
	return Ref()


@pytest.fixture(scope="session")
def engine(tables):
	import torch
	if not torch.cuda.is_available():
		pytest.skip("no CUDA device")
	from muscle_b200.engine import Engine
	e = Engine(0)
	e.set_hmm(tables)
	yield e
	e.close()
