"""Test helper: the Mega emissions (Mega::GetInsScore / GetMatchScore, mega.cpp:273-359) restated with
numpy fp32 operations, and a way to run them through the PINNED plain oracle: for a pair with both
lengths <= 128 the positions themselves become the letters (X position i -> byte i, Y position j ->
byte 128+j) of a synthetic PairHMM table whose match[i][128+j] is the pair emission and whose ins[]
holds the per-position insert emissions, so mo_fwd / mo_bwd / mo_post run unchanged."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_model():
	z = np.load(os.path.join(GOLDEN, "mega_bb11001.npz"))
	model = {k: z[k] for k in ("alpha", "weights", "logprobs", "logprobmx")}
	n = int(z["n"])
	profiles = [z["letters%d" % i] for i in range(n)]
	posts = {(i, j): z["post_%d_%d" % (i, j)] for i in range(n) for j in range(i + 1, n)}
	return model, profiles, posts


def emissions(model, PX, PY):
	"""-> (insx[LX], insy[LY], match[LX,LY]) accumulated from 0 in feature order, fp32 multiply then add"""
	alpha = model["alpha"].astype(np.int64)
	w = model["weights"].astype(np.float32)
	lb = np.concatenate([[0], np.cumsum(alpha)])
	mb = np.concatenate([[0], np.cumsum(alpha*alpha)])
	insx = np.zeros(len(PX), np.float32)
	insy = np.zeros(len(PY), np.float32)
	match = np.zeros((len(PX), len(PY)), np.float32)
	for f in range(len(alpha)):
		lp = model["logprobs"][lb[f]:lb[f + 1]].astype(np.float32)
		mx = model["logprobmx"][mb[f]:mb[f + 1]].astype(np.float32).reshape(alpha[f], alpha[f])
		insx = (insx + lp[PX[:, f]]*w[f]).astype(np.float32)
		insy = (insy + lp[PY[:, f]]*w[f]).astype(np.float32)
		match = (match + mx[PX[:, f]][:, PY[:, f]]*w[f]).astype(np.float32)
	return insx, insy, match


def synthetic_tables(base_tables, model, PX, PY):
	"""PairHMM tables + byte sequences under which the plain oracle computes the Mega pair (PX,PY)"""
	LX, LY = len(PX), len(PY)
	assert LX <= 128 and LY <= 128
	insx, insy, match = emissions(model, PX, PY)
	t = dict(base_tables)
	ins = np.zeros(256, np.float32)
	ins[:LX] = insx
	ins[128:128 + LY] = insy
	m = np.zeros((256, 256), np.float32)
	m[:LX, 128:128 + LY] = match
	t["ins"], t["match"] = ins, m.reshape(-1)
	X = bytes(range(LX))
	Y = bytes(range(128, 128 + LY))
	return t, X, Y
