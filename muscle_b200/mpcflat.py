"""Host-side mirror of the reference's MPCFlat call surface for the hot path (mpcflat.h:16-105),
driving the CUDA library through its C ABI.  Method names, argument meaning and phase order follow
the reference so that tests read like the reference's own code:

    M = MPCFlat(tables); M.InitSeqs(seqs); M.InitPairs(); M.InitDistMx()
    M.CalcPosteriors(); M.Consistency(); path, score = M.AlignAlns(msa1, msa2)

Everything numerical happens on the GPU; host code here is index bookkeeping only (the reference's
guide tree / progressive-alignment control flow stays in the reference's own C++).
With torch.distributed initialised (one process per GPU) the pair loop is sharded and the two
exchange steps of muscle_b200.dist are inserted; results are identical on every rank.
"""
import numpy as np

from .engine import Engine, MB200Error  # noqa: F401

DEFAULT_CONSISTENCY_ITERS_FLAT = 2      # mpcflat.h:12


class MPCFlat:
	def __init__(self, tables, device=0, group=None):
		self.engine = Engine(device)
		self.engine.set_hmm(tables)
		self.group = group
		self.m_ConsistencyIterCount = DEFAULT_CONSISTENCY_ITERS_FLAT
		self.m_MyInputSeqs = None
		self.m_Pairs = []
		self.m_DistMx = None
		self._ranges = None
		self._eranges = None
		self.timings = {}
		self._rank, self._world = 0, 1
		try:
			import torch.distributed as dist
			if dist.is_available() and dist.is_initialized():
				self._rank, self._world = dist.get_rank(group), dist.get_world_size(group)
		except Exception:
			pass

	def close(self):
		self.engine.close()

	# ---- mpcflat.cpp:115-171
	def InitSeqs(self, seqs):
		self.m_MyInputSeqs = [s if isinstance(s, (bytes, bytearray)) else s.encode() for s in seqs]
		self.engine.set_seqs(self.m_MyInputSeqs)

	def GetSeqCount(self):
		return len(self.m_MyInputSeqs)

	def GetSeqLength(self, i):
		return len(self.m_MyInputSeqs[i])

	def InitPairs(self):
		n = self.GetSeqCount()
		self.m_Pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]      # row-major i<j

	def GetPairIndex(self, a, b):
		assert a < b
		n = self.GetSeqCount()
		return a*n - a*(a + 1)//2 + (b - a - 1)

	def InitDistMx(self):
		n = self.GetSeqCount()
		self.m_DistMx = np.full((n, n), np.finfo(np.float32).max, np.float32)
		np.fill_diagonal(self.m_DistMx, 0)

	# ---- mpcflat.cpp:214-252 (the OpenMP pair loop becomes one batch call per rank)
	def CalcPosteriors(self):
		from . import dist as mdist
		n = self.GetSeqCount()
		lens = [len(s) for s in self.m_MyInputSeqs]
		self._ranges, _, _ = mdist.shard_ranges(lens, self._world)
		self._eranges = None
		lo, hi = self._ranges[self._rank]
		ea = self.engine.posteriors_allpairs(lo, hi) if hi > lo else np.zeros(0, np.float32)
		self.timings["posterior_ms"] = self.engine.stats()["last_total_ms"] if hi > lo else 0.0
		if self._world > 1:
			nbytes, secs = mdist.gather_store(self.engine, self.group, have_store=hi > lo)
			self.timings["exchange1_bytes"] = nbytes
			self.timings["exchange1_ms"] = secs*1e3
		m = mdist.gather_ea(ea, self._ranges, n, self.group, distributed=self._world > 1)
		iu = np.triu_indices(n, 1)
		self.m_DistMx[iu] = m[iu]
		self.m_DistMx.T[iu] = m[iu]

	# ---- mpcflat.cpp:173-181, consflat.cpp:5-23
	def Consistency(self):
		if self.GetSeqCount() < 3:
			return
		for it in range(self.m_ConsistencyIterCount):
			self.ConsIter(it)

	def ConsIter(self, it=0):
		from . import dist as mdist
		lo, hi = self._ranges[self._rank] if self._ranges else (0, len(self.m_Pairs))
		self.engine.consistency_iter(lo, hi)
		st = self.engine.stats()
		self.timings["relax_ms"] = st["last_total_ms"]
		self.timings["relax_kernel_ms"] = st["last_kernel_ms"]
		if self._world > 1:
			if self._eranges is None:
				nnz, _ = self.engine.store_nnz()
				base = np.concatenate([[0], np.cumsum(nnz.astype(np.int64))])
				self._eranges = [(int(base[a]), int(base[b])) for (a, b) in self._ranges]
			nbytes, secs = mdist.gather_values(self.engine, self._eranges, self._rank, self.group)
			self.timings["exchange2_bytes"] = nbytes
			self.timings["exchange2_ms"] = secs*1e3

	def GetSparsePost(self, pair_index):
		"""(offsets[LX+1], entries[nnz]) in MySparseMx layout (mpcflat.cpp:88-98)"""
		return self.engine.export_pair(pair_index)

	# ---- alnalnsflat.cpp:7-52 (BuildPost + CalcAlnFlat; the caller inserts the gaps)
	def AlignAlns(self, msa1, msa2):
		"""msa = list of (sequence index, gapped row string); returns (path over B/X/Y, score)"""
		def maps(msa):
			ids = [i for i, _ in msa]
			p2c = [np.array([c for c, ch in enumerate(row) if ch != "-"], np.uint32) for _, row in msa]
			return ids, p2c, len(msa[0][1])
		ia, pa, ca = maps(msa1)
		ib, pb, cb = maps(msa2)
		score, path, _ = self.engine.align_groups(ia, pa, ca, ib, pb, cb)
		return path, score

	def AlignPairs(self, pair_indexes):
		"""batched AlignPairFlat (alignpairflat.cpp:23): (EA-style DP scores, paths)"""
		return self.engine.align_pairs(pair_indexes)


def add_gaps_path(row, path, letter):
	"""Sequence::AddGapsPath (sequence.cpp:115-140): consume one char of `row` on 'B' or `letter`,
	emit '-' on the other insert letter."""
	out = []
	k = 0
	for ch in path:
		if ch == "B" or ch == letter:
			out.append(row[k])
			k += 1
		else:
			out.append("-")
	return "".join(out)
