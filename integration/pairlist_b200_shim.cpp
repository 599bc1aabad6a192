// integration/pairlist_b200_shim.cpp -- reference-side binding of the PAIR-LIST callers of the hot
// path: everything in `muscle -super5` that calls CalcPost on a
// list of sequence pairs outside MPCFlat.  Replaced at link time:
//   PProg::GetPostPairsAlignedFlat  (getpostpairsalignedflat.cpp:5-98)  -> one mb200_posteriors batch
//   CalcEADistMx                    (eadistmx.cpp:7-70)                 -> one mb200_posteriors batch
//   AlignPairFlat_SparsePost / AlignPairFlat (alignpairflat.cpp:3-31)   -> 1-pair batch (+ mb200_align_pairs)
//   UClust::Search                  (uclust.cpp:26-57)                  -> ONE batch over all top candidates of a query,
//                                                                          then the reference's first-accept rule in order
//   EACluster::GetBestCentroid      (eacluster.cpp:93-143)              -> batches of 32 candidates, the reference's
//                                                                          accept / early-out rule applied in candidate order
// Sequences are resolved through the reference's own global label registry (globalinputms.cpp): the
// whole registry is uploaded ONCE (and again only when it grows or is replaced) and pairs are handed to
// the library as registry indexes (GSI), so a call moves no sequence data.
// A context of its own is used so that an MPCFlat store (mpcflat_b200_shim.cpp) is never disturbed;
// the entry points are serialised by a mutex (the reference calls some of them from OpenMP loops).
#include "muscle.h"
#include "pprog.h"
#include "pairhmm.h"
#include "mega.h"
#include "uclust.h"
#include "eacluster.h"
#include "../include/muscle_b200.h"
#include <mutex>
#include <unordered_map>

static mb200_ctx *g_PairCtx = 0;
static std::mutex g_PairMutex;

static void CheckP(int rc, const char *What)
	{
	if (rc != MB200_OK)
		Die("libmuscle_b200 %s failed (%d): %s", What, rc, mb200_last_error(g_PairCtx));
	}

static void EnsurePairCtx()
	{
	// calcpost.cpp:14 switches to the Mega emission functions when a .mega input is loaded; the
	// engine implements the plain pair-HMM only and must not silently compute something else
	if (Mega::m_Loaded)
		Die("muscle_b200: Mega feature profiles (-mega / .mega input) are not implemented by the B200 engine; "
		  "use the CPU build of muscle for this input");
	if (g_PairCtx == 0)
		{
		int Device = 0;
		const char *s = getenv("MB200_DEVICE");
		if (s != 0)
			Device = atoi(s);
		int rc = mb200_create(Device, &g_PairCtx);
		if (rc != MB200_OK)
			Die("libmuscle_b200 mb200_create failed (%d): %s", rc, mb200_last_error(0));
		}
	// the PairHMM tables are process-global statics that can be rewritten between replicates
	// (align.cpp:30-40): upload them when they differ from what the device holds
	static vector<float> Cached;
	vector<float> Now;
	Now.insert(Now.end(), PairHMM::m_StartScore, PairHMM::m_StartScore + HMMSTATE_COUNT);
	Now.insert(Now.end(), &PairHMM::m_TransScore[0][0], &PairHMM::m_TransScore[0][0] + HMMSTATE_COUNT*HMMSTATE_COUNT);
	Now.insert(Now.end(), PairHMM::m_InsScore, PairHMM::m_InsScore + 256);
	Now.insert(Now.end(), &PairHMM::m_MatchScore[0][0], &PairHMM::m_MatchScore[0][0] + 256*256);
	if (Now.size() != Cached.size() || memcmp(Now.data(), Cached.data(), Now.size()*sizeof(float)) != 0)
		{
		CheckP(mb200_set_hmm(g_PairCtx, PairHMM::m_StartScore, &PairHMM::m_TransScore[0][0],
		  PairHMM::m_InsScore, &PairHMM::m_MatchScore[0][0], MIN_SPARSE_SCORE), "mb200_set_hmm");
		Cached.swap(Now);
		}
	}

// The sequences the pair lists name live on the device persistently: the input registry
// (globalinputms.cpp:68-93) is uploaded as a whole the first time, temporary sequences (consensus
// sequences registered with AddGlobalTmpSeq, :60-66, known by label only) are appended when first seen,
// and everything is uploaded again only when something was appended or a label now names other bytes.
struct DevSeq { uint Id; const Sequence *Ptr; uint Length; };
static unordered_map<string, DevSeq> g_LabelToDev;
static vector<byte> g_DevBytes;
static vector<uint64_t> g_DevOffsets(1, 0);
static bool g_DevDirty = false;
static const void *g_RegistryFirst = 0;

static uint AppendDevSeq(const string &Label, const Sequence *Seq)
	{
	const byte *B = Seq->GetBytePtr();
	const uint L = Seq->GetLength();
	DevSeq D;
	D.Id = SIZE(g_DevOffsets) - 1;
	D.Ptr = Seq;
	D.Length = L;
	g_DevBytes.insert(g_DevBytes.end(), B, B + L);
	g_DevOffsets.push_back(g_DevBytes.size());
	g_LabelToDev[Label] = D;
	g_DevDirty = true;
	return D.Id;
	}

static uint DevIdOfLabel(const string &Label)
	{
	const Sequence *Seq = &GetGlobalInputSeqByLabel(Label);          // Die()s on an unknown label, like the reference
	unordered_map<string, DevSeq>::const_iterator it = g_LabelToDev.find(Label);
	if (it != g_LabelToDev.end())
		{
		const DevSeq &D = it->second;
		if (D.Ptr == Seq && D.Length == Seq->GetLength() &&
		  memcmp(g_DevBytes.data() + g_DevOffsets[D.Id], Seq->GetBytePtr(), D.Length) == 0)
			return D.Id;
		}
	return AppendDevSeq(Label, Seq);                                  // new, or the label now names another sequence
	}

static void EnsureRegistry()
	{
	const uint Count = GetGSICount();
	asserta(Count > 0);
	const void *First = GetSequenceByGSI(0);
	if (First == g_RegistryFirst)
		return;
	// first use, or the registry was replaced: start over
	g_LabelToDev.clear();
	g_DevBytes.clear();
	g_DevOffsets.assign(1, 0);
	for (uint GSI = 0; GSI < Count; ++GSI)
		{
		const Sequence *Seq = GetSequenceByGSI(GSI);
		AppendDevSeq(Seq->m_Label, Seq);
		}
	g_RegistryFirst = First;
	}

// Run the posterior stage on a pair list given by labels.  EAs[k] = what CalcAlnFlat(Post)/min(L1,L2)
// gives in the reference (the max-sum DP value of CalcAlnFlat and CalcAlnScoreFlat is the same number).
static void RunPairList(const vector<string> &Labels1, const vector<string> &Labels2, vector<float> &EAs)
	{
	const uint PairCount = SIZE(Labels1);
	asserta(SIZE(Labels2) == PairCount && PairCount > 0);
	EnsureRegistry();
	vector<uint> PX(PairCount), PY(PairCount);
	for (uint k = 0; k < PairCount; ++k)
		{
		PX[k] = DevIdOfLabel(Labels1[k]);
		PY[k] = DevIdOfLabel(Labels2[k]);
		}
	if (g_DevDirty)
		{
		CheckP(mb200_set_seqs(g_PairCtx, SIZE(g_DevOffsets) - 1, g_DevBytes.data(), g_DevOffsets.data()), "mb200_set_seqs");
		g_DevDirty = false;
		}
	EAs.resize(PairCount);
	CheckP(mb200_posteriors(g_PairCtx, PairCount, PX.data(), PY.data(), MB200_POST_DEFAULT, EAs.data()),
	  "mb200_posteriors");
	}

// decoding path of one pair of the current store (CalcAlnFlat + traceback on the device)
static void StorePairPath(uint StorePair, uint LX, uint LY, string &Path)
	{
	uint64_t PathOff[2] = { 0, uint64_t(LX) + LY + 1 };
	vector<char> PathBuf(LX + LY + 2);
	float Score = 0;
	CheckP(mb200_align_pairs(g_PairCtx, 1, &StorePair, PathBuf.data(), PathOff, &Score), "mb200_align_pairs");
	Path = string(PathBuf.data());
	}

// fill host MySparseMx objects (MySparseMx::FromPost layout) for store pairs [0,PairCount)
static void ExportSparse(const vector<string> &Labels1, vector<MySparseMx *> &SparsePosts)
	{
	const uint PairCount = SIZE(SparsePosts);
	vector<uint> Nnz(PairCount);
	uint64_t Total = 0;
	CheckP(mb200_store_nnz(g_PairCtx, Nnz.data(), &Total), "mb200_store_nnz");
	uint64_t Rows = 0;
	vector<uint> LX(PairCount);
	for (uint k = 0; k < PairCount; ++k)
		{
		LX[k] = GetSeqLengthByGlobalLabel(Labels1[k]);
		Rows += LX[k] + 1;
		}
	vector<uint> Offs(Rows);
	vector<mb200_entry> Ents(Total + 1);
	CheckP(mb200_export_all(g_PairCtx, Offs.data(), Ents.data()), "mb200_export_all");
	uint64_t r = 0, e = 0;
	for (uint k = 0; k < PairCount; ++k)
		{
		MySparseMx *S = SparsePosts[k];
		S->AllocLX(LX[k]);
		memcpy(S->m_Offsets, Offs.data() + r, (LX[k] + 1)*sizeof(uint));
		S->m_VecSize = Nnz[k];
		S->AllocVec(Nnz[k]);
		memcpy(S->m_ValueVec, Ents.data() + e, size_t(Nnz[k])*8);
		r += LX[k] + 1;
		e += Nnz[k];
		}
	}

float PProg::GetPostPairsAlignedFlat(const string &aProgressStr,
  const MultiSequence &MSA1, const MultiSequence &MSA2,
  const vector<uint> &SeqIndexes1, const vector<uint> &SeqIndexes2,
  vector<MySparseMx *> &SparsePosts)
	{
	const uint PairCount = SIZE(SeqIndexes1);
	asserta(SIZE(SeqIndexes2) == PairCount);
	asserta(SparsePosts.empty());
	std::lock_guard<std::mutex> Guard(g_PairMutex);
	EnsurePairCtx();
	ProgressStep(0, 2, "%s [%u x %u, %u pairs] (B200)", aProgressStr.substr(0, 20).c_str(),
	  min(MSA1.GetSeqCount(), MSA2.GetSeqCount()), max(MSA1.GetSeqCount(), MSA2.GetSeqCount()), PairCount);
	vector<string> Labels1, Labels2;
	for (uint k = 0; k < PairCount; ++k)
		{
		Labels1.push_back(MSA1.GetLabelStr(SeqIndexes1[k]));
		Labels2.push_back(MSA2.GetLabelStr(SeqIndexes2[k]));
		}
	vector<float> EAs;
	RunPairList(Labels1, Labels2, EAs);
	SparsePosts.resize(PairCount);
	for (uint k = 0; k < PairCount; ++k)
		{
		MySparseMx *S = new MySparseMx;
		S->m_LX = GetSeqLengthByGlobalLabel(Labels1[k]);
		S->m_LY = GetSeqLengthByGlobalLabel(Labels2[k]);
		SparsePosts[k] = S;
		}
	ExportSparse(Labels1, SparsePosts);
	float SumEA = 0;
	for (uint k = 0; k < PairCount; ++k)
		SumEA += EAs[k];
	ProgressStep(1, 2, "%s [%u pairs] (B200)", aProgressStr.substr(0, 20).c_str(), PairCount);
	return SumEA/PairCount;
	}

void CalcEADistMx(FILE *f, MultiSequence *sequences,
  vector<vector<float> > &DistMx, vector<MySparseMx *> *SparsePostVec)
	{
	DistMx.clear();
	const uint SeqCount = sequences->GetSeqCount();
	DistMx.resize(SeqCount);
	for (uint i = 0; i < SeqCount; ++i)
		{
		DistMx[i].resize(SeqCount, 0);
		DistMx[i][i] = 1;
		}
	if (SparsePostVec != 0)
		asserta(SIZE(*SparsePostVec) == 0);
	if (SeqCount < 2)
		return;
	std::lock_guard<std::mutex> Guard(g_PairMutex);
	EnsurePairCtx();
	vector<uint> I1, I2;
	vector<string> Labels1, Labels2;
	for (uint i = 0; i < SeqCount; ++i)                 // GetAllPairs order: i<j row-major (getpairs.cpp)
		for (uint j = i + 1; j < SeqCount; ++j)
			{
			I1.push_back(i);
			I2.push_back(j);
			Labels1.push_back(sequences->GetSequence(i)->m_Label);
			Labels2.push_back(sequences->GetSequence(j)->m_Label);
			}
	const uint PairCount = SIZE(I1);
	ProgressStep(0, 2, "%u consensus seqs (B200)", SeqCount);
	vector<float> EAs;
	RunPairList(Labels1, Labels2, EAs);
	if (SparsePostVec != 0)
		{
		for (uint k = 0; k < PairCount; ++k)
			{
			MySparseMx *S = new MySparseMx;
			S->m_LX = GetSeqLengthByGlobalLabel(Labels1[k]);
			S->m_LY = GetSeqLengthByGlobalLabel(Labels2[k]);
			SparsePostVec->push_back(S);
			}
		ExportSparse(Labels1, *SparsePostVec);
		}
	for (uint k = 0; k < PairCount; ++k)
		{
		DistMx[I1[k]][I2[k]] = EAs[k];
		DistMx[I2[k]][I1[k]] = EAs[k];
		if (f != 0)
			fprintf(f, "%s\t%s\t%.4g\n", Labels1[k].c_str(), Labels2[k].c_str(), EAs[k]);
		}
	ProgressStep(1, 2, "%u consensus seqs (B200)", SeqCount);
	}

float AlignPairFlat_SparsePost(const string &Label1, const string &Label2,
  string &Path, MySparseMx *SparsePost)
	{
	std::lock_guard<std::mutex> Guard(g_PairMutex);
	EnsurePairCtx();
	vector<string> L1(1, Label1), L2(1, Label2);
	vector<float> EAs;
	RunPairList(L1, L2, EAs);
	const uint LX = GetSeqLengthByGlobalLabel(Label1);
	const uint LY = GetSeqLengthByGlobalLabel(Label2);
	// the decoding path: max-sum DP + traceback on the stored sparse posterior.  The reference decodes
	// from the dense thresholded matrix (alignpairflat.cpp:8-13); the two hold the same cells because no
	// score passes the log cut and fails the probability cut (tests/test_threshold_band_cpu.py).
	StorePairPath(0, LX, LY, Path);
	if (SparsePost != 0)
		{
		SparsePost->m_LX = LX;
		SparsePost->m_LY = LY;
		vector<MySparseMx *> One(1, SparsePost);
		ExportSparse(L1, One);
		}
	return EAs[0];
	}

float AlignPairFlat(const string &Label1, const string &Label2, string &Path)
	{
	return AlignPairFlat_SparsePost(Label1, Label2, Path, 0);
	}


// uclust.cpp:26-57.  The reference aligns the query to its top word-count candidates ONE AT A TIME and
// stops at the first with EA >= MinEA; the candidates are known up front and the alignments are pure
// functions, so all of them run as one device batch and the first-accept rule is applied afterwards in
// the same order -- same centroid, same path.
uint UClust::Search(uint SeqIndex, string &Path)
	{
	const Sequence *Seq = m_InputSeqs->GetSequence(SeqIndex);
	const byte *ByteSeq = Seq->GetBytePtr();
	const uint L = Seq->GetLength();

	vector<uint> TopSeqIndexes;
	vector<uint> TopWordCounts;
	m_US.SearchSeq(ByteSeq, L, TopSeqIndexes, TopWordCounts);
	uint TopCount = SIZE(TopSeqIndexes);
	asserta(SIZE(TopWordCounts) == TopCount);
	if (TopCount == 0)
		return UINT_MAX;
	if (TopCount > MAX_REJECTS)
		TopCount = MAX_REJECTS;

	std::lock_guard<std::mutex> Guard(g_PairMutex);
	EnsurePairCtx();
	vector<string> L1, L2;
	for (uint TopIndex = 0; TopIndex < TopCount; ++TopIndex)
		{
		L1.push_back(Seq->m_Label);
		L2.push_back(m_InputSeqs->GetSequence(TopSeqIndexes[TopIndex])->m_Label);
		}
	vector<float> EAs;
	RunPairList(L1, L2, EAs);
	Path.clear();
	for (uint TopIndex = 0; TopIndex < TopCount; ++TopIndex)
		if (EAs[TopIndex] >= m_MinEA)
			{
			const uint TopSeqIndex = TopSeqIndexes[TopIndex];
			StorePairPath(TopIndex, L, m_InputSeqs->GetSequence(TopSeqIndex)->GetLength(), Path);
			return TopSeqIndex;
			}
	return UINT_MAX;
	}

// eacluster.cpp:93-143.  The reference evaluates the candidates in an OpenMP loop with a shared early-out
// flag (its result depends on thread timing); here they are evaluated in device batches of 32 and the
// accept / early-out rule is applied in candidate order, i.e. exactly the reference's `-threads 1` result.
uint EACluster::GetBestCentroid(uint SeqIndex, float MinEA, float &BestEA)
	{
	uint CentroidCount = SIZE(m_CentroidSeqIndexes);
	if (CentroidCount == 0)
		return UINT_MAX;

	uint L;
	const byte *ByteSeq = m_InputSeqs->GetByteSeq(SeqIndex, L);

	vector<uint> TopSeqIndexes;
	vector<uint> TopWordCounts;
	m_US.SearchSeq(ByteSeq, L, TopSeqIndexes, TopWordCounts);
	const uint TopCount = SIZE(TopSeqIndexes);
	asserta(SIZE(TopWordCounts) == TopCount);
	if (TopCount == 0)
		return UINT_MAX;

	std::lock_guard<std::mutex> Guard(g_PairMutex);
	EnsurePairCtx();
	BestEA = 0;
	uint BestCentroidIndex = UINT_MAX;
	bool Done = false;
	const string Label = m_InputSeqs->GetLabel(SeqIndex);
	const uint CHUNK = 32;
	for (uint Top0 = 0; Top0 < TopCount && !Done; Top0 += CHUNK)
		{
		const uint n = min(CHUNK, TopCount - Top0);
		vector<string> L1(n, Label), L2;
		for (uint k = 0; k < n; ++k)
			L2.push_back(m_InputSeqs->GetLabel(TopSeqIndexes[Top0 + k]));
		vector<float> EAs;
		RunPairList(L1, L2, EAs);
		for (uint k = 0; k < n && !Done; ++k)
			{
			const int TopIndex = int(Top0 + k);
			const uint TopSeqIndex = TopSeqIndexes[TopIndex];
			const float EA = EAs[k];
			if (EA > MinEA && EA > BestEA)
				{
				BestEA = EA;
				asserta(TopSeqIndex < SIZE(m_SeqIndexToCentroidIndex));
				uint CentroidIndex = m_SeqIndexToCentroidIndex[TopSeqIndex];
				asserta(CentroidIndex < CentroidCount);
				BestCentroidIndex = CentroidIndex;
				}
			if (BestEA >= MinEA)
				{
				if (BestEA > 0.9)
					Done = true;
				if (BestEA - EA > 0.3)
					Done = true;
				}
			if (BestEA < MinEA - 0.3 && TopIndex > 20)
				Done = true;
			}
		}
	return BestCentroidIndex;
	}
