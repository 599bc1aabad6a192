// integration/mpcflat_b200_shim.cpp -- the reference-side binding a MUSCLE maintainer would add to
// make libmuscle_b200.so the pair engine of `muscle -align` / `-super5`.
//
// It is compiled against the UNMODIFIED reference headers and replaces, at link time, exactly the
// three MPCFlat members that own the hot loops:
//   MPCFlat::CalcPosteriors   (mpcflat.cpp:214-252)  -> mb200_posteriors_allpairs
//   MPCFlat::ConsIter         (consflat.cpp:5-23)    -> mb200_consistency_iter
//   MPCFlat::AlignAlns        (alnalnsflat.cpp:7-52) -> mb200_align_groups (+ the reference's own gap insertion)
// Everything else (FASTA I/O, dereplication, UPGMA guide tree, join order, progressive alignment
// and refinement control flow, MSA output) is the reference's own code, untouched.
// Errors keep the reference convention: Die() -> message + exit(1) (myutils.cpp:883).
#include "muscle.h"
#include "mpcflat.h"
#include "pairhmm.h"
#include "../include/muscle_b200.h"

#include <chrono>
static mb200_ctx *g_Ctx = 0;

// MB200_TRACE=1: wall-time split of the replaced members, printed at exit
static double g_TPrep = 0, g_TLib = 0, g_TGaps = 0, g_TPost = 0, g_TCons = 0;
static uint g_NAlign = 0;
static double Now()
	{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}
static void TraceReport()
	{
	fprintf(stderr, "\n[mb200 trace] CalcPosteriors %.2f s, ConsIter %.2f s, AlignAlns x%u: prepare %.2f s, "
	  "mb200_align_groups %.2f s, AddGapsPath %.2f s\n", g_TPost, g_TCons, g_NAlign, g_TPrep, g_TLib, g_TGaps);
	}

static void Check(int rc, const char *What)
	{
	if (rc != MB200_OK)
		Die("libmuscle_b200 %s failed (%d): %s", What, rc, mb200_last_error(g_Ctx));
	}

static void EnsureCtx()
	{
	if (g_Ctx != 0)
		return;
	int Device = 0;
	const char *s = getenv("MB200_DEVICE");
	if (s != 0)
		Device = atoi(s);
	int rc = mb200_create(Device, &g_Ctx);
	if (rc != MB200_OK)
		Die("libmuscle_b200 mb200_create failed (%d): %s", rc, mb200_last_error(0));
	if (getenv("MB200_TRACE") != 0)
		atexit(TraceReport);
	}

void MPCFlat::CalcPosteriors()
	{
	const double T0 = Now();
	EnsureCtx();
	const uint SeqCount = GetSeqCount();
	const uint PairCount = SIZE(m_Pairs);
	asserta(PairCount > 0);

// PairHMM tables are process-global statics rewritten between replicates (align.cpp:30-40):
// upload them every time, exactly as the host computed them.
	Check(mb200_set_hmm(g_Ctx, PairHMM::m_StartScore, &PairHMM::m_TransScore[0][0],
	  PairHMM::m_InsScore, &PairHMM::m_MatchScore[0][0], MIN_SPARSE_SCORE), "mb200_set_hmm");

	vector<byte> Bytes;
	vector<uint64_t> Offsets;
	Offsets.push_back(0);
	for (uint i = 0; i < SeqCount; ++i)
		{
		const byte *Seq = GetBytePtr(i);
		const uint L = GetSeqLength(i);
		Bytes.insert(Bytes.end(), Seq, Seq + L);
		Offsets.push_back(Bytes.size());
		}
	Check(mb200_set_seqs(g_Ctx, SeqCount, Bytes.data(), Offsets.data()), "mb200_set_seqs");

	ProgressStep(0, 2, "Calc posteriors (B200)");
	vector<float> EAs(PairCount);
	Check(mb200_posteriors_allpairs(g_Ctx, 0, PairCount, EAs.data()), "mb200_posteriors_allpairs");
	for (uint PairIndex = 0; PairIndex < PairCount; ++PairIndex)
		{
		const pair<uint, uint> &Pair = GetPair(PairIndex);
		const float EA = EAs[PairIndex];
		m_DistMx[Pair.first][Pair.second] = EA;			// calcposteriorflat.cpp:89-91
		m_DistMx[Pair.second][Pair.first] = EA;
		}
	ProgressStep(1, 2, "Calc posteriors (B200)");
	g_TPost += Now() - T0;
	// The sparse posteriors stay resident in HBM for ConsIter / AlignAlns (the store replaces
	// m_SparsePosts1/2).  Set MB200_EXPORT_SPARSE=1 to also fill the host MySparseMx objects, e.g.
	// for commands that read GetSparsePost() directly (-profalign).
	if (getenv("MB200_EXPORT_SPARSE") != 0)
		{
		for (uint PairIndex = 0; PairIndex < PairCount; ++PairIndex)
			{
			const pair<uint, uint> &Pair = GetPair(PairIndex);
			const uint LX = GetSeqLength(Pair.first);
			const uint LY = GetSeqLength(Pair.second);
			MySparseMx &S = GetSparsePost(PairIndex);
			S.m_LX = LX;
			S.m_LY = LY;
			S.AllocLX(LX);
			Check(mb200_export_pair(g_Ctx, PairIndex, S.m_Offsets, 0), "mb200_export_pair");
			S.m_VecSize = S.m_Offsets[LX];
			S.AllocVec(S.m_VecSize);
			Check(mb200_export_pair(g_Ctx, PairIndex, S.m_Offsets, (mb200_entry *) S.m_ValueVec), "mb200_export_pair");
			S.m_X = GetBytePtr(Pair.first);
			S.m_Y = GetBytePtr(Pair.second);
			}
		}
	}

void MPCFlat::ConsIter(uint Iter)
	{
	EnsureCtx();
	const uint PairCount = SIZE(m_Pairs);
	asserta(PairCount > 0);
	ProgressStep(0, 2, "Consistency (%u/%u) (B200)", Iter + 1, m_ConsistencyIterCount);
	const double T0 = Now();
	Check(mb200_consistency_iter(g_Ctx, 0, PairCount), "mb200_consistency_iter");
	g_TCons += Now() - T0;
	ProgressStep(1, 2, "Consistency (%u/%u) (B200)", Iter + 1, m_ConsistencyIterCount);
	// the Jacobi buffer swap (consflat.cpp:22) happens inside the library
	}

MultiSequence *MPCFlat::AlignAlns(const MultiSequence &MSA1, const MultiSequence &MSA2, float *ptrScore)
	{
	EnsureCtx();
	const double T0 = Now();
	++g_NAlign;
	const uint SeqCount1 = MSA1.GetSeqCount();
	const uint SeqCount2 = MSA2.GetSeqCount();
	const uint ColCount1 = MSA1.GetColCount();
	const uint ColCount2 = MSA2.GetColCount();

	vector<uint> Ids1, Ids2, P2C1, P2C2, PosToCol;
	for (uint i = 0; i < SeqCount1; ++i)
		{
		const Sequence *Seq = MSA1.GetSequence(i);
		Ids1.push_back(GetMyInputSeqIndex(Seq->m_Label));
		Seq->GetPosToCol(PosToCol);
		P2C1.insert(P2C1.end(), PosToCol.begin(), PosToCol.end());
		}
	for (uint i = 0; i < SeqCount2; ++i)
		{
		const Sequence *Seq = MSA2.GetSequence(i);
		Ids2.push_back(GetMyInputSeqIndex(Seq->m_Label));
		Seq->GetPosToCol(PosToCol);
		P2C2.insert(P2C2.end(), PosToCol.begin(), PosToCol.end());
		}

	vector<char> PathBuf(ColCount1 + ColCount2 + 1);
	float Score = 0;
	const double T1 = Now();
	g_TPrep += T1 - T0;
	Check(mb200_align_groups(g_Ctx, SeqCount1, Ids1.data(), P2C1.data(), ColCount1,
	  SeqCount2, Ids2.data(), P2C2.data(), ColCount2, PathBuf.data(), &Score, 0), "mb200_align_groups");
	const double T2 = Now();
	g_TLib += T2 - T1;
	if (ptrScore != 0)
		*ptrScore = Score;
	const string Path(PathBuf.data());

// gap insertion exactly as alnalnsflat.cpp:36-50
	MultiSequence *result = new MultiSequence();
	for (uint SeqIndex1 = 0; SeqIndex1 < SeqCount1; ++SeqIndex1)
		{
		const Sequence *InputRow = MSA1.GetSequence(SeqIndex1);
		Sequence *AlignedRow = InputRow->AddGapsPath(Path, 'X');
		result->AddSequence(AlignedRow, true);
		}
	for (uint SeqIndex2 = 0; SeqIndex2 < SeqCount2; ++SeqIndex2)
		{
		const Sequence *InputRow = MSA2.GetSequence(SeqIndex2);
		Sequence *AlignedRow = InputRow->AddGapsPath(Path, 'Y');
		result->AddSequence(AlignedRow, true);
		}
	g_TGaps += Now() - T2;
	return result;
	}
