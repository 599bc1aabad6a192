// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Our own glue (no reference code is copied here) that is compiled TOGETHER with the unmodified
// reference sources under /root/reference/src (see oracle/Makefile) into oracle/_ref/libmuscle_ref.so.
// It exposes the reference's hot-path functions through a C ABI so that python tests (ctypes) can
//   * validate the plain-C restatement in oracle/muscle_oracle.c bit for bit,
//   * generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * serve as the "reference" CPU baseline that bench.py times beside the GPU path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Reference entry points wrapped (file:line in /root/reference/src):
//   CalcFwdFlat fwdflat3.cpp:12, CalcBwdFlat bwdflat3.cpp:10, CalcTotalProbFlat totalprobflat.cpp:3,
//   CalcPostFlat calcposteriorflat.cpp:4, MySparseMx::FromPost mysparsemx.cpp:115,
//   CalcAlnScoreFlat calcalnscoreflat.cpp:4, CalcAlnFlat calcalnflat.cpp:6,
//   MPCFlat::{InitSeqs,InitPairs,InitDistMx,ConsIter,CalcGuideTree,CalcJoinOrder,ProgressiveAlign,
//   Refine,SortMSA,BuildPost,AlignAlns} mpcflat.cpp / consflat.cpp / progalnflat.cpp / refineflat.cpp.
#include "muscle.h"
#include "mpcflat.h"
#include "pairhmm.h"
#include "mega.h"
#include <omp.h>
#include <chrono>

void InitProbcons();
string g_Arg1;	// the one global main.cpp (not linked into the .so) defines for the other commands

namespace {
struct Handle
	{
	MPCFlat M;
	MultiSequence *MS = 0;
	uint Iter = 0;
	};

// What MPCFlat::CalcPosterior (calcposteriorflat.cpp:45-92) does for one pair, minus the label
// registry lookups (globalinputms.cpp can only be initialised once per process): the same
// reference functions in the same order.
static void OnePair(MPCFlat &M, uint PairIndex)
	{
	const pair<uint, uint> &Pair = M.GetPair(PairIndex);
	const uint ix = Pair.first, iy = Pair.second;
	const uint LX = M.GetSeqLength(ix), LY = M.GetSeqLength(iy);
	const byte *X = M.GetBytePtr(ix);
	const byte *Y = M.GetBytePtr(iy);
	float *Fwd = AllocFB(LX, LY);
	float *Bwd = AllocFB(LX, LY);
	CalcFwdFlat(X, LX, Y, LY, Fwd);
	CalcBwdFlat(X, LX, Y, LY, Bwd);
	float *Post = AllocPost(LX, LY);
	CalcPostFlat(Fwd, Bwd, LX, LY, Post);
	myfree(Fwd);
	myfree(Bwd);
	MySparseMx &SP = M.GetSparsePost(PairIndex);
	SP.FromPost(Post, LX, LY);
	SP.m_X = X;
	SP.m_Y = Y;
	float *DPRows = AllocDPRows(LX, LY);
	float Score = CalcAlnScoreFlat(Post, LX, LY, DPRows);
	myfree(Post);
	myfree(DPRows);
	float EA = Score/min(LX, LY);
	M.m_DistMx[ix][iy] = EA;
	M.m_DistMx[iy][ix] = EA;
	}
}

extern "C" {

int ref_init(int nucleo, int threads)
	{
	static bool Done = false;
	if (Done)
		return 0;
	opt_quiet = true;
	optset_quiet = true;
	if (threads > 0)
		{
		opt_threads = (unsigned) threads;
		optset_threads = true;
		}
	SetAlpha(nucleo ? ALPHA_Nucleo : ALPHA_Amino);
	InitProbcons();
	Done = true;
	return 0;
	}

int ref_threads() { return (int) GetRequestedThreadCount(); }

void ref_get_hmm(float *start5, float *trans25, float *ins256, float *match65536)
	{
	memcpy(start5, PairHMM::m_StartScore, 5*sizeof(float));
	memcpy(trans25, PairHMM::m_TransScore, 25*sizeof(float));
	memcpy(ins256, PairHMM::m_InsScore, 256*sizeof(float));
	memcpy(match65536, PairHMM::m_MatchScore, 65536*sizeof(float));
	}

void ref_set_hmm(const float *start5, const float *trans25, const float *ins256, const float *match65536)
	{
	memcpy(PairHMM::m_StartScore, start5, 5*sizeof(float));
	memcpy(PairHMM::m_TransScore, trans25, 25*sizeof(float));
	memcpy(PairHMM::m_InsScore, ins256, 256*sizeof(float));
	memcpy(PairHMM::m_MatchScore, match65536, 65536*sizeof(float));
	}

float ref_min_sparse_score() { return MIN_SPARSE_SCORE; }

void ref_fwd(const byte *X, uint LX, const byte *Y, uint LY, float *Flat) { CalcFwdFlat(X, LX, Y, LY, Flat); }
void ref_bwd(const byte *X, uint LX, const byte *Y, uint LY, float *Flat) { CalcBwdFlat(X, LX, Y, LY, Flat); }
float ref_total(const float *Fwd, const float *Bwd, uint LX, uint LY) { return CalcTotalProbFlat(Fwd, Bwd, LX, LY); }
void ref_postflat(const float *Fwd, const float *Bwd, uint LX, uint LY, float *Post) { CalcPostFlat(Fwd, Bwd, LX, LY, Post); }

void ref_calcpost(const byte *X, uint LX, const byte *Y, uint LY, float *Post)
	{
	float *Fwd = AllocFB(LX, LY);
	float *Bwd = AllocFB(LX, LY);
	CalcFwdFlat(X, LX, Y, LY, Fwd);
	CalcBwdFlat(X, LX, Y, LY, Bwd);
	CalcPostFlat(Fwd, Bwd, LX, LY, Post);
	myfree(Fwd);
	myfree(Bwd);
	}

// returns nnz; offsets has LX+1 slots, entries nnz x {float P; uint32 col}
uint ref_frompost(const float *Post, uint LX, uint LY, uint *Offsets, byte *Entries, uint Cap)
	{
	MySparseMx S;
	S.FromPost(Post, LX, LY);
	memcpy(Offsets, S.m_Offsets, (LX + 1)*sizeof(uint));
	uint n = S.m_VecSize;
	if (n <= Cap)
		memcpy(Entries, S.m_ValueVec, size_t(n)*8);
	return n;
	}

float ref_alnscore(const float *Post, uint LX, uint LY)
	{
	float *DPRows = AllocDPRows(LX, LY);
	float s = CalcAlnScoreFlat(Post, LX, LY, DPRows);
	myfree(DPRows);
	return s;
	}

// PathOut needs LX+LY+1 bytes; returns score, writes NUL-terminated path over {B,X,Y}
float ref_calcaln(const float *Post, uint LX, uint LY, char *PathOut)
	{
	float *DPRows = AllocDPRows(LX, LY);
	char *TB = AllocTB(LX, LY);
	string Path;
	float s = CalcAlnFlat(Post, LX, LY, DPRows, TB, Path);
	myfree(DPRows);
	myfree(TB);
	memcpy(PathOut, Path.c_str(), Path.size() + 1);
	return s;
	}

// ---------------------------------------------------------------- MPCFlat handle
void *ref_mpc_create(int nseq, const char *const *seqs)
	{
	Handle *H = new Handle;
	vector<string> Labels, Seqs;
	for (int i = 0; i < nseq; ++i)
		{
		char tmp[32];
		snprintf(tmp, sizeof tmp, "s%d", i);
		Labels.push_back(tmp);
		Seqs.push_back(seqs[i]);
		}
	H->MS = new MultiSequence;
	H->MS->FromStrings(Labels, Seqs);
	MPCFlat &M = H->M;
	M.Clear();
	const uint N = (uint) nseq;
	M.AllocPairCount(N*(N - 1)/2);
	M.InitSeqs(H->MS);
	M.InitPairs();
	M.InitDistMx();
	return H;
	}

void ref_mpc_destroy(void *h)
	{
	Handle *H = (Handle *) h;
	H->M.Clear();
	delete H->MS;
	delete H;
	}

uint ref_mpc_paircount(void *h) { return SIZE(((Handle *) h)->M.m_Pairs); }

// the CalcPosteriors loop (mpcflat.cpp:214-252) with the same OpenMP static schedule
double ref_mpc_posteriors(void *h, int threads)
	{
	Handle *H = (Handle *) h;
	MPCFlat &M = H->M;
	const int PairCount = (int) SIZE(M.m_Pairs);
	if (threads <= 0)
		threads = (int) GetRequestedThreadCount();
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads)
	for (int PairIndex = 0; PairIndex < PairCount; ++PairIndex)
		OnePair(M, (uint) PairIndex);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
	}

// time the posterior stage on a subset of pairs [lo,hi) (bounded CPU-baseline sample)
double ref_mpc_posteriors_range(void *h, int lo, int hi, int threads)
	{
	Handle *H = (Handle *) h;
	MPCFlat &M = H->M;
	if (threads <= 0)
		threads = (int) GetRequestedThreadCount();
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
	for (int PairIndex = lo; PairIndex < hi; ++PairIndex)
		OnePair(M, (uint) PairIndex);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
	}

uint ref_mpc_pair_nnz(void *h, uint PairIndex)
	{
	// m_Offsets[LX], not m_VecSize: UpdateFromPost (mysparsemx.cpp:87-113) never sets m_VecSize
	MySparseMx &S = ((Handle *) h)->M.GetSparsePost(PairIndex);
	return S.m_Offsets[S.m_LX];
	}

void ref_mpc_export(void *h, uint PairIndex, uint *Offsets, byte *Entries)
	{
	MySparseMx &S = ((Handle *) h)->M.GetSparsePost(PairIndex);
	memcpy(Offsets, S.m_Offsets, (S.m_LX + 1)*sizeof(uint));
	memcpy(Entries, S.m_ValueVec, size_t(S.m_Offsets[S.m_LX])*8);
	}

// inject a sparse posterior computed elsewhere (e.g. by the GPU engine) into the reference state
void ref_mpc_import(void *h, uint PairIndex, const uint *Offsets, const byte *Entries)
	{
	MPCFlat &M = ((Handle *) h)->M;
	const pair<uint, uint> &Pair = M.GetPair(PairIndex);
	const uint LX = M.GetSeqLength(Pair.first), LY = M.GetSeqLength(Pair.second);
	MySparseMx &S = M.GetSparsePost(PairIndex);
	S.m_LX = LX;
	S.m_LY = LY;
	S.AllocLX(LX);
	memcpy(S.m_Offsets, Offsets, (LX + 1)*sizeof(uint));
	S.m_VecSize = Offsets[LX];
	S.AllocVec(S.m_VecSize);
	memcpy(S.m_ValueVec, Entries, size_t(S.m_VecSize)*8);
	S.m_X = M.GetBytePtr(Pair.first);
	S.m_Y = M.GetBytePtr(Pair.second);
	}

void ref_mpc_get_distmx(void *h, float *Out)
	{
	MPCFlat &M = ((Handle *) h)->M;
	const uint N = M.GetSeqCount();
	for (uint i = 0; i < N; ++i)
		for (uint j = 0; j < N; ++j)
			Out[i*N + j] = M.m_DistMx[i][j];
	}

void ref_mpc_set_distmx(void *h, const float *In)
	{
	MPCFlat &M = ((Handle *) h)->M;
	const uint N = M.GetSeqCount();
	for (uint i = 0; i < N; ++i)
		for (uint j = 0; j < N; ++j)
			M.m_DistMx[i][j] = In[i*N + j];
	}

// one Jacobi consistency iteration (consflat.cpp:5-23), returns seconds
double ref_mpc_consiter(void *h)
	{
	Handle *H = (Handle *) h;
	auto t0 = std::chrono::steady_clock::now();
	H->M.ConsIter(H->Iter++);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
	}

// time ConsPair on a bounded range of pairs (no swap), for the relax CPU baseline
double ref_mpc_conspairs_range(void *h, int lo, int hi, int threads)
	{
	Handle *H = (Handle *) h;
	MPCFlat &M = H->M;
	if (threads <= 0)
		threads = (int) GetRequestedThreadCount();
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
	for (int PairIndex = lo; PairIndex < hi; ++PairIndex)
		M.ConsPair((uint) PairIndex);
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
	}

// the remainder of MPCFlat::Run (mpcflat.cpp:313-333) after CalcPosteriors, with `consiters`
// consistency and `refineiters` refinement iterations. Rows are returned in final MSA order:
// labels_out[i] = input index of row i, rows_out = concatenated NUL-terminated aligned rows.
// Returns column count.
uint ref_mpc_finish(void *h, int consiters, int refineiters, int *RowSeqIndex, char *RowsOut, uint RowsCap)
	{
	Handle *H = (Handle *) h;
	MPCFlat &M = H->M;
	const uint N = M.GetSeqCount();
	M.m_ConsistencyIterCount = (uint) consiters;
	M.m_RefineIterCount = (uint) refineiters;
	M.CalcGuideTree();
	M.m_Weights.assign(N, 1.0f);		// mpcflat.cpp:324 forces all weights to 1
	M.Consistency();
	M.CalcJoinOrder();
	M.ProgressiveAlign();
	srand(1);						// refineflat.cpp:14 uses un-seeded rand() == srand(1)
	M.Refine();
	M.SortMSA();
	const MultiSequence &A = *M.m_MSA;
	const uint Cols = A.GetColCount();
	asserta(size_t(N)*(Cols + 1) <= RowsCap);
	for (uint i = 0; i < N; ++i)
		{
		const Sequence *S = A.GetSequence(i);
		RowSeqIndex[i] = atoi(S->GetLabel().c_str() + 1);
		memcpy(RowsOut + size_t(i)*(Cols + 1), S->GetCharPtr(), Cols);
		RowsOut[size_t(i)*(Cols + 1) + Cols] = 0;
		}
	return Cols;
	}

// BuildPost + CalcAlnFlat for two groups of single (ungapped) sequences is a degenerate join; the
// general case needs gapped MSAs, so expose AlignAlns on two gapped groups given as strings.
// rows1/rows2: gapped rows ('-' gaps) whose labels are "s<idx>".  Writes dense Post (Cols1*Cols2),
// returns score and path.
float ref_mpc_alignalns(void *h, int n1, const int *idx1, const char *const *rows1,
  int n2, const int *idx2, const char *const *rows2, float *PostOut, char *PathOut)
	{
	MPCFlat &M = ((Handle *) h)->M;
	const uint N = M.GetSeqCount();
	if (SIZE(M.m_Weights) != N)
		M.m_Weights.assign(N, 1.0f);
	vector<string> L1, S1, L2, S2;
	char tmp[32];
	for (int i = 0; i < n1; ++i) { snprintf(tmp, sizeof tmp, "s%d", idx1[i]); L1.push_back(tmp); S1.push_back(rows1[i]); }
	for (int i = 0; i < n2; ++i) { snprintf(tmp, sizeof tmp, "s%d", idx2[i]); L2.push_back(tmp); S2.push_back(rows2[i]); }
	MultiSequence A, B;
	A.FromStrings(L1, S1);
	B.FromStrings(L2, S2);
	const uint C1 = A.GetColCount(), C2 = B.GetColCount();
	float *Post = AllocPost(C1, C2);
	M.BuildPost(A, B, Post);
	if (PostOut != 0)
		memcpy(PostOut, Post, size_t(C1)*C2*sizeof(float));
	float *DPRows = AllocDPRows(C1, C2);
	char *TB = AllocTB(C1, C2);
	string Path;
	float Score = CalcAlnFlat(Post, C1, C2, DPRows, TB, Path);
	myfree(Post);
	myfree(DPRows);
	myfree(TB);
	memcpy(PathOut, Path.c_str(), Path.size() + 1);
	return Score;
	}

// MPCFlat::CalcGuideTree's numeric part (mpcflat.cpp:191-194): UPGMA5::Init + FixEADistMx + Run,
// children / branch lengths read back from the Tree it creates.  ea: N(N-1)/2, row-major i<j.
int ref_upgma(uint n, const float *ea, int linkage, uint *Left, uint *Right, float *LLen, float *RLen)
	{
	vector<string> Labels;
	vector<vector<float> > DistMx(n, vector<float>(n, 0));
	uint p = 0;
	for (uint i = 0; i < n; ++i)
		{
		char tmp[32];
		snprintf(tmp, sizeof tmp, "s%u", i);
		Labels.push_back(tmp);
		for (uint j = i + 1; j < n; ++j, ++p)
			{
			DistMx[i][j] = ea[p];
			DistMx[j][i] = ea[p];
			}
		}
	UPGMA5 U;
	U.Init(Labels, DistMx);
	U.FixEADistMx();
	Tree T;
	U.Run((LINKAGE) linkage, T);
	for (uint k = 0; k + 1 < n; ++k)
		{
		const uint Node = n + k;
		Left[k] = T.GetLeft(Node);
		Right[k] = T.GetRight(Node);
		LLen[k] = (float) T.GetEdgeLength(Node, Left[k]);
		RLen[k] = (float) T.GetEdgeLength(Node, Right[k]);
		}
	return 0;
	}

// ---------------------------------------------------------------- Mega (Muscle-3D feature profiles)
// Mega::FromFile (mega.cpp:113), the model it derives, and CalcPost's mega branch (calcpost.cpp:14-22:
// Mega::CalcFwdFlat_mega + CalcBwdFlat_mega, then CalcPostFlat).  One file per process (the reference
// keeps the model in statics and asserts they are empty).
int ref_mega_load(const char *Path)
	{
	Mega::FromFile(string(Path));
	return (int) Mega::GetProfileCount();
	}

uint ref_mega_nfeat() { return Mega::GetFeatureCount(); }

void ref_mega_model(uint *Alpha, float *Weights, float *LogProbs, float *LogProbMx)
	{
	const uint F = Mega::GetFeatureCount();
	uint a = 0, b = 0;
	for (uint f = 0; f < F; ++f)
		{
		const uint A = Mega::GetAlphaSize(f);
		Alpha[f] = A;
		Weights[f] = Mega::GetWeight(f);
		for (uint x = 0; x < A; ++x)
			{
			LogProbs[a++] = Mega::m_LogProbsVec[f][x];
			for (uint y = 0; y < A; ++y)
				LogProbMx[b++] = Mega::m_LogProbMxVec[f][x][y];
			}
		}
	}

uint ref_mega_profile_len(uint Idx) { return SIZE(Mega::GetProfile(Idx)); }

void ref_mega_profile(uint Idx, byte *Letters, char *Seq)
	{
	const vector<vector<byte> > &P = Mega::GetProfile(Idx);
	const uint F = Mega::GetFeatureCount();
	for (uint i = 0; i < SIZE(P); ++i)
		for (uint f = 0; f < F; ++f)
			Letters[i*F + f] = P[i][f];
	const string &S = Mega::m_Seqs[Idx];
	memcpy(Seq, S.c_str(), S.size() + 1);
	}

void ref_mega_calcpost(uint IdxX, uint IdxY, float *PostOut)
	{
	const vector<vector<byte> > &PX = Mega::GetProfile(IdxX);
	const vector<vector<byte> > &PY = Mega::GetProfile(IdxY);
	const uint LX = SIZE(PX), LY = SIZE(PY);
	float *Fwd = AllocFB(LX, LY);
	float *Bwd = AllocFB(LX, LY);
	Mega::CalcFwdFlat_mega(PX, PY, Fwd);
	Mega::CalcBwdFlat_mega(PX, PY, Bwd);
	float *Post = AllocPost(LX, LY);
	CalcPostFlat(Fwd, Bwd, LX, LY, Post);
	memcpy(PostOut, Post, size_t(LX)*LY*sizeof(float));
	myfree(Fwd);
	myfree(Bwd);
	myfree(Post);
	}

} // extern "C"
