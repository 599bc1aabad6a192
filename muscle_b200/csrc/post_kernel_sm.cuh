// post_kernel_sm.cuh -- k_posterior_sm: the fused Forward -> total -> Backward+posterior -> sparsify
// -> EA kernel.
//
// Replaces, per sequence pair, CalcFwdFlat (fwdflat3.cpp:12-153), CalcTotalProbFlat
// (totalprobflat.cpp:3-16), CalcBwdFlat (bwdflat3.cpp:10-184), CalcPostFlat
// (calcposteriorflat.cpp:4-27), MySparseMx::FromPost (mysparsemx.cpp:115-152) and CalcAlnScoreFlat
// (calcalnscoreflat.cpp:4-32) of the reference.
//
// Mapping.  One warp owns one pair at a time (persistent warps pull pairs, longest first, from an
// atomic cursor).  Lane l owns C consecutive DP columns of Y (C = ceil(LY/32), a runtime value); a
// warp therefore covers a strip of 32*C columns and longer Y are processed strip after strip with
// the strip-edge column handed over through a small per-warp buffer.  Rows of X are swept in
// anti-diagonal (wavefront) order: at step t lane l works on row t-l, the cell to the left lives in
// lane l-1 and was produced one step earlier, so the recurrence dependency is a few warp shuffles
// per step.  The per-column wavefront state lives in SHARED MEMORY as [c][lane] arrays (bank ==
// lane, conflict free) and the loop over the lane's columns is rolled: the hot loops are short and
// stay in the instruction cache (the round-1 register-resident, fully unrolled variant spent 2.6
// stall cycles per issued instruction on instruction fetch, profiles/r01_SUMMARY.md).  Only the
// Forward M-state is spilled (4 B/cell), in [step][c][lane] order so that both the Forward write
// and the Backward read are full 128-byte lines.  Backward walks the same anti-diagonals in
// reverse, fuses the Fwd (*) Bwd posterior, thresholds at log(0.01) and appends survivors to a
// per-row candidate list; the warp then compacts the rows into MySparseMx order and runs the
// expected-accuracy max-sum DP row by row as a warp-wide prefix-max (new[j] = max_{k<=j}
// max(old[k], old[k-1]+P[k]) is exactly the reference's max3 recurrence because every value is an
// exact max of the same sums).
//
// Border handling without special code paths: LOG_ZERO is absorbing under the branch-free LogAdd,
// so the reference's border formulas (first row/column of Forward, fwdflat3.cpp:35-93; last
// row/column of Backward, bwdflat3.cpp:132-176) fall out of the general cell update when the
// out-of-range neighbours are LOG_ZERO.  The only injected values are the start scores at Forward
// (0,0) and the end scores at Backward (LX,LY) (bwdflat3.cpp:53-61).
//
// Emissions.  Plain mode: residue classes index a K x K match table and a K-entry insert table
// (PairHMM::m_MatchScore / m_InsScore, pairhmm.h:26-29).  MEGA mode (Muscle-3D feature profiles,
// Mega::CalcFwdFlat_mega / CalcBwdFlat_mega, fwdflat_mega.cpp:14, bwdflat_mega.cpp): a position is a
// vector of <= 8 feature letters; the match emission of a cell is the sum over features of
// LogProbMx_f[a_f][b_f]*w_f accumulated from 0 in feature order (Mega::GetMatchScore,
// mega.cpp:340-359) -- the products are formed once on the host (same fp32 multiply), the kernel adds
// them in order -- and the insert emission is a per-position value precomputed the same way
// (Mega::GetInsScore, mega.cpp:273-286).  Same recurrences, same kernel, one template flag.
#pragma once
#include "common.cuh"

#define MB_SM_ARRAYS 6            // S/old, M, IX, JX, ycode, eY  (+1 in MEGA mode: second word of feature letters)
#ifndef MB_SM_BLOCKS
#define MB_SM_BLOCKS 5            // resident CTAs/SM the register budget is tuned for (measured best of 4/5/6/8)
#endif
#ifndef MB_SM_UNROLL
#define MB_SM_UNROLL 1            // measured on C3: unroll 1: 51.7, 2: 51.2, 4: 37.7 Gcells/s (I-cache)
#endif
#define MB_PRAGMA(x) _Pragma(#x)
#define MB_UNROLL(n) MB_PRAGMA(unroll n)

struct PostSmemHdr
	{
	float4 coef[17];      // LOGEXP1 pieces indexed by ceil(2*gap), see LogAdd
	float  insT[MB_MAX_K];
	};

// sum over features of the pre-multiplied pair tables, in feature order starting from 0
// (mega.cpp:350-357); xb[f] = offset of row a_f of feature f's table, y0/y1 = packed letters of the column
__device__ __forceinline__ float mb_mega_emit(const float *T, const uint32_t (&xb)[8], uint32_t y0, uint32_t y1, int nf)
	{
	float e = ADD(0.0f, T[xb[0] + (y0 & 255u)]);
	if (nf > 1) e = ADD(e, T[xb[1] + ((y0 >> 8) & 255u)]);
	if (nf > 2) e = ADD(e, T[xb[2] + ((y0 >> 16) & 255u)]);
	if (nf > 3) e = ADD(e, T[xb[3] + (y0 >> 24)]);
	if (nf > 4) e = ADD(e, T[xb[4] + (y1 & 255u)]);
	if (nf > 5) e = ADD(e, T[xb[5] + ((y1 >> 8) & 255u)]);
	if (nf > 6) e = ADD(e, T[xb[6] + ((y1 >> 16) & 255u)]);
	if (nf > 7) e = ADD(e, T[xb[7] + (y1 >> 24)]);
	return e;
	}

template <bool MEGA>
__global__ void __launch_bounds__(32*MB_WARPS_PER_BLOCK, MEGA ? 4 : MB_SM_BLOCKS)
k_posterior_sm(const PostParams P)
	{
	constexpr int NARR = MEGA ? MB_SM_ARRAYS + 1 : MB_SM_ARRAYS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	PostSmemHdr &sm = *reinterpret_cast<PostSmemHdr *>(smem_raw);
	const MbHmm h = P.h;
	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	const int gwarp = blockIdx.x*MB_WARPS_PER_BLOCK + wib;
	const int CM = (int) P.cmax;                             // columns per lane the smem is sized for
	float *matchT = reinterpret_cast<float *>(smem_raw + sizeof(PostSmemHdr));
	const int tsize = MEGA ? (int) P.mega_tsize : h.K*h.KS;
	float *wbase = matchT + ((tsize + 3) & ~3) + (size_t) wib*NARR*CM*32;
	float *aS = wbase;                  // Forward: S ; EA: old row
	float *aM = aS + CM*32;             // M        ; EA: staged posterior row (column-linear)
	float *aIX = aM + CM*32;            //          ; EA: lane-local prefix maxima
	float *aJX = aIX + CM*32;
	int   *aY = reinterpret_cast<int *>(aJX + CM*32);
	float *aE = reinterpret_cast<float *>(aY) + CM*32;
	int   *aY2 = reinterpret_cast<int *>(aE + CM*32);       // MEGA only: letters of features 4..7
	(void) aY2;

	mb_logadd_fill(sm.coef);
	if (!MEGA)
		for (int k = threadIdx.x; k < h.K; k += blockDim.x)
			sm.insT[k] = P.insT[k];
	for (int k = threadIdx.x; k < tsize; k += blockDim.x)
		matchT[k] = P.matchT[k];
	const int nf = (int) P.mega_nf;
	__syncthreads();

	const LogAdd la = mb_logadd_make(sm.coef);
	const float Z = MB_LOG_ZERO;

	float       *fm     = P.fm + (size_t) gwarp*P.fm_stride;
	float4      *edge0  = P.edge + (size_t) gwarp*P.edge_stride;
	float4      *edge1  = edge0 + (P.lxmax + 2);
	mb200_entry *rows   = P.rows + (size_t) gwarp*P.rows_stride;
	uint8_t     *rowcnt = P.rowcnt + (size_t) gwarp*P.rowcnt_stride;

	for (;;)
		{
		uint32_t w = 0;
		if (lane == 0)
			w = atomicAdd(P.counter, 1u);
		w = __shfl_sync(MB_FULL, w, 0);
		if (w >= P.nwork)
			break;
		const uint32_t pair = P.order[w];
		const uint32_t sx = P.px[pair], sy = P.py[pair];
		const int LX = (int) P.seqlen[sx], LY = (int) P.seqlen[sy];
		const uint8_t *Xc = P.codes + P.seqoff[sx];
		const uint8_t *Yc = P.codes + P.seqoff[sy];
		// MEGA: 8 letter bytes and one insert emission per position
		const uint2 *Xl = reinterpret_cast<const uint2 *>(P.codes) + P.seqoff[sx];
		const uint2 *Yl = reinterpret_cast<const uint2 *>(P.codes) + P.seqoff[sy];
		const float *Xi = P.insT + P.seqoff[sx];
		const float *Yi = P.insT + P.seqoff[sy];
		(void) Xc; (void) Yc; (void) Xl; (void) Yl; (void) Xi; (void) Yi;
		const int C = min(CM, (LY + 31) >> 5);            // columns per lane for this pair
		const int W = 32*C;
		const int nstrips = (LY + W - 1)/W;
		const int lastj0 = (nstrips - 1)*W;
		const int lcl = (LY - 1 - lastj0)/C;              // lane / local column of DP column LY
		const int clast = (LY - 1 - lastj0) - lcl*C;

		for (int k = lane; k < LX; k += 32)
			rowcnt[k] = 0;

		// ============================ Forward ============================
		float fin0 = Z, fin1 = Z, fin2 = Z, fin3 = Z, fin4 = Z;
		for (int strip = 0; strip < nstrips; ++strip)
			{
			const int j0 = strip*W;
			const int ncol = min(W, LY - j0);
			const int nl = (ncol + C - 1)/C;
			const bool last = (strip == nstrips - 1);
			const float4 *edgeIn = (strip & 1) ? edge0 : edge1;
			float4 *edgeOut = (strip & 1) ? edge1 : edge0;
			float *fms = fm + (size_t) strip*P.fm_rows*W;

			for (int c = 0; c < C; ++c)
				{
				const int jj = j0 + lane*C + c;
				if (MEGA)
					{
					const uint2 yl = jj < LY ? Yl[jj] : make_uint2(0u, 0u);
					aY[c*32 + lane] = (int) yl.x;
					aY2[c*32 + lane] = (int) yl.y;
					aE[c*32 + lane] = jj < LY ? Yi[jj] : 0.0f;
					}
				else
					{
					const int code = jj < LY ? (int) Yc[jj] : h.pad;
					aY[c*32 + lane] = code;
					aE[c*32 + lane] = sm.insT[code];
					}
				aS[c*32 + lane] = Z; aM[c*32 + lane] = Z; aIX[c*32 + lane] = Z; aJX[c*32 + lane] = Z;
				}
			float outM = Z, outAIY = Z, outAJY = Z, outS = Z;
			float dprev = Z;
			float bIX = Z, bJX = Z;
			const int nsteps = LX + nl;
			int xcPref = h.pad;
			uint2 xlPref = make_uint2(0u, 0u);
			float exPref = 0.0f;
			for (int t = 0; t < nsteps; ++t)
				{
				const int i = t - lane;
				const int xc = xcPref;
				const uint2 xl = xlPref;
				const float exm = exPref;
				if (MEGA)
					{
					const bool in = i >= 0 && i < LX;
					xlPref = in ? Xl[i] : make_uint2(0u, 0u);
					exPref = in ? Xi[i] : 0.0f;
					}
				else
					xcPref = (i >= 0 && i < LX) ? (int) Xc[i] : h.pad;
				(void) xc; (void) xl; (void) exm;
				float Lm = __shfl_up_sync(MB_FULL, outM, 1);
				float Laiy = __shfl_up_sync(MB_FULL, outAIY, 1);
				float Lajy = __shfl_up_sync(MB_FULL, outAJY, 1);
				float Ls = __shfl_up_sync(MB_FULL, outS, 1);
				if (i >= 0 && i <= LX && lane < nl)
					{
					const float ex = MEGA ? exm : sm.insT[xc];
					if (lane == 0)
						{
						if (strip == 0)
							{
							Lm = Z;
							if (i == 0)
								{
								Laiy = h.tSI; Lajy = h.tSJ; Ls = h.tSM;
								}
							else
								{
								if (i == 1)
									{
									bIX = ADD(h.tSI, ex); bJX = ADD(h.tSJ, ex);
									}
								else
									{
									bIX = ADD(ADD(bIX, h.tII), ex); bJX = ADD(ADD(bJX, h.tJJ), ex);
									}
								Laiy = Z; Lajy = Z;
								Ls = la(ADD(bIX, h.tIM), ADD(bJX, h.tJM));
								}
							}
						else
							{
							const float4 e = edgeIn[i];
							Lm = e.x; Laiy = e.y; Lajy = e.z; Ls = e.w;
							}
						}
					float diag = dprev;
					dprev = Ls;
					float lm = Lm, laiy = Laiy, lajy = Lajy, sLast = Z;
					const float *mrow = matchT + (MEGA ? 0 : xc*h.KS);
					uint32_t xb[8];
					if (MEGA)
						{
#pragma unroll
						for (int f = 0; f < 8; ++f)
							xb[f] = P.mega_base[f] + ((f < 4 ? (xl.x >> (8*f)) : (xl.y >> (8*(f - 4)))) & 255u)*P.mega_alpha[f];
						}
					const bool cap = last && i == LX && lane == lcl;
					float *dst = fms + (size_t) t*W + lane;
					const bool dump = P.dbg_fwd != nullptr && i >= 1;
MB_UNROLL(MB_SM_UNROLL)
					for (int c = 0; c < C; ++c)
						{
						const int o = c*32 + lane;
						const float Mo = aM[o], IXo = aIX[o], JXo = aJX[o];
						const float eyc = aE[o];
						const float em = MEGA ? mb_mega_emit(matchT, xb, (uint32_t) aY[o], (uint32_t) aY2[o], nf) : mrow[aY[o]];
						const float m = ADD(diag, em);
						const float ix = ADD(la(ADD(IXo, h.tII), ADD(Mo, h.tMI)), ex);
						const float jx = ADD(la(ADD(JXo, h.tJJ), ADD(Mo, h.tMJ)), ex);
						const float iy = ADD(la(laiy, ADD(lm, h.tMI)), eyc);
						const float jy = ADD(la(lajy, ADD(lm, h.tMJ)), eyc);
						const float s = la(ADD(m, h.tMM), la(ADD(ix, h.tIM), la(ADD(jx, h.tJM),
						  la(ADD(iy, h.tIM), ADD(jy, h.tJM)))));
						diag = aS[o];
						aS[o] = s; aM[o] = m; aIX[o] = ix; aJX[o] = jx;
						lm = m;
						laiy = ADD(iy, h.tII);
						lajy = ADD(jy, h.tJJ);
						sLast = s;
						dst[c*32] = m;
						if (cap && c == clast)
							{
							fin0 = m; fin1 = ix; fin2 = iy; fin3 = jx; fin4 = jy;
							}
						if (dump)
							{
							const int col = j0 + lane*C + c;
							if (col < LY)
								P.dbg_fwd[(size_t)(i - 1)*LY + col] = m;
							}
						}
					outM = lm; outAIY = laiy; outAJY = lajy; outS = sLast;
					if (lane == 31 && !last)
						edgeOut[i] = make_float4(outM, outAIY, outAJY, outS);
					}
				}
			__syncwarp();
			}

		// ============================ total probability ============================
		fin0 = __shfl_sync(MB_FULL, fin0, lcl);
		fin1 = __shfl_sync(MB_FULL, fin1, lcl);
		fin2 = __shfl_sync(MB_FULL, fin2, lcl);
		fin3 = __shfl_sync(MB_FULL, fin3, lcl);
		fin4 = __shfl_sync(MB_FULL, fin4, lcl);
		float total = Z;
		total = la(total, ADD(fin0, h.tSM));
		total = la(total, ADD(fin1, h.tSI));
		total = la(total, ADD(fin2, h.tSI));
		total = la(total, ADD(fin3, h.tSJ));
		total = la(total, ADD(fin4, h.tSJ));
		if (P.dbg_total != nullptr && lane == 0)
			*P.dbg_total = total;

		// ============================ Backward + posterior ============================
		uint32_t kept = 0;
		bool overflow = false;
		for (int strip = nstrips - 1; strip >= 0; --strip)
			{
			const int j0 = strip*W;
			const int ncol = min(W, LY - j0);
			const int nl = (ncol + C - 1)/C;
			const bool last = (strip == nstrips - 1);
			const float4 *edgeIn = (strip & 1) ? edge0 : edge1;
			float4 *edgeOut = (strip & 1) ? edge1 : edge0;
			const float *fms = fm + (size_t) strip*P.fm_rows*W;

			for (int c = 0; c < C; ++c)
				{
				const int jj = j0 + lane*C + c + 1;
				if (MEGA)
					{
					const uint2 yl = jj < LY ? Yl[jj] : make_uint2(0u, 0u);
					aY[c*32 + lane] = (int) yl.x;
					aY2[c*32 + lane] = (int) yl.y;
					aE[c*32 + lane] = jj < LY ? Yi[jj] : 0.0f;
					}
				else
					{
					const int code = jj < LY ? (int) Yc[jj] : h.pad;
					aY[c*32 + lane] = code;
					aE[c*32 + lane] = sm.insT[code];
					}
				aM[c*32 + lane] = Z; aIX[c*32 + lane] = Z; aJX[c*32 + lane] = Z;
				}
			float outM = Z, outIY = Z, outJY = Z;
			float dprev = Z;
			const int nsteps = LX + nl - 1;
			int xcPref = h.pad;
			uint2 xlPref = make_uint2(0u, 0u);
			float exPref = 0.0f;
			for (int u = 0; u < nsteps; ++u)
				{
				const int i = LX - u + (nl - 1 - lane);
				const int xc = xcPref;
				const uint2 xl = xlPref;
				const float exm = exPref;
				if (MEGA)
					{
					const bool in = i >= 2 && i <= LX;
					xlPref = in ? Xl[i - 1] : make_uint2(0u, 0u);
					exPref = in ? Xi[i - 1] : 0.0f;
					}
				else
					xcPref = (i >= 2 && i <= LX) ? (int) Xc[i - 1] : h.pad;
				(void) xc; (void) xl; (void) exm;
				float Rm = __shfl_down_sync(MB_FULL, outM, 1);
				float Riy = __shfl_down_sync(MB_FULL, outIY, 1);
				float Rjy = __shfl_down_sync(MB_FULL, outJY, 1);
				if (i >= 1 && i <= LX && lane < nl)
					{
					if (lane == nl - 1)
						{
						if (last)
							{
							Rm = Z; Riy = Z; Rjy = Z;
							}
						else
							{
							const float4 e = edgeIn[i];
							Rm = e.x; Riy = e.y; Rjy = e.z;
							}
						}
					const float ex = MEGA ? exm : sm.insT[xc];
					const float *mrow = matchT + (MEGA ? 0 : xc*h.KS);
					uint32_t xb[8];
					if (MEGA)
						{
#pragma unroll
						for (int f = 0; f < 8; ++f)
							xb[f] = P.mega_base[f] + ((f < 4 ? (xl.x >> (8*f)) : (xl.y >> (8*(f - 4)))) & 255u)*P.mega_alpha[f];
						}
					uint32_t cnt = rowcnt[i - 1];
					const uint32_t cnt0 = cnt;
					mb200_entry *row = rows + (size_t)(i - 1)*MB_CAP;
					const int t = i + lane;
					const float *src = fms + (size_t) t*W + lane;
					float mdiag = dprev;
					dprev = Rm;
					float riy = Riy, rjy = Rjy, mFirst = Z;
					const bool inj = last && i == LX && lane == lcl;
					const bool dump = P.dbg_bwd != nullptr;
					float fmNext = src[(C - 1)*32];
MB_UNROLL(MB_SM_UNROLL)
					for (int c = C - 1; c >= 0; --c)
						{
						const int o = c*32 + lane;
						const float fmv = fmNext;
						if (c > 0)
							fmNext = src[(c - 1)*32];
						const float Mo = aM[o], IXo = aIX[o], JXo = aJX[o];
						const float eyc = aE[o];
						const float em = MEGA ? mb_mega_emit(matchT, xb, (uint32_t) aY[o], (uint32_t) aY2[o], nf) : mrow[aY[o]];
						const float nM = ADD(mdiag, em);
						const float nIX = ADD(IXo, ex);
						const float nJX = ADD(JXo, ex);
						const float nIY = ADD(riy, eyc);
						const float nJY = ADD(rjy, eyc);
						float m = la(ADD(h.tMM, nM), la(ADD(h.tMI, nIX), la(ADD(h.tMJ, nJX),
						  la(ADD(h.tMI, nIY), ADD(h.tMJ, nJY)))));
						float ix = la(ADD(h.tII, nIX), ADD(h.tIM, nM));
						float jx = la(ADD(h.tJJ, nJX), ADD(h.tJM, nM));
						float iy = la(ADD(h.tII, nIY), ADD(h.tIM, nM));
						float jy = la(ADD(h.tJJ, nJY), ADD(h.tJM, nM));
						if (inj && c == clast)
							{
							m = h.tSM; ix = h.tSI; jx = h.tSJ; iy = h.tSI; jy = h.tSJ;
							}
						mdiag = Mo;
						aM[o] = m; aIX[o] = ix; aJX[o] = jx;
						riy = iy; rjy = jy;
						mFirst = m;
						// posterior (calcposteriorflat.cpp:14-22); candidates arrive with descending column
						const int col = j0 + lane*C + c;
						const float score = __fsub_rn(ADD(fmv, m), total);
						if (col < LY && score >= h.minScore)
							{
							const float p = score >= 0.0f ? 1.0f : mb_expf_glibc(score);
							if (cnt < MB_CAP)
								{
								row[cnt].p = p;
								row[cnt].col = (uint32_t) col;
								}
							else
								overflow = true;
							++cnt;
							kept += (p >= 0.01f) ? 1u : 0u;
							}
						if (dump && col < LY)
							{
							P.dbg_bwd[(size_t)(i - 1)*LY + col] = m;
							P.dbg_post[(size_t)(i - 1)*LY + col] =
							  score < h.minScore ? 0.0f : (score >= 0.0f ? 1.0f : mb_expf_glibc(score));
							}
						}
					outM = mFirst; outIY = riy; outJY = rjy;
					if (lane == 0 && strip > 0)
						edgeOut[i] = make_float4(outM, outIY, outJY, 0.0f);
					if (cnt != cnt0)
						rowcnt[i - 1] = (uint8_t) min(cnt, (uint32_t) MB_CAP);
					}
				__syncwarp();
				}
			}

		// ============================ compaction + expected accuracy ============================
		overflow = __any_sync(MB_FULL, overflow);
		for (int o = 16; o > 0; o >>= 1)
			kept += __shfl_xor_sync(MB_FULL, kept, o);
		unsigned long long base = 0;
		if (lane == 0)
			base = atomicAdd(P.ent_cursor, (unsigned long long) kept);
		base = __shfl_sync(MB_FULL, base, 0);
		const bool fits = (base + kept <= P.ent_cap);
		if (lane == 0)
			{
			if (overflow)
				atomicCAS(P.err, 0, MB200_EOVERFLOW);
			if (!fits)
				atomicCAS(P.err, 0, MB200_ENOMEM);
			P.entbase[pair] = base;
			P.nnz[pair] = kept;
			}
		uint32_t *rowoff = P.rowoff + P.rowbase[pair];
		mb200_entry *out = P.entries + base;

		float eaScore = 0.0f;
		{
		float *old = aS;            // [c][lane]
		float *rowbuf = aM;         // column-linear: [lane*C + c]
		float *pre = aIX;           // [c][lane]
		float *eIn = reinterpret_cast<float *>(edge0);
		float *eOut = reinterpret_cast<float *>(edge1);
		uint32_t written = 0;
		for (int strip = 0; strip < nstrips; ++strip)
			{
			const int j0 = strip*W;
			for (int c = 0; c < C; ++c)
				{
				old[c*32 + lane] = 0.0f;
				rowbuf[c*32 + lane] = 0.0f;
				}
			__syncwarp();
			for (int i = 1; i <= LX; ++i)
				{
				const uint32_t cnt = rowcnt[i - 1];
				const mb200_entry *row = rows + (size_t)(i - 1)*MB_CAP;
				if (strip == 0 && lane == 0)
					rowoff[i - 1] = written;
				for (uint32_t e0 = 0; e0 < cnt; e0 += 32)
					{
					const int idx = (int) cnt - 1 - (int)(e0 + lane);
					mb200_entry ent; ent.p = 0.0f; ent.col = 0;
					if (idx >= 0)
						ent = row[idx];
					const int lc = (int) ent.col - j0;
					if (idx >= 0 && lc >= 0 && lc < W)
						rowbuf[lc] = ent.p;
					if (strip == 0)
						{
						const bool keep = idx >= 0 && ent.p >= 0.01f;          // mysparsemx.cpp:139-141
						const uint32_t b = __ballot_sync(MB_FULL, keep);
						if (keep && fits)
							out[written + __popc(b & ((1u << lane) - 1u))] = ent;
						written += __popc(b);
						}
					}
				__syncwarp();
				// calcalnscoreflat.cpp:13-29 as a prefix-max over the row
				float leftOld = 0.0f, incoming = 0.0f;
				if (lane == 0)
					{
					leftOld = strip == 0 ? 0.0f : eIn[i - 1];
					incoming = strip == 0 ? 0.0f : eIn[i];
					}
				else
					leftOld = old[(C - 1)*32 + lane - 1];
				float run = 0.0f, prevOld = leftOld;
				for (int c = 0; c < C; ++c)
					{
					const float o = old[c*32 + lane];
					const float b = ADD(prevOld, rowbuf[lane*C + c]);
					float v = fmaxf(o, b);
					v = c == 0 ? v : fmaxf(v, run);
					run = v;
					pre[c*32 + lane] = v;
					prevOld = o;
					}
				__syncwarp();
				float scan = run;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
					{
					const float n = __shfl_up_sync(MB_FULL, scan, o);
					if (lane >= o)
						scan = fmaxf(scan, n);
					}
				float excl = __shfl_up_sync(MB_FULL, scan, 1);
				const float inc0 = __shfl_sync(MB_FULL, incoming, 0);
				excl = lane == 0 ? inc0 : fmaxf(excl, inc0);
				float lastNew = 0.0f;
				for (int c = 0; c < C; ++c)
					{
					lastNew = fmaxf(pre[c*32 + lane], excl);
					old[c*32 + lane] = lastNew;
					}
				if (lane == 31 && strip + 1 < nstrips)
					eOut[i] = lastNew;
				__syncwarp();
				for (uint32_t e0 = 0; e0 < cnt; e0 += 32)
					{
					const int idx = (int) cnt - 1 - (int)(e0 + lane);
					if (idx >= 0)
						{
						const int lc = (int) row[idx].col - j0;
						if (lc >= 0 && lc < W)
							rowbuf[lc] = 0.0f;
						}
					}
				__syncwarp();
				}
			if (strip == 0 && lane == 0)
				rowoff[LX] = written;
			if (lane == 31 && strip + 1 < nstrips)
				eOut[0] = 0.0f;
			if (strip == nstrips - 1)
				{
				const int q = LY - 1 - j0;          // column-local index of DP column LY
				eaScore = old[(q % C)*32 + q/C];
				}
			__syncwarp();
			float *tmp = eIn; eIn = eOut; eOut = tmp;
			}
		}
		if (lane == 0)
			P.ea[pair] = __fdiv_rn(eaScore, (float) min(LX, LY));     // calcposteriorflat.cpp:89
		__syncwarp();
		}
	}
