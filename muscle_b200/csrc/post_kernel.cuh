// post_kernel.cuh -- fused Forward -> total -> Backward+posterior -> sparsify -> EA kernel.
//
// Replaces, per sequence pair, CalcFwdFlat (fwdflat3.cpp:12-153), CalcTotalProbFlat
// (totalprobflat.cpp:3-16), CalcBwdFlat (bwdflat3.cpp:10-184), CalcPostFlat
// (calcposteriorflat.cpp:4-27), MySparseMx::FromPost (mysparsemx.cpp:115-152) and CalcAlnScoreFlat
// (calcalnscoreflat.cpp:4-32) of the reference.
//
// Mapping.  One warp owns one pair at a time (persistent warps pull pairs, longest first, from an
// atomic cursor).  Lane l owns C consecutive DP columns of Y; a warp therefore covers a strip of
// 32*C columns and longer Y are processed strip after strip with the strip-edge column handed over
// through a small per-warp buffer.  Rows of X are swept in anti-diagonal (wavefront) order: at
// step t lane l works on row t-l, the cell to the left lives in lane l-1 and was produced one step
// earlier, so the recurrence dependency is three warp shuffles per step.  All five HMM states of
// the wavefront live in registers; only the Forward M-state is spilled (4 B/cell), in
// [step][lane][C] order so that both the Forward write and the Backward read are fully coalesced.
// Backward walks the same anti-diagonals in reverse, fuses the Fwd (*) Bwd posterior, thresholds
// at log(0.01) and appends survivors to a per-row candidate list; the warp then compacts the rows
// into MySparseMx order and runs the expected-accuracy max-sum DP row by row as a warp-wide
// prefix-max (new[j] = max_{k<=j} max(old[k], old[k-1]+P[k]) is exactly the reference's
// max3 recurrence because every value is an exact max of the same sums).
//
// Border handling without special code paths: LOG_ZERO is absorbing under the branch-free LogAdd,
// so the reference's border formulas (first row/column of Forward, last row/column of Backward)
// fall out of the general cell update when the out-of-range neighbours are LOG_ZERO.  The only
// injected values are the start scores at Forward (0,0) and the end scores at Backward (LX,LY).
#pragma once
#include "common.cuh"


template <int C>
struct PostSmem
	{
	float4 coef[4];
	float  insT[MB_MAX_K];
	float  rowbuf[MB_WARPS_PER_BLOCK][32*C];
	// matchT follows (dynamic): K*KS floats
	};

// ---------------------------------------------------------------------------------------------
// Forward step for one lane: C cells of row i.  CAPTURE additionally returns the five states of
// the cell in local column `clast` (the (LX,LY) corner needed by the total probability).
template <int C, bool CAPTURE>
__device__ __forceinline__ void fwd_cells(const MbHmm &h, const LogAdd &la, const float *mrow,
  const int (&yc)[C], const float (&ey)[C], float ex,
  float (&S)[C], float (&M)[C], float (&IX)[C], float (&JX)[C],
  float diag, float &lm, float &laiy, float &lajy, int clast, float (&fin)[5])
	{
#pragma unroll
	for (int c = 0; c < C; ++c)
		{
		// fwdflat3.cpp:113-141
		const float m = ADD(diag, mrow[yc[c]]);
		const float ix = ADD(la(ADD(IX[c], h.tII), ADD(M[c], h.tMI)), ex);
		const float jx = ADD(la(ADD(JX[c], h.tJJ), ADD(M[c], h.tMJ)), ex);
		const float iy = ADD(la(laiy, ADD(lm, h.tMI)), ey[c]);
		const float jy = ADD(la(lajy, ADD(lm, h.tMJ)), ey[c]);
		// what cell (i+1,j+1) will need: LOG_ADD(M+tMM, IX+tIM, JX+tJM, IY+tIM, JY+tJM), right fold
		const float s = la(ADD(m, h.tMM), la(ADD(ix, h.tIM), la(ADD(jx, h.tJM),
		  la(ADD(iy, h.tIM), ADD(jy, h.tJM)))));
		diag = S[c];
		S[c] = s; M[c] = m; IX[c] = ix; JX[c] = jx;
		lm = m;
		laiy = ADD(iy, h.tII);
		lajy = ADD(jy, h.tJJ);
		if (CAPTURE && c == clast)
			{
			fin[0] = m; fin[1] = ix; fin[2] = iy; fin[3] = jx; fin[4] = jy;   // state order M,IX,IY,JX,JY
			}
		}
	}

// Backward step for one lane: C cells of row i, right to left.  INJECT forces the end-of-alignment
// scores into local column `clast` (bwdflat3.cpp:53-61).
template <int C, bool INJECT>
__device__ __forceinline__ void bwd_cells(const MbHmm &h, const LogAdd &la, const float *mrow,
  const int (&yb)[C], const float (&eyb)[C], float ex,
  float (&M)[C], float (&IX)[C], float (&JX)[C],
  float mdiag, float &riy, float &rjy, int clast)
	{
#pragma unroll
	for (int c = C - 1; c >= 0; --c)
		{
		// bwdflat3.cpp:75-121
		const float nM = ADD(mdiag, mrow[yb[c]]);
		const float nIX = ADD(IX[c], ex);
		const float nJX = ADD(JX[c], ex);
		const float nIY = ADD(riy, eyb[c]);
		const float nJY = ADD(rjy, eyb[c]);
		float m = la(ADD(h.tMM, nM), la(ADD(h.tMI, nIX), la(ADD(h.tMJ, nJX),
		  la(ADD(h.tMI, nIY), ADD(h.tMJ, nJY)))));
		float ix = la(ADD(h.tII, nIX), ADD(h.tIM, nM));
		float jx = la(ADD(h.tJJ, nJX), ADD(h.tJM, nM));
		float iy = la(ADD(h.tII, nIY), ADD(h.tIM, nM));
		float jy = la(ADD(h.tJJ, nJY), ADD(h.tJM, nM));
		if (INJECT && c == clast)
			{
			m = h.tSM; ix = h.tSI; jx = h.tSJ; iy = h.tSI; jy = h.tSJ;
			}
		mdiag = M[c];
		M[c] = m; IX[c] = ix; JX[c] = jx;
		riy = iy; rjy = jy;
		}
	}

// resident CTAs per SM the register budget is tuned for (4 warps each)
#ifndef MB_OCC_EXPR
#define MB_OCC_EXPR (C <= 8 ? 5 : (C <= 12 ? 4 : 3))     // measured on C3: +5 % over 3 CTAs/SM, spills <= 120 B
#endif
template <int C> struct PostOcc { static constexpr int kBlocks = MB_OCC_EXPR; };

template <int C>
__global__ void __launch_bounds__(32*MB_WARPS_PER_BLOCK, PostOcc<C>::kBlocks)
k_posterior(const PostParams P)
	{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	PostSmem<C> &sm = *reinterpret_cast<PostSmem<C> *>(smem_raw);
	float *matchT = reinterpret_cast<float *>(smem_raw + sizeof(PostSmem<C>));

	const MbHmm h = P.h;
	const int lane = threadIdx.x & 31;
	const int wib = threadIdx.x >> 5;
	const int gwarp = blockIdx.x*MB_WARPS_PER_BLOCK + wib;

	if (threadIdx.x < 4)
		sm.coef[threadIdx.x] = c_logexp1[threadIdx.x];
	for (int k = threadIdx.x; k < h.K; k += blockDim.x)
		sm.insT[k] = P.insT[k];
	for (int k = threadIdx.x; k < h.K*h.KS; k += blockDim.x)
		matchT[k] = P.matchT[k];
	for (int k = lane; k < 32*C; k += 32)
		sm.rowbuf[wib][k] = 0.0f;
	__syncthreads();

	const LogAdd la = { sm.coef };
	const float Z = MB_LOG_ZERO;
	constexpr int W = 32*C;

	float       *fm     = P.fm + (size_t) gwarp*P.fm_stride;
	float4      *edge0  = P.edge + (size_t) gwarp*P.edge_stride;
	float4      *edge1  = edge0 + (P.lxmax + 2);
	mb200_entry *rows   = P.rows + (size_t) gwarp*P.rows_stride;
	uint8_t     *rowcnt = P.rowcnt + (size_t) gwarp*P.rowcnt_stride;
	float       *rowbuf = sm.rowbuf[wib];

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
		const int nstrips = (LY + W - 1)/W;
		const int lastj0 = (nstrips - 1)*W;
		const int lcl = (LY - 1 - lastj0)/C;           // lane and local column of DP column LY
		const int clast = (LY - 1 - lastj0) - lcl*C;

		for (int k = lane; k < LX; k += 32)
			rowcnt[k] = 0;

		// ============================ Forward ============================
		float fin[5] = { Z, Z, Z, Z, Z };
		for (int strip = 0; strip < nstrips; ++strip)
			{
			const int j0 = strip*W;
			const int ncol = min(W, LY - j0);
			const int nl = (ncol + C - 1)/C;
			const bool last = (strip == nstrips - 1);
			const float4 *edgeIn = (strip & 1) ? edge0 : edge1;
			float4 *edgeOut = (strip & 1) ? edge1 : edge0;
			float *fms = fm + (size_t) strip*P.fm_rows*W;

			int yc[C]; float ey[C];
#pragma unroll
			for (int c = 0; c < C; ++c)
				{
				const int jj = j0 + lane*C + c;
				const int code = jj < LY ? (int) Yc[jj] : h.pad;
				yc[c] = code;
				ey[c] = sm.insT[code];
				}
			float S[C], M[C], IX[C], JX[C];
#pragma unroll
			for (int c = 0; c < C; ++c)
				{
				S[c] = Z; M[c] = Z; IX[c] = Z; JX[c] = Z;
				}
			float outM = Z, outAIY = Z, outAJY = Z, outS = Z;
			float dprev = Z;
			float bIX = Z, bJX = Z;
			const int nsteps = LX + nl;
			int xcPref = h.pad;                       // residue class of the row this lane handles next
			for (int t = 0; t < nsteps; ++t)
				{
				const int i = t - lane;
				const int xc = xcPref;
				xcPref = (i >= 0 && i < LX) ? (int) Xc[i] : h.pad;    // row i+1 uses X[i]; consumed next step
				float Lm = __shfl_up_sync(MB_FULL, outM, 1);
				float Laiy = __shfl_up_sync(MB_FULL, outAIY, 1);
				float Lajy = __shfl_up_sync(MB_FULL, outAJY, 1);
				float Ls = __shfl_up_sync(MB_FULL, outS, 1);
				if (i >= 0 && i <= LX && lane < nl)
					{
					const float ex = sm.insT[xc];
					if (lane == 0)
						{
						if (strip == 0)
							{
							// column j=0 (fwdflat3.cpp:41-45,67-79): start scores, then X-insert chains
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
					const float diag = dprev;
					dprev = Ls;
					float lm = Lm, laiy = Laiy, lajy = Lajy;
					const float *mrow = matchT + xc*h.KS;
					if (last && i == LX && lane == lcl)
						fwd_cells<C, true>(h, la, mrow, yc, ey, ex, S, M, IX, JX, diag, lm, laiy, lajy, clast, fin);
					else
						fwd_cells<C, false>(h, la, mrow, yc, ey, ex, S, M, IX, JX, diag, lm, laiy, lajy, clast, fin);
					outM = lm; outAIY = laiy; outAJY = lajy; outS = S[C - 1];
					float *dst = fms + ((size_t) t*32 + lane)*C;
#pragma unroll
					for (int c = 0; c < C; ++c)
						dst[c] = M[c];
					if (lane == 31 && !last)
						edgeOut[i] = make_float4(outM, outAIY, outAJY, outS);
					if (P.dbg_fwd != nullptr && i >= 1)
						{
#pragma unroll
						for (int c = 0; c < C; ++c)
							{
							const int col = j0 + lane*C + c;
							if (col < LY)
								P.dbg_fwd[(size_t)(i - 1)*LY + col] = M[c];
							}
						}
					}
				}
			__syncwarp();
			}

		// ============================ total probability ============================
		// totalprobflat.cpp:3-16: LOG_PLUS_EQUALS over states M,IX,IY,JX,JY of Fwd+Bwd at (LX,LY);
		// Bwd(LX,LY) are the start scores (bwdflat3.cpp:53-61).
#pragma unroll
		for (int s = 0; s < 5; ++s)
			fin[s] = __shfl_sync(MB_FULL, fin[s], lcl);
		float total = Z;
		total = la(total, ADD(fin[0], h.tSM));
		total = la(total, ADD(fin[1], h.tSI));
		total = la(total, ADD(fin[2], h.tSI));
		total = la(total, ADD(fin[3], h.tSJ));
		total = la(total, ADD(fin[4], h.tSJ));
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

			int yb[C]; float eyb[C];
#pragma unroll
			for (int c = 0; c < C; ++c)
				{
				const int jj = j0 + lane*C + c + 1;       // Y[j] for DP column j (0-based residue j)
				const int code = jj < LY ? (int) Yc[jj] : h.pad;
				yb[c] = code;
				eyb[c] = sm.insT[code];
				}
			float M[C], IX[C], JX[C];
#pragma unroll
			for (int c = 0; c < C; ++c)
				{
				M[c] = Z; IX[c] = Z; JX[c] = Z;
				}
			float outM = Z, outIY = Z, outJY = Z;
			float dprev = Z;
			const int nsteps = LX + nl - 1;
			int xcPref = h.pad;
			for (int u = 0; u < nsteps; ++u)
				{
				const int i = LX - u + (nl - 1 - lane);
				const int xc = xcPref;
				xcPref = (i >= 2 && i <= LX) ? (int) Xc[i - 1] : h.pad;   // row i-1 uses X[i-1]
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
					const float ex = sm.insT[xc];
					const float *mrow = matchT + xc*h.KS;
					uint32_t cnt = rowcnt[i - 1];             // issued early, consumed after the cell updates
					const int t = i + lane;
					const float *src = fms + ((size_t) t*32 + lane)*C;
					float fmv[C];
#pragma unroll
					for (int c = 0; c < C; ++c)
						fmv[c] = src[c];
					const float mdiag = dprev;
					dprev = Rm;
					float riy = Riy, rjy = Rjy;
					if (last && i == LX && lane == lcl)
						bwd_cells<C, true>(h, la, mrow, yb, eyb, ex, M, IX, JX, mdiag, riy, rjy, clast);
					else
						bwd_cells<C, false>(h, la, mrow, yb, eyb, ex, M, IX, JX, mdiag, riy, rjy, clast);
					outM = M[0]; outIY = riy; outJY = rjy;
					if (lane == 0 && strip > 0)
						edgeOut[i] = make_float4(outM, outIY, outJY, 0.0f);

					// posterior (calcposteriorflat.cpp:14-22): candidates arrive with descending column
					const uint32_t cnt0 = cnt;
					mb200_entry *row = rows + (size_t)(i - 1)*MB_CAP;
#pragma unroll
					for (int c = C - 1; c >= 0; --c)
						{
						const int col = j0 + lane*C + c;
						const float score = __fsub_rn(ADD(fmv[c], M[c]), total);
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
						}
					if (cnt != cnt0)
						rowcnt[i - 1] = (uint8_t) min(cnt, (uint32_t) MB_CAP);
					if (P.dbg_bwd != nullptr)
						{
#pragma unroll
						for (int c = 0; c < C; ++c)
							{
							const int col = j0 + lane*C + c;
							if (col < LY)
								{
								P.dbg_bwd[(size_t)(i - 1)*LY + col] = M[c];
								const float score = __fsub_rn(ADD(fmv[c], M[c]), total);
								P.dbg_post[(size_t)(i - 1)*LY + col] =
								  score < h.minScore ? 0.0f : (score >= 0.0f ? 1.0f : mb_expf_glibc(score));
								}
							}
						}
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
		float *eIn = reinterpret_cast<float *>(edge0);       // edge buffers are free again
		float *eOut = reinterpret_cast<float *>(edge1);
		uint32_t written = 0;
		for (int strip = 0; strip < nstrips; ++strip)
			{
			const int j0 = strip*W;
			float old[C];
#pragma unroll
			for (int c = 0; c < C; ++c)
				old[c] = 0.0f;
			for (int i = 1; i <= LX; ++i)
				{
				const uint32_t cnt = rowcnt[i - 1];
				const mb200_entry *row = rows + (size_t)(i - 1)*MB_CAP;
				if (strip == 0 && lane == 0)
					rowoff[i - 1] = written;
				// ascending-column sweep over the candidates of this row
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
				float leftOld = __shfl_up_sync(MB_FULL, old[C - 1], 1);
				float incoming = 0.0f;
				if (lane == 0)
					{
					leftOld = strip == 0 ? 0.0f : eIn[i - 1];
					incoming = strip == 0 ? 0.0f : eIn[i];
					}
				float v[C];
#pragma unroll
				for (int c = 0; c < C; ++c)
					{
					const float pl = rowbuf[lane*C + c];
					const float b = ADD(c == 0 ? leftOld : old[c > 0 ? c - 1 : 0], pl);
					v[c] = fmaxf(old[c], b);
					if (c > 0)
						v[c] = fmaxf(v[c], v[c > 0 ? c - 1 : 0]);
					}
				float run = v[C - 1];
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
					{
					const float n = __shfl_up_sync(MB_FULL, run, o);
					if (lane >= o)
						run = fmaxf(run, n);
					}
				float excl = __shfl_up_sync(MB_FULL, run, 1);
				const float inc0 = __shfl_sync(MB_FULL, incoming, 0);
				excl = lane == 0 ? inc0 : fmaxf(excl, inc0);
#pragma unroll
				for (int c = 0; c < C; ++c)
					old[c] = fmaxf(v[c], excl);
				if (lane == 31 && strip + 1 < nstrips)
					eOut[i] = old[C - 1];
				__syncwarp();
				// clear the staged row
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
#pragma unroll
				for (int c = 0; c < C; ++c)
					rowbuf[lane*C + c] = old[c];
				__syncwarp();
				eaScore = rowbuf[LY - 1 - j0];
				__syncwarp();
#pragma unroll
				for (int c = 0; c < C; ++c)
					rowbuf[lane*C + c] = 0.0f;
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
