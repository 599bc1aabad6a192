// align.cu -- posterior decoding on the device: max-sum DP with traceback (CalcAlnFlat) for batches
// of stored pairs and for the column-posterior matrix of a progressive-alignment join (BuildPost).
//
// Replaces CalcAlnFlat (calcalnflat.cpp:6-46) + Best3 (best3.h:5-28) + TraceBackFlat
// (tracebackflat.cpp:3-37), MPCFlat::BuildPost (buildpostflat.cpp:18-105) and the arithmetic of
// MPCFlat::AlignAlns (alnalnsflat.cpp:7-52).
//
// DP.  One CTA per problem, rows in sequence.  A row is  new[j] = max(old[j-1]+P[j], old[j], new[j-1])
// which equals the running maximum over k<=j of max(old[k], old[k-1]+P[k]); the values are exact
// maxima of the same fp32 sums the reference forms, so a block-wide prefix-max gives bit-identical
// rows, and with old[j-1]+P, old[j] and new[j-1] known the traceback letter of every cell follows
// from Best3's tie rule (B if B>=X and B>=Y; else Y if B>=X; else X if X>=Y else Y) independently.
#include "engine.h"
#include <algorithm>
#include <cstring>

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return mb_fail(ctx, e_ == cudaErrorMemoryAllocation ? MB200_ENOMEM : MB200_ECUDA, \
	  "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define ENSURE(buf, bytes) do { if ((buf).ensure(bytes) != 0) \
	return mb_fail(ctx, MB200_ENOMEM, "device allocation of %zu bytes failed (%s)", (size_t)(bytes), #buf); } while (0)

#define ALN_THREADS 256

// MB200_TRACE=1: wall-time split of mb200_align_groups, printed at process exit
#include <chrono>
#include <cstdlib>
#include <cstdio>
static bool g_trace = false;
static double g_t[6] = { 0, 0, 0, 0, 0, 0 };
static unsigned g_calls = 0;
static double now_s()
	{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}
static void trace_report()
	{
	fprintf(stderr, "[mb200 trace] align_groups x%u: host prep %.2f s, upload+memset %.2f s, gather %.2f s, apply %.2f s, "
	  "decode DP %.2f s, download %.2f s\n", g_calls, g_t[0], g_t[1], g_t[2], g_t[3], g_t[4], g_t[5]);
	}
#define TRACE_MARK(k) do { if (g_trace) { cudaStreamSynchronize(st); const double n_ = now_s(); g_t[k] += n_ - tmark; tmark = n_; } } while (0)

struct AlnProblem
	{
	uint32_t LX, LY;
	const float *dense;              // LX*LY row-major, or nullptr for sparse rows
	const uint32_t *rowoff;          // sparse: CSR of the pair
	const mb200_entry *entries;
	char *tb;                        // (LX+1)*(LY+1) scratch
	char *path;                      // LX+LY+1 output
	float *score;
	};

#define ALN_VPT 4            // columns per thread and pass: one pass covers 1024 columns

__global__ void __launch_bounds__(ALN_THREADS)
k_alnflat(const AlnProblem *probs)
	{
	extern __shared__ float sh[];        // old[LY+1], prow[LY+1]
	__shared__ float warpmax[ALN_THREADS/32];
	__shared__ float carry_s;
	const AlnProblem pr = probs[blockIdx.x];
	const uint32_t LX = pr.LX, LY = pr.LY;
	const uint32_t LY1 = LY + 1;
	float *old = sh;
	float *prow = sh + LY1;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

	for (uint32_t j = tid; j <= LY; j += ALN_THREADS)
		{
		old[j] = 0.0f;
		prow[j] = 0.0f;
		pr.tb[j] = 'Y';                                   // calcalnflat.cpp:15-19
		}
	__syncthreads();
	for (uint32_t i = 1; i <= LX; ++i)
		{
		// stage P[i-1][*] into prow[1..LY]
		if (pr.dense != nullptr)
			{
			const float *src = pr.dense + (size_t)(i - 1)*LY;
			for (uint32_t j = tid; j < LY; j += ALN_THREADS)
				prow[j + 1] = src[j];
			}
		else
			{
			const uint32_t b = pr.rowoff[i - 1], e = pr.rowoff[i];
			for (uint32_t k = b + tid; k < e; k += ALN_THREADS)
				prow[pr.entries[k].col + 1] = pr.entries[k].p;
			}
		__syncthreads();
		char *tbrow = pr.tb + (size_t) i*LY1;
		if (tid == 0)
			tbrow[0] = 'X';                                   // calcalnflat.cpp:25
		float carry = 0.0f;                                   // new[0] = 0
		for (uint32_t j0 = 1; j0 <= LY; j0 += ALN_THREADS*ALN_VPT)
			{
			const uint32_t jb = j0 + tid*ALN_VPT;             // this thread's first column
			float B[ALN_VPT], X[ALN_VPT], r[ALN_VPT];
			float left = jb <= LY ? old[jb - 1] : 0.0f;
#pragma unroll
			for (int q = 0; q < ALN_VPT; ++q)
				{
				const uint32_t j = jb + q;
				const bool in = j <= LY;
				X[q] = in ? old[j] : 0.0f;
				B[q] = in ? __fadd_rn(left, prow[j]) : 0.0f;
				left = X[q];
				const float v = fmaxf(B[q], X[q]);
				r[q] = q == 0 ? v : fmaxf(r[q > 0 ? q - 1 : 0], v);
				}
			// inclusive prefix max of the per-thread maxima across the block
			float run = r[ALN_VPT - 1];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
				{
				const float t = __shfl_up_sync(MB_FULL, run, o);
				if (lane >= (uint32_t) o)
					run = fmaxf(run, t);
				}
			if (lane == 31)
				warpmax[wid] = run;
			__syncthreads();
			float pre = carry;                                // max of everything left of this pass
			for (uint32_t w = 0; w < wid; ++w)
				pre = fmaxf(pre, warpmax[w]);
			float excl = __shfl_up_sync(MB_FULL, run, 1);     // new[jb-1] inside the warp
			excl = lane == 0 ? pre : fmaxf(excl, pre);
			float Y = excl;
#pragma unroll
			for (int q = 0; q < ALN_VPT; ++q)
				{
				const uint32_t j = jb + q;
				const float nw = fmaxf(r[q], excl);               // new[j]
				if (j <= LY)
					{
					char t;
					if (B[q] >= X[q])
						t = (B[q] >= Y) ? 'B' : 'Y';                 // best3.h:5-28 tie order
					else
						t = (X[q] >= Y) ? 'X' : 'Y';
					tbrow[j] = t;
					// old[] must stay the previous row until every thread has read it: park in prow[]
					prow[j] = nw;
					}
				Y = nw;
				}
			if (j0 + ALN_THREADS*ALN_VPT <= LY)
				{
				// another pass follows: hand the running maximum over
				if (tid == ALN_THREADS - 1)
					carry_s = Y;
				__syncthreads();
				carry = carry_s;
				}
			}
		__syncthreads();
		for (uint32_t j = tid + 1; j <= LY; j += ALN_THREADS)
			{
			old[j] = prow[j];
			prow[j] = 0.0f;
			}
		__syncthreads();
		}
	if (tid == 0)
		{
		*pr.score = old[LY];
		// tracebackflat.cpp:3-37
		uint32_t n = 0;
		int64_t i = LX, j = LY;
		while (i != 0 || j != 0)
			{
			const char t = pr.tb[(size_t) i*LY1 + (size_t) j];
			pr.path[n++] = t;
			if (t == 'B') { --i; --j; }
			else if (t == 'X') --i;
			else --j;
			}
		for (uint32_t a = 0, b = n; a + 1 < b; ++a, --b)
			{
			const char t = pr.path[a]; pr.path[a] = pr.path[b - 1]; pr.path[b - 1] = t;
			}
		pr.path[n] = 0;
		}
	}

// ---------------------------------------------------------------------------------------------
// BuildPost: one 8-lane group owns one column (row of Post) of alignment A and walks (s,t) in the
// reference's s-major, t-minor order; a cell receives at most one term per (s,t), so the owner's
// sequential adds reproduce the reference's fp32 sums exactly.
struct BuildPostParams
	{
	uint32_t n;                                   // sequences in the store
	uint32_t na, nb, cols_a, cols_b;
	const uint32_t *ids_a, *ids_b;
	const int32_t *col2pos_a;                     // [na][cols_a], -1 = gap
	const uint32_t *p2c_b;                        // concatenated pos->col maps of B
	const uint64_t *p2c_b_off;                    // [nb]
	const uint64_t *rowbase;  const uint32_t *rowoff;  const mb200_entry *entries;
	const uint64_t *trbase;   const uint32_t *troff;   const mb200_entry *trentries;
	const uint64_t *entbase;
	float *post;                                  // cols_a*cols_b, zeroed
	};

__global__ void __launch_bounds__(128)
k_buildpost(const BuildPostParams P)
	{
	const uint32_t grp = (blockIdx.x*blockDim.x + threadIdx.x) >> 3;      // Post row = column of A
	const uint32_t gl = threadIdx.x & 7;
	if (grp >= P.cols_a)
		return;
	float *prow = P.post + (size_t) grp*P.cols_b;
	const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7u);
	for (uint32_t s = 0; s < P.na; ++s)
		{
		const int32_t pos = P.col2pos_a[(size_t) s*P.cols_a + grp];
		if (pos < 0)
			continue;
		const uint32_t a = P.ids_a[s];
		for (uint32_t t = 0; t < P.nb; ++t)
			{
			const uint32_t b = P.ids_b[t];
			const uint32_t *ro;
			const mb200_entry *en;
			if (a < b)
				{
				const uint32_t q = (uint32_t)((uint64_t) a*P.n - (uint64_t) a*(a + 1)/2 + (b - a - 1));
				ro = P.rowoff + P.rowbase[q]; en = P.entries + P.entbase[q];
				}
			else
				{
				const uint32_t q = (uint32_t)((uint64_t) b*P.n - (uint64_t) b*(b + 1)/2 + (a - b - 1));
				ro = P.troff + P.trbase[q]; en = P.trentries + P.entbase[q];
				}
			const uint32_t e0 = ro[pos], e1 = ro[pos + 1];
			const uint32_t *p2c = P.p2c_b + P.p2c_b_off[t];
			for (uint32_t e = e0 + gl; e < e1; e += 8)
				{
				const mb200_entry v = en[e];
				const uint32_t c2 = p2c[v.col];
				prow[c2] = __fadd_rn(prow[c2], v.p);          // += w1*w2*P with unit weights
				}
			__syncwarp(gmask);
			}
		}
	}

#define BP_WARPS 4
#define BP_W 16           // staged entries per sparse row (rows are 7.2 +- 3 long: 25 % exceed 8, ~0.1 % exceed 16)

// ---------------------------------------------------------------------------------------------
// Two-phase BuildPost (default).  The direct formulation (k_buildpost: one lane group per row walking all (s,t)) is
// latency bound: every (s,t) step chains 4-5 dependent random loads into the 10 GB store and only
// cols_a groups exist (C3 trace: 44 ms per AlignAlns call, 49 s of a 110 s `muscle -align`).  Here the random gathers are done by a
// massively parallel pre-pass -- one THREAD per (row, s, t) copies the <= BP_W entries of the needed
// sparse row, already mapped to columns of B, into a dense staging slot -- and the order-sensitive
// accumulation then streams the staging area with coalesced loads.  The fp32 sum order of the
// reference (s-major, t-minor) is unchanged, so the result stays bit-identical.
struct BpStage
	{
	uint2   *slots;      // [row][s_local][t][BP_W]  (column of B, bits of P)
	uint8_t *cnt;        // [row][s_local][t]; 255 = row longer than BP_W (applied by direct gather)
	uint32_t s_lo, s_n;  // batch of sequences of A
	};

__global__ void __launch_bounds__(256)
k_bp_gather(const BuildPostParams P, const BpStage G)
	{
	const uint64_t total = (uint64_t) P.cols_a*G.s_n*P.nb;
	for (uint64_t idx = blockIdx.x*(uint64_t) blockDim.x + threadIdx.x; idx < total; idx += (uint64_t) gridDim.x*blockDim.x)
		{
		const uint32_t t = (uint32_t)(idx % P.nb);
		const uint64_t rs = idx/P.nb;
		const uint32_t sl = (uint32_t)(rs % G.s_n);
		const uint32_t row = (uint32_t)(rs/G.s_n);
		const uint32_t s = G.s_lo + sl;
		const int32_t pos = P.col2pos_a[(size_t) s*P.cols_a + row];
		uint32_t n = 0;
		if (pos >= 0)
			{
			const uint32_t a = P.ids_a[s], b = P.ids_b[t];
			const uint32_t *ro;
			const mb200_entry *en;
			if (a < b)
				{
				const uint32_t q = (uint32_t)((uint64_t) a*P.n - (uint64_t) a*(a + 1)/2 + (b - a - 1));
				ro = P.rowoff + P.rowbase[q]; en = P.entries + P.entbase[q];
				}
			else
				{
				const uint32_t q = (uint32_t)((uint64_t) b*P.n - (uint64_t) b*(b + 1)/2 + (a - b - 1));
				ro = P.troff + P.trbase[q]; en = P.trentries + P.entbase[q];
				}
			const uint32_t e0 = ro[pos];
			n = ro[pos + 1] - e0;
			if (n <= BP_W)
				{
				const uint32_t *p2c = P.p2c_b + P.p2c_b_off[t];
				uint2 *dst = G.slots + idx*BP_W;
				for (uint32_t k = 0; k < n; ++k)
					{
					const mb200_entry v = en[e0 + k];
					dst[k] = make_uint2(p2c[v.col], __float_as_uint(v.p));
					}
				}
			else
				n = 255;
			}
		G.cnt[idx] = (uint8_t) n;
		}
	}

__global__ void __launch_bounds__(32*BP_WARPS)
k_bp_apply(const BuildPostParams P, const BpStage G)
	{
	extern __shared__ __align__(16) unsigned char bp_smem[];
	const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t row = blockIdx.x*BP_WARPS + wib;
	float *acc = reinterpret_cast<float *>(bp_smem) + (size_t) wib*P.cols_b;
	// per-warp copy of the 32 staged rows of the current chunk (the ordered apply must not wait on
	// one global load per (s,t))
	uint2 *stage = reinterpret_cast<uint2 *>(bp_smem + (((size_t) BP_WARPS*P.cols_b*sizeof(float) + 15) & ~(size_t) 15))
	  + (size_t) wib*32*BP_W;
	if (row >= P.cols_a)
		return;
	float *prow = P.post + (size_t) row*P.cols_b;
	for (uint32_t c = lane; c < P.cols_b; c += 32)
		acc[c] = prow[c];
	__syncwarp();
	for (uint32_t sl = 0; sl < G.s_n; ++sl)
		{
		const uint32_t s = G.s_lo + sl;
		const int32_t pos = P.col2pos_a[(size_t) s*P.cols_a + row];
		if (pos < 0)
			continue;
		const uint64_t base = ((uint64_t) row*G.s_n + sl)*P.nb;
		for (uint32_t t0 = 0; t0 < P.nb; t0 += 32)
			{
			const uint32_t t = t0 + lane;
			const uint32_t n = t < P.nb ? (uint32_t) G.cnt[base + t] : 0u;
			const uint32_t any = __ballot_sync(MB_FULL, n != 0);
			if (any == 0)
				continue;
			if (n != 0 && n != 255)
				{
				const uint4 *src = reinterpret_cast<const uint4 *>(G.slots + (base + t)*BP_W);
				uint4 *dst = reinterpret_cast<uint4 *>(stage + lane*BP_W);
#pragma unroll
				for (int k = 0; k < (int)(BP_W*sizeof(uint2)/sizeof(uint4)); ++k)
					dst[k] = src[k];
				}
			__syncwarp();
			uint32_t rest = any;
			while (rest)
				{
				const uint32_t l = __ffs(rest) - 1;
				rest &= rest - 1;
				const uint32_t nl = __shfl_sync(MB_FULL, n, l);
				if (nl != 255)
					{
					if (lane < nl)
						{
						const uint2 v = stage[l*BP_W + lane];
						acc[v.x] = __fadd_rn(acc[v.x], __uint_as_float(v.y));     // += w1*w2*P, unit weights
						}
					}
				else if (lane == 0)
					{
					// rare: more than BP_W entries in the sparse row -> gather directly, still in order
					const uint32_t a = P.ids_a[s], b = P.ids_b[t0 + l];
					const uint32_t *ro;
					const mb200_entry *en;
					if (a < b)
						{
						const uint32_t q = (uint32_t)((uint64_t) a*P.n - (uint64_t) a*(a + 1)/2 + (b - a - 1));
						ro = P.rowoff + P.rowbase[q]; en = P.entries + P.entbase[q];
						}
					else
						{
						const uint32_t q = (uint32_t)((uint64_t) b*P.n - (uint64_t) b*(b + 1)/2 + (a - b - 1));
						ro = P.troff + P.trbase[q]; en = P.trentries + P.entbase[q];
						}
					const uint32_t *p2c = P.p2c_b + P.p2c_b_off[t0 + l];
					for (uint32_t e = ro[pos]; e < ro[pos + 1]; ++e)
						{
						const uint32_t c2 = p2c[en[e].col];
						acc[c2] = __fadd_rn(acc[c2], en[e].p);
						}
					}
				__syncwarp();
				}
			}
		}
	for (uint32_t c = lane; c < P.cols_b; c += 32)
		prow[c] = acc[c];
	}

extern "C" {

int mb200_align_pairs(mb200_ctx *ctx, uint32_t n, const uint32_t *store_pairs, char *paths_out,
  const uint64_t *path_off, float *scores_out)
	{
	if (!ctx || n == 0 || !store_pairs || !paths_out || !path_off || !scores_out)
		return mb_fail(ctx, MB200_EINVAL, "mb200_align_pairs: bad argument");
	int rc = mb_store_pack_inplace(ctx);
	if (rc != MB200_OK)
		return rc;
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	const uint32_t np = (uint32_t) ctx->h_px.size();
	std::vector<AlnProblem> probs(n);
	uint64_t tb_total = 0, path_total = path_off[n];
	uint32_t lymax = 0;
	for (uint32_t k = 0; k < n; ++k)
		{
		const uint32_t sp = store_pairs[k];
		if (sp >= np)
			return mb_fail(ctx, MB200_EINVAL, "store pair %u out of range", sp);
		const uint32_t LX = ctx->h_len[ctx->h_px[sp]], LY = ctx->h_len[ctx->h_py[sp]];
		tb_total += (uint64_t)(LX + 1)*(LY + 1);
		lymax = std::max(lymax, LY);
		}
	ENSURE(ctx->d_tmp, tb_total + 16);
	ENSURE(ctx->d_tmp2, path_total + n*sizeof(float) + n*sizeof(AlnProblem) + 64);
	char *d_paths = (char *) ctx->d_tmp2.p;
	float *d_scores = (float *)(d_paths + ((path_total + 15)/16)*16);
	AlnProblem *d_probs = (AlnProblem *)(d_scores + ((n + 3)/4)*4);
	uint64_t tboff = 0;
	for (uint32_t k = 0; k < n; ++k)
		{
		const uint32_t sp = store_pairs[k];
		AlnProblem &p = probs[k];
		p.LX = ctx->h_len[ctx->h_px[sp]];
		p.LY = ctx->h_len[ctx->h_py[sp]];
		p.dense = nullptr;
		p.rowoff = (const uint32_t *) ctx->d_rowoff.p + ctx->h_rowbase[sp];
		p.entries = (const mb200_entry *) ctx->d_entries.p + ctx->h_entbase[sp];
		p.tb = (char *) ctx->d_tmp.p + tboff;
		p.path = d_paths + path_off[k];
		p.score = d_scores + k;
		tboff += (uint64_t)(p.LX + 1)*(p.LY + 1);
		}
	CU(cudaMemcpyAsync(d_probs, probs.data(), n*sizeof(AlnProblem), cudaMemcpyHostToDevice, st));
	const size_t smem = 2*(size_t)(lymax + 1)*sizeof(float);
	if (smem > 200*1024)
		return mb_fail(ctx, MB200_EOVERFLOW, "sequence of length %u too long for the decoding kernel", lymax);
	CU(cudaFuncSetAttribute(k_alnflat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_alnflat<<<n, ALN_THREADS, smem, st>>>(d_probs);
	CU(cudaGetLastError());
	ctx->stats.kernel_launches++;
	CU(cudaMemcpyAsync(paths_out, d_paths, path_total, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(scores_out, d_scores, n*sizeof(float), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	ctx->stats.d2h_bytes += path_total + n*sizeof(float);
	return MB200_OK;
	}

int mb200_align_groups(mb200_ctx *ctx, uint32_t na, const uint32_t *ids_a, const uint32_t *pos2col_a, uint32_t cols_a,
  uint32_t nb, const uint32_t *ids_b, const uint32_t *pos2col_b, uint32_t cols_b,
  char *path_out, float *score_out, float *post_out)
	{
	if (!ctx || na == 0 || nb == 0 || !ids_a || !ids_b || !pos2col_a || !pos2col_b || !path_out || cols_a == 0 || cols_b == 0)
		return mb_fail(ctx, MB200_EINVAL, "mb200_align_groups: bad argument");
	if (!ctx->store_valid || !ctx->store_allpairs)
		return mb_fail(ctx, MB200_EINVAL, "mb200_align_groups: the store must hold all N(N-1)/2 pairs");
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	static bool trace_init = false;
	if (!trace_init)
		{
		trace_init = true;
		g_trace = getenv("MB200_TRACE") != nullptr;
		if (g_trace)
			atexit(trace_report);
		}
	int rc = mb_store_build_transposed(ctx);
	if (rc != MB200_OK)
		return rc;
	double tmark = now_s();
	++g_calls;
	if (ctx->tr_values_stale)
		{
		rc = mb_store_refresh_transposed(ctx);
		if (rc != MB200_OK)
			return rc;
		ctx->tr_values_stale = false;
		}
	// host-side index maps (tiny): col->pos for A, concatenated pos->col for B
	std::vector<int32_t> c2p((size_t) na*cols_a, -1);
	uint64_t off = 0;
	for (uint32_t s = 0; s < na; ++s)
		{
		if (ids_a[s] >= ctx->nseq)
			return mb_fail(ctx, MB200_EINVAL, "group A sequence id out of range");
		const uint32_t L = ctx->h_len[ids_a[s]];
		for (uint32_t i = 0; i < L; ++i)
			{
			const uint32_t c = pos2col_a[off + i];
			if (c >= cols_a)
				return mb_fail(ctx, MB200_EINVAL, "group A pos2col out of range");
			c2p[(size_t) s*cols_a + c] = (int32_t) i;
			}
		off += L;
		}
	std::vector<uint64_t> boff(nb);
	uint64_t btot = 0;
	for (uint32_t t = 0; t < nb; ++t)
		{
		if (ids_b[t] >= ctx->nseq)
			return mb_fail(ctx, MB200_EINVAL, "group B sequence id out of range");
		boff[t] = btot;
		btot += ctx->h_len[ids_b[t]];
		}
	const size_t post_bytes = (size_t) cols_a*cols_b*sizeof(float);
	const size_t tb_bytes = (size_t)(cols_a + 1)*(cols_b + 1);
	const size_t path_bytes = (size_t) cols_a + cols_b + 16;
	auto al16 = [](size_t b) { return (b + 15)/16*16; };
	const size_t need = al16(post_bytes) + al16(tb_bytes) + al16(path_bytes) + 16 + al16(sizeof(AlnProblem))
	  + al16(c2p.size()*4) + al16(btot*4) + al16(nb*8) + al16(na*4) + al16(nb*4) + 64;
	ENSURE(ctx->d_tmp, need);
	char *base = (char *) ctx->d_tmp.p;          // cudaMalloc base is 256-byte aligned; every slice 16-byte aligned
	float *d_post = (float *) base;                  base += al16(post_bytes);
	char *d_tb = base;                               base += al16(tb_bytes);
	char *d_path = base;                             base += al16(path_bytes);
	float *d_score = (float *) base;                 base += 16;
	AlnProblem *d_prob = (AlnProblem *) base;        base += al16(sizeof(AlnProblem));
	int32_t *d_c2p = (int32_t *) base;               base += al16(c2p.size()*4);
	uint32_t *d_p2cb = (uint32_t *) base;            base += al16(btot*4);
	uint64_t *d_boff = (uint64_t *) base;            base += al16(nb*8);
	uint32_t *d_ida = (uint32_t *) base;             base += al16(na*4);
	uint32_t *d_idb = (uint32_t *) base;
	TRACE_MARK(0);
	CU(cudaMemsetAsync(d_post, 0, post_bytes, st));
	CU(cudaMemcpyAsync(d_c2p, c2p.data(), c2p.size()*4, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_p2cb, pos2col_b, btot*4, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_boff, boff.data(), nb*8, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_ida, ids_a, na*4, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_idb, ids_b, nb*4, cudaMemcpyHostToDevice, st));
	ctx->stats.h2d_bytes += c2p.size()*4 + btot*4 + nb*8 + (na + nb)*4;
	BuildPostParams P;
	P.n = ctx->nseq; P.na = na; P.nb = nb; P.cols_a = cols_a; P.cols_b = cols_b;
	P.ids_a = d_ida; P.ids_b = d_idb; P.col2pos_a = d_c2p; P.p2c_b = d_p2cb; P.p2c_b_off = d_boff;
	P.rowbase = (const uint64_t *) ctx->d_rowbase.p; P.rowoff = (const uint32_t *) ctx->d_rowoff.p;
	P.entries = (const mb200_entry *) ctx->d_entries.p;
	P.trbase = (const uint64_t *) ctx->d_tr_rowbase.p; P.troff = (const uint32_t *) ctx->d_tr_rowoff.p;
	P.trentries = (const mb200_entry *) ctx->d_tr_entries.p;
	P.entbase = (const uint64_t *) ctx->d_entbase.p;
	P.post = d_post;
	TRACE_MARK(1);
	const size_t acc_smem = (((size_t) BP_WARPS*cols_b*sizeof(float) + 15) & ~(size_t) 15) + (size_t) BP_WARPS*32*BP_W*sizeof(uint2);
	if (acc_smem <= 160*1024)
		{
		// batches of sequences of A sized so that the staging area stays below ~512 MB
		const uint64_t per_s = (uint64_t) cols_a*nb*(BP_W*sizeof(uint2) + 1);
		uint32_t sb = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(na, (512ull << 20)/std::max<uint64_t>(per_s, 1)));
		ENSURE(ctx->d_tmp2, (uint64_t) cols_a*sb*nb*BP_W*sizeof(uint2) + (uint64_t) cols_a*sb*nb + 64);
		BpStage G;
		G.slots = (uint2 *) ctx->d_tmp2.p;
		G.cnt = (uint8_t *)(G.slots + (uint64_t) cols_a*sb*nb*BP_W);
		CU(cudaFuncSetAttribute(k_bp_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) acc_smem));
		for (uint32_t s_lo = 0; s_lo < na; s_lo += sb)
			{
			G.s_lo = s_lo;
			G.s_n = std::min(sb, na - s_lo);
			const uint64_t total = (uint64_t) cols_a*G.s_n*nb;
			const uint32_t gblocks = (uint32_t) std::min<uint64_t>((total + 255)/256, (uint64_t) ctx->prop.multiProcessorCount*32);
			k_bp_gather<<<gblocks, 256, 0, st>>>(P, G);
			TRACE_MARK(2);
			k_bp_apply<<<(cols_a + BP_WARPS - 1)/BP_WARPS, 32*BP_WARPS, acc_smem, st>>>(P, G);
			TRACE_MARK(3);
			ctx->stats.kernel_launches += 2;
			}
		}
	else
		{
		// very wide alignments: row accumulators stay in global memory
		const uint32_t groups_per_block = 128/8;
		k_buildpost<<<(cols_a + groups_per_block - 1)/groups_per_block, 128, 0, st>>>(P);
		}
	CU(cudaGetLastError());
	AlnProblem pr;
	pr.LX = cols_a; pr.LY = cols_b; pr.dense = d_post; pr.rowoff = nullptr; pr.entries = nullptr;
	pr.tb = d_tb; pr.path = d_path; pr.score = d_score;
	CU(cudaMemcpyAsync(d_prob, &pr, sizeof pr, cudaMemcpyHostToDevice, st));
	const size_t smem = 2*(size_t)(cols_b + 1)*sizeof(float);
	if (smem > 200*1024)
		return mb_fail(ctx, MB200_EOVERFLOW, "alignment with %u columns too wide for the decoding kernel", cols_b);
	CU(cudaFuncSetAttribute(k_alnflat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_alnflat<<<1, ALN_THREADS, smem, st>>>(d_prob);
	CU(cudaGetLastError());
	ctx->stats.kernel_launches += 2;
	TRACE_MARK(4);
	CU(cudaMemcpyAsync(path_out, d_path, cols_a + cols_b + 1, cudaMemcpyDeviceToHost, st));
	if (score_out)
		CU(cudaMemcpyAsync(score_out, d_score, sizeof(float), cudaMemcpyDeviceToHost, st));
	if (post_out)
		CU(cudaMemcpyAsync(post_out, d_post, post_bytes, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	TRACE_MARK(5);
	ctx->stats.d2h_bytes += cols_a + cols_b + 1 + 4 + (post_out ? post_bytes : 0);
	return MB200_OK;
	}

} // extern "C"
