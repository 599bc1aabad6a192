// relax.cu -- the consistency transform (one Jacobi iteration) on the device-resident store.
//
// Replaces MPCFlat::ConsIter (consflat.cpp:5-23) / MPCFlat::ConsPair (conspairflat.cpp:10-110) with
// RelaxFlat_ZX_ZY, RelaxFlat_XZ_ZY, RelaxFlat_XZ_YZ (relaxflat.cpp:4-94) and
// MySparseMx::UpdateFromPost (mysparsemx.cpp:87-113).
//
// Formulation.  The reference scatters P_XZ[i,k]*P_ZY[k,j] into a dense LX*LY scratch and then
// keeps only XY's own pattern.  Here every stored entry (i,j) of XY is owned by one thread that
// GATHERS its own sum: with both orientations of every pair in HBM, row i of S[x->z] and row j of
// S[y->z] are two short column-sorted lists and the contribution of z is their sparse dot product
// (a merge).  For a fixed entry the reference adds the products in ascending z and, inside one z,
// in ascending k for all three of its loop nests; the merge visits k ascending, z runs ascending,
// every product is one __fmul_rn and every accumulation one __fadd_rn, so the result is
// bit-identical (verified against the oracle).  Relax is a sparse-sparse contraction (about 7
// non-zeros per row, 2 % density): tensor cores do not apply (DESIGN.md "relax is not a GEMM").
#include "engine.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return mb_fail(ctx, e_ == cudaErrorMemoryAllocation ? MB200_ENOMEM : MB200_ECUDA, \
	  "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define ENSURE(buf, bytes) do { if ((buf).ensure(bytes) != 0) \
	return mb_fail(ctx, MB200_ENOMEM, "device allocation of %zu bytes failed (%s)", (size_t)(bytes), #buf); } while (0)

struct RelaxParams
	{
	uint32_t n;                       // sequences
	uint32_t p_lo, p_hi;              // pairs to update
	const uint32_t *seqlen;
	const uint64_t *rowbase;  const uint32_t *rowoff;  const mb200_entry *entries;      // forward
	const uint64_t *trbase;   const uint32_t *troff;   const mb200_entry *trentries;    // transposed
	const uint64_t *entbase;
	mb200_entry *out;                 // new values (same layout as entries)
	};

struct ZDesc { const uint32_t *roA; const mb200_entry *enA; const uint32_t *roB; const mb200_entry *enB; };

#define RELAX_THREADS 256
#define RELAX_ZCHUNK 64
#define RELAX_EPT 4          // entries per thread held in registers per sweep
#ifndef RELAX_BAND_GUESS
#define RELAX_BAND_GUESS 0      // measured on C2: merge 285 ms, band-guess lookup 440 ms per iteration
#endif

__device__ __forceinline__ uint32_t pair_index(uint32_t n, uint32_t a, uint32_t b)
	{
	// row-major i<j (mpcflat.cpp:139-159): p = a*n - a(a+1)/2 + (b-a-1)
	return (uint32_t)((uint64_t) a*n - (uint64_t) a*(a + 1)/2 + (b - a - 1));
	}

__global__ void __launch_bounds__(RELAX_THREADS)
k_relax(const RelaxParams P)
	{
	__shared__ ZDesc zd[RELAX_ZCHUNK];
	const uint32_t p = P.p_lo + blockIdx.x;
	if (p >= P.p_hi)
		return;
	// invert p -> (x,y): find x with base(x) <= p < base(x+1)
	const uint32_t n = P.n;
	uint32_t x = 0;
		{
		// base(x) = x*n - x(x+1)/2 ; solve by a short search from the analytic estimate
		const double nn = (double) n - 0.5;
		double est = nn - sqrt(fmax(0.0, nn*nn - 2.0*(double) p));
		x = (uint32_t) fmin(fmax(est, 0.0), (double)(n - 2));
		while (x > 0 && (uint64_t) x*n - (uint64_t) x*(x + 1)/2 > p)
			--x;
		while (x + 1 < n - 1 && (uint64_t)(x + 1)*n - (uint64_t)(x + 1)*(x + 2)/2 <= p)
			++x;
		}
	const uint32_t y = x + 1 + (p - (uint32_t)((uint64_t) x*n - (uint64_t) x*(x + 1)/2));
	const uint32_t LX = P.seqlen[x];
	const uint32_t *ro = P.rowoff + P.rowbase[p];
	const uint64_t eb = P.entbase[p];
	const mb200_entry *en = P.entries + eb;
	mb200_entry *out = P.out + eb;
	const uint32_t nnz = ro[LX];
	const float fn = (float) n;

	for (uint32_t e0 = 0; e0 < nnz; e0 += RELAX_THREADS*RELAX_EPT)
		{
		// this thread's entries of the sweep: (row i, col j, accumulator)
		uint32_t ei[RELAX_EPT], ej[RELAX_EPT];
		float acc[RELAX_EPT];
		bool live[RELAX_EPT];
#pragma unroll
		for (int q = 0; q < RELAX_EPT; ++q)
			{
			const uint32_t e = e0 + q*RELAX_THREADS + threadIdx.x;
			live[q] = e < nnz;
			ei[q] = 0; ej[q] = 0; acc[q] = 0.0f;
			if (live[q])
				{
				const mb200_entry v = en[e];
				ej[q] = v.col;
				acc[q] = __fmul_rn(v.p, 2.0f);          // Z=X and Z=Y (conspairflat.cpp:26-30)
				// row of entry e: largest i with ro[i] <= e (binary search over LX+1 offsets)
				uint32_t lo = 0, hi = LX;
				while (hi - lo > 1)
					{
					const uint32_t mid = (lo + hi) >> 1;
					if (ro[mid] <= e) lo = mid; else hi = mid;
					}
				ei[q] = lo;
				}
			}
		for (uint32_t z0 = 0; z0 < n; z0 += RELAX_ZCHUNK)
			{
			__syncthreads();
			if (threadIdx.x < RELAX_ZCHUNK)
				{
				const uint32_t z = z0 + threadIdx.x;
				ZDesc d = { nullptr, nullptr, nullptr, nullptr };
				if (z < n && z != x && z != y)
					{
					// S[x->z]: rows are positions of x
					if (x < z)
						{
						const uint32_t q = pair_index(n, x, z);
						d.roA = P.rowoff + P.rowbase[q]; d.enA = P.entries + P.entbase[q];
						}
					else
						{
						const uint32_t q = pair_index(n, z, x);
						d.roA = P.troff + P.trbase[q]; d.enA = P.trentries + P.entbase[q];
						}
					// S[y->z]: rows are positions of y
					if (y < z)
						{
						const uint32_t q = pair_index(n, y, z);
						d.roB = P.rowoff + P.rowbase[q]; d.enB = P.entries + P.entbase[q];
						}
					else
						{
						const uint32_t q = pair_index(n, z, y);
						d.roB = P.troff + P.trbase[q]; d.enB = P.trentries + P.entbase[q];
						}
					}
				zd[threadIdx.x] = d;
				}
			__syncthreads();
			const uint32_t zn = min((uint32_t) RELAX_ZCHUNK, n - z0);
			for (uint32_t zz = 0; zz < zn; ++zz)
				{
				const ZDesc d = zd[zz];
				if (d.roA == nullptr)
					continue;
#pragma unroll
				for (int q = 0; q < RELAX_EPT; ++q)
					{
					if (!live[q])
						continue;
					uint32_t a = d.roA[ei[q]];
					const uint32_t aend = d.roA[ei[q] + 1];
					const uint32_t b0 = d.roB[ej[q]];
					const uint32_t bend = d.roB[ej[q] + 1];
					if (a == aend || b0 == bend)
						continue;
					float s = acc[q];
#if RELAX_BAND_GUESS
					// posterior rows are (nearly) contiguous column bands: the entry with column k sits at
					// index k-firstcol unless the row has holes below k (then walk back a few slots)
					const uint32_t kb0 = d.enB[b0].col;
					for (; a < aend; ++a)
						{
						const mb200_entry ea = d.enA[a];
						if (ea.col < kb0)
							continue;
						uint32_t idx = min(b0 + (ea.col - kb0), bend - 1);
						mb200_entry eb2 = d.enB[idx];
						while (eb2.col > ea.col && idx > b0)
							eb2 = d.enB[--idx];
						if (eb2.col == ea.col)
							s = __fadd_rn(s, __fmul_rn(ea.p, eb2.p));      // relaxflat.cpp:27,56,90
						}
#else
					// branch-free merge step: both cursors advance by predicate, the product is added only on
					// a column match (lanes stay converged inside the loop; profiles: the 3-way if/else of the
					// textbook merge ran at 12.6/32 active lanes)
					uint32_t b = b0;
					mb200_entry ea = d.enA[a], ebv = d.enB[b];
					for (;;)
						{
						const bool adva = ea.col <= ebv.col;
						const bool advb = ebv.col <= ea.col;
						if (adva && advb)
							s = __fadd_rn(s, __fmul_rn(ea.p, ebv.p));      // relaxflat.cpp:27,56,90
						a += adva ? 1u : 0u;
						b += advb ? 1u : 0u;
						if (a >= aend || b >= bend)
							break;
						if (adva)
							ea = d.enA[a];
						if (advb)
							ebv = d.enB[b];
						}
#endif
					acc[q] = s;
					}
				}
			}
#pragma unroll
		for (int q = 0; q < RELAX_EPT; ++q)
			{
			if (live[q])
				{
				const uint32_t e = e0 + q*RELAX_THREADS + threadIdx.x;
				mb200_entry v;
				v.col = ej[q];
				v.p = __fdiv_rn(acc[q], fn);                    // mysparsemx.cpp:108
				out[e] = v;
				}
			}
		}
	}

// copy entries of pairs outside [p_lo,p_hi) unchanged into the new buffer
__global__ void k_copy_entries(uint64_t lo, uint64_t hi, const mb200_entry *__restrict__ src, mb200_entry *__restrict__ dst)
	{
	for (uint64_t k = lo + blockIdx.x*(uint64_t) blockDim.x + threadIdx.x; k < hi; k += (uint64_t) gridDim.x*blockDim.x)
		dst[k] = src[k];
	}

extern "C" {

int mb200_consistency_iter(mb200_ctx *ctx, uint32_t p_lo, uint32_t p_hi)
	{
	if (!ctx)
		return MB200_EINVAL;
	if (!ctx->store_valid || !ctx->store_allpairs)
		return mb_fail(ctx, MB200_EINVAL, "mb200_consistency_iter: the store must hold all N(N-1)/2 pairs "
		  "(mb200_posteriors_allpairs(0,npairs) or mb200_store_load_allpairs)");
	const uint32_t n = ctx->nseq;
	const uint32_t np = (uint32_t) ctx->h_px.size();
	if (p_lo > p_hi || p_hi > np)
		return mb_fail(ctx, MB200_EINVAL, "pair range [%u,%u) invalid", p_lo, p_hi);
	if (n < 3 || p_lo == p_hi)
		return MB200_OK;                       // MPCFlat::Consistency skips N<3 (mpcflat.cpp:176)
	cudaSetDevice(ctx->device);
	cudaStream_t st = ctx->stream;
	CU(cudaEventRecord(ctx->ev0, st));
	int rc = mb_store_build_transposed(ctx);
	if (rc != MB200_OK)
		return rc;
	if (ctx->tr_values_stale)
		{
		rc = mb_store_refresh_transposed(ctx);
		if (rc != MB200_OK)
			return rc;
		ctx->tr_values_stale = false;
		}
	ENSURE(ctx->d_entries2, (ctx->store_nnz + 64)*sizeof(mb200_entry));
	RelaxParams P;
	P.n = n; P.p_lo = p_lo; P.p_hi = p_hi;
	P.seqlen = (const uint32_t *) ctx->d_seqlen.p;
	P.rowbase = (const uint64_t *) ctx->d_rowbase.p;
	P.rowoff = (const uint32_t *) ctx->d_rowoff.p;
	P.entries = (const mb200_entry *) ctx->d_entries.p;
	P.trbase = (const uint64_t *) ctx->d_tr_rowbase.p;
	P.troff = (const uint32_t *) ctx->d_tr_rowoff.p;
	P.trentries = (const mb200_entry *) ctx->d_tr_entries.p;
	P.entbase = (const uint64_t *) ctx->d_entbase.p;
	P.out = (mb200_entry *) ctx->d_entries2.p;
	CU(cudaEventRecord(ctx->ev1, st));
	k_relax<<<p_hi - p_lo, RELAX_THREADS, 0, st>>>(P);
	CU(cudaGetLastError());
	ctx->stats.kernel_launches++;
	CU(cudaEventRecord(ctx->ev2, st));
	// pairs outside the range keep their old values until the peers' results arrive
	const uint64_t e_lo = ctx->h_entbase[p_lo];
	const uint64_t e_hi = p_hi < np ? ctx->h_entbase[p_hi] : ctx->store_nnz;
	const int blocks = ctx->prop.multiProcessorCount*8;
	if (e_lo > 0)
		{
		k_copy_entries<<<blocks, 256, 0, st>>>(0, e_lo, (const mb200_entry *) ctx->d_entries.p, (mb200_entry *) ctx->d_entries2.p);
		ctx->stats.kernel_launches++;
		}
	if (e_hi < ctx->store_nnz)
		{
		k_copy_entries<<<blocks, 256, 0, st>>>(e_hi, ctx->store_nnz, (const mb200_entry *) ctx->d_entries.p, (mb200_entry *) ctx->d_entries2.p);
		ctx->stats.kernel_launches++;
		}
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	std::swap(ctx->d_entries, ctx->d_entries2);            // consflat.cpp:22
	ctx->tr_values_stale = true;
	CU(cudaEventRecord(ctx->ev3, st));
	CU(cudaStreamSynchronize(st));
	cudaEventElapsedTime(&ctx->stats.last_kernel_ms, ctx->ev1, ctx->ev2);
	cudaEventElapsedTime(&ctx->stats.last_total_ms, ctx->ev0, ctx->ev3);
	return MB200_OK;
	}

} // extern "C"
